from tensorflow import group, no_op  # noqa: F401
