from tensorflow import variable_scope, get_variable, get_variable_scope  # noqa: F401
