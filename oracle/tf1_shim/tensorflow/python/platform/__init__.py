from . import gfile  # noqa: F401
