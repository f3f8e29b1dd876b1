import contextlib

from tensorflow import Tensor, convert_to_tensor, device, get_default_graph, name_scope  # noqa: F401


@contextlib.contextmanager
def op_scope(values, name, default_name=None):
    yield name or default_name


class GraphKeys(object):
    GLOBAL_VARIABLES = 'variables'
    TRAINABLE_VARIABLES = 'trainable_variables'
