from . import variable_scope, init_ops, embedding_ops, array_ops, math_ops, nn_ops, control_flow_ops, rnn  # noqa: F401
