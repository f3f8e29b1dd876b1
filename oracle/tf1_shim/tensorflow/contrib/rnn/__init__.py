"""tf.contrib.rnn subset (TF-1.0): static_rnn and core_rnn_cell.{LSTMCell, DropoutWrapper, MultiRNNCell,
LSTMStateTuple}.  LSTMCell: one fused weight [d_in + H, 4H] on concat([x, h]), zero bias, gate order
i, j, f, o, forget_bias 1.0, c' = sigmoid(f + 1) c + sigmoid(i) tanh(j), h' = sigmoid(o) tanh(c')."""
from . import core_rnn_cell  # noqa: F401
from .core_rnn_cell import LSTMCell, DropoutWrapper, MultiRNNCell, LSTMStateTuple  # noqa: F401
import tensorflow as _tf


def static_rnn(cell, inputs, initial_state=None, dtype=None, sequence_length=None, scope=None):
    assert sequence_length is None, 'not used by the reference'
    outputs = []
    with _tf.variable_scope(scope or 'rnn') as vs:
        state = initial_state
        if state is None:
            batch = inputs[0].get_shape().as_list()[0]
            state = cell.zero_state(batch, dtype)
        for t, x in enumerate(inputs):
            if t > 0:
                vs.reuse_variables()
            out, state = cell(x, state)
            outputs.append(out)
    return outputs, state
