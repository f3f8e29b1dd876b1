import collections

import torch as _t

import tensorflow as _tf
from tensorflow import _op

_LSTMStateTuple = collections.namedtuple('LSTMStateTuple', ('c', 'h'))


class LSTMStateTuple(_LSTMStateTuple):
    __slots__ = ()


class RNNCell(object):
    def zero_state(self, batch_size, dtype):
        raise NotImplementedError


class LSTMCell(RNNCell):
    def __init__(self, num_units, input_size=None, use_peepholes=False, cell_clip=None, initializer=None,
                 num_proj=None, proj_clip=None, forget_bias=1.0, state_is_tuple=True, activation=None, reuse=None):
        assert not use_peepholes and num_proj is None and cell_clip is None and state_is_tuple
        self._n = num_units
        self._fb = forget_bias
        self._init = initializer

    @property
    def state_size(self):
        return LSTMStateTuple(self._n, self._n)

    @property
    def output_size(self):
        return self._n

    def zero_state(self, batch_size, dtype):
        return LSTMStateTuple(_tf.zeros([batch_size, self._n]), _tf.zeros([batch_size, self._n]))

    def __call__(self, inputs, state, scope=None):
        c_prev, h_prev = state
        d_in = inputs.get_shape().as_list()[1]
        n, fb = self._n, self._fb
        with _tf.variable_scope(scope or 'lstm_cell'):
            w = _tf.get_variable('weights', [d_in + n, 4 * n], initializer=self._init)
            b = _tf.get_variable('biases', [4 * n], initializer=_tf.constant_initializer(0.0))

        def f(x, h, c, w_, b_):
            z = _t.cat([x, h], 1) @ w_ + b_
            i, j, fg, o = z[:, :n], z[:, n:2 * n], z[:, 2 * n:3 * n], z[:, 3 * n:]
            c2 = _t.sigmoid(fg + fb) * c + _t.sigmoid(i) * _t.tanh(j)
            h2 = _t.sigmoid(o) * _t.tanh(c2)
            return _t.stack([c2, h2])
        both = _op(f, [inputs, h_prev, c_prev, w, b], 'lstm_cell')
        bs = inputs.get_shape().as_list()[0]
        c2 = _op(lambda r: r[0], [both], 'lstm_c', [bs, n])
        h2 = _op(lambda r: r[1], [both], 'lstm_h', [bs, n])
        return h2, LSTMStateTuple(c2, h2)


BasicLSTMCell = LSTMCell


class DropoutWrapper(RNNCell):
    def __init__(self, cell, input_keep_prob=1.0, output_keep_prob=1.0, seed=None):
        self._cell, self._ik, self._ok = cell, input_keep_prob, output_keep_prob

    @property
    def state_size(self):
        return self._cell.state_size

    @property
    def output_size(self):
        return self._cell.output_size

    def zero_state(self, batch_size, dtype):
        return self._cell.zero_state(batch_size, dtype)

    def __call__(self, inputs, state, scope=None):
        # TF-1.0: dropout is skipped only when keep_prob is the python float 1.0
        if not (isinstance(self._ik, float) and self._ik == 1.0):
            inputs = _tf.nn.dropout(inputs, self._ik)
        out, new_state = self._cell(inputs, state, scope)
        if not (isinstance(self._ok, float) and self._ok == 1.0):
            out = _tf.nn.dropout(out, self._ok)
        return out, new_state


class MultiRNNCell(RNNCell):
    def __init__(self, cells, state_is_tuple=True):
        assert state_is_tuple
        self._cells = list(cells)

    @property
    def state_size(self):
        return tuple(c.state_size for c in self._cells)

    @property
    def output_size(self):
        return self._cells[-1].output_size

    def zero_state(self, batch_size, dtype):
        return tuple(c.zero_state(batch_size, dtype) for c in self._cells)

    def __call__(self, inputs, state, scope=None):
        cur = inputs
        new_states = []
        with _tf.variable_scope(scope or 'multi_rnn_cell'):
            for i, cell in enumerate(self._cells):
                with _tf.variable_scope('cell_%d' % i):
                    cur, ns = cell(cur, state[i])
                    new_states.append(ns)
        return cur, tuple(new_states)
