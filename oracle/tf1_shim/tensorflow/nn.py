"""tf.nn.* subset (see the package docstring: test infrastructure, semantics restated from TF-1.0 docs)."""
import numpy as _np
import torch as _t

import tensorflow as _tf
from tensorflow import _op, _shape_of


def relu(x, name=None):
    return _op(_t.relu, [x], 'relu', _shape_of(x))


def sigmoid(x, name=None):
    return _tf.sigmoid(x)


def tanh(x, name=None):
    return _tf.tanh(x)


def softmax(logits, dim=-1, name=None):
    return _op(lambda a: _t.softmax(a, dim=dim), [logits], 'softmax', _shape_of(logits))


def log_softmax(logits, dim=-1, name=None):
    return _op(lambda a: _t.log_softmax(a, dim=dim), [logits], 'log_softmax', _shape_of(logits))


def embedding_lookup(params, ids, partition_strategy='mod', name=None, validate_indices=True, max_norm=None):
    """params[ids] along axis 0 (single, unpartitioned params)."""
    if isinstance(params, (list, tuple)) and len(params) == 1:
        params = params[0]
    return _tf._lookup(params, ids, name or 'embedding_lookup', getattr(params, 'dtype', None))


def dropout(x, keep_prob, noise_shape=None, seed=None, name=None):
    """x / keep_prob * floor(keep_prob + U[0,1)).  The 0/1 mask comes from the graph's dropout hook when
    one is installed (so a test can inject the same masks into the CUDA path), else from torch's RNG."""
    node = None

    def f(a, k):
        k = float(k)
        if k == 1.0:                       # floor(1 + U) == 1: identity, no mask drawn
            return a
        g = _tf.get_default_graph()
        if g.dropout_hook is not None:
            m = _t.from_numpy(_np.asarray(g.dropout_hook(tuple(a.shape), k, node), dtype=_np.float32))
        else:
            m = _t.floor(k + _t.rand(a.shape))
        return a / k * m
    node = _op(f, [x, keep_prob], 'dropout', _shape_of(x))
    return node


def sparse_softmax_cross_entropy_with_logits(_sentinel=None, labels=None, logits=None, name=None):
    def f(lg, lb):
        return _t.logsumexp(lg, dim=-1) - lg.gather(-1, lb.to(_t.int64).unsqueeze(-1)).squeeze(-1)
    return _op(f, [logits, labels], 'sparse_xent', (_shape_of(logits) or [None])[:-1] or None)


def top_k(input, k=1, sorted=True, name=None):      # noqa: A002
    """values, indices of the k largest entries along the last axis; ties -> lower index first."""
    both = _op(lambda a: _t.sort(a, dim=-1, descending=True, stable=True), [input], 'top_k')
    vals = _op(lambda r: r[0][..., :k], [both], 'top_k_values')
    idx = _op(lambda r: r[1][..., :k], [both], 'top_k_indices', dtype=_t.int32)
    return vals, idx
