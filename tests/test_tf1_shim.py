"""Unit tests of oracle/tf1_shim (the TF-1.0 op restatement the golden generator runs the reference's model
code on): every semantic the pin relies on, against hand-computed values (SURVEY.md 8(c) KATs, Appendix C)."""
import os
import sys

import numpy as np
import pytest

SHIM = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'oracle', 'tf1_shim')


@pytest.fixture()
def tf():
    sys.path.insert(0, SHIM)
    saved = {k: v for k, v in sys.modules.items() if k == 'tensorflow' or k.startswith('tensorflow.')}
    for k in saved:
        del sys.modules[k]
    import tensorflow as tf_
    assert 'tf1_shim' in tf_.__file__
    tf_.reset_default_graph()
    yield tf_
    for k in [k for k in sys.modules if k == 'tensorflow' or k.startswith('tensorflow.')]:
        del sys.modules[k]
    sys.modules.update(saved)
    sys.path.remove(SHIM)


def test_kat1_pool_through_lookup_and_segment_sum(tf):
    """mulhot pooling = lookup + unsorted_segment_sum + div."""
    E = tf.Variable(np.array([[1, 2], [3, 4], [5, 6], [7, 8], [9, 10]], dtype=np.float32))
    flat = tf.constant([3, 4, 4, 0, 2, 1, 1], dtype=tf.int32)
    seg = tf.constant([0, 0, 0, 1, 1, 2, 3], dtype=tf.int32)
    lens = tf.constant(np.array([[3.], [1.], [1.], [1.]], dtype=np.float32))       # ids [2,0,1,3] -> lengths
    pooled = tf.div(tf.unsorted_segment_sum(tf.nn.embedding_lookup(E, flat), seg, 4), lens)
    out = tf.Session().run(pooled)
    # KAT-1 states the pooled rows for bag lengths (3, 2, 1, 1): recompute the segment sums by hand
    want = np.array([[7 + 9 + 9, 8 + 10 + 10], [1 + 5, 2 + 6], [3, 4], [3, 4]], dtype=np.float32) / np.array([[3.], [1.], [1.], [1.]])
    np.testing.assert_allclose(out, want, rtol=1e-6)


def test_kat2_sparse_softmax_cross_entropy(tf):
    logits = tf.constant(np.array([[3.2, 3.2, 8.8], [-2.3, -2.3, -4.7]], dtype=np.float32))
    ce = tf.nn.sparse_softmax_cross_entropy_with_logits(logits=logits, labels=tf.constant([2, 0], dtype=tf.int32))
    out = tf.Session().run(ce)
    np.testing.assert_allclose(out, [0.0073685, 0.7375075], rtol=2e-5, atol=1e-6)   # fp32 cancellation at 8.8


def test_kat5_adagrad_accumulator_and_duplicate_rows(tf):
    """acc0 = 0.1, acc += g^2, w -= lr g / sqrt(acc); rows hit twice get the SUMMED gradient, untouched rows stay."""
    E = tf.get_variable('E', [3, 2], dtype=tf.float32, initializer=tf.constant_initializer(0.0))
    x = tf.placeholder(tf.int32, [3], name='x')
    w = tf.constant(np.array([[0.5, -2.0], [0.5, -2.0], [1.0, 1.0]], dtype=np.float32))
    loss = tf.reduce_sum(tf.nn.embedding_lookup(E, x) * w)
    g = tf.gradients(loss, [E])
    up = tf.train.AdagradOptimizer(1.0).apply_gradients(zip(g, [E]))
    s = tf.Session()
    s.run(up, {x.name: [0, 2, 2]})                      # row 0: g = (.5, -2); row 2: g = (.5 + 1, -2 + 1); row 1 untouched
    got = E.numpy()
    np.testing.assert_allclose(got[0], [-0.5 / np.sqrt(0.35), 2.0 / np.sqrt(4.1)], rtol=1e-6)     # KAT-5
    np.testing.assert_allclose(got[0], [-0.8451542, 0.9877296], rtol=1e-6)
    np.testing.assert_allclose(got[1], [0.0, 0.0])
    np.testing.assert_allclose(got[2], [-1.5 / np.sqrt(0.1 + 2.25), 1.0 / np.sqrt(1.1)], rtol=1e-6)


def test_fetching_loss_with_update_returns_pre_update_loss(tf):
    v = tf.Variable(np.array([2.0], dtype=np.float32))
    loss = tf.reduce_sum(v * v)
    up = tf.train.GradientDescentOptimizer(0.25).apply_gradients(zip(tf.gradients(loss, [v]), [v]))
    s = tf.Session()
    _, l0 = s.run([up, loss])
    assert abs(float(l0) - 4.0) < 1e-6 and abs(float(v.numpy()[0]) - 1.0) < 1e-6     # 2 - 0.25 * 4


def test_top_k_ties_lower_index_first(tf):
    x = tf.constant(np.array([[1.0, 3.0, 3.0, 0.0, 3.0]], dtype=np.float32))
    vals, idx = tf.nn.top_k(x, 3)
    v, i = tf.Session().run([vals, idx])
    assert i.tolist() == [[1, 2, 4]] and v.tolist() == [[3.0, 3.0, 3.0]]


def test_dropout_formula_and_hook(tf):
    x = tf.constant(np.ones((2, 3), dtype=np.float32))
    kp = tf.placeholder(tf.float32, name='kp')
    y = tf.nn.dropout(x, kp)
    mask = np.array([[1, 0, 1], [0, 0, 1]], dtype=np.float32)
    tf.set_dropout_hook(lambda shape, keep, node=None: mask)
    s = tf.Session()
    np.testing.assert_allclose(s.run(y, {kp.name: 0.5}), mask / 0.5)
    np.testing.assert_allclose(s.run(y, {kp.name: 1.0}), np.ones((2, 3)))            # identity, hook not consulted


def test_clip_by_global_norm_indexed_slices_vs_dense(tf):
    """A table reached only through lookups contributes the norm of its UN-MERGED slices (IndexedSlices.values);
    a dense parameter its ordinary norm; the clipped gradients are g * clip / max(norm, clip)."""
    E = tf.Variable(np.zeros((4, 1), dtype=np.float32))
    W = tf.Variable(np.array([1.0], dtype=np.float32))
    ids = tf.constant([1, 1, 3], dtype=tf.int32)
    coef = tf.constant(np.array([[3.0], [4.0], [12.0]], dtype=np.float32))
    loss = tf.reduce_sum(tf.nn.embedding_lookup(E, ids) * coef) + tf.reduce_sum(W * 2.0)
    grads = tf.gradients(loss, [E, W])
    clipped, norm = tf.clip_by_global_norm(grads, 6.5)
    n, cE, cW = tf.Session().run([norm, clipped[0], clipped[1]])
    assert abs(float(n) - np.sqrt(9 + 16 + 144 + 4)) < 1e-5            # un-merged: 3^2 + 4^2, not (3 + 4)^2
    scale = 6.5 / np.sqrt(173.0)
    np.testing.assert_allclose(cE[:, 0], np.array([0, 7, 0, 12]) * scale, rtol=1e-6)    # the applied gradient IS merged
    np.testing.assert_allclose(cW, [2.0 * scale], rtol=1e-6)


def test_lstm_cell_gate_order_and_forget_bias(tf):
    from tensorflow.contrib import rnn
    cell = rnn.LSTMCell(2, state_is_tuple=True)
    x = tf.constant(np.array([[0.5, -1.0, 2.0]], dtype=np.float32))
    outs, state = rnn.static_rnn(cell, [x, x], dtype=tf.float32)
    g = tf.get_default_graph()
    W = np.linspace(-0.5, 0.5, 5 * 8).reshape(5, 8).astype(np.float32)
    g.by_name['rnn/lstm_cell/weights'].load(W)
    g.by_name['rnn/lstm_cell/biases'].load(np.zeros(8, dtype=np.float32))
    h2 = tf.Session().run(outs[1])
    sig = lambda z: 1.0 / (1.0 + np.exp(-z))                                           # noqa: E731
    h = np.zeros((1, 2)); c = np.zeros((1, 2))
    for _ in range(2):
        z = np.concatenate([x_np := np.array([[0.5, -1.0, 2.0]]), h], 1) @ W
        i, j, f, o = z[:, :2], z[:, 2:4], z[:, 4:6], z[:, 6:]
        c = sig(f + 1.0) * c + sig(i) * np.tanh(j)
        h = sig(o) * np.tanh(c)
    np.testing.assert_allclose(h2, h, rtol=1e-5, atol=1e-6)


def test_scatter_update_and_assign_are_staged(tf):
    m = tf.Variable([True] * 6, dtype=tf.bool, trainable=False)
    idx = tf.placeholder(tf.int32, shape=[None])
    val = tf.placeholder(tf.bool, shape=[None])
    setop = tf.scatter_update(m, idx, val)
    s = tf.Session()
    s.run(setop, {idx.name: [1, 4], val.name: [False, False]})
    assert m.numpy().tolist() == [True, False, True, True, False, True]
    lr = tf.Variable(2.0, trainable=False)
    decay = lr.assign(lr * 0.5)
    s.run(decay)
    assert abs(float(lr.numpy()) - 1.0) < 1e-7
