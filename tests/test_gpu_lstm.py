"""GPU parity of the LSTM path: the LSTM layer against a float64 autograd
restatement of TF-1.0 LSTMCell, and whole SeqModel training steps (embedding inputs -> LSTM ->
per-step catalog scoring -> sequence loss -> clip_by_global_norm -> Adagrad) against
oracle/torch_cpu_ref.py::TorchRefSeq with injected weights and dropout masks."""
import numpy as np
import pytest
import torch

from helpers import small_dataset, random_params, positives

pytestmark = pytest.mark.gpu


def _ref_lstm(X, W, b, keep, in_mask, out_mask):
    X = torch.tensor(X, dtype=torch.float64, requires_grad=True)
    W = torch.tensor(W, dtype=torch.float64, requires_grad=True)
    b = torch.tensor(b, dtype=torch.float64, requires_grad=True)
    T, mb, _ = X.shape
    H = W.shape[1] // 4
    h = torch.zeros((mb, H), dtype=torch.float64); c = torch.zeros((mb, H), dtype=torch.float64)
    outs = []
    for t in range(T):
        x = X[t] if keep == 1.0 else X[t] / keep * torch.tensor(in_mask[t], dtype=torch.float64)
        z = torch.cat([x, h], 1) @ W + b
        i, j, f, o = torch.split(z, H, 1)
        c = torch.sigmoid(f + 1.0) * c + torch.sigmoid(i) * torch.tanh(j)
        h = torch.sigmoid(o) * torch.tanh(c)
        outs.append(h if keep == 1.0 else h / keep * torch.tensor(out_mask[t], dtype=torch.float64))
    return torch.stack(outs, 0), X, W, b


@pytest.mark.parametrize('exact', [True, False])
@pytest.mark.parametrize('T,mb,d_in,H,keep', [(5, 16, 8, 8, 1.0), (7, 33, 12, 16, 0.5), (4, 130, 64, 64, 0.5),
                                              (1, 40, 32, 32, 1.0), (6, 200, 32, 32, 0.5), (9, 300, 128, 128, 0.5),
                                              (50, 512, 64, 64, 0.5), (12, 1024, 96, 128, 1.0)])
def test_lstm_layer_forward_backward(cuda, T, mb, d_in, H, keep, exact):
    """H in {32, 64, 128} on the tensor-core path runs the persistent cluster kernels (arx_lstm_seq_fwd / _bwd):
    1, 2 and 4 CTAs per cluster, ragged last row tile, T = 1."""
    import arecsys_b200  # noqa: F401
    from arecsys_b200 import _lib
    from arecsys_b200.lstm.lstm_layer import LSTMLayer
    rng = np.random.default_rng(T * 100 + mb)
    X = rng.standard_normal((T, mb, d_in)).astype(np.float32)
    W = (rng.standard_normal((d_in + H, 4 * H)) * (1.0 / np.sqrt(d_in + H))).astype(np.float32)   # |z| ~ 1 like a trained cell
    b = (rng.standard_normal(4 * H) * 0.1).astype(np.float32)
    im = np.floor(rng.random((T, mb, d_in)) + keep).astype(np.float32)
    om = np.floor(rng.random((T, mb, H)) + keep).astype(np.float32)
    dO = rng.standard_normal((T, mb, H)).astype(np.float32)
    out_r, Xr, Wr, br = _ref_lstm(X, W, b, keep, im, om)
    (out_r * torch.tensor(dO, dtype=torch.float64)).sum().backward()
    _lib.exact_fp32 = exact
    try:
        layer = LSTMLayer(d_in, H, cuda, W=W, b=b)
        dev = lambda a: torch.tensor(a, device='cuda')
        out = layer.forward(dev(X), keep, dev(im), dev(om))
        dX = layer.backward(dev(dO))
    finally:
        _lib.exact_fp32 = False
    tol = 2e-5 if exact else 3e-3
    assert bool(getattr(layer, '_seq', False)) == ((not exact) and H in (32, 64, 128)), 'persistent-kernel dispatch'

    def close(a, r, name):
        r = r.detach().numpy()
        err = np.abs(a.cpu().numpy() - r).max() / max(np.abs(r).max(), 1e-6)
        assert err < tol, (name, err)
    close(out, out_r, 'out'); close(dX, Xr.grad, 'dX'); close(layer.dW, Wr.grad, 'dW'); close(layer.db, br.grad, 'db')


def _seq_setup(loss, use_concat, no_user_id, sep=False, dim=8, mb=12, T=5, ns=None, keep=0.5, seed=0, n_items=30):
    import arecsys_b200  # noqa: F401
    from arecsys_b200.attributes.embed_attribute import EmbeddingAttribute
    from arecsys_b200.lstm.seqModel import SeqModel
    from oracle.torch_cpu_ref import TorchRefSeq
    n_users = 40
    ua, ia, _, l2i = small_dataset(n_users, n_items, 2, 15, 3, 5, seed, None, dim)
    params = random_params(ua, ia, dim, seed + 1, scale=0.4, item_output=sep)
    rng = np.random.default_rng(seed + 5)
    Fu = ua.num_features_cat + ua.num_features_mulhot - (1 if no_user_id else 0)
    Fi = ia.num_features_cat + ia.num_features_mulhot
    d_in = dim
    if use_concat:
        params['w_input_user'] = rng.uniform(-.4, .4, (Fu * dim, dim)).astype(np.float32)
        params['w_input_item'] = rng.uniform(-.4, .4, (Fi * dim, dim)).astype(np.float32)
    params['lstm_w'] = rng.uniform(-.4, .4, (d_in + dim, 4 * dim)).astype(np.float32)
    params['lstm_b'] = np.zeros(4 * dim, dtype=np.float32)
    START = n_items                                        # the trailing pseudo-entity
    l2i_d = {int(v): int(l2i[v]) for v in range(len(l2i))}
    i2l_d = {v: k for k, v in l2i_d.items()}
    i2l_d[START] = 0                                       # lstm/run.py:276-278
    emb = EmbeddingAttribute(ua, ia, mb, ns, T, sep, i2l_d, l2i_d, params=params)
    model = SeqModel([3, T], dim, 1, 5.0, mb, 0.5, 0.83, emb, withAdagrad=True, dropoutRate=keep, START_ID=START,
                     loss=loss, use_concat=use_concat, no_user_id=no_user_id, topk_n=5, params=params)
    ref = TorchRefSeq(ua, ia, params, l2i_d, i2l_d, loss=loss, keep_prob=keep, learning_rate=0.5, n_sampled=ns,
                      dtype=torch.float64, size=dim, use_concat=use_concat, no_user_id=no_user_id,
                      max_gradient_norm=5.0, item_output=sep)
    return model, ref, emb, rng, (n_users, n_items, START, dim, mb, T)


def _batch(rng, n_users, n_items, START, mb, T):
    users = rng.integers(0, n_users, mb).tolist()
    seqs = [rng.integers(0, n_items, rng.integers(1, T + 1)).tolist() for _ in range(mb)]
    inp = [[START] + s[:-1] + [START] * (T - len(s)) for s in seqs]
    out = [s + [START] * (T - len(s)) for s in seqs]
    w = [[1.0] * len(s) + [0.0] * (T - len(s)) for s in seqs]
    tm = lambda l: [[l[j][i] for j in range(mb)] for i in range(T)]
    return users, tm(inp), tm(out), tm(w), seqs


@pytest.mark.parametrize('exact', [True, False])
@pytest.mark.parametrize('loss,use_concat,no_user_id,sep', [('ce', False, False, False), ('ce', False, True, False),
                                                            ('ce', True, False, False), ('warp', False, False, False),
                                                            ('mw', False, False, False), ('ce', False, False, True),
                                                            ('ce', True, False, True)])
def test_seqmodel_training_steps_match_reference(cuda, loss, use_concat, no_user_id, sep, exact):
    from arecsys_b200 import _lib
    ns = 10 if loss == 'mw' else None
    model, ref, emb, rng, (n_users, n_items, START, dim, mb, T) = _seq_setup(loss, use_concat, no_user_id, sep, ns=ns)
    ltol, ptol = (2e-4, 2e-3) if exact else (2e-3, 2e-2)
    _lib.exact_fp32 = exact
    try:
        for it in range(3):
            users, inp, out, w, seqs = _batch(rng, n_users, n_items, START, mb, T)
            pos = {u: sorted(set(s)) for u, s in zip(users, seqs)}
            emb.prepare_warp(pos, pos); ref.pos, ref.pos_eval = pos, pos
            sampled = [int(v) for v in rng.permutation(n_items)[:ns]] if (ns and it != 1) else None
            d_in = dim
            im = np.floor(rng.random((T, mb, d_in)) + 0.5).astype(np.float32)
            om = np.floor(rng.random((T, mb, dim)) + 0.5).astype(np.float32)
            lg = model.step(None, users, inp, out, w, 1, item_sampled=sampled,
                            masks=(torch.tensor(im, device='cuda'), torch.tensor(om, device='cuda')))
            lr_ = ref.step_seq(users, inp, out, w, item_sampled=sampled, masks=(im, om))
            assert abs(lg - lr_) <= ltol * max(1.0, abs(lr_)), (it, lg, lr_)
            assert abs(float(model.last_gnorm) - ref.last_gnorm) <= 5 * ltol * max(1.0, ref.last_gnorm), (float(model.last_gnorm), ref.last_gnorm)
            dense = model.dense_params()
            for k, v in ref.p.items():
                got = (emb.params[k] if k in emb.params else dense[k][0]).cpu().numpy()
                want = v.detach().numpy()
                err = np.abs(got.reshape(want.shape) - want).max()
                assert err <= ptol * max(1.0, np.abs(want).max()), (k, it, err)
        users, inp, out, w, seqs = _batch(rng, n_users, n_items, START, mb, T)
        pos = {u: sorted(set(s)) for u, s in zip(users, seqs)}
        emb.prepare_warp(pos, pos); ref.pos, ref.pos_eval = pos, pos
        eg = model.step(None, users, inp, out, w, 1, forward_only=True)
        er = ref.step_seq(users, inp, out, w, forward_only=True)
        assert abs(eg - er) <= ltol * max(1.0, abs(er))
    finally:
        _lib.exact_fp32 = False


def test_seqmodel_recommend_and_batching(cuda):
    model, ref, emb, rng, (n_users, n_items, START, dim, mb, T) = _seq_setup('ce', False, False)
    data = [[(1, [3, 4]), (2, [5])], [(3, [1, 2, 3, 4]), (4, [7, 8, 9, 1, 2])]]
    model.batch_size = 2
    users, inp, out, w, fin = model.get_batch(data, 0, start_id=0)
    assert users == [1, 2] and inp == [[START, START], [3, START], [START, START]] and out[0] == [3, 5]
    assert w == [[1.0, 1.0], [1.0, 0.0], [0.0, 0.0]] and fin
    users, inp, positions, valids, fin = model.get_batch_recommend(data, 1, start_id=1)
    assert users == [4, 0] and positions == [4, T - 1] and valids == [1, 0] and fin
    res = model.step_recommend(None, [3, 4], [[1, 7], [2, 8], [3, 9], [4, 1], [START, 2]], [3, 4], 1)
    assert len(res) == 2 and res[0][0] == 3 and res[0][1].shape == (5,) and res[0][2].shape == (5,)
    assert (np.diff(res[0][1]) <= 1e-7).all() and 0 < res[0][1].sum() <= 1.0 + 1e-5      # sorted probabilities


@pytest.mark.parametrize('use_concat,sep', [(False, False), (True, True)])
def test_seqmodel_ce_fused_tensor_core_path(cuda, use_concat, sep):
    """loss=ce with shapes the fused kernels take (T*mb and V multiples of 4, dim 32): all positions are scored
    and reduced by arx_ce_fwd / arx_ce_bwd without materialising [T*mb, V] logits; clip norm through the
    merged-row path (arx_pool_bwd_apply OPT_NONE + arx_rows_sumsq)."""
    from arecsys_b200 import _lib
    model, ref, emb, rng, (n_users, n_items, START, dim, mb, T) = _seq_setup('ce', use_concat, False, sep, dim=32,
                                                                             n_items=36)
    assert _lib.ce_supported(T * mb, n_items, dim)
    calls = []
    orig = _lib.ce_fwd
    _lib.ce_fwd = lambda *a, **k: (calls.append(1), orig(*a, **k))[1]
    try:
        for it in range(3):
            users, inp, out, w, seqs = _batch(rng, n_users, n_items, START, mb, T)
            im = np.floor(rng.random((T, mb, dim)) + 0.5).astype(np.float32)
            om = np.floor(rng.random((T, mb, dim)) + 0.5).astype(np.float32)
            lg = model.step(None, users, inp, out, w, 1, masks=(torch.tensor(im, device='cuda'), torch.tensor(om, device='cuda')))
            lr_ = ref.step_seq(users, inp, out, w, masks=(im, om))
            assert abs(lg - lr_) <= 2e-3 * max(1.0, abs(lr_)), (it, lg, lr_)
            assert abs(float(model.last_gnorm) - ref.last_gnorm) <= 1e-2 * max(1.0, ref.last_gnorm)
            dense = model.dense_params()
            for k, v in ref.p.items():
                got = (emb.params[k] if k in emb.params else dense[k][0]).cpu().numpy()
                want = v.detach().numpy()
                err = np.abs(got.reshape(want.shape) - want).max()
                assert err <= 2e-2 * max(1.0, np.abs(want).max()), (k, it, err)
        users, inp, out, w, seqs = _batch(rng, n_users, n_items, START, mb, T)
        eg = model.step(None, users, inp, out, w, 1, forward_only=True)
        er = ref.step_seq(users, inp, out, w, forward_only=True)
        assert abs(eg - er) <= 2e-3 * max(1.0, abs(er))
    finally:
        _lib.ce_fwd = orig
    assert len(calls) == 4, 'the fused CE path did not run'


@pytest.mark.parametrize('T,mb,H', [(7, 260, 32), (20, 512, 64), (50, 4096, 128)])
def test_lstm_persistent_kernels_match_per_step_kernels(cuda, T, mb, H, monkeypatch):
    """The cluster kernels against the round-1 per-step path (one tcgen05 GEMM + one gate kernel per time step), both on
    tf32 operands: same recurrence, so they agree far inside the parity bar — at the north-star shape too
    (mb 4096, d = H = 128, T 50)."""
    import arecsys_b200  # noqa: F401
    from arecsys_b200.lstm.lstm_layer import LSTMLayer
    rng = np.random.default_rng(T + mb)
    d_in = H
    X = torch.tensor(rng.standard_normal((T, mb, d_in)).astype(np.float32), device='cuda')
    W = (rng.standard_normal((d_in + H, 4 * H)) * (1.0 / np.sqrt(d_in + H))).astype(np.float32)
    b = (rng.standard_normal(4 * H) * 0.1).astype(np.float32)
    dO = torch.tensor(rng.standard_normal((T, mb, H)).astype(np.float32), device='cuda')
    res = []
    for seq in ('1', '0'):
        monkeypatch.setenv('ARX_LSTM_SEQ', seq)
        layer = LSTMLayer(d_in, H, cuda, W=W, b=b)
        out = layer.forward(X, 1.0)
        dX = layer.backward(dO)
        assert bool(layer._seq) == (seq == '1')
        res.append((out.clone(), dX.clone(), layer.dW.clone(), layer.db.clone()))
    for a, r, name in zip(res[0], res[1], ('out', 'dX', 'dW', 'db')):
        err = float((a - r).abs().max() / r.abs().max().clamp_min(1e-6))
        assert err < 2e-3, (name, err)
