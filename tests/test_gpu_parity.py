"""GPU parity tests: every C-ABI kernel and the HMF training step against the NumPy oracle on
identical seeded inputs and injected weights / dropout masks.  Bars: bit-exact for index
results; fp32 within the stated tolerance (north_star: 1e-3 relative on loss/logits; the
exact-fp32 SIMT path is held to 1e-5)."""
import numpy as np
import pytest
import torch

from helpers import small_dataset, random_params, positives, kat_item_attributes
from oracle import np_oracle as O

pytestmark = pytest.mark.gpu

RTOL = 2e-5      # fp32 kernels vs float64 oracle


def _dicts(l2i):
    l2i_d = {int(v): int(l2i[v]) for v in range(len(l2i))}
    i2l_d = {int(l2i[v]): int(v) for v in range(len(l2i))}
    return i2l_d, l2i_d


def _build(dim, mb=16, n_users=60, n_items=50, n_mulhot=3, vocab_m=25, max_len=7, seed=0, n_sampled=None,
           loss='ce', nonlinear='linear', keep_prob=1.0, lr=0.3, logit_size=None, hidden=5, loss_func='log',
           exp_p=1.005):
    from arecsys_b200.hmf.hmf_model import LatentProductModel
    ua, ia, i2l, l2i = small_dataset(n_users, n_items, n_mulhot, vocab_m, 3, max_len, seed, logit_size, dim)
    params = random_params(ua, ia, dim, seed + 1, mlp_hidden=hidden if nonlinear != 'linear' else None)
    i2l_d, l2i_d = _dicts(l2i)
    model = LatentProductModel(n_users, n_items, dim, 1, mb, lr, 1.0, ua, ia, i2l_d, l2i_d, loss_function=loss,
                               nonlinear=nonlinear, dropout=keep_prob, n_sampled=n_sampled, hidden_size=hidden,
                               loss_func=loss_func, loss_exp_p=exp_p, params=params, top_N_items=10)
    emb = O.OracleEmbeddingAttribute(ua, ia, mb, n_sampled, params, item_ind2logit_ind=i2l_d,
                                     logit_ind2item_ind=l2i_d, dtype=np.float64)
    omodel = O.OracleHMF(emb, loss=loss, nonlinear=nonlinear, keep_prob=keep_prob, learning_rate=lr,
                         loss_func=loss_func, loss_exp_p=exp_p)
    return model, omodel, ua, ia


@pytest.mark.parametrize('dim', [4, 6, 8, 20, 32, 64, 128, 256])
def test_pool_fwd_matches_oracle(cuda, dim):
    model, om, ua, ia = _build(dim)
    m, e = model.att_emb, om.emb
    rng = np.random.default_rng(3)
    ids = np.concatenate([rng.integers(0, 50, 37), [50]])            # includes the START pseudo-entity
    from arecsys_b200._lib import POOL_MEAN, POOL_CONCAT
    out, b, _ = m.pool('item', m._ids(ids), POOL_MEAN, True)
    cat, mul = e._tables('item', ia); bc, bm = e._biases('item', ia)
    P, bp = O.pool_entities(e, cat, mul, bc, bm, ia, ids)
    np.testing.assert_allclose(out.cpu().numpy(), P, rtol=RTOL, atol=1e-6)
    np.testing.assert_allclose(b.cpu().numpy(), bp, rtol=RTOL, atol=1e-6)
    outc, _, _ = m.pool('item', m._ids(ids), POOL_CONCAT, True)
    cc, bb = e._get_embedded(cat, mul, bc, bm, ids, ia, concatenation=True)
    np.testing.assert_allclose(outc.cpu().numpy(), cc, rtol=RTOL, atol=1e-6)
    # no_attribute keeps categorical attribute 0 only; no_id drops it (embed_attribute.py:356-373)
    o2, _, _ = m.pool('item', m._ids(ids), POOL_MEAN, True, no_attribute=True)
    c2, m2, _ = e._get_embedded(cat, mul, bc, bm, ids, ia, concatenation=False, no_attribute=True)
    np.testing.assert_allclose(o2.cpu().numpy(), np.mean(np.stack(c2 + m2), 0), rtol=RTOL, atol=1e-6)
    ucat, umul = e._tables('user', ua)
    uids = rng.integers(0, 60, 20)
    o3, _, _ = m.pool('user', m._ids(uids), POOL_CONCAT, False, no_id=True)
    c3, _ = e._get_embedded(ucat, umul, None, None, uids, ua, concatenation=True, no_id=False)
    np.testing.assert_allclose(o3.cpu().numpy(), c3[:, dim:], rtol=RTOL, atol=1e-6)   # attribute 0 dropped
    # reference quirk (:356-366): no_id with a single categorical attribute returns zeros [mb, dim]
    z, _ = m.get_batch_user(1.0, concat=True, no_id=True, u_inds=uids)
    z0, _ = e._get_embedded(ucat, umul, None, None, uids, ua, concatenation=True, no_id=True)
    assert z.shape == z0.shape and float(z.abs().sum()) == 0.0 and not z0.any()


def test_reference_facing_api_get_batch_user_item(cuda):
    model, om, ua, ia = _build(8)
    m, e = model.att_emb, om.emb
    m.input_steps = 1
    users = list(range(16)); items = [list(range(16, 32))]
    m.add_input({}, users, items)
    u, ub = m.get_batch_user(1.0, concat=False)
    ou, _ = e.get_batch_user(users, 1.0, False)
    np.testing.assert_allclose(u.cpu().numpy(), ou, rtol=RTOL, atol=1e-6)
    lst, b = m.get_batch_item('input0', 16, concat=False)
    olst, ob = e.get_batch_item(items[0], concat=False)
    assert len(lst) == len(olst) == ia.num_features_cat + ia.num_features_mulhot
    for a, c in zip(lst, olst):
        np.testing.assert_allclose(a.cpu().numpy(), c, rtol=RTOL, atol=1e-6)
    np.testing.assert_allclose(b.cpu().numpy(), ob, rtol=RTOL, atol=1e-6)
    assert m.get_user_model_size(concat=True) == 8 * (ua.num_features_cat + ua.num_features_mulhot)
    assert m.get_item_model_size(concat=False) == 8


def test_flat_index_bit_exact_kat1_and_random(cuda):
    model, om, ua, ia = _build(8)
    m, e = model.att_emb, om.emb
    rng = np.random.default_rng(5)
    ids = rng.integers(0, 51, 64)
    for f in range(ia.num_features_mulhot):
        flat, seg = m.flat_indices('item', ia.num_features_cat + f, ids)
        oi, os_ = e.flat_indices(ia, f, ids)
        assert np.array_equal(flat.cpu().numpy().astype(np.int64), oi)
        assert np.array_equal(seg.cpu().numpy().astype(np.int64), os_)
    # KAT-1
    from arecsys_b200.attributes.embed_attribute import EmbeddingAttribute
    k = kat_item_attributes()
    ke = EmbeddingAttribute(k, k, 4, None, logit_ind2item_ind={0: 0, 1: 1, 2: 2}, item_ind2logit_ind={0: 0, 1: 1, 2: 2},
                            params={'userembed_mulhot_0': [[1, 2], [3, 4], [5, 6], [7, 8], [9, 10]],
                                    'itemembed_mulhot_0': [[1, 2], [3, 4], [5, 6], [7, 8], [9, 10]],
                                    'item_bias_mulhot_0': [[.1], [.2], [.3], [.4], [.5]]})
    flat, seg = ke.flat_indices('item', 0, [2, 0, 1, 3])
    assert flat.cpu().tolist() == [3, 4, 4, 0, 2, 1, 1] and seg.cpu().tolist() == [0, 0, 0, 1, 1, 2, 3]
    out, b, _ = ke.pool('item', ke._ids([2, 0, 1, 3]), 0, True)
    np.testing.assert_allclose(out.cpu().numpy(), [[25 / 3, 28 / 3], [3, 4], [3, 4], [3, 4]], rtol=1e-6)
    # KAT-2
    U = torch.tensor([[1, 0], [.5, -1]], dtype=torch.float32, device='cuda')
    logits = ke.get_prediction(U)
    np.testing.assert_allclose(logits.cpu().numpy(), [[3.2, 3.2, 8.8], [-2.3, -2.3, -4.7]], rtol=1e-6)
    ce = ke.compute_loss(logits, [2, 0], 'ce', want_grad=False)
    np.testing.assert_allclose(ce.cpu().numpy(), [0.0073685, 0.7375075], atol=2e-6)


@pytest.mark.parametrize('dim,mode', [(8, 0), (8, 1), (20, 0), (6, 0), (128, 0), (128, 1)])
def test_pool_bwd_row_gradients_match_oracle(cuda, dim, mode):
    model, om, ua, ia = _build(dim)
    m, e = model.att_emb, om.emb
    rng = np.random.default_rng(7)
    ids1 = rng.integers(0, 51, 40); ids2 = rng.integers(0, 51, 24)    # two lookups, duplicates across both
    F = ia.num_features_cat + ia.num_features_mulhot
    w = dim * (F if mode == 1 else 1)
    d1 = rng.standard_normal((40, w)).astype(np.float32); d2 = rng.standard_normal((24, w)).astype(np.float32)
    b1 = rng.standard_normal(40).astype(np.float32); b2 = rng.standard_normal(24).astype(np.float32)
    rngA = m.sets['item'].attr_range()
    m.push_grad('item', rngA, m._ids(ids1), mode, torch.tensor(d1, device='cuda'), torch.tensor(b1, device='cuda'))
    m.push_grad('item', rngA, m._ids(ids2), mode, torch.tensor(d2, device='cuda'), torch.tensor(b2, device='cuda'))
    got = m.row_gradients('item')
    cat, mul = e._tables('item', ia)
    sc, sm = [t.shape for t in cat], [t.shape for t in mul]
    tot = None
    for ids, d, b in ((ids1, d1, b1), (ids2, d2, b2)):
        if mode == 0:
            g = O.pool_backward(ia, ids, d.astype(np.float64), b.astype(np.float64), sc, sm)
        else:   # concat: attribute f reads its own column block and is not divided by F
            g = None
            for f in range(F):
                blk = np.zeros((len(ids), dim)); blk[:] = d[:, f * dim:(f + 1) * dim]
                gi = O.pool_backward(ia, ids, blk * F, b.astype(np.float64), sc, sm)   # bias stays a mean
                keep = lambda lst, k, off: [a if (i + off) == k else np.zeros_like(a) for i, a in enumerate(lst)]
                gi = (keep(gi[0], f, 0), keep(gi[1], f, len(sc)), keep(gi[2], f, 0), keep(gi[3], f, len(sc)))
                g = gi if g is None else tuple([a + c for a, c in zip(x, y)] for x, y in zip(g, gi))
        tot = g if tot is None else tuple([a + c for a, c in zip(x, y)] for x, y in zip(tot, g))
    gc, gm, gbc, gbm = tot
    for i in range(ia.num_features_cat):
        np.testing.assert_allclose(got['itemembed_cat_%d' % i].cpu().numpy(), gc[i], rtol=1e-4, atol=1e-5)
        np.testing.assert_allclose(got['item_bias_cat_%d' % i].cpu().numpy(), gbc[i], rtol=1e-4, atol=1e-5)
    for i in range(ia.num_features_mulhot):
        np.testing.assert_allclose(got['itemembed_mulhot_%d' % i].cpu().numpy(), gm[i], rtol=1e-4, atol=1e-5)
        np.testing.assert_allclose(got['item_bias_mulhot_%d' % i].cpu().numpy(), gbm[i], rtol=1e-4, atol=1e-5)
    for t in m.touch.values():                     # the plan leaves the scratch arrays clean
        assert int(t.abs().sum().item()) == 0


def test_deterministic_scatter_mode_fixes_the_bucket_order(cuda):
    """EmbeddingAttribute.deterministic (ARX_DETERMINISTIC=1): every bucket of the backward plan is put into ascending
    order of the contributing arena row, so the de-duplicated row gradients (hot rows with thousands of contributions
    included) are bit-identical from run to run, and still equal to the default mode within fp32 rounding."""
    model, om, ua, ia = _build(128)
    m = model.att_emb
    rng = np.random.default_rng(3)
    ids = rng.integers(0, 8, 6000)                   # 6000 lookups over 8 items: every row is hot (chunked) or heavy
    d = torch.tensor(rng.standard_normal((6000, 128)).astype(np.float32), device='cuda')
    b = torch.tensor(rng.standard_normal(6000).astype(np.float32), device='cuda')
    rngA = m.sets['item'].attr_range()
    runs = []
    for det in (True, True, True, False):
        m.deterministic = det
        m.push_grad('item', rngA, m._ids(ids), 0, d, b)
        runs.append({k: v.clone() for k, v in m.row_gradients('item').items()})
        if det:
            plan = m.sets['item']._scratch                 # the plan row_gradients just built (buckets stay in place)
            nu, occ = int(plan.counters[0].item()), int(plan.counters[1].item())
            base, cnt = plan.row_base[:nu].cpu().numpy(), plan.row_cnt[:nu].cpu().numpy()
            src = plan.bucket_src[:occ].cpu().numpy()
            assert cnt.max() > 64                          # hot rows present
            assert all(np.all(np.diff(src[b0:b0 + c]) >= 0) for b0, c in zip(base, cnt))
    m.deterministic = False
    for k in runs[0]:
        assert torch.equal(runs[0][k], runs[1][k]) and torch.equal(runs[0][k], runs[2][k]), k
        assert torch.allclose(runs[0][k], runs[3][k], rtol=1e-4, atol=1e-4), k


@pytest.mark.parametrize('shape', [(5, 7, 3), (64, 64, 16), (130, 77, 33), (16, 300, 128)])
def test_gemm_variants(cuda, shape):
    from arecsys_b200._lib import call
    M, N, K = shape
    rng = np.random.default_rng(11)
    A = rng.standard_normal((M, K)).astype(np.float32); B = rng.standard_normal((N, K)).astype(np.float32)
    bias = rng.standard_normal(N).astype(np.float32)
    dA, dB, db = (torch.tensor(x, device='cuda') for x in (A, B, bias))
    C = torch.empty((M, N), device='cuda')
    call('arx_gemm', dA.data_ptr(), dB.data_ptr(), C.data_ptr(), M, N, K, 0, 1, db.data_ptr(), 1.0, 0.0)
    np.testing.assert_allclose(C.cpu().numpy(), A.astype(np.float64) @ B.T.astype(np.float64) + bias, rtol=1e-4, atol=1e-4)
    Bn = torch.tensor(np.ascontiguousarray(B.T), device='cuda')          # [K, N]
    call('arx_gemm', dA.data_ptr(), Bn.data_ptr(), C.data_ptr(), M, N, K, 0, 0, None, 2.0, 0.0)
    np.testing.assert_allclose(C.cpu().numpy(), 2 * (A.astype(np.float64) @ B.T), rtol=1e-4, atol=1e-4)
    At = torch.tensor(np.ascontiguousarray(A.T), device='cuda')          # [K, M]
    call('arx_gemm', At.data_ptr(), Bn.data_ptr(), C.data_ptr(), M, N, K, 1, 0, None, 1.0, 0.0)
    np.testing.assert_allclose(C.cpu().numpy(), A.astype(np.float64) @ B.T, rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize('loss', ['ce', 'warp', 'rs', 'rs-sig', 'rs-sig2', 'bbpr', 'mw'])
@pytest.mark.parametrize('V', [9, 300, 40000])
def test_loss_rows_match_oracle(cuda, loss, V):
    from arecsys_b200 import _lib
    rng = np.random.default_rng(13)
    mb = 6
    x = rng.standard_normal((mb, V)).astype(np.float32) * 2
    tgt = rng.integers(0, V, mb).astype(np.int32)
    ts = rng.standard_normal(mb).astype(np.float32)
    n_rows = 10
    pos = [np.unique(rng.integers(0, V, rng.integers(0, 6))) for _ in range(n_rows)]
    prow = rng.integers(0, n_rows, mb).astype(np.int32)
    if loss != 'mw':
        for b in range(mb):                                  # training mask contains the target itself
            pos[prow[b]] = np.unique(np.append(pos[prow[b]], tgt[b]))
    pptr = np.concatenate([[0], np.cumsum([len(p) for p in pos])]).astype(np.int32)
    pidx = np.concatenate(pos + [np.zeros(0, dtype=np.int64)]).astype(np.int32)
    mask = np.ones((mb, V), dtype=bool)
    for b in range(mb):
        mask[b, pos[prow[b]]] = False
    scale = rng.uniform(0.1, 1.0, mb).astype(np.float32)
    l, d, dts = O.loss_and_grad(x.astype(np.float64), ts.astype(np.float64) if loss == 'mw' else tgt.astype(np.int64),
                                loss, mask, 'log', 1.005)
    dx = torch.tensor(x, device='cuda'); out = torch.empty(mb, device='cuda'); dd = torch.empty_like(dx)
    dt = torch.empty(mb, device='cuda'); rank = torch.empty(mb, dtype=torch.int64, device='cuda')
    # keep every device buffer alive until the kernel has run (the caching allocator would
    # hand a freed temporary to the next torch.tensor() before the launch)
    g_tgt, g_ts, g_prow, g_pptr, g_scale = (torch.tensor(a, device='cuda') for a in (tgt, ts, prow, pptr, scale))
    g_pidx = torch.tensor(pidx if len(pidx) else np.zeros(1, dtype=np.int32), device='cuda')
    _lib.call('arx_loss_rows', dx.data_ptr(), mb, V, V, g_tgt.data_ptr(), g_ts.data_ptr(), g_prow.data_ptr(),
              g_pptr.data_ptr(), g_pidx.data_ptr(), _lib.LOSS_KIND[loss], 0, 1.005,
              g_scale.data_ptr(), out.data_ptr(), dd.data_ptr(), dt.data_ptr(), rank.data_ptr())
    np.testing.assert_allclose(out.cpu().numpy(), l, rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(dd.cpu().numpy(), d * scale[:, None], rtol=1e-3, atol=2e-6)
    if loss == 'mw':
        np.testing.assert_allclose(dt.cpu().numpy(), dts * scale, rtol=1e-3, atol=1e-5)
    if loss == 'warp':
        xt = x[np.arange(mb), tgt][:, None]
        assert np.array_equal(rank.cpu().numpy(), ((x > xt) & mask).sum(1))


@pytest.mark.parametrize('loss,nonlinear,dim', [('ce', 'linear', 8), ('ce', 'linear', 32), ('warp', 'linear', 8),
                                                ('rs', 'linear', 8), ('bbpr', 'linear', 20), ('mw', 'linear', 8),
                                                ('mw', 'linear', 128), ('ce', 'relu', 8), ('warp', 'tanh', 8),
                                                ('mw', 'tanh', 8), ('ce', 'linear', 128)])
@pytest.mark.parametrize('exact', [True, False])
def test_hmf_training_steps_match_oracle(cuda, loss, nonlinear, dim, exact):
    """exact=True: every contraction on the exact-fp32 SIMT kernel (tight bar);
    exact=False: the product default, tcgen05 tf32 contractions (north-star bar 1e-3)."""
    from arecsys_b200 import _lib
    _lib.exact_fp32 = exact
    try:
        _hmf_steps(loss, nonlinear, dim, 1e-4 if exact else 1e-3, 1e-3 if exact else 1e-2, 2e-5 if exact else 2e-3)
    finally:
        _lib.exact_fp32 = False


def _hmf_steps(loss, nonlinear, dim, ltol, prtol, patol):
    ns = 12 if loss == 'mw' else None
    mb = 16
    model, om, ua, ia = _build(dim, mb=mb, n_sampled=ns, loss=loss, nonlinear=nonlinear, keep_prob=0.5)
    rng = np.random.default_rng(17)
    for it in range(4):
        users = rng.integers(0, 60, mb); items = rng.integers(0, 50, mb)
        if it == 3:
            users[:4] = users[4]; items[:3] = items[5]           # duplicates inside the batch
        pos = positives(users, items, 60, rng, n_items=50)
        model.prepare_warp(pos, pos); om.emb.prepare_warp(pos, pos)
        sampled = [int(v) for v in rng.permutation(50)[:ns]] if (ns and it % 2 == 0) else None
        id2idx = {v: k for k, v in enumerate(sampled)} if sampled else None
        shapes = [(mb, dim), (mb, 5), (mb, dim)] if nonlinear != 'linear' else [(mb, dim)]
        masks = [np.floor(rng.random(s) + 0.5) for s in shapes]
        tm = [torch.tensor(k, dtype=torch.float32, device='cuda') for k in masks]
        l_gpu = model.step(None, list(users), list(items), None, sampled, id2idx, loss=loss, masks=tm)
        l_ora = om.step(list(users), list(items), sampled, id2idx, masks=masks)
        assert abs(l_gpu - l_ora) <= ltol * max(1.0, abs(l_ora)), (it, l_gpu, l_ora)
        for k, v in om.emb.p.items():
            got = (model.att_emb.params[k] if k in model.att_emb.params else model.dense[k].data).cpu().numpy()
            np.testing.assert_allclose(got.reshape(v.shape), v, rtol=prtol, atol=patol, err_msg='%s step %d' % (k, it))
    users = rng.integers(0, 60, mb); items = rng.integers(0, 50, mb)
    pos = positives(users, items, 60, rng, n_items=50)
    model.prepare_warp(pos, pos); om.emb.prepare_warp(pos, pos)
    e_gpu = model.step(None, list(users), list(items), forward_only=True, loss=loss)
    e_ora = om.step(list(users), list(items), forward_only=True)
    assert abs(e_gpu - e_ora) <= ltol * max(1.0, abs(e_ora))
    assert model.global_step.eval() == 4


def test_topk_and_recommend_match_oracle(cuda):
    model, om, ua, ia = _build(8, mb=16, loss='ce')
    users = list(range(16))
    rec = model.step(None, users, None, forward_only=True, recommend=True)
    idx, _ = om.top_k(users, 10)
    assert rec.shape == (16, 10) and np.array_equal(rec, idx)
    # ties -> lower index first; large V exercises the radix select
    from arecsys_b200._lib import call
    x = torch.zeros((3, 5000), device='cuda'); x[1, 4000] = 1.0; x[2, ::7] = 2.0
    out = torch.empty((3, 100), dtype=torch.int32, device='cuda'); val = torch.empty((3, 100), device='cuda')
    call('arx_topk_rows', x.data_ptr(), 3, 5000, 5000, 100, out.data_ptr(), val.data_ptr())
    o = out.cpu().numpy()
    assert o[0].tolist() == list(range(100))
    assert o[1].tolist() == [4000] + list(range(99))
    assert o[2].tolist() == list(range(0, 700, 7))
    xr = torch.randn((4, 30011), device='cuda')
    out = torch.empty((4, 64), dtype=torch.int32, device='cuda')
    call('arx_topk_rows', xr.data_ptr(), 4, 30011, 30011, 64, out.data_ptr(), None)
    ref = torch.topk(xr, 64, dim=1).indices.cpu().numpy()
    assert np.array_equal(out.cpu().numpy(), ref)


def test_checkpoint_roundtrip(cuda, tmp_path):
    model, om, ua, ia = _build(8, loss='ce', keep_prob=1.0)
    users = list(range(16)); items = list(range(16))
    model.step(None, users, items, loss='ce')
    p = model.saver.save(None, str(tmp_path / 'best.ckpt'), global_step=0)
    before = {k: v.clone() for k, v in model.att_emb.params.items()}
    model.step(None, users, items, loss='ce')
    model.saver.restore(None, p)
    for k, v in before.items():
        assert torch.equal(model.att_emb.params[k], v)
    assert model.global_step.eval() == 1


def test_cuda_graph_replay_matches_eager_steps(cuda):
    """The captured step (bench.py's launch mode) must train exactly like the eager step."""
    from arecsys_b200 import _lib
    _lib.exact_fp32 = True
    try:
        a, om, ua, ia = _build(128, mb=32, n_sampled=12, loss='mw', keep_prob=1.0)
        b, _, _, _ = _build(128, mb=32, n_sampled=12, loss='mw', keep_prob=1.0)
        rng = np.random.default_rng(23)
        users = [rng.integers(0, 60, 32).astype(np.int32) for _ in range(5)]
        items = [rng.integers(0, 50, 32).astype(np.int32) for _ in range(5)]
        pos = positives(np.concatenate(users), np.concatenate(items), 60, rng, n_items=50)
        a.prepare_warp(pos, pos); b.prepare_warp(pos, pos)
        s1 = [int(v) for v in rng.permutation(50)[:12]]; s2 = [int(v) for v in rng.permutation(50)[:12]]
        T = lambda x: torch.tensor(x, device='cuda')
        la = [a.step(None, users[0].tolist(), items[0].tolist(), None, s1, None, loss='mw')]
        b.att_emb.pass_sampled_items(s1)
        b.capture_step(T(users[0]), T(items[0]), loss='mw')          # runs step 0 eagerly, then captures
        lb = [float(b._g_loss.item())] if False else [None]
        for k in range(1, 5):
            smp = s2 if k == 3 else None
            la.append(a.step(None, users[k].tolist(), items[k].tolist(), None, smp, None, loss='mw'))
            if smp is not None:
                b.att_emb.pass_sampled_items(smp)
            lb.append(b.replay_step(T(users[k]), T(items[k])))
        for k in range(1, 5):
            assert abs(la[k] - lb[k]) <= 1e-5 * max(1.0, abs(la[k])), (k, la, lb)
        for k2, v in a.att_emb.params.items():
            torch.testing.assert_close(b.att_emb.params[k2], v, rtol=1e-5, atol=1e-6)
        assert a.global_step.eval() == b.global_step.eval() == 5
    finally:
        _lib.exact_fp32 = False


@pytest.mark.parametrize('dim,G', [(8, 2), (128, 3), (128, 4)])
def test_row_sharded_tables_on_one_device_match_unsharded(cuda, dim, G):
    """SURVEY 8(e) on one GPU: G row-sharded copies (row t on shard t % G, pre-partitioned bags) — the sum of
    their partial pooled vectors is the unsharded result, and their local Adagrad updates reassemble
    to the unsharded tables."""
    from arecsys_b200.attributes.embed_attribute import EmbeddingAttribute
    from arecsys_b200._lib import POOL_MEAN, OPT_ADAGRAD
    ua, ia, i2l, l2i = small_dataset(60, 50, 3, 25, 3, 7, 0, None, dim)
    params = random_params(ua, ia, dim, 5)
    i2l_d, l2i_d = _dicts(l2i)
    mk = lambda shard: EmbeddingAttribute(ua, ia, 16, None, item_ind2logit_ind=i2l_d, logit_ind2item_ind=l2i_d,
                                          params=params, shard=shard)
    full = mk(None)
    parts = [mk((G, r)) for r in range(G)]
    rng = np.random.default_rng(11)
    ids = np.concatenate([rng.integers(0, 51, 45), [50, 50, 3, 3]]).astype(np.int32)
    dout = torch.tensor(rng.standard_normal((len(ids), dim)).astype(np.float32), device='cuda')
    dbias = torch.tensor(rng.standard_normal(len(ids)).astype(np.float32), device='cuda')
    ref, refb, _ = full.pool('item', full._ids(ids), POOL_MEAN, True)
    acc, accb = torch.zeros_like(ref), torch.zeros_like(refb)
    for m in parts:
        o, b, _ = m.pool('item', m._ids(ids), POOL_MEAN, True)
        acc += o; accb += b
    np.testing.assert_allclose(acc.cpu().numpy(), ref.cpu().numpy(), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(accb.cpu().numpy(), refb.cpu().numpy(), rtol=1e-5, atol=1e-6)
    for m in [full] + parts:
        m.push_grad('item', m.sets['item'].attr_range(), m._ids(ids), POOL_MEAN, dout.clone(), dbias.clone())
        m.apply_gradients(0.3, OPT_ADAGRAD)
    for name in full.sets['item'].names + [b for b in full.sets['item'].bias_names if b]:
        want = full.params[name].cpu().numpy()
        got = np.empty_like(want)
        for r, m in enumerate(parts):
            got[r::G] = m.params[name].cpu().numpy()
        np.testing.assert_allclose(got, want, rtol=1e-5, atol=1e-6, err_msg=name)
        assert np.abs(want - np.asarray(params[name]).reshape(want.shape)).max() > 0    # something moved


@pytest.mark.parametrize('dim,n_items', [(32, 52), (64, 200), (128, 1000)])
def test_hmf_ce_fused_tensor_core_path_matches_oracle(cuda, dim, n_items):
    """loss=ce with a catalog size the fused kernels take (V % 4 == 0, dim in 32..128): scoring + softmax CE +
    gradients run in arx_ce_fwd / arx_ce_rowloss / arx_ce_bwd without materialising the logits."""
    from arecsys_b200 import _lib
    mb = 16
    model, om, ua, ia = _build(dim, mb=mb, n_items=n_items, loss='ce', keep_prob=0.5)
    assert _lib.ce_supported(mb, n_items, dim)
    calls = []
    orig = _lib.ce_fwd
    _lib.ce_fwd = lambda *a, **k: (calls.append(1), orig(*a, **k))[1]
    try:
        rng = np.random.default_rng(5)
        for it in range(4):
            users = rng.integers(0, 60, mb); items = rng.integers(0, n_items, mb)
            if it == 2:
                users[:4] = users[4]; items[:3] = items[5]
            masks = [np.floor(rng.random((mb, dim)) + 0.5)]
            tm = [torch.tensor(k, dtype=torch.float32, device='cuda') for k in masks]
            l_gpu = model.step(None, list(users), list(items), loss='ce', masks=tm)
            l_ora = om.step(list(users), list(items), masks=masks)
            assert abs(l_gpu - l_ora) <= 1e-3 * max(1.0, abs(l_ora)), (it, l_gpu, l_ora)
            for k, v in om.emb.p.items():
                got = model.att_emb.params[k].cpu().numpy()
                np.testing.assert_allclose(got.reshape(v.shape), v, rtol=1e-2, atol=2e-3, err_msg='%s step %d' % (k, it))
        e_gpu = model.step(None, list(users), list(items), forward_only=True, loss='ce')
        e_ora = om.step(list(users), list(items), forward_only=True)
        assert abs(e_gpu - e_ora) <= 1e-3 * max(1.0, abs(e_ora))
    finally:
        _lib.ce_fwd = orig
    assert len(calls) == 5, 'the fused CE path did not run'


def test_mw_prep_post_fused_glue_and_philox_dropout(cuda):
    """arx_mw_prep / arx_mw_post against their unfused definitions; the in-kernel Philox dropout mask is 0/1 with
    P(keep) = keep_prob, reproducible for a given (seed, step) and different from step to step."""
    from arecsys_b200._lib import call
    rng = np.random.default_rng(41)
    M, S, d, keep = 1000, 260, 128, 0.5
    f = lambda *sh: torch.tensor(rng.standard_normal(sh).astype(np.float32), device='cuda')
    u0, Pt, Ps, bt = f(M, d), f(M, d), f(S, d), f(M)
    mask = torch.tensor(np.floor(rng.random((M, d)) + keep).astype(np.float32), device='cuda')
    E = lambda *sh: torch.empty(sh, dtype=torch.float32, device='cuda')
    u, U_r, UT, ts, P_r, PT = E(M, d), E(M, d), E(d, M), E(M), E(S, d), E(d, S)
    call('arx_mw_prep', u0.data_ptr(), mask.data_ptr(), 1 / keep, None, None, Pt.data_ptr(), bt.data_ptr(), Ps.data_ptr(),
         M, S, d, u.data_ptr(), U_r.data_ptr(), UT.data_ptr(), ts.data_ptr(), P_r.data_ptr(), PT.data_ptr())
    want_u = u0 / keep * mask
    assert torch.equal(u, want_u)
    Ur_want = torch.empty_like(u); Pr_want = torch.empty_like(Ps)
    call('arx_round_tf32', want_u.data_ptr(), Ur_want.data_ptr(), want_u.numel())
    call('arx_round_tf32', Ps.data_ptr(), Pr_want.data_ptr(), Ps.numel())
    assert torch.equal(U_r, Ur_want) and torch.equal(P_r, Pr_want)
    assert torch.equal(UT, Ur_want.t().contiguous()) and torch.equal(PT, Pr_want.t().contiguous())
    torch.testing.assert_close(ts, (want_u.double() * Pt.double()).sum(1).float() + bt, rtol=1e-5, atol=1e-5)
    # adjoint glue
    dU, dts = f(M, d), f(M)
    du0, dPt = E(M, d), E(M, d)
    call('arx_mw_post', dU.data_ptr(), dts.data_ptr(), Pt.data_ptr(), u.data_ptr(), mask.data_ptr(), 1 / keep, M, d,
         du0.data_ptr(), dPt.data_ptr(), None)
    torch.testing.assert_close(du0, (dU + dts[:, None] * Pt) / keep * mask, rtol=1e-6, atol=1e-6)
    torch.testing.assert_close(dPt, dts[:, None] * u, rtol=1e-6, atol=1e-6)
    # in-kernel dropout draw
    state = torch.tensor([1234, 0], dtype=torch.int64, device='cuda')
    m1, m2, m3 = E(M, d), E(M, d), E(M, d)
    for mo in (m1, m2):
        call('arx_mw_prep', u0.data_ptr(), None, 1 / keep, state.data_ptr(), mo.data_ptr(), Pt.data_ptr(), bt.data_ptr(),
             Ps.data_ptr(), M, S, d, u.data_ptr(), U_r.data_ptr(), UT.data_ptr(), ts.data_ptr(), P_r.data_ptr(), PT.data_ptr())
    assert torch.equal(m1, m2) and torch.equal(u, u0 / keep * m1)                 # same (seed, step): same mask
    assert set(torch.unique(m1).tolist()) <= {0.0, 1.0}
    n = M * d
    assert abs(float(m1.mean()) - keep) < 5 * (keep * (1 - keep) / n) ** 0.5
    assert abs(float(m1.mean(0).std()) - (keep * (1 - keep) / M) ** 0.5) < 0.3 * (keep * (1 - keep) / M) ** 0.5   # no column structure
    call('arx_mw_post', dU.data_ptr(), dts.data_ptr(), Pt.data_ptr(), u.data_ptr(), m1.data_ptr(), 1 / keep, M, d,
         du0.data_ptr(), dPt.data_ptr(), state.data_ptr())
    assert state.tolist() == [1234, 1]                                              # the step counter advanced
    call('arx_mw_prep', u0.data_ptr(), None, 1 / keep, state.data_ptr(), m3.data_ptr(), Pt.data_ptr(), bt.data_ptr(),
         Ps.data_ptr(), M, S, d, u.data_ptr(), U_r.data_ptr(), UT.data_ptr(), ts.data_ptr(), P_r.data_ptr(), PT.data_ptr())
    agree = float((m1 == m3).float().mean())
    assert 0.45 < agree < 0.55                                                      # independent of the previous step's mask
