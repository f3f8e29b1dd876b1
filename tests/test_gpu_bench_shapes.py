"""GPU parity AT THE BENCHMARK'S OWN SHAPES (BASELINE.json configs C2 / C3 / C5): the code paths only those
shapes reach — flat-forward sub-group splits (> 1024 rows per CTA group), 64-long bags, entities straddling several
warps, several hash rounds of the block-aggregated plan kernels, Zipf head rows with thousands of contributions
(hot-row chunk phase of the apply kernel), pre-partitioned bags at G = 8 — against oracle/np_oracle.py
(float64) / oracle/torch_cpu_ref.py on the same seeded inputs, injected weights and dropout masks.

Tables are vocab 10^4 (not 10^5 / 10^6) so that the literal-order float64 oracle stays at seconds per step; batch,
dim, attribute count, bag-length law, token law, pool size and keep_prob are BASELINE's.
Reference lines: attributes/embed_attribute.py:350-417, mulhot_index.py:48-67, hmf/hmf_model.py:146-151.
"""
import numpy as np
import pytest
import torch

from helpers import random_params, positives
from oracle import np_oracle as O

pytestmark = pytest.mark.gpu

MB, DIM, S = 4096, 128, 1024
N_ENT, VOCAB = 10000, 10000


def _dataset(mean_len=12, seed=0, n_mulhot=8):
    import arecsys_b200  # noqa: F401
    from arecsys_b200.utils import synthetic
    ua, ia, i2l, l2i = synthetic.make_dataset(N_ENT, N_ENT, n_mulhot, VOCAB, mean_len, 64, 1.05, seed=seed)
    ua.set_model_size(DIM)
    ia.set_model_size(DIM)
    return ua, ia, i2l, l2i


def _dicts(l2i):
    l2i_d = {int(v): int(l2i[v]) for v in range(len(l2i))}
    return {v: k for k, v in l2i_d.items()}, l2i_d


def _embedding(ua, ia, l2i, params, shard=None, plan_agg=3, n_sampled=None):
    from arecsys_b200.attributes.embed_attribute import EmbeddingAttribute
    i2l_d, l2i_d = _dicts(l2i)
    m = EmbeddingAttribute(ua, ia, MB, n_sampled, item_ind2logit_ind=i2l_d, logit_ind2item_ind=l2i_d, params=params,
                           shard=shard)
    m.plan_agg = plan_agg
    return m


def _batch_ids(rng, n, start_item=True):
    """users uniform, items Zipf over a permutation (bench.py's interaction law) + the START pseudo-entity
    (which has no logit index: not as a training target)."""
    from arecsys_b200.utils import synthetic
    u, i = synthetic.make_interactions(N_ENT, N_ENT, n, seed=int(rng.integers(1 << 30)))
    u[-1] = N_ENT
    if start_item:
        i[-1] = N_ENT
    return u, i


@pytest.mark.parametrize('mean_len', [12, 40])
def test_c2_pool_fwd_matches_oracle(cuda, mean_len):
    """mean_len 12 = C2; mean_len 40 makes a 7-entity group hold ~2300 rows (> kFlatRows = 1024: the sub-group split
    of pool_fwd_flat_kernel) and clips a good share of the bags at 64 tokens."""
    ua, ia, i2l, l2i = _dataset(mean_len)
    params = random_params(ua, ia, DIM, 1, scale=0.1)
    m = _embedding(ua, ia, l2i, params)
    i2l_d, l2i_d = _dicts(l2i)
    e = O.OracleEmbeddingAttribute(ua, ia, MB, None, params, item_ind2logit_ind=i2l_d, logit_ind2item_ind=l2i_d)
    rng = np.random.default_rng(3)
    uids, iids = _batch_ids(rng, MB)
    if mean_len == 40:
        assert max(int(a.max()) for a in ia.mulhot_lengths) == 64
    from arecsys_b200._lib import POOL_MEAN
    for prefix, att, ids, bias in (('user', ua, uids, False), ('item', ia, iids, True)):
        out, b, _ = m.pool(prefix, m._ids(ids), POOL_MEAN, bias)
        cat, mul = e._tables(prefix, att)
        bc, bm = e._biases(prefix, att) if bias else (None, None)
        P, bp = O.pool_entities(e, cat, mul, bc, bm, att, ids)
        np.testing.assert_allclose(out.cpu().numpy(), P, rtol=2e-5, atol=1e-6)
        if bias:
            np.testing.assert_allclose(b.cpu().numpy(), bp, rtol=2e-5, atol=1e-6)
    # the integer part stays bit-exact at this size too (mulhot_index.py:48-67)
    flat, seg = m.flat_indices('item', ia.num_features_cat + 3, iids)
    oi, os_ = e.flat_indices(ia, 3, iids)
    assert np.array_equal(flat.cpu().numpy().astype(np.int64), oi)
    assert np.array_equal(seg.cpu().numpy().astype(np.int64), os_)


@pytest.mark.parametrize('plan_agg', [3, 0])
@pytest.mark.parametrize('mean_len', [12, 40])
def test_c2_row_gradients_match_oracle(cuda, plan_agg, mean_len):
    """De-duplicated row gradients (ARX_OPT_NONE) of the two item-side lookups of an `mw` step (sampled pool +
    targets, de-duplicated together) for both plan-kernel flavours."""
    ua, ia, i2l, l2i = _dataset(mean_len)
    params = random_params(ua, ia, DIM, 1, scale=0.1)
    m = _embedding(ua, ia, l2i, params, plan_agg=plan_agg)
    rng = np.random.default_rng(7)
    _, iids = _batch_ids(rng, MB)
    sids = rng.permutation(N_ENT)[:S].astype(np.int32)
    from arecsys_b200._lib import POOL_MEAN
    rngA = m.sets['item'].attr_range()
    grads_in = []
    for ids in (sids, iids):
        d = rng.standard_normal((len(ids), DIM)).astype(np.float32)
        b = rng.standard_normal(len(ids)).astype(np.float32)
        grads_in.append((ids, d, b))
        m.push_grad('item', rngA, m._ids(ids), POOL_MEAN, torch.tensor(d, device='cuda'), torch.tensor(b, device='cuda'))
    got = m.row_gradients('item')
    shapes_c = [(v, DIM) for v in ia._embedding_classes_list_cat]
    shapes_m = [(v, DIM) for v in ia._embedding_classes_list_mulhot]
    tot = None
    for ids, d, b in grads_in:
        g = O.pool_backward(ia, ids, d.astype(np.float64), b.astype(np.float64), shapes_c, shapes_m)
        tot = g if tot is None else tuple([a + c for a, c in zip(x, y)] for x, y in zip(tot, g))
    gc, gm, gbc, gbm = tot
    # Zipf head rows sum thousands of contributions: the bar is relative to the row's own magnitude
    def close(a, want, name):
        scale = np.abs(want).max(axis=-1, keepdims=True) + 1e-3
        err = (np.abs(a - want) / scale).max()
        assert err < 1e-4, (name, err)       # fp32 sums of thousands of terms in atomics-dependent order
    for i in range(ia.num_features_cat):
        close(got['itemembed_cat_%d' % i].cpu().numpy(), gc[i], 'cat%d' % i)
        close(got['item_bias_cat_%d' % i].cpu().numpy(), gbc[i], 'bcat%d' % i)
    for i in range(ia.num_features_mulhot):
        close(got['itemembed_mulhot_%d' % i].cpu().numpy(), gm[i], 'mul%d' % i)
        close(got['item_bias_mulhot_%d' % i].cpu().numpy(), gbm[i], 'bmul%d' % i)
    for t in m.touch.values():
        assert int(t.abs().sum().item()) == 0


@pytest.mark.parametrize('exact,plan_agg', [(False, 3), (True, 0)])
def test_c2_mw_training_steps_match_oracle(cuda, exact, plan_agg):
    """Three full `mw` training steps at C2's shape (mb 4096, dim 128, 9 attributes per side, pool 1024,
    keep_prob 0.5 with injected masks): loss <= 1e-3, and every table and Adagrad accumulator moves as the
    oracle's (compared on the UPDATE, which is what the step computes)."""
    from arecsys_b200 import _lib
    from arecsys_b200.hmf.hmf_model import LatentProductModel
    ua, ia, i2l, l2i = _dataset(12)
    params = random_params(ua, ia, DIM, 1, scale=0.1)
    i2l_d, l2i_d = _dicts(l2i)
    lr = 0.1
    model = LatentProductModel(N_ENT, N_ENT, DIM, 1, MB, lr, 1.0, ua, ia, i2l_d, l2i_d, loss_function='mw',
                               dropout=0.5, n_sampled=S, params=params)
    model.att_emb.plan_agg = plan_agg
    emb = O.OracleEmbeddingAttribute(ua, ia, MB, S, params, item_ind2logit_ind=i2l_d, logit_ind2item_ind=l2i_d)
    om = O.OracleHMF(emb, loss='mw', keep_prob=0.5, learning_rate=lr)
    init = {k: np.asarray(v, dtype=np.float64).copy() for k, v in emb.p.items()}
    rng = np.random.default_rng(17)
    ltol, utol = (1e-4, 1e-4) if exact else (1e-3, 1e-2)
    _lib.exact_fp32 = exact
    try:
        for it in range(3):
            users, items = _batch_ids(rng, MB, start_item=False)
            pos = positives(users, items, N_ENT + 1, rng, n_items=N_ENT)
            model.prepare_warp(pos, pos)
            om.emb.prepare_warp(pos, pos)
            sampled = [int(v) for v in rng.permutation(N_ENT)[:S]] if it != 1 else None
            id2idx = {v: k for k, v in enumerate(sampled)} if sampled else None
            mask = np.floor(rng.random((MB, DIM)) + 0.5)
            lg = model.step(None, users.tolist(), items.tolist(), None, sampled, id2idx, loss='mw',
                            masks=[torch.tensor(mask, dtype=torch.float32, device='cuda')])
            lo = om.step(users.tolist(), items.tolist(), sampled, id2idx, masks=[mask])
            assert abs(lg - lo) <= ltol * max(1.0, abs(lo)), (it, lg, lo)
        model.att_emb.check_plans()
        for k, want in om.emb.p.items():
            got = model.att_emb.params[k].cpu().numpy().astype(np.float64).reshape(want.shape)
            upd = np.abs(want - init[k]).max()
            assert upd > 0, k
            assert np.abs(got - want).max() <= utol * upd + 1e-7, (k, np.abs(got - want).max(), upd)
            acc_got = model.att_emb.accs[k].cpu().numpy().astype(np.float64).reshape(want.shape)
            acc_upd = np.abs(om.acc[k] - 0.1).max()
            assert np.abs(acc_got - om.acc[k]).max() <= 2 * utol * acc_upd + 3e-8, (k, 'acc')      # fp32(0.1) is 1.5e-9 off
    finally:
        _lib.exact_fp32 = False


@pytest.mark.parametrize('plan_agg', [3, 0])
def test_c2_row_sharded_g8_matches_unsharded(cuda, plan_agg):
    """SURVEY 8(e) at C2's shape on one device: 8 row-sharded copies (row t on shard t % 8, bags pre-partitioned by
    owner with `lengths_full` as the mean divisor).  Partial pooled vectors sum to the unsharded result; local
    plan + Adagrad applies reassemble to the unsharded tables."""
    from arecsys_b200._lib import POOL_MEAN, OPT_ADAGRAD
    G = 8
    ua, ia, i2l, l2i = _dataset(12)
    params = random_params(ua, ia, DIM, 5, scale=0.1)
    full = _embedding(ua, ia, l2i, params, plan_agg=plan_agg)
    parts = [_embedding(ua, ia, l2i, params, shard=(G, r), plan_agg=plan_agg) for r in range(G)]
    rng = np.random.default_rng(11)
    _, ids = _batch_ids(rng, MB)
    dout = torch.tensor(rng.standard_normal((MB, DIM)).astype(np.float32), device='cuda')
    dbias = torch.tensor(rng.standard_normal(MB).astype(np.float32), device='cuda')
    ref, refb, _ = full.pool('item', full._ids(ids), POOL_MEAN, True)
    acc, accb = torch.zeros_like(ref), torch.zeros_like(refb)
    for m in parts:
        o, b, _ = m.pool('item', m._ids(ids), POOL_MEAN, True)
        acc += o
        accb += b
    np.testing.assert_allclose(acc.cpu().numpy(), ref.cpu().numpy(), rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(accb.cpu().numpy(), refb.cpu().numpy(), rtol=2e-5, atol=2e-6)
    for m in [full] + parts:
        m.push_grad('item', m.sets['item'].attr_range(), m._ids(ids), POOL_MEAN, dout.clone(), dbias.clone())
        m.apply_gradients(0.3, OPT_ADAGRAD)
        m.check_plans()
    for name in full.sets['item'].names + [b for b in full.sets['item'].bias_names if b]:
        want = full.params[name].cpu().numpy()
        got = np.empty_like(want)
        for r, m in enumerate(parts):
            got[r::G] = m.params[name].cpu().numpy()
        np.testing.assert_allclose(got, want, rtol=1e-5, atol=1e-6, err_msg=name)
        assert np.abs(want - np.asarray(params[name]).reshape(want.shape)).max() > 0


def test_c3_lstm_training_step_matches_reference(cuda):
    """C3's shape: T = 50, mb = 512, d = H = 64, id + 2 multi-hot attributes per side, loss ce (fused scoring + CE
    over T*mb = 25 600 rows), clip_by_global_norm 5, Adagrad; catalog 2000 items so that the float64 autograd
    reference (oracle/torch_cpu_ref.py::TorchRefSeq) finishes in seconds."""
    import arecsys_b200  # noqa: F401
    from arecsys_b200.attributes.embed_attribute import EmbeddingAttribute
    from arecsys_b200.lstm.seqModel import SeqModel
    from arecsys_b200.utils import synthetic
    from oracle.torch_cpu_ref import TorchRefSeq
    dim, mb, T, n_users, n_items = 64, 512, 50, 3000, 2000
    ua, ia, _, l2i = synthetic.make_dataset(n_users, n_items, 2, 1000, 4, 16, 1.05, seed=2)
    ua.set_model_size(dim)
    ia.set_model_size(dim)
    params = random_params(ua, ia, dim, 3, scale=0.2)
    rng = np.random.default_rng(5)
    params['lstm_w'] = rng.uniform(-.15, .15, (2 * dim, 4 * dim)).astype(np.float32)
    params['lstm_b'] = np.zeros(4 * dim, dtype=np.float32)
    START = n_items
    l2i_d = {int(v): int(l2i[v]) for v in range(len(l2i))}
    i2l_d = {v: k for k, v in l2i_d.items()}
    i2l_d[START] = 0
    emb = EmbeddingAttribute(ua, ia, mb, None, T, False, i2l_d, l2i_d, params=params)
    model = SeqModel([T], dim, 1, 5.0, mb, 0.5, 0.83, emb, withAdagrad=True, dropoutRate=0.5, START_ID=START,
                     loss='ce', use_concat=False, no_user_id=False, topk_n=5, params=params)
    ref = TorchRefSeq(ua, ia, params, l2i_d, i2l_d, loss='ce', keep_prob=0.5, learning_rate=0.5, n_sampled=None,
                      dtype=torch.float64, size=dim, use_concat=False, no_user_id=False, max_gradient_norm=5.0,
                      item_output=False)
    init = {k: v.detach().clone().numpy() for k, v in ref.p.items()}
    for it in range(2):
        users = rng.integers(0, n_users, mb).tolist()
        lens = rng.integers(5, T + 1, mb)
        seqs = [rng.integers(0, n_items, int(n)).tolist() for n in lens]
        inp = [[START] + s[:-1] + [START] * (T - len(s)) for s in seqs]
        out = [s + [START] * (T - len(s)) for s in seqs]
        w = [[1.0] * len(s) + [0.0] * (T - len(s)) for s in seqs]
        tm = lambda l: [[l[j][i] for j in range(mb)] for i in range(T)]
        im = np.floor(rng.random((T, mb, dim)) + 0.5).astype(np.float32)
        omk = np.floor(rng.random((T, mb, dim)) + 0.5).astype(np.float32)
        lg = model.step(None, users, tm(inp), tm(out), tm(w), 0,
                        masks=(torch.tensor(im, device='cuda'), torch.tensor(omk, device='cuda')))
        lr_ = ref.step_seq(users, tm(inp), tm(out), tm(w), masks=(im, omk))
        assert abs(lg - lr_) <= 1e-3 * max(1.0, abs(lr_)), (it, lg, lr_)
        assert abs(float(model.last_gnorm) - ref.last_gnorm) <= 5e-3 * max(1.0, ref.last_gnorm)
    dense = model.dense_params()
    for k, v in ref.p.items():
        got = (emb.params[k] if k in emb.params else dense[k][0]).cpu().numpy()
        want = v.detach().numpy()
        upd = np.abs(want - init[k]).max()
        assert np.abs(got.reshape(want.shape) - want).max() <= 2e-2 * upd + 1e-6, (k, upd)


def test_c5_cbow_training_step_matches_reference(cuda):
    """C5's shape: CBOW, d = 128, ni = 3 context items, mb = 4096, loss mw over a 1024-item pool, separate output
    tables; 10^4-row tables."""
    import arecsys_b200  # noqa: F401
    from arecsys_b200.word2vec.cbow_model import Model
    from arecsys_b200.utils import synthetic
    from oracle.torch_cpu_ref import TorchRefCbow
    dim, mb, ni, n_users, n_items = 128, 4096, 3, N_ENT, N_ENT
    ua, ia, _, l2i = synthetic.make_dataset(n_users, n_items, 2, VOCAB, 4, 16, 1.05, seed=4)
    ua.set_model_size(dim)
    ia.set_model_size(dim)
    params = random_params(ua, ia, dim, 1, scale=0.1, item_output=True)
    l2i_d = {int(v): int(l2i[v]) for v in range(len(l2i))}
    i2l_d = {v: k for k, v in l2i_d.items()}
    model = Model(n_users, n_items, dim, mb, 0.5, 1.0, ua, ia, i2l_d, l2i_d, n_input_items=ni, loss_function='mw',
                  dropout=0.5, top_N_items=5, use_sep_item=True, n_sampled=S, params=params)
    ref = TorchRefCbow(ua, ia, params, l2i_d, i2l_d, loss='mw', keep_prob=0.5, learning_rate=0.5, n_sampled=S,
                       dtype=torch.float64, size=dim, item_output=True, ni=ni)
    init = {k: v.detach().clone().numpy() for k, v in ref.p.items()}
    rng = np.random.default_rng(3)
    for it in range(2):
        users = rng.integers(0, n_users, mb)
        outs = rng.integers(0, n_items, mb)
        ins = [rng.integers(0, n_items + 1, mb) for _ in range(ni)]
        pos = positives(users, outs, n_users, rng, n_items=n_items)
        model.prepare_warp(pos, pos)
        ref.pos, ref.pos_eval = pos, pos
        sampled = [int(v) for v in rng.permutation(n_items)[:S]] if it == 0 else None
        mask = np.floor(rng.random((mb, dim)) + 0.5).astype(np.float32)
        lg = model.step(None, users.tolist(), [x.tolist() for x in ins], outs.tolist(), sampled, None, loss='mw',
                        masks=[torch.tensor(mask, device='cuda')])
        lr_ = ref.step_cbow(users, ins, outs, item_sampled=sampled, mask=mask)
        assert abs(lg - lr_) <= 1e-3 * max(1.0, abs(lr_)), (it, lg, lr_)
    for k, v in ref.p.items():
        got = model.att_emb.params[k].cpu().numpy()
        want = v.detach().numpy()
        upd = np.abs(want - init[k]).max()
        assert np.abs(got.reshape(want.shape) - want).max() <= 2e-2 * upd + 1e-6, (k, upd)


def test_warp_eval_through_model_step_matches_reference_golden(cuda):
    """loss 'warp_eval' through LatentProductModel.step on the CUDA path vs what the reference's own graph reports
    (tests/golden/ref_hmfeval_warp_eval.npz): margin rank within 1e-3, TRUE rank bit-exact."""
    import os
    from test_ref_golden import GOLD, _attributes
    from arecsys_b200 import _lib
    from arecsys_b200.hmf.hmf_model import LatentProductModel
    d = np.load(os.path.join(GOLD, 'ref_hmfeval_warp_eval.npz'))
    dim, mb = int(d['dim']), int(d['mb'])
    ua, ia = _attributes(d, 'u_', dim), _attributes(d, 'i_', dim)
    l2i = d['l2i']
    ia.set_target_prediction_from_map(l2i)
    l2i_d = {int(v): int(l2i[v]) for v in range(len(l2i))}
    i2l_d = {int(l2i[v]): int(v) for v in range(len(l2i))}
    params = {k[len('init/'):]: d[k] for k in d.files if k.startswith('init/')}
    pu, ptr, it = d['pos_users'], d['pos_ptr'], d['pos_items']
    pos = {int(u): [int(v) for v in it[ptr[j]:ptr[j + 1]]] for j, u in enumerate(pu)}
    _lib.exact_fp32 = True               # true ranks compare scores: exact-fp32 contraction
    try:
        model = LatentProductModel(ua.num_entities, ia.num_entities, dim, 1, mb, 0.3, 1.0, ua, ia, i2l_d, l2i_d,
                                   loss_function='warp_eval', dropout=0.5, params=params)
        model.prepare_warp({}, pos)
        margin, rank = model.step(None, d['users'].tolist(), d['items'].tolist(), forward_only=True, loss='warp_eval')
    finally:
        _lib.exact_fp32 = False
    np.testing.assert_allclose(margin, d['margin_rank'], rtol=1e-3, atol=1e-4)
    assert np.array_equal(np.asarray(rank), d['true_rank'])
