"""CPU tests that pin the oracle: the hand-checked known-answer vectors of SURVEY.md 8(c),
the algebraic identities the CUDA path relies on, and an independent autograd derivation of
every hand-derived gradient (oracle/np_oracle.py vs oracle/torch_cpu_ref.py)."""
import numpy as np
import pytest
import torch

from helpers import kat_item_attributes, small_dataset, random_params, positives
from oracle import np_oracle as O
from oracle.torch_cpu_ref import TorchRefHMF

E = np.array([[1, 2], [3, 4], [5, 6], [7, 8], [9, 10]], dtype=np.float64)
BETA = np.array([[.1], [.2], [.3], [.4], [.5]])


def kat_emb():
    ia = kat_item_attributes()
    params = {'userembed_mulhot_0': E, 'itemembed_mulhot_0': E, 'item_bias_mulhot_0': BETA}
    return O.OracleEmbeddingAttribute(ia, ia, 2, None, params, logit_ind2item_ind={0: 0, 1: 1, 2: 2},
                                      item_ind2logit_ind={0: 0, 1: 1, 2: 2})


def test_kat1_pool_indices_and_values():
    emb = kat_emb()
    ids = [2, 0, 1, 3]
    idx, seg = emb.flat_indices(emb.item_attributes, 0, ids)
    assert idx.tolist() == [3, 4, 4, 0, 2, 1, 1]           # bit-exact
    assert seg.tolist() == [0, 0, 0, 1, 1, 2, 3]
    c, m, b = emb._get_embedded([], [E], None, None, ids, emb.item_attributes, concatenation=False)
    np.testing.assert_allclose(m[0], [[25 / 3, 28 / 3], [3, 4], [3, 4], [3, 4]], rtol=1e-15)


def test_kat2_scoring_and_ce():
    emb = kat_emb()
    U = np.array([[1, 0], [.5, -1]])
    logits = emb.get_prediction(U)
    np.testing.assert_allclose(logits, [[3.2, 3.2, 8.8], [-2.3, -2.3, -4.7]], rtol=1e-12)
    ce = emb.compute_loss(logits, np.array([2, 0]), 'ce')
    np.testing.assert_allclose(ce, [0.0073685, 0.7375075], atol=5e-8)
    assert abs(ce.mean() - 0.3724380) < 5e-8


def test_kat3_wmrb():
    scores = np.array([[2.0, 0.5, 1.5, 3.0]])
    mask = np.array([[True, False, True, False]])          # positives {1, 3}
    l = O.compute_loss(scores, np.array([1]), 'warp', mask)
    assert abs(l[0] - np.log(5.5)) < 1e-14 and abs(l[0] - 1.7047482) < 2e-7   # SURVEY rounds log 5.5 up
    mr, tr = O.compute_loss(scores, np.array([1]), 'warp_eval', mask)
    assert abs(mr[0] - 4.5) < 1e-12 and tr[0] == 2


def test_kat4_mw():
    scores = np.array([[0.2, 1.4, -0.3]])
    mask = np.array([[True, True, False]])
    l = O.compute_loss(scores, np.array([0.9]), 'mw', mask)
    assert abs(l[0] - np.log(2.8)) < 1e-14 and abs(l[0] - 1.0296195) < 2e-7


def test_kat5_adagrad():
    th, acc = O.adagrad_update(np.zeros(2), np.full(2, 0.1), np.array([0.5, -2.0]), 1.0)
    np.testing.assert_allclose(th, [-0.8451542, 0.9877296], atol=5e-8)
    np.testing.assert_allclose(acc, [0.35, 4.1])


def test_start_pseudo_entity_and_lstm_cell():
    ia = kat_item_attributes()
    assert ia.features_mulhot[0][ia.mulhot_starts[0][3]] == O.START_ID and ia.mulhot_lengths[0][3] == 1
    x = np.array([[0.5, -1.0]]); h = np.zeros((1, 1)); c = np.zeros((1, 1))
    W = np.arange(12, dtype=np.float64).reshape(3, 4) / 10; b = np.zeros(4)
    h2, c2 = O.lstm_cell(x, h, c, W, b)
    z = x @ W[:2]
    i, j, f, o = z[0]
    cc = O.sigmoid(i) * np.tanh(j)
    np.testing.assert_allclose(c2, [[cc]]); np.testing.assert_allclose(h2, [[O.sigmoid(o) * np.tanh(cc)]])


def _setup(loss, nonlinear='linear', n_sampled=None, seed=0, dtype=np.float64):
    dim = 6
    ua, ia, i2l, l2i = small_dataset(n_users=30, n_items=25, n_mulhot=2, vocab_m=12, dim=dim, seed=seed)
    params = random_params(ua, ia, dim, seed=seed + 1, mlp_hidden=5 if nonlinear != 'linear' else None)
    rng = np.random.default_rng(seed + 2)
    mb = 7
    users = rng.integers(0, 30, mb)
    items = rng.integers(0, 25, mb)
    pos = positives(users, items, 30, rng, n_items=25)
    l2i_d = {int(v): int(l2i[v]) for v in range(len(l2i))}
    i2l_d = {int(l2i[v]): int(v) for v in range(len(l2i))}
    emb = O.OracleEmbeddingAttribute(ua, ia, mb, n_sampled, params, item_ind2logit_ind=i2l_d,
                                     logit_ind2item_ind=l2i_d, dtype=dtype)
    emb.prepare_warp(pos, pos)
    model = O.OracleHMF(emb, loss=loss, nonlinear=nonlinear, keep_prob=0.5, learning_rate=0.3)
    ref = TorchRefHMF(ua, ia, params, l2i_d, i2l_d, loss=loss, nonlinear=nonlinear, keep_prob=0.5,
                      learning_rate=0.3, n_sampled=n_sampled, dtype=torch.float64)
    ref.pos, ref.pos_eval = pos, pos
    nm = 3 if nonlinear != 'linear' else 1
    shapes = [(mb, dim), (mb, 5), (mb, dim)] if nm == 3 else [(mb, dim)]
    masks = [np.floor(rng.random(s) + 0.5) for s in shapes]
    sampled = list(rng.permutation(25)[:n_sampled]) if n_sampled else None
    return model, ref, users, items, masks, sampled, rng


@pytest.mark.parametrize('loss,nonlinear', [('ce', 'linear'), ('warp', 'linear'), ('rs', 'linear'),
                                            ('rs-sig', 'linear'), ('rs-sig2', 'linear'), ('bbpr', 'linear'),
                                            ('ce', 'relu'), ('warp', 'tanh'), ('mw', 'linear'), ('mw', 'tanh')])
def test_hand_gradients_match_autograd_over_training_steps(loss, nonlinear):
    ns = 9 if loss == 'mw' else None
    model, ref, users, items, masks, sampled, rng = _setup(loss, nonlinear, ns)
    for it in range(3):
        id2idx = {int(v): k for k, v in enumerate(sampled)} if sampled else None
        l1 = model.step(list(users), list(items), item_sampled=sampled, item_sampled_id2idx=id2idx, masks=masks)
        l2 = ref.step(list(users), list(items), item_sampled=sampled, masks=masks)
        assert abs(l1 - l2) <= 1e-10 * max(1, abs(l2)), (it, l1, l2)
        for k in model.emb.p:
            np.testing.assert_allclose(model.emb.p[k], ref.p[k].detach().numpy(), rtol=1e-9, atol=1e-12,
                                       err_msg='%s after step %d' % (k, it))
        users = rng.integers(0, 30, len(users)); items = rng.integers(0, 25, len(items))
        pos = positives(users, items, 30, rng, n_items=25)
        model.emb.prepare_warp(pos, pos); ref.pos, ref.pos_eval = pos, pos
    e1 = model.step(list(users), list(items), forward_only=True)
    e2 = ref.step(list(users), list(items), forward_only=True)
    assert abs(e1 - e2) < 1e-10 * max(1, abs(e2))


@pytest.mark.parametrize('lf', ['log', 'exp', 'poly', 'poly2', 'linear', 'square'])
def test_rs_transforms_match_autograd(lf):
    model, ref, users, items, masks, _, _ = _setup('rs')
    model.loss_func = ref.loss_func = lf
    model.loss_exp_p = ref.exp_p = 1.3
    l1 = model.step(list(users), list(items), masks=masks)
    l2 = ref.step(list(users), list(items), masks=masks)
    assert abs(l1 - l2) < 1e-10 * max(1, abs(l2))
    for k in model.emb.p:
        np.testing.assert_allclose(model.emb.p[k], ref.p[k].detach().numpy(), rtol=1e-9, atol=1e-12)


def test_linearity_target_score_equals_prediction_column():
    model, _, users, items, _, _, _ = _setup('ce')
    e = model.emb
    u, _ = e.get_batch_user(users, 1.0, False)
    logits = e.get_prediction(u)
    ts = e.get_target_score(u, items)
    cols = [e.item_ind2logit_ind[int(v)] for v in items]
    np.testing.assert_allclose(ts, logits[np.arange(len(items)), cols], rtol=1e-12)


def test_pool_then_gemm_equals_gemm_then_pool():
    model, _, users, items, _, _, _ = _setup('ce')
    e = model.emb
    ia = e.item_attributes
    u, _ = e.get_batch_user(users, 1.0, False)
    cat, mul = e._tables('item', ia); bc, bm = e._biases('item', ia)
    ids = [e.logit_ind2item_ind[v] for v in range(e.logit_size)]
    P, bp = O.pool_entities(e, cat, mul, bc, bm, ia, ids)
    np.testing.assert_allclose(u @ P.T + bp, e.get_prediction(u), rtol=1e-12)


def test_bag_permutation_invariance_and_segment_ids_sorted():
    ua, ia, i2l, l2i = small_dataset()
    for f in range(ia.num_features_mulhot):
        seg = np.asarray(ia.full_segids_tr[f])
        assert (np.diff(seg) >= 0).all() and seg[0] == 0 and seg[-1] == len(l2i) - 1
        assert (np.bincount(seg) == np.asarray(ia.full_lengths_tr[f]).ravel()).all()


def test_topk_ties_lower_index_first():
    model, _, users, _, _, _, _ = _setup('ce')
    for k in model.emb.p:
        if 'item' in k:
            model.emb.p[k][:] = 0.0        # all logits equal -> order must be 0,1,2,...
    idx, _ = model.top_k(users, 5)
    assert (idx == np.arange(5)).all()
