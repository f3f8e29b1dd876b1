"""GPU parity of the full-catalog WMRB on the tensor-core pipeline (EmbeddingAttribute.fused_warp: 'warp' and the hinge
members of the rs family, attributes/embed_attribute.py:551-618 of the reference) against a float64 autograd restatement
of the reference's order (get_prediction -> mask -> relu(1 + s - s_t) -> sum -> transform), and of the blocked
scoring + top-k (EmbeddingAttribute.score_topk, hmf/hmf_model.py:154) against arx_topk_rows over the materialised scores."""
import numpy as np
import pytest
import torch

from helpers import small_dataset, random_params, positives

pytestmark = pytest.mark.gpu


def _setup(cuda, dim=64, nu=200, ni=3000, mb=64, seed=0):
    from arecsys_b200.attributes.embed_attribute import EmbeddingAttribute
    ua, ia, i2l, l2i = small_dataset(nu, ni, 3, 80, 4, 8, seed, None, dim)
    params = random_params(ua, ia, dim, seed + 1, scale=0.3)
    emb = EmbeddingAttribute(ua, ia, mb, None, 0, False, i2l, l2i, params=params)
    rng = np.random.default_rng(seed + 2)
    users = rng.integers(0, nu, mb)
    items = rng.integers(0, ni, mb)
    pos = positives(users, items, nu, rng, extra=20, n_items=ni)
    emb.prepare_warp(pos, pos)
    emb.add_input({}, users.tolist(), items.tolist(), loss='warp')
    latent = torch.tensor(rng.standard_normal((mb, dim)).astype(np.float32) * 0.5, device=cuda)
    targets = emb.item2logit_dev[emb._ids(items).long()].contiguous()
    return emb, latent, targets, users, pos, rng


def _transform(S, loss, lf, p):
    if loss == 'warp' or lf == 'log':
        return torch.log1p(S)
    if lf == 'exp':
        return 1.0 - torch.pow(torch.tensor(p, dtype=S.dtype), -S)
    if lf == 'poly':
        return torch.pow(S, p)
    if lf == 'poly2':
        return torch.pow(1.0 + S, p)
    if lf == 'linear':
        return S
    return S * S


def _reference(emb, latent, targets, users, pos, loss, lf, p, scale, unmasked):
    P, beta, _ = emb.pool_catalog('full', 1)
    U = latent.double().cpu().requires_grad_(True)
    P64 = P.double().cpu().requires_grad_(True)
    b64 = beta.double().cpu().requires_grad_(True)
    s = U @ P64.t() + b64
    t = targets.long().cpu()
    st = s.gather(1, t.unsqueeze(1))
    keep = torch.ones_like(s)
    if not unmasked:
        l2i = emb.item2logit_dev.cpu().numpy() if hasattr(emb, 'item2logit_dev') else None
        for r, u in enumerate(users):
            cols = [int(l2i[i]) for i in pos[int(u)] if l2i[i] >= 0]
            keep[r, cols] = 0.0
    S = (torch.relu(1.0 + s - st) * keep).sum(1)
    l = _transform(S, loss, lf, p)
    (l * scale.double().cpu()).sum().backward()
    return l.detach(), U.grad, P64.grad, b64.grad


@pytest.mark.parametrize('loss,lf', [('warp', 'log'), ('rs', 'log'), ('rs', 'exp'), ('rs', 'poly'), ('rs', 'poly2'),
                                     ('rs', 'linear'), ('rs', 'square')])
@pytest.mark.parametrize('unmasked,blocks', [(False, 1), (True, 1), (False, 4)])
def test_fused_warp_matches_float64(cuda, loss, lf, unmasked, blocks):
    emb, latent, targets, users, pos, rng = _setup(cuda)
    mb = latent.shape[0]
    p = 1.005 if lf == 'exp' else 1.3
    scale = torch.tensor(rng.random(mb).astype(np.float32) / mb, device=cuda)
    want_l, want_dU, want_dP, want_db = _reference(emb, latent, targets, users, pos, loss, lf, p, scale, unmasked)
    ld_bytes = 4 * ((3000 + 127) // 128 * 4)
    got = emb.fused_warp(latent, targets, loss, lf, p, scale, True, unmasked=unmasked,
                         max_mask_bytes=(mb // blocks) * ld_bytes)
    assert got is not None, 'shape must be covered by the tensor-core path'
    l, (dU, dP, db) = got
    tol = lambda w: 4e-3 * max(1e-6, float(w.abs().max()))
    np.testing.assert_allclose(l.cpu().double().numpy(), want_l.numpy(), rtol=3e-3, atol=3e-3)
    assert float((dU.cpu().double() - want_dU).abs().max()) <= tol(want_dU)
    assert float((dP.cpu().double() - want_dP).abs().max()) <= tol(want_dP)
    assert float((db.cpu().double() - want_db).abs().max()) <= tol(want_db)


def test_fused_warp_declines_what_it_does_not_cover(cuda):
    emb, latent, targets, users, pos, rng = _setup(cuda)
    assert emb.fused_warp(latent, targets, 'rs-sig') is None
    assert emb.fused_warp(latent, targets, 'bbpr') is None


@pytest.mark.parametrize('exact', [True, False])
@pytest.mark.parametrize('k,chunk', [(10, 1024), (100, 512), (7, 4096), (300, 256)])
def test_blocked_score_topk_equals_topk_of_the_full_scores(cuda, exact, k, chunk):
    from arecsys_b200 import _lib
    from arecsys_b200._lib import call
    emb, latent, targets, users, pos, rng = _setup(cuda, seed=3)
    _lib.exact_fp32 = exact
    try:
        logits = emb.get_prediction(latent)
        mb, V = logits.shape
        idx_w = torch.empty((mb, k), dtype=torch.int32, device=cuda)
        val_w = torch.empty((mb, k), dtype=torch.float32, device=cuda)
        call('arx_topk_rows', logits.data_ptr(), mb, V, logits.stride(0), k, idx_w.data_ptr(), val_w.data_ptr())
        idx, val, lse = emb.score_topk(latent, k, want_lse=True, chunk=chunk)
        assert torch.equal(idx, idx_w)
        assert torch.equal(val, val_w)
        np.testing.assert_allclose(lse.cpu().numpy(), torch.logsumexp(logits, 1).cpu().numpy(), rtol=1e-5, atol=1e-5)
    finally:
        _lib.exact_fp32 = False


def test_blocked_score_topk_keeps_tie_order(cuda):
    """Duplicate items (identical attribute bags and a shared id row) score identically: ties must come back lower
    index first across block boundaries, as tf.nn.top_k does."""
    from arecsys_b200._lib import call
    emb, latent, targets, users, pos, rng = _setup(cuda, seed=4)
    P, beta, _ = emb.pool_catalog('full', 1)
    P[1500:3000] = P[0:1500]               # make the second half of the catalog a copy of the first
    beta[1500:3000] = beta[0:1500]
    emb.pool_catalog = lambda pool='full', output_feat=1: (P, beta, emb.catalog_ids)
    logits = emb.get_prediction(latent)
    k = 20
    idx_w = torch.empty((latent.shape[0], k), dtype=torch.int32, device=cuda)
    call('arx_topk_rows', logits.data_ptr(), logits.shape[0], logits.shape[1], logits.stride(0), k, idx_w.data_ptr(), None)
    idx, _, _ = emb.score_topk(latent, k, chunk=512)
    assert torch.equal(idx, idx_w)


@pytest.mark.parametrize('opt_name', ['adagrad', 'sgd'])
def test_catalog_gradient_in_column_slabs_equals_the_row_kernel(cuda, opt_name, monkeypatch):
    """arx_pool_bwd_apply_slab (+ arx_bwd_plan_alloc_h): the de-duplicated optimizer step for the gradient of the whole
    pooled catalog, 16 columns at a time from an L2-resident slab, against the row-at-a-time kernel on the same dP —
    hot rows (chunked), rows with one contribution (id table), biases, and a clip scale."""
    from arecsys_b200 import _lib
    from arecsys_b200._lib import POOL_MEAN, OPT_ADAGRAD, OPT_SGD
    from arecsys_b200.attributes.embed_attribute import EmbeddingAttribute
    dim, ni = 128, 70000
    ua, ia, i2l, l2i = small_dataset(50, ni, 3, 3000, 6, 12, 1, None, dim)
    params = random_params(ua, ia, dim, 2, scale=0.3)
    rng = np.random.default_rng(0)
    dP = torch.tensor(rng.standard_normal((ni, dim)).astype(np.float32) * 1e-2, device=cuda)
    db = torch.tensor(rng.standard_normal(ni).astype(np.float32) * 1e-2, device=cuda)
    scale = torch.tensor([0.37], dtype=torch.float32, device=cuda)
    opt = OPT_ADAGRAD if opt_name == 'adagrad' else OPT_SGD
    out = {}
    for slabs in ('0', '1'):
        monkeypatch.setenv('ARX_CATALOG_SLABS', slabs)
        emb = EmbeddingAttribute(ua, ia, 16, None, 0, False, i2l, l2i, params={k: v.copy() for k, v in params.items()})
        pre = emb._out_prefix()
        for _ in range(2):                                   # twice: the accumulators of the first step feed the second
            emb.push_grad(pre, emb.sets[pre].attr_range(), emb.catalog_ids, POOL_MEAN, dP, db, plan_key='catalog')
            emb.apply_gradients(0.3, opt, grad_scale=scale)
        emb.check_plans()
        assert (getattr(emb.sets[pre], '_slab_plans', None) is not None) == (slabs == '1')
        out[slabs] = {k: v.clone() for k, v in emb.params.items() if k.startswith('item')}
        out[slabs].update({'acc/' + k: v.clone() for k, v in emb.accs.items() if k.startswith('item')})
    for k in out['0']:
        a, b = out['0'][k], out['1'][k]
        assert float((a - b).abs().max()) <= 2e-5 * max(1.0, float(a.abs().max())), k
