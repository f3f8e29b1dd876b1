"""Fused full-catalog scoring + softmax cross-entropy (arx_ce_fwd / arx_ce_bwd, tcgen05) against a float64
restatement of embed_attribute.py:148-206 + :530 on the same tf32-rounded operands."""
import numpy as np
import pytest
import torch

import arecsys_b200  # noqa: F401
from arecsys_b200 import _lib

pytestmark = pytest.mark.gpu


def _rounded(x):
    out = torch.empty_like(x)
    _lib.call('arx_round_tf32', x.data_ptr(), out.data_ptr(), x.numel())
    return out


@pytest.mark.parametrize('M,N,d', [(128, 64, 32), (4, 8, 32), (200, 1000, 64), (132, 4100, 96), (4096, 5000, 128),
                                   (512, 100000, 64)])
def test_fused_ce_forward_and_backward(cuda, M, N, d):
    g = torch.Generator(device='cpu').manual_seed(M + N + d)
    U = _rounded((torch.randn(M, d, generator=g) * 0.7).cuda())
    P = _rounded((torch.randn(N, d, generator=g) * 0.7).cuda())
    beta = (torch.randn(N, generator=g) * 0.5).cuda()
    tgt = torch.randint(0, N, (M,), generator=g).to(torch.int32).cuda()
    gr = (torch.rand(M, generator=g) / M).cuda()
    assert _lib.ce_supported(M, N, d)
    lse = _lib.ce_fwd(U, P, beta, M, N, d)
    assert lse is not None
    Ud, Pd, bd = U.double(), P.double(), beta.double()
    ref_lse = torch.empty(M, dtype=torch.float64, device='cuda')
    dU = torch.zeros(M, d, dtype=torch.float64, device='cuda')
    dP = torch.zeros(N, d, dtype=torch.float64, device='cuda')
    db = torch.zeros(N, dtype=torch.float64, device='cuda')
    step = max(1, (1 << 24) // N)
    for a in range(0, M, step):                                # row blocks: [M, N] float64 does not fit at once
        b = min(M, a + step)
        logits = Ud[a:b] @ Pd.T + bd
        ref_lse[a:b] = torch.logsumexp(logits, 1)
        D = torch.softmax(logits, 1)
        D[torch.arange(b - a, device='cuda'), tgt[a:b].long()] -= 1.0
        D *= gr[a:b].double()[:, None]
        dU[a:b] = D @ Pd
        dP += D.T @ Ud[a:b]
        db += D.sum(0)
    np.testing.assert_allclose(lse.cpu().numpy(), ref_lse.cpu().numpy(), rtol=2e-6, atol=2e-6)
    out = _lib.ce_bwd(U, P, beta, lse, gr, tgt, M, N, d)
    assert out is not None
    for got, want, name in zip(out, (dU, dP, db), ('dU', 'dP', 'dbeta')):
        w = want.cpu().numpy()
        err = np.abs(got.cpu().numpy() - w).max()
        assert err <= 2e-3 * np.abs(w).max() + 1e-9, (name, err, np.abs(w).max())


@pytest.mark.parametrize('M,N,d', [(128, 64, 32), (16, 12, 128), (200, 1000, 64), (4096, 1024, 128), (132, 4100, 96)])
def test_fused_mw_forward_and_backward(cuda, M, N, d):
    """arx_mw_mask_build / arx_mw_fwd / arx_mw_bwd against a float64 restatement of embed_attribute.py:641-649
    (sampled WMRB) and its autograd gradients on the same tf32-rounded operands."""
    g = torch.Generator(device='cpu').manual_seed(M * 3 + N + d)
    U = _rounded((torch.randn(M, d, generator=g) * 0.7).cuda())
    P = _rounded((torch.randn(N, d, generator=g) * 0.7).cuda())
    beta = (torch.randn(N, generator=g) * 0.5).cuda()
    ts = (torch.randn(M, generator=g) * 2.0).cuda()
    gr = (torch.rand(M, generator=g) / M).cuda()
    # per-row excluded columns as a CSR with some -1 (positives outside the pool)
    n_users = M + 5
    lens = torch.randint(0, 6, (n_users,), generator=g)
    ptr_ = torch.zeros(n_users + 1, dtype=torch.int64); ptr_[1:] = torch.cumsum(lens, 0)
    idx = torch.randint(-1, N, (int(ptr_[-1]),), generator=g)
    rows = torch.randperm(n_users, generator=g)[:M]
    excl = torch.zeros(M, N, dtype=torch.bool)
    for b in range(M):
        for p_ in range(int(ptr_[rows[b]]), int(ptr_[rows[b] + 1])):
            if idx[p_] >= 0:
                excl[b, idx[p_]] = True
    ld = _lib.mw_mask_words(N)
    mask = torch.empty((M, ld), dtype=torch.int32, device='cuda')
    rows_d, ptr_d, idx_d = rows.to(torch.int32).cuda(), ptr_.to(torch.int32).cuda(), idx.to(torch.int32).cuda()
    _lib.call('arx_mw_mask_build', rows_d.data_ptr(), ptr_d.data_ptr(), idx_d.data_ptr(), M, N, mask.data_ptr(), ld)
    fw = _lib.mw_fwd(U, P, beta, ts, mask, ld, M, N, d)
    assert fw is not None
    hsum, loss = fw
    Ud = U.double().requires_grad_(True); Pd = P.double().requires_grad_(True); bd = beta.double().requires_grad_(True)
    td = ts.double().requires_grad_(True)
    S = Ud @ Pd.T + bd
    hinge = torch.relu(1.0 + S - td[:, None]) * (~excl).cuda().double()
    ref_loss = torch.log(1.0 + hinge.sum(1))
    np.testing.assert_allclose(loss.cpu().numpy(), ref_loss.detach().cpu().numpy(), rtol=2e-5, atol=2e-6)
    (ref_loss * gr.double()).sum().backward()
    out = _lib.mw_bwd(U, P, beta, ts, mask, ld, hsum, gr, M, N, d)
    assert out is not None
    for got, want, name in zip(out, (Ud.grad, Pd.grad, bd.grad, td.grad), ('dU', 'dP', 'dbeta', 'dts')):
        w = want.cpu().numpy()
        err = np.abs(got.cpu().numpy() - w).max()
        assert err <= 2e-3 * np.abs(w).max() + 1e-9, (name, err, np.abs(w).max())
