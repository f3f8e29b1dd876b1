"""GPU parity of the non-linear catalog pooling (K3m, attributes/embed_attribute.py:194-200, output_feat 2 / 3):
arx_score_max / arx_token_pool_fwd / arx_token_pool_bwd / arx_rowsum against a float64 autograd restatement of the
reference's literal order (segment_max; score_max + log(1 + segment_sum(exp(s - score_max)))), and get_prediction /
token_prediction_backward of the host mirror against oracle/torch_cpu_ref.py on a small attribute store."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _csr(rng, V, Vf, max_len):
    lens = rng.integers(1, max_len + 1, size=V)
    ptr = np.zeros(V + 1, dtype=np.int64)
    ptr[1:] = np.cumsum(lens)
    vals = rng.integers(0, Vf, size=int(ptr[-1])).astype(np.int32)
    return vals, ptr


def _ref_pool(S, bias, vals, ptr, mode):
    """float64 restatement: S [Vf, mb] -> [V, mb]"""
    s = S + bias.reshape(-1, 1)
    V = len(ptr) - 1
    seg = torch.tensor(np.repeat(np.arange(V), np.diff(ptr)), dtype=torch.int64)
    rows = s.index_select(0, torch.tensor(vals, dtype=torch.int64))
    if mode == 1:
        return torch.zeros((V, S.shape[1]), dtype=S.dtype).index_add(0, seg, rows)
    if mode == 2:
        out = torch.full((V, S.shape[1]), float('-inf'), dtype=S.dtype)
        return out.scatter_reduce(0, seg.reshape(-1, 1).expand_as(rows), rows, 'amax', include_self=True)
    m = s.max()
    return m + torch.log(1 + torch.zeros((V, S.shape[1]), dtype=S.dtype).index_add(0, seg, torch.exp(rows - m)))


@pytest.mark.parametrize('mode', [1, 2, 3])
@pytest.mark.parametrize('V,Vf,mb,max_len', [(7, 5, 3, 4), (300, 40, 33, 9), (1000, 500, 260, 64), (50, 2000, 512, 1)])
def test_token_pool_fwd_bwd_match_float64(cuda, mode, V, Vf, mb, max_len):
    from arecsys_b200._lib import call, ptr as p_
    rng = np.random.default_rng(V * 7 + mode)
    vals, ptr = _csr(rng, V, Vf, max_len)
    S = rng.standard_normal((Vf, mb)).astype(np.float32)
    bias = (rng.standard_normal(Vf) * 0.3).astype(np.float32)
    dOut = rng.standard_normal((V, mb)).astype(np.float32)
    scale = 0.25

    S64 = torch.tensor(S, dtype=torch.float64, requires_grad=True)
    b64 = torch.tensor(bias, dtype=torch.float64, requires_grad=True)
    want = _ref_pool(S64, b64, vals, ptr, mode) * scale
    (want * torch.tensor(dOut, dtype=torch.float64)).sum().backward()

    dev = lambda a: torch.tensor(a, device=cuda)
    S_d, b_d, v_d, p_d, g_d = dev(S), dev(bias), dev(vals), dev(ptr), dev(dOut)
    out = torch.full((V, mb), 0.5, dtype=torch.float32, device=cuda)             # the kernel ACCUMULATES into out
    arg = torch.empty((V, mb), dtype=torch.int32, device=cuda) if mode == 2 else None
    den = torch.empty((V, mb), dtype=torch.float32, device=cuda) if mode == 3 else None
    pk = None
    if mode == 3:
        pk = torch.zeros(1, dtype=torch.int64, device=cuda)
        call('arx_score_max', S_d.data_ptr(), b_d.data_ptr(), Vf, mb, pk.data_ptr())
    call('arx_token_pool_fwd', S_d.data_ptr(), b_d.data_ptr(), mb, v_d.data_ptr(), p_d.data_ptr(), V, mode, p_(pk), scale,
         out.data_ptr(), p_(arg), p_(den))
    got = out.cpu().double() - 0.5
    np.testing.assert_allclose(got.numpy(), want.detach().numpy(), rtol=2e-5, atol=2e-5)

    dS = torch.zeros((Vf, mb), dtype=torch.float32, device=cuda)
    scratch = torch.zeros(1, dtype=torch.float32, device=cuda)
    call('arx_token_pool_bwd', g_d.data_ptr(), S_d.data_ptr(), b_d.data_ptr(), mb, v_d.data_ptr(), p_d.data_ptr(), V, mode,
         p_(pk), scale, p_(arg), p_(den), dS.data_ptr(), scratch.data_ptr())
    np.testing.assert_allclose(dS.cpu().numpy(), S64.grad.numpy(), rtol=2e-4, atol=2e-5)
    db = torch.zeros(Vf, dtype=torch.float32, device=cuda)
    call('arx_rowsum', dS.data_ptr(), Vf, mb, db.data_ptr(), 0)
    np.testing.assert_allclose(db.cpu().numpy(), b64.grad.numpy(), rtol=2e-4, atol=5e-5)
    call('arx_rowsum', dS.data_ptr(), Vf, mb, db.data_ptr(), 1)                  # accumulate flag
    np.testing.assert_allclose(db.cpu().numpy(), 2 * b64.grad.numpy(), rtol=2e-4, atol=1e-4)


def test_categorical_attribute_is_a_bag_of_one(cuda):
    """ptr = NULL: item v holds the single token values[v] (embed_attribute.py:172)."""
    from arecsys_b200._lib import call
    rng = np.random.default_rng(3)
    V, Vf, mb = 90, 17, 40
    vals = rng.integers(0, Vf, size=V).astype(np.int32)
    S = rng.standard_normal((Vf, mb)).astype(np.float32)
    bias = rng.standard_normal(Vf).astype(np.float32)
    dev = lambda a: torch.tensor(a, device=cuda)
    S_d, b_d, v_d = dev(S), dev(bias), dev(vals)
    out = torch.zeros((V, mb), dtype=torch.float32, device=cuda)
    call('arx_token_pool_fwd', S_d.data_ptr(), b_d.data_ptr(), mb, v_d.data_ptr(), None, V, 1, None, 1.0, out.data_ptr(),
         None, None)
    np.testing.assert_allclose(out.cpu().numpy(), (S + bias[:, None])[vals], rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize('output_feat', [2, 3])
@pytest.mark.parametrize('sep', [False, True])
def test_get_prediction_nonlinear_pooling_matches_oracle(cuda, output_feat, sep):
    """EmbeddingAttribute.get_prediction(output_feat = 2 / 3) and its adjoint against the oracle's literal
    get_prediction (torch_cpu_ref.py::_pred) on the small heterogeneous store of tests/helpers.py."""
    from helpers import small_dataset, random_params
    from arecsys_b200 import _lib
    from arecsys_b200.attributes.embed_attribute import EmbeddingAttribute
    from oracle.torch_cpu_ref import TorchRefSeq
    ua, ia, i2l, l2i = small_dataset(seed=11, dim=16)
    dim, mb = 16, 12
    params = random_params(ua, ia, dim, seed=5, item_output=sep)
    rng = np.random.default_rng(9)
    u = rng.standard_normal((mb, dim)).astype(np.float32)
    V = len(l2i)
    dlog = rng.standard_normal((mb, V)).astype(np.float32)

    ref = TorchRefSeq(ua, ia, {k: v.copy() for k, v in params.items()}, l2i, i2l, loss='ce', keep_prob=1.0, learning_rate=0.1,
                      n_sampled=None, dtype=torch.float64, size=dim, item_output=sep, output_feat=output_feat)
    u64 = torch.tensor(u, dtype=torch.float64, requires_grad=True)
    for v in ref.p.values():
        v.grad = None
        v.requires_grad_(True)
    want = ref._pred(u64, 'full')
    (want * torch.tensor(dlog, dtype=torch.float64)).sum().backward()

    _lib.exact_fp32 = True
    try:
        emb = EmbeddingAttribute(ua, ia, mb, None, 0, sep, i2l, l2i, params={k: v.copy() for k, v in params.items()})
        got = emb.get_prediction(torch.tensor(u, device=cuda), output_feat=output_feat)
        np.testing.assert_allclose(got.cpu().numpy(), want.detach().numpy(), rtol=2e-5, atol=2e-5)
        du = emb.token_prediction_backward(torch.tensor(dlog, device=cuda))
        np.testing.assert_allclose(du.cpu().numpy(), u64.grad.numpy(), rtol=2e-4, atol=2e-5)
        pre = emb._out_prefix()
        for name, g in emb.dense_table_grads[pre].items():
            w = ref.p[name].grad
            assert w is not None, name
            np.testing.assert_allclose(g.cpu().numpy().reshape(-1), w.numpy().reshape(-1), rtol=2e-4, atol=2e-5, err_msg=name)
    finally:
        _lib.exact_fp32 = False
