"""CPU unit tests of host-side pieces that sit between the C ABI and the models (no kernel is called): gradient-arena
adjacency, the transforms of the WMRB hinge sum (attributes/embed_attribute.py:580-594 of the reference), the canonical
bucket order of the deterministic scatter mode, shard-aware checkpoint paths."""
import types

import numpy as np
import pytest
import torch

import arecsys_b200  # noqa: F401
from arecsys_b200.attributes.embed_attribute import EmbeddingAttribute as E


def test_adjacent_row_blocks_are_used_in_place():
    a = torch.arange(40.).view(10, 4)
    r = E._adjacent([a[:3], a[3:7], a[7:]])
    assert r.shape == (10, 4) and r.data_ptr() == a.data_ptr()
    assert E._adjacent([a[2:5], a[5:9]]).shape == (7, 4)
    b = torch.arange(10.)
    assert E._adjacent([b[:4], b[4:]]).shape == (10,)
    assert E._adjacent([a[:3], a[4:]]) is None                      # gap
    assert E._adjacent([a[:3], torch.zeros(2, 4)]) is None          # another allocation
    assert E._adjacent([a[:3, :2], a[3:, :2]]) is None              # not contiguous


@pytest.mark.parametrize('loss,lf,p', [('warp', 'square', 1.3), ('rs', 'log', 1.3), ('rs', 'exp', 1.005), ('rs', 'poly', 1.3),
                                       ('rs', 'poly2', 1.7), ('rs', 'linear', 1.0), ('rs', 'square', 1.0)])
def test_hinge_sum_transforms_and_their_derivatives(loss, lf, p):
    S = torch.tensor([0.0, 0.3, 2.0, 57.0], dtype=torch.float64).requires_grad_(True)
    l, dl = E._rs_transform(S, loss, lf, p)
    want = {'log': lambda s: torch.log(1 + s), 'exp': lambda s: 1 - p ** (-s), 'poly': lambda s: s ** p,
            'poly2': lambda s: (1 + s) ** p, 'linear': lambda s: s, 'square': lambda s: s * s}['log' if loss == 'warp' else lf]
    assert torch.allclose(l, want(S))
    g, = torch.autograd.grad(l.sum(), S)
    ok = torch.isfinite(g)
    assert torch.allclose(dl[ok].detach(), g[ok])


def test_canonical_bucket_order_sorts_inside_rows_only():
    rng = np.random.default_rng(0)
    cnt = np.array([3, 1, 70, 5, 2], dtype=np.int64)
    order = np.array([2, 0, 4, 1, 3])                               # rows are NOT laid out in list order
    base = np.zeros(5, dtype=np.int64)
    pos = 0
    for u in order:
        base[u] = pos
        pos += cnt[u]
    occ = int(cnt.sum())
    src = rng.integers(0, 50, occ).astype(np.int32)
    w = rng.random(occ).astype(np.float32)
    plan = types.SimpleNamespace(counters=torch.tensor([5, occ, 0, 0, 0, 0, 0, 0], dtype=torch.int32),
                                 row_base=torch.tensor(base, dtype=torch.int32), row_cnt=torch.tensor(cnt, dtype=torch.int32),
                                 bucket_src=torch.tensor(np.concatenate([src, [99, 99]]).astype(np.int32)),
                                 bucket_w=torch.tensor(np.concatenate([w, [9., 9.]]).astype(np.float32)))
    E._canonical_buckets(plan)
    s2, w2 = plan.bucket_src.numpy(), plan.bucket_w.numpy()
    assert s2[-2:].tolist() == [99, 99] and w2[-2:].tolist() == [9.0, 9.0]          # nothing beyond the used part moves
    for u in range(5):
        a, b = int(base[u]), int(base[u] + cnt[u])
        assert np.all(np.diff(s2[a:b]) >= 0)
        assert sorted(zip(src[a:b].tolist(), w[a:b].tolist())) == sorted(zip(s2[a:b].tolist(), w2[a:b].tolist()))


def test_checkpoint_paths_follow_the_shard():
    from arecsys_b200.hmf.hmf_model import _Saver
    mk = lambda shard: _Saver(types.SimpleNamespace(att_emb=types.SimpleNamespace(shard=shard)))
    assert mk(None)._shard_path('/t/best.ckpt-0') == '/t/best.ckpt-0'
    assert mk((4, 1))._shard_path('/t/best.ckpt-0') == '/t/best.ckpt-0.shard1of4'
    assert mk((4, 1))._shard_path('/t/best.ckpt-0.shard3of4') == '/t/best.ckpt-0.shard1of4'      # an index written by another rank
    assert mk(None)._shard_path('/t/best.ckpt-0.shard3of4') == '/t/best.ckpt-0'
