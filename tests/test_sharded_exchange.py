"""World-size-2 tests of the row-sharded path.
CPU (gloo): the exchange layer + the sharding arithmetic, with partial pooled vectors computed by
the NumPy oracle restricted to the rows each rank owns.
GPU (needs 2 devices, NCCL): ShardedLatentProductModel against the unsharded oracle."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))


def _free_port():
    s = socket.socket(); s.bind(('127.0.0.1', 0)); p = s.getsockname()[1]; s.close(); return p


def _setup():
    from helpers import small_dataset, random_params
    dim = 8
    ua, ia, _, l2i = small_dataset(40, 30, 2, 15, 3, 5, 0, None, dim)
    params = random_params(ua, ia, dim, 1, scale=0.4)
    return ua, ia, l2i, params, dim


def _owned_only(params, G, r):
    """zero every table row this rank does not own: pooling the result = the rank's partial sum."""
    out = {}
    for k, v in params.items():
        w = np.zeros_like(v)
        w[r::G] = v[r::G]
        out[k] = w
    return out


def _cpu_worker(rank, world, port, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'; os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    import arecsys_b200  # noqa: F401
    from arecsys_b200.hmf.exchange import RowShardExchange
    from oracle import np_oracle as O
    ua, ia, l2i, params, dim = _setup()
    ex = RowShardExchange()
    mb = 6
    rng = np.random.default_rng(0)
    users_g = rng.integers(0, 40, world * mb)
    full = O.OracleEmbeddingAttribute(ua, ia, mb, None, params)
    part = O.OracleEmbeddingAttribute(ua, ia, mb, None, _owned_only(params, world, rank))
    cat, mul = part._tables('user', ua)
    p_partial, _ = O.pool_entities(part, cat, mul, None, None, ua, users_g)
    got = ex.reduce_scatter_rows(torch.tensor(p_partial))
    fc, fm = full._tables('user', ua)
    want, _ = O.pool_entities(full, fc, fm, None, None, ua, users_g[rank * mb:(rank + 1) * mb])
    ok1 = np.allclose(got.numpy(), want, rtol=1e-12)
    g = ex.all_gather_rows(torch.full((mb, 3), float(rank)))
    ok2 = g.shape == (world * mb, 3) and all(float(g[k * mb, 0]) == k for k in range(world))
    a = ex.all_reduce(torch.ones(4) * (rank + 1))
    ok3 = float(a[0]) == sum(range(1, world + 1))
    q.put((rank, ok1, ok2, ok3))
    dist.destroy_process_group()


def test_exchange_and_sharding_arithmetic_gloo_world2():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_cpu_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(r[0] for r in res) == [0, 1] and all(all(r[1:]) for r in res), res


def _gpu_worker(rank, world, port, q, nonlinear='linear'):
    os.environ['MASTER_ADDR'] = '127.0.0.1'; os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    import arecsys_b200  # noqa: F401
    from arecsys_b200 import _lib
    from arecsys_b200.hmf.sharded import ShardedLatentProductModel
    from oracle import np_oracle as O
    from helpers import positives
    ua, ia, l2i, params, dim = _setup()
    hidden = 5
    if nonlinear != 'linear':        # replicated MLP tower: its gradients take the one dense all-reduce of the step
        from helpers import random_params
        params = random_params(ua, ia, dim, 1, scale=0.4, mlp_hidden=hidden)
    l2i_d = {int(v): int(l2i[v]) for v in range(len(l2i))}
    i2l_d = {v: k for k, v in l2i_d.items()}
    mb, ns = 8, 10
    _lib.exact_fp32 = True
    model = ShardedLatentProductModel(40, 30, dim, 1, mb, 0.3, 1.0, ua, ia, i2l_d, l2i_d, loss_function='mw',
                                      nonlinear=nonlinear, hidden_size=hidden,
                                      dropout=0.5, n_sampled=ns, params={k: v.copy() for k, v in params.items()})
    emb = O.OracleEmbeddingAttribute(ua, ia, world * mb, ns, params, item_ind2logit_ind=i2l_d, logit_ind2item_ind=l2i_d)
    om = O.OracleHMF(emb, loss='mw', nonlinear=nonlinear, keep_prob=0.5, learning_rate=0.3)
    rng = np.random.default_rng(5)
    ok = True
    for it in range(3):
        users = rng.integers(0, 40, world * mb); items = rng.integers(0, 30, world * mb)
        pos = positives(users, items, 40, rng, n_items=30)
        model.prepare_warp(pos, pos); emb.prepare_warp(pos, pos)
        sampled = [int(v) for v in rng.permutation(30)[:ns]] if it != 1 else None
        id2idx = {v: k for k, v in enumerate(sampled)} if sampled else None
        shapes = [(world * mb, dim)] if nonlinear == 'linear' else [(world * mb, dim), (world * mb, hidden), (world * mb, dim)]
        masks = [np.floor(rng.random(sh) + 0.5) for sh in shapes]
        lg = model.step(None, users.tolist(), items.tolist(), None, sampled, id2idx, loss='mw',
                        masks=[torch.tensor(mk[rank * mb:(rank + 1) * mb], dtype=torch.float32, device='cuda') for mk in masks])
        lo = om.step(users.tolist(), items.tolist(), sampled, id2idx, masks=masks)
        ok = ok and abs(lg - lo) <= 1e-4 * max(1.0, abs(lo))
        for k, v in om.emb.p.items():
            if k in model.att_emb.params:
                got = model.att_emb.params[k].cpu().numpy()
                ok = ok and np.allclose(got, v[rank::world].reshape(got.shape), rtol=1e-3, atol=2e-5)
            else:                                                   # replicated dense parameter: identical on every rank
                got = model.dense[k].detach().cpu().numpy()
                ok = ok and np.allclose(got, v.reshape(got.shape), rtol=1e-3, atol=2e-5)
    # evaluation loss (loss_eval) and recommendation on the sharded tables against the oracle at the global batch
    users = rng.integers(0, 40, world * mb); items = rng.integers(0, 30, world * mb)
    ev = model.step(None, users.tolist(), items.tolist(), forward_only=True, loss='mw')
    eo = om.step(users.tolist(), items.tolist(), forward_only=True)
    ok = ok and abs(ev - eo) <= 1e-4 * max(1.0, abs(eo))
    model.top_N_items = 7
    top = model.step(None, users.tolist(), None, recommend=True)
    want, _ = om.top_k(users.tolist(), model.top_N_items)
    ok = ok and top.shape == want.shape and bool((top == want).all())
    q.put((rank, bool(ok), lg, lo, ev, eo))
    dist.destroy_process_group()


@pytest.mark.gpu
@pytest.mark.parametrize('nonlinear', ['linear', 'relu'])
def test_sharded_hmf_matches_oracle_two_gpus(nonlinear):
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs (run with gpurun --gpus 2)')
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gpu_worker, args=(r, 2, port, q, nonlinear)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] for r in res), res


def _peer_worker(rank, world, port, q):
    """NVLink peer-memory exchange (arx_pool_fwd_many_push + arx_peer_push_rows + arx_peer_barrier) against the NCCL
    collectives on the same sharded model, and against the oracle at the global batch."""
    os.environ['MASTER_ADDR'] = '127.0.0.1'; os.environ['MASTER_PORT'] = str(port)
    os.environ['ARX_PEER_TIMEOUT_S'] = '30'
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    try:
        import arecsys_b200  # noqa: F401
        from arecsys_b200.hmf.sharded import ShardedLatentProductModel
        from oracle import np_oracle as O
        from helpers import small_dataset, random_params, positives
        dim, nu, ni = 128, 300, 200
        ua, ia, _, l2i = small_dataset(nu, ni, 3, 50, 4, 8, 0, None, dim)
        params = random_params(ua, ia, dim, 1, scale=0.1)
        l2i_d = {int(v): int(l2i[v]) for v in range(len(l2i))}
        i2l_d = {v: k for k, v in l2i_d.items()}
        mb, ns = 64, 64
        mk = lambda peer: ShardedLatentProductModel(nu, ni, dim, 1, mb, 0.3, 1.0, ua, ia, i2l_d, l2i_d, loss_function='mw',
                                                    dropout=0.5, n_sampled=ns, params={k: v.copy() for k, v in params.items()},
                                                    peer=peer)
        A, B = mk(True), mk(False)
        emb = O.OracleEmbeddingAttribute(ua, ia, world * mb, ns, {k: v.copy() for k, v in params.items()},
                                         item_ind2logit_ind=i2l_d, logit_ind2item_ind=l2i_d)
        om = O.OracleHMF(emb, loss='mw', keep_prob=0.5, learning_rate=0.3)
        rng = np.random.default_rng(5)
        ok, detail = True, []
        for it in range(4):
            users = rng.integers(0, nu, world * mb); items = rng.integers(0, ni, world * mb)
            pos = positives(users, items, nu, rng, n_items=ni)
            for mdl in (A, B):
                mdl.prepare_warp(pos, pos)
            emb.prepare_warp(pos, pos)
            sampled = [int(v) for v in rng.permutation(ni)[:ns]] if it != 1 else None
            id2idx = {v: k for k, v in enumerate(sampled)} if sampled else None
            mask = np.floor(rng.random((world * mb, dim)) + 0.5)
            mk_ = [torch.tensor(mask[rank * mb:(rank + 1) * mb], dtype=torch.float32, device='cuda')]
            la = A.step(None, users.tolist(), items.tolist(), None, sampled, id2idx, loss='mw', masks=mk_)
            lb = B.step(None, users.tolist(), items.tolist(), None, sampled, id2idx, loss='mw', masks=mk_)
            lo = om.step(users.tolist(), items.tolist(), sampled, id2idx, masks=[mask])
            A.px.check()
            ok = ok and getattr(B, 'px', None) is None and abs(la - lb) <= 1e-6 * max(1.0, abs(lb)) and abs(la - lo) <= 2e-3 * max(1.0, abs(lo))
            detail.append((la, lb, lo))
            for k in A.att_emb.params:
                a_, b_ = A.att_emb.params[k], B.att_emb.params[k]
                ok = ok and bool(torch.allclose(a_, b_, rtol=1e-5, atol=1e-7))
                v = om.emb.p[k]
                ok = ok and np.allclose(a_.cpu().numpy(), v[rank::world].reshape(a_.shape), rtol=2e-2, atol=2e-4)
        q.put((rank, bool(ok), detail))
    except Exception as e:           # noqa: BLE001  (the parent must hear about it instead of waiting for the queue)
        import traceback
        q.put((rank, False, traceback.format_exc()))
    dist.destroy_process_group()


@pytest.mark.gpu
def test_peer_memory_exchange_matches_nccl_and_oracle_two_gpus():
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs (run with gpurun --gpus 2)')
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_peer_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] for r in res), res
