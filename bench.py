#!/usr/bin/env python
"""bench.py — interactions/sec of the HMF training hot path (BASELINE.json metric) on B200.

Workload (BASELINE.json configs[1]): HMF, dim 128, batch 4096, synthetic
1M users / 1M items, per side 1 id attribute + 8 multi-hot attributes (mean bag 12, vocab 1e5,
Zipf tokens), loss `mw` (WMRB over a 1024-item sampled pool refreshed every 50 steps, the
reference's sampled ranking loss), keep_prob 0.5, Adagrad lr 0.1.  A "step" is one full
training step: user/item attribute pooling, scoring, loss, backward, de-duplicated sparse Adagrad.

  value : device-timed throughput, inputs (the step's user/item ids) already resident in HBM
  e2e   : the same through the public model.step() with HOST (pinned) id buffers, the H2D copy of
          the ids and the D2H read of the loss inside the timed region
  roofline     : the dominant kernel, ALGORITHMIC bytes / CUDA-event duration vs MEASURED_PEAKS.json
  cpu_baseline : oracle/torch_cpu_ref.py (the reference's literal op sequence on the host cores)
                 on a bounded sample (rank 0, N=1)
`--impl reference` times that CPU restatement alone (TF-1/Python-2 cannot run here).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=400)
    ap.add_argument('--warmup', type=int, default=20)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--loss', default='mw')
    ap.add_argument('--mb', type=int, default=4096)
    ap.add_argument('--dim', type=int, default=128)
    ap.add_argument('--n-users', type=int, default=1000000)
    ap.add_argument('--n-items', type=int, default=1000000)
    ap.add_argument('--n-mulhot', type=int, default=8)
    ap.add_argument('--vocab-m', type=int, default=100000)
    ap.add_argument('--mean-len', type=int, default=12)
    ap.add_argument('--n-sampled', type=int, default=1024)
    ap.add_argument('--n-resample', type=int, default=50)
    ap.add_argument('--keep-prob', type=float, default=0.5)
    ap.add_argument('--lr', type=float, default=0.1)
    ap.add_argument('--cpu-mb', type=int, default=256, help='rows per CPU-baseline step (bounded sample)')
    ap.add_argument('--cpu-steps', type=int, default=3)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-graph', action='store_true', help='launch every kernel eagerly instead of replaying a CUDA graph')
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.isfile(p):
        d = json.load(open(p))
        return float(d['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    return 6650.0, 'fallback (B200_PROFILING.md 6.65 TB/s)'


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
         'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile('w+', suffix='.csv', delete=False)
        self.p = None
        self.idx = gpu_index

    def start(self):
        self.skip = 0
        if self.idx is None:
            return
        try:
            self.p = subprocess.Popen(['nvidia-smi', '-i', str(self.idx), '--query-gpu=' + self.Q,
                                       '--format=csv,noheader,nounits', '-lms', '25'], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def mark(self):
        """Forget the samples taken so far (warm-up): only what follows is reported."""
        if self.p is None:
            return
        try:
            self.skip = len([r for r in open(self.f.name).read().splitlines() if r.strip()])
        except OSError:
            self.skip = 0

    def stop(self):
        if self.p is None:
            try:
                os.unlink(self.f.name)
            except OSError:
                pass
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.p.terminate()
        self.p.wait()
        self.f.flush()
        rows = [r.strip().split(', ') for r in open(self.f.name).read().strip().splitlines() if r.strip()]
        rows = rows[getattr(self, 'skip', 0):] or rows[-1:]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except ValueError:
                continue
            for n, v in zip(names, r[5:9]):
                if v.strip().lower().startswith('active'):
                    reasons.add(n)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': sorted(reasons), 'samples': len(sm)}


def build_workload(a, rank, world):
    import arecsys_b200  # noqa: F401
    from arecsys_b200.utils import synthetic, prepare_train
    t0 = time.time()
    ua, ia, i2l, l2i = synthetic.make_dataset(a.n_users, a.n_items, a.n_mulhot, a.vocab_m, a.mean_len, 64,
                                              1.05, seed=0)
    n_batches = a.warmup + a.steps
    users, items = synthetic.make_interactions(a.n_users, a.n_items, a.mb * world * n_batches * 2, seed=rank)
    pop, counts = np.unique(items, return_counts=True)
    p = np.power(counts / counts.sum(), 0.5)
    p = p / p.sum()
    pos = prepare_train.positives_csr(users, items, ua.num_entities)
    if rank == 0:
        print('[bench] workload built in %.1fs' % (time.time() - t0), file=sys.stderr)
    return ua, ia, i2l, l2i, users, items, pop, p, pos


def algorithmic_bytes(att, ids, dim, with_bias):
    """bytes one pooling launch must move, nominal (per occurrence) and conservative
    (unique (table,row) pairs; duplicates hit L2)."""
    n = len(ids)
    occ = uniq = 0
    idx_bytes = 0
    for f in range(att.num_features_cat):
        tok = att.features_cat[f][ids]
        occ += n
        uniq += len(np.unique(tok))
        idx_bytes += 4 * n
    for f in range(att.num_features_mulhot):
        s = att.mulhot_starts[f][ids].astype(np.int64)
        l = att.mulhot_lengths[f][ids].astype(np.int64)
        off = np.concatenate([[0], np.cumsum(l)[:-1]])
        pos = np.arange(int(l.sum()), dtype=np.int64) - np.repeat(off, l) + np.repeat(s, l)
        tok = att.features_mulhot[f][pos]
        occ += len(tok)
        uniq += len(np.unique(tok))
        idx_bytes += 4 * len(tok) + 8 * n
    row = dim * 4 + (4 if with_bias else 0)
    fwd_nom = occ * row + idx_bytes + n * 4 + n * row
    fwd_uni = uniq * row + idx_bytes + n * 4 + n * row
    # backward + fused Adagrad: dOut read + plan (16 B per row + 8 B per occurrence) + RMW of E and acc
    bwd_nom = n * row + 16 * uniq + 8 * occ + 4 * occ * row
    bwd_uni = n * row + 16 * uniq + 8 * occ + 4 * uniq * row
    return dict(occ=occ, uniq=uniq, fwd_nominal=fwd_nom, fwd_unique=fwd_uni, bwd_nominal=bwd_nom, bwd_unique=bwd_uni)


def cpu_baseline(a, ua, ia, l2i, users, items, pos, sampler_pop, sampler_p, steps, warm=1, threads=None):
    """The reference's literal op sequence on the host cores (oracle/torch_cpu_ref.py), bounded sample."""
    from oracle.torch_cpu_ref import TorchRefHMF
    torch.set_num_threads(threads or os.cpu_count() or 1)
    rng = np.random.default_rng(1)
    dim = a.dim
    lim = 0.05
    params = {}
    for prefix, att, bias in (('user', ua, False), ('item', ia, True)):
        for tag, n, V in (('cat', att.num_features_cat, att._embedding_classes_list_cat),
                          ('mulhot', att.num_features_mulhot, att._embedding_classes_list_mulhot)):
            for k in range(n):
                params['%sembed_%s_%d' % (prefix, tag, k)] = rng.uniform(-lim, lim, (V[k], dim)).astype(np.float32)
                if bias:
                    params['%s_bias_%s_%d' % (prefix, tag, k)] = rng.uniform(-lim, lim, (V[k], 1)).astype(np.float32)

    class _Id(object):
        def __getitem__(self, k):
            return int(k)
    ref = TorchRefHMF(ua, ia, params, l2i, _Id(), loss=a.loss, keep_prob=a.keep_prob, learning_rate=a.lr,
                      n_sampled=a.n_sampled if a.loss == 'mw' else None)
    ptr_, pit = pos
    mb = a.cpu_mb
    times = []
    for s in range(warm + steps):
        u = users[s * mb:(s + 1) * mb]
        it = items[s * mb:(s + 1) * mb]
        pd = {int(x): pit[ptr_[x]:ptr_[x + 1]].tolist() for x in u}
        ref.pos = ref.pos_eval = pd
        sampled = None
        if a.loss == 'mw' and s == 0:
            sampled = rng.choice(sampler_pop, a.n_sampled, replace=False, p=sampler_p)
        t0 = time.perf_counter()
        ref.step(u, it, item_sampled=sampled)
        if s >= warm:
            times.append(time.perf_counter() - t0)
    ms = 1e3 * float(np.mean(times))
    return mb / (ms / 1e3), ms


def main():
    a = parse()
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    cfg = {'workload': ('C4' if a.n_items >= 10000000 else 'C2') + ': HMF dim=%d batch=%d synthetic %d users / %d items, per side id + %d multi-hot attrs '
                       '(mean bag %d, vocab %d, Zipf 1.05), loss=%s%s, keep_prob=%.2f, Adagrad' % (
                           a.dim, a.mb, a.n_users, a.n_items, a.n_mulhot, a.mean_len, a.vocab_m, a.loss,
                           (' n_sampled=%d n_resample=%d' % (a.n_sampled, a.n_resample)) if a.loss == 'mw' else '',
                           a.keep_prob),
           'global_batch': a.mb * world, 'parallelism': ('tables row-sharded x%d (row t on rank t %% N), RS/AR/AG of pooled vectors over NCCL' % world)
           if world > 1 else 'single GPU',
           'l2_policy': 'inputs larger than L2: tables+accumulators 3.7 GB, rows gathered at random each step'}

    if a.impl == 'reference':
        if rank != 0:
            return
        ua, ia, i2l, l2i, users, items, pop, p, pos = build_workload(a, 0, 1)
        # each timed step = one bounded SAMPLE of the C2 step: a.cpu_mb rows of the 4096-row batch through the
        # reference's literal op order (at 4096 rows that order materialises a [10^6, 4096] fp32 token-score
        # matrix per id table = 16 GB; the sample keeps the arm within minutes).  Exactly --steps timed steps
        # after --warmup untimed ones; throughput = rows really processed / time.
        v, ms = cpu_baseline(a, ua, ia, l2i, users, items, pos, pop, p, max(1, a.steps), warm=max(0, a.warmup))
        sample = ('each step = %d rows of the %d-row C2 batch, literal token-score order; %d timed steps after %d '
                  'warm-up steps; all %d host threads' % (a.cpu_mb, a.mb, max(1, a.steps), max(0, a.warmup),
                                                          os.cpu_count() or 1))
        print(json.dumps({'impl': 'reference', 'metric': 'interactions/sec', 'value': v, 'unit': 'interactions/s',
                          'n_gpus': a.gpus, 'steps': a.steps, 'warmup': a.warmup, 'ms_per_step': ms,
                          'rows_per_step': a.cpu_mb, 'higher_is_better': True, 'scaling': 'weak',
                          'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic', 'config': cfg, 'sample': sample,
                          'cpu_baseline': {'value': v, 'unit': 'interactions/s', 'cores': os.cpu_count(),
                                           'kind': 'port', 'sample': sample},
                          'e2e': {'value': v, 'unit': 'interactions/s', 'h2d_bytes_per_step': 0,
                                  'd2h_bytes_per_step': 0}}))
        return

    assert torch.cuda.is_available(), 'bench.py needs a GPU (no CPU fallback)'
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=dev)

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    import arecsys_b200  # noqa: F401
    from arecsys_b200 import _lib
    from arecsys_b200.hmf.hmf_model import LatentProductModel
    from arecsys_b200.utils.prepare_train import DeviceItemSampler
    # N > 1: tables row-sharded over the ranks (row t on rank t % N), weak scaling: the global batch
    # is N * mb interactions, identical on every rank (same seed); rank r scores rows [r*mb,(r+1)*mb)
    ua, ia, i2l, l2i, users, items, pop, p, pos = build_workload(a, 0, world)
    n_sampled = a.n_sampled if a.loss == 'mw' else None
    if world > 1:
        from arecsys_b200.hmf.sharded import ShardedLatentProductModel
        model = ShardedLatentProductModel(a.n_users, a.n_items, a.dim, 1, a.mb, a.lr, 1.0, ua, ia, i2l, l2i,
                                          loss_function=a.loss, dropout=a.keep_prob, n_sampled=n_sampled, seed=1)
    else:
        model = LatentProductModel(a.n_users, a.n_items, a.dim, 1, a.mb, a.lr, 1.0, ua, ia, i2l, l2i,
                                   loss_function=a.loss, dropout=a.keep_prob, n_sampled=n_sampled, seed=1)
    if a.loss != 'ce':
        model.prepare_warp(pos, pos)
    sampler = DeviceItemSampler(pop, p, dev, seed=0)
    mb = a.mb * world                       # rows per (global) step
    nb = a.warmup + a.steps
    u_host = torch.from_numpy(users[:2 * nb * mb].reshape(2 * nb, mb)).pin_memory()
    i_host = torch.from_numpy(items[:2 * nb * mb].reshape(2 * nb, mb)).pin_memory()
    u_dev = u_host[:nb].to(dev)
    i_dev = i_host[:nb].to(dev)

    state = {'step': 0, 'graph': False}
    use_graph = not a.no_graph           # N > 1: the NCCL exchanges are captured with the kernels

    # pool refresh every n_resample steps, phased so that a refresh falls INSIDE the timed window even when the
    # driver times only 20 steps (then 1 refresh per 20 steps: above the real 1-in-50 rate, i.e. conservative)
    refresh_at = (a.warmup + 2 + min(10, a.steps // 2)) % a.n_resample

    def run_step(u, it, sync):
        sampled = None
        if a.loss == 'mw' and (state['step'] == 0 or state['step'] % a.n_resample == refresh_at):
            sampled = sampler.sample(a.n_sampled)
            if world > 1:
                torch.distributed.broadcast(sampled, 0)       # every rank must score the same pool
        state['step'] += 1
        if state['graph'] and _lib.timeline is None:
            if sampled is not None:
                model.att_emb.pass_sampled_items(sampled)      # in place: the graph reads the same buffers
            return model.replay_step(u, it, sync=sync)
        return model.step(None, u, it, None, sampled, None, loss=a.loss, sync=sync)

    # ---------------- N > 1: the sharded step must compute what the single-GPU step computes ----------
    sharded_check = None
    if world > 1:
        ones = torch.ones((a.mb, a.dim), dtype=torch.float32, device=dev)           # dropout mask injected: all kept
        pool0 = sampler.sample(a.n_sampled)
        torch.distributed.broadcast(pool0, 0)
        l_sh = model.step(None, u_dev[0], i_dev[0], None, pool0, None, loss=a.loss, masks=[ones], sync=True)
        if rank == 0:
            single = LatentProductModel(a.n_users, a.n_items, a.dim, 1, mb, a.lr, 1.0, ua, ia, i2l, l2i,
                                        loss_function=a.loss, dropout=a.keep_prob, n_sampled=n_sampled, seed=1)
            single.prepare_warp(pos, pos)
            l_1 = single.step(None, u_dev[0], i_dev[0], None, pool0, None, loss=a.loss,
                              masks=[torch.ones((mb, a.dim), dtype=torch.float32, device=dev)], sync=True)
            ok = abs(l_sh - l_1) <= 1e-3 * max(1.0, abs(l_1))
            sharded_check = {'loss_sharded': l_sh, 'loss_single_gpu': l_1, 'ok': bool(ok),
                             'what': 'step 0, same global batch / pool / weights, dropout mask of ones'}
            del single
            torch.cuda.empty_cache()
            assert ok, 'sharded step disagrees with the single-GPU step: %r' % (sharded_check,)
        state['step'] += 1
        barrier()
        if getattr(model, 'px', None) is not None:
            cfg['parallelism'] = ('tables row-sharded x%d (row t on rank t %% N); pooled vectors and gradient rows exchanged '
                                  'over NVLink peer memory from inside the kernels (lookup + reduce-scatter / all-reduce fused, '
                                  'all-gather by push, two device-side barriers per step; ARX_PEER=0: NCCL collectives)' % world)
    # ---------------- value: ids resident in HBM ---------------------------------------
    # The clock sampler (an nvidia-smi loop) is started BEFORE the warm-up steps and on rank 0 only: its start-up
    # (process spawn + NVML initialisation, which takes driver-wide locks) used to land at the head of the timed window
    # and, with one sampler per rank, stalled graph launches for 3-17 ms of a 15 ms window at N > 1.
    clocks = ClockSampler(local_rank if rank == 0 else None)
    clocks.start()
    for s in range(a.warmup):
        run_step(u_dev[s], i_dev[s], False)
    if use_graph:
        model.capture_step(u_dev[0], i_dev[0], loss=a.loss)    # whole step -> one CUDA graph launch
        state['graph'] = True
        run_step(u_dev[1], i_dev[1], False)
    barrier()
    clocks.mark()                                              # samples from here on are "during the timed region"
    l0 = _lib.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for s in range(a.warmup, nb):
        run_step(u_dev[s], i_dev[s], False)
    ev1.record()
    barrier()
    launches = _lib.launch_count - l0
    ms_total = ev0.elapsed_time(ev1)
    if world > 1 and getattr(model, 'px', None) is not None:
        model.px.check()                                       # the device-side barrier's watchdog never fired
    # ---------------- e2e: host ids -> H2D -> step -> loss D2H -------------------------
    # The training-loop form of the public API: every step copies its ids from pinned host memory, replays the
    # step and copies its loss back to the host; the host reads the loss one step late (replay_step(sync='lag')),
    # so it never stalls on the step it has just launched.  The last loss is read before the clock stops.
    def e2e_step(s):
        u = u_host[nb + s].to(dev, non_blocking=True)
        it = i_host[nb + s].to(dev, non_blocking=True)
        if state['graph']:
            return run_step(u, it, 'lag')
        return run_step(u, it, True)
    for s in range(min(a.warmup, 5)):
        e2e_step(s)
    if state['graph']:
        model.flush_loss()
    barrier()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    last = 0.0
    for s in range(a.warmup, nb):
        l = e2e_step(s)
        last = l if l is not None else last
    if state['graph']:
        last = model.flush_loss()
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1)
    wall_e2e = (time.perf_counter() - t0) * 1e3
    # the reference-facing call as a reference user makes it (B3): model.step(session, [python ints], [python ints])
    # -> python float, eager launches, synchronous loss; a bounded number of steps
    n_list = min(a.steps, 20)
    lists = [(u_host[nb + s].tolist(), i_host[nb + s].tolist()) for s in range(n_list)]
    g_was, state['graph'] = state['graph'], False
    run_step(lists[0][0], lists[0][1], True)
    barrier()
    t0 = time.perf_counter()
    for s in range(n_list):
        run_step(lists[s][0], lists[s][1], True)
    barrier()
    wall_lists = (time.perf_counter() - t0) * 1e3
    state['graph'] = g_was
    clk = clocks.stop()
    # ---------------- per-kernel CUDA-event durations (separate pass: the event pairs would
    # otherwise perturb `value`); same batches, same stream -------------------------------
    _lib.timeline = []
    for s in range(a.warmup, nb):
        run_step(u_dev[s], i_dev[s], False)
    barrier()
    tl, _lib.timeline = _lib.timeline, None

    if world > 1:
        t = torch.tensor([ms_total, ms_e2e, wall_e2e, wall_lists], device=dev, dtype=torch.float64)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        ms_total, ms_e2e, wall_e2e, wall_lists = (float(x) for x in t)
    if rank != 0:
        return
    ms_step = ms_total / a.steps
    value = mb * a.steps / (ms_total / 1e3)               # mb already counts all ranks' rows
    e2e_v = mb * a.steps / (max(ms_e2e, wall_e2e) / 1e3)

    # ---------------- per-kernel durations and roofline -----------------------------------
    agg = {}
    for name, tag, s0, s1 in tl:
        k = '%s:%s' % (name, tag) if name in ('arx_pool_fwd', 'arx_pool_bwd_apply', 'arx_pool_fwd_many', 'arx_pool_bwd_apply_many') else name
        d = agg.setdefault(k, [0.0, 0])
        d[0] += s0.elapsed_time(s1)
        d[1] += 1
    per_kernel = {k: {'ms_per_step': v[0] / a.steps, 'launches_per_step': v[1] / a.steps,
                      'avg_us': 1e3 * v[0] / max(v[1], 1)} for k, v in agg.items()}
    peak, peak_src = peaks()
    sb = a.warmup
    ub = algorithmic_bytes(ua, users[sb * mb:(sb + 1) * mb], a.dim, False)
    item_ids = items[sb * mb:(sb + 1) * mb]
    ib = algorithmic_bytes(ia, item_ids, a.dim, True)
    roofs = {}

    def roof(key, nbytes_unique, nbytes_nominal, what):
        if key not in per_kernel or per_kernel[key]['launches_per_step'] == 0:
            return
        us = per_kernel[key]['avg_us']
        roofs[key] = {'bound': 'hbm', 'achieved': nbytes_unique / us / 1e3, 'peak': peak, 'unit': 'GB/s',
                      'frac': nbytes_unique / us / 1e3 / peak, 'traffic': None,
                      'achieved_nominal': nbytes_nominal / us / 1e3, 'algorithmic_bytes': nbytes_unique,
                      'algorithmic_bytes_nominal': nbytes_nominal, 'avg_us': us, 'kernel': what, 'peak_source': peak_src}
    step_roof = None
    if world == 1:       # per-rank byte counts under sharding are 1/N of these: roofline only at N=1
        roof('arx_pool_fwd:user', ub['fwd_unique'], ub['fwd_nominal'], 'pool_fwd_flat_kernel<1> user side, %d bags' % (a.mb * (a.n_mulhot + 1)))
        roof('arx_pool_bwd_apply:user', ub['bwd_unique'], ub['bwd_nominal'], 'pool_bwd_apply_kernel<4> user side')
        # item side: the sampled pool and the target items are two lookups whose gradients are de-duplicated
        # together (one plan, one apply launch); byte counts use the pool that is current at the end of the run
        if a.loss == 'mw' and model.att_emb.sampled_ids is not None:
            pool_ids = model.att_emb.sampled_ids.cpu().numpy().astype(np.int64)
            ib2 = algorithmic_bytes(ia, np.concatenate([pool_ids, item_ids.astype(np.int64)]), a.dim, True)
            # the three lookups of the step (users, pool, targets) run as ONE launch (arx_pool_fwd_many)
            roof('arx_pool_fwd_many:many', ub['fwd_unique'] + ib2['fwd_unique'], ub['fwd_nominal'] + ib2['fwd_nominal'],
                 'pool_fwd_flat_many_kernel<1>: users + sampled pool + target items, %d entities' % (2 * a.mb + len(pool_ids)))
            roof('arx_pool_bwd_apply:item', ib2['bwd_unique'], ib2['bwd_nominal'],
                 'pool_bwd_apply_kernel<4> item side (pool + targets, %d entity rows)' % (len(pool_ids) + len(item_ids)))
            # both applies as ONE launch (arx_pool_bwd_apply_many): user set + item set
            roof('arx_pool_bwd_apply_many:user+item', ub['bwd_unique'] + ib2['bwd_unique'], ub['bwd_nominal'] + ib2['bwd_nominal'],
                 'pool_bwd_apply2_kernel: user tables + item tables (pool + targets) in one launch')
            # whole step against the HBM roofline: embedding forward + backward bytes of both sides,
            # unique-row accounting, over the measured step time (everything else in the step counts as overhead)
            step_bytes = ub['fwd_unique'] + ub['bwd_unique'] + ib2['fwd_unique'] + ib2['bwd_unique']
            step_roof = {'bound': 'hbm', 'algorithmic_bytes': step_bytes, 'achieved': step_bytes / (ms_step * 1e-3) / 1e9,
                         'peak': peak, 'unit': 'GB/s', 'frac': step_bytes / (ms_step * 1e-3) / 1e9 / peak,
                         'what': 'embedding fwd+bwd bytes of both sides (unique rows) / ms_per_step', 'peak_source': peak_src}
    dom = max(roofs, key=lambda k: per_kernel[k]['ms_per_step']) if roofs else None
    ku, ki = 'arx_pool_bwd_apply:user', 'arx_pool_bwd_apply:item'
    if ku in roofs and ki in roofs:
        # the dominant KERNEL is pool_bwd_apply_kernel<4>: two launches per step (user tables, item tables).  Its
        # roofline entry is what the contract defines — algorithmic bytes per launch / the kernel's average launch
        # duration, over BOTH launches; the per-side entries stay in roofline_all.
        ru, ri = roofs[ku], roofs[ki]
        us = 0.5 * (ru['avg_us'] + ri['avg_us'])
        nb_u = 0.5 * (ru['algorithmic_bytes'] + ri['algorithmic_bytes'])
        nb_n = 0.5 * (ru['algorithmic_bytes_nominal'] + ri['algorithmic_bytes_nominal'])
        key = 'arx_pool_bwd_apply:both'
        roofs[key] = {'bound': 'hbm', 'achieved': nb_u / us / 1e3, 'peak': peak, 'unit': 'GB/s', 'frac': nb_u / us / 1e3 / peak,
                      'traffic': None, 'achieved_nominal': nb_n / us / 1e3, 'algorithmic_bytes': nb_u,
                      'algorithmic_bytes_nominal': nb_n, 'avg_us': us, 'launches_per_step': 2,
                      'kernel': 'pool_bwd_apply_kernel<4>: average over its two launches per step (user tables: %.1f us, '
                                'frac %.3f; item tables (pool + targets): %.1f us, frac %.3f)'
                                % (ru['avg_us'], ru['frac'], ri['avg_us'], ri['frac']), 'peak_source': peak_src}
        dom = key
    tpath = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'profiles', 'traffic.json')
    # dram bytes per launch from the committed ncu --set full capture (profiles/r1_ncu_summary.md; the first
    # launch of each kernel in the eager step is the user side)
    tkeys = {'arx_pool_fwd:user': 'void pool_fwd_flat_kernel<1> grid %d #1' % ((a.mb + 6) // 7),
             'arx_pool_bwd_apply:user': 'void pool_bwd_apply_kernel<4> grid 592 #1',
             'arx_pool_bwd_apply:item': 'void pool_bwd_apply_kernel<4> grid 592 #2',
             'arx_pool_fwd_many:many': 'void pool_fwd_flat_many_kernel<1> grid 1319 #1'}
    if os.path.exists(tpath):
        tj = json.load(open(tpath))
        for k, r in roofs.items():
            t = tj.get(tkeys.get(k, ''))
            if t:
                r['traffic'] = t['dram_bytes_read'] + t['dram_bytes_write']
                r['traffic_source'] = t['source']
        if 'arx_pool_bwd_apply:both' in roofs and roofs[ku].get('traffic') and roofs[ki].get('traffic'):
            roofs['arx_pool_bwd_apply:both']['traffic'] = 0.5 * (roofs[ku]['traffic'] + roofs[ki]['traffic'])
            roofs['arx_pool_bwd_apply:both']['traffic_source'] = roofs[ki].get('traffic_source')

    out = {'metric': 'interactions/sec', 'value': value, 'unit': 'interactions/s', 'n_gpus': world,
           'steps': a.steps, 'warmup': a.warmup, 'ms_per_step': ms_step, 'higher_is_better': True,
           'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic', 'config': cfg,
           'clocks': clk,
           'e2e': {'value': e2e_v, 'unit': 'interactions/s', 'h2d_bytes_per_step': 2 * mb * 4,
                   'd2h_bytes_per_step': 4, 'ms_per_step': max(ms_e2e, wall_e2e) / a.steps, 'last_loss': last,
                   'api': "model.replay_step(pinned host ids, sync='lag'): per step H2D of the ids, the captured "
                          "step, D2H of the loss (read by the host one step late; the last one before the clock stops)",
                   'via_model_step_lists': {'value': mb * n_list / (wall_lists / 1e3), 'unit': 'interactions/s',
                                            'ms_per_step': wall_lists / n_list, 'steps': n_list,
                                            'api': 'model.step(None, [python ints], [python ints]) -> float, as the '
                                                   'reference runner calls it (eager launches, synchronous loss)'}},
           'roofline_step': step_roof, 'sharded_check': sharded_check,
           'gpu_launches': launches, 'launch_mode': 'cuda graph replay (1 graph launch per step)' if use_graph else 'eager',
           'roofline': roofs.get(dom), 'roofline_all': roofs, 'per_kernel': per_kernel,
           'batch_stats': {'user_occurrences': ub['occ'], 'user_unique_rows': ub['uniq'],
                           'item_occurrences': ib['occ'], 'item_unique_rows': ib['uniq']}}
    if world == 1 and not a.no_cpu_baseline:
        del model
        torch.cuda.empty_cache()
        v, ms = cpu_baseline(a, ua, ia, l2i, users, items, pos, pop, p, a.cpu_steps)
        v1, ms1 = cpu_baseline(a, ua, ia, l2i, users, items, pos, pop, p, 1, warm=1, threads=1)
        out['cpu_baseline_1thread'] = {'value': v1, 'unit': 'interactions/s', 'cores': 1, 'kind': 'port',
                                       'ms_per_step': ms1, 'sample': '%d-row batch, 1 timed step after 1 warm-up, '
                                       'torch.set_num_threads(1)' % a.cpu_mb}
        out['cpu_baseline'] = {'value': v, 'unit': 'interactions/s', 'cores': os.cpu_count(), 'kind': 'port',
                               'ms_per_step': ms,
                               'sample': '%d-row batches (of the %d-row step), %d timed steps after 1 warm-up, '
                                         'reference op order (token scores for every table row, then gather+pool)'
                                         % (a.cpu_mb, a.mb, a.cpu_steps)}
    print(json.dumps(out))


if __name__ == '__main__':
    main()
