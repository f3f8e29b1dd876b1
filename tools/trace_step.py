#!/usr/bin/env python
"""Kernel timeline of one CUDA-graph replay of the C2 training step (torch.profiler / CUPTI):
prints every kernel of the last profiled step with start offset, duration and stream, so that the
critical path and the gaps between dependent kernels can be read off.

    python tools/trace_step.py [--loss mw] > gpurun_out/trace_step.txt
"""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--loss', default='mw')
    ap.add_argument('--mb', type=int, default=4096)
    a = ap.parse_args()
    import bench
    sys.argv = ['bench.py', '--loss', a.loss, '--mb', str(a.mb), '--steps', '30', '--warmup', '5']
    args = bench.parse()
    import arecsys_b200  # noqa: F401
    from arecsys_b200.hmf.hmf_model import LatentProductModel
    from arecsys_b200.utils.prepare_train import DeviceItemSampler
    dev = torch.device('cuda', 0)
    ua, ia, i2l, l2i, users, items, pop, p, pos = bench.build_workload(args, 0, 1)
    model = LatentProductModel(args.n_users, args.n_items, args.dim, 1, args.mb, args.lr, 1.0, ua, ia, i2l, l2i,
                               loss_function=args.loss, dropout=args.keep_prob,
                               n_sampled=args.n_sampled if args.loss == 'mw' else None, seed=1)
    if args.loss != 'ce':
        model.prepare_warp(pos, pos)
    sampler = DeviceItemSampler(pop, p, dev, seed=0)
    nb = 40
    u_dev = torch.from_numpy(users[:nb * args.mb].reshape(nb, args.mb)).to(dev)
    i_dev = torch.from_numpy(items[:nb * args.mb].reshape(nb, args.mb)).to(dev)
    sampled = sampler.sample(args.n_sampled) if args.loss == 'mw' else None
    for s in range(5):
        model.step(None, u_dev[s], i_dev[s], None, sampled if s == 0 else None, None, loss=args.loss, sync=False)
    model.capture_step(u_dev[0], i_dev[0], loss=args.loss)
    for s in range(5, 15):
        model.replay_step(u_dev[s], i_dev[s], sync=False)
    torch.cuda.synchronize()
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for s in range(15, 19):
            model.replay_step(u_dev[s], i_dev[s], sync=False)
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    evs.sort(key=lambda e: e.time_range.start)
    # split into steps at the big gaps / by the id-copy memcpy that starts every replay
    starts = [i for i, e in enumerate(evs) if 'Memcpy' in e.name]
    # every replay begins with two DtoD id copies: take the last pair as the start of the last step
    first = starts[-2] if len(starts) >= 2 else 0
    last = evs[first:]
    t0 = last[0].time_range.start
    print('# %d kernels in the last replay; columns: start_us dur_us end_us stream name' % len(last))
    for e in last:
        st = e.time_range.start - t0
        print('%8.1f %7.1f %8.1f  s%-3s %s' % (st, e.time_range.elapsed_us(), st + e.time_range.elapsed_us(),
                                              getattr(e, 'device_index', '?') if False else '', e.name[:90]))
    print('# step span: %.1f us' % (last[-1].time_range.end - t0))
    # device idle time between consecutive replays (end of the last kernel of a step -> first id copy of the next)
    pairs = [starts[i] for i in range(0, len(starts) - 1, 2)]
    for a_, b_ in zip(pairs[:-1], pairs[1:]):
        seg = evs[a_:b_]
        end = max(e.time_range.end for e in seg)
        print('# replay: span %.1f us, idle before the next replay %.1f us' % (
            end - seg[0].time_range.start, evs[b_].time_range.start - end))
    cpu = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CPU and 'cudaGraphLaunch' in e.name]
    if cpu:
        print('# cudaGraphLaunch host time: %s us' % [round(e.time_range.elapsed_us(), 1) for e in cpu])


if __name__ == '__main__':
    main()
