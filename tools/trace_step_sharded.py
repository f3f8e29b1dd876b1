#!/usr/bin/env python
"""Kernel timeline of one CUDA-graph replay of the ROW-SHARDED C2 training step on rank 0 (torch.profiler / CUPTI):

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/trace_step_sharded.py > gpurun_out/trace_sharded.txt
"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import bench
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    sys.argv = ['bench.py', '--gpus', str(world), '--steps', '30', '--warmup', '5']
    args = bench.parse()
    import arecsys_b200  # noqa: F401
    from arecsys_b200.hmf.sharded import ShardedLatentProductModel
    from arecsys_b200.utils.prepare_train import DeviceItemSampler
    dev = torch.device('cuda', local)
    ua, ia, i2l, l2i, users, items, pop, p, pos = bench.build_workload(args, 0, world)
    model = ShardedLatentProductModel(args.n_users, args.n_items, args.dim, 1, args.mb, args.lr, 1.0, ua, ia, i2l, l2i,
                                      loss_function='mw', dropout=args.keep_prob, n_sampled=args.n_sampled, seed=1)
    model.prepare_warp(pos, pos)
    sampler = DeviceItemSampler(pop, p, dev, seed=0)
    gb = args.mb * world
    nb = 24
    u_dev = torch.from_numpy(users[:nb * gb].reshape(nb, gb)).to(dev)
    i_dev = torch.from_numpy(items[:nb * gb].reshape(nb, gb)).to(dev)
    sampled = sampler.sample(args.n_sampled)
    for s in range(4):
        model.step(None, u_dev[s], i_dev[s], None, sampled if s == 0 else None, None, loss='mw', sync=False)
    model.capture_step(u_dev[0], i_dev[0], loss='mw')
    for s in range(4, 12):
        model.replay_step(u_dev[s], i_dev[s], sync=False)
    torch.cuda.synchronize()
    dist.barrier()
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for s in range(12, 16):
            model.replay_step(u_dev[s], i_dev[s], sync=False)
        torch.cuda.synchronize()
    dist.barrier()
    if rank == 0:
        evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
        evs.sort(key=lambda e: e.time_range.start)
        starts = [i for i, e in enumerate(evs) if 'Memcpy' in e.name]
        first = starts[-2] if len(starts) >= 2 else 0
        last = evs[first:]
        t0 = last[0].time_range.start
        print('# rank 0 of %d; %d kernels in the last replay; columns: start_us dur_us end_us name' % (world, len(last)))
        for e in last:
            st = e.time_range.start - t0
            print('%8.1f %7.1f %8.1f  %s' % (st, e.time_range.elapsed_us(), st + e.time_range.elapsed_us(), e.name[:100]))
        print('# step span: %.1f us' % (max(e.time_range.end for e in last) - t0))
    if getattr(model, 'px', None) is not None:
        model.px.check()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
