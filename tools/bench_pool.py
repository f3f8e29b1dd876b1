#!/usr/bin/env python
"""Kernel-level sweep of the embedding pair on the C2 workload (user side, 4096 bags):
arx_pool_fwd and arx_pool_bwd_apply timed alone with CUDA events, L2 flushed between launches,
for every launch-shape setting of arx_set_tuning.  Prints one JSON line per setting with the
achieved GB/s on the conservative (unique-row) and nominal byte counts of SURVEY 8(d).

    python tools/bench_pool.py [--reps 20] [--mb 4096]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--reps', type=int, default=20)
    ap.add_argument('--mb', type=int, default=4096)
    ap.add_argument('--dim', type=int, default=128)
    ap.add_argument('--n', type=int, default=1000000)
    a = ap.parse_args()
    import arecsys_b200  # noqa: F401
    from arecsys_b200 import _lib
    from arecsys_b200._lib import call, POOL_MEAN, OPT_ADAGRAD
    from arecsys_b200.attributes.embed_attribute import EmbeddingAttribute
    from arecsys_b200.utils import synthetic
    from bench import algorithmic_bytes, peaks
    dev = torch.device('cuda', 0)
    ua, ia, i2l, l2i = synthetic.make_dataset(a.n, 1000, 8, 100000, 12, 64, 1.05, seed=0)
    ua.set_model_size(a.dim); ia.set_model_size(a.dim)
    emb = EmbeddingAttribute(ua, ia, a.mb, None, 0, False, i2l, l2i, seed=1)
    lib = _lib.load()
    peak, _ = peaks()
    flush = torch.empty(64 << 20, dtype=torch.float32, device=dev)      # 256 MB > L2
    rng = np.random.default_rng(0)
    batches = [rng.integers(0, a.n, a.mb).astype(np.int32) for _ in range(a.reps)]
    ids_dev = [torch.from_numpy(b).to(dev) for b in batches]
    ab = algorithmic_bytes(ua, batches[0], a.dim, False)
    ts = emb.sets['user']
    a0, na = ts.attr_range()
    out = torch.empty((a.mb, a.dim), dtype=torch.float32, device=dev)
    dout = torch.randn((a.mb, a.dim), dtype=torch.float32, device=dev) * 1e-3
    max_rows = int(sum(ts.max_len[a0:a0 + na]))

    def timed(fn, prep=None):
        us = []
        for r in range(a.reps):
            if prep is not None:
                prep(r)
            flush.add_(1.0)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(r); e1.record()
            torch.cuda.synchronize()
            us.append(e0.elapsed_time(e1) * 1e3)
        us = sorted(us[2:])
        return float(np.median(us)), float(us[0])

    def fwd(r):
        call('arx_pool_fwd', ts.desc_ptr(a0), na, a.dim, ids_dev[r].data_ptr(), a.mb, out.data_ptr(), out.stride(0),
             POOL_MEAN, None, max_rows)

    for epb in (0, 7, 5, 4, 3, 2):
        assert lib.arx_set_tuning(b'flat_epb', epb) == 0
        med, best = timed(fwd)
        print(json.dumps({'kernel': 'pool_fwd_flat', 'flat_epb': epb, 'median_us': med, 'best_us': best,
                          'GBs_unique': ab['fwd_unique'] / med / 1e3, 'GBs_nominal': ab['fwd_nominal'] / med / 1e3,
                          'frac_unique': ab['fwd_unique'] / med / 1e3 / peak}))
    lib.arx_set_tuning(b'flat_epb', 0)

    plans = {}

    def prep(r):
        plans[0] = emb._build_plan(ts, [(a0, na, ids_dev[r], POOL_MEAN)], None)

    def apply(r):
        call('arx_pool_bwd_apply', ts.desc_ptr(0), ts.n_attr, a.dim, plans[0].c, dout.data_ptr(), dout.stride(0), None,
             0.1, None, OPT_ADAGRAD, None, None)

    for cps in (1, 2, 3, 4):
        assert lib.arx_set_tuning(b'apply_ctas_per_sm', cps) == 0
        med, best = timed(apply, prep)
        print(json.dumps({'kernel': 'pool_bwd_apply', 'ctas_per_sm': cps, 'median_us': med, 'best_us': best,
                          'GBs_unique': ab['bwd_unique'] / med / 1e3, 'GBs_nominal': ab['bwd_nominal'] / med / 1e3,
                          'frac_unique': ab['bwd_unique'] / med / 1e3 / peak}))
    lib.arx_set_tuning(b'apply_ctas_per_sm', 2)
    print(json.dumps({'batch': ab}))


if __name__ == '__main__':
    main()
