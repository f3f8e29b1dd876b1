#!/usr/bin/env python
"""Secondary workloads of BASELINE.json (not the bench.py contract line): one JSON line each.

  python bench_extra.py --workload c3   # LSTM seq2seq dim 64, T 50, batch 512, 100k items, loss ce
  python bench_extra.py --workload c5   # CBOW dim 128, window 5 (ni = 3 inputs), batch 4096, loss mw
  python bench_extra.py --workload c4   # HMF + mw, 10 M items (1 GPU slice of config 4)

Same timing rules as bench.py: warm-up >= 3, CUDA events on the launching stream, ids resident in HBM,
tables far larger than L2.  Per-kernel durations come from a separate eager pass.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def timeline_table(_lib, tl, steps):
    agg = {}
    for name, tag, s0, s1 in tl:
        k = '%s:%s' % (name, tag) if name in ('arx_pool_fwd', 'arx_pool_bwd_apply') else name
        d = agg.setdefault(k, [0.0, 0])
        d[0] += s0.elapsed_time(s1)
        d[1] += 1
    return {k: {'ms_per_step': round(v[0] / steps, 4), 'launches_per_step': v[1] / steps,
                'avg_us': round(1e3 * v[0] / max(v[1], 1), 1)} for k, v in agg.items()}


def run(name, step_fn, units_per_step, unit, steps, warmup, cfg, flops_per_step=None):
    from arecsys_b200 import _lib
    for s in range(warmup):
        step_fn(s)
    torch.cuda.synchronize()
    l0 = _lib.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in range(warmup, warmup + steps):
        step_fn(s)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    launches = _lib.launch_count - l0
    _lib.timeline = []
    for s in range(warmup, warmup + min(steps, 5)):
        step_fn(s)
    torch.cuda.synchronize()
    tl, _lib.timeline = _lib.timeline, None
    out = {'workload': name, 'metric': unit, 'value': units_per_step * steps / (ms / 1e3), 'ms_per_step': ms / steps,
           'steps': steps, 'warmup': warmup, 'gpu_launches': launches, 'config': cfg,
           'per_kernel': timeline_table(_lib, tl, min(steps, 5))}
    if flops_per_step:
        out['dense_tflops'] = flops_per_step * steps / (ms / 1e3) / 1e12
    print(json.dumps(out))


def c3(a):
    """BASELINE config 3: LSTM seq2seq dim=64 seqlen=50 batch=512 synthetic sessions (run_lstm.sh path)."""
    import arecsys_b200  # noqa: F401
    from arecsys_b200.utils import synthetic
    from arecsys_b200.attributes.embed_attribute import EmbeddingAttribute
    from arecsys_b200.lstm.seqModel import SeqModel
    n_users, n_items, T, mb, d = a.n_users or 1000000, a.n_items or 100000, 50, a.mb or 512, a.dim or 64
    ua = synthetic.make_side(n_users, 0, 2, 1, 1, seed=0)
    ia = synthetic.make_side(n_items, 2, 10000, 4, 16, seed=1)
    ua.set_model_size(d); ia.set_model_size(d)
    START = n_items
    l2i = np.arange(n_items, dtype=np.int64)
    ia.set_target_prediction_from_map(l2i)

    class SeqMap(synthetic.IdentityMap):           # START / padding -> logit 0 (lstm/run.py:276-278)
        def __contains__(self, k):
            return 0 <= int(k) <= self.V

        def __getitem__(self, k):
            return 0 if int(k) == self.V else synthetic.IdentityMap.__getitem__(self, k)

        def as_array(self, n):
            arr = synthetic.IdentityMap.as_array(self, n)
            arr[self.V] = 0
            return arr
    i2l = SeqMap(n_items)
    emb = EmbeddingAttribute(ua, ia, mb, None, T, False, i2l, l2i, seed=1)
    model = SeqModel([T], d, 1, 5.0, mb, 0.5, 0.83, emb, withAdagrad=True, dropoutRate=a.keep_prob, START_ID=START,
                     loss='ce', use_concat=False, no_user_id=True, topk_n=30, seed=2)
    rng = np.random.default_rng(0)
    nb = a.warmup + a.steps
    batches = []
    for _ in range(min(nb, 8)):
        lens = rng.integers(5, T + 1, mb)
        seq = (rng.zipf(1.2, (mb, T)) % n_items).astype(np.int32)
        valid = np.arange(T)[None, :] < lens[:, None]
        out = np.where(valid, seq, START).astype(np.int32)
        inp = np.concatenate([np.full((mb, 1), START, np.int32), out[:, :-1]], 1)
        inp = np.where(np.arange(T)[None, :] < lens[:, None], inp, START).astype(np.int32)
        users = rng.integers(0, n_users, mb).astype(np.int32)
        batches.append((users, np.ascontiguousarray(inp.T), np.ascontiguousarray(out.T),
                        np.ascontiguousarray(valid.T.astype(np.float32))))
    targets = float(np.mean([b[3].sum() for b in batches]))

    def step(s):
        u, i, o, w = batches[s % len(batches)]
        model.step(None, u, i, o, w, 0, sync=False)
    V = n_items
    flops = 3 * 2.0 * T * mb * d * V + 3 * 2.0 * T * mb * (2 * d) * (4 * d)
    run('C3: LSTM dim=%d T=%d batch=%d, %d users / %d items (id + 2 multi-hot, mean bag 4), loss=ce, keep_prob=%.2f'
        % (d, T, mb, n_users, n_items, a.keep_prob), step, targets, 'targets/s', a.steps, a.warmup,
        {'targets_per_step': targets, 'note': 'targets = sum of target weights (lstm/run.py:470)'}, flops)


def c5(a):
    """BASELINE config 5: CBOW (run_w2v.sh) dim=128 window=5 -> ni=3 input items, synthetic event stream."""
    import arecsys_b200  # noqa: F401
    from arecsys_b200.utils import synthetic
    from arecsys_b200.utils.prepare_train import DeviceItemSampler, positives_csr
    from arecsys_b200.word2vec.cbow_model import Model
    n_users, n_items, mb, d, ni = a.n_users or 2000000, a.n_items or 1000000, a.mb or 4096, a.dim or 128, 3
    ua, ia, i2l, l2i = synthetic.make_dataset(n_users, n_items, 2, 100000, 12, 64, 1.05, seed=0)
    nb = a.warmup + a.steps
    users, items = synthetic.make_interactions(n_users, n_items, mb * nb * (ni + 1), seed=0)
    pop, counts = np.unique(items, return_counts=True)
    p = np.power(counts / counts.sum(), 0.5); p /= p.sum()
    model = Model(n_users, n_items, d, mb, a.lr, 1.0, ua, ia, i2l, l2i, n_input_items=ni, loss_function='mw',
                  dropout=a.keep_prob, top_N_items=30, use_sep_item=True, n_sampled=1024, seed=1)
    model.prepare_warp(positives_csr(users[:mb * nb], items[:mb * nb], ua.num_entities),
                       positives_csr(users[:mb * nb], items[:mb * nb], ua.num_entities))
    dev = torch.device('cuda:0')
    sampler = DeviceItemSampler(pop, p, dev, seed=0)
    u_all = users[:mb * nb].reshape(nb, mb)
    it = items.reshape(nb, ni + 1, mb)
    state = {'n': 0}

    def step(s):
        sampled = sampler.sample(1024) if state['n'] % 50 == 0 else None
        state['n'] += 1
        model.step(None, u_all[s % nb], it[s % nb, :ni], it[s % nb, ni], sampled, None, loss='mw', sync=False)
    run('C5: CBOW dim=%d ni=%d batch=%d, %d users / %d items (id + 2 multi-hot, mean bag 12), separate output tables, '
        'loss=mw n_sampled=1024, keep_prob=%.2f' % (d, ni, mb, n_users, n_items, a.keep_prob), step, mb,
        'events/s', a.steps, a.warmup, {})


def c4(a):
    """One-GPU slice of BASELINE config 4: HMF + mw, 10 M items, 1024 negatives (tables 10.2 GB + accumulators)."""
    import arecsys_b200  # noqa: F401
    from arecsys_b200.utils import synthetic
    from arecsys_b200.utils.prepare_train import DeviceItemSampler, positives_csr
    from arecsys_b200.hmf.hmf_model import LatentProductModel
    n_users, n_items, mb, d = a.n_users or 1000000, a.n_items or 10000000, a.mb or 4096, a.dim or 128
    ua, ia, i2l, l2i = synthetic.make_dataset(n_users, n_items, 8, 100000, 12, 64, 1.05, seed=0)
    nb = a.warmup + a.steps
    users, items = synthetic.make_interactions(n_users, n_items, mb * nb, seed=0)
    pop, counts = np.unique(items, return_counts=True)
    p = np.power(counts / counts.sum(), 0.5); p /= p.sum()
    model = LatentProductModel(n_users, n_items, d, 1, mb, a.lr, 1.0, ua, ia, i2l, l2i, loss_function='mw',
                               dropout=a.keep_prob, n_sampled=1024, seed=1)
    pos = positives_csr(users, items, ua.num_entities)
    model.prepare_warp(pos, pos)
    dev = torch.device('cuda:0')
    sampler = DeviceItemSampler(pop, p, dev, seed=0)
    u_dev = torch.from_numpy(users.reshape(nb, mb)).to(dev)
    i_dev = torch.from_numpy(items.reshape(nb, mb)).to(dev)
    state = {'n': 0, 'graph': False}

    def step(s):
        sampled = sampler.sample(1024) if state['n'] % 50 == 0 else None
        state['n'] += 1
        from arecsys_b200 import _lib
        if state['graph'] and _lib.timeline is None:
            if sampled is not None:
                model.att_emb.pass_sampled_items(sampled)
            return model.replay_step(u_dev[s % nb], i_dev[s % nb], sync=False)
        return model.step(None, u_dev[s % nb], i_dev[s % nb], None, sampled, None, loss='mw', sync=False)
    for s in range(3):
        step(s)
    model.capture_step(u_dev[0], i_dev[0], loss='mw')
    state['graph'] = True
    run('C4 (1 GPU): HMF dim=%d batch=%d, %d users / %d items, 8 multi-hot per side, loss=mw n_sampled=1024'
        % (d, mb, n_users, n_items), step, mb, 'interactions/s', a.steps, a.warmup, {'launch_mode': 'cuda graph replay'})


def lstm_cell(a):
    """The LSTM layer alone at the north-star shape (batch 4096, d = H = 128, T = 50): x-projection GEMM + the persistent
    recurrence kernels (arx_lstm_seq_fwd / _bwd) + the three large weight / input gradient contractions; forward and
    backward timed separately with CUDA events; `roofline` = dense tf32 FLOP/s of the whole layer step against half
    the measured bf16 peak, `recurrence` = the two persistent kernels alone (they are HBM-bound on the gate tensors:
    bytes = read x-projection + write gates + write h, c forward; read gates, c, dH + write dZ backward)."""
    import arecsys_b200  # noqa: F401
    from arecsys_b200 import _lib
    from arecsys_b200.lstm.lstm_layer import LSTMLayer
    T, mb, H = 50, a.mb or 4096, a.dim or 128
    d_in = H
    dev = torch.device('cuda:0')
    rng = np.random.default_rng(0)
    X = torch.tensor(rng.standard_normal((T, mb, d_in)).astype(np.float32), device=dev)
    dO = torch.tensor(rng.standard_normal((T, mb, H)).astype(np.float32), device=dev)
    W = (rng.standard_normal((d_in + H, 4 * H)) / np.sqrt(d_in + H)).astype(np.float32)
    layer = LSTMLayer(d_in, H, dev, W=W, b=np.zeros(4 * H, np.float32))
    res = {}
    for seq in ('1', '0'):
        os.environ['ARX_LSTM_SEQ'] = seq
        for _ in range(a.warmup):
            layer.forward(X, 1.0); layer.backward(dO)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        tf = tb = 0.0
        for _ in range(a.steps):
            ev[0].record(); layer.forward(X, 1.0); ev[1].record(); layer.backward(dO); ev[2].record()
            torch.cuda.synchronize()
            tf += ev[0].elapsed_time(ev[1]); tb += ev[1].elapsed_time(ev[2])
        res[seq] = (tf / a.steps, tb / a.steps)
        if seq == '1':
            _lib.timeline = []
            layer.forward(X, 1.0); layer.backward(dO)
            torch.cuda.synchronize()
            tl, _lib.timeline = _lib.timeline, None
            pk = timeline_table(_lib, tl, 1)
    os.environ['ARX_LSTM_SEQ'] = '1'
    peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json'))) if os.path.isfile(os.path.join(ROOT, 'MEASURED_PEAKS.json')) else {}
    tf32_peak = peaks.get('bf16_tflops_sustained', 1400.0) / 2
    hbm = peaks.get('hbm_gbs', 6650.0)
    fl_fwd = 2.0 * T * mb * (d_in + H) * 4 * H
    fl_bwd = 2.0 * fl_fwd                                  # dX + dW_x, dh + dW_h
    ms_f, ms_b = res['1']
    rec = {}
    for k, nbytes, fl in (('arx_lstm_seq_fwd', T * mb * (4 * H + 4 * H + 2 * H) * 4.0, 2.0 * T * mb * H * 4 * H),
                          ('arx_lstm_seq_bwd', T * mb * (4 * H + 2 * H + H + 4 * H) * 4.0, 2.0 * T * mb * H * 4 * H)):
        if k in pk:
            us = pk[k]['avg_us']
            rec[k] = {'us': us, 'us_per_time_step': us / T, 'hbm_gbs': nbytes / us / 1e3, 'hbm_frac': nbytes / us / 1e3 / hbm,
                      'tf32_tflops': fl / us / 1e6, 'tensor_frac_of_tf32_peak': fl / us / 1e6 / tf32_peak,
                      'ctas': (H // 32) * ((mb + 127) // 128)}
    out = {'workload': 'LSTM layer: T=%d batch=%d d_in=H=%d, forward + backward (persistent cluster kernels)' % (T, mb, H),
           'metric': 'targets/s (layer only)', 'value': T * mb / ((ms_f + ms_b) / 1e3), 'ms_fwd': ms_f, 'ms_bwd': ms_b,
           'ms_fwd_per_step_kernels': res['0'][0], 'ms_bwd_per_step_kernels': res['0'][1],
           'speedup_vs_per_step_path': (res['0'][0] + res['0'][1]) / (ms_f + ms_b), 'steps': a.steps, 'warmup': a.warmup,
           'roofline': {'bound': 'tensor', 'achieved': (fl_fwd + fl_bwd) / ((ms_f + ms_b) / 1e3) / 1e12, 'peak': tf32_peak,
                        'unit': 'TFLOP/s', 'frac': (fl_fwd + fl_bwd) / ((ms_f + ms_b) / 1e3) / 1e12 / tf32_peak,
                        'peak_source': 'half of MEASURED_PEAKS.json bf16_tflops_sustained (tf32 operands)', 'traffic': None},
           'recurrence': rec, 'per_kernel': pk}
    print(json.dumps(out))


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--workload', required=True, choices=['c3', 'c4', 'c5', 'lstm_cell'])
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--n_users', type=int, default=0)
    ap.add_argument('--n_items', type=int, default=0)
    ap.add_argument('--mb', type=int, default=0)
    ap.add_argument('--dim', type=int, default=0)
    ap.add_argument('--lr', type=float, default=0.1)
    ap.add_argument('--keep_prob', type=float, default=0.5)
    a = ap.parse_args()
    if not torch.cuda.is_available():
        raise SystemExit('bench_extra.py needs a CUDA device (no CPU fallback)')
    t0 = time.time()
    {'c3': c3, 'c4': c4, 'c5': c5, 'lstm_cell': lstm_cell}[a.workload](a)
    print('[bench_extra] %s done in %.0fs' % (a.workload, time.time() - t0), file=sys.stderr)
