"""Debug aid: persistent LSTM kernels vs the per-step path, error per time step."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import arecsys_b200  # noqa
from arecsys_b200.lstm.lstm_layer import LSTMLayer

def run(T, mb, H, seed=0):
    rng = np.random.default_rng(seed)
    d_in = H
    X = torch.tensor(rng.standard_normal((T, mb, d_in)).astype(np.float32), device='cuda')
    W = (rng.standard_normal((d_in + H, 4 * H)) * (1.0 / np.sqrt(d_in + H))).astype(np.float32)
    b = (rng.standard_normal(4 * H) * 0.1).astype(np.float32)
    dO = torch.tensor(rng.standard_normal((T, mb, H)).astype(np.float32), device='cuda')
    res = {}
    for seq in ('1', '0'):
        os.environ['ARX_LSTM_SEQ'] = seq
        layer = LSTMLayer(d_in, H, torch.device('cuda'), W=W, b=b)
        out = layer.forward(X, 1.0)
        Xd, G, Hs, Cs = layer._ctx[:4]
        gates = G.clone()
        dX = layer.backward(dO)
        torch.cuda.synchronize()
        res[seq] = (out.clone(), gates, Cs.clone(), G.clone(), dX.clone(), layer.dW.clone(), bool(layer._seq))
    a, r = res['1'], res['0']
    print('T %d mb %d H %d  seq used: %s / %s' % (T, mb, H, a[6], r[6]))
    for name, k in (('h', 0), ('gates', 1), ('c', 2), ('dZ', 3), ('dX', 4)):
        x, y = a[k], r[k]
        per_t = [(float((x[t] - y[t]).abs().max()), float(y[t].abs().max())) for t in range(x.shape[0])]
        bad = [(t, '%.2e' % e) for t, (e, m) in enumerate(per_t) if e > 3e-3 * max(m, 1e-6)]
        print('  %-5s max err %.3e  first bad steps: %s' % (name, max(e for e, _ in per_t), bad[:6]))
        if bad and name in ('h', 'dZ'):
            t = bad[0][0]
            d = (x[t] - y[t]).abs()
            rows = torch.nonzero(d.max(1).values > 3e-3 * float(y[t].abs().max())).flatten()
            cols = torch.nonzero(d.max(0).values > 3e-3 * float(y[t].abs().max())).flatten()
            print('        step %d: %d bad rows (first %s), %d bad cols (first %s)' % (
                t, rows.numel(), rows[:8].tolist(), cols.numel(), cols[:8].tolist()))
    print('  dW max rel err %.3e' % float((a[5] - r[5]).abs().max() / r[5].abs().max()))



def ref64(T, mb, H, seed=0):
    """fp64 restatement on the same tf32-rounded operands is not needed: report each path against exact fp64."""
    rng = np.random.default_rng(seed)
    d_in = H
    X = rng.standard_normal((T, mb, d_in)).astype(np.float32).astype(np.float64)
    W = (rng.standard_normal((d_in + H, 4 * H)) * (1.0 / np.sqrt(d_in + H))).astype(np.float32).astype(np.float64)
    b = (rng.standard_normal(4 * H) * 0.1).astype(np.float32).astype(np.float64)
    h = np.zeros((mb, H)); c = np.zeros((mb, H)); hs = []
    sig = lambda x: 1 / (1 + np.exp(-x))
    for t in range(T):
        z = np.concatenate([X[t], h], 1) @ W + b
        i, j, f, o = np.split(z, 4, 1)
        c = sig(f + 1) * c + sig(i) * np.tanh(j)
        h = sig(o) * np.tanh(c)
        hs.append(h)
    return np.stack(hs)


def vs64(T, mb, H, seed=0):
    want = ref64(T, mb, H, seed)
    rng = np.random.default_rng(seed)
    d_in = H
    X = torch.tensor(rng.standard_normal((T, mb, d_in)).astype(np.float32), device='cuda')
    W = (rng.standard_normal((d_in + H, 4 * H)) * (1.0 / np.sqrt(d_in + H))).astype(np.float32)
    b = (rng.standard_normal(4 * H) * 0.1).astype(np.float32)
    for seq in ('1', '0', '1', '0'):
        os.environ['ARX_LSTM_SEQ'] = seq
        layer = LSTMLayer(d_in, H, torch.device('cuda'), W=W, b=b)
        out = layer.forward(X, 1.0).cpu().numpy().astype(np.float64)
        e = np.abs(out - want)
        per_t = e.reshape(T, -1).max(1)
        t = int(np.argmax(per_t > 5e-3)) if (per_t > 5e-3).any() else -1
        loc = np.argwhere(e[t] > 5e-3)[:6].tolist() if t >= 0 else []
        print('  vs fp64: seq=%s max err %.3e first step over 5e-3: %d at %s; err by step %s' % (
            seq, e.max(), t, loc, np.round(per_t[:12], 4).tolist()))


def count_glitches(T, mb, H, runs=12):
    want = ref64(T, mb, H, 0)
    rng = np.random.default_rng(0)
    d_in = H
    X = torch.tensor(rng.standard_normal((T, mb, d_in)).astype(np.float32), device='cuda')
    W = (rng.standard_normal((d_in + H, 4 * H)) * (1.0 / np.sqrt(d_in + H))).astype(np.float32)
    b = (rng.standard_normal(4 * H) * 0.1).astype(np.float32)
    os.environ['ARX_LSTM_SEQ'] = '1'
    bad = 0
    where = []
    for _ in range(runs):
        if os.environ.get('INTERLEAVE', '1') == '1':
            os.environ['ARX_LSTM_SEQ'] = '0'
            LSTMLayer(d_in, H, torch.device('cuda'), W=W, b=b).forward(X, 1.0).cpu()
            os.environ['ARX_LSTM_SEQ'] = '1'
        layer = LSTMLayer(d_in, H, torch.device('cuda'), W=W, b=b)
        out = layer.forward(X, 1.0).cpu().numpy().astype(np.float64)
        e = np.abs(out - want)
        if e.max() > 5e-3:
            bad += 1
            t = int(np.argmax(e.reshape(T, -1).max(1) > 5e-3))
            loc = np.argwhere(e[t] > 5e-3)
            where.append((t, loc[:, 0].min(), loc[:, 0].max(), loc[:, 1].min(), loc[:, 1].max(), len(loc)))
    print('T %d mb %d H %d dbg=%s: %d / %d runs glitched; (step, row range, col range, count): %s' % (
        T, mb, H, os.environ.get('ARX_LSTM_DBG', '0'), bad, runs, where[:8]))


if __name__ == '__main__':
    mode = sys.argv[1] if len(sys.argv) > 1 else 'count'
    if mode == 'full':
        for cfg in ((4, 130, 64), (50, 512, 64), (9, 300, 128), (50, 4096, 128), (20, 260, 32)):
            run(*cfg)
        print('--- against fp64')
        vs64(50, 512, 64)
        vs64(50, 512, 64, seed=3)
    elif mode == 'bwdfirst':
        run(9, 300, 128)
        count_glitches(50, 512, 64)
    elif mode == 'bigfirst':
        os.environ['ARX_LSTM_SEQ'] = '1'
        rng = np.random.default_rng(0)
        X = torch.tensor(rng.standard_normal((50, 4096, 128)).astype(np.float32), device='cuda')
        W = (rng.standard_normal((256, 512)) * 0.06).astype(np.float32)
        LSTMLayer(128, 128, torch.device('cuda'), W=W, b=np.zeros(512, np.float32)).forward(X, 1.0).cpu()
        del X
        count_glitches(50, 512, 64)
    else:
        count_glitches(50, 512, 64)
        count_glitches(50, 512, 128)
