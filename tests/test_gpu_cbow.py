"""GPU parity of the CBOW tower against oracle/torch_cpu_ref.py::TorchRefCbow, the
window batcher's invariants, and word2vec/run_w2v.py driven like examples/run_w2v.sh."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from helpers import small_dataset, random_params, positives
from test_gpu_runner_hmf import _write_dataset, ROOT


def test_cbow_batcher_invariants():
    import arecsys_b200  # noqa: F401
    from arecsys_b200.word2vec.data_iterator import DataIterator
    PAD = 99
    seq = []
    for u in range(6):
        seq.append((u, PAD))
        seq.extend((u, 10 * u + k) for k in range(1 + u))
    np.random.seed(0)
    it = DataIterator(seq, PAD, 8, 2, 3, False).get_next_cbow()
    items = np.array([s[1] for s in seq]); users_all = np.array([s[0] for s in seq])
    seen = set()
    for _ in range(20):
        users, inputs, outputs = next(it)
        assert len(inputs) == 2 and all(len(x) == 8 for x in inputs) and len(users) == len(outputs) == 8
        assert (outputs != PAD).all()                                   # PAD events are never targets
        for b in range(8):
            p = int(np.nonzero((items == outputs[b]) & (users_all == users[b]))[0][0])
            window = set(items[(p - 3 + np.arange(3)) % len(seq)].tolist())
            assert inputs[0][b] in window and inputs[1][b] in window     # inputs come from the 3 previous stream events
            seen.add(int(outputs[b]))
    assert seen == set(items[items != PAD].tolist())                    # the sweep covers every event


@pytest.mark.gpu
@pytest.mark.parametrize('exact', [True, False])
@pytest.mark.parametrize('loss,sep,ni', [('ce', True, 2), ('warp', True, 3), ('bbpr', False, 1), ('mw', True, 2), ('ce', False, 0)])
def test_cbow_training_steps_match_reference(cuda, loss, sep, ni, exact):
    import arecsys_b200  # noqa: F401
    from arecsys_b200 import _lib
    from arecsys_b200.word2vec.cbow_model import Model
    from oracle.torch_cpu_ref import TorchRefCbow
    dim, mb, n_users, n_items = 8, 16, 40, 30
    ns = 10 if loss == 'mw' else None
    ua, ia, _, l2i = small_dataset(n_users, n_items, 2, 15, 3, 5, 0, None, dim)
    params = random_params(ua, ia, dim, 1, scale=0.4, item_output=sep)
    l2i_d = {int(v): int(l2i[v]) for v in range(len(l2i))}
    i2l_d = {v: k for k, v in l2i_d.items()}
    model = Model(n_users, n_items, dim, mb, 0.5, 1.0, ua, ia, i2l_d, l2i_d, n_input_items=ni, loss_function=loss,
                  dropout=0.5, top_N_items=5, use_sep_item=sep, n_sampled=ns, params=params)
    ref = TorchRefCbow(ua, ia, params, l2i_d, i2l_d, loss=loss, keep_prob=0.5, learning_rate=0.5, n_sampled=ns,
                       dtype=torch.float64, size=dim, item_output=sep, ni=ni)
    rng = np.random.default_rng(3)
    ltol, ptol = (2e-4, 2e-3) if exact else (2e-3, 2e-2)
    _lib.exact_fp32 = exact
    try:
        nin = max(ni, 1)
        for it in range(3):
            users = rng.integers(0, n_users, mb); outs = rng.integers(0, n_items, mb)
            ins = [rng.integers(0, n_items + 1, mb) for _ in range(nin)]          # may include the PAD pseudo-item
            pos = positives(users, outs, n_users, rng, n_items=n_items)
            model.prepare_warp(pos, pos); ref.pos, ref.pos_eval = pos, pos
            sampled = [int(v) for v in rng.permutation(n_items)[:ns]] if (ns and it != 1) else None
            mask = np.floor(rng.random((mb, dim)) + 0.5).astype(np.float32)
            lg = model.step(None, users.tolist(), [x.tolist() for x in ins], outs.tolist(), sampled, None, loss=loss,
                            masks=[torch.tensor(mask, device='cuda')])
            lr_ = ref.step_cbow(users, ins, outs, item_sampled=sampled, mask=mask)
            assert abs(lg - lr_) <= ltol * max(1.0, abs(lr_)), (it, lg, lr_)
            for k, v in ref.p.items():
                got = model.att_emb.params[k].cpu().numpy(); want = v.detach().numpy()
                assert np.abs(got.reshape(want.shape) - want).max() <= ptol * max(1.0, np.abs(want).max()), (k, it)
        eg = model.step(None, users.tolist(), [x.tolist() for x in ins], outs.tolist(), forward_only=True, loss=loss)
        er = ref.step_cbow(users, ins, outs, forward_only=True)
        assert abs(eg - er) <= ltol * max(1.0, abs(er))
        rec = model.step(None, users.tolist(), [x.tolist() for x in ins], forward_only=True, recommend=True)
        assert rec.shape == (mb, 5)
    finally:
        _lib.exact_fp32 = False


@pytest.mark.gpu
def test_run_w2v_like_the_launcher_script(cuda, tmp_path):
    raw = str(tmp_path / 'dataset') + '/'
    _write_dataset(raw, n_users=120, n_items=80, n_tr=3000)
    train_dir = str(tmp_path / 'train' / 'cbow')
    os.makedirs(str(tmp_path / 'train')); os.makedirs(str(tmp_path / 'cache'))
    base = [sys.executable, 'run_w2v.py', '--model', 'cbow', '--dataset', 'ml1m', '--raw_data', raw, '--data_dir',
            str(tmp_path / 'cache' / 'ml'), '--train_dir', train_dir, '--item_vocab_size', '40', '--vocab_min_thresh', '1',
            '--steps_per_checkpoint', '20', '--loss', 'ce', '--learning_rate', '1', '--size', '16', '--n_epoch', '2',
            '--skip_window', '5', '--ni', '3', '--num_skips', '3', '--test', 'False', '--top_N_items', '30']
    r = subprocess.run(base + ['--recommend', 'False'], cwd=os.path.join(ROOT, 'word2vec'), capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    log = open(os.path.join(train_dir, 'log.txt')).read()
    assert 'perplexity' in log and '  dev: perplexity' in log and 'Saving best model...' in log
    assert os.path.isfile(os.path.join(train_dir, 'best.ckpt-0'))
    r = subprocess.run(base + ['--recommend', 'True'], cwd=os.path.join(ROOT, 'word2vec'), capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert 'METRIC_FORMAT (self):' in r.stdout
