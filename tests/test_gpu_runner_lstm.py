"""GPU test of lstm/run.py driven like examples/run_lstm.sh (incl. --steps_per_checkpoint, which the
runner does not define, and --test defaulting to True => data_dir + '_test')."""
import os
import subprocess
import sys

import pytest

from test_gpu_runner_hmf import _write_dataset, ROOT

pytestmark = pytest.mark.gpu


def test_run_lstm_like_the_launcher_script(cuda, tmp_path):
    raw = str(tmp_path / 'dataset') + '/'
    _write_dataset(raw, n_users=150, n_items=80, n_tr=4000)
    train_dir = str(tmp_path / 'train' / 'lstm_h16')
    os.makedirs(str(tmp_path / 'train')); os.makedirs(str(tmp_path / 'cache'))
    base = [sys.executable, 'run.py', '--dataset', 'ml1m', '--raw_data', raw, '--data_dir', str(tmp_path / 'cache' / 'ml'),
            '--train_dir', train_dir, '--item_vocab_size', '40', '--vocab_min_thresh', '1', '--steps_per_checkpoint', '5',
            '--loss', 'ce', '--learning_rate', '1', '--size', '16', '--batch_size', '32', '--L', '12', '--n_bucket', '3',
            '--topk', '20', '--max_steps', '60']
    r = subprocess.run(base + ['--recommend', 'False'], cwd=os.path.join(ROOT, 'lstm'), capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    log = open(os.path.join(train_dir, 'log.txt')).read()
    assert 'Speed:' in log and 'targets / sec in total' in log                  # the reference's throughput line
    assert 'dev: ppx:' in log and 'Saving best model....' in log and 'perplexity' in log
    assert os.path.isfile(os.path.join(train_dir, 'best.ckpt-0'))
    assert os.path.isdir(str(tmp_path / 'cache' / 'ml_test'))                   # --test defaults to True
    r = subprocess.run(base + ['--recommend', 'True'], cwd=os.path.join(ROOT, 'lstm'), capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert 'METRIC_FORMAT (self):' in r.stdout
    assert os.path.isfile(os.path.join(train_dir, 'top20_index.npy')) and os.path.isfile(os.path.join(train_dir, 'top20_value.npy'))
    for extra in (['--loss', 'warp'], ['--loss', 'mw', '--n_sampled', '16'], ['--use_concat', 'True', '--use_sep_item', 'True']):
        td = str(tmp_path / 'train' / '_'.join(extra).replace('-', ''))
        r = subprocess.run(base[:9] + [td] + base[10:] + extra + ['--max_steps', '20'], cwd=os.path.join(ROOT, 'lstm'),
                           capture_output=True, text=True, timeout=900)
        assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
