"""Device-side batchers: the interaction sampler of the HMF runner, the CBOW sliding-window batcher
and the LSTM pad/bucket batcher produce their batches as device tensors through libarx_b200.so
(arx_gather_pairs / arx_cbow_window_batch / arx_lstm_pad_batch), ready for model.step() / replay_step().

What decides WHICH examples enter a batch stays bit-compatible with the reference's own functions where its RNG allows:
  * HMF `random`  : LatentProductModel.get_batch (hmf/hmf_model.py:230-241) draws mb x random.choice(data).  PyRandomStream
                    reproduces Python's Mersenne-Twister stream of those draws vectorised in NumPy (same generator state in,
                    same indices out, same state left behind) — 4096 interpreter calls per step become three array ops.
  * HMF `permute` : get_permuted_batch (:243-260): np.random.permutation per sweep, identical calls.
  * LSTM          : SeqModel.get_batch (lstm/seqModel.py:356-404): random.choice per slot (same stream) or the
                    deterministic window of the evaluation sweep; padding / START shifting on the device.
  * CBOW          : get_next_cbow (word2vec/data_iterator.py:108-169): the (user, target) stream is the reference's; the
                    context draws come from the device generator (same distribution, not the same stream).
"""
import random

import numpy as np
import torch

from .. import _lib
from .._lib import call


class PyRandomStream(object):
    """random.randrange(n) / random.choice(seq) draws of Python's global generator, vectorised.

    CPython: choice(seq) = seq[_randbelow(len(seq))], _randbelow(n): k = n.bit_length(); r = getrandbits(k) until r < n;
    getrandbits(k <= 32) = genrand_uint32() >> (32 - k).  NumPy's legacy RandomState is the same MT19937: its raw 32-bit
    words are Python's.  draw() takes the state from `random`, produces the indices, and puts the advanced state back."""

    @staticmethod
    def _to_numpy():
        ver, st, _gauss = random.getstate()
        rs = np.random.RandomState()
        rs.set_state(('MT19937', np.asarray(st[:-1], dtype=np.uint32), int(st[-1])))
        return rs, ver

    @staticmethod
    def _to_python(rs, ver):
        _, keys, pos = rs.get_state()[:3]
        random.setstate((ver, tuple(int(x) for x in keys) + (int(pos),), None))

    @classmethod
    def draw(cls, n, count):
        """`count` draws of random.randrange(n) (== the index random.choice takes for a sequence of length n)."""
        assert 0 < n < (1 << 32)
        k = int(n).bit_length()
        rs, ver = cls._to_numpy()
        start = rs.get_state()
        m = int(count * (1 << k) / n * 1.1) + 64
        while True:
            words = rs.randint(0, 1 << 32, size=m, dtype=np.uint32)
            r = words >> np.uint32(32 - k)
            ok = np.nonzero(r < n)[0]
            if len(ok) >= count:
                break
            rs.set_state(start)
            m *= 2
        used = int(ok[count - 1]) + 1
        rs.set_state(start)                       # consume exactly the words Python would have consumed
        if used:
            rs.randint(0, 1 << 32, size=used, dtype=np.uint32)
        cls._to_python(rs, ver)
        return r[ok[:count]].astype(np.int64)


class DeviceInteractionSampler(object):
    """HMF batches on the device: the training interactions live in HBM as two int32 arrays; a step's batch is a 32 KB
    index upload + one gather kernel.  `sample_type` as run_hmf.py's flag: 'random' | 'permute'."""

    def __init__(self, data, batch_size, device, sample_type='random'):
        arr = np.asarray([(d[0], d[1]) for d in data], dtype=np.int32) if not isinstance(data, np.ndarray) else data
        self.n = int(arr.shape[0])
        self.users = torch.from_numpy(np.ascontiguousarray(arr[:, 0])).to(device)
        self.items = torch.from_numpy(np.ascontiguousarray(arr[:, 1])).to(device)
        self.batch_size, self.device, self.sample_type = batch_size, device, sample_type
        self._idx_host = torch.empty(batch_size, dtype=torch.int64).pin_memory()
        self._idx_dev = torch.empty(batch_size, dtype=torch.int64, device=device)
        self.out_users = torch.empty(batch_size, dtype=torch.int32, device=device)
        self.out_items = torch.empty(batch_size, dtype=torch.int32, device=device)
        self.train_permutation, self.start_index = None, None

    def indices(self):
        mb = self.batch_size
        if self.sample_type == 'random':
            return PyRandomStream.draw(self.n, mb)                                   # hmf_model.py:230-241
        if self.train_permutation is None or self.start_index + mb >= self.n:        # :243-260
            self.start_index = 0
            self.train_permutation = np.random.permutation(self.n)
        idx = self.train_permutation[self.start_index:self.start_index + mb]
        self.start_index += mb
        return idx.astype(np.int64)

    def next(self):
        """(users, items): int32 device tensors [batch_size] (the same buffers every call)."""
        self._idx_host.copy_(torch.from_numpy(self.indices()))
        self._idx_dev.copy_(self._idx_host, non_blocking=True)
        call('arx_gather_pairs', self.users.data_ptr(), self.items.data_ptr(), self._idx_dev.data_ptr(), self.batch_size,
             self.out_users.data_ptr(), self.out_items.data_ptr())
        return self.out_users, self.out_items


class DeviceCbowBatcher(object):
    """word2vec/data_iterator.py:108-169 on the device.  seq: [(user, item)] stream with a PAD event before each user
    (word2vec/run_w2v.py:118-131)."""

    def __init__(self, seq, end_ind, batch_size, n_skips, window, device, seed=0):
        seq = np.asarray(seq, dtype=np.int64).reshape(-1, 2)
        items = seq[:, 1]
        self.l_seq = len(seq)
        is_pad = items == end_ind
        idx = np.arange(self.l_seq)
        last_pad = np.maximum.accumulate(np.where(is_pad, idx, -1))
        u_seq_len = np.minimum(idx - last_pad, window)
        targets = np.nonzero(~is_pad)[0]
        t = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a.astype(dt))).to(device)
        self.users, self.items = t(seq[:, 0], np.int32), t(items, np.int32)
        self.u_seq_len, self.targets = t(u_seq_len, np.int32), t(targets, np.int64)
        self.n_targets = len(targets)
        self.cursor = int(np.searchsorted(targets, window))           # first centre = stream position `window`
        self.mb, self.ni, self.window = batch_size, n_skips, window
        self.rng = torch.tensor([seed, 0], dtype=torch.int64, device=device)
        self.out_users = torch.empty(batch_size, dtype=torch.int32, device=device)
        self.out_inputs = torch.empty((n_skips, batch_size), dtype=torch.int32, device=device)
        self.out_targets = torch.empty(batch_size, dtype=torch.int32, device=device)

    def next(self):
        """(users [mb], inputs [ni, mb], targets [mb]) int32 device tensors."""
        call('arx_cbow_window_batch', self.users.data_ptr(), self.items.data_ptr(), self.u_seq_len.data_ptr(),
             self.targets.data_ptr(), self.l_seq, self.n_targets, self.cursor, self.mb, self.ni, self.window,
             self.rng.data_ptr(), self.out_users.data_ptr(), self.out_inputs.data_ptr(), self.out_targets.data_ptr())
        self.cursor = (self.cursor + self.mb) % self.n_targets
        return self.out_users, self.out_inputs, self.out_targets


class DeviceSeqBatcher(object):
    """SeqModel.get_batch (lstm/seqModel.py:356-404) on the device: every bucket's sequences as one CSR in HBM; a batch is
    mb sequence indices (random.choice per slot, or a window of the sweep) + one padding kernel."""

    def __init__(self, data_set, buckets, batch_size, start_id, device, user_pad_id=0):
        self.buckets, self.mb, self.device = buckets, batch_size, device
        self.start_id, self.pad_id, self.user_pad = start_id, start_id, user_pad_id
        self.csr = []
        for b, seqs in enumerate(data_set):
            lens = np.asarray([len(s[1]) for s in seqs], dtype=np.int64)
            ptr = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
            flat = np.asarray([v for s in seqs for v in s[1]], dtype=np.int32) if len(seqs) else np.zeros(0, np.int32)
            users = np.asarray([s[0] for s in seqs], dtype=np.int32)
            t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(device)
            self.csr.append((t(ptr), t(np.concatenate([flat, [0]]).astype(np.int32)), t(np.concatenate([users, [0]]).astype(np.int32)),
                             len(seqs)))
        self._sel_host = torch.empty(batch_size, dtype=torch.int64).pin_memory()
        self._sel_dev = torch.empty(batch_size, dtype=torch.int64, device=device)

    def next(self, bucket_id, start_id=None):
        """(users [mb], inputs [T, mb], targets [T, mb], weights [T, mb], finished) — device tensors; the selection follows
        get_batch: random.choice per slot when start_id is None, else sequences start_id .. start_id + mb - 1."""
        ptr, flat, users, n = self.csr[bucket_id]
        T, mb = self.buckets[bucket_id], self.mb
        if start_id is None:
            sel = PyRandomStream.draw(n, mb)
        else:
            sel = start_id + np.arange(mb, dtype=np.int64)
            sel[sel >= n] = -1
        self._sel_host.copy_(torch.from_numpy(sel))
        self._sel_dev.copy_(self._sel_host, non_blocking=True)
        ou = torch.empty(mb, dtype=torch.int32, device=self.device)
        oi = torch.empty((T, mb), dtype=torch.int32, device=self.device)
        ot = torch.empty((T, mb), dtype=torch.int32, device=self.device)
        ow = torch.empty((T, mb), dtype=torch.float32, device=self.device)
        call('arx_lstm_pad_batch', ptr.data_ptr(), flat.data_ptr(), users.data_ptr(), self._sel_dev.data_ptr(), mb, T,
             self.start_id, self.pad_id, self.user_pad, ou.data_ptr(), oi.data_ptr(), ot.data_ptr(), ow.data_ptr())
        finished = start_id is not None and start_id + mb >= n
        return ou, oi, ot, ow, finished
