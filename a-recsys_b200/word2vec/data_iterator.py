"""CBOW window batcher (reference: word2vec/data_iterator.py:108-169), vectorised, and the skip-gram pair generator
(:60-106), kept as the reference's loop (it consumes np.random exactly like the reference: same seed, same pairs).

The training corpus is the concatenation of every user's time-ordered items with a PAD event
before each user (word2vec/run_w2v.py:118-131).  For every non-PAD event (u, o) the reference keeps
a deque of the previous `skip_window` stream events and draws `ni` of them (without replacement
once the user has >= ni earlier events in the window, with replacement before) as the inputs —
drawn from the stream window, so they may belong to the previous user or be the PAD pseudo-item,
exactly as in the reference.  Same distribution, NumPy instead of a Python deque per event.
"""
import numpy as np


class DataIterator(object):
    def __init__(self, seq, end_ind, batch_size, n_skips, window, sequence):
        seq = np.asarray(seq, dtype=np.int64).reshape(-1, 2)
        self.users, self.items = seq[:, 0], seq[:, 1]
        self.l_seq = len(seq)
        self.end_ind = end_ind
        self.batch_size = batch_size
        self.num_skips = n_skips
        self.skip_window = window
        self.sequence = sequence
        self.seq = [tuple(int(v) for v in r) for r in seq]
        self.index = 0
        if sequence:
            print('error: not implemented')
            exit(1)
        is_pad = self.items == end_ind
        # events since the user's PAD marker (1 for the first real event), capped at the window
        idx = np.arange(self.l_seq)
        last_pad = np.maximum.accumulate(np.where(is_pad, idx, -1))
        self.u_seq_len = np.minimum(idx - last_pad, window)
        self.targets = np.nonzero(~is_pad)[0]
        self.cursor = int(np.searchsorted(self.targets, window))      # first centre = stream position `window`

    def get_next_cbow(self):
        mb, ni, c = self.batch_size, self.num_skips, self.skip_window
        nt = len(self.targets)
        while True:
            sel = self.targets[(self.cursor + np.arange(mb)) % nt]
            self.cursor = (self.cursor + mb) % nt
            with_repl = self.u_seq_len[sel] < ni
            offs = np.argsort(np.random.random((mb, c)), axis=1)[:, :ni] if ni <= c else \
                np.random.randint(0, c, (mb, ni))
            offs_r = np.random.randint(0, c, (mb, ni))
            offs = np.where(with_repl[:, None], offs_r, offs)
            pos = (sel[:, None] - c + offs) % self.l_seq
            i_items = self.items[pos]                                   # [mb, ni]
            yield (self.users[sel].astype(np.int32), [i_items[:, k].astype(np.int32) for k in range(ni)],
                   self.items[sel].astype(np.int32))

    def get_next_sg(self):
        """word2vec/data_iterator.py:60-106: a window of span = 2 * skip_window + 1 stream events slides one event at a
        time; its OLDEST event is the input (centre = 0: "only predict future based on history"); num_skips attempts draw a
        later position of the window, and an attempt that lands on another user's event or on PAD yields no pair."""
        import collections
        seq, mb = self.seq, self.batch_size
        users = np.ndarray(shape=[mb], dtype=np.int32)
        i_items = np.ndarray(shape=[mb], dtype=np.int32)
        o_items = np.ndarray(shape=[mb], dtype=np.int32)
        span = 2 * self.skip_window + 1
        b = collections.deque(maxlen=span)
        center = 0
        for _ in range(span):
            b.append(seq[self.index])
            self.index = (self.index + 1) % self.l_seq
        while True:
            ind = 0
            while ind < mb:
                u, i_i = b[center]
                targets_to_avoid = [center]
                for _ in range(self.num_skips):
                    t = np.random.randint(0, span)
                    while t in targets_to_avoid:
                        t = np.random.randint(0, span)
                    o_i = b[t][1]
                    if b[t][0] != u or o_i == self.end_ind:
                        continue
                    targets_to_avoid.append(t)
                    users[ind], i_items[ind], o_items[ind] = u, i_i, o_i
                    ind += 1
                    if ind >= mb:
                        break
                b.append(seq[self.index])
                self.index = (self.index + 1) % self.l_seq
            yield users, [i_items], o_items
