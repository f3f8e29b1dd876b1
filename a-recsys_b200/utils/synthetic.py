"""Synthetic attribute stores and interaction streams.

HET layout per side: categorical attribute 0 = identity id (vocab N+2, token = 2 + entity
index), then `n_mulhot` multi-hot attributes with vocab `vocab_m`+2, bag length
clip(1 + Poisson(mean_len-1), 1, max_len), token ids 2 + (Zipf(a) mod vocab_m) (heavy head,
duplicates allowed).  Every array carries the trailing START pseudo-entity exactly like
utils/preprocess.py:198,223-226 of the reference (cat token 1, bag [1]).
"""
import numpy as np

from ..attributes.attribute import Attributes

UNK_ID, START_ID = 0, 1


def make_side(n_entities, n_mulhot=8, vocab_m=100000, mean_len=12, max_len=64, zipf_a=1.05,
              seed=0, with_id=True):
    rng = np.random.default_rng(seed)
    cats, vc = [], []
    if with_id:
        ids = np.arange(n_entities, dtype=np.int64) + 2
        cats.append(np.append(ids, START_ID).astype(np.int32))
        vc.append(n_entities + 2)
    vals, starts, lens, vm = [], [], [], []
    for _ in range(n_mulhot):
        l = np.clip(1 + rng.poisson(mean_len - 1, size=n_entities), 1, max_len).astype(np.int64)
        nnz = int(l.sum())
        tok = 2 + (rng.zipf(zipf_a, size=nnz) - 1) % vocab_m
        l = np.append(l, 1)
        tok = np.append(tok, START_ID)
        s = np.concatenate([[0], np.cumsum(l)])
        vals.append(tok.astype(np.int32))
        starts.append(s.astype(np.int32))
        lens.append(l.astype(np.int32))
        vm.append(vocab_m + 2)
    return Attributes(len(cats), cats, n_mulhot, vals, [int(max_len)] * n_mulhot, starts, lens, vc, vm)


def make_dataset(n_users, n_items, n_mulhot=8, vocab_m=100000, mean_len=12, max_len=64,
                 zipf_a=1.05, seed=0, logit_size=None):
    """Returns (u_attr, i_attr, item_ind2logit_ind, logit_ind2item_ind) with every item a logit
    (identity map) unless logit_size < n_items (then the first logit_size items)."""
    u_attr = make_side(n_users, n_mulhot, vocab_m, mean_len, max_len, zipf_a, seed)
    i_attr = make_side(n_items, n_mulhot, vocab_m, mean_len, max_len, zipf_a, seed + 1)
    V = n_items if logit_size is None else logit_size
    l2i = np.arange(V, dtype=np.int64)
    i_attr.set_target_prediction_from_map(l2i)
    return u_attr, i_attr, IdentityMap(V), l2i


class IdentityMap(object):
    """item index -> logit index for the first V items, dict-like without 10^6 Python ints."""

    def __init__(self, V):
        self.V = V

    def __contains__(self, k):
        return 0 <= int(k) < self.V

    def __getitem__(self, k):
        k = int(k)
        if not 0 <= k < self.V:
            raise KeyError(k)
        return k

    def __len__(self):
        return self.V

    def as_array(self, n_items):
        a = np.full(n_items + 1, -1, dtype=np.int32)
        a[:self.V] = np.arange(self.V, dtype=np.int32)
        return a


def make_interactions(n_users, n_items, n, seed=0, item_zipf=1.0):
    """users uniform; items ~ 1/rank^item_zipf over a random permutation of the catalog."""
    rng = np.random.default_rng(seed + 7)
    users = rng.integers(0, n_users, size=n, dtype=np.int64)
    p = 1.0 / np.power(np.arange(1, n_items + 1, dtype=np.float64), item_zipf)
    cdf = np.cumsum(p)
    cdf /= cdf[-1]
    ranks = np.searchsorted(cdf, rng.random(n), side='right')
    ranks = np.minimum(ranks, n_items - 1)
    perm = rng.permutation(n_items)
    return users.astype(np.int32), perm[ranks].astype(np.int32)
