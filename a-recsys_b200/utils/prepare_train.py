"""Sampling helpers of the training loop (reference: utils/prepare_train.py:7-57), vectorised.

sample_items / item_frequency / positive_items keep the reference's names, arguments and return
values; DeviceItemSampler draws the sampled pool on the GPU (Gumbel top-k == sequential sampling
without replacement with probabilities p, i.e. what np.random.choice(replace=False, p=p) does)
because np.random.choice over 10^6-10^7 items costs more than a whole training step.
"""
import numpy as np
import torch


def sample_items(items, n, p=None, replace=False):
    """utils/prepare_train.py:7-17."""
    if p is not None and len(p):
        item_sampled = np.random.choice(items, n, replace=replace, p=p)
    else:
        item_sampled = np.random.choice(items, n, replace=replace)
    item_sampled_id2idx = {int(item): i for i, item in enumerate(item_sampled)}
    return item_sampled, item_sampled_id2idx


def item_frequency(data_tr, power):
    """utils/prepare_train.py:19-35: population of training items and p ~ (count/total)^power."""
    items = np.fromiter((d[1] for d in data_tr), dtype=np.int64, count=len(data_tr))
    item_population, counts = np.unique(items, return_counts=True)
    p = np.power(counts / float(counts.sum()), power)
    p = p / p.sum()
    return [int(v) for v in item_population], [float(v) for v in p]


def positive_items(data_tr, data_va):
    """utils/prepare_train.py:37-57: user -> list of distinct positive items (train, validation)."""
    def build(data):
        hist = {}
        for u, i, _ in data:
            hist.setdefault(u, set()).add(i)
        return {u: list(v) for u, v in hist.items()}
    return build(data_tr), build(data_va)


def positives_csr(users, items, n_users):
    """Per-user CSR (ptr, items) of distinct positives from parallel arrays — the vectorised form
    of positive_items() for large synthetic streams (EmbeddingAttribute.prepare_warp accepts it)."""
    users = np.asarray(users, dtype=np.int64)
    items = np.asarray(items, dtype=np.int64)
    key = np.unique(users * (int(items.max()) + 1) + items)
    u = key // (int(items.max()) + 1)
    it = key % (int(items.max()) + 1)
    ptr = np.zeros(n_users + 1, dtype=np.int64)
    np.add.at(ptr, u + 1, 1)
    return np.cumsum(ptr).astype(np.int32), it.astype(np.int32)


class DeviceItemSampler(object):
    """Draw n distinct items with P(order) as in np.random.choice(population, n, False, p) (utils/prepare_train.py:7-17):
    Gumbel-top-k — key_i = log p_i - log(-log U_i) with U from the library's counter-based generator (arx_gumbel_keys), the
    n largest keys selected by arx_topk_rows (radix select, n <= 1024).  Everything runs in libarx_b200.so on the current
    stream: O(V) device work per refresh instead of np.random.choice's O(V) host work + a 10^6-10^7 element upload."""

    CHUNKS = 128          # two-stage selection for large populations: top-n per chunk, then top-n of the candidates

    def __init__(self, item_population, p_item, device, seed=0):
        from .. import _lib
        self._lib = _lib
        self.population = torch.as_tensor(np.asarray(item_population, dtype=np.int32)).to(device)
        logp = torch.log(torch.as_tensor(np.asarray(p_item, dtype=np.float64))).float()
        self.V = logp.numel()
        self.C = self.CHUNKS if self.V >= (1 << 16) else 1
        self.chunk = (self.V + self.C - 1) // self.C
        pad = self.chunk * self.C - self.V
        # padding entries have probability 0 (log p = -inf): never selected
        self.logp = torch.cat([logp, torch.full((pad,), float('-inf'))]).to(device).contiguous()
        self.keys = torch.empty_like(self.logp)
        self.rng = torch.tensor([seed, 0], dtype=torch.int64, device=device)

    def sample(self, n):
        if n > 1024 or n > self.V:
            raise ValueError('DeviceItemSampler draws at most min(1024, population) items per call')
        call, dev = self._lib.call, self.logp.device
        call('arx_gumbel_keys', self.logp.data_ptr(), self.logp.numel(), self.rng.data_ptr(), self.keys.data_ptr())
        if self.C == 1:
            idx = torch.empty((1, n), dtype=torch.int32, device=dev)
            call('arx_topk_rows', self.keys.data_ptr(), 1, self.V, self.V, n, idx.data_ptr(), None)
            return self.population[idx[0].long()].contiguous()
        k1 = min(n, self.chunk)
        idx1 = torch.empty((self.C, k1), dtype=torch.int32, device=dev)
        val1 = torch.empty((self.C, k1), dtype=torch.float32, device=dev)
        call('arx_topk_rows', self.keys.data_ptr(), self.C, self.chunk, self.chunk, k1, idx1.data_ptr(), val1.data_ptr())
        idx2 = torch.empty((1, n), dtype=torch.int32, device=dev)
        call('arx_topk_rows', val1.data_ptr(), 1, self.C * k1, self.C * k1, n, idx2.data_ptr(), None)
        j = idx2[0].long()
        glob = (j // k1) * self.chunk + idx1.reshape(-1)[j].long()
        return self.population[glob].contiguous()
