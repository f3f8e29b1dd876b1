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
    """Draw n distinct items with P(order) as in np.random.choice(population, n, False, p)."""

    def __init__(self, item_population, p_item, device, seed=0):
        self.population = torch.as_tensor(np.asarray(item_population, dtype=np.int32)).to(device)
        self.logp = torch.log(torch.as_tensor(np.asarray(p_item, dtype=np.float64)).to(device)).float()
        self.gen = torch.Generator(device=device)
        self.gen.manual_seed(seed)

    def sample(self, n):
        u = torch.rand(self.logp.shape, generator=self.gen, device=self.logp.device).clamp_(1e-20, 1.0)
        keys = self.logp - torch.log(-torch.log(u))
        idx = torch.topk(keys, n).indices
        return self.population[idx].contiguous()
