"""Shared builders for the parity tests: small seeded attribute stores + injected weights."""
import numpy as np

import arecsys_b200  # noqa: F401  (registers the package)
from arecsys_b200.attributes.attribute import Attributes
from arecsys_b200.utils import synthetic


def small_dataset(n_users=50, n_items=40, n_mulhot=2, vocab_m=30, mean_len=3, max_len=6, seed=0,
                  logit_size=None, dim=8):
    u, i, i2l, l2i = synthetic.make_dataset(n_users, n_items, n_mulhot, vocab_m, mean_len, max_len,
                                            seed=seed, logit_size=logit_size)
    u.set_model_size(dim)
    i.set_model_size(dim)
    return u, i, i2l, l2i


def random_params(u_attr, i_attr, dim, seed=1, scale=0.5, item_output=False, mlp_hidden=None):
    rng = np.random.default_rng(seed)
    p = {}

    def add(prefix, att, bias):
        for tag, n, V in (('cat', att.num_features_cat, att._embedding_classes_list_cat),
                          ('mulhot', att.num_features_mulhot, att._embedding_classes_list_mulhot)):
            for k in range(n):
                p['%sembed_%s_%d' % (prefix, tag, k)] = rng.uniform(-scale, scale, (V[k], dim)).astype(np.float32)
                if bias:
                    p['%s_bias_%s_%d' % (prefix, tag, k)] = rng.uniform(-scale, scale, (V[k], 1)).astype(np.float32)
    add('user', u_attr, False)
    add('item', i_attr, True)
    if item_output:
        add('item_output', i_attr, True)
    if mlp_hidden:
        p['w1'] = rng.uniform(-scale, scale, (dim, mlp_hidden)).astype(np.float32)
        p['b1'] = rng.uniform(-scale, scale, (mlp_hidden,)).astype(np.float32)
        p['w2'] = rng.uniform(-scale, scale, (mlp_hidden, dim)).astype(np.float32)
        p['b2'] = rng.uniform(-scale, scale, (dim,)).astype(np.float32)
    return p


def positives(users, items, n_users, rng, extra=3, n_items=None):
    """per-user positive item sets containing the batch targets plus a few random others."""
    pos = {}
    for u, i in zip(users, items):
        pos.setdefault(int(u), set()).add(int(i))
    for u in list(pos.keys()):
        for v in rng.integers(0, n_items, size=extra):
            pos[u].add(int(v))
    return {u: sorted(v) for u, v in pos.items()}


def kat_item_attributes():
    """one multi-hot attribute, vocab 5, 3 items + START."""
    att = Attributes(0, [], 1, [np.array([0, 2, 1, 3, 4, 4, 1])], [3], [np.array([0, 2, 3, 6, 7])],
                     [np.array([2, 1, 3, 1])], [], [5])
    att.set_model_size(2)
    att.set_target_prediction([], [np.array([0, 2, 1, 3, 4, 4])], [np.array([0, 0, 1, 2, 2, 2])],
                              [np.array([[2.], [1.], [3.]])])
    return att
