"""On-disk attribute store: every array of the `Attributes` containers, the interaction lists and the index maps as
plain `.npy` files + one JSON manifest, memory-mapped on load.

The reference caches a pickle of Python lists (`data_dir/data`, attributes/input_attribute.py:19-24,59-62), which neither
scales to 10^7 entities (a Python int per token) nor maps into memory.  Layout under `<data_dir>/store/`:

    manifest.json                     format version, counts, vocabulary sizes, list lengths
    {u,i}_cat_<f>.npy                 features_cat[f]            int32 [N+1]
    {u,i}_mul_<f>_values.npy          features_mulhot[f]         int32 [nnz+1]
    {u,i}_mul_<f>_starts.npy / _lengths.npy                      int32 [N+2] / [N+1]
    i_full_cat_<f>.npy, i_full_values_<f>.npy, i_full_segids_<f>.npy, i_full_lengths_<f>.npy   catalog-ordered copies
    data_tr.npy / data_va.npy         int64 [n, 3]   (user index, item index, timestamp)
    item_ind2logit.npy / logit_ind2item.npy          int64 [n, 2] pairs / int64 [V]
    user_index.json / item_index.json                [[raw id, entity index], ...] (ids keep their int / str type)

`save_store` / `load_store` are exact inverses for everything the hot path reads; `read_data` (input_attribute.py) writes
the store next to the first preprocessing run and loads it afterwards (the pickle is still read if only it exists).
"""
import json
import os

import numpy as np

from .attribute import Attributes

FORMAT = 1


def _save(d, name, a, dtype):
    np.save(os.path.join(d, name + '.npy'), np.ascontiguousarray(np.asarray(a, dtype=dtype)))


def _save_side(d, tag, att, meta):
    m = {'n_cat': att.num_features_cat, 'n_mul': att.num_features_mulhot,
         'v_cat': [int(v) for v in att._embedding_classes_list_cat],
         'v_mul': [int(v) for v in att._embedding_classes_list_mulhot],
         'mulhot_max_length': [int(v) for v in (att.mulhot_max_length or [])],
         'has_full': hasattr(att, 'full_cat_tr')}
    for f in range(att.num_features_cat):
        _save(d, '%s_cat_%d' % (tag, f), att.features_cat[f], np.int32)
    for f in range(att.num_features_mulhot):
        _save(d, '%s_mul_%d_values' % (tag, f), att.features_mulhot[f], np.int32)
        _save(d, '%s_mul_%d_starts' % (tag, f), att.mulhot_starts[f], np.int32)
        _save(d, '%s_mul_%d_lengths' % (tag, f), att.mulhot_lengths[f], np.int32)
    if m['has_full']:
        for f in range(att.num_features_cat):
            _save(d, '%s_full_cat_%d' % (tag, f), att.full_cat_tr[f], np.int32)
        for f in range(att.num_features_mulhot):
            _save(d, '%s_full_values_%d' % (tag, f), att.full_values_tr[f], np.int32)
            _save(d, '%s_full_segids_%d' % (tag, f), att.full_segids_tr[f], np.int32)
            _save(d, '%s_full_lengths_%d' % (tag, f), att.full_lengths_tr[f], np.float32)
    meta[tag] = m


def save_store(data_dir, data_tr, data_va, u_attr, i_attr, item_ind2logit_ind, logit_ind2item_ind, user_index,
               item_index):
    d = os.path.join(data_dir, 'store')
    os.makedirs(d, exist_ok=True)
    meta = {'format': FORMAT}
    _save_side(d, 'u', u_attr, meta)
    _save_side(d, 'i', i_attr, meta)
    for name, data in (('data_tr', data_tr), ('data_va', data_va)):
        arr = np.asarray([(x[0], x[1], x[2] if len(x) > 2 else 0) for x in data], dtype=np.int64).reshape(-1, 3)
        _save(d, name, arr, np.int64)
    i2l = np.asarray(sorted((int(k), int(v)) for k, v in item_ind2logit_ind.items()), dtype=np.int64).reshape(-1, 2)
    _save(d, 'item_ind2logit', i2l, np.int64)
    V = len(logit_ind2item_ind)
    _save(d, 'logit_ind2item', [logit_ind2item_ind[v] for v in range(V)], np.int64)
    for name, idx in (('user_index', user_index), ('item_index', item_index)):
        with open(os.path.join(d, name + '.json'), 'w') as f:
            json.dump([[k if isinstance(k, str) else (int(k) if float(k) == int(k) else float(k)), int(v)]
                       for k, v in idx.items()], f)
    tmp = os.path.join(d, 'manifest.json.tmp')
    with open(tmp, 'w') as f:
        json.dump(meta, f)
    os.replace(tmp, os.path.join(d, 'manifest.json'))          # the manifest appears last: a partial store is never loaded
    return d


def store_exists(data_dir):
    return os.path.isfile(os.path.join(data_dir, 'store', 'manifest.json'))


def _load_side(d, tag, m, mmap):
    mode = 'r' if mmap else None
    ld = lambda name: np.load(os.path.join(d, name + '.npy'), mmap_mode=mode)
    att = Attributes.__new__(Attributes)
    att.num_features_cat, att.num_features_mulhot = m['n_cat'], m['n_mul']
    att.features_cat = [ld('%s_cat_%d' % (tag, f)) for f in range(m['n_cat'])]
    att.features_mulhot = [ld('%s_mul_%d_values' % (tag, f)) for f in range(m['n_mul'])]
    att.mulhot_starts = [ld('%s_mul_%d_starts' % (tag, f)) for f in range(m['n_mul'])]
    att.mulhot_lengths = [ld('%s_mul_%d_lengths' % (tag, f)) for f in range(m['n_mul'])]
    att.mulhot_max_length = list(m['mulhot_max_length']) or None
    att._embedding_classes_list_cat, att._embedding_classes_list_mulhot = list(m['v_cat']), list(m['v_mul'])
    att._check()
    if m['has_full']:
        att.full_cat_tr = [ld('%s_full_cat_%d' % (tag, f)) for f in range(m['n_cat'])]
        att.full_values_tr = [ld('%s_full_values_%d' % (tag, f)) for f in range(m['n_mul'])]
        att.full_segids_tr = [ld('%s_full_segids_%d' % (tag, f)) for f in range(m['n_mul'])]
        att.full_lengths_tr = [ld('%s_full_lengths_%d' % (tag, f)) for f in range(m['n_mul'])]
    return att


def load_store(data_dir, mmap=True):
    """-> (data_tr, data_va, u_attr, i_attr, item_ind2logit_ind, logit_ind2item_ind, user_index, item_index) with the
    attribute arrays memory-mapped read-only (they go to HBM once; nothing is materialised as Python objects) and the
    interaction lists as lists of (user, item, timestamp) int tuples, as the runners index them."""
    d = os.path.join(data_dir, 'store')
    meta = json.load(open(os.path.join(d, 'manifest.json')))
    if meta.get('format') != FORMAT:
        raise ValueError('attribute store format %r, expected %r' % (meta.get('format'), FORMAT))
    u_attr, i_attr = _load_side(d, 'u', meta['u'], mmap), _load_side(d, 'i', meta['i'], mmap)
    tolist = lambda name: [tuple(r) for r in np.load(os.path.join(d, name + '.npy')).tolist()]
    data_tr, data_va = tolist('data_tr'), tolist('data_va')
    i2l = np.load(os.path.join(d, 'item_ind2logit.npy'))
    item_ind2logit_ind = {int(k): int(v) for k, v in i2l}
    l2i = np.load(os.path.join(d, 'logit_ind2item.npy'))
    logit_ind2item_ind = {v: int(l2i[v]) for v in range(len(l2i))}
    user_index = {k: v for k, v in json.load(open(os.path.join(d, 'user_index.json')))}
    item_index = {k: v for k, v in json.load(open(os.path.join(d, 'item_index.json')))}
    return data_tr, data_va, u_attr, i_attr, item_ind2logit_ind, logit_ind2item_ind, user_index, item_index
