"""Raw-data loading (reference: utils/load_data.py:5-96), Python 3 / pandas.

Input format (README.md:28-35 of the reference): TSV with a header row.
  u.csv / i.csv        : column 0 = id, remaining columns = attribute values; multi-hot values are
                         comma-joined tokens; i.csv holds Latin-1 bytes -> read as latin-1
  u_attr.csv/i_attr.csv: optional; header + one row of type codes (0 categorical, 1 multi-hot,
                         >1 ignored); absent => all categorical (load_data.py:42-43)
  obs_{tr,va,te}.csv   : user id, item id, optional timestamp (zeros appended if absent,
                         load_data.py:66-69), further columns ignored
Entity index = row number in u.csv / i.csv.
"""
from os.path import join, isfile

import numpy as np
import pandas as pd


def build_index(values):
    """id (column 0) -> row number (load_data.py:5-14)."""
    return {v: k for k, v in enumerate(values[:, 0].tolist())}


def load_csv(filename, indexing=True, sep='\t', header=0):
    if not isfile(filename):
        return ([], None, None) if indexing else ([], None)
    data = pd.read_csv(filename, delimiter=sep, header=header, encoding='latin-1', keep_default_na=False)
    columns = list(data.columns)
    values = data.values.astype(object)
    if indexing:
        return values, columns, build_index(values)
    return values, columns


def file_check(filename):
    if not isfile(filename):
        print("Error: user file {} does not exit!".format(filename))
        exit(1)


def _load_entities(data_dir, stem):
    filename = join(data_dir, stem + '.csv')
    file_check(filename)
    values, attr_names, index = load_csv(filename)
    tfile = join(data_dir, stem + '_attr.csv')
    if isfile(tfile):
        vals, _ = load_csv(tfile, False)
        attr_types = [int(v) for v in vals.flatten().tolist()]
    else:
        attr_types = [0] * len(attr_names)
    return values, (attr_names, attr_types), index


def load_users(data_dir, sep='\t'):
    return _load_entities(data_dir, 'u')


def load_items(data_dir, sep='\t'):
    return _load_entities(data_dir, 'i')


def load_interactions(data_dir, sep='\t'):
    ints, names = [], []
    for s in ['tr.csv', 'va.csv', 'te.csv']:
        interact, name = load_csv(join(data_dir, 'obs_' + s), False)
        assert interact.shape[1] >= 2
        if interact.shape[1] == 2:
            interact = np.append(interact, np.zeros((interact.shape[0], 1), dtype=int), 1)
        ints.append(interact)
        names.append(name)
    return ints, names[0]


def load_raw_data(data_dir, _submit=0):
    """load_data.py:73-96.  data_tr / data_va are lists of (user index, item index, time)."""
    users, u_attr, user_index = load_users(data_dir)
    items, i_attr, item_index = load_items(data_dir)
    ints, _ = load_interactions(data_dir)
    mapped = []
    for v in ints:
        u = np.fromiter((user_index[x] for x in v[:, 0].tolist()), dtype=np.int64, count=len(v))
        i = np.fromiter((item_index[x] for x in v[:, 1].tolist()), dtype=np.int64, count=len(v))
        t = np.asarray(v[:, 2].tolist())
        mapped.append((u, i, t))
    tr, va, te = mapped
    if _submit == 1:                                   # --test True: train on tr+va, validate on te
        tr = tuple(np.concatenate([a, b]) for a, b in zip(tr, va))
        va = te
    data_tr = list(zip(tr[0].tolist(), tr[1].tolist(), tr[2].tolist()))
    data_va = list(zip(va[0].tolist(), va[1].tolist(), va[2].tolist()))
    return users, items, data_tr, data_va, u_attr, i_attr, user_index, item_index
