"""Vocabulary building and tokenisation (reference: utils/preprocess.py:51-326), Python 3.

Same files and array layouts as the reference; deterministic where Python-2 dict order decided
(SURVEY Appendix B.3): tokens are ordered by descending count over the training interactions and
ties keep first-occurrence order (entity-index order, then position inside the entity's token
list); MIX 'uid*' tokens come first in first-occurrence order.
"""
from os import listdir
from os.path import join

import numpy as np

_UNK = "_UNK"
_START = "_START"
UNK_ID = 0
START_ID = 1
_START_VOCAB = [_UNK, _START]


def pickle_save(m, filename):
    import pickle
    with open(filename, 'wb') as f:
        pickle.dump(m, f, protocol=pickle.HIGHEST_PROTOCOL)


def initialize_vocabulary(vocabulary_path):
    """One token per line -> ({token: id}, [token]) (preprocess.py:23-49)."""
    with open(vocabulary_path, 'r', encoding='latin-1') as f:
        rev_vocab = [line.rstrip('\n') for line in f]
    return {x: y for (y, x) in enumerate(rev_vocab)}, rev_vocab


def _tokens(value):
    if isinstance(value, list):
        return value
    if not isinstance(value, str):
        value = str(value)
    return value.split(',')


def _interaction_counts(inds, n_entities):
    return np.bincount(np.asarray(inds, dtype=np.int64), minlength=n_entities)


def _write_vocab(data_dir, prefix, i, max_size, vocab_list):
    with open(join(data_dir, "%s_vocab%d_%d" % (prefix, i, max_size)), 'w', encoding='latin-1') as f:
        for w in vocab_list:
            f.write(str(w) + "\n")


def _ranked(counts, order):
    """tokens by descending count, ties in first-occurrence order (stable sort)."""
    return sorted(order, key=lambda t: -counts[t])


def create_dictionary(data_dir, inds, features, feature_types, feature_names, max_vocabulary_size=50000,
                      logits_size_tr=50000, threshold=2, prefix='user'):
    """HET: one vocabulary per attribute (preprocess.py:51-117).  An entity's tokens are counted once
    per training interaction it appears in."""
    num_uf = len(feature_names)
    assert len(feature_types) == num_uf
    mult = _interaction_counts(inds, len(features))
    minimum_occurance = []
    max_size = max_vocabulary_size
    for i in range(num_uf):
        if feature_types[i] > 1:
            continue
        counts, order = {}, []
        for u in np.nonzero(mult)[0]:
            toks = [features[u, i]] if feature_types[i] == 0 else _tokens(features[u, i])
            for t in toks:
                if t not in counts:
                    counts[t] = 0
                    order.append(t)
                counts[t] += int(mult[u])
        if prefix == 'item' and i == 0:
            max_size = logits_size_tr + len(_START_VOCAB)                     # :86-87
        else:
            max_size = max_vocabulary_size
        vocab_list = _START_VOCAB + [t for t in _ranked(counts, order) if counts[t] >= threshold]
        if len(vocab_list) > max_size:
            print("vocabulary {}_{} longer than max_vocabulary_size {}. Truncate the tail".format(
                prefix, len(vocab_list), max_size))
            vocab_list = vocab_list[:max_size]
        _write_vocab(data_dir, prefix, i, max_size, vocab_list)
        minimum_occurance.append(counts[vocab_list[-1]] if vocab_list[-1] in counts else 0)
    with open(join(data_dir, "%s_minimum_occurance_%d" % (prefix, max_size)), 'w') as f:
        f.write('\n'.join(str(v) for v in minimum_occurance))


def create_dictionary_mix(data_dir, inds, features, feature_types, feature_names, max_vocabulary_size=50000,
                          logits_size_tr=50000, threshold=2, prefix='user'):
    """MIX: one vocabulary over '<column><value>' tokens (preprocess.py:119-167); 'uid*' tokens first."""
    mult = _interaction_counts(inds, len(features))
    cu, ou, cv, ov = {}, [], {}, []
    for u in np.nonzero(mult)[0]:
        for t in _tokens(features[u, 0]):
            c, o = (cu, ou) if t.startswith('uid') else (cv, ov)
            if t not in c:
                c[t] = 0
                o.append(t)
            c[t] += int(mult[u])
    max_size = max_vocabulary_size
    vocab_list = _START_VOCAB + [t for t in ou if cu[t] >= threshold] + \
        [t for t in _ranked(cv, ov) if cv[t] >= threshold]
    if len(vocab_list) > max_size:
        print("vocabulary {}_{} longer than max_vocabulary_size {}. Truncate the tail".format(
            prefix, len(vocab_list), max_size))
        vocab_list = vocab_list[:max_size]
    _write_vocab(data_dir, prefix, 0, max_size, vocab_list)
    last = vocab_list[-1]
    with open(join(data_dir, "%s_minimum_occurance_%d" % (prefix, max_size)), 'w') as f:
        f.write(str(cv[last] if last in cv else cu.get(last, 0)))


def _vocab_for(data_dir, prefix, i):
    path = "%s_vocab%d_" % (prefix, i)
    paths = [f for f in listdir(data_dir) if f.startswith(path)]
    assert len(paths) == 1, 'more than one dictionaries found! delete unnecessary ones to fix this.'
    return initialize_vocabulary(join(data_dir, paths[0]))[0]


def tokenize_attribute_map(data_dir, features, feature_types, max_vocabulary_size, logits_size_tr=50000,
                           prefix='user'):
    """Entities -> Attributes arrays (preprocess.py:169-238): categorical token per entity (+ START),
    multi-hot CSR with UNK tokens dropped, empty bag -> [UNK], trailing START bag [1]."""
    features_cat, features_mulhot = [], []
    v_sizes_cat, v_sizes_mulhot = [], []
    mulhot_max_leng, mulhot_starts, mulhot_lengs = [], [], []
    N = len(features)
    for i in range(len(feature_types)):
        ut = feature_types[i]
        if ut > 1:
            continue
        vocab = _vocab_for(data_dir, prefix, i)
        col = features[:, i]
        if ut == 0:
            v_sizes_cat.append(len(vocab))
            uf = np.fromiter((vocab.get(str(v), UNK_ID) for v in col), dtype=np.int32, count=N)
            features_cat.append(np.append(uf, START_ID).astype(np.int32))
        else:
            v_sizes_mulhot.append(len(vocab))
            vals, lengs = [], []
            for n in range(N):
                val_ = [x for x in (vocab.get(str(v), UNK_ID) for v in _tokens(col[n])) if x != UNK_ID]
                if not val_:
                    val_ = [UNK_ID]
                vals.extend(val_)
                lengs.append(len(val_))
            vals.append(START_ID)
            lengs.append(1)
            lengs = np.asarray(lengs, dtype=np.int32)
            mulhot_max_leng.append(int(lengs[:-1].max()) if N else 0)
            mulhot_starts.append(np.concatenate([[0], np.cumsum(lengs)]).astype(np.int32))
            mulhot_lengs.append(lengs)
            features_mulhot.append(np.asarray(vals, dtype=np.int32))
    num_features_cat = sum(v == 0 for v in feature_types)
    num_features_mulhot = sum(v == 1 for v in feature_types)
    return (num_features_cat, features_cat, num_features_mulhot, features_mulhot, mulhot_max_leng,
            mulhot_starts, mulhot_lengs, v_sizes_cat, v_sizes_mulhot)
