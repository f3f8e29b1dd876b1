"""Attribute store of one entity type (users or items).

Same field names and meaning as the reference container (attributes/attribute.py:7-47)
so code written against it keeps working; arrays are NumPy int32 (CSR-like) instead of
Python lists so they upload to HBM without conversion:

  features_cat[f]    int32[N+1]   token id of categorical attribute f per entity; the last
                                   entry is the START/PAD pseudo-entity (utils/preprocess.py:198)
  features_mulhot[f] int32[nnz+1] flat token ids of multi-hot attribute f
  mulhot_starts[f]   int32[N+2]   bag offsets, mulhot_lengths[f] int32[N+1] bag sizes
  full_cat_tr / full_values_tr / full_segids_tr / full_lengths_tr: the catalog-ordered
  copies built by filter_cat / filter_mulhot (utils/preprocess.py:240-326).
"""
import numpy as np


class Attributes(object):
    def __init__(self, num_feature_cat=0, feature_cat=None, num_text_feat=0, feature_mulhot=None,
                 mulhot_max_length=None, mulhot_starts=None, mulhot_lengths=None,
                 v_sizes_cat=None, v_sizes_mulhot=None, embedding_size_list_cat=None):
        self.num_features_cat = num_feature_cat
        self.num_features_mulhot = num_text_feat
        self.features_cat = [np.asarray(a, dtype=np.int32) for a in (feature_cat or [])]
        self.features_mulhot = [np.asarray(a, dtype=np.int32) for a in (feature_mulhot or [])]
        self.mulhot_starts = [np.asarray(a, dtype=np.int32) for a in (mulhot_starts or [])]
        self.mulhot_lengths = [np.asarray(a, dtype=np.int32) for a in (mulhot_lengths or [])]
        self.mulhot_max_length = mulhot_max_length
        self._embedding_classes_list_cat = list(v_sizes_cat or [])
        self._embedding_classes_list_mulhot = list(v_sizes_mulhot or [])
        self._check()

    def _check(self):
        assert len(self.features_cat) == self.num_features_cat
        assert len(self.features_mulhot) == self.num_features_mulhot
        for i in range(self.num_features_mulhot):
            s, l, v = self.mulhot_starts[i], self.mulhot_lengths[i], self.features_mulhot[i]
            assert len(s) == len(l) + 1, 'mulhot_starts must have one more entry than mulhot_lengths'
            assert int(s[-1]) == len(v), 'CSR inconsistent: starts[-1] != len(values)'
            assert (l >= 1).all(), 'every bag holds at least one token (UNK if empty)'

    @property
    def num_entities(self):
        """N+1: real entities plus the trailing START/PAD pseudo-entity."""
        if self.num_features_cat:
            return len(self.features_cat[0])
        return len(self.mulhot_lengths[0])

    def set_model_size(self, sizes, opt=0):
        """attributes/attribute.py:24-38."""
        if isinstance(sizes, list):
            if opt == 0:
                assert len(sizes) == self.num_features_cat
                self._embedding_size_list_cat = sizes
            else:
                assert len(sizes) == self.num_features_mulhot
                self._embedding_size_list_mulhot = sizes
        elif isinstance(sizes, int):
            self._embedding_size_list_cat = [sizes] * self.num_features_cat
            self._embedding_size_list_mulhot = [sizes] * self.num_features_mulhot
        else:
            print('error: sizes need to be list or int')
            exit(0)

    def set_target_prediction(self, features_cat_tr, full_values_tr, full_segids_tr, full_lengths_tr):
        """attributes/attribute.py:40-47."""
        self.full_cat_tr = [np.asarray(a, dtype=np.int32) for a in features_cat_tr]
        self.full_values_tr = [np.asarray(a, dtype=np.int32) for a in full_values_tr]
        self.full_segids_tr = [np.asarray(a, dtype=np.int32) for a in full_segids_tr]
        self.full_lengths_tr = [np.asarray(a, dtype=np.float32).reshape(-1, 1) for a in full_lengths_tr]

    def set_target_prediction_from_map(self, logit_ind2item_ind):
        """Vectorised filter_cat + filter_mulhot (utils/preprocess.py:240-326): catalog-ordered
        copies of the attribute arrays for logit index v -> entity logit_ind2item_ind[v]."""
        V = len(logit_ind2item_ind)
        ids = np.asarray([logit_ind2item_ind[v] for v in range(V)], dtype=np.int64)
        cat_tr = [self.features_cat[i][ids] for i in range(self.num_features_cat)]
        vals, segs, lens = [], [], []
        for i in range(self.num_features_mulhot):
            s = self.mulhot_starts[i][ids].astype(np.int64)
            l = self.mulhot_lengths[i][ids].astype(np.int64)
            seg = np.repeat(np.arange(V, dtype=np.int64), l)
            off = np.concatenate([[0], np.cumsum(l)[:-1]])
            pos = np.arange(int(l.sum()), dtype=np.int64) - np.repeat(off, l) + np.repeat(s, l)
            vals.append(self.features_mulhot[i][pos])
            segs.append(seg)
            lens.append(l.astype(np.float32).reshape(V, 1))
        self.set_target_prediction(cat_tr, vals, segs, lens)

    def overview(self, out=None):
        p = out if out else print
        p('# of categorical attributes: {}'.format(self.num_features_cat))
        p('# of multi-hot   attributes: {}'.format(self.num_features_mulhot))
        p('vocab sizes cat {} mulhot {}'.format(self._embedding_classes_list_cat,
                                                self._embedding_classes_list_mulhot))
