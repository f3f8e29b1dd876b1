"""HET / MIX attribute-combination strategies (reference: attributes/comb_attribute.py:6-176).

HET: one embedding table per column, typed by *_attr.csv.  MIX: every value becomes the token
'<column-name><value>' (user id column renamed 'uid') and all tokens of an entity form ONE
multi-hot attribute.  Also builds the item-index <-> logit-index maps.
"""
import numpy as np

from . import attribute
from ..utils.preprocess import create_dictionary, create_dictionary_mix, tokenize_attribute_map


class Comb_Attributes(object):
    def get_attributes(self, users, items, data_tr, user_features, item_features):
        """comb_attribute.py:10-71."""
        user_feature_names, user_feature_types = user_features
        item_feature_names, item_feature_types = item_features
        u_inds = [p[0] for p in data_tr]
        self.create_dictionary(self.data_dir, u_inds, users, user_feature_types, user_feature_names,
                               self.max_vocabulary_size, self.logits_size_tr, prefix='user',
                               threshold=self.threshold)
        t = tokenize_attribute_map(self.data_dir, users, user_feature_types, self.max_vocabulary_size,
                                   self.logits_size_tr, prefix='user')
        u_attributes = attribute.Attributes(*t)

        i_inds_tr = [p[1] for p in data_tr]
        self.create_dictionary(self.data_dir, i_inds_tr, items, item_feature_types, item_feature_names,
                               self.max_vocabulary_size, self.logits_size_tr, prefix='item',
                               threshold=self.threshold)
        t2 = tokenize_attribute_map(self.data_dir, items, item_feature_types, self.max_vocabulary_size,
                                    self.logits_size_tr, prefix='item')
        features_cat2 = t2[1]
        item2fea0 = features_cat2[0] if len(features_cat2) > 0 else None
        item_ind2logit_ind, logit_ind2item_ind = self.index_mapping(item2fea0, i_inds_tr, len(items))
        i_attributes = attribute.Attributes(*t2)
        # filter_cat + filter_mulhot (preprocess.py:240-326): catalog-ordered copies
        i_attributes.set_target_prediction_from_map(logit_ind2item_ind)
        return u_attributes, i_attributes, item_ind2logit_ind, logit_ind2item_ind


class MIX(Comb_Attributes):
    def __init__(self, data_dir, max_vocabulary_size=500000, logits_size_tr=50000, threshold=2):
        self.data_dir = data_dir
        self.max_vocabulary_size = max_vocabulary_size
        self.logits_size_tr = logits_size_tr
        self.threshold = threshold
        self.create_dictionary = create_dictionary_mix

    def index_mapping(self, item2fea0, i_inds, M=None):
        """Top logits_size_tr items by training count (comb_attribute.py:83-98); ties keep
        first-occurrence order (the reference left them to Python-2 dict order)."""
        i_inds = np.asarray(i_inds, dtype=np.int64)
        uniq, first, counts = np.unique(i_inds, return_index=True, return_counts=True)
        assert self.logits_size_tr <= len(uniq), 'Item_vocab_size should be smaller than # of appeared items'
        order = np.lexsort((first, -counts))[:self.logits_size_tr]
        ind_list = uniq[order]
        item_ind2logit_ind = {int(e): k for k, e in enumerate(ind_list)}
        logit_ind2item_ind = {k: int(e) for k, e in enumerate(ind_list)}
        return item_ind2logit_ind, logit_ind2item_ind

    def mix_attr(self, users, items, user_features, item_features):
        """comb_attribute.py:100-148."""
        user_feature_names, user_feature_types = user_features
        item_feature_names, item_feature_types = item_features
        user_feature_names = list(user_feature_names)
        user_feature_names[0] = 'uid'

        def mix(rows, names, types):
            out = np.zeros((len(rows), 1), dtype=object)
            for i in range(len(rows)):
                v = []
                for j, t in enumerate(types):
                    if t == 0:
                        v.append(names[j] + str(rows[i, j]))
                    elif t == 1:
                        v.extend([names[j] + s for s in str(rows[i, j]).split(',')])
                out[i, 0] = ','.join(v)
            return out
        users2 = mix(users, user_feature_names, user_feature_types)
        items2 = mix(items, item_feature_names, item_feature_types)
        uf = (['mix'], [0]) if (len(user_feature_types) == 1 and user_feature_types[0] == 0) else (['mix'], [1])
        itf = (['mix'], [0]) if (len(item_feature_types) == 1 and item_feature_types[0] == 0) else (['mix'], [1])
        return users2, items2, uf, itf


class HET(Comb_Attributes):
    def __init__(self, data_dir, max_vocabulary_size=50000, logits_size_tr=50000, threshold=2):
        self.data_dir = data_dir
        self.max_vocabulary_size = max_vocabulary_size
        self.logits_size_tr = logits_size_tr
        self.threshold = threshold
        self.create_dictionary = create_dictionary

    def index_mapping(self, item2fea0, i_inds, M):
        """Items whose id token survived the vocabulary filter, in entity order (comb_attribute.py:162-176)."""
        keep = np.nonzero(np.asarray(item2fea0[:M]) != 0)[0]
        ind = len(keep)
        assert ind == self.logits_size_tr, ('Item_vocab_size %d too large! need to be no greater than %d\n'
                                            'Fix: --item_vocab_size [smaller item_vocab_size]\n' % (self.logits_size_tr, ind))
        item_ind2logit_ind = {int(e): k for k, e in enumerate(keep)}
        logit_ind2item_ind = {k: int(e) for k, e in enumerate(keep)}
        return item_ind2logit_ind, logit_ind2item_ind
