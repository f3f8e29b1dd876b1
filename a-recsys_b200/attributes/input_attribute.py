"""read_data: raw TSVs -> (interactions, Attributes, index maps), cached in data_dir
(reference: attributes/input_attribute.py:10-71).  The cache is the `.npy` + manifest store of attribute_store.py
(memory-mapped on load); a reference-style pickle `data_dir/data` is still read when it is all there is."""
import os
import pickle

from . import attribute_store
from .comb_attribute import HET, MIX
from ..utils.load_data import load_raw_data
from ..utils.preprocess import pickle_save


def read_data(raw_data_dir='../raw_data/data/', data_dir='../cache/data/', combine_att='mix',
              logits_size_tr='10000', thresh=2, use_user_feature=True, use_item_feature=True,
              no_user_id=False, test=False, mylog=None):
    if not mylog:
        mylog = print
    data_filename = os.path.join(data_dir, 'data')
    if attribute_store.store_exists(data_dir):
        mylog("attribute store {} exists! loading cached data (memory-mapped). \nCaution: change cached data dir "
              "(--data_dir) if new data (or new preprocessing) is used.".format(os.path.join(data_dir, 'store')))
        (data_tr, data_va, u_attr, i_attr, item_ind2logit_ind, logit_ind2item_ind, user_index,
         item_index) = attribute_store.load_store(data_dir)
    elif os.path.isfile(data_filename):
        mylog("data file {} exists! loading cached data. \nCaution: change cached data dir (--data_dir) "
              "if new data (or new preprocessing) is used.".format(data_filename))
        with open(data_filename, 'rb') as f:
            (data_tr, data_va, u_attr, i_attr, item_ind2logit_ind, logit_ind2item_ind, user_index,
             item_index) = pickle.load(f)
    else:
        if not os.path.exists(data_dir):
            os.makedirs(data_dir)
        _submit = 1 if test else 0
        (users, items, data_tr, data_va, user_features, item_features, user_index,
         item_index) = load_raw_data(data_dir=raw_data_dir, _submit=_submit)
        if not use_user_feature:
            users = users[:, 0].reshape(len(users), 1)
            user_features = ([user_features[0][0]], [user_features[1][0]])
        if not use_item_feature:
            items = items[:, 0].reshape(len(items), 1)
            item_features = ([item_features[0][0]], [item_features[1][0]])
        if no_user_id:
            users[:, 0] = 0                                                   # :43-44
        if combine_att == 'het':
            het = HET(data_dir=data_dir, logits_size_tr=logits_size_tr, threshold=thresh)
            u_attr, i_attr, item_ind2logit_ind, logit_ind2item_ind = het.get_attributes(
                users, items, data_tr, user_features, item_features)
        elif combine_att == 'mix':
            mix = MIX(data_dir=data_dir, logits_size_tr=logits_size_tr, threshold=thresh)
            users2, items2, user_features, item_features = mix.mix_attr(users, items, user_features, item_features)
            u_attr, i_attr, item_ind2logit_ind, logit_ind2item_ind = mix.get_attributes(
                users2, items2, data_tr, user_features, item_features)
        mylog("saving data format to data directory")
        attribute_store.save_store(data_dir, data_tr, data_va, u_attr, i_attr, item_ind2logit_ind, logit_ind2item_ind,
                                   user_index, item_index)
    mylog('length of item_ind2logit_ind: {}'.format(len(item_ind2logit_ind)))
    return (data_tr, data_va, u_attr, i_attr, item_ind2logit_ind, logit_ind2item_ind, user_index, item_index)
