"""Generates tests/golden/ml1m_facts.json and ml1m_mix.npz from the reference's bundled
MovieLens-1m TSVs (/root/reference/examples/dataset) through THIS repo's data layer.
Run in the authoring container only (the GPU box has no /root/reference):
    python tests/golden/make_ml1m_golden.py
The facts are cross-checked in tests/test_data_layer.py against the numbers SURVEY.md 8(c)
derived independently with pandas."""
import hashlib
import json
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import arecsys_b200  # noqa: E402,F401
from arecsys_b200.attributes.input_attribute import read_data  # noqa: E402

RAW = '/root/reference/examples/dataset/'


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def main():
    facts = {}
    for comb in ('mix', 'het'):
        d = tempfile.mkdtemp()
        (data_tr, data_va, ua, ia, i2l, l2i, uidx, iidx) = read_data(RAW, d, comb, 3100, 2, mylog=lambda s: None)
        facts[comb] = {
            'n_users': len(uidx), 'n_items': len(iidx), 'n_train': len(data_tr), 'n_valid': len(data_va),
            'distinct_train_items': len(set(p[1] for p in data_tr)),
            'train_kept': sum(1 for p in data_tr if p[1] in i2l), 'valid_kept': sum(1 for p in data_va if p[1] in i2l),
            'user_vocab_cat': list(ua._embedding_classes_list_cat), 'user_vocab_mulhot': list(ua._embedding_classes_list_mulhot),
            'item_vocab_cat': list(ia._embedding_classes_list_cat), 'item_vocab_mulhot': list(ia._embedding_classes_list_mulhot),
            'user_values': [len(v) for v in ua.features_mulhot], 'item_values': [len(v) for v in ia.features_mulhot],
            'catalog_nnz': [len(v) for v in ia.full_values_tr],
            'sha_item_values': [sha(v) for v in ia.features_mulhot], 'sha_user_values': [sha(v) for v in ua.features_mulhot],
        }
        if comb == 'mix':
            tr = np.asarray([(p[0], p[1]) for p in data_tr if p[1] in i2l], dtype=np.int32)
            va = np.asarray([(p[0], p[1]) for p in data_va if p[1] in i2l], dtype=np.int32)
            np.savez_compressed(os.path.join(HERE, 'ml1m_mix.npz'),
                                u_values=ua.features_mulhot[0], u_starts=ua.mulhot_starts[0], u_lengths=ua.mulhot_lengths[0],
                                i_values=ia.features_mulhot[0], i_starts=ia.mulhot_starts[0], i_lengths=ia.mulhot_lengths[0],
                                u_vocab=np.int64(ua._embedding_classes_list_mulhot[0]),
                                i_vocab=np.int64(ia._embedding_classes_list_mulhot[0]),
                                logit2item=np.asarray([l2i[k] for k in range(len(l2i))], dtype=np.int32),
                                train=tr, valid=va)
    with open(os.path.join(HERE, 'ml1m_facts.json'), 'w') as f:
        json.dump(facts, f, indent=1, sort_keys=True)
    print(json.dumps(facts['mix'], indent=1))


if __name__ == '__main__':
    main()
