"""CPU tests of the host data layer: the Python-3 port of the
reference's load / vocabulary / tokenise / HET-MIX code, pinned on
  * the facts SURVEY.md 8(c) derived independently (pandas) from the bundled MovieLens-1m files,
  * a hand-written miniature dataset that exercises every edge the reference handles."""
import json
import os

import numpy as np
import pytest

import arecsys_b200  # noqa: F401
from arecsys_b200.attributes.input_attribute import read_data
from arecsys_b200.utils import prepare_train
from arecsys_b200.utils.preprocess import UNK_ID, START_ID

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
REF_DATA = '/root/reference/examples/dataset/'


def test_ml1m_facts_match_survey():
    f = json.load(open(os.path.join(GOLD, 'ml1m_facts.json')))
    m, h = f['mix'], f['het']
    # SURVEY.md 8(c) "Golden vectors / fixtures available in the reference"
    assert (m['n_users'], m['n_items']) == (6040, 3883)
    assert (m['n_train'], m['n_valid']) == (342436, 114886)
    assert m['distinct_train_items'] == 3391
    assert (m['train_kept'], m['valid_kept']) == (342057, 114554)
    assert m['item_vocab_mulhot'] == [7297] and m['user_vocab_mulhot'] == [5982]
    assert m['catalog_nnz'] == [20883]
    assert m['user_values'] == [24071] and m['item_values'] == [24481]
    assert h['item_vocab_mulhot'] == [20, 4091] and h['item_values'] == [6409, 14885]
    assert h['item_vocab_cat'] == [3102] and h['user_vocab_cat'][1:] == [4, 9, 23]


def test_ml1m_fixture_is_consistent_csr():
    z = np.load(os.path.join(GOLD, 'ml1m_mix.npz'))
    for p in ('u', 'i'):
        v, s, l = z[p + '_values'], z[p + '_starts'], z[p + '_lengths']
        assert s[-1] == len(v) and (np.diff(s) == l).all() and (l >= 1).all()
        assert v[-1] == START_ID and l[-1] == 1                      # trailing START pseudo-entity
        assert v.max() < int(z[p + '_vocab'])
    assert len(z['logit2item']) == 3100 and len(np.unique(z['logit2item'])) == 3100
    assert np.isin(z['train'][:, 1], z['logit2item']).all()


@pytest.mark.skipif(not os.path.isdir(REF_DATA), reason='reference dataset only exists in the authoring container')
def test_read_data_reproduces_fixture(tmp_path):
    (data_tr, data_va, ua, ia, i2l, l2i, uidx, iidx) = read_data(REF_DATA, str(tmp_path / 'c'), 'mix', 3100, 2,
                                                                  mylog=lambda s: None)
    z = np.load(os.path.join(GOLD, 'ml1m_mix.npz'))
    assert np.array_equal(ua.features_mulhot[0], z['u_values']) and np.array_equal(ia.features_mulhot[0], z['i_values'])
    assert np.array_equal(ia.mulhot_starts[0], z['i_starts'])
    assert [l2i[k] for k in range(3100)] == z['logit2item'].tolist()
    # cache hit returns the same objects' content
    again = read_data(REF_DATA, str(tmp_path / 'c'), 'mix', 3100, 2, mylog=lambda s: None)
    assert np.array_equal(again[2].features_mulhot[0], ua.features_mulhot[0]) and again[4] == i2l
    assert os.path.isfile(tmp_path / 'c' / 'item_vocab0_500000') and os.path.isfile(tmp_path / 'c' / 'store' / 'manifest.json')


def _mini(tmp_path, with_types=True):
    d = tmp_path / 'raw'
    d.mkdir()
    (d / 'u.csv').write_text('id\tgender\tage\n10\tF\t1\n11\tM\t2\n12\tM\t1\n13\tF\t3\n')
    (d / 'i.csv').write_bytes('id\tgenres\ttitle\n100\tA,B\tCaf\xe9,One\n101\tB\tTwo\n102\tC\tOne,Two\n103\tZ\tRare\n'.encode('latin-1'))
    if with_types:
        (d / 'u_attr.csv').write_text('id\tgender\tage\n0\t0\t0\n')
        (d / 'i_attr.csv').write_text('id\tgenres\ttitle\n0\t1\t1\n')
    (d / 'obs_tr.csv').write_text('user\titem\n10\t100\n11\t100\n12\t101\n10\t101\n11\t102\n12\t102\n10\t103\n')
    (d / 'obs_va.csv').write_text('user\titem\ttime\n13\t100\t5\n')
    (d / 'obs_te.csv').write_text('user\titem\ttime\n13\t101\t6\n')
    return str(d) + '/'


def test_het_miniature(tmp_path):
    raw = _mini(tmp_path)
    (data_tr, data_va, ua, ia, i2l, l2i, uidx, iidx) = read_data(raw, str(tmp_path / 'c'), 'het', 3, 2, mylog=lambda s: None)
    assert data_tr[0] == (0, 0, 0) and len(data_tr) == 7 and data_va == [(3, 0, 5)]      # zeros appended as time
    # item id vocab: items with >= 2 training interactions: 100,101,102 (103 has one) -> vocab size 3+2
    assert ia._embedding_classes_list_cat == [5]
    assert ia.features_cat[0].tolist()[:3] != [0, 0, 0] and ia.features_cat[0][3] == UNK_ID and ia.features_cat[0][4] == START_ID
    assert i2l == {0: 0, 1: 1, 2: 2} and l2i == {0: 0, 1: 1, 2: 2}
    # genres counted once per interaction: A:2 B:4 C:2 Z:1 -> vocab [_UNK,_START,B,A,C] (ties: first occurrence)
    v = open(tmp_path / 'c' / 'item_vocab1_50000', encoding='latin-1').read().split('\n')[:-1]
    assert v == ['_UNK', '_START', 'B', 'A', 'C']
    g = ia.features_mulhot[0]; s = ia.mulhot_starts[0]; l = ia.mulhot_lengths[0]
    assert g.tolist() == [3, 2, 2, 4, UNK_ID, START_ID]        # item 103: only 'Z' (filtered) -> [_UNK]
    assert l.tolist() == [2, 1, 1, 1, 1] and s.tolist() == [0, 2, 3, 4, 5, 6]
    # Latin-1 token survives the round trip through the vocab file
    t = open(tmp_path / 'c' / 'item_vocab2_50000', encoding='latin-1').read()
    assert 'Caf\xe9' in t
    # catalog-ordered copies (filter_mulhot): segment ids sorted, lengths float [V,1]
    assert ia.full_segids_tr[0].tolist() == [0, 0, 1, 2] and ia.full_lengths_tr[0].ravel().tolist() == [2.0, 1.0, 1.0]
    # user 13 never trains: its id token is UNK; cold users keep their attribute tokens
    assert ua.features_cat[0][3] == UNK_ID and ua.num_features_cat == 3 and ua.num_features_mulhot == 0


def test_mix_miniature_and_flags(tmp_path):
    raw = _mini(tmp_path)
    (data_tr, data_va, ua, ia, i2l, l2i, _, _) = read_data(raw, str(tmp_path / 'c'), 'mix', 3, 2, mylog=lambda s: None)
    assert ua.num_features_cat == 0 and ua.num_features_mulhot == 1 and ia.num_features_mulhot == 1
    uv = open(tmp_path / 'c' / 'user_vocab0_500000').read().split('\n')[:-1]
    assert uv[:2] == ['_UNK', '_START'] and uv[2:5] == ['uid10', 'uid11', 'uid12']     # uid tokens first
    assert 'uid13' not in uv and 'genderM' in uv
    # top-3 items by training count, ties by first occurrence
    assert [l2i[k] for k in range(3)] == [0, 1, 2]
    # --test True: train on tr+va, validate on te
    r2 = read_data(raw, str(tmp_path / 'c2'), 'mix', 3, 2, test=True, mylog=lambda s: None)
    assert len(r2[0]) == 8 and r2[1] == [(3, 1, 6)]
    # use_item_feature False keeps only the id column -> one categorical attribute
    r3 = read_data(raw, str(tmp_path / 'c3'), 'mix', 3, 2, use_item_feature=False, mylog=lambda s: None)
    assert r3[3].num_features_cat == 1 and r3[3].num_features_mulhot == 0
    # no_user_id overwrites the id column before tokenisation (input_attribute.py:43-44)
    r4 = read_data(raw, str(tmp_path / 'c4'), 'het', 3, 2, no_user_id=True, mylog=lambda s: None)
    assert len(set(r4[2].features_cat[0][:-1].tolist())) == 1


def test_missing_attr_files_mean_all_categorical(tmp_path):
    raw = _mini(tmp_path, with_types=False)
    r = read_data(raw, str(tmp_path / 'c'), 'het', 3, 2, mylog=lambda s: None)
    assert r[3].num_features_cat == 3 and r[3].num_features_mulhot == 0


def test_sampling_helpers():
    data = [(0, 5, 0), (1, 5, 0), (2, 7, 0), (0, 9, 0), (0, 5, 0)]
    pop, p = prepare_train.item_frequency(data, 0.5)
    assert pop == [5, 7, 9]
    w = np.sqrt(np.array([3, 1, 1]) / 5.0)
    np.testing.assert_allclose(p, w / w.sum())
    np.random.seed(0)
    s, m = prepare_train.sample_items(pop, 2, p)
    assert len(set(s)) == 2 and all(m[int(v)] == k for k, v in enumerate(s))
    tr, va = prepare_train.positive_items(data, [(0, 7, 0)])
    assert sorted(tr[0]) == [5, 9] and va == {0: [7]}
    ptr, it = prepare_train.positives_csr([0, 1, 2, 0, 0], [5, 5, 7, 9, 5], 4)
    assert ptr.tolist() == [0, 2, 3, 4, 4] and it.tolist() == [5, 9, 5, 7]


def test_attribute_store_roundtrip_and_mmap(tmp_path):
    """the `.npy` + manifest store is an exact inverse pair, loads memory-mapped, and read_data's cache
    hit goes through it (no pickle of Python lists)."""
    from arecsys_b200.attributes import attribute_store
    from arecsys_b200.utils import synthetic
    ua, ia, i2l, l2i = synthetic.make_dataset(300, 200, 3, 50, 4, 9, seed=3, logit_size=150)
    i2l_d = {int(l2i[v]): v for v in range(len(l2i))}
    l2i_d = {v: int(l2i[v]) for v in range(len(l2i))}
    data_tr = [(1, 2, 0), (5, 7, 11), (299, 149, 3)]
    data_va = [(4, 4, 9)]
    uidx = {'u%d' % k: k for k in range(300)}
    iidx = {1000 + k: k for k in range(200)}                      # int raw ids keep their type
    d = attribute_store.save_store(str(tmp_path), data_tr, data_va, ua, ia, i2l_d, l2i_d, uidx, iidx)
    assert attribute_store.store_exists(str(tmp_path)) and os.path.isfile(os.path.join(d, 'i_mul_0_values.npy'))
    tr, va, ua2, ia2, i2l2, l2i2, uidx2, iidx2 = attribute_store.load_store(str(tmp_path))
    assert tr == data_tr and va == data_va and i2l2 == i2l_d and l2i2 == l2i_d and uidx2 == uidx and iidx2 == iidx
    for a, b in ((ua, ua2), (ia, ia2)):
        assert a.num_features_cat == b.num_features_cat and a.num_features_mulhot == b.num_features_mulhot
        assert a._embedding_classes_list_cat == b._embedding_classes_list_cat
        assert a._embedding_classes_list_mulhot == b._embedding_classes_list_mulhot
        for x, y in zip(a.features_cat + a.features_mulhot + a.mulhot_starts + a.mulhot_lengths,
                        b.features_cat + b.features_mulhot + b.mulhot_starts + b.mulhot_lengths):
            assert isinstance(y, np.memmap) and y.dtype == np.int32 and np.array_equal(x, y)
    for x, y in zip(ia.full_cat_tr + ia.full_values_tr + ia.full_segids_tr + ia.full_lengths_tr,
                    ia2.full_cat_tr + ia2.full_values_tr + ia2.full_segids_tr + ia2.full_lengths_tr):
        assert np.array_equal(np.asarray(x), np.asarray(y))
    ia2.set_model_size(8)                                           # a loaded container behaves like a built one
    assert ia2.num_entities == ia.num_entities
