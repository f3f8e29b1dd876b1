"""CPU tests of the evaluation helpers used by --recommend True (utils/evaluate.py)."""
import numpy as np

import arecsys_b200  # noqa: F401
from arecsys_b200.utils.evaluate import metrics, combine_sub, format_submit, load_submit


def test_metrics_known_values():
    T = {1: ['a', 'b'], 2: ['c']}
    X = {1: ['a', 'x', 'b', 'y', 'z'], 2: ['q', 'c']}
    r = metrics(X, T, Ns=(2, 5))
    # user 1: hits at positions 1,3; user 2: hit at position 2
    np.testing.assert_allclose(r['prec'], [(0.5 + 0.5) / 2, (2 / 5 + 1 / 2) / 2])
    np.testing.assert_allclose(r['recall'], [(0.5 + 1.0) / 2, (1.0 + 1.0) / 2])
    ap1_2 = 1.0 / 2; ap1_5 = (1.0 + 2 / 3) / 2; ap2 = 0.5 / 1
    np.testing.assert_allclose(r['map'], [(ap1_2 + ap2) / 2, (ap1_5 + ap2) / 2])
    d = [1 / np.log2(2 + n) for n in range(5)]
    n1_2 = d[0] / (d[0] + d[1]); n1_5 = (d[0] + d[2]) / (d[0] + d[1]); n2 = d[1] / d[0]
    np.testing.assert_allclose(r['ndcg'], [(n1_2 + n2) / 2, (n1_5 + n2) / 2])
    # users of T missing from X count as zeros
    assert metrics({}, T)['prec'] == [0.0] * 5


def test_combine_and_roundtrip(tmp_path):
    users = np.array([[1], [2], [3]], dtype=object)
    hist = {1: ['a'], 2: ['c']}
    rec = {1: ['a', 'b', 'b'], 3: ['z']}
    assert combine_sub(hist, rec, 1, users=users) == {1: ['b'], 2: [], 3: ['z']}
    assert combine_sub(hist, rec, 0, users=users) == {1: ['a', 'b'], 2: ['c'], 3: ['z']}
    format_submit({1: ['5', '6'], 2: []}, 's.csv', str(tmp_path))
    assert load_submit('s.csv', str(tmp_path)) == {1: ['5', '6'], 2: []}
