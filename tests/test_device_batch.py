"""Device-side batchers (a-recsys_b200/utils/device_batch.py) against the host functions they replace,
which tests/test_host_logic_vs_reference.py pins to the reference's own code under the same seeds."""
import random

import numpy as np
import pytest
import torch

from helpers import small_dataset, random_params


def test_py_random_stream_is_pythons_own_stream():
    """random.choice / randrange draws reproduced vectorised: same indices, same generator state afterwards."""
    import arecsys_b200  # noqa: F401
    from arecsys_b200.utils.device_batch import PyRandomStream
    for n in (1, 2, 3, 7, 1000, 342057, 1 << 20, (1 << 20) + 1, (1 << 31) - 1):
        random.seed(123)
        want = [random.randrange(n) for _ in range(3000)]
        after_want = [random.random() for _ in range(3)]
        random.seed(123)
        got = PyRandomStream.draw(n, 3000)
        assert list(got) == want and [random.random() for _ in range(3)] == after_want, n
    seq = [(i, 2 * i, 0) for i in range(777)]
    random.seed(5)
    want = [random.choice(seq) for _ in range(64)]
    random.seed(5)
    assert [seq[i] for i in PyRandomStream.draw(len(seq), 64)] == want


@pytest.mark.gpu
@pytest.mark.parametrize('sample_type', ['random', 'permute'])
def test_device_interaction_sampler_equals_get_batch(cuda, sample_type):
    """Same seeds -> the device batches are LatentProductModel.get_batch / get_permuted_batch (hmf_model.py:230-260)."""
    from arecsys_b200.hmf.hmf_model import LatentProductModel
    from arecsys_b200.utils.device_batch import DeviceInteractionSampler
    ua, ia, i2l, l2i = small_dataset(60, 50, 2, 25, 3, 6, 0, None, 8)
    l2i_d = {int(v): int(l2i[v]) for v in range(len(l2i))}
    i2l_d = {v: k for k, v in l2i_d.items()}
    mb = 32
    model = LatentProductModel(60, 50, 8, 1, mb, 0.1, 1.0, ua, ia, i2l_d, l2i_d, loss_function='ce',
                               params=random_params(ua, ia, 8, 1))
    rng = np.random.default_rng(0)
    data = [(int(u), int(i), 0) for u, i in zip(rng.integers(0, 60, 1000), rng.integers(0, 50, 1000))]
    random.seed(11); np.random.seed(11)
    want = []
    for _ in range(40):                          # 40 x 32 > 1000: the permutation sweep restarts inside
        u, i, _ = model.get_batch(data) if sample_type == 'random' else model.get_permuted_batch(data)
        want.append((list(u), list(i)))
    random.seed(11); np.random.seed(11)
    ds = DeviceInteractionSampler(data, mb, cuda, sample_type)
    for k in range(40):
        u, i = ds.next()
        assert u.cpu().tolist() == want[k][0] and i.cpu().tolist() == want[k][1], k
    # and the device tensors feed the step directly
    loss = model.step(None, u, i, loss='ce')
    assert np.isfinite(loss)


@pytest.mark.gpu
def test_device_seq_batcher_equals_get_batch(cuda):
    """SeqModel.get_batch (lstm/seqModel.py:356-404), random and sweep selection, against the device padding kernel."""
    from test_gpu_lstm import _seq_setup
    from arecsys_b200.utils.device_batch import DeviceSeqBatcher
    model, ref, emb, rng, (n_users, n_items, START, dim, mb, T) = _seq_setup('ce', False, False)
    buckets = model.buckets                                  # [3, T]
    data = [[], []]
    for _ in range(57):
        L = int(rng.integers(1, T + 1))
        data[0 if L <= buckets[0] else 1].append((int(rng.integers(0, n_users)), rng.integers(0, n_items, L).tolist()))
    sb = DeviceSeqBatcher(data, buckets, mb, START, cuda, user_pad_id=model.USER_PAD_ID)
    for b in (0, 1):
        random.seed(3)
        wu, wi, wo, ww, _ = model.get_batch(data, b)
        random.seed(3)
        u, i, o, w, _ = sb.next(b)
        assert u.cpu().tolist() == wu and i.cpu().tolist() == wi and o.cpu().tolist() == wo and w.cpu().tolist() == ww
        for start in (0, mb, (len(data[b]) // mb) * mb):      # the last window is ragged: empty slots
            wu, wi, wo, ww, wf = model.get_batch(data, b, start_id=start)
            u, i, o, w, f = sb.next(b, start_id=start)
            assert (u.cpu().tolist(), i.cpu().tolist(), o.cpu().tolist(), w.cpu().tolist(), f) == (wu, wi, wo, ww, wf), (b, start)
    # device tensors straight into the step == the same batch as Python lists
    random.seed(9)
    wu, wi, wo, ww, _ = model.get_batch(data, 1)
    random.seed(9)
    u, i, o, w, _ = sb.next(1)
    l_dev = model.step(None, u, i, o, w, 1, forward_only=True)
    l_lst = model.step(None, wu, wi, wo, ww, 1, forward_only=True)
    assert abs(l_dev - l_lst) <= 1e-6 * max(1.0, abs(l_lst))


@pytest.mark.gpu
def test_device_cbow_batcher_stream_and_windows(cuda):
    """(user, target) stream identical to the host batcher's (= the reference's get_next_cbow); every input comes from the
    `window` stream positions before its target, distinct positions once the user has >= ni events in the window."""
    import arecsys_b200  # noqa: F401
    from arecsys_b200.word2vec.data_iterator import DataIterator
    from arecsys_b200.utils.device_batch import DeviceCbowBatcher
    PAD, mb, ni, c = 999, 16, 3, 5
    seq = []
    rng = np.random.default_rng(1)
    for u in range(40):
        seq.append((u, PAD))
        seq.extend((u, int(1000 * u + k)) for k in range(int(rng.integers(1, 12))))     # distinct items: positions identifiable
    items = np.array([s[1] for s in seq]); users_all = np.array([s[0] for s in seq])
    pos_of = {int(v): p for p, v in enumerate(items) if v != PAD}
    host = DataIterator(seq, PAD, mb, ni, c, False).get_next_cbow()
    dev = DeviceCbowBatcher(seq, PAD, mb, ni, c, cuda, seed=4)
    np.random.seed(0)
    seen_inputs = set()
    for step in range(60):
        hu, _, ho = next(host)
        u, inp, o = dev.next()
        assert u.cpu().tolist() == hu.tolist() and o.cpu().tolist() == ho.tolist(), step
        inp = inp.cpu().numpy()
        for b in range(mb):
            p = pos_of[int(o[b])]
            window = [(p - c + k) % len(seq) for k in range(c)]
            chosen = []
            for k in range(ni):
                cand = [q for q in window if items[q] == inp[k, b]]
                assert cand, (step, b, k)
                chosen.append(cand[0] if items[cand[0]] != PAD else None)
            n_before = min(p - max(q for q in range(p + 1) if items[q] == PAD), c)
            real = [q for q in chosen if q is not None]
            if n_before >= ni and len(real) == ni:
                assert len(set(real)) == ni                  # without replacement
            seen_inputs.update(int(v) for v in inp[:, b])
    assert len(seen_inputs) > 50                             # the draws vary


@pytest.mark.gpu
def test_device_item_sampler_matches_numpy_choice_distribution(cuda):
    """Gumbel-top-k through arx_gumbel_keys + arx_topk_rows == np.random.choice(items, n, replace=False, p): distinct
    items, and the same inclusion probabilities and first-draw distribution (chi-square-sized bars over 20 000 draws)."""
    import arecsys_b200  # noqa: F401
    from arecsys_b200.utils.prepare_train import DeviceItemSampler
    pop = np.arange(100, 108)
    p = np.array([0.30, 0.22, 0.15, 0.10, 0.09, 0.07, 0.05, 0.02])
    n, draws = 3, 20000
    s = DeviceItemSampler(pop, p, cuda, seed=1)
    got = torch.stack([s.sample(n) for _ in range(draws)]).cpu().numpy()
    assert all(len(set(r)) == n for r in got[:500])
    rs = np.random.RandomState(0)
    ref = np.stack([rs.choice(pop, n, replace=False, p=p) for _ in range(draws)])
    for name, a, b in (('inclusion', got, ref), ('first draw', got[:, :1], ref[:, :1])):
        fa = np.array([(a == v).any(1).mean() for v in pop]); fb = np.array([(b == v).any(1).mean() for v in pop])
        sigma = np.sqrt(fb * (1 - fb) / draws * 2) + 1e-4
        assert (np.abs(fa - fb) < 5 * sigma).all(), (name, fa, fb)
    # large population: the two-stage selection returns the n largest keys (distinct, valid ids), heavy items dominate
    V = 200000
    pl = np.ones(V); pl[:10] = 5000.0; pl /= pl.sum()
    sl = DeviceItemSampler(np.arange(V), pl, cuda, seed=2)
    out = sl.sample(1024).cpu().numpy()
    assert len(set(out.tolist())) == 1024 and out.min() >= 0 and out.max() < V
    assert sum(int(v) < 10 for v in out) >= 9                # P(each heavy item in the pool) ~ 1
    s2 = DeviceItemSampler(np.arange(V), pl, cuda, seed=2)
    assert np.array_equal(s2.sample(1024).cpu().numpy(), out)        # reproducible for a given (seed, step)
    assert not np.array_equal(s2.sample(1024).cpu().numpy(), out)    # and fresh on the next call
