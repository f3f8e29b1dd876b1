"""Parity against golden vectors produced by the reference's OWN model code.

`tests/golden/ref_hmf_*.npz` were written by `tests/golden/make_ref_golden.py`, which imports the
unmodified `/root/reference/hmf/hmf_model.py` + `attributes/embed_attribute.py` + `mulhot_index.py`
on top of `oracle/tf1_shim` (a TF-1.0 op restatement) and records what that graph computes.
  * not gpu: the NumPy oracle (`oracle/np_oracle.py`) must reproduce the reference's losses,
    parameters after training, eval loss and top-k on the same inputs / weights / dropout masks;
  * gpu: the CUDA path through the C ABI must do the same (exact-fp32 contractions: 1e-4;
    tcgen05 tf32 contractions: the north star's 1e-3).
Nothing here reads /root/reference at run time, except the provenance test at the end (skipped where the
reference sources are absent), which regenerates fixtures from them and compares.
"""
import glob
import os

import numpy as np
import pytest

import arecsys_b200  # noqa: F401
from arecsys_b200.attributes.attribute import Attributes
from oracle import np_oracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
CASES = sorted(os.path.basename(p)[len('ref_hmf_'):-4] for p in glob.glob(os.path.join(GOLD, 'ref_hmf_*.npz')))


def _attributes(d, prefix, dim):
    nc, nm = int(d[prefix + 'n_cat']), int(d[prefix + 'n_mulhot'])
    a = Attributes(nc, [d['%scat_%d' % (prefix, i)] for i in range(nc)], nm,
                   [d['%svalues_%d' % (prefix, i)] for i in range(nm)], None,
                   [d['%sstarts_%d' % (prefix, i)] for i in range(nm)],
                   [d['%slengths_%d' % (prefix, i)] for i in range(nm)],
                   d[prefix + 'v_cat'].tolist(), d[prefix + 'v_mulhot'].tolist())
    a.set_model_size(dim)
    return a


class Case(object):
    def __init__(self, name):
        d = np.load(os.path.join(GOLD, 'ref_hmf_%s.npz' % name))
        self.d = d
        self.name = name
        self.loss = str(d['loss']); self.nonlinear = str(d['nonlinear']); self.loss_func = str(d['loss_func'])
        self.dim = int(d['dim']); self.mb = int(d['mb']); self.hidden = int(d['hidden'])
        self.ns = None if int(d['n_sampled']) < 0 else int(d['n_sampled'])
        self.lr = float(d['lr']); self.keep = float(d['keep_prob']); self.exp_p = float(d['exp_p'])
        self.n_users = int(d['n_users']); self.n_items = int(d['n_items']); self.top_n = int(d['top_n'])
        self.n_steps = int(d['n_steps'])
        self.ua = _attributes(d, 'u_', self.dim)
        self.ia = _attributes(d, 'i_', self.dim)
        l2i = d['l2i']
        self.ia.set_target_prediction_from_map(l2i)
        self.l2i_d = {int(v): int(l2i[v]) for v in range(len(l2i))}
        self.i2l_d = {int(l2i[v]): int(v) for v in range(len(l2i))}
        self.params = {k[len('init/'):]: d[k] for k in d.files if k.startswith('init/')}
        self.final = {k[len('final/'):]: d[k] for k in d.files if k.startswith('final/')}

    def positives(self, tag):
        pu, ptr, it = self.d[tag + '/pos_users'], self.d[tag + '/pos_ptr'], self.d[tag + '/pos_items']
        return {int(u): [int(v) for v in it[ptr[j]:ptr[j + 1]]] for j, u in enumerate(pu)}

    def step_inputs(self, it):
        tag = 'step%d' % it
        sampled = [int(v) for v in self.d[tag + '/sampled']]
        masks = [self.d['%s/mask%d' % (tag, j)] for j in range(int(self.d[tag + '/n_masks']))]
        return (self.d[tag + '/users'].tolist(), self.d[tag + '/items'].tolist(), sampled or None,
                self.positives(tag), masks)


def test_fixtures_present():
    assert len(CASES) >= 12, CASES


@pytest.mark.parametrize('name', CASES)
def test_oracle_reproduces_reference_run(name):
    c = Case(name)
    emb = O.OracleEmbeddingAttribute(c.ua, c.ia, c.mb, c.ns, {k: v.copy() for k, v in c.params.items()},
                                     item_ind2logit_ind=c.i2l_d, logit_ind2item_ind=c.l2i_d, dtype=np.float64)
    om = O.OracleHMF(emb, loss=c.loss, nonlinear=c.nonlinear, keep_prob=c.keep, learning_rate=c.lr,
                     loss_func=c.loss_func, loss_exp_p=c.exp_p)
    id2idx = None
    for it in range(c.n_steps):
        users, items, sampled, pos, masks = c.step_inputs(it)
        om.emb.prepare_warp(pos, pos)
        if sampled:
            id2idx = {v: k for k, v in enumerate(sampled)}
        l = om.step(users, items, sampled, id2idx, masks=masks)
        ref = float(c.d['losses'][it])
        assert abs(l - ref) <= 2e-5 * max(1.0, abs(ref)), (name, it, l, ref)
    for k, v in c.final.items():
        np.testing.assert_allclose(np.asarray(om.emb.p[k]).reshape(v.shape), v, rtol=2e-4, atol=2e-5, err_msg=k)
    pos = c.positives('eval')
    om.emb.prepare_warp(pos, pos)
    ev = om.step(c.d['eval/users'].tolist(), c.d['eval/items'].tolist(), None, id2idx, forward_only=True)
    ref = float(c.d['eval/loss'])
    assert abs(ev - ref) <= 2e-5 * max(1.0, abs(ref)), (ev, ref)
    idx, _ = om.top_k(c.d['recommend/users'].tolist(), c.top_n)
    assert np.array_equal(np.asarray(idx), c.d['recommend/indices'])


@pytest.mark.gpu
@pytest.mark.parametrize('exact', [True, False])
@pytest.mark.parametrize('name', CASES)
def test_cuda_path_reproduces_reference_run(cuda, name, exact):
    import torch
    from arecsys_b200 import _lib
    from arecsys_b200.hmf.hmf_model import LatentProductModel
    c = Case(name)
    ltol, prtol, patol = (1e-4, 1e-3, 2e-5) if exact else (1e-3, 1e-2, 2e-3)
    _lib.exact_fp32 = exact
    try:
        model = LatentProductModel(c.n_users, c.n_items, c.dim, 1, c.mb, c.lr, 1.0, c.ua, c.ia, c.i2l_d, c.l2i_d,
                                   loss_function=c.loss, nonlinear=c.nonlinear, dropout=c.keep, n_sampled=c.ns,
                                   hidden_size=c.hidden, loss_func=c.loss_func, loss_exp_p=c.exp_p,
                                   params={k: v.copy() for k, v in c.params.items()}, top_N_items=c.top_n)
        id2idx = None
        for it in range(c.n_steps):
            users, items, sampled, pos, masks = c.step_inputs(it)
            model.prepare_warp(pos, pos)
            if sampled:
                id2idx = {v: k for k, v in enumerate(sampled)}
            tm = [torch.tensor(m, dtype=torch.float32, device='cuda') for m in masks]
            l = model.step(None, users, items, None, sampled, id2idx, loss=c.loss, masks=tm)
            ref = float(c.d['losses'][it])
            assert abs(l - ref) <= ltol * max(1.0, abs(ref)), (name, it, l, ref)
        for k, v in c.final.items():
            got = (model.att_emb.params[k] if k in model.att_emb.params else model.dense[k].data).cpu().numpy()
            np.testing.assert_allclose(got.reshape(v.shape), v, rtol=prtol, atol=patol, err_msg=k)
        pos = c.positives('eval')
        model.prepare_warp(pos, pos)
        ev = model.step(None, c.d['eval/users'].tolist(), c.d['eval/items'].tolist(), None, None, id2idx,
                        forward_only=True, loss=c.loss)
        ref = float(c.d['eval/loss'])
        assert abs(ev - ref) <= ltol * max(1.0, abs(ref)), (ev, ref)
        if exact:        # rank order is only bit-stable with exact-fp32 scores
            rec = model.step(None, c.d['recommend/users'].tolist(), None, forward_only=True, recommend=True)
            assert np.array_equal(np.asarray(rec), c.d['recommend/indices'])
        assert model.global_step.eval() == int(c.d['global_step'])
    finally:
        _lib.exact_fp32 = False


# ------------------------------------------------------------------ LSTM (lstm/seqModel.py) ---------
LSTM_CASES = sorted(os.path.basename(p)[len('ref_lstm_'):-4] for p in glob.glob(os.path.join(GOLD, 'ref_lstm_*.npz')))


class SeqCase(object):
    def __init__(self, name, kind='lstm'):
        d = np.load(os.path.join(GOLD, 'ref_%s_%s.npz' % (kind, name)))
        self.d = d
        self.output_feat = int(d['output_feat']) if 'output_feat' in d.files else 1
        self.no_input_item_feature = bool(d['no_input_item_feature']) if 'no_input_item_feature' in d.files else False
        self.num_layers = int(d['num_layers']) if 'num_layers' in d.files else 1
        self.loss = str(d['loss'])
        self.use_concat, self.sep, self.adagrad = bool(d['use_concat']), bool(d['sep']), bool(d['adagrad'])
        self.dim, self.mb, self.T = int(d['dim']), int(d['mb']), int(d['T'])
        self.buckets = [int(b) for b in d['buckets']]
        self.lr, self.keep, self.clip = float(d['lr']), float(d['keep_prob']), float(d['clip'])
        self.START, self.topk, self.n_steps = int(d['START']), int(d['topk']), int(d['n_steps'])
        self.ua = _attributes(d, 'u_', self.dim)
        self.ia = _attributes(d, 'i_', self.dim)
        l2i = d['l2i']
        self.ia.set_target_prediction_from_map(l2i)
        self.l2i_d = {int(v): int(l2i[v]) for v in range(len(l2i))}
        self.i2l_d = {v: k for k, v in self.l2i_d.items()}
        self.i2l_d[self.START] = 0
        self.params = {k[len('init/'):]: d[k] for k in d.files if k.startswith('init/')}
        self.final = {k[len('final/'):]: d[k] for k in d.files if k.startswith('final/')}

    def masks(self, it):
        """(input masks of layer 0, output masks of the stack[, input masks of layer 1, ...]) of training step `it`."""
        d = self.d
        return (d['step%d/in_masks' % it], d['step%d/out_masks' % it]) + tuple(
            d['step%d/in_masks_%d' % (it, l)] for l in range(1, self.num_layers))

    def batch(self, tag):
        d = self.d
        pu, ptr, it = d[tag + '/pos_users'], d[tag + '/pos_ptr'], d[tag + '/pos_items']
        pos = {int(u): [int(v) for v in it[ptr[j]:ptr[j + 1]]] for j, u in enumerate(pu)}
        return (d[tag + '/users'].tolist(), d[tag + '/inputs'].tolist(), d[tag + '/targets'].tolist(),
                d[tag + '/weights'].tolist(), pos)


def test_lstm_fixtures_present():
    assert len(LSTM_CASES) >= 6, LSTM_CASES


@pytest.mark.parametrize('name', LSTM_CASES)
def test_oracle_reproduces_reference_lstm_run(name):
    import torch
    from oracle.torch_cpu_ref import TorchRefSeq
    c = SeqCase(name)
    ref = TorchRefSeq(c.ua, c.ia, {k: v.copy() for k, v in c.params.items()}, c.l2i_d, c.i2l_d, loss=c.loss,
                      keep_prob=c.keep, learning_rate=c.lr, n_sampled=None, dtype=torch.float64, size=c.dim,
                      use_concat=c.use_concat, no_user_id=False, max_gradient_norm=c.clip, item_output=c.sep,
                      withAdagrad=c.adagrad)
    for it in range(c.n_steps):
        users, inp, tgt, w, pos = c.batch('step%d' % it)
        ref.pos, ref.pos_eval = pos, pos
        l = ref.step_seq(users, inp, tgt, w, masks=(c.d['step%d/in_masks' % it], c.d['step%d/out_masks' % it]))
        want = float(c.d['losses'][it])
        assert abs(l - want) <= 2e-5 * max(1.0, abs(want)), (name, it, l, want)
        gn = float(c.d['gnorms'][it])
        assert abs(ref.last_gnorm - gn) <= 1e-4 * max(1.0, gn), (name, it, ref.last_gnorm, gn)
    for k, v in c.final.items():
        np.testing.assert_allclose(ref.p[k].detach().numpy().reshape(v.shape), v, rtol=2e-4, atol=2e-5, err_msg=k)
    users, inp, tgt, w, pos = c.batch('eval')
    ref.pos, ref.pos_eval = pos, pos
    ev = ref.step_seq(users, inp, tgt, w, forward_only=True)
    want = float(c.d['eval/loss'])
    assert abs(ev - want) <= 2e-5 * max(1.0, abs(want)), (ev, want)
    # per-position top-k of softmax(logits) (seqModel.py:514-519) on the same eval batch
    tk = ref.topk_seq(users, inp, c.topk)
    assert np.array_equal(tk, c.d['eval/topk_indexes'])


@pytest.mark.gpu
@pytest.mark.parametrize('exact', [True, False])
@pytest.mark.parametrize('name', LSTM_CASES)
def test_cuda_path_reproduces_reference_lstm_run(cuda, name, exact):
    import torch
    from arecsys_b200 import _lib
    from arecsys_b200.attributes.embed_attribute import EmbeddingAttribute
    from arecsys_b200.lstm.seqModel import SeqModel
    c = SeqCase(name)
    ltol, ptol = (2e-4, 2e-3) if exact else (2e-3, 2e-2)
    _lib.exact_fp32 = exact
    try:
        params = {k: v.copy() for k, v in c.params.items()}
        emb = EmbeddingAttribute(c.ua, c.ia, c.mb, None, c.T, c.sep, c.i2l_d, c.l2i_d, params=params)
        model = SeqModel(c.buckets, c.dim, 1, c.clip, c.mb, c.lr, 0.83, emb, withAdagrad=c.adagrad,
                         dropoutRate=c.keep, START_ID=c.START, loss=c.loss, use_concat=c.use_concat,
                         no_user_id=False, topk_n=c.topk, params=params)
        for it in range(c.n_steps):
            users, inp, tgt, w, pos = c.batch('step%d' % it)
            emb.prepare_warp(pos, pos)
            im = torch.tensor(c.d['step%d/in_masks' % it], device='cuda')
            om = torch.tensor(c.d['step%d/out_masks' % it], device='cuda')
            l = model.step(None, users, inp, tgt, w, int(c.d['step%d/bucket' % it]), masks=(im, om))
            want = float(c.d['losses'][it])
            assert abs(l - want) <= ltol * max(1.0, abs(want)), (name, it, l, want)
            gn = float(c.d['gnorms'][it])
            assert abs(float(model.last_gnorm) - gn) <= 5 * ltol * max(1.0, gn), (float(model.last_gnorm), gn)
        dense = model.dense_params()
        for k, v in c.final.items():
            got = (emb.params[k] if k in emb.params else dense[k][0]).cpu().numpy()
            err = np.abs(got.reshape(v.shape) - v).max()
            assert err <= ptol * max(1.0, np.abs(v).max()), (k, err)
        users, inp, tgt, w, pos = c.batch('eval')
        emb.prepare_warp(pos, pos)
        ev = model.step(None, users, inp, tgt, w, 1, forward_only=True)
        want = float(c.d['eval/loss'])
        assert abs(ev - want) <= ltol * max(1.0, abs(want)), (ev, want)
    finally:
        _lib.exact_fp32 = False


# ------------------------------------------------------------------ CBOW (word2vec/cbow_model.py) ---
CBOW_CASES = sorted(os.path.basename(p)[len('ref_cbow_'):-4] for p in glob.glob(os.path.join(GOLD, 'ref_cbow_*.npz')))


class CbowCase(object):
    def __init__(self, name):
        d = np.load(os.path.join(GOLD, 'ref_cbow_%s.npz' % name))
        self.d = d
        self.loss, self.sep, self.ni = str(d['loss']), bool(d['sep']), int(d['ni'])
        self.sg = name.startswith('sg_')        # fixtures of the reference's skipgram_model.py
        self.dim, self.mb, self.n_users, self.n_items = int(d['dim']), int(d['mb']), int(d['n_users']), int(d['n_items'])
        self.lr, self.keep, self.top_n, self.n_steps = float(d['lr']), float(d['keep_prob']), int(d['top_n']), int(d['n_steps'])
        self.ua = _attributes(d, 'u_', self.dim)
        self.ia = _attributes(d, 'i_', self.dim)
        l2i = d['l2i']
        self.ia.set_target_prediction_from_map(l2i)
        self.l2i_d = {int(v): int(l2i[v]) for v in range(len(l2i))}
        self.i2l_d = {v: k for k, v in self.l2i_d.items()}
        self.params = {k[len('init/'):]: d[k] for k in d.files if k.startswith('init/')}
        self.final = {k[len('final/'):]: d[k] for k in d.files if k.startswith('final/')}

    def batch(self, it):
        d, tag = self.d, 'step%d' % it
        pu, ptr, pit = d[tag + '/pos_users'], d[tag + '/pos_ptr'], d[tag + '/pos_items']
        pos = {int(u): [int(v) for v in pit[ptr[j]:ptr[j + 1]]] for j, u in enumerate(pu)}
        return d[tag + '/users'], [x for x in d[tag + '/inputs']], d[tag + '/outputs'], d[tag + '/mask'], pos


def test_cbow_fixtures_present():
    assert len(CBOW_CASES) >= 4, CBOW_CASES


@pytest.mark.parametrize('name', CBOW_CASES)
def test_oracle_reproduces_reference_cbow_run(name):
    import torch
    from oracle.torch_cpu_ref import TorchRefCbow
    c = CbowCase(name)
    ref = TorchRefCbow(c.ua, c.ia, {k: v.copy() for k, v in c.params.items()}, c.l2i_d, c.i2l_d, loss=c.loss,
                       keep_prob=c.keep, learning_rate=c.lr, n_sampled=None, dtype=torch.float64, size=c.dim,
                       item_output=c.sep, ni=c.ni, sg=c.sg)
    for it in range(c.n_steps):
        users, ins, outs, mask, pos = c.batch(it)
        ref.pos, ref.pos_eval = pos, pos
        l = ref.step_cbow(users, ins, outs, mask=mask)
        want = float(c.d['losses'][it])
        assert abs(l - want) <= 2e-5 * max(1.0, abs(want)), (name, it, l, want)
    for k, v in c.final.items():
        np.testing.assert_allclose(ref.p[k].detach().numpy().reshape(v.shape), v, rtol=2e-4, atol=2e-5, err_msg=k)
    ev = ref.step_cbow(users, ins, outs, forward_only=True)
    want = float(c.d['eval/loss'])
    assert abs(ev - want) <= 2e-5 * max(1.0, abs(want)), (ev, want)
    assert np.array_equal(ref.recommend(users, ins, c.top_n), c.d['recommend/indices'])


@pytest.mark.gpu
@pytest.mark.parametrize('exact', [True, False])
@pytest.mark.parametrize('name', CBOW_CASES)
def test_cuda_path_reproduces_reference_cbow_run(cuda, name, exact):
    import torch
    from arecsys_b200 import _lib
    c = CbowCase(name)
    if c.sg:
        from arecsys_b200.word2vec.skipgram_model import Model
    else:
        from arecsys_b200.word2vec.cbow_model import Model
    ltol, ptol = (2e-4, 2e-3) if exact else (2e-3, 2e-2)
    _lib.exact_fp32 = exact
    try:
        model = Model(c.n_users, c.n_items, c.dim, c.mb, c.lr, 1.0, c.ua, c.ia, c.i2l_d, c.l2i_d, n_input_items=c.ni,
                      loss_function=c.loss, dropout=c.keep, top_N_items=c.top_n, use_sep_item=c.sep, n_sampled=None,
                      params={k: v.copy() for k, v in c.params.items()})
        for it in range(c.n_steps):
            users, ins, outs, mask, pos = c.batch(it)
            model.prepare_warp(pos, pos)
            l = model.step(None, users.tolist(), [x.tolist() for x in ins], outs.tolist(), None, None, loss=c.loss,
                           masks=[torch.tensor(mask, device='cuda')])
            want = float(c.d['losses'][it])
            assert abs(l - want) <= ltol * max(1.0, abs(want)), (name, it, l, want)
        for k, v in c.final.items():
            got = model.att_emb.params[k].cpu().numpy()
            assert np.abs(got.reshape(v.shape) - v).max() <= ptol * max(1.0, np.abs(v).max()), k
        ev = model.step(None, users.tolist(), [x.tolist() for x in ins], outs.tolist(), forward_only=True, loss=c.loss)
        want = float(c.d['eval/loss'])
        assert abs(ev - want) <= ltol * max(1.0, abs(want)), (ev, want)
        if exact:
            rec = model.step(None, users.tolist(), [x.tolist() for x in ins], forward_only=True, recommend=True)
            assert np.array_equal(np.asarray(rec), c.d['recommend/indices'])
    finally:
        _lib.exact_fp32 = False


# ------------------------------------------------------------------ provenance of the fixtures ------
@pytest.mark.skipif(not os.path.isdir('/root/reference/hmf'), reason='reference sources only exist in the authoring container')
@pytest.mark.parametrize('spec', ['mw_linear', 'rs_sig_exp', 'lstm:ce_adagrad_clipped', 'cbow:warp_sep_ni3'])
def test_committed_fixture_regenerates_from_the_reference_sources(tmp_path, spec):
    """The committed .npz really is what the unmodified reference code computes on the shim: regenerate it in a
    fresh interpreter and compare every array."""
    import subprocess
    import sys
    env = dict(os.environ, ARX_GOLDEN_OUT=str(tmp_path))
    subprocess.check_call([sys.executable, os.path.join(GOLD, 'make_ref_golden.py'), spec], env=env,
                          stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    kind, name = spec.split(':') if ':' in spec else ('hmf', spec)
    fn = 'ref_%s_%s.npz' % (kind, name)
    new, old = np.load(os.path.join(str(tmp_path), fn)), np.load(os.path.join(GOLD, fn))
    assert sorted(new.files) == sorted(old.files)
    for k in old.files:
        assert np.array_equal(new[k], old[k]), k


def test_oracle_reproduces_reference_warp_eval_ranks():
    """loss 'warp_eval' (embed_attribute.py:620-639) as the reference's own graph reports it: per-row margin rank
    (float) and true rank (integer, must be exact)."""
    d = np.load(os.path.join(GOLD, 'ref_hmfeval_warp_eval.npz'))
    dim, mb = int(d['dim']), int(d['mb'])
    ua, ia = _attributes(d, 'u_', dim), _attributes(d, 'i_', dim)
    l2i = d['l2i']
    ia.set_target_prediction_from_map(l2i)
    l2i_d = {int(v): int(l2i[v]) for v in range(len(l2i))}
    i2l_d = {int(l2i[v]): int(v) for v in range(len(l2i))}
    params = {k[len('init/'):]: d[k] for k in d.files if k.startswith('init/')}
    emb = O.OracleEmbeddingAttribute(ua, ia, mb, None, params, item_ind2logit_ind=i2l_d, logit_ind2item_ind=l2i_d,
                                     dtype=np.float64)
    om = O.OracleHMF(emb, loss='warp_eval', keep_prob=0.5, learning_rate=0.3)
    pu, ptr, it = d['pos_users'], d['pos_ptr'], d['pos_items']
    pos = {int(u): [int(v) for v in it[ptr[j]:ptr[j + 1]]] for j, u in enumerate(pu)}
    emb.prepare_warp({}, pos)
    users, items = d['users'].tolist(), d['items'].tolist()
    u, _ = om.user_tower(users, 1.0, None)
    logits = emb.get_prediction(u)
    targets = np.asarray(emb.target_mapping([items])[0], dtype=np.int64)
    mask = emb.build_mask(users, 'warp_eval', True)
    margin, rank = O.compute_loss(logits, targets, 'warp_eval', mask)
    np.testing.assert_allclose(margin, d['margin_rank'], rtol=2e-5, atol=1e-5)
    assert np.array_equal(np.asarray(rank), d['true_rank'])


# ------------------------------------------------------------------ LSTM, non-default flags -----------
LSTMX_CASES = sorted(os.path.basename(p)[len('ref_lstmx_'):-4] for p in glob.glob(os.path.join(GOLD, 'ref_lstmx_*.npz')))


def SeqCaseX(name):
    return SeqCase(name, 'lstmx')


@pytest.mark.parametrize('name', LSTMX_CASES)
def test_oracle_reproduces_reference_lstm_run_nondefault_flags(name):
    """output_feat = 0 (score with the id table only) and no_input_item_feature (only the id embedding feeds the
    LSTM): SURVEY 8(f) row 4, pinned on the CPU oracle."""
    import torch
    from oracle.torch_cpu_ref import TorchRefSeq
    c = SeqCaseX(name)
    ref = TorchRefSeq(c.ua, c.ia, {k: v.copy() for k, v in c.params.items()}, c.l2i_d, c.i2l_d, loss=c.loss,
                      keep_prob=c.keep, learning_rate=c.lr, n_sampled=None, dtype=torch.float64, size=c.dim,
                      use_concat=c.use_concat, no_user_id=False, max_gradient_norm=c.clip, item_output=c.sep,
                      withAdagrad=c.adagrad, output_feat=c.output_feat, no_input_item_feature=c.no_input_item_feature,
                      num_layers=c.num_layers)
    for it in range(c.n_steps):
        users, inp, tgt, w, pos = c.batch('step%d' % it)
        ref.pos, ref.pos_eval = pos, pos
        l = ref.step_seq(users, inp, tgt, w, masks=c.masks(it))
        want = float(c.d['losses'][it])
        assert abs(l - want) <= 2e-5 * max(1.0, abs(want)), (name, it, l, want)
        gn = float(c.d['gnorms'][it])
        assert abs(ref.last_gnorm - gn) <= 1e-4 * max(1.0, gn), (name, it, ref.last_gnorm, gn)
    for k, v in c.final.items():
        np.testing.assert_allclose(ref.p[k].detach().numpy().reshape(v.shape), v, rtol=2e-4, atol=2e-5, err_msg=k)
    users, inp, tgt, w, pos = c.batch('eval')
    ref.pos, ref.pos_eval = pos, pos
    ev = ref.step_seq(users, inp, tgt, w, forward_only=True)
    want = float(c.d['eval/loss'])
    assert abs(ev - want) <= 2e-5 * max(1.0, abs(want)), (ev, want)
    assert np.array_equal(ref.topk_seq(users, inp, c.topk), c.d['eval/topk_indexes'])


@pytest.mark.gpu
@pytest.mark.parametrize('name', LSTMX_CASES)
def test_cuda_path_reproduces_reference_lstm_run_nondefault_flags(cuda, name):
    import torch
    from arecsys_b200 import _lib
    from arecsys_b200.attributes.embed_attribute import EmbeddingAttribute
    from arecsys_b200.lstm.seqModel import SeqModel
    c = SeqCaseX(name)
    _lib.exact_fp32 = True
    try:
        params = {k: v.copy() for k, v in c.params.items()}
        emb = EmbeddingAttribute(c.ua, c.ia, c.mb, None, c.T, c.sep, c.i2l_d, c.l2i_d, params=params)
        model = SeqModel(c.buckets, c.dim, c.num_layers, c.clip, c.mb, c.lr, 0.83, emb, withAdagrad=c.adagrad,
                         dropoutRate=c.keep, START_ID=c.START, loss=c.loss, use_concat=c.use_concat,
                         no_user_id=False, topk_n=c.topk, params=params, output_feat=c.output_feat,
                         no_input_item_feature=c.no_input_item_feature)
        for it in range(c.n_steps):
            users, inp, tgt, w, pos = c.batch('step%d' % it)
            emb.prepare_warp(pos, pos)
            l = model.step(None, users, inp, tgt, w, int(c.d['step%d/bucket' % it]),
                           masks=tuple(torch.tensor(m, device='cuda') for m in c.masks(it)))
            want = float(c.d['losses'][it])
            assert abs(l - want) <= 2e-4 * max(1.0, abs(want)), (name, it, l, want)
        dense = model.dense_params()
        for k, v in c.final.items():
            got = (emb.params[k] if k in emb.params else dense[k][0]).cpu().numpy()
            assert np.abs(got.reshape(v.shape) - v).max() <= 2e-3 * max(1.0, np.abs(v).max()), k
    finally:
        _lib.exact_fp32 = False
