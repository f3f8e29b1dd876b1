"""Parity against golden vectors produced by the reference's OWN model code.

`tests/golden/ref_hmf_*.npz` were written by `tests/golden/make_ref_golden.py`, which imports the
unmodified `/root/reference/hmf/hmf_model.py` + `attributes/embed_attribute.py` + `mulhot_index.py`
on top of `oracle/tf1_shim` (a TF-1.0 op restatement) and records what that graph computes.
  * not gpu: the NumPy oracle (`oracle/np_oracle.py`) must reproduce the reference's losses,
    parameters after training, eval loss and top-k on the same inputs / weights / dropout masks;
  * gpu: the CUDA path through the C ABI must do the same (exact-fp32 contractions: 1e-4;
    tcgen05 tf32 contractions: the north star's 1e-3).
Nothing here reads /root/reference at run time.
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
