"""CBOW "linear sequence" model on B200 — same constructor and step() protocol as the reference
(word2vec/cbow_model.py:14-140, word2vec/linear_seq.py:70-120), eager through libarx_b200.so.

  h = dropout( mean( user_emb, mean_{k<ni} mean_f item_emb_f(input_k) ) )     K1+K2   (:81-91)
  logits = h P^T + beta over the catalog, separate output tables by default     K3     (:92, use_sep_item)
  loss ce / warp / bbpr (mw additionally; the reference's mw branch dies with a NameError at :124)
The ni input lookups of a batch are ONE pooling launch over ni*mb bags.
"""
import numpy as np
import torch

from .. import _lib
from .._lib import POOL_MEAN, OPT_ADAGRAD, call
from ..attributes import embed_attribute
from ..hmf.hmf_model import _Var, _Saver


class Model(object):
    # skip-gram (word2vec/skipgram_model.py:86-88) trains on the FIRST input item only; its test tower is CBOW's
    train_first_input_only = False

    def __init__(self, user_size, item_size, size, batch_size, learning_rate, learning_rate_decay_factor,
                 user_attributes=None, item_attributes=None, item_ind2logit_ind=None, logit_ind2item_ind=None,
                 n_input_items=0, loss_function='ce', logit_size_test=None, dropout=1.0, top_N_items=100,
                 use_sep_item=True, n_sampled=None, output_feat=1, indices_item=None, dtype=torch.float32,
                 seed=None, params=None):
        self.user_size, self.item_size = user_size, item_size
        self.top_N_items = top_N_items
        user_attributes.set_model_size(size)
        item_attributes.set_model_size(size)
        self.user_attributes, self.item_attributes = user_attributes, item_attributes
        self.item_ind2logit_ind, self.logit_ind2item_ind = item_ind2logit_ind, logit_ind2item_ind
        self.logit_size = len(logit_ind2item_ind)
        self.indices_item = indices_item if indices_item is not None else range(self.logit_size)
        self.loss_function = loss_function
        self.n_input_items = n_input_items
        self.n_sampled = n_sampled
        self.batch_size = batch_size
        self.size = size
        self.output_feat = output_feat
        if loss_function not in ('warp', 'ce', 'bbpr', 'mw'):
            print("not implemented!")
            exit(-1)
        self.learning_rate = _Var(float(learning_rate))
        self._decay = learning_rate_decay_factor
        self.learning_rate_decay_op = lambda: self.learning_rate.assign(self.learning_rate.eval() * self._decay)
        self.global_step = _Var(0)
        self.dropout = dropout
        self.n_input = max(n_input_items, 1)
        self.att_emb = embed_attribute.EmbeddingAttribute(user_attributes, item_attributes, batch_size, n_sampled,
                                                          self.n_input, use_sep_item, item_ind2logit_ind,
                                                          logit_ind2item_ind, seed=seed, params=params)
        self.device = self.att_emb.device
        self.dense, self.dense_acc = {}, {}
        if loss_function in ('warp', 'mw', 'bbpr'):
            self.set_mask, self.reset_mask = self.att_emb.get_warp_mask()
        self._row_scale = {}
        self.saver = _Saver(self)

    def prepare_warp(self, pos_item_set, pos_item_set_eval):
        self.att_emb.prepare_warp(pos_item_set, pos_item_set_eval)

    def _scale(self, mb):
        if mb not in self._row_scale:
            self._row_scale[mb] = torch.full((mb,), 1.0 / mb, dtype=torch.float32, device=self.device)
        return self._row_scale[mb]

    def step(self, session, user_input, item_input=None, item_output=None, item_sampled=None,
             item_sampled_id2idx=None, forward_only=False, recommend=False, recommend_new=False, loss=None,
             run_op=None, run_meta=None, masks=None, sync=True):
        """linear_seq.py:70-120.  item_input: [ni][mb] lists (time-major), item_output: [mb]."""
        m = self.att_emb
        ni = self.n_input
        if self.train_first_input_only and not (forward_only or recommend):
            ni = 1
        m.add_input({}, user_input, None, item_sampled=item_sampled, item_sampled_id2idx=item_sampled_id2idx,
                    forward_only=forward_only, recommend=recommend, loss=loss)
        users = m.u_indices['input']
        mb = users.numel()
        in_ids = m._ids(np.asarray(item_input[:ni], dtype=np.int32).reshape(-1))             # [ni*mb]
        uemb, _, urng = m.pool('user', users, POOL_MEAN, False)                               # :81
        iemb, _, irng = m.pool('item', in_ids, POOL_MEAN, False)                              # :83-86
        imean = torch.empty((mb, self.size), dtype=torch.float32, device=self.device)
        call('arx_sum_over_steps', iemb.data_ptr(), ni, mb, self.size, 1.0 / ni, imean.data_ptr())
        x = torch.empty_like(imean)
        test_user_only = (forward_only or recommend) and self.n_input_items == 0              # :97-98
        if test_user_only:
            x = uemb
        else:
            call('arx_axpby_rows', imean.data_ptr(), uemb.data_ptr(), 0.5, 0.5, mb, mb, self.size, x.data_ptr())
        keep = 1.0 if (forward_only or recommend) else self.dropout                           # logits_test :103
        h = m.dropout(x, keep, masks[0] if masks else None)
        mask = getattr(m, '_last_dropout_mask', None) if keep != 1.0 else None
        if recommend:
            if recommend_new:
                raise AttributeError("'Model' object has no attribute 'indices_test'")       # linear_seq.py:96
            idx, _, _ = m.score_topk(h, self.top_N_items, self.output_feat)      # column blocks: no [mb, V] scores
            return idx.cpu().numpy()
        out_ids = m._ids(item_output)
        targets = m.item2logit_dev[out_ids.long()].contiguous()
        eff = loss if loss is not None else self.loss_function
        if eff == 'mw' and forward_only:
            eff = 'warp'
        train = not forward_only
        scale = self._scale(mb)
        pre = m._out_prefix()
        fused = None
        if eff == 'mw':
            Ps, bs, sids = m.pool_catalog('sampled', self.output_feat)
            S = Ps.shape[0]
            tscore = m.get_target_score(h, out_ids)
            Pt = m._last_target[1]
            fused = m.fused_mw(h, Ps, bs, tscore, scale, train)     # scoring + WMRB + adjoints on the tensor cores
            if fused is not None:
                batch_loss, fgrads = fused
            else:
                logits = torch.empty((mb, S), dtype=torch.float32, device=self.device)
                _lib.gemm(h, Ps, logits, mb, S, self.size, 0, 1, bs)
                batch_loss = m.compute_loss(logits, tscore, 'mw', row_scale=scale, want_grad=train)
            P = Ps
        else:
            fused = m.fused_ce(h, targets, scale, train, 'full', self.output_feat) if eff == 'ce' else None
            if fused is None and eff in ('warp', 'rs'):                          # full-catalog WMRB, same pipeline
                fused = m.fused_warp(h, targets, eff, 'log', 1.005, scale, train, forward_only=forward_only,
                                     output_feat=self.output_feat)
            if fused is not None:        # scoring + softmax CE / WMRB + adjoints on the tensor cores, no [mb, V] logits
                batch_loss, fgrads = fused
            else:
                logits = m.get_prediction(h, output_feat=self.output_feat)
                batch_loss = m.compute_loss(logits, targets, eff, row_scale=scale, want_grad=train,
                                            forward_only=forward_only)
            P, sids = m._last_pred[1], m._last_pred[3]
        loss_val = batch_loss.mean()
        if train:
            if fused is not None:
                dH, dP, db = fgrads[:3]
            else:
                D = logits
                N = D.shape[1]
                dH = torch.empty_like(h)
                _lib.gemm(D, P, dH, mb, self.size, N, 0, 0)
                dP = torch.empty_like(P)
                _lib.gemm(D, h, dP, N, self.size, mb, 1, 0)
                db = torch.empty((N,), dtype=torch.float32, device=self.device)
                call('arx_colsum', D.data_ptr(), mb, N, D.stride(0), db.data_ptr())
            rng_out = m.sets[pre].attr_range(no_attribute=(self.output_feat == 0))
            m.push_grad(pre, rng_out, sids, POOL_MEAN, dP, db, plan_key='catalog' if eff != 'mw' else None)
            if eff == 'mw':
                dts = fgrads[3] if fused is not None else m._last_dtarget
                dPt = torch.empty_like(Pt)
                call('arx_rowdot_bwd', h.data_ptr(), Pt.data_ptr(), dts.data_ptr(), mb, self.size, dH.data_ptr(),
                     dPt.data_ptr())
                m.push_grad(pre, m.sets[pre].attr_range(), out_ids, POOL_MEAN, dPt, dts)
            if keep != 1.0:
                dx = torch.empty_like(dH)
                call('arx_scale_mask', dH.data_ptr(), mask.data_ptr(), 1.0 / keep, dH.numel(), dx.data_ptr())
            else:
                dx = dH
            du = torch.empty_like(dx)
            call('arx_axpby_rows', dx.data_ptr(), None, 0.5, 0.0, mb, mb, self.size, du.data_ptr())
            m.push_grad('user', urng, users, POOL_MEAN, du)
            di = (dx * (0.5 / ni)).repeat(ni, 1).contiguous()                 # every input gets dx / (2 ni)
            m.push_grad('item', irng, in_ids, POOL_MEAN, di)
            m.apply_gradients(self.learning_rate.eval(), OPT_ADAGRAD)
            self.global_step.assign(self.global_step.eval() + 1)
        return float(loss_val.item()) if sync else loss_val
