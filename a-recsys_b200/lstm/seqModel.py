"""Bucketed LSTM next-item model on B200 — same constructor, step / step_recommend / get_batch /
get_batch_recommend protocol as the reference SeqModel (lstm/seqModel.py:24-521), executed
eagerly through libarx_b200.so.

Per bucket of length T (all T steps batched where the maths allows it):
  x_t   = mean(user_emb, mean_f item_emb_f(input_t))          K1+K2 over T*mb bags  (:148-156)
          or  user_cat W_u + item_cat_t W_i  with use_concat    (:131-146)
  h_t   = LSTM(dropout(x_t)); out_t = dropout(h_t)             K8  (:99-103,:477)
  s_t   = out_t P^T + beta  (pooled catalog / sampled pool)    K3  (:480-493)
  loss  = sum_b [ sum_t w_bt l_bt / (sum_t w_bt + 1e-12) ]     K5/K6, sequence_loss (:571-604)
  clip_by_global_norm(5.0) -> Adagrad / SGD                    (:173-182)
One persistent-shape kernel set serves every bucket: buckets are only a host batching policy.
"""
import os
import random

import numpy as np
import torch

from .. import _lib
from .._lib import POOL_MEAN, POOL_CONCAT, OPT_ADAGRAD, OPT_SGD, call, ptr
from ..attributes import embed_attribute
from ..hmf.hmf_model import _Var
from .lstm_layer import LSTMLayer, LSTMStack


class _Saver(object):
    def __init__(self, model):
        self.model = model

    def save(self, sess, path, global_step=None, write_meta_graph=False):
        m = self.model
        if global_step is not None:
            path = '%s-%d' % (path, global_step)
        e = m.embeddingAttribute
        torch.save({'params': {k: v.cpu() for k, v in e.params.items()},
                    'accs': {k: v.cpu() for k, v in e.accs.items()},
                    'dense': {k: v[0].cpu() for k, v in m.dense_params().items()},
                    'dense_acc': {k: v.cpu() for k, v in m.dense_acc.items()},
                    'learning_rate': m.learning_rate.eval(), 'global_step': m.global_step.eval()}, path)
        with open(os.path.join(os.path.dirname(path), 'checkpoint'), 'w') as f:
            f.write('model_checkpoint_path: "%s"\n' % os.path.basename(path))
        return path

    def restore(self, sess, path):
        m = self.model
        st = torch.load(path, map_location='cpu')
        e = m.embeddingAttribute
        for k, v in st['params'].items():
            e.params[k].copy_(v)
        for k, v in st['accs'].items():
            e.accs[k].copy_(v)
        for k, v in st['dense'].items():
            m.dense_params()[k][0].copy_(v)
        for k, v in st['dense_acc'].items():
            m.dense_acc[k].copy_(v)
        m.learning_rate.assign(st['learning_rate'])
        m.global_step.assign(st['global_step'])


class SeqModel(object):
    def __init__(self, buckets, size, num_layers, max_gradient_norm, batch_size, learning_rate,
                 learning_rate_decay_factor, embeddingAttribute, withAdagrad=True, num_samples=512,
                 forward_only=False, dropoutRate=1.0, START_ID=0, loss="ce", devices="", run_options=None,
                 run_metadata=None, use_concat=True, output_feat=1, no_input_item_feature=False,
                 no_user_id=True, topk_n=30, dtype=torch.float32, seed=None, params=None):
        self.embeddingAttribute = embeddingAttribute
        self.buckets = buckets
        self.START_ID = START_ID
        self.PAD_ID = START_ID
        self.USER_PAD_ID = 0
        self.batch_size = batch_size
        self.loss = loss
        self.devices = devices
        self.output_feat = output_feat
        self.no_input_item_feature = no_input_item_feature
        self.no_user_id = no_user_id
        self.topk_n = topk_n
        self.use_concat = use_concat
        self.size = size
        self.max_gradient_norm = max_gradient_norm
        self.withAdagrad = withAdagrad
        self.num_layers = max(1, int(num_layers))
        if loss not in ('ce', 'warp', 'mw'):
            print('Error: not implemented other loss!!')
            exit(1)
        self.dropoutRate = float(dropoutRate)
        self._keep_train = float(dropoutRate)
        self.learning_rate = _Var(float(learning_rate))
        self._decay = learning_rate_decay_factor
        self.learning_rate_decay_op = lambda: self.learning_rate.assign(self.learning_rate.eval() * self._decay)
        self.global_step = _Var(0)
        # dropout10_op / dropoutAssign_op of the reference (seqModel.py:88-91)
        self.dropout10_op = lambda: setattr(self, 'dropoutRate', 1.0)
        self.dropoutAssign_op = lambda: setattr(self, 'dropoutRate', self._keep_train)

        m = embeddingAttribute
        self.device = m.device
        gen = torch.Generator(device='cpu')
        gen.manual_seed(2 if seed is None else seed + 2)
        p = params or {}
        self.dense, self.dense_grad = {}, {}
        if use_concat:
            ue = m.get_user_model_size(no_id=no_user_id, concat=True)
            ie = m.get_item_model_size(concat=True) if not no_input_item_feature else m.dim
            for name, shape in (('w_input_user', (ue, size)), ('w_input_item', (ie, size))):
                if name in p:
                    w = torch.as_tensor(np.asarray(p[name]), dtype=torch.float32).reshape(shape).clone()
                else:
                    lim = (6.0 / (shape[0] + shape[1])) ** 0.5
                    w = (torch.rand(shape, generator=gen) * 2 - 1) * lim
                self.dense[name] = w.to(self.device).contiguous()
                self.dense_grad[name] = torch.zeros_like(self.dense[name])
            d_in = size
        else:
            d_in = m.dim
        self.cell = LSTMStack(d_in, size, self.num_layers, self.device, gen, p)       # MultiRNNCell (seqModel.py:99-103)
        self.dense_acc = {k: torch.full_like(v[0], embed_attribute.ADAGRAD_INIT_ACC)
                          for k, v in self.dense_params().items()}
        if self.loss in ["warp", "mw"]:
            self.set_mask, self.reset_mask = m.get_warp_mask()
        self.saver = _Saver(self)
        self.max_score_bytes = 2 << 30          # scores are materialised in time chunks of <= 2 GB

    def dense_params(self):
        d = {k: (v, self.dense_grad[k]) for k, v in self.dense.items()}
        d.update(self.cell.parameters())
        return d

    # ------------------------------------------------------------------ forward pieces --
    def _inputs(self, users, item_ids, T, mb):
        """LSTM inputs [T, mb, d_in] (seqModel.py:126-156) + the context the backward needs."""
        m = self.embeddingAttribute
        ua = m.user_attributes
        zero_user = self.no_user_id and ua.num_features_cat == 1            # embed_attribute.py:356-366
        if self.use_concat:
            ucat = None
            if not zero_user:
                ucat, _, urng = m.pool('user', users, POOL_CONCAT, False, no_id=self.no_user_id)
            icat, _, irng = m.pool('item', item_ids, POOL_CONCAT, False, no_attribute=self.no_input_item_feature)
            x = torch.empty((T * mb, self.size), dtype=torch.float32, device=self.device)
            _lib.gemm(icat, self.dense['w_input_item'], x, T * mb, self.size, icat.shape[1], 0, 0)
            if ucat is not None:
                ux = torch.empty((mb, self.size), dtype=torch.float32, device=self.device)
                _lib.gemm(ucat, self.dense['w_input_user'], ux, mb, self.size, ucat.shape[1], 0, 0)
                call('arx_axpby_rows', x.data_ptr(), ux.data_ptr(), 1.0, 1.0, T * mb, mb, self.size, x.data_ptr())
            ctx = ('concat', ucat, icat, None if zero_user else urng, irng)
        else:
            iemb, _, irng = m.pool('item', item_ids, POOL_MEAN, False, no_attribute=self.no_input_item_feature)
            x = torch.empty_like(iemb)
            if zero_user:
                call('arx_axpby_rows', iemb.data_ptr(), None, 0.5, 0.0, T * mb, mb, m.dim, x.data_ptr())
                urng = None
            else:
                uemb, _, urng = m.pool('user', users, POOL_MEAN, False, no_id=self.no_user_id)
                call('arx_axpby_rows', iemb.data_ptr(), uemb.data_ptr(), 0.5, 0.5, T * mb, mb, m.dim, x.data_ptr())
            ctx = ('mean', None, None, urng, irng)
        return x.view(T, mb, -1), ctx

    def _inputs_backward(self, ctx, dX, users, item_ids, T, mb):
        m = self.embeddingAttribute
        kind, ucat, icat, urng, irng = ctx
        dX = dX.reshape(T * mb, -1)
        if kind == 'concat':
            di = torch.empty_like(icat)
            _lib.gemm(dX, self.dense['w_input_item'], di, T * mb, icat.shape[1], self.size, 0, 1)
            _lib.gemm(icat, dX, self.dense_grad['w_input_item'], icat.shape[1], self.size, T * mb, 1, 0)
            m.push_grad('item', irng, item_ids, POOL_CONCAT, di)
            if ucat is not None:
                dux = torch.empty((mb, self.size), dtype=torch.float32, device=self.device)
                call('arx_sum_over_steps', dX.data_ptr(), T, mb, self.size, 1.0, dux.data_ptr())
                du = torch.empty_like(ucat)
                _lib.gemm(dux, self.dense['w_input_user'], du, mb, ucat.shape[1], self.size, 0, 1)
                _lib.gemm(ucat, dux, self.dense_grad['w_input_user'], ucat.shape[1], self.size, mb, 1, 0)
                m.push_grad('user', urng, users, POOL_CONCAT, du)
            else:
                self.dense_grad['w_input_user'].zero_()
        else:
            di = torch.empty_like(dX)
            call('arx_axpby_rows', dX.data_ptr(), None, 0.5, 0.0, T * mb, mb, m.dim, di.data_ptr())
            m.push_grad('item', irng, item_ids, POOL_MEAN, di)
            if urng is not None:
                du = torch.empty((mb, m.dim), dtype=torch.float32, device=self.device)
                call('arx_sum_over_steps', dX.data_ptr(), T, mb, m.dim, 0.5, du.data_ptr())
                m.push_grad('user', urng, users, POOL_MEAN, du)

    # ------------------------------------------------------------------ step -----------
    def step(self, session, user_input, item_inputs, targets, target_weights, bucket_id, item_sampled=None,
             item_sampled_id2idx=None, forward_only=False, recommend=False, masks=None, sync=True):
        """seqModel.py:289-324.  item_inputs / targets / target_weights: time-major lists [T][mb].
        Returns the summed sequence loss (a Python float).  masks = (in_mask, out_mask) injects the
        dropout masks for parity runs."""
        m = self.embeddingAttribute
        T = self.buckets[bucket_id]
        mb = len(user_input)
        dev = self.device
        m.add_input({}, user_input, None, item_sampled=item_sampled, item_sampled_id2idx=item_sampled_id2idx,
                    forward_only=forward_only, recommend=recommend, loss=self.loss)
        users = m.u_indices['input']
        # time-major lists [T][mb] as the reference feeds them, or [T, mb] device tensors (utils/device_batch.py)
        as_ids = lambda x: (x[:T].reshape(-1) if isinstance(x, torch.Tensor) else np.asarray(x[:T], dtype=np.int32).reshape(-1))
        item_ids = m._ids(as_ids(item_inputs))
        tgt_items = m._ids(as_ids(targets))
        tgt = m.item2logit_dev[tgt_items.long()].contiguous()                 # target_mapping (:294)
        if isinstance(target_weights, torch.Tensor):
            w = target_weights[:T].to(device=dev, dtype=torch.float32)
        else:
            w = torch.as_tensor(np.asarray(target_weights[:T], dtype=np.float32)).to(dev)     # [T, mb]
        row_scale = (w / (w.sum(0, keepdim=True) + 1e-12)).reshape(-1).contiguous()        # sequence_loss
        keep = 1.0 if forward_only else self.dropoutRate
        train = not forward_only

        X, ictx = self._inputs(users, item_ids, T, mb)
        # masks = (input masks of layer 0, output masks of the stack[, input masks of layer 1, ...])
        in_mask, out_mask = (masks[0], masks[1]) if masks is not None else (None, None)
        Hout = self.cell.forward(X, keep, in_mask, out_mask, tuple(masks[2:]) if masks is not None else ())   # [T, mb, H]
        Hf = Hout.reshape(T * mb, self.size)

        eff = self.loss
        if eff == 'mw' and forward_only:
            eff = 'warp'                                                      # losses_full (:311,:510)
        pre = m._out_prefix()
        pool = 'sampled' if eff == 'mw' else 'full'
        nonlinear = self.output_feat in (2, 3)      # max / log-sum-exp pooling of the token scores: literal order (K3m)
        fused = m.fused_ce(Hf, tgt, row_scale, train, pool, self.output_feat) if (eff == 'ce' and not nonlinear) else None
        if fused is None and eff in ('warp', 'rs') and not nonlinear:
            # full-catalog WMRB of all T*mb positions without the [T*mb, V] scores (embed_attribute.py:551-618)
            fused = m.fused_warp(Hf, tgt, eff, 'log', 1.005, row_scale, train, forward_only=forward_only,
                                 pos_rows=users.repeat(T).contiguous(), output_feat=self.output_feat)
        if fused is not None:
            # all T*mb positions scored against the catalog and reduced to the softmax CE on the tensor
            # cores without materialising [T*mb, V] logits (seqModel.py:480-493 + sequence_loss)
            bl, grads = fused
            _, P, beta, cids, _, _ = m._last_pred
            total = (bl * row_scale).sum()
            if not train:
                return float(total.item()) if sync else total
            dH, dP, dbeta = grads
            return self._finish_step(m, pre, cids, dP, dbeta, dH, ictx, users, item_ids, T, mb, total, sync)
        if nonlinear:
            P = beta = cids = None
            N = (m.catalog_ids if pool == 'full' else m.sampled_ids).numel()
        else:
            P, beta, cids = m.pool_catalog(pool, self.output_feat)
            N = P.shape[0]
        users_rep = users.repeat(T) if eff != 'ce' else None
        tscore = dts_all = Pt = None
        if eff == 'mw':
            tscore = m.get_target_score(Hf, tgt_items)                        # :493
            Pt = m._last_target[1]
            dts_all = torch.empty((T * mb,), dtype=torch.float32, device=dev)
        total = torch.zeros((), dtype=torch.float32, device=dev)
        dP = dbeta = None
        if train:
            dH = torch.empty_like(Hf)
            if not nonlinear:
                dP = torch.zeros_like(P)
                dbeta = torch.zeros((N,), dtype=torch.float32, device=dev)
        rows_per_chunk = max(mb, int(self.max_score_bytes // (4 * N)) // mb * mb)
        if nonlinear:
            # the reference scores one time step per get_prediction call (seqModel.py:480-493), and the log-sum-exp
            # pooling is anchored at the maximum of THAT call's token scores (embed_attribute.py:197): log(e^m + sum e^s)
            # depends on m, so the calls must not be merged
            rows_per_chunk = mb
        for r0 in range(0, T * mb, rows_per_chunk):
            r1 = min(T * mb, r0 + rows_per_chunk)
            n = r1 - r0
            if nonlinear:
                S = m.get_prediction(Hf[r0:r1], pool, output_feat=self.output_feat)
            else:
                S = torch.empty((n, N), dtype=torch.float32, device=dev)
                _lib.gemm(Hf[r0:r1], P, S, n, N, self.size, 0, 1, beta)
            bl = m.compute_loss(S, tscore[r0:r1] if eff == 'mw' else tgt[r0:r1], eff,
                                row_scale=row_scale[r0:r1], want_grad=train, forward_only=forward_only,
                                pos_rows=users_rep[r0:r1].contiguous() if users_rep is not None else None)
            total += (bl * row_scale[r0:r1]).sum()
            if train and nonlinear:
                dH[r0:r1] = m.token_prediction_backward(S)         # table / bias gradients accumulate in m.dense_table_grads
                if eff == 'mw':
                    dts_all[r0:r1] = m._last_dtarget
            elif train:
                _lib.gemm(S, P, dH[r0:r1], n, self.size, N, 0, 0)
                dPc = torch.empty_like(P)
                _lib.gemm(S, Hf[r0:r1], dPc, N, self.size, n, 1, 0)
                dP += dPc
                dbc = torch.empty((N,), dtype=torch.float32, device=dev)
                call('arx_colsum', S.data_ptr(), n, N, S.stride(0), dbc.data_ptr())
                dbeta += dbc
                if eff == 'mw':
                    dts_all[r0:r1] = m._last_dtarget
            del S
        if not train:
            return float(total.item()) if sync else total

        if eff == 'mw':
            dPt = torch.empty_like(Pt)
            call('arx_rowdot_bwd', Hf.data_ptr(), Pt.data_ptr(), dts_all.data_ptr(), T * mb, self.size,
                 dH.data_ptr(), dPt.data_ptr())
            m.push_grad(pre, m.sets[pre].attr_range(), tgt_items, POOL_MEAN, dPt, dts_all)
        return self._finish_step(m, pre, cids, dP, dbeta, dH, ictx, users, item_ids, T, mb, total, sync)

    def _finish_step(self, m, pre, cids, dP, dbeta, dH, ictx, users, item_ids, T, mb, total, sync):
        """Backward through the catalog pooling, the LSTM and the input embeddings, then the clipped
        optimizer step (seqModel.py:173-182)."""
        dev = self.device
        if dP is not None:
            rng_out = m.sets[pre].attr_range(no_attribute=(self.output_feat == 0))
            m.push_grad(pre, rng_out, cids, POOL_MEAN, dP, dbeta, plan_key=None)
        dX = self.cell.backward(dH.view(T, mb, self.size))
        self._inputs_backward(ictx, dX, users, item_ids, T, mb)

        # clip_by_global_norm over dense + table gradients (:179-182)
        sumsq = torch.zeros((1,), dtype=torch.float32, device=dev)
        for name, (wt, g) in self.dense_params().items():
            sumsq += (g * g).sum()
        m.sparse_sumsq(sumsq, dense_semantics=(pre,))
        gnorm = torch.sqrt(sumsq)
        clip = float(self.max_gradient_norm)
        scale = (clip / torch.clamp(gnorm, min=clip)).contiguous()
        self.last_gnorm = gnorm
        lr = self.learning_rate.eval()
        opt = OPT_ADAGRAD if self.withAdagrad else OPT_SGD
        m.apply_gradients(lr, opt, grad_scale=scale)
        for name, (wt, g) in self.dense_params().items():
            call('arx_dense_update', wt.data_ptr(), ptr(self.dense_acc[name]), g.contiguous().data_ptr(), wt.numel(),
                 float(lr), scale.data_ptr(), opt)
        self.global_step.assign(self.global_step.eval() + 1)
        return float(total.item()) if sync else total

    def step_recommend(self, session, user_input, item_inputs, positions, bucket_id):
        """seqModel.py:326-353: softmax + top-k at the last valid position of every sequence.
        Returns [(uid, values[topk], indexes[topk])]."""
        m = self.embeddingAttribute
        T = self.buckets[bucket_id]
        mb = len(user_input)
        m.add_input({}, user_input, None, forward_only=True, recommend=True, loss=self.loss)
        users = m.u_indices['input']
        item_ids = m._ids(np.asarray(item_inputs[:T], dtype=np.int32).reshape(-1))
        X, _ = self._inputs(users, item_ids, T, mb)
        Hout = self.cell.forward(X, 1.0)
        pos = torch.as_tensor(np.asarray(positions, dtype=np.int64)).to(self.device)
        hsel = Hout[pos, torch.arange(mb, device=self.device)].contiguous()            # [mb, H]
        if self.output_feat not in (2, 3):
            # scoring + top-k in column blocks of the catalog, softmax denominator accumulated alongside (:514-519)
            idx, val, lse = m.score_topk(hsel, self.topk_n, self.output_feat, want_lse=True)
            prob = torch.exp(val - lse.unsqueeze(1))
            idx, prob = idx.cpu().numpy(), prob.cpu().numpy()
            return [(user_input[i], prob[i, :], idx[i, :]) for i in range(mb)]
        # output_feat 2 / 3: one get_prediction per time step, as the reference's graph has it (the output_feat 3 pooling
        # depends on the maximum token score of the step's whole batch); every sequence keeps the row of its last position
        N = m.catalog_ids.numel()
        S = torch.empty((mb, N), dtype=torch.float32, device=self.device)
        for t in sorted(set(int(p) for p in positions)):
            sel = (pos == t).nonzero().reshape(-1)
            S[sel] = m.get_prediction(Hout[t].contiguous(), 'full', output_feat=self.output_feat)[sel]
        idx = torch.empty((mb, self.topk_n), dtype=torch.int32, device=self.device)
        val = torch.empty((mb, self.topk_n), dtype=torch.float32, device=self.device)
        call('arx_topk_rows', S.data_ptr(), mb, N, S.stride(0), self.topk_n, idx.data_ptr(), val.data_ptr())
        prob = torch.exp(val - torch.logsumexp(S, 1, keepdim=True))                    # tf.nn.softmax then top_k (:515)
        idx, prob = idx.cpu().numpy(), prob.cpu().numpy()
        return [(user_input[i], prob[i, :], idx[i, :]) for i in range(mb)]

    # ------------------------------------------------------------------ batching -------
    def _batch_major(self, l):
        return [[l[j][i] for j in range(len(l))] for i in range(len(l[0]))]

    def get_batch(self, data_set, bucket_id, start_id=None):
        """seqModel.py:356-404: inputs = [START] + seq[:-1] + pad, targets = seq + pad, weights 1/0."""
        length = self.buckets[bucket_id]
        users, item_inputs, item_outputs, weights = [], [], [], []
        for i in range(self.batch_size):
            if start_id is None:
                user, item_seq = random.choice(data_set[bucket_id])
            elif start_id + i < len(data_set[bucket_id]):
                user, item_seq = data_set[bucket_id][start_id + i]
            else:
                user, item_seq = self.USER_PAD_ID, []
            pad_seq = [self.PAD_ID] * (length - len(item_seq))
            if len(item_seq) == 0:
                item_input_seq = [self.START_ID] + pad_seq[1:]
            else:
                item_input_seq = [self.START_ID] + item_seq[:-1] + pad_seq
            users.append(user)
            item_inputs.append(item_input_seq)
            item_outputs.append(item_seq + pad_seq)
            weights.append([1.0] * len(item_seq) + [0.0] * len(pad_seq))
        finished = start_id is not None and start_id + self.batch_size >= len(data_set[bucket_id])
        return users, self._batch_major(item_inputs), self._batch_major(item_outputs), self._batch_major(weights), finished

    def get_batch_recommend(self, data_set, bucket_id, start_id=None):
        """seqModel.py:407-451."""
        length = self.buckets[bucket_id]
        users, item_inputs, positions, valids = [], [], [], []
        for i in range(self.batch_size):
            if start_id is None:
                user, item_seq = random.choice(data_set[bucket_id])
                valid, position = 1, len(item_seq) - 1
            elif start_id + i < len(data_set[bucket_id]):
                user, item_seq = data_set[bucket_id][start_id + i]
                valid, position = 1, len(item_seq) - 1
            else:
                user, item_seq, valid, position = self.USER_PAD_ID, [], 0, length - 1
            users.append(user)
            positions.append(position)
            valids.append(valid)
            item_inputs.append(item_seq + [self.PAD_ID] * (length - len(item_seq)))
        finished = start_id is not None and start_id + self.batch_size >= len(data_set[bucket_id])
        return users, self._batch_major(item_inputs), positions, valids, finished
