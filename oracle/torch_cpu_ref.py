"""PyTorch-CPU restatement of the reference HMF training step.  TEST / BASELINE INFRASTRUCTURE
ONLY (see oracle/np_oracle.py header for how the oracle is pinned: golden vectors produced by the
reference's own model code on oracle/tf1_shim, tests/test_ref_golden.py).

Executes the reference's op sequence *literally* (index_select -> segment-sum (index_add_) ->
div -> mean over attributes -> dropout -> token-score matmul -> index_select -> segment-sum ->
div -> mean -> transpose -> loss -> autograd -> Adagrad), i.e. attributes/embed_attribute.py:
148-220, :350-417, :525-649 and hmf/hmf_model.py:75-151, on the host cores.  Two uses:
  * an independent (autograd) derivation of every gradient the NumPy oracle derives by hand;
  * the timed CPU baseline of bench.py (`cpu_baseline`, `--impl reference`): it omits the TF-1
    session / feed_dict overhead and the two extra mask session.runs, so it is faster than the
    real reference and every reported speed-up is conservative.
"""
import numpy as np
import torch


def _t(a, dtype):
    return torch.as_tensor(np.asarray(a), dtype=dtype)


def _flat(att, i, inds):
    """mulhot_index.py:48-67, vectorised."""
    s = att.mulhot_starts[i][inds].astype(np.int64)
    l = att.mulhot_lengths[i][inds].astype(np.int64)
    seg = np.repeat(np.arange(len(inds), dtype=np.int64), l)
    off = np.concatenate([[0], np.cumsum(l)[:-1]])
    pos = np.arange(int(l.sum()), dtype=np.int64) - np.repeat(off, l) + np.repeat(s, l)
    return torch.from_numpy(att.features_mulhot[i][pos].astype(np.int64)), torch.from_numpy(seg), l


def seg_sum(data, seg, n):
    out = torch.zeros((n,) + tuple(data.shape[1:]), dtype=data.dtype)
    return out.index_add_(0, seg, data)


class TorchRefHMF(object):
    def __init__(self, user_attributes, item_attributes, params, logit_ind2item_ind, item_ind2logit_ind,
                 loss='ce', nonlinear='linear', keep_prob=1.0, learning_rate=0.1, n_sampled=None,
                 loss_func='log', loss_exp_p=1.005, dtype=torch.float32):
        self.ua, self.ia = user_attributes, item_attributes
        self.dtype = dtype
        self.p = {k: _t(v, dtype).clone().requires_grad_(True) for k, v in params.items()}
        self.acc = {k: torch.full_like(v, 0.1) for k, v in self.p.items()}
        self.l2i, self.i2l = logit_ind2item_ind, item_ind2logit_ind
        self.V = len(logit_ind2item_ind)
        self.loss, self.nonlinear, self.keep_prob, self.lr = loss, nonlinear, keep_prob, learning_rate
        self.n_sampled, self.loss_func, self.exp_p = n_sampled, loss_func, loss_exp_p
        ia = item_attributes
        self.full = ([torch.from_numpy(np.asarray(a, dtype=np.int64)) for a in ia.full_cat_tr],
                     [torch.from_numpy(np.asarray(a, dtype=np.int64)) for a in ia.full_values_tr],
                     [torch.from_numpy(np.asarray(a, dtype=np.int64)) for a in ia.full_segids_tr],
                     [_t(a, dtype).reshape(-1, 1) for a in ia.full_lengths_tr])
        self.sampled = None
        self.pos, self.pos_eval = None, None
        self.touched = {}

    # embed_attribute.py:350-417
    def get_embedded(self, prefix, att, inds, with_bias):
        inds = np.asarray(inds, dtype=np.int64)
        mb = len(inds)
        outs, biases = [], []
        for i in range(att.num_features_cat):
            tok = torch.from_numpy(att.features_cat[i][inds].astype(np.int64))
            self.touched['%sembed_cat_%d' % (prefix, i)] = tok
            outs.append(self.p['%sembed_cat_%d' % (prefix, i)].index_select(0, tok))
            if with_bias:
                biases.append(self.p['%s_bias_cat_%d' % (prefix, i)].index_select(0, tok))
        for i in range(att.num_features_mulhot):
            idx, seg, l = _flat(att, i, inds)
            lengs = torch.from_numpy(l).to(self.dtype).reshape(mb, 1)
            self.touched['%sembed_mulhot_%d' % (prefix, i)] = idx
            flat = self.p['%sembed_mulhot_%d' % (prefix, i)].index_select(0, idx)
            outs.append(seg_sum(flat, seg, mb) / lengs)
            if with_bias:
                bflat = self.p['%s_bias_mulhot_%d' % (prefix, i)].index_select(0, idx)
                biases.append(seg_sum(bflat, seg, mb) / lengs)
        bias = torch.stack(biases, 0).mean(0).reshape(-1) if with_bias else None
        return outs, bias

    # embed_attribute.py:320-348
    def pass_sampled_items(self, item_sampled):
        ia = self.ia
        inds = np.asarray(item_sampled, dtype=np.int64)
        cat = [torch.from_numpy(ia.features_cat[i][inds].astype(np.int64)) for i in range(ia.num_features_cat)]
        val, seg, leng = [], [], []
        for i in range(ia.num_features_mulhot):
            a, b, l = _flat(ia, i, inds)
            val.append(a); seg.append(b)
            leng.append(torch.from_numpy(l).to(self.dtype).reshape(-1, 1))
        self.sampled = (cat, val, seg, leng)
        self.sampled_id2idx = {int(v): k for k, v in enumerate(inds)}

    # embed_attribute.py:148-206 (output_feat = 1)
    def get_prediction(self, u, pool='full'):
        ia = self.ia
        cat, val, seg, leng = self.full if pool == 'full' else self.sampled
        V = self.V if pool == 'full' else self.n_sampled
        innerps = []
        for i in range(ia.num_features_cat):
            innerp = self.p['itemembed_cat_%d' % i] @ u.t() + self.p['item_bias_cat_%d' % i]
            innerps.append(innerp.index_select(0, cat[i]))
        for i in range(ia.num_features_mulhot):
            innerp = self.p['itemembed_mulhot_%d' % i] @ u.t() + self.p['item_bias_mulhot_%d' % i]
            innerps.append(seg_sum(innerp.index_select(0, val[i]), seg[i], V) / leng[i])
        return torch.stack(innerps, 0).mean(0).t()

    # embed_attribute.py:208-220
    def get_target_score(self, u, inds):
        outs, bias = self.get_embedded('item', self.ia, inds, True)
        return (u * torch.stack(outs, 0).mean(0)).sum(1) + bias

    def build_mask(self, user_input, loss, forward_only):
        V = self.n_sampled if loss == 'mw' else self.V
        mask = np.ones((len(user_input), V), dtype=bool)
        item_set = self.pos_eval if forward_only else self.pos
        for c, uu in enumerate(user_input):
            if uu in item_set:
                for v in item_set[uu]:
                    if loss == 'mw':
                        if v in self.sampled_id2idx:
                            mask[c, self.sampled_id2idx[v]] = False
                    else:
                        mask[c, self.i2l[v]] = False
        return torch.from_numpy(mask)

    # embed_attribute.py:525-649
    def compute_loss(self, logits, target, loss, mask):
        mb = logits.shape[0]
        if loss == 'ce':
            return torch.nn.functional.cross_entropy(logits, target, reduction='none')
        zero = torch.zeros_like(logits)
        if loss == 'mw':
            t = torch.where(mask, logits - target.reshape(mb, 1) + 1, zero)
            return torch.log(1 + torch.relu(t).sum(1))
        tl = logits.gather(1, target.reshape(mb, 1))
        if loss == 'warp':
            t = torch.where(mask, logits - tl + 1, zero)
            return torch.log(1 + torch.relu(t).sum(1))
        if loss in ('rs', 'rs-sig'):
            err = torch.relu(logits - tl + 1)
        else:
            err = torch.sigmoid(logits - tl)
        em = torch.where(mask, err, zero)
        if loss == 'rs-sig':
            em = torch.sigmoid(em) * 2 - 1
        s = em.sum(1)
        if loss == 'bbpr':
            return s
        f, p = self.loss_func, self.exp_p
        if f == 'log':
            return torch.log(1 + s)
        if f == 'exp':
            return 1 - torch.pow(torch.tensor(p, dtype=self.dtype), -s)
        if f == 'poly':
            return torch.pow(s, p)
        if f == 'poly2':
            return torch.pow(1 + s, p)
        if f == 'linear':
            return s
        return s * s

    def _drop(self, x, keep, mask):
        if keep == 1.0:
            return x
        if mask is None:
            mask = torch.floor(torch.rand_like(x) + keep)
        return x / keep * _t(mask, self.dtype)

    def forward(self, user_input, item_input, forward_only=False, masks=None):
        keep = 1.0 if forward_only else self.keep_prob
        outs, _ = self.get_embedded('user', self.ua, user_input, False)
        u = torch.stack(outs, 0).mean(0)
        mk = (lambda k: masks[k] if masks is not None else None)
        if self.nonlinear in ('relu', 'tanh'):
            act = torch.relu if self.nonlinear == 'relu' else torch.tanh
            h0 = self._drop(act(u), keep, mk(0))
            h1 = self._drop(act(h0 @ self.p['w1'] + self.p['b1']), keep, mk(1))
            u = self._drop(act(h1 @ self.p['w2'] + self.p['b2']), keep, mk(2))
        else:
            u = self._drop(u, keep, mk(0))
        eff = 'warp' if (self.loss == 'mw' and forward_only) else self.loss
        mask = self.build_mask(user_input, eff, forward_only) if eff != 'ce' else None
        if self.loss == 'mw' and forward_only:       # reference quirk: the 'warp' mask is never set for mw eval
            mask = torch.ones((len(user_input), self.V), dtype=torch.bool)
        if eff == 'mw':
            logits = self.get_prediction(u, 'sampled')
            ts = self.get_target_score(u, item_input)
            bl = self.compute_loss(logits, ts, 'mw', mask)
        else:
            logits = self.get_prediction(u)
            tgt = torch.as_tensor([self.i2l[int(v)] for v in item_input], dtype=torch.int64)
            bl = self.compute_loss(logits, tgt, eff, mask)
        return bl.mean(), logits

    def step(self, user_input, item_input, item_sampled=None, forward_only=False, masks=None):
        if item_sampled is not None and self.loss == 'mw':
            self.pass_sampled_items(item_sampled)
        if forward_only:
            with torch.no_grad():
                return float(self.forward(user_input, item_input, True, masks)[0])
        loss, _ = self.forward(user_input, item_input, False, masks)
        names = list(self.p.keys())
        grads = torch.autograd.grad(loss, [self.p[k] for k in names], allow_unused=True)
        self.last_grads = {}
        with torch.no_grad():
            for k, g in zip(names, grads):
                if g is None:
                    continue
                self.last_grads[k] = g
                if k.startswith('userembed') and k in self.touched:
                    # tables reached only through lookups get IndexedSlices -> sparse apply in TF
                    rows = torch.unique(self.touched[k])
                    gr = g.index_select(0, rows)
                    a = self.acc[k].index_select(0, rows) + gr * gr
                    self.acc[k].index_copy_(0, rows, a)
                    self.p[k].index_copy_(0, rows, self.p[k].index_select(0, rows) - self.lr * gr / torch.sqrt(a))
                    continue
                self.acc[k] += g * g
                self.p[k] -= self.lr * g / torch.sqrt(self.acc[k])
        return float(loss.detach())


class TorchRefSeq(TorchRefHMF):
    """lstm/seqModel.py:24-604 restated with autograd: inputs = mean(user, item) (or the concat
    projections), DropoutWrapper(in) -> LSTMCell -> DropoutWrapper(out), static unroll from a zero
    state, per-step literal get_prediction + loss, sequence_loss, clip_by_global_norm, Adagrad/SGD.
    params additionally hold 'lstm_w' [d_in+H, 4H], 'lstm_b' [4H] (+ 'w_input_user'/'w_input_item')."""

    def __init__(self, *a, size=8, use_concat=False, no_user_id=False, no_input_item_feature=False,
                 max_gradient_norm=5.0, withAdagrad=True, item_output=False, output_feat=1, num_layers=1, **kw):
        super(TorchRefSeq, self).__init__(*a, **kw)
        self.num_layers = num_layers            # MultiRNNCell([DropoutWrapper(LSTMCell, in)] * n) then DropoutWrapper(out): seqModel.py:99-103
        self.output_feat = output_feat          # 0: score with the id table only (embed_attribute.py:164-165)
        self.size, self.use_concat, self.no_user_id = size, use_concat, no_user_id
        self.no_input_item_feature = no_input_item_feature
        self.clip, self.withAdagrad, self.item_output = max_gradient_norm, withAdagrad, item_output
        self.slices = []

    def _emb(self, prefix, att, inds, concat, no_id=False, no_attribute=False):
        """_get_embedded (embed_attribute.py:350-417) keeping every gathered block as a leaf-like
        tensor so that the IndexedSlices values of tf.gradients can be read back."""
        inds = np.asarray(inds, dtype=np.int64)
        mb = len(inds)
        if no_id and att.num_features_cat == 1:
            d = self.p['%sembed_cat_0' % prefix].shape[1]
            return torch.zeros((mb, d), dtype=self.dtype)
        outs = []
        n1 = 1 if no_attribute else att.num_features_cat
        n2 = 0 if no_attribute else att.num_features_mulhot
        for i in range(n1):
            if no_id and i == 0:
                continue
            name = '%sembed_cat_%d' % (prefix, i)
            tok = torch.from_numpy(att.features_cat[i][inds].astype(np.int64))
            rows = self.p[name].index_select(0, tok)
            if rows.requires_grad:
                rows.retain_grad()
            self.slices.append((name, rows))
            outs.append(rows)
        for i in range(n2):
            name = '%sembed_mulhot_%d' % (prefix, i)
            idx, seg, l = _flat(att, i, inds)
            rows = self.p[name].index_select(0, idx)
            if rows.requires_grad:
                rows.retain_grad()
            self.slices.append((name, rows))
            outs.append(seg_sum(rows, seg, mb) / torch.from_numpy(l).to(self.dtype).reshape(mb, 1))
        return torch.cat(outs, 1) if concat else torch.stack(outs, 0).mean(0)

    def _pred(self, u, pool):
        pre = 'item_output' if self.item_output else 'item'
        ia = self.ia
        cat, val, seg, leng = self.full if pool == 'full' else self.sampled
        V = self.V if pool == 'full' else self.n_sampled
        innerps = []
        n1 = 1 if self.output_feat == 0 else ia.num_features_cat
        n2 = 0 if self.output_feat == 0 else ia.num_features_mulhot
        if not hasattr(self, '_pred_tables'):
            self._pred_tables = set()        # tables that feed the scoring matmul: TF holds their gradient dense
        self._pred_tables.update(['%sembed_cat_%d' % (pre, i) for i in range(n1)] +
                                 ['%sembed_mulhot_%d' % (pre, i) for i in range(n2)])
        for i in range(n1):
            innerp = self.p['%sembed_cat_%d' % (pre, i)] @ u.t() + self.p['%s_bias_cat_%d' % (pre, i)]
            innerps.append(innerp.index_select(0, cat[i]))
        for i in range(n2):
            innerp = self.p['%sembed_mulhot_%d' % (pre, i)] @ u.t() + self.p['%s_bias_mulhot_%d' % (pre, i)]
            if self.output_feat in (0, 1):
                innerps.append(seg_sum(innerp.index_select(0, val[i]), seg[i], V) / leng[i])       # :192
            elif self.output_feat == 2:                                                            # :195 segment_max
                rows = innerp.index_select(0, val[i])
                idx = seg[i].reshape(-1, 1).expand_as(rows)
                out = torch.full((V, rows.shape[1]), float('-inf'), dtype=rows.dtype)
                innerps.append(out.scatter_reduce(0, idx, rows, 'amax', include_self=True))
            else:                                                                                  # :197-199
                score_max = innerp.max()
                innerps.append(score_max + torch.log(1 + seg_sum(torch.exp(innerp - score_max).index_select(0, val[i]),
                                                                 seg[i], V)))
        return torch.stack(innerps, 0).mean(0).t()

    def _tscore(self, u, inds):
        pre = 'item_output' if self.item_output else 'item'
        outs, bias = self.get_embedded(pre, self.ia, inds, True)
        return (u * torch.stack(outs, 0).mean(0)).sum(1) + bias

    def forward_seq(self, users, item_inputs, targets, weights, forward_only=False, masks=None):
        T, mb = len(item_inputs), len(users)
        keep = 1.0 if forward_only else self.keep_prob
        self.slices = []
        if self.use_concat:
            ue = self._emb('user', self.ua, users, True, no_id=self.no_user_id)
            uproj = ue @ self.p['w_input_user']
        else:
            ue = self._emb('user', self.ua, users, False, no_id=self.no_user_id)
        L = self.num_layers
        Ws = [(self.p['lstm_w' + ('_%d' % l if l else '')], self.p['lstm_b' + ('_%d' % l if l else '')]) for l in range(L)]
        H = self.size
        hs = [torch.zeros((mb, H), dtype=self.dtype) for _ in range(L)]
        cs = [torch.zeros((mb, H), dtype=self.dtype) for _ in range(L)]
        eff = 'warp' if (self.loss == 'mw' and forward_only) else self.loss
        num = torch.zeros(mb, dtype=self.dtype)
        den = torch.zeros(mb, dtype=self.dtype)
        mask = None
        if eff != 'ce':
            mask = self.build_mask(list(users), eff, forward_only)
        for t in range(T):
            if self.use_concat:
                ie = self._emb('item', self.ia, item_inputs[t], True, no_attribute=self.no_input_item_feature)
                x = uproj + ie @ self.p['w_input_item']
            else:
                ie = self._emb('item', self.ia, item_inputs[t], False, no_attribute=self.no_input_item_feature)
                x = torch.stack([ue, ie], 0).mean(0)
            # masks = (in_masks of layer 0, out_masks[, in_masks of layer 1, ...]): every layer has its own input
            # dropout, the stack one output dropout
            for l in range(L):
                mk = None
                if masks is not None:
                    mk = masks[0][t] if l == 0 else masks[1 + l][t]
                x = self._drop(x, keep, mk)
                W, b = Ws[l]
                z = torch.cat([x, hs[l]], 1) @ W + b
                i, j, f, o = torch.split(z, H, dim=1)
                cs[l] = torch.sigmoid(f + 1.0) * cs[l] + torch.sigmoid(i) * torch.tanh(j)
                hs[l] = torch.sigmoid(o) * torch.tanh(cs[l])
                x = hs[l]
            out = self._drop(x, keep, masks[1][t] if masks is not None else None)
            wt = torch.as_tensor(np.asarray(weights[t]), dtype=self.dtype)
            if eff == 'mw':
                logits = self._pred(out, 'sampled')
                bl = self.compute_loss(logits, self._tscore(out, targets[t]), 'mw', mask)
            else:
                logits = self._pred(out, 'full')
                tg = torch.as_tensor([self.i2l[int(v)] for v in targets[t]], dtype=torch.int64)
                bl = self.compute_loss(logits, tg, eff, mask)
            num = num + bl * wt
            den = den + wt
        return (num / (den + 1e-12)).sum()

    def topk_seq(self, users, item_inputs, k):
        """Per-position top-k of softmax(full logits) (seqModel.py:514-519), keep_prob 1: int64 [T, mb, k];
        ties -> lower index first (tf.nn.top_k)."""
        T, mb = len(item_inputs), len(users)
        with torch.no_grad():
            if self.use_concat:
                uproj = self._emb('user', self.ua, users, True, no_id=self.no_user_id) @ self.p['w_input_user']
            else:
                ue = self._emb('user', self.ua, users, False, no_id=self.no_user_id)
            H, L = self.size, self.num_layers
            Ws = [(self.p['lstm_w' + ('_%d' % l if l else '')], self.p['lstm_b' + ('_%d' % l if l else '')]) for l in range(L)]
            hs = [torch.zeros((mb, H), dtype=self.dtype) for _ in range(L)]
            cs = [torch.zeros((mb, H), dtype=self.dtype) for _ in range(L)]
            out = []
            for t in range(T):
                if self.use_concat:
                    x = uproj + self._emb('item', self.ia, item_inputs[t], True,
                                          no_attribute=self.no_input_item_feature) @ self.p['w_input_item']
                else:
                    ie = self._emb('item', self.ia, item_inputs[t], False, no_attribute=self.no_input_item_feature)
                    x = torch.stack([ue, ie], 0).mean(0)
                for l in range(L):
                    W, b = Ws[l]
                    z = torch.cat([x, hs[l]], 1) @ W + b
                    i, j, f, o = torch.split(z, H, dim=1)
                    cs[l] = torch.sigmoid(f + 1.0) * cs[l] + torch.sigmoid(i) * torch.tanh(j)
                    hs[l] = torch.sigmoid(o) * torch.tanh(cs[l])
                    x = hs[l]
                prob = torch.softmax(self._pred(x, 'full'), 1)
                out.append(torch.sort(prob, dim=1, descending=True, stable=True)[1][:, :k])
            return torch.stack(out, 0).numpy()

    def step_seq(self, users, item_inputs, targets, weights, item_sampled=None, forward_only=False, masks=None):
        if item_sampled is not None and self.loss == 'mw':
            self.pass_sampled_items(item_sampled)
        if forward_only:
            with torch.no_grad():
                return float(self.forward_seq(users, item_inputs, targets, weights, True, masks))
        for v in self.p.values():
            v.grad = None
        loss = self.forward_seq(users, item_inputs, targets, weights, False, masks)
        loss.backward()
        pre = 'item_output' if self.item_output else 'item'
        # tables reached ONLY through lookups keep IndexedSlices gradients: their norm is taken over
        # the un-merged slice values; every other gradient is dense
        sparse_only = set(n for n, _ in self.slices if n not in getattr(self, '_pred_tables', ()))
        sumsq = 0.0
        for k, v in self.p.items():
            if v.grad is None or k in sparse_only:
                continue
            sumsq += float((v.grad * v.grad).sum())
        for n, rows in self.slices:
            if n in sparse_only and rows.grad is not None:
                sumsq += float((rows.grad * rows.grad).sum())
        norm = sumsq ** 0.5
        scale = self.clip / max(norm, self.clip)
        self.last_gnorm = norm
        with torch.no_grad():
            for k, v in self.p.items():
                if v.grad is None:
                    continue
                g = v.grad * scale
                if self.withAdagrad:
                    self.acc[k] += g * g
                    v -= self.lr * g / torch.sqrt(self.acc[k])
                else:
                    v -= self.lr * g
        return float(loss.detach())


class TorchRefCbow(TorchRefSeq):
    """word2vec/cbow_model.py:76-136 restated with autograd: h = dropout(mean(user, mean_k item_k)),
    literal get_prediction over the (separate) output tables, loss, Adagrad."""

    def __init__(self, *a, ni=2, sg=False, **kw):
        super(TorchRefCbow, self).__init__(*a, **kw)
        self.ni = ni
        self.sg = sg          # skip-gram (word2vec/skipgram_model.py:86-88): the TRAINING tower sees input 0 only

    def recommend(self, users, item_inputs, k):
        """top-k logit indices of logits_test (cbow_model.py:95-104,138-139): no dropout, user embedding alone
        when n_input_items == 0; ties -> lower index first."""
        n_input = max(self.ni, 1)
        with torch.no_grad():
            ue = self._emb('user', self.ua, users, False)
            if self.ni == 0:
                x = ue
            else:
                its = torch.stack([self._emb('item', self.ia, item_inputs[j], False) for j in range(n_input)], 0).mean(0)
                x = torch.stack([ue, its], 0).mean(0)
            return torch.sort(self._pred(x, 'full'), dim=1, descending=True, stable=True)[1][:, :k].numpy()

    def step_cbow(self, users, item_inputs, item_outputs, item_sampled=None, forward_only=False, mask=None):
        if item_sampled is not None and self.loss == 'mw':
            self.pass_sampled_items(item_sampled)
        n_input = max(self.ni, 1)
        keep = 1.0 if forward_only else self.keep_prob

        def fwd():
            self.slices = []
            ue = self._emb('user', self.ua, users, False)
            n_in = 1 if (self.sg and not forward_only) else n_input
            its = torch.stack([self._emb('item', self.ia, item_inputs[k], False) for k in range(n_in)], 0).mean(0)
            if forward_only and self.ni == 0:
                x = ue
            else:
                x = torch.stack([ue, its], 0).mean(0)
            h = self._drop(x, keep, mask)
            eff = 'warp' if (self.loss == 'mw' and forward_only) else self.loss
            m = self.build_mask(list(users), eff, forward_only) if eff != 'ce' else None
            if eff == 'mw':
                bl = self.compute_loss(self._pred(h, 'sampled'), self._tscore(h, item_outputs), 'mw', m)
            else:
                tg = torch.as_tensor([self.i2l[int(v)] for v in item_outputs], dtype=torch.int64)
                bl = self.compute_loss(self._pred(h, 'full'), tg, eff, m)
            return bl.mean()
        if forward_only:
            with torch.no_grad():
                return float(fwd())
        for v in self.p.values():
            v.grad = None
        loss = fwd()
        loss.backward()
        with torch.no_grad():
            for k, v in self.p.items():
                if v.grad is None:
                    continue
                self.acc[k] += v.grad * v.grad
                v -= self.lr * v.grad / torch.sqrt(self.acc[k])
        return float(loss.detach())
