"""NumPy oracle for the A-RecSys training hot path.  TEST INFRASTRUCTURE ONLY.

HOW THIS ORACLE IS PINNED: the reference (skywaLKer518/A-Recsys) ships no tests and no
golden vectors and cannot run as published (Python 2 + TensorFlow 1.0, neither
installable offline).  Its model code parses under Python 3, so
tests/golden/make_ref_golden.py runs the UNMODIFIED reference sources
(hmf/hmf_model.py, lstm/seqModel.py, word2vec/cbow_model.py, attributes/
embed_attribute.py, mulhot_index.py, attribute.py) on oracle/tf1_shim (a lazy-graph
restatement of the TF-1.0 ops they call) and commits what the reference's own graphs
compute as tests/golden/ref_*.npz; tests/test_ref_golden.py holds this oracle (and
the CUDA path) to those vectors.  Scope of the pin: op ORDER and model wiring are the
reference's own code; the per-op TensorFlow-1.0 semantics are
restated in the shim from the public TF docs and are not verifiable offline.
This file is a CPU restatement of the reference's *op sequence* (cited file:line below,
paths relative to the reference root), additionally anchored on (a) the hand-checked
known-answer vectors of SURVEY.md section 8(c) (tests/test_oracle_kat.py), (b)
dataset-derived facts of the bundled MovieLens-1m files (tests/golden/ml1m_facts.json)
and (c) a second, independent derivation of every gradient through torch-CPU float64
autograd (oracle/torch_cpu_ref.py).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this package.  The product (a-recsys_b200/) never does.

Conventions: every function takes a ``dtype`` (float64 = gold, float32 = mimic the
reference's arithmetic type).  Indices are int64 inside the oracle, int32 at the
CUDA boundary; the tests compare index results bit-exactly after casting.
"""
from __future__ import annotations

import numpy as np

UNK_ID = 0    # utils/preprocess.py:11
START_ID = 1  # utils/preprocess.py:12


# --------------------------------------------------------------------------
# attributes/mulhot_index.py
# --------------------------------------------------------------------------
def batch_slice2(target, b, s):
    """mulhot_index.py:48-52 — concat(target[b[i]:b[i]+s[i]] for i in range(l))."""
    target = np.asarray(target)
    if len(b) == 0:
        return np.zeros((0,), dtype=np.int64)
    return np.concatenate([target[int(bi):int(bi) + int(si)] for bi, si in zip(b, s)]).astype(np.int64)


def batch_segids2(s):
    """mulhot_index.py:62-67 — concat(tile([i], s[i]))."""
    s = np.asarray(s, dtype=np.int64)
    return np.repeat(np.arange(len(s), dtype=np.int64), s)


def unsorted_segment_sum(data, seg, n):
    """tf.unsorted_segment_sum: out[seg[k]] += data[k]."""
    out = np.zeros((n,) + data.shape[1:], dtype=data.dtype)
    np.add.at(out, seg, data)
    return out


def segment_max(data, seg, n):
    """tf.segment_max over sorted segment ids (embed_attribute.py:195)."""
    out = np.full((n,) + data.shape[1:], -np.inf, dtype=data.dtype)
    np.maximum.at(out, seg, data)
    return out


# --------------------------------------------------------------------------
# attributes/embed_attribute.py :: EmbeddingAttribute
# --------------------------------------------------------------------------
class OracleEmbeddingAttribute(object):
    """Restates EmbeddingAttribute (embed_attribute.py:19-747) eagerly in NumPy.

    ``params`` is a dict of arrays with the reference's variable names
    (embed_attribute.py:265-306): '{user,item,item_output}embed_{cat,mulhot}_{i}' and
    '{item,item_output}_bias_{cat,mulhot}_{i}' (bias shape [V,1]).
    """

    def __init__(self, user_attributes, item_attributes, mb, n_sampled, params,
                 input_steps=0, item_output=False, item_ind2logit_ind=None,
                 logit_ind2item_ind=None, dtype=np.float64):
        self.user_attributes = user_attributes
        self.item_attributes = item_attributes
        self.batch_size = mb
        self.n_sampled = n_sampled
        self.input_steps = input_steps
        self.item_output = item_output
        self.item_ind2logit_ind = item_ind2logit_ind
        self.logit_ind2item_ind = logit_ind2item_ind
        if logit_ind2item_ind is not None:
            self.logit_size = len(logit_ind2item_ind)
        self.dtype = dtype
        self.p = {k: np.asarray(v, dtype=dtype) for k, v in params.items()}
        self.pos_item_set = None
        self.pos_item_set_eval = None
        self.sampled = None

    # -- parameter access (embed_attribute.py:58-70) -----------------------
    def _tables(self, prefix, att):
        cat = [self.p['%sembed_cat_%d' % (prefix, i)] for i in range(att.num_features_cat)]
        mul = [self.p['%sembed_mulhot_%d' % (prefix, i)] for i in range(att.num_features_mulhot)]
        return cat, mul

    def _biases(self, prefix, att):
        cat = [self.p['%s_bias_cat_%d' % (prefix, i)] for i in range(att.num_features_cat)]
        mul = [self.p['%s_bias_mulhot_%d' % (prefix, i)] for i in range(att.num_features_mulhot)]
        return cat, mul

    def _out_prefix(self):
        return 'item_output' if self.item_output else 'item'

    # -- embed_attribute.py:350-417 ----------------------------------------
    def _get_embedded(self, embs_cat, embs_mulhot, b_cat, b_mulhot, inds, att,
                      concatenation=True, no_id=False, no_attribute=False):
        inds = np.asarray(inds, dtype=np.int64)
        mb = len(inds)
        cat_list, mulhot_list, bias_cat_list, bias_mulhot_list = [], [], [], []
        if no_id and att.num_features_cat == 1:          # :356-366
            assert b_cat is None and b_mulhot is None, 'error: not implemented'
            dim = embs_cat[0].shape[1]
            z = np.zeros((mb, dim), dtype=self.dtype)
            return (z, None) if concatenation else ([z], [], None)
        n1 = 1 if no_attribute else att.num_features_cat      # :368
        n2 = 0 if no_attribute else att.num_features_mulhot   # :369
        for i in range(n1):
            if no_id and i == 0:
                continue
            cat_indices = np.asarray(att.features_cat[i], dtype=np.int64)[inds]   # :374
            cat_list.append(embs_cat[i][cat_indices])                               # :375
            if b_cat is not None:
                bias_cat_list.append(b_cat[i][cat_indices])                         # :379
        for i in range(n2):
            begin_ = np.asarray(att.mulhot_starts[i], dtype=np.int64)[inds]        # :383
            size_ = np.asarray(att.mulhot_lengths[i], dtype=np.int64)[inds]        # :384
            mulhot_indices = batch_slice2(att.features_mulhot[i], begin_, size_)    # :394
            mulhot_segids = batch_segids2(size_)                                    # :396
            embedded_flat = embs_mulhot[i][mulhot_indices]                          # :397
            embedded_sum = unsorted_segment_sum(embedded_flat, mulhot_segids, mb)   # :398
            lengs = size_.astype(self.dtype).reshape(mb, 1)                         # :399
            mulhot_list.append(embedded_sum / lengs)                                # :400
            if b_mulhot is not None:
                b_flat = b_mulhot[i][mulhot_indices]                                # :403
                b_sum = unsorted_segment_sum(b_flat, mulhot_segids, mb)             # :404
                bias_mulhot_list.append(b_sum / lengs)                              # :406
        if b_cat is None and b_mulhot is None:
            bias = None
        else:
            bias = np.mean(np.stack(bias_cat_list + bias_mulhot_list, 0), 0).reshape(-1)  # :412
        if concatenation:
            return np.concatenate(cat_list + mulhot_list, 1), bias                  # :415
        return cat_list, mulhot_list, bias

    def flat_indices(self, att, i, inds):
        """The integer part of :383-396 alone (bit-exact parity target)."""
        inds = np.asarray(inds, dtype=np.int64)
        begin_ = np.asarray(att.mulhot_starts[i], dtype=np.int64)[inds]
        size_ = np.asarray(att.mulhot_lengths[i], dtype=np.int64)[inds]
        return batch_slice2(att.features_mulhot[i], begin_, size_), batch_segids2(size_)

    # -- embed_attribute.py:222-237 ----------------------------------------
    def get_batch_user(self, u_inds, keep_prob=1.0, concat=True, no_id=False, dropout_mask=None):
        cat, mul = self._tables('user', self.user_attributes)
        if concat:
            emb, user_b = self._get_embedded(cat, mul, None, None, u_inds, self.user_attributes,
                                             concatenation=True, no_id=no_id)
        else:
            c, m, user_b = self._get_embedded(cat, mul, None, None, u_inds, self.user_attributes,
                                              concatenation=False, no_id=no_id)
            emb = np.mean(np.stack(c + m, 0), 0)                                    # :235
        emb = dropout(emb, keep_prob, dropout_mask)                                 # :236
        return emb, user_b

    # -- embed_attribute.py:239-254 ----------------------------------------
    def get_batch_item(self, i_inds, concat=False, no_attribute=False):
        cat, mul = self._tables('item', self.item_attributes)
        bc, bm = self._biases('item', self.item_attributes)
        if concat:
            return self._get_embedded(cat, mul, bc, bm, i_inds, self.item_attributes,
                                      concatenation=True, no_attribute=no_attribute)
        c, m, b = self._get_embedded(cat, mul, bc, bm, i_inds, self.item_attributes,
                                     concatenation=False, no_attribute=no_attribute)
        return c + m, b

    # -- embed_attribute.py:320-348 ----------------------------------------
    def pass_sampled_items(self, item_sampled):
        """Materialise the CSR of the sampled pool (the `update_sampled` ops)."""
        att = self.item_attributes
        inds = np.asarray(item_sampled, dtype=np.int64)
        cat_ind = [np.asarray(att.features_cat[i], dtype=np.int64)[inds]
                   for i in range(att.num_features_cat)]
        mul_ind, mul_seg, mul_len = [], [], []
        for i in range(att.num_features_mulhot):
            idx, seg = self.flat_indices(att, i, inds)
            mul_ind.append(idx)
            mul_seg.append(seg)
            size_ = np.asarray(att.mulhot_lengths[i], dtype=np.int64)[inds]
            mul_len.append(size_.astype(self.dtype).reshape(len(inds), 1))
        self.sampled = (cat_ind, mul_ind, mul_seg, mul_len)
        return self.sampled

    def full_indices(self):
        """embed_attribute.py:97-108 — catalog-ordered constants."""
        ia = self.item_attributes
        cat = [np.asarray(ia.full_cat_tr[i], dtype=np.int64) for i in range(ia.num_features_cat)]
        val = [np.asarray(ia.full_values_tr[i], dtype=np.int64) for i in range(ia.num_features_mulhot)]
        seg = [np.asarray(ia.full_segids_tr[i], dtype=np.int64) for i in range(ia.num_features_mulhot)]
        leng = [np.asarray(ia.full_lengths_tr[i], dtype=self.dtype).reshape(-1, 1)
                for i in range(ia.num_features_mulhot)]
        return cat, val, seg, leng

    # -- embed_attribute.py:148-206 ----------------------------------------
    def get_prediction(self, latent, pool='full', output_feat=1):
        """Literal order: score every table token, then gather+pool per catalog item."""
        ia = self.item_attributes
        if pool == 'full':
            indices_cat, indices_mulhot, segids_mulhot, lengths_mulhot = self.full_indices()
            V = self.logit_size
        else:
            indices_cat, indices_mulhot, segids_mulhot, lengths_mulhot = self.sampled
            V = self.n_sampled
        cat, mul = self._tables(self._out_prefix(), ia)
        bc, bm = self._biases(self._out_prefix(), ia)
        innerps = []
        n1 = 1 if output_feat == 0 else ia.num_features_cat          # :163
        n2 = 0 if output_feat == 0 else ia.num_features_mulhot       # :164
        for i in range(n1):
            u = latent[i] if isinstance(latent, list) else latent
            innerp = cat[i] @ u.T + bc[i]                            # :171  [V_f, mb]
            innerps.append(innerp[indices_cat[i]])                   # :172
        offset = ia.num_features_cat
        for i in range(n2):
            u = latent[i + offset] if isinstance(latent, list) else latent
            innerp = mul[i] @ u.T + bm[i]                            # :188
            inds, segids, lengs = indices_mulhot[i], segids_mulhot[i], lengths_mulhot[i]
            if output_feat == 1:
                innerps.append(unsorted_segment_sum(innerp[inds], segids, V) / lengs)     # :192
            elif output_feat == 2:
                innerps.append(segment_max(innerp[inds], segids, V))                      # :195
            elif output_feat == 3:
                score_max = innerp.max()                                                  # :197
                innerp = innerp - score_max
                innerps.append(score_max + np.log(1 + unsorted_segment_sum(
                    np.exp(innerp[inds]), segids, V)))                                    # :199
            else:
                raise SystemExit('Error: Attribute combination not implemented!')
        return np.mean(np.stack(innerps, 0), 0).T                    # :205  [mb, V]

    # -- embed_attribute.py:208-220 ----------------------------------------
    def get_target_score(self, latent, inds):
        ia = self.item_attributes
        cat, mul = self._tables(self._out_prefix(), ia)
        bc, bm = self._biases(self._out_prefix(), ia)
        c, m, i_bias = self._get_embedded(cat, mul, bc, bm, inds, ia, concatenation=False)
        target_item_emb = np.mean(np.stack(c + m, 0), 0)             # :219
        return np.sum(latent * target_item_emb, 1) + i_bias          # :220

    # -- embed_attribute.py:674-684 ----------------------------------------
    def prepare_warp(self, pos_item_set, pos_item_set_eval):
        self.pos_item_set = pos_item_set
        self.pos_item_set_eval = pos_item_set_eval

    def target_mapping(self, item_target):
        m = self.item_ind2logit_ind
        return [[m[v] for v in items] for items in item_target]

    # -- embed_attribute.py:721-745 + :651-672 -----------------------------
    def build_mask(self, user_input, loss, forward_only=False, item_sampled_id2idx=None):
        """Dense bool mask [mb, V]; False marks 'another positive of this user'."""
        V = self.n_sampled if loss == 'mw' else self.logit_size
        mb = len(user_input)
        mask = np.ones((mb, V), dtype=bool)
        item_set = self.pos_item_set_eval if forward_only else self.pos_item_set
        s_2idx = self.item_ind2logit_ind if loss != 'mw' else item_sampled_id2idx
        for c, u in enumerate(user_input):
            if u in item_set:
                for v in item_set[u]:
                    if loss == 'mw':
                        if v in s_2idx:                              # :739-740
                            mask[c, s_2idx[v]] = False
                    else:
                        mask[c, s_2idx[v]] = False                   # :733
        return mask

    # -- embed_attribute.py:525-649 ----------------------------------------
    def compute_loss(self, logits, item_target, loss='ce', mask=None, loss_func='log',
                     exp_p=1.005, true_rank=False):
        return compute_loss(logits, item_target, loss, mask, loss_func, exp_p, true_rank)


def dropout(x, keep_prob, mask=None):
    """tf.nn.dropout: x/keep * floor(keep + U) — the 0/1 mask is injected (SURVEY
    headline fact: the reference's masks are unseeded, parity is by injection)."""
    if keep_prob == 1.0 or mask is None:
        return x
    return x / keep_prob * mask


def sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def compute_loss(logits, item_target, loss='ce', mask=None, loss_func='log', exp_p=1.005,
                 true_rank=False):
    """embed_attribute.py:525-649.  Returns per-row loss [mb] (or the reference's
    list for warp_eval / true_rank)."""
    mb = logits.shape[0]
    rows = np.arange(mb)
    if loss == 'ce':                                                 # :530
        mx = logits.max(1, keepdims=True)
        lse = mx[:, 0] + np.log(np.exp(logits - mx).sum(1))
        return lse - logits[rows, item_target]
    if loss in ('rs', 'rs-sig', 'rs-sig2', 'bbpr'):                  # :551-603
        target_logits = logits[rows, item_target].reshape(mb, 1)
        if loss in ('rs', 'rs-sig'):
            errors = np.maximum(logits - target_logits + 1, 0)       # :566-567
        else:
            errors = sigmoid(logits - target_logits)                 # :569
        errors_masked = np.where(mask, errors, 0.0)                  # :573
        if loss == 'rs-sig':
            errors_masked = sigmoid(errors_masked) * 2 - 1           # :577-578
        s = errors_masked.sum(1)
        if true_rank:                                                # :596-601
            nomargin = np.where(mask, np.maximum(logits - target_logits, 0), 0.0)
            return [errors_masked, np.count_nonzero(nomargin, 1)]
        if loss == 'bbpr':
            return s                                                 # :594
        return _rs_transform(s, loss_func, exp_p)[0]
    if loss == 'warp':                                               # :605-618
        target_logits = logits[rows, item_target].reshape(mb, 1)
        t = np.where(mask, logits - target_logits + 1, 0.0)
        return np.log(1 + np.maximum(t, 0).sum(1))
    if loss == 'warp_eval':                                          # :620-639
        target_logits = logits[rows, item_target].reshape(mb, 1)
        t = np.where(mask, logits - target_logits + 1, 0.0)
        margin_rank = np.maximum(t, 0).sum(1)
        t2 = np.where(mask, np.maximum(logits - target_logits, 0), 0.0)
        return [margin_rank, np.count_nonzero(t2, 1)]
    if loss == 'mw':                                                 # :641-649
        t = np.where(mask, logits - np.asarray(item_target).reshape(mb, 1) + 1, 0.0)
        return np.log(1 + np.maximum(t, 0).sum(1))
    if loss == 'bpr':                                                # :542
        return np.log(1 + np.exp(logits))
    if loss == 'bpr-hinge':                                          # :544
        return np.maximum(1 + logits, 0)
    raise SystemExit('Error: not implemented other loss!!')


def _rs_transform(s, loss_func, exp_p):
    """embed_attribute.py:580-592: returns (l, dl/ds)."""
    if loss_func == 'log':
        return np.log(1 + s), 1.0 / (1 + s)
    if loss_func == 'exp':
        return 1 - np.power(exp_p, -s), np.log(exp_p) * np.power(exp_p, -s)
    if loss_func == 'poly':
        with np.errstate(divide='ignore', invalid='ignore'):
            g = exp_p * np.power(s, exp_p - 1)
        return np.power(s, exp_p), g
    if loss_func == 'poly2':
        return np.power(1 + s, exp_p), exp_p * np.power(1 + s, exp_p - 1)
    if loss_func == 'linear':
        return s, np.ones_like(s)
    if loss_func == 'square':
        return np.square(s), 2 * s
    raise SystemExit('unknown loss_func')


def loss_and_grad(logits, item_target, loss='ce', mask=None, loss_func='log', exp_p=1.005):
    """Per-row loss and d(sum_b loss_b)/d logits (and d/d target-score for 'mw').

    Hand-derived; cross-checked against torch float64 autograd in tests.  ReLU'(0)=0
    as in TF's ReluGrad.
    Returns (loss[mb], dlogits[mb,V], dtarget[mb] or None).
    """
    mb, V = logits.shape
    rows = np.arange(mb)
    if loss == 'ce':
        mx = logits.max(1, keepdims=True)
        e = np.exp(logits - mx)
        ssum = e.sum(1, keepdims=True)
        l = mx[:, 0] + np.log(ssum[:, 0]) - logits[rows, item_target]
        d = e / ssum
        d[rows, item_target] -= 1.0
        return l, d, None
    if loss == 'mw':
        t = np.asarray(item_target).reshape(mb, 1)
        z = logits - t + 1
        act = (mask & (z > 0)).astype(logits.dtype)
        s = (np.maximum(z, 0) * mask).sum(1)
        l = np.log(1 + s)
        g = 1.0 / (1 + s)
        d = act * g[:, None]
        return l, d, -d.sum(1)
    tl = logits[rows, item_target].reshape(mb, 1)
    if loss in ('warp', 'rs', 'rs-sig'):
        z = logits - tl + 1
        r = np.where(mask, np.maximum(z, 0), 0.0)
        dr = (mask & (z > 0)).astype(logits.dtype)
        if loss == 'rs-sig':
            sg = sigmoid(r)
            e = sg * 2 - 1
            de = 2 * sg * (1 - sg) * dr
        else:
            e, de = r, dr
    elif loss in ('rs-sig2', 'bbpr'):
        sg = sigmoid(logits - tl)
        e = np.where(mask, sg, 0.0)
        de = np.where(mask, sg * (1 - sg), 0.0)
    else:
        raise SystemExit('loss_and_grad: unsupported loss %s' % loss)
    s = e.sum(1)
    if loss == 'warp':
        l, g = _rs_transform(s, 'log', exp_p)
    elif loss == 'bbpr':
        l, g = s, np.ones_like(s)
    else:
        l, g = _rs_transform(s, loss_func, exp_p)
    d = de * np.asarray(g).reshape(mb, 1)
    np.subtract.at(d, (rows, item_target), d.sum(1))   # d/d target logit = -sum_v (uses pre-update row sums)
    return l, d, None


# --------------------------------------------------------------------------
# optimiser (tf.train.AdagradOptimizer, acc0 = 0.1, no eps)
# --------------------------------------------------------------------------
def adagrad_update(theta, acc, g, lr):
    """Dense rule; rows with g == 0 are unchanged, so it equals the sparse rule with
    duplicate indices summed first (hmf_model.py:146-151)."""
    acc = acc + g * g
    theta = theta - lr * g / np.sqrt(acc)
    return theta, acc


def clip_by_global_norm(grads, clip):
    """tf.clip_by_global_norm (lstm/seqModel.py:180)."""
    norm = np.sqrt(sum(float((g * g).sum()) for g in grads))
    scale = clip / max(norm, clip)
    return [g * scale for g in grads], norm


# --------------------------------------------------------------------------
# pooled ("rewritten") forms used by the CUDA path; the tests prove them equal
# to the literal forms above (pool-then-GEMM == GEMM-then-pool).
# --------------------------------------------------------------------------
def pool_entities(emb, tables_cat, tables_mulhot, b_cat, b_mulhot, att, inds):
    """mean over attributes of the per-attribute pooled vectors (+ pooled bias)."""
    c, m, bias = emb._get_embedded(tables_cat, tables_mulhot, b_cat, b_mulhot, inds, att,
                                   concatenation=False)
    return np.mean(np.stack(c + m, 0), 0), bias


def pool_backward(att, inds, dpooled, dbias, shapes_cat, shapes_mulhot, dtype=np.float64):
    """Adjoint of pool_entities: dense gradients for every table (and [V,1] bias)."""
    inds = np.asarray(inds, dtype=np.int64)
    F = att.num_features_cat + att.num_features_mulhot
    g_cat = [np.zeros(s, dtype=dtype) for s in shapes_cat]
    g_mul = [np.zeros(s, dtype=dtype) for s in shapes_mulhot]
    gb_cat = [np.zeros((s[0], 1), dtype=dtype) for s in shapes_cat]
    gb_mul = [np.zeros((s[0], 1), dtype=dtype) for s in shapes_mulhot]
    for i in range(att.num_features_cat):
        tok = np.asarray(att.features_cat[i], dtype=np.int64)[inds]
        np.add.at(g_cat[i], tok, dpooled / F)
        if dbias is not None:
            np.add.at(gb_cat[i][:, 0], tok, dbias / F)
    for i in range(att.num_features_mulhot):
        begin_ = np.asarray(att.mulhot_starts[i], dtype=np.int64)[inds]
        size_ = np.asarray(att.mulhot_lengths[i], dtype=np.int64)[inds]
        idx = batch_slice2(att.features_mulhot[i], begin_, size_)
        seg = batch_segids2(size_)
        w = 1.0 / (F * size_.astype(dtype))
        np.add.at(g_mul[i], idx, dpooled[seg] * w[seg][:, None])
        if dbias is not None:
            np.add.at(gb_mul[i][:, 0], idx, dbias[seg] * w[seg])
    return g_cat, g_mul, gb_cat, gb_mul


# --------------------------------------------------------------------------
# hmf/hmf_model.py :: LatentProductModel  (forward + hand-derived backward)
# --------------------------------------------------------------------------
class OracleHMF(object):
    """hmf_model.py:20-156 with step() semantics of :162-228, eager NumPy.

    Dense float parameters live in ``self.emb.p``; Adagrad accumulators in
    ``self.acc`` (0.1-initialised).  MLP weights 'w1','b1','w2','b2' if nonlinear.
    """

    def __init__(self, emb, loss='ce', nonlinear='linear', keep_prob=1.0, learning_rate=0.1,
                 loss_func='log', loss_exp_p=1.005):
        self.emb = emb
        self.loss_function = loss
        self.nonlinear = nonlinear
        self.keep_prob = keep_prob
        self.lr = learning_rate
        self.loss_func = loss_func
        self.loss_exp_p = loss_exp_p
        self.acc = {k: np.full_like(v, 0.1) for k, v in emb.p.items()}
        self.global_step = 0

    # forward pieces ---------------------------------------------------------
    def _act(self, x):
        return np.maximum(x, 0) if self.nonlinear == 'relu' else np.tanh(x)

    def _dact(self, y, x):
        return (x > 0).astype(x.dtype) if self.nonlinear == 'relu' else 1 - y * y

    def user_tower(self, user_input, keep_prob, masks):
        """hmf_model.py:78-94.  masks: list of injected 0/1 dropout masks (1 for linear,
        3 for MLP), or None when keep_prob == 1."""
        e = self.emb
        cache = {}
        if self.nonlinear in ('relu', 'tanh'):
            u0, _ = e.get_batch_user(user_input, 1.0, False)                 # :87
            a0 = self._act(u0)
            h0 = dropout(a0, keep_prob, masks[0] if masks else None)         # :88
            z1 = h0 @ e.p['w1'] + e.p['b1']
            a1 = self._act(z1)                                               # :90
            h1 = dropout(a1, keep_prob, masks[1] if masks else None)         # :91
            z2 = h1 @ e.p['w2'] + e.p['b2']
            a2 = self._act(z2)                                               # :93
            u = dropout(a2, keep_prob, masks[2] if masks else None)          # :94
            cache = dict(u0=u0, a0=a0, h0=h0, z1=z1, a1=a1, h1=h1, z2=z2, a2=a2)
        else:
            u0, _ = e.get_batch_user(user_input, 1.0, False)
            u = dropout(u0, keep_prob, masks[0] if masks else None)          # :78 / embed :236
            cache = dict(u0=u0)
        return u, cache

    def forward(self, user_input, item_input, forward_only=False, masks=None,
                item_sampled_id2idx=None, literal=True):
        """Returns mean loss (hmf_model.py:140,144) and a cache for backward()."""
        e = self.emb
        loss = self.loss_function
        keep = 1.0 if forward_only else self.keep_prob                       # :167-170
        u, cache = self.user_tower(user_input, keep, masks)
        targets = np.asarray(e.target_mapping([item_input])[0], dtype=np.int64)  # :173
        eff = loss
        if loss == 'mw' and forward_only:
            eff = 'warp'                                                     # :130,:144
        mask = None
        if loss == 'mw' and forward_only:
            # hmf_model.py:209-211 runs set_mask['mw'] only: the 'warp' mask Variable behind loss_eval
            # (:130) is never written, so the reference's mw eval loss is WARP with NO positives masked
            # (pinned by tests/golden/ref_hmf_mw_*.npz, produced by the reference's own code)
            mask = np.ones((len(user_input), e.logit_size), dtype=bool)
        elif eff in ('warp', 'warp_eval', 'rs', 'rs-sig', 'rs-sig2', 'bbpr', 'mw'):
            mask = e.build_mask(user_input, eff, forward_only, item_sampled_id2idx)
        if eff == 'mw':
            logits = e.get_prediction(u, 'sampled')                          # :112
            tscore = e.get_target_score(u, item_input)                       # :115
            l, dlog, dts = loss_and_grad(logits, tscore, 'mw', mask)
        else:
            logits = e.get_prediction(u)                                     # :118
            l, dlog, dts = loss_and_grad(logits, targets, eff, mask, self.loss_func,
                                         self.loss_exp_p)
        cache.update(u=u, keep=keep, masks=masks, logits=logits, dlog=dlog, dts=dts,
                     user_input=user_input, item_input=item_input, eff=eff, batch_loss=l)
        return float(np.mean(l)), cache                                      # :140

    # backward (hand-derived, pool-first form) -------------------------------
    def backward(self, cache):
        e = self.emb
        ia, ua = e.item_attributes, e.user_attributes
        dt = e.dtype
        mb = len(cache['user_input'])
        dlog = cache['dlog'] / mb                     # d mean / d logits
        u = cache['u']
        grads = {k: np.zeros_like(v) for k, v in e.p.items()}
        pre = e._out_prefix()
        cat, mul = e._tables(pre, ia)
        bc, bm = e._biases(pre, ia)
        shapes_cat = [t.shape for t in cat]
        shapes_mul = [t.shape for t in mul]
        if cache['eff'] == 'mw':
            n_s = e.n_sampled
            # reconstruct the sampled pool ids from the stored CSR is not possible in
            # general; the caller stores them:
            ids = np.asarray(self.sampled_ids, dtype=np.int64)
        else:
            ids = np.asarray([e.logit_ind2item_ind[v] for v in range(e.logit_size)], dtype=np.int64)
        P, bp = pool_entities(e, cat, mul, bc, bm, ia, ids)
        du = dlog @ P                                  # [mb,d]
        dP = dlog.T @ u                                # [N,d]
        dbp = dlog.sum(0)
        gc, gm, gbc, gbm = pool_backward(ia, ids, dP, dbp, shapes_cat, shapes_mul, dt)
        if cache['dts'] is not None:
            dts = cache['dts'] / mb
            tid = np.asarray(cache['item_input'], dtype=np.int64)
            Pt, _ = pool_entities(e, cat, mul, bc, bm, ia, tid)
            du = du + dts[:, None] * Pt
            gc2, gm2, gbc2, gbm2 = pool_backward(ia, tid, dts[:, None] * u, dts,
                                                 shapes_cat, shapes_mul, dt)
            gc = [a + b for a, b in zip(gc, gc2)]
            gm = [a + b for a, b in zip(gm, gm2)]
            gbc = [a + b for a, b in zip(gbc, gbc2)]
            gbm = [a + b for a, b in zip(gbm, gbm2)]
        for i in range(ia.num_features_cat):
            grads['%sembed_cat_%d' % (pre, i)] += gc[i]
            grads['%s_bias_cat_%d' % (pre, i)] += gbc[i]
        for i in range(ia.num_features_mulhot):
            grads['%sembed_mulhot_%d' % (pre, i)] += gm[i]
            grads['%s_bias_mulhot_%d' % (pre, i)] += gbm[i]
        # user tower
        keep, masks = cache['keep'], cache['masks']

        def ddrop(g, k):
            if keep == 1.0 or masks is None:
                return g
            return g / keep * masks[k]
        if self.nonlinear in ('relu', 'tanh'):
            g = ddrop(du, 2)
            g = g * self._dact(cache['a2'], cache['z2'])
            grads['w2'] += cache['h1'].T @ g
            grads['b2'] += g.sum(0)
            g = g @ e.p['w2'].T
            g = ddrop(g, 1)
            g = g * self._dact(cache['a1'], cache['z1'])
            grads['w1'] += cache['h0'].T @ g
            grads['b1'] += g.sum(0)
            g = g @ e.p['w1'].T
            g = ddrop(g, 0)
            du0 = g * self._dact(cache['a0'], cache['u0'])
        else:
            du0 = ddrop(du, 0)
        ucat, umul = e._tables('user', ua)
        gc, gm, _, _ = pool_backward(ua, cache['user_input'], du0, None,
                                     [t.shape for t in ucat], [t.shape for t in umul], dt)
        for i in range(ua.num_features_cat):
            grads['userembed_cat_%d' % i] += gc[i]
        for i in range(ua.num_features_mulhot):
            grads['userembed_mulhot_%d' % i] += gm[i]
        return grads

    def apply_gradients(self, grads):
        """hmf_model.py:146-151."""
        for k, g in grads.items():
            self.emb.p[k], self.acc[k] = adagrad_update(self.emb.p[k], self.acc[k], g, self.lr)
        self.global_step += 1

    def step(self, user_input, item_input, item_sampled=None, item_sampled_id2idx=None,
             forward_only=False, masks=None):
        """hmf_model.py:162-228 (training / eval legs)."""
        if item_sampled is not None and self.loss_function == 'mw':
            self.emb.pass_sampled_items(item_sampled)                        # :206-207
            self.sampled_ids = list(item_sampled)
            self.sampled_id2idx = item_sampled_id2idx
        id2idx = item_sampled_id2idx if item_sampled_id2idx is not None else getattr(
            self, 'sampled_id2idx', None)
        loss, cache = self.forward(user_input, item_input, forward_only, masks, id2idx)
        if not forward_only:
            self.apply_gradients(self.backward(cache))
        return loss

    def top_k(self, user_input, k):
        """hmf_model.py:154 — tf.nn.top_k(sorted=True): descending, ties -> lower index."""
        u, _ = self.user_tower(user_input, 1.0, None)
        logits = self.emb.get_prediction(u)
        order = np.argsort(-logits, axis=1, kind='stable')[:, :k]
        return order, np.take_along_axis(logits, order, 1)


# --------------------------------------------------------------------------
# TF1.0 LSTMCell (lstm/seqModel.py:99-103)
# --------------------------------------------------------------------------
def lstm_cell(x, h, c, W, b, forget_bias=1.0):
    """i,j,f,o = split([x,h]W + b, 4); c' = sig(f+1)c + sig(i)tanh(j); h' = sig(o)tanh(c')."""
    z = np.concatenate([x, h], 1) @ W + b
    i, j, f, o = np.split(z, 4, axis=1)
    c2 = sigmoid(f + forget_bias) * c + sigmoid(i) * np.tanh(j)
    h2 = sigmoid(o) * np.tanh(c2)
    return h2, c2


def lstm_seq(X, W, b, in_masks=None, out_masks=None, keep=1.0):
    """static_rnn unroll (seqModel.py:477) with DropoutWrapper in/out (:100-103).
    X: [T, mb, d_in]; returns outputs [T, mb, H] (after output dropout), final (h, c)."""
    T, mb, _ = X.shape
    H = W.shape[1] // 4
    h = np.zeros((mb, H), dtype=X.dtype)
    c = np.zeros((mb, H), dtype=X.dtype)
    outs = []
    for t in range(T):
        x = dropout(X[t], keep, in_masks[t] if in_masks is not None else None)
        h, c = lstm_cell(x, h, c, W, b)
        outs.append(dropout(h, keep, out_masks[t] if out_masks is not None else None))
    return np.stack(outs, 0), (h, c)


def sequence_loss(per_step_losses, weights):
    """seqModel.py:524-604: sum_b [ sum_t w*l / (sum_t w + 1e-12) ] (not batch-averaged)."""
    num = sum(l * w for l, w in zip(per_step_losses, weights))
    den = sum(weights) + 1e-12
    return float((num / den).sum())
