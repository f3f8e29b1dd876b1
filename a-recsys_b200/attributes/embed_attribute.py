"""EmbeddingAttribute on B200: same public surface as the reference class
(attributes/embed_attribute.py:19-747) — get_batch_user / get_batch_item / get_prediction /
get_target_score / get_sampled_item / compute_loss / get_warp_mask / prepare_warp /
target_mapping / add_input / get_user_model_size / get_item_model_size — executed eagerly by
the sm_100a kernels of libarx_b200.so instead of a TF-1 graph.

Differences that are deliberate:
  * placeholders become device index tensors filled by add_input();
  * catalog scoring pools the catalog first and then contracts (pooling is linear), the
    literal "score tokens, then pool" order lives only in the oracle;
  * the dense bool mask Variable [mb*V] (:651-672) becomes a per-user CSR of positives;
  * backward is explicit: every lookup registers its output gradient with
    `push_grad`, and `apply_gradients` de-duplicates all touched rows of a table set and
    applies Adagrad once per row (TF sums duplicate indices before the sparse apply).
There is no CPU path: constructing this class without CUDA raises.
"""
import ctypes
import math

import os

import numpy as np
import torch

from .. import _lib
from .._lib import AttrDesc, BwdPlan, POOL_MEAN, POOL_CONCAT, OPT_ADAGRAD, OPT_SGD, OPT_NONE, call, ptr

ADAGRAD_INIT_ACC = 0.1     # tf.train.AdagradOptimizer(initial_accumulator_value=0.1)


class _TableSet(object):
    """All tables of one variable prefix ('user' | 'item' | 'item_output') plus the device
    descriptor array the kernels read."""

    def __init__(self, owner, prefix, att, att_dev, with_bias):
        self.prefix = prefix
        self.att = att
        self.n_cat = att.num_features_cat
        self.n_mul = att.num_features_mulhot
        self.n_attr = self.n_cat + self.n_mul
        self.with_bias = with_bias
        self.names, self.bias_names = [], []
        descs = (AttrDesc * self.n_attr)()
        feats_cat, feats_mul, starts, lengths = att_dev[:4]
        lengths_full = att_dev[4] if len(att_dev) > 4 else None
        k = 0
        for kind, cnt in ((0, self.n_cat), (1, self.n_mul)):
            for i in range(cnt):
                tag = 'cat' if kind == 0 else 'mulhot'
                name = '%sembed_%s_%d' % (prefix, tag, i)
                bname = '%s_bias_%s_%d' % (prefix, tag, i)
                d = descs[k]
                d.table = owner.params[name].data_ptr()
                d.table_acc = owner.accs[name].data_ptr()
                if with_bias:
                    d.bias = owner.params[bname].data_ptr()
                    d.bias_acc = owner.accs[bname].data_ptr()
                d.values = (feats_cat[i] if kind == 0 else feats_mul[i]).data_ptr()
                if kind == 1:
                    d.starts = starts[i].data_ptr()
                    d.lengths = lengths[i].data_ptr()
                    if lengths_full is not None:
                        d.lengths_full = lengths_full[i].data_ptr()
                d.touch = owner.touch[name].data_ptr()
                d.vocab = owner.touch[name].shape[0]            # global vocabulary (touch is global-sized)
                assert d.vocab < (1 << 26), 'backward plan packs (attribute, row) in 32 bits: vocab < 2^26'
                d.kind = kind
                if owner.shard is not None:
                    d.reserved = (owner.shard[0] << 16) | owner.shard[1]
                self.names.append(name)
                self.bias_names.append(bname if with_bias else None)
                k += 1
        raw = np.frombuffer(bytes(descs), dtype=np.uint8).copy()
        self.descs_dev = torch.from_numpy(raw).to(owner.device)
        self.desc_size = ctypes.sizeof(AttrDesc)
        self.total_vocab = sum(owner.touch[n].shape[0] for n in self.names)
        self.max_len = [1] * self.n_cat + [int(att.mulhot_lengths[i].max()) for i in range(self.n_mul)]
        self.pending = []          # (attr_begin, n_attr, ids, mode, dout[n, w], dbias or None, plan_key)
        self.plans = {}            # plan_key -> (_Plan, rows)

    def desc_ptr(self, attr_begin=0):
        return self.descs_dev.data_ptr() + attr_begin * self.desc_size

    def attr_range(self, no_id=False, no_attribute=False):
        """embed_attribute.py:356-373: no_id drops categorical attribute 0, no_attribute keeps it only."""
        if no_attribute:
            return 0, 1
        if no_id:
            return 1, self.n_attr - 1
        return 0, self.n_attr


class _Plan(object):
    """Device buffers of one arx_bwd_plan."""

    def __init__(self, device, cap_rows, cap_occ, dim):
        i32 = dict(dtype=torch.int32, device=device)
        _lib.load()
        cap_chunks = 2 * cap_occ // _lib.TUNING.get('heavy', _lib.HEAVY_DEFAULT) + 1     # rows above `heavy` entries: < 2 c / heavy chunks each
        self.chunk_row = torch.empty(cap_chunks, **i32)
        self.row_chunk0 = torch.empty(cap_rows, **i32)
        self.row_done = torch.zeros(cap_rows, **i32)
        self.partials = torch.empty(cap_chunks * (dim + 1), dtype=torch.float32, device=device)
        self.counters = torch.zeros(8, **i32)
        self.uniq_tok = torch.empty(cap_rows, **i32)
        self.uniq_attr = torch.empty(cap_rows, **i32)
        self.row_base = torch.empty(cap_rows, **i32)
        self.row_cnt = torch.empty(cap_rows, **i32)
        self.bucket_src = torch.empty(cap_occ, **i32)
        self.bucket_w = torch.empty(cap_occ, dtype=torch.float32, device=device)
        self.cap_rows, self.cap_occ = cap_rows, cap_occ
        self.c = BwdPlan(self.counters.data_ptr(), self.uniq_tok.data_ptr(), self.uniq_attr.data_ptr(),
                         self.row_base.data_ptr(), self.row_cnt.data_ptr(), self.bucket_src.data_ptr(),
                         self.bucket_w.data_ptr(), self.chunk_row.data_ptr(), self.row_chunk0.data_ptr(),
                         self.row_done.data_ptr(), self.partials.data_ptr(), cap_rows, cap_occ, cap_chunks)


class EmbeddingAttribute(object):
    def __init__(self, user_attributes, item_attributes, mb, n_sampled, input_steps=0,
                 item_output=False, item_ind2logit_ind=None, logit_ind2item_ind=None,
                 indices_item=None, devices=['/gpu:0'], device=None, seed=None, params=None, shard=None):
        if not torch.cuda.is_available():
            raise RuntimeError('arecsys_b200 needs a CUDA device (sm_100a); there is no CPU fallback')
        _lib.load()
        # shard = (G, r): this GPU stores rows t with t % G == r of every table
        self.shard = shard if (shard is not None and shard[0] > 1) else None
        # Plan-kernel flavour, per model instance (set on the library right before each plan build): block-aggregated
        # count / fill by default; ARX_TUNE=plan_agg=0 selects the warp-per-bag kernels for A/B runs.  Both key their
        # (attribute, row) pairs in 32 bits: _TableSet asserts vocab < 2^26.
        tune = dict(kv.split('=') for kv in os.environ.get('ARX_TUNE', '').split(',') if '=' in kv)
        self.plan_agg = int(tune.get('plan_agg', 3))
        self.device = torch.device(device if device is not None else 'cuda:%d' % torch.cuda.current_device())
        self.user_attributes = user_attributes
        self.item_attributes = item_attributes
        self.batch_size = mb
        self.n_sampled = n_sampled
        self.input_steps = input_steps
        self.item_output = item_output
        self.num_item_features = item_attributes.num_features_cat + item_attributes.num_features_mulhot
        self.item_ind2logit_ind = item_ind2logit_ind
        self.logit_ind2item_ind = logit_ind2item_ind
        if logit_ind2item_ind is not None:
            self.logit_size = len(logit_ind2item_ind)
        self.indices_item = indices_item if indices_item is not None else range(getattr(self, 'logit_size', 0))
        self.devices = devices
        self.mask, self.pos_indices = {}, {}
        ua, ia = user_attributes, item_attributes
        sizes = set(ua._embedding_size_list_cat + ua._embedding_size_list_mulhot +
                    ia._embedding_size_list_cat + ia._embedding_size_list_mulhot)
        assert len(sizes) == 1, 'all attribute embeddings share one size (set_model_size(int))'
        self.dim = sizes.pop()

        # ---- attribute arrays -> HBM (embed_attribute.py:308-318) ------------------------
        self.att = {'user': self._init_attributes(ua), 'item': self._init_attributes(ia)}
        if item_output:
            self.att['item_output'] = self.att['item']

        # ---- variables (embed_attribute.py:265-306) --------------------------------------
        gen = torch.Generator(device='cpu')
        gen.manual_seed(0 if seed is None else seed)
        self.params, self.accs, self.touch = {}, {}, {}
        self._embedded(ua, 'user', gen, params)
        self._embedded(ia, 'item', gen, params)
        self._embedded_bias(ia, 'item', gen, params)
        if item_output:
            self._embedded(ia, 'item_output', gen, params)
            self._embedded_bias(ia, 'item_output', gen, params)
        self.sets = {'user': _TableSet(self, 'user', ua, self.att['user'], False),
                     'item': _TableSet(self, 'item', ia, self.att['item'], True)}
        if item_output:
            self.sets['item_output'] = _TableSet(self, 'item_output', ia, self.att['item'], True)

        # ---- feeds (placeholders in the reference, :72-93) -------------------------------
        self.u_indices, self.i_indices = {}, {}
        # catalog ids in logit order (the 'full' constants of :97-108 in CSR form)
        if logit_ind2item_ind is not None:
            V = self.logit_size
            ids = np.asarray([logit_ind2item_ind[v] for v in range(V)], dtype=np.int32) \
                if not isinstance(logit_ind2item_ind, np.ndarray) else logit_ind2item_ind.astype(np.int32)
            self.catalog_ids = torch.from_numpy(ids).to(self.device)
            n_items = ia.num_entities
            if hasattr(item_ind2logit_ind, 'as_array'):
                i2l = item_ind2logit_ind.as_array(n_items - 1)
            else:
                i2l = np.full(n_items, -1, dtype=np.int32)
                i2l[ids] = np.arange(V, dtype=np.int32)
                if isinstance(item_ind2logit_ind, dict):
                    # extra entries the runners add, e.g. START_ID -> logit 0 (lstm/run.py:276-278)
                    for k, v in item_ind2logit_ind.items():
                        if 0 <= k < n_items and i2l[k] < 0:
                            i2l[k] = v
            self.item2logit_dev = torch.from_numpy(i2l).to(self.device)
        self.sampled_ids = None
        self.sampled_pos_dev = None
        self.dense_table_grads = {}     # prefix -> {variable name: dense gradient}  (output_feat 2 / 3 scoring)
        self.deterministic = os.environ.get('ARX_DETERMINISTIC', '0') == '1'     # canonical bucket order (see _canonical_buckets)
        self.pos_csr = {}          # (kind) -> (ptr, idx) per-user positives, device
        self.pos_item_set = None
        self.pos_item_set_eval = None

    # ------------------------------------------------------------------ construction --
    def _init_attributes(self, att):
        dev = self.device
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.int32)).to(dev)
        if self.shard is None:
            return ([t(a) for a in att.features_cat], [t(a) for a in att.features_mulhot],
                    [t(a) for a in att.mulhot_starts], [t(a) for a in att.mulhot_lengths])
        # row-sharded tables: list, per bag, only the rows this rank owns (as local row numbers), so
        # that a rank walks 1/G of every bag instead of skipping foreign rows ("CSR
        # pre-partitioned by owner").  The full bag length stays the mean divisor.
        G, r = self.shard
        vals, sts, lens = [], [], []
        for f in range(att.num_features_mulhot):
            v, s, l = self._partition_bags(np.asarray(att.features_mulhot[f]), np.asarray(att.mulhot_starts[f]),
                                           np.asarray(att.mulhot_lengths[f]), G, r)
            vals.append(t(v)); sts.append(t(s)); lens.append(t(l))
        return ([t(a) for a in att.features_cat], vals, sts, lens, [t(a) for a in att.mulhot_lengths])

    @staticmethod
    def _partition_bags(values, starts, lengths, G, r):
        n1 = len(lengths)
        L = lengths.astype(np.int64)
        first = np.cumsum(L) - L
        seg = np.repeat(np.arange(n1, dtype=np.int64), L)
        tok = values[starts[:n1].astype(np.int64)[seg] + (np.arange(int(L.sum()), dtype=np.int64) - first[seg])]
        own = (tok % G) == r
        l_loc = np.bincount(seg[own], minlength=n1).astype(np.int32)
        s_loc = np.full(len(starts), int(l_loc.sum()), dtype=np.int64)
        s_loc[:n1] = np.cumsum(l_loc, dtype=np.int64) - l_loc
        v_loc = np.concatenate([tok[own] // G, [0]]).astype(np.int32)      # trailing pad as in the reference arrays
        return v_loc, s_loc.astype(np.int32), l_loc

    def _new_var(self, name, shape, gen, params):
        if params is not None and name in params:
            w = torch.as_tensor(np.asarray(params[name]), dtype=torch.float32).reshape(shape).clone()
        else:
            # tf.get_variable default in TF1.0: glorot_uniform; unseeded there.
            limit = math.sqrt(6.0 / (shape[0] + shape[1]))
            if self.shard is not None and shape[0] > (1 << 20):
                # large sharded table: draw only the local rows (a 10^7-row global draw costs 5 GB)
                G, r = self.shard
                local = (shape[0] - r + G - 1) // G
                w = (torch.rand((local, shape[1]), generator=gen, dtype=torch.float32) * 2 - 1) * limit
                self.params[name] = w.to(self.device).contiguous()
                self.accs[name] = torch.full(tuple(w.shape), ADAGRAD_INIT_ACC, dtype=torch.float32, device=self.device)
                return
            w = (torch.rand(shape, generator=gen, dtype=torch.float32) * 2 - 1) * limit
        if self.shard is not None:
            G, r = self.shard
            w = w[r::G]                       # the global initialisation, rows this GPU owns
        self.params[name] = w.to(self.device).contiguous()
        self.accs[name] = torch.full(tuple(w.shape), ADAGRAD_INIT_ACC, dtype=torch.float32, device=self.device)

    def _embedded(self, att, prefix, gen, params):
        for tag, n, V in (('cat', att.num_features_cat, att._embedding_classes_list_cat),
                          ('mulhot', att.num_features_mulhot, att._embedding_classes_list_mulhot)):
            for i in range(n):
                name = '%sembed_%s_%d' % (prefix, tag, i)
                self._new_var(name, (V[i], self.dim), gen, params)
                self.touch[name] = torch.zeros(V[i], dtype=torch.int32, device=self.device)

    def _embedded_bias(self, att, prefix, gen, params):
        for tag, n, V in (('cat', att.num_features_cat, att._embedding_classes_list_cat),
                          ('mulhot', att.num_features_mulhot, att._embedding_classes_list_mulhot)):
            for i in range(n):
                self._new_var('%s_bias_%s_%d' % (prefix, tag, i), (V[i], 1), gen, params)

    # ------------------------------------------------------------------ lookups -------
    def _ids(self, x):
        if isinstance(x, torch.Tensor):
            return x.to(device=self.device, dtype=torch.int32).contiguous()
        return torch.as_tensor(np.asarray(x, dtype=np.int32)).to(self.device, non_blocking=True)

    def pool(self, prefix, ids, mode=POOL_MEAN, want_bias=False, no_id=False, no_attribute=False,
             out=None, bias_out=None):
        """K1+K2 over `ids` for the table set `prefix`.  Returns (out, bias, (attr_begin, n_attr))."""
        ts = self.sets[prefix]
        a0, na = ts.attr_range(no_id, no_attribute)
        n = ids.numel()
        width = self.dim if mode == POOL_MEAN else self.dim * na
        if out is None:
            out = torch.empty((n, width), dtype=torch.float32, device=self.device)
        if want_bias and bias_out is None:
            bias_out = torch.empty((n,), dtype=torch.float32, device=self.device)
        _lib.tag = prefix
        call('arx_pool_fwd', ts.desc_ptr(a0), na, self.dim, ids.data_ptr(), n, out.data_ptr(),
             out.stride(0), mode, ptr(bias_out) if want_bias else None, int(sum(ts.max_len[a0:a0 + na])))
        return out, (bias_out if want_bias else None), (a0, na)

    def flat_indices(self, prefix, attr, ids):
        """Integer part of K2 (mulhot_index.py:48-67): (flat token index, segment ids)."""
        ts = self.sets[prefix]
        ids = self._ids(ids)
        f = attr - ts.n_cat
        lens = self.att[prefix][3][f][ids.long()].long() if attr >= ts.n_cat else torch.ones_like(ids).long()
        offs = torch.zeros(ids.numel() + 1, dtype=torch.int64, device=self.device)
        offs[1:] = torch.cumsum(lens, 0)
        total = int(offs[-1].item())
        flat = torch.empty(total, dtype=torch.int32, device=self.device)
        seg = torch.empty(total, dtype=torch.int32, device=self.device)
        call('arx_mulhot_flat_index', ts.desc_ptr(0), attr, ids.data_ptr(), ids.numel(), offs.data_ptr(),
             flat.data_ptr(), seg.data_ptr())
        return flat, seg

    # -- embed_attribute.py:222-237 --------------------------------------------------------
    def get_batch_user(self, keep_prob, concat=True, no_id=False, device='/gpu:0', dropout_mask=None,
                       u_inds=None):
        ua = self.user_attributes
        ids = self.u_indices['input'] if u_inds is None else self._ids(u_inds)
        if no_id and ua.num_features_cat == 1:                         # :356-366
            z = torch.zeros((ids.numel(), self.dim), dtype=torch.float32, device=self.device)
            self._last_user = None
            return z, None
        mode = POOL_CONCAT if concat else POOL_MEAN
        out, _, rng = self.pool('user', ids, mode, False, no_id=no_id)
        self._last_user = ('user', rng, ids, mode)
        out = self.dropout(out, keep_prob, dropout_mask)               # :236
        return out, None

    def dropout(self, x, keep_prob, mask=None):
        """tf.nn.dropout: x / keep * floor(keep + U[0,1)); identity at keep_prob == 1."""
        if keep_prob == 1.0:
            return x
        if mask is None:
            mask = getattr(self, '_premade_mask', None)
            self._premade_mask = None
            if mask is None or mask.shape != x.shape:
                mask = torch.floor(torch.rand_like(x) + keep_prob)
        y = torch.empty_like(x)
        call('arx_scale_mask', x.data_ptr(), mask.data_ptr(), 1.0 / keep_prob, x.numel(), y.data_ptr())
        self._last_dropout_mask = mask
        return y

    def premake_dropout_mask(self, shape, keep_prob):
        """Draw the next dropout mask now (it depends on nothing): the three tiny generator kernels then sit
        in front of the lookups instead of on the dependent chain behind them."""
        self._premade_mask = None
        if keep_prob != 1.0:
            self._premade_mask = torch.floor(torch.rand(shape, dtype=torch.float32, device=self.device) + keep_prob)

    # -- embed_attribute.py:239-254 --------------------------------------------------------
    def get_batch_item(self, name, batch_size, concat=False, keep_prob=1.0, no_attribute=False,
                       device='/gpu:0', reduce=None):
        """reduce=None mirrors the reference (list of per-attribute tensors, or the concat);
        reduce='mean' returns the fused mean over attributes directly."""
        assert name in self.i_indices
        assert keep_prob == 1.0, 'otherwise not implemented'
        ids = self.i_indices[name]
        if reduce == 'mean':
            out, b, rng = self.pool('item', ids, POOL_MEAN, True, no_attribute=no_attribute)
            return out, b
        out, b, (a0, na) = self.pool('item', ids, POOL_CONCAT, True, no_attribute=no_attribute)
        if concat:
            return out, b
        return [out[:, f * self.dim:(f + 1) * self.dim] for f in range(na)], b

    def get_sampled_item(self, n_sampled, device='/gpu:0'):
        """embed_attribute.py:256-263 (unused by HMF)."""
        out, b, _ = self.pool('item', self.sampled_ids, POOL_MEAN, True)
        return out, b

    # -- embed_attribute.py:320-348 --------------------------------------------------------
    def pass_sampled_items(self, item_sampled):
        """`update_sampled`: remember the sampled pool (its CSR is read straight from the
        attribute store by the kernels, no materialisation) and the item -> pool-position map."""
        new_ids = self._ids(item_sampled)
        if self.sampled_ids is not None and self.sampled_ids.shape == new_ids.shape:
            self.sampled_ids.copy_(new_ids)      # in place: captured CUDA graphs keep reading this buffer
        else:
            self.sampled_ids = new_ids.clone()
        n_items = self.item_attributes.num_entities
        if self.sampled_pos_dev is None:
            self.sampled_pos_dev = torch.full((n_items,), -1, dtype=torch.int32, device=self.device)
        else:
            self.sampled_pos_dev.fill_(-1)
        self.sampled_pos_dev[self.sampled_ids.long()] = torch.arange(
            self.sampled_ids.numel(), dtype=torch.int32, device=self.device)
        self._sampled_version = getattr(self, '_sampled_version', 0) + 1
        for k in ('mw_train', 'mw_eval'):       # masked columns = pool positions: refresh in place
            if k in self.pos_csr:
                items_d = self._pos_items_dev[k.split('_')[1]]
                self.pos_csr[k][1].copy_(self.sampled_pos_dev[items_d])
        for ts in self.sets.values():
            ts.plans.pop('sampled', None)

    # -- pooled catalog (the rewritten K3) ---------------------------------------------------
    def _out_prefix(self):
        return 'item_output' if self.item_output else 'item'

    def pool_catalog(self, pool='full', output_feat=1):
        ids = self.catalog_ids if pool == 'full' else self.sampled_ids
        P, beta, _ = self.pool(self._out_prefix(), ids, POOL_MEAN, True, no_attribute=(output_feat == 0))
        return P, beta, ids

    # -- embed_attribute.py:148-206 --------------------------------------------------------
    def get_prediction(self, latent, pool='full', device='/gpu:0', output_feat=1):
        """logits [mb, V] (or [mb, S]).  output_feat 0 (id only) and 1 (mean pooling) are linear
        in the tables and use the pool-first form; 2 / 3 (max / log-sum-exp pooling of the token scores) are not, and
        take the reference's literal order (token_prediction)."""
        if output_feat not in (0, 1, 2, 3):
            print('Error: Attribute combination not implemented!')
            exit(1)
        if isinstance(latent, list):
            raise NotImplementedError('per-attribute latent lists are not supported on the CUDA path')
        if output_feat in (2, 3):
            return self.token_prediction(latent, pool, output_feat)
        P, beta, ids = self.pool_catalog(pool, output_feat)
        mb, N = latent.shape[0], P.shape[0]
        logits = torch.empty((mb, N), dtype=torch.float32, device=self.device)
        _lib.gemm(latent, P, logits, mb, N, self.dim, 0, 1, beta)
        self._last_pred = (latent, P, beta, ids, pool, output_feat)
        return logits

    # -- embed_attribute.py:163-205 for output_feat 2 / 3 ------------------------------------------------
    def _catalog_csr(self, pool):
        """Per attribute, the catalog in logit (or pool) order as (token ids, ptr): the `full_*_tr` / sampled constants
        of the reference (:97-108, :320-348) in CSR form, built once per catalog / pool refresh."""
        key = ('csr', pool, getattr(self, '_sampled_version', 0) if pool != 'full' else 0)
        cache = getattr(self, '_csr_cache', {})
        if key in cache:
            return cache[key]
        ids = (self.catalog_ids if pool == 'full' else self.sampled_ids).long()
        feats_cat, feats_mul, starts, lengths = self.att['item'][:4]
        cats = [fc[ids].contiguous() for fc in feats_cat]
        muls = []
        for f in range(len(feats_mul)):
            l = lengths[f][ids].long()
            st = starts[f][ids].long()
            ptr_ = torch.zeros(ids.numel() + 1, dtype=torch.int64, device=self.device)
            ptr_[1:] = torch.cumsum(l, 0)
            seg = torch.repeat_interleave(torch.arange(ids.numel(), device=self.device), l)
            pos = torch.arange(int(ptr_[-1].item()), device=self.device) - ptr_[:-1][seg] + st[seg]
            muls.append((feats_mul[f][pos].contiguous(), ptr_))
        self._csr_cache = {k: v for k, v in cache.items() if k[1] != pool}
        self._csr_cache[key] = (cats, muls)
        return cats, muls

    def token_prediction(self, latent, pool='full', output_feat=2):
        """logits [n, V] in the reference's literal order: per attribute f the token scores s_f = E_f latent^T + beta_f
        ([V_f, n], one contraction), pooled per catalog item by segment max (2) or m + log(1 + sum exp(s - m)) (3), a
        categorical attribute contributing the score of its single token; mean over attributes (:205).  Keeps what
        token_prediction_backward needs in self._last_tokpred."""
        assert self.shard is None, 'output_feat 2 / 3 is not provided for row-sharded tables'
        pre = self._out_prefix()
        ts = self.sets[pre]
        n = latent.shape[0]
        V = (self.catalog_ids if pool == 'full' else self.sampled_ids).numel()
        cats, muls = self._catalog_csr(pool)
        F = ts.n_attr
        out_vm = torch.zeros((V, n), dtype=torch.float32, device=self.device)
        lat = latent.contiguous()
        saved = []
        for k in range(F):
            E, b = self.params[ts.names[k]], self.params[ts.bias_names[k]].reshape(-1)
            Vf = E.shape[0]
            S = torch.empty((Vf, n), dtype=torch.float32, device=self.device)
            _lib.gemm(E, lat, S, Vf, n, self.dim, 0, 1)
            if k < ts.n_cat:
                vals, ptr_, mode = cats[k], None, 1
            else:
                (vals, ptr_), mode = muls[k - ts.n_cat], output_feat
            arg = torch.empty((V, n), dtype=torch.int32, device=self.device) if mode == 2 else None
            den = torch.empty((V, n), dtype=torch.float32, device=self.device) if mode == 3 else None
            pk = None
            if mode == 3:
                pk = torch.zeros(1, dtype=torch.int64, device=self.device)
                call('arx_score_max', S.data_ptr(), b.data_ptr(), Vf, n, pk.data_ptr())
            call('arx_token_pool_fwd', S.data_ptr(), b.data_ptr(), n, vals.data_ptr(), ptr(ptr_), V, mode, ptr(pk), 1.0 / F,
                 out_vm.data_ptr(), ptr(arg), ptr(den))
            saved.append((k, S, vals, ptr_, mode, pk, arg, den))
        logits = torch.empty((n, V), dtype=torch.float32, device=self.device)
        call('arx_transpose', out_vm.data_ptr(), V, n, logits.data_ptr(), 0)
        self._last_tokpred = (lat, saved, V, pool, output_feat)
        self._last_pred = (latent, None, None, self.catalog_ids if pool == 'full' else self.sampled_ids, pool, output_feat)
        return logits

    def token_prediction_backward(self, dlogits):
        """Adjoint of token_prediction for d(loss)/d(logits) [n, V]: returns d(latent) [n, dim] and ACCUMULATES the dense
        gradients of the output tables and biases (every token row scores every batch row) into self.dense_table_grads —
        applied, merged with any sparse look-up gradients of the same tables, by apply_gradients."""
        lat, saved, V, pool, output_feat = self._last_tokpred
        pre = self._out_prefix()
        ts = self.sets[pre]
        n, F = lat.shape[0], ts.n_attr
        d_vm = torch.empty((V, n), dtype=torch.float32, device=self.device)
        call('arx_transpose', dlogits.data_ptr(), n, V, d_vm.data_ptr(), 0)
        dlat = torch.zeros((n, self.dim), dtype=torch.float32, device=self.device)
        grads = self.dense_table_grads.setdefault(pre, {})
        scratch = torch.zeros(1, dtype=torch.float32, device=self.device)
        for (k, S, vals, ptr_, mode, pk, arg, den) in saved:
            E, b = self.params[ts.names[k]], self.params[ts.bias_names[k]].reshape(-1)
            Vf = E.shape[0]
            dS = torch.zeros((Vf, n), dtype=torch.float32, device=self.device)
            call('arx_token_pool_bwd', d_vm.data_ptr(), S.data_ptr(), b.data_ptr(), n, vals.data_ptr(), ptr(ptr_), V, mode,
                 ptr(pk), 1.0 / F, ptr(arg), ptr(den), dS.data_ptr(), scratch.data_ptr())
            name, bname = ts.names[k], ts.bias_names[k]
            first = name not in grads
            if first:
                grads[name] = torch.empty_like(E)
                grads[bname] = torch.empty((Vf,), dtype=torch.float32, device=self.device)
            _lib.gemm(dS, lat, grads[name], Vf, self.dim, n, 0, 0, None, 1.0, 0.0 if first else 1.0)      # dE_f (+)= dS latent
            call('arx_rowsum', dS.data_ptr(), Vf, n, grads[bname].data_ptr(), 0 if first else 1)
            _lib.gemm(dS, E, dlat, n, self.dim, Vf, 1, 0, None, 1.0, 1.0)                                  # d latent += dS^T E_f
        return dlat

    def fused_ce(self, latent, targets, row_scale=None, want_grad=True, pool='full', output_feat=1):
        """get_prediction (:148-206) + compute_loss(..., 'ce') (:530) + their gradients without ever writing
        the [rows, V] logits: arx_ce_fwd / arx_ce_rowloss / arx_ce_bwd on tf32-rounded operands.
        Returns (loss_rows [rows], (dU, dP, dbeta) or None), or None when the shape is not supported
        (the caller then takes the materialised path)."""
        if output_feat not in (0, 1) or isinstance(latent, list):
            return None
        M, d = latent.shape[0], self.dim
        ids = self.catalog_ids if pool == 'full' else self.sampled_ids
        N = ids.numel()
        if not _lib.ce_supported(M, N, d):
            return None
        P, beta, ids = self.pool_catalog(pool, output_feat)
        U_r = _lib.round_tf32(latent if latent.is_contiguous() else latent.contiguous())
        P_r = _lib.round_tf32(P)
        lse = _lib.ce_fwd(U_r, P_r, beta, M, N, d)
        if lse is None:
            return None
        tgt = targets if isinstance(targets, torch.Tensor) else self._ids(targets)
        loss = _lib.ce_rowloss(U_r, P_r, beta, tgt, lse, M, N, d)
        self._last_pred = (latent, P, beta, ids, pool, output_feat)
        if not want_grad:
            return loss, None
        g = row_scale if row_scale is not None else torch.ones(M, dtype=torch.float32, device=self.device)
        grads = _lib.ce_bwd(U_r, P_r, beta, lse, g, tgt, M, N, d)
        if grads is None:
            return None
        return loss, grads

    def mw_mask(self, M, S, forward_only=False, pos_rows=None):
        """Bit matrix of the positives that sit in the sampled pool (arx_mw_mask_build): depends on the batch's user ids
        and the pool only, so the step builds it on a side stream under the lookups."""
        key = 'mw' + ('_eval' if forward_only else '_train')
        pos_ptr, pos_idx = self._positives(key)
        pos_row = pos_rows if pos_rows is not None else self.u_indices['input']
        ld = _lib.mw_mask_words(S)
        mask = torch.empty((M, ld), dtype=torch.int32, device=self.device)
        call('arx_mw_mask_build', pos_row.data_ptr(), pos_ptr.data_ptr(), pos_idx.data_ptr(), M, S, mask.data_ptr(), ld)
        return mask, ld

    def fused_mw(self, latent, Ps, bs, tscore, row_scale, want_grad, forward_only=False, pos_rows=None, dP=None,
                 prepared=None, mask=None, outputs=None):
        """Sampled-pool scoring (:148-206, pool='sampled') + _compute_mw_loss (:641-649) + their gradients on the
        tensor cores without writing the [rows, S] scores: arx_mw_mask_build / arx_mw_fwd / arx_mw_bwd.
        Returns (loss_rows, (dU, dPs, dbs, dts) or None), or None when the shape is not supported."""
        M, d, S = latent.shape[0], self.dim, Ps.shape[0]
        if not _lib.ce_supported(M, S, d):
            return None
        if mask is None:
            mask, ld = self.mw_mask(M, S, forward_only, pos_rows)
        else:
            mask, ld = mask
        UT = PT = None
        if prepared is not None:               # arx_mw_prep already rounded / transposed the operands
            U_r, P_r, UT, PT = prepared
        else:
            U_r = _lib.round_tf32(latent if latent.is_contiguous() else latent.contiguous())
            P_r = _lib.round_tf32(Ps)
        fw = _lib.mw_fwd(U_r, P_r, bs, tscore, mask, ld, M, S, d)
        if fw is None:
            return None
        hsum, loss = fw
        if not want_grad:
            return loss, None
        g = row_scale if row_scale is not None else torch.ones(M, dtype=torch.float32, device=self.device)
        grads = _lib.mw_bwd(U_r, P_r, bs, tscore, mask, ld, hsum, g, M, S, d, dP=dP, UT=UT, PT=PT, outputs=outputs)
        if grads is None:
            return None
        return loss, grads

    # the transforms of the masked hinge sum S (embed_attribute.py:580-594): (l(S), dl/dS) as device vectors
    @staticmethod
    def _rs_transform(S, loss, loss_func, p):
        lf = 'log' if loss == 'warp' else loss_func
        if lf == 'log':
            return torch.log1p(S), 1.0 / (1.0 + S)
        if lf == 'exp':
            q = torch.pow(torch.full_like(S, p), -S)
            return 1.0 - q, float(np.log(p)) * q
        if lf == 'poly':
            return torch.pow(S, p), p * torch.pow(S, p - 1.0)
        if lf == 'poly2':
            return torch.pow(1.0 + S, p), p * torch.pow(1.0 + S, p - 1.0)
        if lf == 'linear':
            return S, torch.ones_like(S)
        return S * S, 2.0 * S                                            # 'square'

    def fused_warp(self, latent, targets, loss='warp', loss_func='log', exp_p=1.005, row_scale=None, want_grad=True,
                   forward_only=False, pos_rows=None, unmasked=False, output_feat=1, max_mask_bytes=1 << 30):
        """Full-catalog WMRB without the [rows, V] scores: get_prediction (:148-206) + _compute_warp_loss (:605-618) /
        the hinge members of the rs family (:551-603: 'rs' with any loss_func) + their gradients, on the sampled-WMRB
        tensor-core kernels run over the WHOLE catalog (arx_mw_fwd / arx_mw_bwd2 with N = V): the hinge sum per row comes
        from the forward pass, its transform l(S) and l'(S) are per-row scalars applied here, and the gradient that
        reaches the target's own score (column t of the reference's logits) is scattered back to U, P[t] and beta[t].
        The positives of a row travel as one bit per (row, item), built per row block (<= max_mask_bytes) from the
        per-user CSR.  Returns (loss_rows [rows], (dU, dP, dbeta) or None), or None when the shape / loss is not covered
        (the caller then materialises the scores: arx_gemm_tc + arx_loss_rows)."""
        if loss not in ('warp', 'rs') or output_feat not in (0, 1) or isinstance(latent, list):
            return None
        M, d = latent.shape[0], self.dim
        N = self.catalog_ids.numel()
        if not _lib.ce_supported(M, N, d):
            return None
        P, beta, ids = self.pool_catalog('full', output_feat)
        U_r = _lib.round_tf32(latent if latent.is_contiguous() else latent.contiguous())
        P_r = _lib.round_tf32(P)
        tgt = (targets if isinstance(targets, torch.Tensor) else self._ids(targets)).long()
        # the target's own score from the same rounded operands the kernels contract (it is one of their columns)
        Pt_r = P_r[tgt]
        tscore = torch.empty((M,), dtype=torch.float32, device=self.device)
        call('arx_rowdot_fwd', U_r.data_ptr(), Pt_r.data_ptr(), beta[tgt].contiguous().data_ptr(), M, d, tscore.data_ptr())
        pos_ptr = pos_idx = pos_row = None
        if not unmasked:
            pos_ptr, pos_idx = self._positives('full' + ('_eval' if forward_only else '_train'))
            pos_row = pos_rows if pos_rows is not None else self.u_indices['input']
        ld = _lib.mw_mask_words(N)
        rows_per_block = max(4, min(M, (max_mask_bytes // (4 * ld)) // 4 * 4))
        g = row_scale if row_scale is not None else torch.ones(M, dtype=torch.float32, device=self.device)
        loss_rows = torch.empty((M,), dtype=torch.float32, device=self.device)
        dU = dP = dbeta = None
        if want_grad:
            dU = torch.zeros((M, d), dtype=torch.float32, device=self.device)
            dP = torch.zeros((N, d), dtype=torch.float32, device=self.device)
            dbeta = torch.zeros((N,), dtype=torch.float32, device=self.device)
            PT = torch.empty((d, N), dtype=torch.float32, device=self.device)
            call('arx_transpose', P_r.data_ptr(), N, d, PT.data_ptr(), 0)
        for r0 in range(0, M, rows_per_block):
            r1 = min(M, r0 + rows_per_block)
            n = r1 - r0
            mask = torch.empty((n, ld), dtype=torch.int32, device=self.device)
            if unmasked:
                mask.zero_()
            else:
                call('arx_mw_mask_build', pos_row[r0:r1].contiguous().data_ptr(), pos_ptr.data_ptr(), pos_idx.data_ptr(), n, N,
                     mask.data_ptr(), ld)
            fw = _lib.mw_fwd(U_r[r0:r1], P_r, beta, tscore[r0:r1], mask, ld, n, N, d)
            if fw is None:
                return None
            hsum, _ = fw
            l, dl = self._rs_transform(hsum, loss, loss_func, float(exp_p))
            loss_rows[r0:r1] = l
            if not want_grad:
                continue
            # arx_mw_bwd2 scales the hinge indicator by g / (1 + hsum): hand it g * l'(S) * (1 + S)
            gk = (g[r0:r1] * dl * (1.0 + hsum)).contiguous()
            dts = torch.zeros((n,), dtype=torch.float32, device=self.device)
            out = _lib.mw_bwd(U_r[r0:r1], P_r, beta, tscore[r0:r1], mask, ld, hsum, gk, n, N, d, PT=PT,
                              outputs=(dU[r0:r1], dP, dbeta, dts), accumulate=True)     # adds into the zeroed outputs
            if out is None:
                return None
            # the target column: d loss / d s_t = dts (minus the row sum of the score gradients)
            t = tgt[r0:r1]
            dU[r0:r1].addcmul_(dts.unsqueeze(1), P[t])
            dP.index_add_(0, t, dts.unsqueeze(1) * latent[r0:r1])
            dbeta.index_add_(0, t, dts)
            del mask
        self._last_pred = (latent, P, beta, ids, 'full', output_feat)
        return loss_rows, ((dU, dP, dbeta) if want_grad else None)

    def score_topk(self, latent, k, output_feat=1, want_lse=False, chunk=1 << 16):
        """get_prediction (:148-206) + tf.nn.top_k(sorted=True) (hmf/hmf_model.py:154, lstm/seqModel.py:514-519) without
        the [rows, V] scores: the catalog is scored in column blocks of `chunk` items on the tensor cores, every block
        is reduced to its k best by arx_topk_rows, and the candidates (block order = index order) are reduced once more —
        bit-equal to arx_topk_rows over the full matrix, ties -> lower index.  Returns (idx int32 [rows, k],
        val [rows, k], lse [rows] or None) with lse = log sum_v exp(score) for the softmax probabilities of :515."""
        if output_feat in (2, 3):
            logits = self.get_prediction(latent, 'full', output_feat=output_feat)
            P = beta = None
            N = logits.shape[1]
        else:
            P, beta, _ = self.pool_catalog('full', output_feat)
            N = P.shape[0]
        M = latent.shape[0]
        k = int(k)
        lat = latent if latent.is_contiguous() else latent.contiguous()
        n_blk = (N + chunk - 1) // chunk
        if n_blk == 1 or k > chunk:
            if P is not None:
                logits = torch.empty((M, N), dtype=torch.float32, device=self.device)
                _lib.gemm(lat, P, logits, M, N, self.dim, 0, 1, beta)
            idx = torch.empty((M, k), dtype=torch.int32, device=self.device)
            val = torch.empty((M, k), dtype=torch.float32, device=self.device)
            call('arx_topk_rows', logits.data_ptr(), M, N, logits.stride(0), k, idx.data_ptr(), val.data_ptr())
            return idx, val, (torch.logsumexp(logits, 1) if want_lse else None)
        cand_v = torch.empty((M, n_blk * k), dtype=torch.float32, device=self.device)
        cand_i = torch.empty((M, n_blk * k), dtype=torch.int32, device=self.device)
        lse = torch.full((M,), float('-inf'), dtype=torch.float32, device=self.device) if want_lse else None
        S = torch.empty((M, min(chunk, N)), dtype=torch.float32, device=self.device)
        bi = torch.empty((M, k), dtype=torch.int32, device=self.device)
        bv = torch.empty((M, k), dtype=torch.float32, device=self.device)
        for b in range(n_blk):
            c0, c1 = b * chunk, min(N, (b + 1) * chunk)
            w = c1 - c0
            kb = min(k, w)
            if P is not None:
                Sv = S.view(-1)[:M * w].view(M, w)                       # dense [M, w] block (the last one is narrower)
                _lib.gemm(lat, P[c0:c1], Sv, M, w, self.dim, 0, 1, beta[c0:c1])
            else:
                Sv = logits[:, c0:c1]
            call('arx_topk_rows', Sv.data_ptr(), M, w, Sv.stride(0), kb, bi.data_ptr(), bv.data_ptr())
            cand_v[:, b * k:b * k + kb] = bv.view(-1)[:M * kb].view(M, kb) if kb < k else bv
            cand_i[:, b * k:b * k + kb] = (bi.view(-1)[:M * kb].view(M, kb) if kb < k else bi) + c0
            if kb < k:
                cand_v[:, b * k + kb:(b + 1) * k] = float('-inf')
                cand_i[:, b * k + kb:(b + 1) * k] = 0
            if want_lse:
                lse = torch.logaddexp(lse, torch.logsumexp(Sv, 1))
        pos = torch.empty((M, k), dtype=torch.int32, device=self.device)
        val = torch.empty((M, k), dtype=torch.float32, device=self.device)
        call('arx_topk_rows', cand_v.data_ptr(), M, n_blk * k, cand_v.stride(0), k, pos.data_ptr(), val.data_ptr())
        idx = torch.gather(cand_i, 1, pos.long())
        return idx.contiguous(), val, lse

    # -- embed_attribute.py:208-220 --------------------------------------------------------
    def get_target_score(self, latent, inds, device='/gpu:0'):
        ids = self._ids(inds)
        Pt, bt, _ = self.pool(self._out_prefix(), ids, POOL_MEAN, True)
        out = torch.empty((ids.numel(),), dtype=torch.float32, device=self.device)
        call('arx_rowdot_fwd', latent.data_ptr(), Pt.data_ptr(), bt.data_ptr(), ids.numel(), self.dim,
             out.data_ptr())
        self._last_target = (latent, Pt, ids)
        return out

    # -- embed_attribute.py:510-523 --------------------------------------------------------
    def get_user_model_size(self, no_id=False, concat=True):
        ua = self.user_attributes
        if concat:
            cat_start = 1 if no_id else 0
            return (sum(ua._embedding_size_list_cat[cat_start:ua.num_features_cat]) +
                    sum(ua._embedding_size_list_mulhot[0:ua.num_features_mulhot]))
        return ua._embedding_size_list_cat[0]

    def get_item_model_size(self, concat=True):
        ia = self.item_attributes
        if concat:
            return (sum(ia._embedding_size_list_cat[0:ia.num_features_cat]) +
                    sum(ia._embedding_size_list_mulhot[0:ia.num_features_mulhot]))
        return ia._embedding_size_list_cat[0]

    # -- embed_attribute.py:525-649 --------------------------------------------------------
    def compute_loss(self, logits, item_target, loss='ce', true_rank=False, loss_func='log',
                     exp_p=1.005, device='/gpu:0', row_scale=None, want_grad=True, pos_rows=None,
                     forward_only=False, unmasked=False):
        """Per-row loss [mb].  With want_grad the gradient d(sum_b row_scale_b * loss_b)/d logits
        overwrites `logits` in place (the scores are not needed afterwards) and, for 'mw',
        d/d target-score is returned through self._last_dtarget."""
        assert loss in ['ce', 'mce', 'warp', 'warp_eval', 'rs', 'rs-sig', 'rs-sig2', 'mw', 'bbpr',
                        'bpr', 'bpr-hinge']
        if loss in ('mce', 'bpr', 'bpr-hinge'):
            # 'mce' has no branch in the reference (:527 vs :529-549); bpr/bpr-hinge inputs are
            # never fed there (embed_attribute.py:704-706 commented out).
            print('Error: not implemented other loss!!')
            exit(1)
        mb, V = logits.shape
        out = torch.empty((mb,), dtype=torch.float32, device=self.device)
        tgt = ts = None
        if loss == 'mw':
            ts = item_target
        else:
            tgt = item_target if isinstance(item_target, torch.Tensor) else self._ids(item_target)
        pos_row = pos_ptr = pos_idx = None
        if loss != 'ce' and not unmasked:
            key = ('mw' if loss == 'mw' else 'full') + ('_eval' if forward_only else '_train')
            pos_ptr, pos_idx = self._positives(key)
            pos_row = pos_rows if pos_rows is not None else self.u_indices['input']
        rank = torch.empty((mb,), dtype=torch.int64, device=self.device) if (true_rank or loss == 'warp_eval') else None
        dts = torch.empty((mb,), dtype=torch.float32, device=self.device) if (loss == 'mw' and want_grad) else None
        kind, func = _lib.LOSS_KIND[loss], _lib.LOSS_FUNC[loss_func]
        if loss == 'warp_eval':
            # _compute_warp_eval_loss (:620-639) reports the RAW margin rank sum_v mask * relu(1 + s_v - s_t), not its
            # log: the hinge sum with the identity transform (found by the golden test through model.step)
            kind, func = _lib.LOSS_KIND['rs'], _lib.LOSS_FUNC['linear']
        call('arx_loss_rows', logits.data_ptr(), mb, V, logits.stride(0), ptr(tgt), ptr(ts), ptr(pos_row),
             ptr(pos_ptr), ptr(pos_idx), kind, func, float(exp_p),
             ptr(row_scale), out.data_ptr(), logits.data_ptr() if want_grad else None, ptr(dts), ptr(rank))
        self._last_dtarget = dts
        if loss == 'warp_eval' or true_rank:
            return [out, rank]
        return out

    # -- positives: embed_attribute.py:651-677, :721-747 -------------------------------------
    def get_warp_mask(self, device='/gpu:0'):
        """The reference returns scatter_update ops on a dense bool Variable; here the mask is the
        per-user CSR consumed inside the loss kernel, so there is nothing to run."""
        self.set_mask, self.reset_mask = {}, {}
        return self.set_mask, self.reset_mask

    def prepare_warp(self, pos_item_set, pos_item_set_eval):
        self.pos_item_set = pos_item_set
        self.pos_item_set_eval = pos_item_set_eval
        self.pos_csr = {}
        self._pos_host = {}
        self._pos_items_dev = {}

    def _positives_host(self, which):
        """CSR over users of positive ITEM indices (host, int32), built once per item set."""
        if which in self._pos_host:
            return self._pos_host[which]
        item_set = self.pos_item_set_eval if which == 'eval' else self.pos_item_set
        n_users = self.user_attributes.num_entities
        if isinstance(item_set, tuple):          # already CSR (ptr, items) — large synthetic sets
            ptr_, items = item_set
        else:
            lens = np.zeros(n_users, dtype=np.int64)
            for u, v in item_set.items():
                lens[u] = len(v)
            ptr_ = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
            items = np.empty(int(ptr_[-1]), dtype=np.int32)
            for u, v in item_set.items():
                items[ptr_[u]:ptr_[u + 1]] = np.fromiter(v, dtype=np.int32, count=len(v))
        self._pos_host[which] = (np.asarray(ptr_, dtype=np.int32), np.asarray(items, dtype=np.int32))
        return self._pos_host[which]

    def _positives(self, key):
        """Device CSR (ptr, idx) for key in {full,mw}_{train,eval}: idx are masked COLUMN ids —
        logit indices (sorted per user) for the full losses, sampled-pool positions (or -1) for mw."""
        if key in self.pos_csr:
            return self.pos_csr[key]
        kind, which = key.split('_')
        ptr_h, items_h = self._positives_host(which)
        ptr_d = torch.from_numpy(ptr_h).to(self.device)
        items_d = torch.from_numpy(items_h).to(self.device).long()
        if not hasattr(self, '_pos_items_dev'):
            self._pos_items_dev = {}
        self._pos_items_dev[which] = items_d
        if kind == 'full':
            cols = self.item2logit_dev[items_d]
            # sort columns inside each user's slice: one global sort on (user, col) keys
            seg = torch.repeat_interleave(torch.arange(len(ptr_h) - 1, device=self.device),
                                          (ptr_d[1:] - ptr_d[:-1]).long())
            keyv = seg * (int(self.logit_size) + 1) + (cols.long() + 1)
            cols = cols[torch.argsort(keyv)]
        else:
            cols = self.sampled_pos_dev[items_d]
        self.pos_csr[key] = (ptr_d, cols.to(torch.int32).contiguous())
        return self.pos_csr[key]

    # -- embed_attribute.py:679-684 --------------------------------------------------------
    def target_mapping(self, item_target):
        m = self.item_ind2logit_ind
        target = []
        for items in item_target:
            if isinstance(items, torch.Tensor):
                target.append(self.item2logit_dev[items.long()])
            elif isinstance(items, np.ndarray) and hasattr(m, 'as_array'):
                target.append(items)
            else:
                target.append([m[v] for v in items])
        return target

    # -- embed_attribute.py:697-747 --------------------------------------------------------
    def add_input(self, input_feed, user_input, item_input, neg_item_input=None, item_sampled=None,
                  item_sampled_id2idx=None, forward_only=False, recommend=False, loss=None):
        """Feed the 'placeholders'.  Returns the reference's triple; the sampled-pool refresh is
        performed right here (it was a separate session.run in hmf_model.py:206-207)."""
        if self.user_attributes is not None:
            self.u_indices['input'] = self._ids(user_input)
        if self.item_attributes is not None and self.input_steps > 0 and item_input is not None:
            for step in range(len(item_input)):
                self.i_indices['input{}'.format(step)] = self._ids(item_input[step])
        update_sampled = []
        if (self.item_attributes is not None and recommend is False and item_sampled is not None
                and loss in ['mw', 'mce']):
            self.pass_sampled_items(item_sampled)
            update_sampled = ['update_sampled']
        return update_sampled, {}, {}

    # ------------------------------------------------------------------ backward ------
    def push_grad(self, prefix, rng, ids, mode, dout, dbias=None, plan_key=None):
        """Register d(loss)/d(pooled output) of one lookup (the IndexedSlices of tf.gradients)."""
        a0, na = rng
        self.sets[prefix].pending.append((a0, na, ids, mode, dout, dbias, plan_key))

    def _plan_for(self, ts, entries):
        """Build (or fetch the cached / prefetched) backward plan for the pending lookups of one table set."""
        single_key = entries[0][6] if len(entries) == 1 else None
        if single_key is not None and single_key in ts.plans:
            return ts.plans[single_key]
        specs = [(a0, na, ids, mode) for (a0, na, ids, mode, dout, dbias, key) in entries]
        pre = getattr(ts, '_prefetched', None)
        if pre is not None:
            plan, sig, ev = pre
            ts._prefetched = None
            torch.cuda.current_stream().wait_event(ev)           # joins the side stream in any case
            if sig == self._plan_sig(specs):
                return plan
        return self._build_plan(ts, specs, single_key)

    @staticmethod
    def _plan_sig(specs):
        return [(a0, na, ids.data_ptr(), ids.numel(), mode) for (a0, na, ids, mode) in specs]

    def prefetch_plans(self, requests):
        """A backward plan depends only on the entity ids of the lookups, not on any gradient: start
        building it at the top of the step on a side stream, so that the latency-bound plan kernels
        overlap the forward pass.  requests: {prefix: [((a0, na), ids, mode), ...]} in the order the
        matching push_grad() calls will come; apply_gradients() picks the plan up when the signature
        matches and rebuilds otherwise."""
        if _lib.timeline is not None:
            return                                # the per-kernel timing pass runs everything serially
        main = torch.cuda.current_stream()
        for k, (prefix, reqs) in enumerate(requests.items()):
            ts = self.sets[prefix]
            specs = [(rng[0], rng[1], ids, mode) for (rng, ids, mode) in reqs]
            side = self.side_stream(8 + k)
            side.wait_stream(main)
            with torch.cuda.stream(side):
                plan = self._build_plan(ts, specs, None)
                ev = torch.cuda.Event()
                ev.record(side)
            ts._prefetched = (plan, self._plan_sig(specs), ev)

    def _build_plan(self, ts, specs, single_key, heavy=None):
        cap_occ = 0
        for (a0, na, ids, mode) in specs:
            cap_occ += int(ids.numel()) * sum(ts.max_len[a0:a0 + na])
        if single_key is not None and single_key.startswith('catalog'):
            ia = self.item_attributes     # exact: the catalog CSR is static
            a0, na = specs[0][0], specs[0][1]
            cap_occ = sum(int(specs[0][2].numel()) if a < ts.n_cat else len(ia.full_values_tr[a - ts.n_cat])
                          for a in range(a0, a0 + na))
        cap_occ = max(cap_occ, 1)
        cap_rows = max(min(cap_occ, ts.total_vocab), 1)
        plan = _Plan(self.device, cap_rows, cap_occ, self.dim) if single_key is not None else self._scratch_plan(ts, cap_rows, cap_occ)
        _lib.tag = ts.prefix
        _lib.load().arx_set_tuning(b'plan_agg', self.plan_agg)     # host-side launch-shape choice of THIS model
        # consecutive lookups over the same attribute range and mode (the sampled pool and the target items of an
        # `mw` step) are walked as ONE id list: their arena rows are consecutive in the same order
        merged = []
        for sp in specs:
            if merged and merged[-1][0] == sp[0] and merged[-1][1] == sp[1] and merged[-1][3] == sp[3]:
                merged[-1] = (sp[0], sp[1], merged[-1][2] + [sp[2]], sp[3])
            else:
                merged.append((sp[0], sp[1], [sp[2]], sp[3]))
        specs = [(a0, na, (l[0] if len(l) == 1 else torch.cat(l)), mode) for (a0, na, l, mode) in merged]
        plan._keep = specs                                  # the concatenated id lists live as long as the plan
        call('arx_bwd_plan_begin', plan.c)
        for (a0, na, ids, mode) in specs:
            call('arx_bwd_plan_count', ts.desc_ptr(0), a0, na, ids.data_ptr(), ids.numel(), plan.c)
        if heavy is None:
            call('arx_bwd_plan_alloc', ts.desc_ptr(0), plan.c)
        else:                                               # explicit chunk size (the column-slab apply repeats it)
            call('arx_bwd_plan_alloc_h', ts.desc_ptr(0), plan.c, int(heavy))
        row = 0
        for (a0, na, ids, mode) in specs:
            call('arx_bwd_plan_fill', ts.desc_ptr(0), a0, na, ids.data_ptr(), ids.numel(), mode, row, plan.c)
            row += ids.numel() * (na if mode == POOL_CONCAT else 1)
        call('arx_bwd_plan_end', ts.desc_ptr(0), plan.c)
        if self.deterministic:
            self._canonical_buckets(plan)
        if single_key is not None:
            ts.plans[single_key] = plan
        return plan

    @staticmethod
    def _canonical_buckets(plan):
        """Deterministic scatter mode: arx_bwd_plan_fill places the contributions of a table row into its
        bucket in the order its atomics happen to retire, so the fp32 sum of a row's gradient can differ in the last bit
        from run to run.  Here every bucket is put into ascending order of the contributing arena row (ties carry equal
        weights), which fixes the summation order of arx_pool_bwd_apply — also for hot rows, whose chunks are consecutive
        slices of the bucket folded in chunk order.  Needs the counters on the host: eager steps only (a debugging and
        regression-test mode; capture_step refuses it)."""
        if plan.counters.is_cuda and torch.cuda.is_current_stream_capturing():
            raise RuntimeError('deterministic scatter mode reads the plan counters on the host: not capturable')
        nu, occ = int(plan.counters[0].item()), int(plan.counters[1].item())
        if nu == 0 or occ == 0:
            return
        base = plan.row_base[:nu].long()
        cnt = plan.row_cnt[:nu].long()
        order = torch.argsort(base)
        seg = torch.repeat_interleave(torch.arange(nu, device=base.device), cnt[order])      # bucket index by position
        key = seg * (1 << 31) + plan.bucket_src[:occ].long()
        perm = torch.argsort(key)
        plan.bucket_src[:occ] = plan.bucket_src[:occ][perm]
        plan.bucket_w[:occ] = plan.bucket_w[:occ][perm]

    def _scratch_plan(self, ts, cap_rows, cap_occ):
        p = getattr(ts, '_scratch', None)
        if p is None or p.cap_rows < cap_rows or p.cap_occ < cap_occ:
            p = _Plan(self.device, cap_rows, cap_occ, self.dim)
            ts._scratch = p
        return p

    def _arena(self, entries):
        """Row arena [R, dim] (+ aligned bias grads [R]) of the pending lookups."""
        rows, biases, any_bias = [], [], any(e[5] is not None for e in entries)
        for (a0, na, ids, mode, dout, dbias, key) in entries:
            r = dout.reshape(-1, self.dim) if mode == POOL_CONCAT else dout
            rows.append(r)
            if any_bias:
                if dbias is None:
                    biases.append(torch.zeros(r.shape[0], dtype=torch.float32, device=self.device))
                else:
                    # the pooled bias is always the MEAN over attributes (:412), also in concat mode
                    biases.append((dbias / na).repeat_interleave(na) if mode == POOL_CONCAT else dbias)
        if len(rows) == 1:
            r0 = rows[0]
            strided_ok = (r0.dim() == 2 and r0.stride(1) == 1 and r0.stride(0) % 4 == 0 and r0.data_ptr() % 16 == 0)
            arena = r0 if (r0.is_contiguous() or strided_ok) else r0.contiguous()     # the kernels take a row stride
            bias = biases[0].contiguous() if any_bias else None
        else:
            arena = self._adjacent(rows)
            if arena is None:
                arena = torch.cat(rows, 0)
            bias = None
            if any_bias:
                bias = self._adjacent(biases)
                if bias is None:
                    bias = torch.cat(biases, 0)
        return arena, bias

    @staticmethod
    def _adjacent(parts):
        """If the tensors are consecutive contiguous row blocks of one allocation (the caller wrote its gradients
        straight into a shared arena), return the covering view instead of concatenating."""
        p0 = parts[0]
        st = p0.untyped_storage().data_ptr()
        p, n = p0.data_ptr(), 0
        for t in parts:
            if (not t.is_contiguous() or t.shape[1:] != p0.shape[1:] or t.dtype != p0.dtype or t.data_ptr() != p
                    or t.untyped_storage().data_ptr() != st):
                return None
            p += t.numel() * t.element_size()
            n += t.shape[0]
        return torch.as_strided(p0, (n,) + tuple(p0.shape[1:]), p0.stride())

    def _merge_dense(self):
        """Tables that hold a dense gradient (token_prediction_backward) AND pending look-up gradients get ONE summed dense
        gradient, as TensorFlow sums a variable's IndexedSlices and dense gradients before the norm and the update."""
        for prefix, grads in self.dense_table_grads.items():
            ts = self.sets[prefix]
            if grads and ts.pending:
                rows = self.row_gradients(prefix)                 # densified look-up gradients; clears ts.pending
                for name, g in rows.items():
                    if name in grads:
                        grads[name] += g.reshape(grads[name].shape)

    def apply_dense_table_grads(self, lr, opt=OPT_ADAGRAD, grad_scale=None):
        for prefix, grads in self.dense_table_grads.items():
            for name, g in grads.items():
                w, acc = self.params[name], self.accs[name]
                call('arx_dense_update', w.data_ptr(), acc.data_ptr(), g.contiguous().data_ptr(), w.numel(), float(lr),
                     ptr(grad_scale), opt)
        self.dense_table_grads = {}

    def sparse_sumsq(self, out, dense_semantics=()):
        """Add the table-gradient terms of clip_by_global_norm into out[0].  Table sets named in
        `dense_semantics` also feed the scoring matmul, so TF holds their gradient as ONE dense
        tensor (duplicates merged before the norm); the others are IndexedSlices (norm over the
        un-merged slices)."""
        self._merge_dense()
        for grads in self.dense_table_grads.values():
            for g in grads.values():
                out += (g * g).sum()
        for ts in self.sets.values():
            if not ts.pending:
                continue
            plan = self._plan_for(ts, ts.pending)
            arena, bias = self._arena(ts.pending)
            ts._ready = (plan, arena, bias)
            if ts.prefix in dense_semantics:
                # merged rows through the same segment-reduce as the update (ARX_OPT_NONE), then their norm
                rows = self._norm_scratch(ts, plan.cap_rows)
                brow = rows[1] if bias is not None else None
                call('arx_pool_bwd_apply', ts.desc_ptr(0), ts.n_attr, self.dim, plan.c, arena.data_ptr(),
                     arena.stride(0), ptr(bias), 0.0, None, OPT_NONE, rows[0].data_ptr(), ptr(brow))
                call('arx_rows_sumsq', rows[0].data_ptr(), ptr(brow), plan.c, self.dim, out.data_ptr())
            else:
                call('arx_pool_bwd_sumsq', ts.desc_ptr(0), self.dim, plan.c, arena.data_ptr(), arena.stride(0),
                     ptr(bias), out.data_ptr(), 0)

    def _norm_scratch(self, ts, cap_rows):
        sc = getattr(ts, '_norm_rows', None)
        if sc is None or sc[0].shape[0] < cap_rows:
            sc = (torch.empty((cap_rows, self.dim), dtype=torch.float32, device=self.device),
                  torch.empty((cap_rows,), dtype=torch.float32, device=self.device))
            ts._norm_rows = sc
        return sc

    def apply_gradients(self, lr, opt=OPT_ADAGRAD, grad_scale=None):
        """De-duplicated sparse optimizer step on every table set with pending gradients
        (hmf_model.py:146-151; lstm/seqModel.py:173-182)."""
        # The table sets are independent and each of their kernels is latency-bound on its own
        # (random 512-byte rows, L2 atomics): run them on parallel streams (parallel branches of the
        # captured CUDA graph) so that their memory traffic overlaps.
        if self.dense_table_grads:
            self._merge_dense()
            self.apply_dense_table_grads(lr, opt, grad_scale)
        main = torch.cuda.current_stream()
        busy = [ts for ts in self.sets.values() if ts.pending]
        if (busy and self.dim % 4 == 0 and opt in (OPT_ADAGRAD, OPT_SGD)
                and os.environ.get('ARX_APPLY_MANY', '0') == '1'):
            # every table set of the step in one launch (two per call): arx_pool_bwd_apply_many.  OFF by default:
            # measured at C2 the merged launch is slower than two launches on parallel streams (334-384 us vs
            # 117 + 138 us; the doubled body costs registers and spills) — kept as an A/B switch only.
            ready_all = []
            for ts in busy:
                ready = getattr(ts, '_ready', None)
                if ready is not None:
                    ts._ready = None
                else:
                    ready = (self._plan_for(ts, ts.pending), ) + self._arena(ts.pending)
                ready_all.append(ready)
            ok = True
            for k0 in range(0, len(busy), 2):
                grp = list(zip(busy[k0:k0 + 2], ready_all[k0:k0 + 2]))
                sets = (_lib.ApplySet * len(grp))()
                for q, (ts, (plan, arena, bias)) in zip(sets, grp):
                    q.attrs, q.dout, q.dbias = ts.desc_ptr(0), arena.data_ptr(), ptr(bias)
                    q.dout_stride, q.plan, q.n_attr = arena.stride(0), plan.c, ts.n_attr
                _lib.tag = '+'.join(ts.prefix for ts, _ in grp)
                if call('arx_pool_bwd_apply_many', ctypes.addressof(sets), len(grp), self.dim, float(lr), ptr(grad_scale),
                        opt) != 0:
                    ok = False
                    break
            if ok:
                for ts in busy:
                    ts.pending = []
                self._after_apply()
                return
            for ts, ready in zip(busy, ready_all):           # not supported: per-set launches below
                ts._ready = ready
        forks = []
        for k, ts in enumerate(busy):
            side = self.side_stream(k) if (len(busy) > 1 and k > 0 and _lib.timeline is None) else None
            if side is not None:
                side.wait_stream(main)
                forks.append(side)
            with torch.cuda.stream(side if side is not None else main):
                ready = getattr(ts, '_ready', None)
                _lib.tag = ts.prefix
                if ready is None and self._apply_catalog_slabs(ts, lr, grad_scale, opt):
                    ts.pending = []                      # gradient of the whole pooled catalog: column-slab passes
                    continue
                if ready is not None:
                    plan, arena, bias = ready
                    ts._ready = None
                else:
                    plan = self._plan_for(ts, ts.pending)
                    arena, bias = self._arena(ts.pending)
                call('arx_pool_bwd_apply', ts.desc_ptr(0), ts.n_attr, self.dim, plan.c, arena.data_ptr(),
                     arena.stride(0), ptr(bias), float(lr), ptr(grad_scale), opt, None, None)
            ts.pending = []
        for side in forks:
            main.wait_stream(side)
        self._after_apply()

    CATALOG_SLAB_HEAVY = 512

    def _apply_catalog_slabs(self, ts, lr, grad_scale, opt):
        """The gradient of the WHOLE pooled catalog (loss = ce / full-catalog WMRB: push_grad(..., plan_key='catalog') as
        the only pending lookup of its table set) in column slabs of 16 floats (arx_pool_bwd_apply_slab): the slab-major
        copy of dP is L2-resident, so the ~100 gathers per table row hit L2 instead of DRAM.  The id table (one
        contribution per row: pure streaming) keeps the row-at-a-time kernel on its own static plan.  Returns False when
        the case is not this one (the caller then runs the general kernel)."""
        if len(ts.pending) != 1 or os.environ.get('ARX_CATALOG_SLABS', '1') != '1':
            return False
        a0, na, ids, mode, dout, dbias, key = ts.pending[0]
        if (key != 'catalog' or mode != POOL_MEAN or self.dim % 16 != 0 or self.shard is not None or na < 2 or a0 != 0
                or ts.n_cat != 1 or opt not in (OPT_ADAGRAD, OPT_SGD) or not dout.is_contiguous()
                or dout.numel() * 4 < (32 << 20) or self.deterministic):
            return False
        V = ids.numel()
        plans = getattr(ts, '_slab_plans', None)
        if plans is None:
            # static plans: id table | attribute tables; the mean over the F = na attributes of the LOOKUP stays in the
            # weights (the sub-plans were filled with 1 / their own attribute count)
            p_id = self._build_plan(ts, [(0, 1, ids, mode)], 'catalog_id')
            p_at = self._build_plan(ts, [(1, na - 1, ids, mode)], 'catalog_attr', heavy=self.CATALOG_SLAB_HEAVY)
            p_id.bucket_w[:int(p_id.counters[1].item())] *= 1.0 / na
            p_at.bucket_w[:int(p_at.counters[1].item())] *= (na - 1.0) / na
            plans = ts._slab_plans = (p_id, p_at)
        p_id, p_at = plans
        lr = float(lr)
        call('arx_pool_bwd_apply', ts.desc_ptr(0), ts.n_attr, self.dim, p_id.c, dout.data_ptr(), dout.stride(0), ptr(dbias),
             lr, ptr(grad_scale), opt, None, None)
        ns = self.dim // 16
        slabs = dout.view(V, ns, 16).permute(1, 0, 2).contiguous()                 # [ns][V][16]: 64 B per item and slab
        for s_ in range(ns):
            call('arx_pool_bwd_apply_slab', ts.desc_ptr(0), ts.n_attr, self.dim, p_at.c, slabs[s_].data_ptr(), 16 * s_,
                 ptr(dbias), lr, ptr(grad_scale), opt, self.CATALOG_SLAB_HEAVY)
        return True

    def _after_apply(self):
        # A plan that ran out of capacity makes plan_fill / apply return without touching anything: surface it instead
        # of training on silently.  Reading the flag synchronises, so: every 256th eager call, never during capture.
        self._apply_calls = getattr(self, '_apply_calls', 0) + 1
        if self._apply_calls % 256 == 1 and not torch.cuda.is_current_stream_capturing():
            self.check_plans()

    def check_plans(self):
        """Raise ARX_E_CAPACITY if any backward plan of this model overflowed its buffers (counters[2])."""
        for ts in self.sets.values():
            plans = list(ts.plans.values()) + ([ts._scratch] if getattr(ts, '_scratch', None) is not None else [])
            for p in plans:
                if int(p.counters[2].item()) != 0:
                    for n in ts.names:
                        self.touch[n].zero_()              # plan_reset only cleared the rows it had recorded
                    p.counters.zero_()
                    raise RuntimeError('backward plan of table set %r exceeded its capacity (ARX_E_CAPACITY): '
                                       'cap_rows=%d cap_occ=%d' % (ts.prefix, p.cap_rows, p.cap_occ))

    def side_stream(self, k):
        """Streams 8.. carry the backward-plan builds: thousands of small latency-bound CTAs that would
        otherwise take the SM slots the dependent chain of the step is waiting for (torch.profiler timeline,
        tools/trace_step.py) — they get the LOW priority, every other stream of the step the high one."""
        if not hasattr(self, '_side_streams'):
            self._side_streams = {}
        if k not in self._side_streams:
            # (single-GPU step only: the row-sharded step keeps default priorities — its NCCL exchanges are part
            # of the captured graph and were validated at N = 2 / 4 with this scheduling)
            # ... and the peer-memory step, whose graph holds no NCCL kernel: hmf/sharded.py sets use_priorities)
            prio = (0 if k >= 8 else -1) if (self.shard is None or getattr(self, 'use_priorities', False)) else 0
            self._side_streams[k] = torch.cuda.Stream(device=self.device, priority=prio)
        return self._side_streams[k]

    def pool_many(self, requests):
        """Several independent pooling launches on parallel streams.  requests: list of
        (prefix, ids, mode, want_bias, kwargs); outputs are allocated on the calling stream first.
        Returns the list of pool() results."""
        main = torch.cuda.current_stream()
        outs = []
        for prefix, ids, mode, want_bias, kw in requests:
            a0, na = self.sets[prefix].attr_range(kw.get('no_id', False), kw.get('no_attribute', False))
            width = self.dim if mode == POOL_MEAN else self.dim * na
            o = kw.get('out')              # optional: pool straight into a (row-strided) view of a caller's buffer
            if kw.get('push') is not None:
                outs.append((None, None))  # results are ADDED into the owner ranks' receive blocks (hmf/exchange.py::PeerExchange)
                continue
            if o is None:
                o = torch.empty((ids.numel(), width), dtype=torch.float32, device=self.device)
            outs.append((o, torch.empty((ids.numel(),), dtype=torch.float32, device=self.device) if want_bias else None))
        pushing = any(r[4].get('push') is not None for r in requests)
        if pushing and not (len(requests) <= 4 and self.dim in (128, 256) and all(r[2] == POOL_MEAN for r in requests)):
            raise NotImplementedError('push-mode lookups need mean pooling and dim 128 / 256')
        if pushing or (len(requests) > 1 and len(requests) <= 4 and self.dim in (128, 256)
                and all(r[2] == POOL_MEAN for r in requests) and os.environ.get('ARX_POOL_MANY', '1') == '1'):
            # all lookups of the step in ONE launch (arx_pool_fwd_many): one ramp, one tail, balanced waves
            reqs = (_lib.PoolReq * len(requests))()
            push = (_lib.PoolPush * len(requests))()
            rngs = []
            for k, ((prefix, ids, mode, want_bias, kw), (o, b)) in enumerate(zip(requests, outs)):
                ts = self.sets[prefix]
                a0, na = ts.attr_range(kw.get('no_id', False), kw.get('no_attribute', False))
                rngs.append((a0, na))
                q = reqs[k]
                q.attrs, q.ent_ids = ts.desc_ptr(a0), ids.data_ptr()
                q.n, q.n_attr = ids.numel(), na
                q.max_rows_per_entity = int(sum(ts.max_len[a0:a0 + na]))
                if kw.get('push') is not None:
                    peers, peers_bias, rows_per_rank, pitch, n_ranks = kw['push']
                    push[k].peer_out, push[k].rows_per_rank, push[k].stride = peers.data_ptr(), rows_per_rank, pitch
                    push[k].peer_bias = peers_bias.data_ptr() if (want_bias and peers_bias is not None) else None
                    push[k].n_ranks = n_ranks
                    q.out, q.bias_out, q.out_stride = None, None, pitch
                else:
                    q.out, q.out_stride = o.data_ptr(), o.stride(0)
                    q.bias_out = b.data_ptr() if want_bias else None
            _lib.tag = 'many'
            if pushing:
                call('arx_pool_fwd_many_push', ctypes.addressof(reqs), ctypes.addressof(push), len(requests), self.dim)
                return [(o, b, rg) for (o, b), rg in zip(outs, rngs)]
            if call('arx_pool_fwd_many', ctypes.addressof(reqs), len(requests), self.dim) == 0:
                return [(o, b, rg) for (o, b), rg in zip(outs, rngs)]
        res, forks = [], []
        for k, ((prefix, ids, mode, want_bias, kw), (o, b)) in enumerate(zip(requests, outs)):
            side = self.side_stream(k) if (k > 0 and _lib.timeline is None) else None
            if side is not None:
                side.wait_stream(main)
                forks.append(side)
            with torch.cuda.stream(side if side is not None else main):
                res.append(self.pool(prefix, ids, mode, want_bias, out=o, bias_out=b,
                                     **{k_: v_ for k_, v_ in kw.items() if k_ != 'out'}))
        for side in forks:
            main.wait_stream(side)
        return res

    def row_gradients(self, prefix):
        """Oracle check: dense gradients of every table of `prefix` from the pending lookups,
        through the same plan + segment-reduce kernels (ARX_OPT_NONE).  Clears the pending list."""
        ts = self.sets[prefix]
        plan = self._plan_for(ts, ts.pending)
        arena, bias = self._arena(ts.pending)
        nu = int(plan.counters[0].item())
        assert int(plan.counters[2].item()) == 0, 'backward plan capacity exceeded'
        rows = torch.zeros((max(nu, 1), self.dim), dtype=torch.float32, device=self.device)
        brow = torch.zeros((max(nu, 1),), dtype=torch.float32, device=self.device)
        call('arx_pool_bwd_apply', ts.desc_ptr(0), ts.n_attr, self.dim, plan.c, arena.data_ptr(),
             arena.stride(0), ptr(bias), 0.0, None, OPT_NONE, rows.data_ptr(), brow.data_ptr())
        grads = {n: torch.zeros_like(self.params[n]) for n in ts.names}
        bgrads = {n: torch.zeros_like(self.params[n]) for n in ts.bias_names if n}
        tok = plan.uniq_tok[:nu].long()
        attr = plan.uniq_attr[:nu].long()
        for f, name in enumerate(ts.names):
            sel = attr == f
            lt = tok[sel] // self.shard[0] if self.shard is not None else tok[sel]
            grads[name][lt] = rows[:nu][sel]
            if ts.bias_names[f] and bias is not None:
                bgrads[ts.bias_names[f]][lt, 0] = brow[:nu][sel]
        ts.pending = []
        grads.update(bgrads)
        return grads
