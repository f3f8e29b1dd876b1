#!/usr/bin/env python
"""Generate golden vectors by running the UNMODIFIED reference sources.

    python tests/golden/make_ref_golden.py            # needs /root/reference (this container only)

The reference (Python 2 + TensorFlow 1.0) cannot run as published: neither interpreter nor library
is installable offline.  Its model code however parses under Python 3, so this script imports
`/root/reference/hmf/hmf_model.py`, `attributes/embed_attribute.py`, `attributes/mulhot_index.py`
and `attributes/attribute.py` AS THEY LIE (nothing copied, nothing edited) on top of
`oracle/tf1_shim/tensorflow` — a small lazy-graph re-statement of the TF-1.0 ops those files call —
and records what the reference's own graph computes: per-step training losses, the parameters after
the last step, the eval loss and the top-k recommendation.  The fixtures land in
`tests/golden/ref_hmf_<case>.npz`; `tests/test_ref_golden.py` holds the oracle (CPU) and the CUDA
path (GPU) to them.

Honest scope of the pin: op ORDER and model wiring are the reference's; per-op numerics are the
shim's reading of the TF-1.0 documentation (see the shim's docstring).
"""
import builtins
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get('ARX_REFERENCE', '/root/reference')
OUT = os.environ.get('ARX_GOLDEN_OUT', HERE)          # where the .npz fixtures are written (default: next to this script)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
sys.path.insert(0, os.path.join(ROOT, 'oracle', 'tf1_shim'))
for sub in ('attributes', 'hmf', 'lstm', 'word2vec', 'utils'):
    sys.path.append(os.path.join(REF, sub))
builtins.xrange = range          # python-2 builtin used without six in places

import tensorflow as tf  # noqa: E402  (the shim)
assert 'tf1_shim' in tf.__file__

import attribute as ref_attribute  # noqa: E402  /root/reference/attributes/attribute.py
from helpers import small_dataset, random_params, positives  # noqa: E402


def to_ref_attributes(a, dim):
    """our Attributes (same field names) -> the reference's own container class, python lists of
    ints exactly as utils/preprocess.py builds them."""
    r = ref_attribute.Attributes(a.num_features_cat, [x.tolist() for x in a.features_cat],
                                 a.num_features_mulhot, [x.tolist() for x in a.features_mulhot], None,
                                 [x.tolist() for x in a.mulhot_starts], [x.tolist() for x in a.mulhot_lengths],
                                 list(a._embedding_classes_list_cat), list(a._embedding_classes_list_mulhot))
    r.set_model_size(dim)
    if hasattr(a, 'full_cat_tr'):
        r.set_target_prediction([x.tolist() for x in a.full_cat_tr], [x.tolist() for x in a.full_values_tr],
                                [x.tolist() for x in a.full_segids_tr],
                                [np.asarray(x, dtype=np.float32).reshape(-1, 1).tolist() for x in a.full_lengths_tr])
    return r


def pack_attributes(prefix, a, out):
    out[prefix + 'n_cat'] = a.num_features_cat
    out[prefix + 'n_mulhot'] = a.num_features_mulhot
    out[prefix + 'v_cat'] = np.asarray(a._embedding_classes_list_cat, dtype=np.int64)
    out[prefix + 'v_mulhot'] = np.asarray(a._embedding_classes_list_mulhot, dtype=np.int64)
    for i in range(a.num_features_cat):
        out['%scat_%d' % (prefix, i)] = a.features_cat[i]
    for i in range(a.num_features_mulhot):
        out['%svalues_%d' % (prefix, i)] = a.features_mulhot[i]
        out['%sstarts_%d' % (prefix, i)] = a.mulhot_starts[i]
        out['%slengths_%d' % (prefix, i)] = a.mulhot_lengths[i]


class MaskQueue(object):
    """dropout hook: draws floor(keep + U) from a seeded stream and remembers every mask in call order."""

    def __init__(self, seed):
        self.rng = np.random.default_rng(seed)
        self.taken = []
        self.by_node = []      # (graph creation index of the dropout node, mask)

    def __call__(self, shape, keep, node=None):
        m = np.floor(self.rng.random(shape) + keep).astype(np.float32)
        self.taken.append(m)
        self.by_node.append((node.seq if node is not None else len(self.by_node), m))
        return m


HMF_CASES = {
    # name: loss, nonlinear, dim, n_sampled, loss_func, exp_p
    'ce_linear': ('ce', 'linear', 8, None, 'log', 1.005),
    'ce_relu': ('ce', 'relu', 8, None, 'log', 1.005),
    'warp_linear': ('warp', 'linear', 8, None, 'log', 1.005),
    'warp_tanh': ('warp', 'tanh', 8, None, 'log', 1.005),
    'rs_log': ('rs', 'linear', 8, None, 'log', 1.005),
    'rs_poly2': ('rs', 'linear', 8, None, 'poly2', 1.3),
    'rs_sig_exp': ('rs-sig', 'linear', 8, None, 'exp', 1.2),
    'rs_sig2_square': ('rs-sig2', 'linear', 8, None, 'square', 1.005),
    'bbpr': ('bbpr', 'linear', 20, None, 'log', 1.005),
    'mw_linear': ('mw', 'linear', 8, 12, 'log', 1.005),
    'mw_tanh': ('mw', 'tanh', 8, 12, 'log', 1.005),
    'mw_dim128': ('mw', 'linear', 128, 12, 'log', 1.005),
    'ce_dim32_logit40': ('ce', 'linear', 32, None, 'log', 1.005),
}


def run_hmf_case(name, n_steps=6):
    import hmf_model as ref_hmf          # /root/reference/hmf/hmf_model.py
    loss, nonlinear, dim, ns, loss_func, exp_p = HMF_CASES[name]
    n_users, n_items, mb, hidden, lr, keep, topn = 60, 50, 16, 5, 0.3, 0.5, 10
    logit_size = 40 if 'logit40' in name else None
    ua, ia, i2l, l2i = small_dataset(n_users, n_items, 3, 25, 3, 7, 0, logit_size, dim)
    params = random_params(ua, ia, dim, 1, mlp_hidden=hidden if nonlinear != 'linear' else None)
    V = len(l2i)
    l2i_d = {int(v): int(l2i[v]) for v in range(V)}
    i2l_d = {int(l2i[v]): int(v) for v in range(V)}

    tf.reset_default_graph()
    rua, ria = to_ref_attributes(ua, dim), to_ref_attributes(ia, dim)
    model = ref_hmf.LatentProductModel(n_users, n_items, dim, 1, mb, lr, 1.0, rua, ria, i2l_d, l2i_d,
                                       loss_function=loss, nonlinear=None if nonlinear == 'linear' else nonlinear,
                                       dropout=keep, n_sampled=ns, top_N_items=topn, hidden_size=hidden,
                                       loss_func=loss_func, loss_exp_p=exp_p)
    g = tf.get_default_graph()
    for k, v in params.items():          # inject the weights by the reference's own variable names
        g.by_name[k].load(v)
    trainable = sorted(v._name for v in tf.trainable_variables())
    assert trainable == sorted(params.keys()), (trainable, sorted(params.keys()))
    masks = MaskQueue(99)
    tf.set_dropout_hook(masks)
    sess = tf.Session()

    out = {'case': name, 'loss': loss, 'nonlinear': nonlinear, 'dim': dim, 'n_sampled': -1 if ns is None else ns,
           'loss_func': loss_func, 'exp_p': exp_p, 'n_users': n_users, 'n_items': n_items, 'mb': mb,
           'hidden': hidden, 'lr': lr, 'keep_prob': keep, 'top_n': topn, 'n_steps': n_steps,
           'l2i': np.asarray(l2i, dtype=np.int64)}
    pack_attributes('u_', ua, out)
    pack_attributes('i_', ia, out)
    for k, v in params.items():
        out['init/' + k] = v
    rng = np.random.default_rng(17)
    losses = []
    id2idx = None
    for it in range(n_steps):
        users = rng.integers(0, n_users, mb)
        items = rng.integers(0, V, mb)          # targets must be catalog items
        items = np.asarray([l2i_d[int(v)] for v in items])
        if it == 3:
            users[:4] = users[4]
            items[:3] = items[5]                  # duplicates inside the batch
        pos = positives(users, items, n_users, rng, n_items=V)
        pos = {u: [l2i_d[v] if v in l2i_d else v for v in vs] for u, vs in pos.items()}
        model.prepare_warp(pos, pos)
        sampled = None
        if ns and it % 2 == 0:
            sampled = [l2i_d[int(v)] for v in rng.permutation(V)[:ns]]
            id2idx = {v: k for k, v in enumerate(sampled)}
        n_before = len(masks.taken)
        lval = model.step(sess, [int(u) for u in users], [int(i) for i in items], None, sampled, id2idx,
                          loss=loss)
        losses.append(float(lval))
        out['step%d/users' % it] = users.astype(np.int64)
        out['step%d/items' % it] = items.astype(np.int64)
        out['step%d/sampled' % it] = np.asarray(sampled if sampled is not None else [], dtype=np.int64)
        pu = sorted(pos.keys())
        out['step%d/pos_users' % it] = np.asarray(pu, dtype=np.int64)
        out['step%d/pos_ptr' % it] = np.cumsum([0] + [len(pos[u]) for u in pu]).astype(np.int64)
        out['step%d/pos_items' % it] = np.asarray([v for u in pu for v in pos[u]], dtype=np.int64)
        taken = masks.taken[n_before:]
        out['step%d/n_masks' % it] = len(taken)
        for j, m in enumerate(taken):
            out['step%d/mask%d' % (it, j)] = m
    out['losses'] = np.asarray(losses, dtype=np.float64)
    for v in tf.trainable_variables():
        out['final/' + v._name] = v.numpy()
    # eval loss (forward_only -> keep_prob fed as 1.0, eval positives) and recommendation
    users = rng.integers(0, n_users, mb)
    items = np.asarray([l2i_d[int(v)] for v in rng.integers(0, V, mb)])
    pos = positives(users, items, n_users, rng, n_items=V)
    model.prepare_warp(pos, pos)
    ev = model.step(sess, [int(u) for u in users], [int(i) for i in items], None, None, id2idx,
                    forward_only=True, loss=loss)
    out['eval/users'] = users.astype(np.int64)
    out['eval/items'] = items.astype(np.int64)
    pu = sorted(pos.keys())
    out['eval/pos_users'] = np.asarray(pu, dtype=np.int64)
    out['eval/pos_ptr'] = np.cumsum([0] + [len(pos[u]) for u in pu]).astype(np.int64)
    out['eval/pos_items'] = np.asarray([v for u in pu for v in pos[u]], dtype=np.int64)
    out['eval/loss'] = float(ev)
    rec = model.step(sess, list(range(mb)), None, forward_only=True, recommend=True)
    out['recommend/users'] = np.arange(mb, dtype=np.int64)
    out['recommend/indices'] = np.asarray(rec, dtype=np.int64)
    out['global_step'] = int(model.global_step.numpy())
    tf.set_dropout_hook(None)
    path = os.path.join(OUT, 'ref_hmf_%s.npz' % name)
    np.savez_compressed(path, **out)
    print('%-18s losses %s  eval %.6f  -> %s' % (name, np.round(losses, 5).tolist(), ev, os.path.basename(path)))


def run_hmf_warp_eval():
    """loss_function='warp_eval' (hmf_model.py:123-124, embed_attribute.py:620-639): the per-row margin rank and
    true rank the reference reports at evaluation time (keep_prob 1, dev positives)."""
    import hmf_model as ref_hmf
    dim, n_users, n_items, mb = 8, 60, 50, 16
    ua, ia, i2l, l2i = small_dataset(n_users, n_items, 3, 25, 3, 7, 0, None, dim)
    params = random_params(ua, ia, dim, 1)
    V = len(l2i)
    l2i_d = {int(v): int(l2i[v]) for v in range(V)}
    i2l_d = {int(l2i[v]): int(v) for v in range(V)}
    tf.reset_default_graph()
    model = ref_hmf.LatentProductModel(n_users, n_items, dim, 1, mb, 0.3, 1.0, to_ref_attributes(ua, dim),
                                       to_ref_attributes(ia, dim), i2l_d, l2i_d, loss_function='warp_eval',
                                       dropout=0.5, top_N_items=10)
    g = tf.get_default_graph()
    for k, v in params.items():
        g.by_name[k].load(v)
    sess = tf.Session()
    rng = np.random.default_rng(23)
    users = rng.integers(0, n_users, mb)
    items = rng.integers(0, V, mb)
    users[:3] = users[3]
    pos = positives(users, items, n_users, rng, n_items=V)
    model.prepare_warp({}, pos)                        # forward_only reads pos_item_set_eval
    margin, rank = model.step(sess, [int(u) for u in users], [int(i) for i in items], forward_only=True,
                              loss='warp_eval')
    out = {'dim': dim, 'n_users': n_users, 'n_items': n_items, 'mb': mb, 'l2i': np.asarray(l2i, dtype=np.int64),
           'users': users.astype(np.int64), 'items': items.astype(np.int64),
           'margin_rank': np.asarray(margin, dtype=np.float64), 'true_rank': np.asarray(rank, dtype=np.int64)}
    pu = sorted(pos.keys())
    out['pos_users'] = np.asarray(pu, dtype=np.int64)
    out['pos_ptr'] = np.cumsum([0] + [len(pos[u]) for u in pu]).astype(np.int64)
    out['pos_items'] = np.asarray([v for u in pu for v in pos[u]], dtype=np.int64)
    pack_attributes('u_', ua, out)
    pack_attributes('i_', ia, out)
    for k, v in params.items():
        out['init/' + k] = v
    path = os.path.join(OUT, 'ref_hmfeval_warp_eval.npz')
    np.savez_compressed(path, **out)
    print('hmf warp_eval margin %s rank %s -> %s' % (np.round(margin[:4], 3).tolist(), np.asarray(rank)[:8].tolist(),
                                                     os.path.basename(path)))


LSTM_CASES = {
    # name: loss, use_concat, use_sep_item, withAdagrad, lr, max_gradient_norm   (lstm/run.py always passes
    # no_user_id=False)
    # loss 'mw' is not pinned: lstm/seqModel.py:301 never hands item_sampled to add_input, so the
    # reference's sampled-pool Variables are never written on this path (it scores garbage there).
    'ce_mean': ('ce', False, False, True, 0.5, 5.0),
    'ce_concat': ('ce', True, False, True, 0.5, 5.0),
    'ce_sep_item': ('ce', False, True, True, 0.5, 5.0),
    'ce_concat_sep': ('ce', True, True, True, 0.5, 5.0),
    'warp_mean': ('warp', False, False, True, 0.5, 5.0),
    'ce_sgd_clipped': ('ce', False, False, False, 2.0, 0.5),        # norm > max_gradient_norm: clipping active
    'ce_adagrad_clipped': ('ce', False, True, True, 0.5, 0.25),
}
# non-default flags, fixtures named ref_lstmx_*: (loss, use_concat, sep, adagrad, lr, clip, extra kwargs)
LSTMX_CASES = {
    'ce_outfeat0': ('ce', False, False, True, 0.5, 5.0, {'output_feat': 0}),
    'ce_outfeat0_sep': ('ce', True, True, True, 0.5, 5.0, {'output_feat': 0}),
    'ce_noitemfeat': ('ce', False, False, True, 0.5, 5.0, {'no_input_item_feature': True}),
    # (use_concat + no_input_item_feature is not usable in the reference: w_input_item keeps the full concat width,
    #  seqModel.py:135-137 vs embed_attribute.py:368-369)
    'warp_noitemfeat': ('warp', False, False, True, 0.5, 5.0, {'no_input_item_feature': True}),
    # MultiRNNCell([DropoutWrapper(LSTMCell, in)] * 2) + DropoutWrapper(out) (seqModel.py:99-103)
    'ce_2layers': ('ce', False, False, True, 0.5, 5.0, {'num_layers': 2}),
    # non-linear attribute pooling of the token scores (embed_attribute.py:194-200): segment max / log-sum-exp
    'ce_outfeat2': ('ce', False, False, True, 0.5, 5.0, {'output_feat': 2}),
    'ce_outfeat3_sep': ('ce', False, True, True, 0.5, 5.0, {'output_feat': 3}),
    'warp_outfeat2_sep': ('warp', True, True, True, 0.5, 5.0, {'output_feat': 2}),
}


def run_lstm_case(name, n_steps=4, extended=False):
    import seqModel as ref_seq            # /root/reference/lstm/seqModel.py
    import embed_attribute as ref_emb     # /root/reference/attributes/embed_attribute.py
    extra = {}
    if extended:
        loss, use_concat, sep, adagrad, lr, clip, extra = LSTMX_CASES[name]
        extra = dict(extra)
    else:
        loss, use_concat, sep, adagrad, lr, clip = LSTM_CASES[name]
    n_users, n_items, dim, mb, T, keep, topk = 40, 30, 8, 12, 5, 0.5, 5
    buckets = [3, T]
    ua, ia, _, l2i = small_dataset(n_users, n_items, 2, 15, 3, 5, 0, None, dim)
    params = random_params(ua, ia, dim, 1, scale=0.4, item_output=sep)
    rng = np.random.default_rng(5)
    Fu = ua.num_features_cat + ua.num_features_mulhot
    Fi = ia.num_features_cat + ia.num_features_mulhot
    if extra.get('no_input_item_feature'):
        Fi = 1                                          # only the id embedding feeds the LSTM (embed_attribute.py:368-369)
    if use_concat:
        params['w_input_user'] = rng.uniform(-.4, .4, (Fu * dim, dim)).astype(np.float32)
        params['w_input_item'] = rng.uniform(-.4, .4, (Fi * dim, dim)).astype(np.float32)
    params['lstm_w'] = rng.uniform(-.4, .4, (2 * dim, 4 * dim)).astype(np.float32)
    params['lstm_b'] = np.zeros(4 * dim, dtype=np.float32)
    n_layers = int(extra.pop('num_layers', 1))
    for l in range(1, n_layers):
        params['lstm_w_%d' % l] = rng.uniform(-.4, .4, (2 * dim, 4 * dim)).astype(np.float32)
        params['lstm_b_%d' % l] = np.zeros(4 * dim, dtype=np.float32)
    START = n_items
    l2i_d = {int(v): int(l2i[v]) for v in range(len(l2i))}
    i2l_d = {v: k for k, v in l2i_d.items()}
    i2l_d[START] = 0                                  # lstm/run.py:276-278

    tf.reset_default_graph()
    devices = ['/cpu:0'] * 3
    rua, ria = to_ref_attributes(ua, dim), to_ref_attributes(ia, dim)
    emb = ref_emb.EmbeddingAttribute(rua, ria, mb, None, buckets[-1], sep, i2l_d, l2i_d, devices=devices)
    model = ref_seq.SeqModel(buckets, dim, n_layers, clip, mb, lr, 0.83, emb, withAdagrad=adagrad, dropoutRate=keep,
                             START_ID=START, loss=loss, devices=devices, use_concat=use_concat, no_user_id=False,
                             topk_n=topk, **extra)
    g = tf.get_default_graph()
    tfname = {'lstm_w': 'rnn/multi_rnn_cell/cell_0/lstm_cell/weights',
              'lstm_b': 'rnn/multi_rnn_cell/cell_0/lstm_cell/biases'}
    for l in range(1, n_layers):
        tfname['lstm_w_%d' % l] = 'rnn/multi_rnn_cell/cell_%d/lstm_cell/weights' % l
        tfname['lstm_b_%d' % l] = 'rnn/multi_rnn_cell/cell_%d/lstm_cell/biases' % l
    for k, v in params.items():
        g.by_name[tfname.get(k, k)].load(v)
    back = {v: k for k, v in tfname.items()}
    trainable = sorted(back.get(v._name, v._name) for v in tf.trainable_variables())
    assert trainable == sorted(params.keys()), (trainable, sorted(params.keys()))
    out_extra = {'output_feat': int(extra.get('output_feat', 1)),
                 'no_input_item_feature': bool(extra.get('no_input_item_feature', False))}
    if n_layers > 1:
        out_extra['num_layers'] = n_layers
    masks = MaskQueue(7)
    tf.set_dropout_hook(masks)
    sess = tf.Session()

    out = {'case': name, 'loss': loss, 'use_concat': use_concat, 'sep': sep, 'adagrad': adagrad, 'dim': dim,
           'n_users': n_users, 'n_items': n_items, 'mb': mb, 'T': T, 'buckets': np.asarray(buckets), 'lr': lr,
           'keep_prob': keep, 'topk': topk, 'n_steps': n_steps, 'START': START, 'clip': clip,
           'l2i': np.asarray(l2i, dtype=np.int64)}
    out.update(out_extra)
    pack_attributes('u_', ua, out)
    pack_attributes('i_', ia, out)
    for k, v in params.items():
        out['init/' + k] = v

    def batch(Tb):
        users = rng.integers(0, n_users, mb).tolist()
        seqs = [rng.integers(0, n_items, rng.integers(1, Tb + 1)).tolist() for _ in range(mb)]
        inp = [[START] + s_[:-1] + [START] * (Tb - len(s_)) for s_ in seqs]
        tgt = [s_ + [START] * (Tb - len(s_)) for s_ in seqs]
        w = [[1.0] * len(s_) + [0.0] * (Tb - len(s_)) for s_ in seqs]
        tm = lambda l: [[l[j][i] for j in range(mb)] for i in range(Tb)]   # noqa: E731  time-major
        return users, tm(inp), tm(tgt), tm(w), seqs

    def put(tag, users, inp, tgt, w, pos):
        out[tag + '/users'] = np.asarray(users, dtype=np.int64)
        out[tag + '/inputs'] = np.asarray(inp, dtype=np.int64)
        out[tag + '/targets'] = np.asarray(tgt, dtype=np.int64)
        out[tag + '/weights'] = np.asarray(w, dtype=np.float32)
        pu = sorted(pos.keys())
        out[tag + '/pos_users'] = np.asarray(pu, dtype=np.int64)
        out[tag + '/pos_ptr'] = np.cumsum([0] + [len(pos[u]) for u in pu]).astype(np.int64)
        out[tag + '/pos_items'] = np.asarray([v for u in pu for v in pos[u]], dtype=np.int64)

    losses, norms = [], []
    for it in range(n_steps):
        b = 0 if it == 2 else 1
        Tb = buckets[b]
        users, inp, tgt, w, seqs = batch(Tb)
        pos = {u: sorted(set(s_)) for u, s_ in zip(users, seqs)}
        emb.prepare_warp(pos, pos)
        masks.by_node = []
        lval = model.step(sess, users, inp, tgt, w, b)
        fetched = [f for f in sess.fetch_log if isinstance(f, list) and len(f) == 3][-1]   # [loss, update, norm]
        losses.append(float(lval))
        norms.append(float(fetched[2]))
        order = sorted(masks.by_node, key=lambda kv: kv[0])     # graph order per time step: in (layer 0), [in (layer 1), ...], out
        per = n_layers + 1
        assert len(order) == per * Tb, len(order)
        put('step%d' % it, users, inp, tgt, w, pos)
        out['step%d/bucket' % it] = b
        out['step%d/in_masks' % it] = np.stack([m for _, m in order[0::per]])
        for l in range(1, n_layers):
            out['step%d/in_masks_%d' % (it, l)] = np.stack([m for _, m in order[l::per]])
        out['step%d/out_masks' % it] = np.stack([m for _, m in order[n_layers::per]])
    out['losses'] = np.asarray(losses, dtype=np.float64)
    out['gnorms'] = np.asarray(norms, dtype=np.float64)
    for v in tf.trainable_variables():
        out['final/' + back.get(v._name, v._name)] = v.numpy()
    # evaluation as lstm/run.py does it: dropout10_op, forward_only step on losses_full
    users, inp, tgt, w, seqs = batch(T)
    pos = {u: sorted(set(s_)) for u, s_ in zip(users, seqs)}
    emb.prepare_warp(pos, pos)
    sess.run(model.dropout10_op)
    ev = model.step(sess, users, inp, tgt, w, 1, forward_only=True)
    put('eval', users, inp, tgt, w, pos)
    out['eval/loss'] = float(ev)
    # per-position top-k of softmax(full logits) (seqModel.py:514-519), same feeds
    feed = {}
    for l in range(T):
        feed[model.targets[l].name] = emb.target_mapping(tgt)[l]
        feed[model.target_weights[l].name] = w[l]
    emb.add_input(feed, users, inp, forward_only=True, recommend=True, loss=loss)
    tk = sess.run(model.topk_indexes[1], feed)
    out['eval/topk_indexes'] = np.asarray(tk, dtype=np.int64)      # [T, mb, topk]
    sess.run(model.dropoutAssign_op)
    out['global_step'] = int(model.global_step.numpy())
    tf.set_dropout_hook(None)
    path = os.path.join(OUT, 'ref_%s_%s.npz' % ('lstmx' if extended else 'lstm', name))
    np.savez_compressed(path, **out)
    print('lstm %-14s losses %s gnorm %s eval %.5f -> %s' % (name, np.round(losses, 4).tolist(),
                                                          np.round(norms, 3).tolist(), ev, os.path.basename(path)))


CBOW_CASES = {
    # name: loss, use_sep_item, n_input_items.  loss 'mw' is not pinned: word2vec/cbow_model.py:126 references
    # batch_loss_test, which the 'mw' branch (:118-120) never defines -> the reference raises NameError.
    'ce_sep_ni2': ('ce', True, 2),
    'warp_sep_ni3': ('warp', True, 3),
    'bbpr_shared_ni1': ('bbpr', False, 1),
    'ce_shared_ni0': ('ce', False, 0),
    # skip-gram tower (word2vec/skipgram_model.py: trains on the first input item only), fixtures ref_cbow_sg_*
    'sg_ce_sep_ni2': ('ce', True, 2),
    'sg_warp_shared_ni1': ('warp', False, 1),
}


def run_cbow_case(name, n_steps=4):
    if name.startswith('sg_'):
        import skipgram_model as ref_cbow     # /root/reference/word2vec/skipgram_model.py
    else:
        import cbow_model as ref_cbow         # /root/reference/word2vec/cbow_model.py
    loss, sep, ni = CBOW_CASES[name]
    dim, mb, n_users, n_items, lr, keep, topn = 8, 16, 40, 30, 0.5, 0.5, 5
    ua, ia, _, l2i = small_dataset(n_users, n_items, 2, 15, 3, 5, 0, None, dim)
    params = random_params(ua, ia, dim, 1, scale=0.4, item_output=sep)
    l2i_d = {int(v): int(l2i[v]) for v in range(len(l2i))}
    i2l_d = {v: k for k, v in l2i_d.items()}
    tf.reset_default_graph()
    rua, ria = to_ref_attributes(ua, dim), to_ref_attributes(ia, dim)
    model = ref_cbow.Model(n_users, n_items, dim, mb, lr, 1.0, rua, ria, i2l_d, l2i_d, n_input_items=ni,
                           loss_function=loss, dropout=keep, top_N_items=topn, use_sep_item=sep, n_sampled=None)
    g = tf.get_default_graph()
    for k, v in params.items():
        g.by_name[k].load(v)
    assert sorted(v._name for v in tf.trainable_variables()) == sorted(params.keys())
    masks = MaskQueue(11)
    tf.set_dropout_hook(masks)
    sess = tf.Session()
    out = {'case': name, 'loss': loss, 'sep': sep, 'ni': ni, 'dim': dim, 'mb': mb, 'n_users': n_users,
           'n_items': n_items, 'lr': lr, 'keep_prob': keep, 'top_n': topn, 'n_steps': n_steps,
           'l2i': np.asarray(l2i, dtype=np.int64)}
    pack_attributes('u_', ua, out)
    pack_attributes('i_', ia, out)
    for k, v in params.items():
        out['init/' + k] = v
    rng = np.random.default_rng(3)
    nin = max(ni, 1)
    losses = []

    def put_pos(tag, pos):
        pu = sorted(pos.keys())
        out[tag + '/pos_users'] = np.asarray(pu, dtype=np.int64)
        out[tag + '/pos_ptr'] = np.cumsum([0] + [len(pos[u]) for u in pu]).astype(np.int64)
        out[tag + '/pos_items'] = np.asarray([v for u in pu for v in pos[u]], dtype=np.int64)

    for it in range(n_steps):
        users = rng.integers(0, n_users, mb)
        outs = rng.integers(0, n_items, mb)
        ins = [rng.integers(0, n_items + 1, mb) for _ in range(nin)]      # may include the PAD pseudo-item
        pos = positives(users, outs, n_users, rng, n_items=n_items)
        model.prepare_warp(pos, pos)
        n0 = len(masks.taken)
        lval = model.step(sess, users.tolist(), [x.tolist() for x in ins], outs.tolist(), None, None, loss=loss)
        assert len(masks.taken) == n0 + 1
        losses.append(float(lval))
        tag = 'step%d' % it
        out[tag + '/users'] = users.astype(np.int64)
        out[tag + '/outputs'] = outs.astype(np.int64)
        out[tag + '/inputs'] = np.stack(ins).astype(np.int64)
        out[tag + '/mask'] = masks.taken[-1]
        put_pos(tag, pos)
    out['losses'] = np.asarray(losses, dtype=np.float64)
    for v in tf.trainable_variables():
        out['final/' + v._name] = v.numpy()
    ev = model.step(sess, users.tolist(), [x.tolist() for x in ins], outs.tolist(), forward_only=True, loss=loss)
    out['eval/loss'] = float(ev)                    # loss_test on the last training batch (same positives)
    rec = model.step(sess, users.tolist(), [x.tolist() for x in ins], forward_only=True, recommend=True)
    out['recommend/indices'] = np.asarray(rec, dtype=np.int64)
    tf.set_dropout_hook(None)
    path = os.path.join(OUT, 'ref_cbow_%s.npz' % name)
    np.savez_compressed(path, **out)
    print('cbow %-16s losses %s eval %.5f -> %s' % (name, np.round(losses, 4).tolist(), ev, os.path.basename(path)))


def main():
    assert os.path.isdir(REF), 'reference sources not found at %s' % REF
    names = sys.argv[1:] or (list(HMF_CASES) + ['hmfeval:warp_eval'] + ['lstm:' + n for n in LSTM_CASES] +
                             ['lstmx:' + n for n in LSTMX_CASES] +
                             ['cbow:' + n for n in CBOW_CASES])
    for n in names:
        if n == 'hmfeval:warp_eval':
            run_hmf_warp_eval()
        elif n.startswith('lstmx:'):
            run_lstm_case(n[6:], extended=True)
        elif n.startswith('lstm:'):
            run_lstm_case(n[5:])
        elif n.startswith('cbow:'):
            run_cbow_case(n[5:])
        else:
            run_hmf_case(n)


if __name__ == '__main__':
    main()
