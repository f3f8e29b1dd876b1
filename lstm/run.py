"""run.py — LSTM next-item runner with the reference's flag surface (lstm/run.py:37-117, 49 flags),
data preparation (:170-319: sequence forming, 5 % dev split, bucket search), training loop with
the reference's own throughput line ("Speed: ... targets / sec", :468-470), dev perplexity, best
checkpoint, patience, and --recommend.  `examples/run_lstm.sh` (cd ../lstm; python run.py --flags,
including --steps_per_checkpoint which this runner does not define) runs unchanged.
`--ensemble` / `--beam_search` are dead/broken in the reference (undefined names at
lstm/run.py:704,718-719) and are rejected here.
"""
import logging
import math
import os
import random
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import arecsys_b200  # noqa: E402,F401
from arecsys_b200.utils.flags import Flags  # noqa: E402

FLAGS = Flags()
FLAGS.DEFINE_string("dataset", "xing", "dataset name")
FLAGS.DEFINE_string("raw_data", "../raw_data", "input data directory")
FLAGS.DEFINE_string("data_dir", "./cache0/", "Data directory")
FLAGS.DEFINE_string("train_dir", "./train", "Training directory.")
FLAGS.DEFINE_boolean("test", True, "Test on test splits")
FLAGS.DEFINE_string("combine_att", 'mix', "method to combine attributes: het or mix")
FLAGS.DEFINE_boolean("use_item_feature", True, "RT")
FLAGS.DEFINE_boolean("use_user_feature", True, "RT")
FLAGS.DEFINE_integer("item_vocab_size", 50000, "Item vocabulary size.")
FLAGS.DEFINE_integer("vocab_min_thresh", 2, "filter inactive tokens.")
FLAGS.DEFINE_string("loss", 'ce', "loss function: ce, warp, (mw, mce, bpr)")
FLAGS.DEFINE_float("learning_rate", 0.5, "Learning rate.")
FLAGS.DEFINE_float("learning_rate_decay_factor", 0.83, "Learning rate decays by this much.")
FLAGS.DEFINE_float("max_gradient_norm", 5.0, "Clip gradients to this norm.")
FLAGS.DEFINE_float("keep_prob", 0.5, "dropout rate.")
FLAGS.DEFINE_float("power", 0.5, "related to sampling rate.")
FLAGS.DEFINE_integer("batch_size", 64, "Batch size to use during training/evaluation.")
FLAGS.DEFINE_integer("size", 128, "Size of each model layer.")
FLAGS.DEFINE_integer("num_layers", 1, "Number of layers in the model.")
FLAGS.DEFINE_integer("n_epoch", 500, "Maximum number of epochs in training.")
FLAGS.DEFINE_integer("L", 30, "max length")
FLAGS.DEFINE_integer("n_bucket", 10, "num of buckets to run.")
FLAGS.DEFINE_integer("patience", 10, "exit if the model can't improve for $patence evals")
FLAGS.DEFINE_boolean("recommend", False, "Set to True for recommending.")
FLAGS.DEFINE_boolean("recommend_new", False, "TODO.")
FLAGS.DEFINE_integer("topk", 100, "recommend items with the topk values")
FLAGS.DEFINE_boolean("ensemble", False, "to ensemble")
FLAGS.DEFINE_string("ensemble_suffix", "", "multiple models suffix: 1,2,3,4,5")
FLAGS.DEFINE_integer("seed", 0, "mini batch sampling random seed.")
FLAGS.DEFINE_integer("output_feat", 1, "0: no use, 1: use, mean-mulhot, 2: use, max-pool")
FLAGS.DEFINE_boolean("use_sep_item", False, "use separate embedding parameters for output items.")
FLAGS.DEFINE_boolean("no_input_item_feature", False, "not using attributes at input layer")
FLAGS.DEFINE_boolean("use_concat", False, "use concat or mean")
FLAGS.DEFINE_boolean("no_user_id", True, "use user id or not")
FLAGS.DEFINE_string("N", "000", "GPU layer distribution: [input_embedding, lstm, output_embedding]")
FLAGS.DEFINE_boolean("withAdagrad", True, "withAdagrad.")
FLAGS.DEFINE_boolean("fromScratch", True, "fromScratch.")
FLAGS.DEFINE_boolean("saveCheckpoint", False, "save Model at each checkpoint.")
FLAGS.DEFINE_boolean("profile", False, "False = no profile, True = profile")
FLAGS.DEFINE_integer("ta", 1, "part of target_active")
FLAGS.DEFINE_float("user_sample", 1.0, "user sample rate.")
FLAGS.DEFINE_boolean("after40", False, "whether use items after week 40 only.")
FLAGS.DEFINE_string("split", "last", "last: last maxlen only; overlap: overlap 1 / 3 of maxlen")
FLAGS.DEFINE_integer("n_sampled", 1024, "sampled softmax/warp loss.")
FLAGS.DEFINE_integer("n_resample", 30, "iterations before resample.")
FLAGS.DEFINE_boolean("beam_search", False, "to beam_search")
FLAGS.DEFINE_integer("beam_size", 10, "the beam size")
FLAGS.DEFINE_integer("max_train_data_size", 0, "Limit on the size of training data (0: no limit).")
FLAGS.DEFINE_boolean("old_att", False, "tmp: use attribute_0.8.csv")
FLAGS.DEFINE_integer("max_steps", 0, "not in the reference: stop after this many steps (0 = n_epoch decides)")

_buckets = []


def mylog(msg):
    print(msg)
    sys.stdout.flush()
    logging.info(msg)


def get_buckets_id(l, buckets):
    for i, b in enumerate(buckets):
        if l <= b:
            return i
    return -1


def split_buckets(array, buckets):
    d = [[] for _ in buckets]
    for u, items in array:
        index = get_buckets_id(len(items), buckets)
        if index >= 0:
            d[index].append((u, items))
    return d


def form_sequence_prediction(data, uids, maxlen, START_ID):
    m = {uid: items for uid, items in data}
    return [(uid, [START_ID] + m[uid][-(maxlen - 1):]) if uid in m else (uid, [START_ID]) for uid in uids]


def form_sequence(data, maxlen=100):
    """lstm/run.py:170-209: per-user time-ordered histories, chunked to maxlen (a tail of <= 7
    items is merged with the 10 items before it)."""
    d = {}
    for u, i, week in data:
        d.setdefault(u, []).append((i, week))
    dd = []
    for u in d:
        tmp = sorted(d[u], key=lambda x: x[1])
        while True:
            new_tmp = [x[0] for x in tmp][:maxlen]
            if len(new_tmp) > 0:
                dd.append((u, new_tmp))
            if len(tmp) <= maxlen:
                break
            tmp = tmp[maxlen - 10:] if len(tmp) - maxlen <= 7 else tmp[maxlen:]
    return dd


def prepare_warp(embAttr, data_tr, data_va):
    pos_item_list = {u: list(set(i_list)) for u, i_list in data_tr}
    pos_item_list_val = {u: list(set(i_list)) for u, i_list in data_va}
    embAttr.prepare_warp(pos_item_list, pos_item_list_val)


def split_train_dev(seq_all, ratio=0.05):
    random.seed(FLAGS.seed)
    seq_tr, seq_va = [], []
    for item in seq_all:
        (seq_va if random.random() < ratio else seq_tr).append(item)
    return seq_tr, seq_va


def get_data(raw_data, data_dir, recommend=False):
    """lstm/run.py:243-319."""
    global _buckets
    from arecsys_b200.attributes.input_attribute import read_data
    from arecsys_b200.attributes import embed_attribute
    from arecsys_b200.utils.prepare_train import item_frequency
    from arecsys_b200.lstm.best_buckets import calculate_buckets
    (data_tr, data_va, u_attr, i_attr, item_ind2logit_ind, logit_ind2item_ind, user_index, item_index) = read_data(
        raw_data_dir=raw_data, data_dir=data_dir, combine_att=FLAGS.combine_att,
        logits_size_tr=FLAGS.item_vocab_size, thresh=FLAGS.vocab_min_thresh,
        use_user_feature=FLAGS.use_user_feature, use_item_feature=FLAGS.use_item_feature,
        no_user_id=FLAGS.no_user_id, test=FLAGS.test, mylog=mylog)
    data_tr = [p for p in data_tr if (p[1] in item_ind2logit_ind)]                     # remove unk
    item_population, p_item = item_frequency(data_tr, FLAGS.power)
    START_ID = len(item_index)
    item_ind2logit_ind[START_ID] = 0                                                    # :276-278
    seq_all = form_sequence(data_tr, maxlen=FLAGS.L)
    seq_tr0, seq_va0 = split_train_dev(seq_all, ratio=0.05)
    _buckets = sorted(calculate_buckets(seq_tr0 + seq_va0, FLAGS.L, FLAGS.n_bucket))
    seq_tr = split_buckets(seq_tr0, _buckets)
    seq_va = split_buckets(seq_va0, _buckets)
    if recommend:
        from arecsys_b200.utils.evaluate import Evaluation
        evaluation = Evaluation(raw_data, test=FLAGS.test)
        uinds = evaluation.get_uinds()
        seq_test = form_sequence_prediction(seq_all, uinds, FLAGS.L, START_ID)
        _buckets = sorted(calculate_buckets(seq_test, FLAGS.L, FLAGS.n_bucket))
        seq_test = split_buckets(seq_test, _buckets)
    else:
        seq_test, evaluation, uinds = [], None, []
    u_attr.set_model_size(FLAGS.size)
    i_attr.set_model_size(FLAGS.size)
    embAttr = embed_attribute.EmbeddingAttribute(u_attr, i_attr, FLAGS.batch_size, FLAGS.n_sampled, _buckets[-1],
                                                 FLAGS.use_sep_item, item_ind2logit_ind, logit_ind2item_ind,
                                                 seed=FLAGS.seed)
    if FLAGS.loss in ["warp", 'mw']:
        prepare_warp(embAttr, seq_tr0, seq_va0)
    return (seq_tr, seq_va, seq_test, embAttr, START_ID, item_population, p_item, evaluation, uinds, user_index,
            item_index, logit_ind2item_ind)


def create_model(session, embAttr, START_ID, run_options=None, run_metadata=None):
    """lstm/run.py:321-358 (no_user_id is hard-coded False there, quirk (7) of SURVEY 8a)."""
    from arecsys_b200.lstm.seqModel import SeqModel
    model = SeqModel(_buckets, FLAGS.size, FLAGS.num_layers, FLAGS.max_gradient_norm, FLAGS.batch_size,
                     FLAGS.learning_rate, FLAGS.learning_rate_decay_factor, embAttr,
                     withAdagrad=FLAGS.withAdagrad, num_samples=FLAGS.n_sampled, dropoutRate=FLAGS.keep_prob,
                     START_ID=START_ID, loss=FLAGS.loss, use_concat=FLAGS.use_concat, no_user_id=False,
                     output_feat=FLAGS.output_feat, no_input_item_feature=FLAGS.no_input_item_feature,
                     topk_n=FLAGS.topk, seed=FLAGS.seed)
    ckpt = os.path.join(FLAGS.train_dir, 'checkpoint')
    if FLAGS.recommend or ((not FLAGS.fromScratch) and os.path.isfile(ckpt)):
        path = os.path.join(FLAGS.train_dir, open(ckpt).read().split('"')[1])
        mylog("Reading model parameters from %s" % path)
        model.saver.restore(session, path)
    else:
        mylog("Created model with fresh parameters.")
    return model


def evaluate(sess, model, data_set, item_sampled_id2idx=None):
    """lstm/run.py:523-548: dev perplexity with dropout off."""
    from arecsys_b200.lstm.data_iterator import DataIterator
    model.dropout10_op()
    loss, n_valids = 0.0, 0
    ite = DataIterator(model, data_set, len(_buckets), FLAGS.batch_size, None).next_sequence(stop=True)
    for users, inputs, outputs, weights, bucket_id in ite:
        loss += model.step(sess, users, inputs, outputs, weights, bucket_id, forward_only=True)
        n_valids += np.sum(np.sign(weights[0]))
    loss = loss / max(n_valids, 1)
    ppx = math.exp(loss) if loss < 300 else float("inf")
    model.dropoutAssign_op()
    return loss, ppx


def train(raw_data=None):
    from arecsys_b200.lstm.data_iterator import DataIterator
    from arecsys_b200.utils.prepare_train import sample_items
    raw_data = FLAGS.raw_data if raw_data is None else raw_data
    mylog("Reading Data...")
    (train_set, dev_set, test_set, embAttr, START_ID, item_population, p_item, _, _, _, _, _) = get_data(
        raw_data, data_dir=FLAGS.data_dir)
    n_targets_train = int(np.sum([np.sum([len(items) for uid, items in x]) for x in train_set]))
    train_bucket_sizes = [len(train_set[b]) for b in range(len(_buckets))]
    train_total_size = float(sum(train_bucket_sizes))
    train_buckets_scale = [sum(train_bucket_sizes[:i + 1]) / train_total_size for i in range(len(train_bucket_sizes))]
    dev_bucket_sizes = [len(dev_set[b]) for b in range(len(_buckets))]
    dev_total_size = int(sum(dev_bucket_sizes))
    batch_size = FLAGS.batch_size
    steps_per_epoch = int(train_total_size / batch_size)
    steps_per_checkpoint = max(int(steps_per_epoch / 2), 1)                             # :385
    total_steps = steps_per_epoch * FLAGS.n_epoch
    if FLAGS.max_steps:
        total_steps = min(total_steps, FLAGS.max_steps)
        steps_per_checkpoint = min(steps_per_checkpoint, total_steps)
    mylog(_buckets)
    mylog("Train:")
    mylog("total: {}".format(train_total_size))
    mylog("bucket sizes: {}".format(train_bucket_sizes))
    mylog("Dev:")
    mylog("total: {}".format(dev_total_size))
    mylog("bucket sizes: {}".format(dev_bucket_sizes))
    mylog("")
    mylog("Steps_per_epoch: {}".format(steps_per_epoch))
    mylog("Total_steps:{}".format(total_steps))
    mylog("Steps_per_checkpoint: {}".format(steps_per_checkpoint))
    sess = None
    mylog("Creating Model.. (this can take a few minutes)")
    model = create_model(sess, embAttr, START_ID)
    for name in list(embAttr.params) + list(model.dense_params()):
        mylog(name)
    ite = DataIterator(model, train_set, len(train_buckets_scale), batch_size, train_buckets_scale).next_random()
    mylog("withRandom")
    np.random.seed(FLAGS.seed)
    step_time, loss, current_step = 0.0, 0.0, 0
    low_ppx = float("inf")
    steps_per_report = 30
    n_targets_report, report_time, n_valid_sents = 0, 0, 0
    patience = FLAGS.patience
    item_sampled, item_sampled_id2idx = None, None
    while current_step < total_steps:
        start_time = time.time()
        if FLAGS.loss in ['mw', 'mce'] and current_step % FLAGS.n_resample == 0:
            item_sampled, item_sampled_id2idx = sample_items(item_population, FLAGS.n_sampled, p_item)
        else:
            item_sampled = None
        users, inputs, outputs, weights, bucket_id = next(ite)
        L = model.step(sess, users, inputs, outputs, weights, bucket_id, item_sampled=item_sampled,
                       item_sampled_id2idx=item_sampled_id2idx)
        step_time += (time.time() - start_time) / steps_per_checkpoint
        loss += L
        current_step += 1
        n_valid_sents += np.sum(np.sign(weights[0]))
        report_time += (time.time() - start_time)
        n_targets_report += np.sum(weights)
        if current_step % steps_per_report == 0:
            mylog("--------------------" + "Report" + str(current_step) + "-------------------")
            mylog("StepTime: {} Speed: {} targets / sec in total {} targets".format(
                report_time / steps_per_report, n_targets_report * 1.0 / report_time, n_targets_train))
            report_time, n_targets_report = 0, 0
        if current_step % steps_per_checkpoint == 0:
            mylog("--------------------" + "TRAIN" + str(current_step) + "-------------------")
            loss = loss / max(n_valid_sents, 1)
            perplexity = math.exp(float(loss)) if loss < 300 else float("inf")
            mylog("global step %d learning rate %.4f step-time %.2f perplexity " "%.2f" % (
                model.global_step.eval(), model.learning_rate.eval(), step_time, perplexity))
            step_time, loss, n_valid_sents = 0.0, 0.0, 0
            mylog("--------------------" + "DEV" + str(current_step) + "-------------------")
            eval_loss, eval_ppx = evaluate(sess, model, dev_set, item_sampled_id2idx=item_sampled_id2idx)
            mylog("dev: ppx: {}".format(eval_ppx))
            if eval_ppx < low_ppx:
                patience = FLAGS.patience
                low_ppx = eval_ppx
                mylog("Saving best model....")
                s = time.time()
                model.saver.save(sess, os.path.join(FLAGS.train_dir, "best.ckpt"), global_step=0, write_meta_graph=False)
                mylog("Best model saved using {} sec".format(time.time() - s))
            else:
                patience -= 1
            if patience <= 0:
                mylog("Training finished. Running out of patience.")
                break
            sys.stdout.flush()


def recommend(raw_data=None):
    """lstm/run.py:550-640."""
    from arecsys_b200.lstm.data_iterator import DataIterator
    raw_data = FLAGS.raw_data if raw_data is None else raw_data
    mylog("recommend")
    mylog("Reading Data...")
    (_, _, test_set, embAttr, START_ID, _, _, evaluation, uinds, user_index, item_index, logit_ind2item_ind) = get_data(
        raw_data, data_dir=FLAGS.data_dir, recommend=True)
    test_bucket_sizes = [len(test_set[b]) for b in range(len(_buckets))]
    mylog(_buckets)
    mylog("Test:")
    mylog("total: {}".format(int(sum(test_bucket_sizes))))
    mylog("buckets: {}".format(test_bucket_sizes))
    mylog("Creating Model")
    model = create_model(None, embAttr, START_ID)
    model.dropout10_op()
    ite = DataIterator(model, test_set, len(_buckets), FLAGS.batch_size, None).next_sequence(stop=True, recommend=True)
    n_total_user = len(uinds)
    uind2rank = {uind: r for r, uind in enumerate(uinds)}
    rec = np.zeros((n_total_user, FLAGS.topk), dtype=int)
    rec_value = np.zeros((n_total_user, FLAGS.topk), dtype=float)
    start, n_steps, n_recommended = time.time(), 0, 0
    for users, inputs, positions, valids, bucket_id in ite:
        results = model.step_recommend(None, users, inputs, positions, bucket_id)
        for i, valid in enumerate(valids):
            if valid == 1:
                n_recommended += 1
                if n_recommended % 1000 == 0:
                    mylog("Evaluating n {} bucket_id {}".format(n_recommended, bucket_id))
                uind, topk_values, topk_indexes = results[i]
                rec[uind2rank[uind], :] = topk_indexes
                rec_value[uind2rank[uind], :] = topk_values
        n_steps += 1
    mylog("Time used {} sec for {} steps {} users ".format(time.time() - start, n_steps, n_recommended))
    ind2id = {}
    for iid, iind in item_index.items():
        assert iind not in ind2id
        ind2id[iind] = iid
    uind2id = {}
    for uid, uind in user_index.items():
        assert uind not in uind2id
        uind2id[uind] = uid
    R = {uind2id[uinds[i]]: [ind2id[logit_ind2item_ind[v]] for v in list(rec[i, :])] for i in range(n_total_user)}
    evaluation.eval_on(R)
    scores_self, scores_ex = evaluation.get_scores()
    mylog("====evaluation scores (NDCG, RECALL, PRECISION, MAP) @ 2,5,10,20,30====")
    mylog("METRIC_FORMAT (self): {}".format(scores_self))
    mylog("METRIC_FORMAT (ex  ): {}".format(scores_ex))
    np.save(os.path.join(FLAGS.train_dir, "top{}_index.npy".format(FLAGS.topk)), rec)
    np.save(os.path.join(FLAGS.train_dir, "top{}_value.npy".format(FLAGS.topk)), rec_value)


def main(_=None):
    FLAGS.parse()
    print("V 2017-03-22")
    if FLAGS.test:
        FLAGS.data_dir = (FLAGS.data_dir[:-1] if FLAGS.data_dir[-1] == '/' else FLAGS.data_dir) + '_test'
    if not os.path.exists(FLAGS.train_dir):
        os.makedirs(FLAGS.train_dir)
    if FLAGS.beam_search or FLAGS.ensemble:
        print('ensemble / beam_search are broken in the reference (undefined names) and not provided')
        exit(1)
    if FLAGS.recommend:
        logging.basicConfig(filename=os.path.join(FLAGS.train_dir, "log.recommend.txt.{}".format(FLAGS.topk)),
                            level=logging.DEBUG, filemode="w")
        recommend()
    else:
        logging.basicConfig(filename=os.path.join(FLAGS.train_dir, "log.txt"), level=logging.DEBUG,
                            filemode="w" if FLAGS.fromScratch else "a")
        train()


if __name__ == "__main__":
    main()
