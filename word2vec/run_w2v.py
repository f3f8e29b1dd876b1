"""run_w2v.py — CBOW runner with the reference's flag surface (word2vec/run_w2v.py:24-88, 39 flags),
corpus construction (:101-190), training loop (:232-422) and --recommend (:424-496).
`examples/run_w2v.sh` (cd ../word2vec; python run_w2v.py --flags) runs unchanged for --model cbow;
the skip-gram tower is out of scope and exits like an unknown model.
"""
import logging
import math
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import arecsys_b200  # noqa: E402,F401
from arecsys_b200.utils.flags import Flags  # noqa: E402

FLAGS = Flags()
FLAGS.DEFINE_string("model", "cbow", "cbow or sg (skip-gram: not provided)")
FLAGS.DEFINE_string("dataset", "xing", ".")
FLAGS.DEFINE_string("raw_data", "../raw_data", "input data directory")
FLAGS.DEFINE_string("data_dir", "./data0", "Data directory")
FLAGS.DEFINE_string("train_dir", "./test0", "Training directory.")
FLAGS.DEFINE_boolean("test", False, "Test on test splits")
FLAGS.DEFINE_string("combine_att", 'mix', "method to combine attributes: het or mix")
FLAGS.DEFINE_boolean("use_user_feature", True, "RT")
FLAGS.DEFINE_boolean("use_item_feature", True, "RT")
FLAGS.DEFINE_integer("user_vocab_size", 150000, "User vocabulary size.")
FLAGS.DEFINE_integer("item_vocab_size", 50000, "Item vocabulary size.")
FLAGS.DEFINE_integer("vocab_min_thresh", 2, "filter inactive tokens.")
FLAGS.DEFINE_string("loss", 'ce', "loss function: ce, warp, (mw, mce, bpr)")
FLAGS.DEFINE_float("learning_rate", 0.1, "Learning rate.")
FLAGS.DEFINE_float("keep_prob", 0.5, "dropout rate.")
FLAGS.DEFINE_float("learning_rate_decay_factor", 1.0, "Learning rate decays by this much.")
FLAGS.DEFINE_integer("batch_size", 64, "Batch size to use during training.")
FLAGS.DEFINE_integer("size", 20, "Size of each embedding.")
FLAGS.DEFINE_integer("patience", 20, "exit if the model can't improve for $patence evals")
FLAGS.DEFINE_integer("n_epoch", 1000, "How many epochs to train.")
FLAGS.DEFINE_integer("steps_per_checkpoint", 4000, "How many training steps to do per checkpoint.")
FLAGS.DEFINE_boolean("recommend", False, "Set to True for recommend items.")
FLAGS.DEFINE_integer("top_N_items", 100, "number of items output")
FLAGS.DEFINE_boolean("recommend_new", False, "Set to True for recommend new items that were not used to train.")
FLAGS.DEFINE_float("power", 0.5, "related to sampling rate.")
FLAGS.DEFINE_integer("n_resample", 50, "iterations before resample.")
FLAGS.DEFINE_integer("n_sampled", 1024, "sampled softmax/warp loss.")
FLAGS.DEFINE_float("user_sample", 1.0, "user sample rate.")
FLAGS.DEFINE_integer("output_feat", 1, "0: no use, 1: use, mean-mulhot, 2: use, max-pool")
FLAGS.DEFINE_boolean("use_sep_item", True, "use separate embedding parameters for output items.")
FLAGS.DEFINE_boolean("no_user_id", False, "use user id or not")
FLAGS.DEFINE_integer("ni", 2, "# of input items.")
FLAGS.DEFINE_integer("num_skips", 3, "# of output context words for each input word.")
FLAGS.DEFINE_integer("skip_window", 5, "Size of each model layer.")
FLAGS.DEFINE_boolean("device_log", False, "Set to True for logging device usages.")
FLAGS.DEFINE_boolean("eval", True, "Set to True for evaluation.")
FLAGS.DEFINE_boolean("use_more_train", False, "Set true if use non-appearred items to train.")
FLAGS.DEFINE_boolean("profile", False, "False = no profile, True = profile")
FLAGS.DEFINE_boolean("after40", False, "whether use items after week 40 only.")
FLAGS.DEFINE_integer("max_steps", 0, "not in the reference: stop after this many steps (0 = n_epoch decides)")


def mylog(msg):
    print(msg)
    logging.info(msg)


def get_user_items_seq(data):
    """run_w2v.py:101-116: user -> time-ordered item list."""
    d = {}
    for u, i, t in data:
        d.setdefault(u, []).append((i, t))
    return {u: [x[0] for x in sorted(v, key=lambda x: x[1])] for u, v in d.items()}


def form_train_seq(x, pad_token, opt=1):
    """run_w2v.py:118-131: concatenated corpus with a PAD event before each user."""
    seq = []
    for u in x:
        l = [(u, i) for i in x[u]]
        if opt == 0:
            seq.extend(l)
            seq.append((u, pad_token))
        else:
            seq.append((u, pad_token))
            seq.extend(l)
    return seq


def prepare_valid(data_va, u_i_seq_tr, end_ind, n=0):
    """run_w2v.py:133-157: the last n training items of every validation user (PAD padded)."""
    res, processed = {}, set()
    for u, _, _ in data_va:
        if u in processed:
            continue
        processed.add(u)
        if u in u_i_seq_tr:
            if n == 0:
                res[u] = []
            elif n == -1:
                res[u] = [end_ind]
            else:
                items = list(u_i_seq_tr[u][-n:])
                res[u] = items + [end_ind] * (n - len(items))
        else:
            res[u] = [end_ind] if n == -1 else [end_ind] * n
    return res


def get_data(raw_data, data_dir):
    from arecsys_b200.attributes.input_attribute import read_data
    (data_tr0, data_va0, u_attr, i_attr, item_ind2logit_ind, logit_ind2item_ind, user_index, item_index) = read_data(
        raw_data_dir=raw_data, data_dir=data_dir, combine_att=FLAGS.combine_att,
        logits_size_tr=FLAGS.item_vocab_size, thresh=FLAGS.vocab_min_thresh,
        use_user_feature=FLAGS.use_user_feature, use_item_feature=FLAGS.use_item_feature,
        no_user_id=FLAGS.no_user_id, test=FLAGS.test, mylog=mylog)
    mylog('length of item_ind2logit_ind: {}'.format(len(item_ind2logit_ind)))
    mylog("original train/dev size: %d/%d" % (len(data_tr0), len(data_va0)))
    data_tr = [p for p in data_tr0 if (p[1] in item_ind2logit_ind)]
    data_va = [p for p in data_va0 if (p[1] in item_ind2logit_ind)]
    mylog("new train/dev size: %d/%d" % (len(data_tr), len(data_va)))
    u_i_seq_tr = get_user_items_seq(data_tr)
    PAD_ID = len(item_index)
    seq_tr = form_train_seq(u_i_seq_tr, PAD_ID)
    items_dev = prepare_valid(data_va0, u_i_seq_tr, PAD_ID, max(FLAGS.ni, 1) if FLAGS.ni != 0 else 1)
    return (seq_tr, items_dev, data_tr, data_va, u_attr, i_attr, item_ind2logit_ind, logit_ind2item_ind, PAD_ID,
            user_index, item_index)


def create_model(session, u_attributes=None, i_attributes=None, item_ind2logit_ind=None,
                 logit_ind2item_ind=None, loss=None, ind_item=None):
    n_sampled = FLAGS.n_sampled if FLAGS.loss in ['mw', 'mce'] else None
    if FLAGS.model == 'cbow':
        from arecsys_b200.word2vec import cbow_model as w2v_model
    elif FLAGS.model == 'sg':
        from arecsys_b200.word2vec import skipgram_model as w2v_model          # run_w2v.py:202-203
    else:
        mylog('not implemented error')
        exit(1)
    model = w2v_model.Model(FLAGS.user_vocab_size, FLAGS.item_vocab_size, FLAGS.size, FLAGS.batch_size,
                            FLAGS.learning_rate, FLAGS.learning_rate_decay_factor, u_attributes, i_attributes,
                            item_ind2logit_ind, logit_ind2item_ind, loss_function=loss or FLAGS.loss,
                            n_input_items=FLAGS.ni, use_sep_item=FLAGS.use_sep_item, dropout=FLAGS.keep_prob,
                            top_N_items=FLAGS.top_N_items, output_feat=FLAGS.output_feat, n_sampled=n_sampled)
    if not os.path.isdir(FLAGS.train_dir):
        os.mkdir(FLAGS.train_dir)
    ckpt = os.path.join(FLAGS.train_dir, 'checkpoint')
    if os.path.isfile(ckpt):
        path = os.path.join(FLAGS.train_dir, open(ckpt).read().split('"')[1])
        mylog("Reading model parameters from %s" % path)
        model.saver.restore(session, path)
    else:
        mylog("Created model with fresh parameters.")
    return model


def train(raw_data=None):
    from arecsys_b200.utils.prepare_train import item_frequency, sample_items, positive_items
    from arecsys_b200.word2vec.data_iterator import DataIterator
    raw_data = FLAGS.raw_data if raw_data is None else raw_data
    if FLAGS.profile:
        FLAGS.steps_per_checkpoint = 30
    mylog("reading data")
    (seq_tr, items_dev, data_tr, data_va, u_attributes, i_attributes, item_ind2logit_ind, logit_ind2item_ind,
     end_ind, _, _) = get_data(raw_data, data_dir=FLAGS.data_dir)
    item_pop, p_item = item_frequency(data_tr, FLAGS.power)
    item_population = list(range(len(item_ind2logit_ind))) if FLAGS.use_more_train else item_pop
    sess = None
    model = create_model(sess, u_attributes, i_attributes, item_ind2logit_ind, logit_ind2item_ind,
                         loss=FLAGS.loss, ind_item=item_population)
    if FLAGS.loss in ['warp', 'mw', 'bbpr']:
        model.prepare_warp(*positive_items(data_tr, data_va))
    np.random.seed(0)
    if FLAGS.model == 'sg':                                                     # run_w2v.py:259-266
        ite = DataIterator(seq_tr, end_ind, FLAGS.batch_size, FLAGS.num_skips, FLAGS.skip_window, False).get_next_sg()
    else:
        ite = DataIterator(seq_tr, end_ind, FLAGS.batch_size, max(FLAGS.ni, 1), FLAGS.skip_window, False).get_next_cbow()
    mylog('started training')
    step_time, loss, current_step = 0.0, 0.0, 0
    patience = FLAGS.patience
    previous_losses, losses_dev = [], []
    best_auc, best_loss = -1, 1000000                                         # run_w2v.py:284 (the auc is never computed)
    item_sampled, item_sampled_id2idx = None, None
    train_total_size = float(len(data_tr))
    steps_per_epoch = int(1.0 * train_total_size / FLAGS.batch_size)
    total_steps = steps_per_epoch * FLAGS.n_epoch
    if FLAGS.max_steps:
        total_steps = min(total_steps, FLAGS.max_steps)
    mylog("Train:")
    mylog("total: {}".format(train_total_size))
    mylog("Steps_per_epoch: {}".format(steps_per_epoch))
    mylog("Total_steps:{}".format(total_steps))
    mylog("Dev:")
    mylog("total: {}".format(len(data_va)))
    while True:
        start_time = time.time()
        (user_input, input_items, output_items) = next(ite)
        if current_step < 5:                                                  # run_w2v.py:305-312: the first batches, verbatim
            as_list = lambda x: x.tolist() if hasattr(x, 'tolist') else x
            mylog("current step is {}".format(current_step))
            mylog('user')
            mylog(as_list(user_input))
            mylog('input_item')
            mylog(as_list(input_items))
            mylog('output_item')
            mylog(as_list(output_items))
        if FLAGS.loss in ['mw', 'mce'] and current_step % FLAGS.n_resample == 0:
            item_sampled, item_sampled_id2idx = sample_items(item_population, FLAGS.n_sampled, p_item)
        else:
            item_sampled = None
        step_loss = model.step(sess, user_input, input_items, output_items, item_sampled, item_sampled_id2idx,
                               loss=FLAGS.loss)
        step_time += (time.time() - start_time) / FLAGS.steps_per_checkpoint
        loss += step_loss / FLAGS.steps_per_checkpoint
        current_step += 1
        if current_step > total_steps:
            mylog("Training reaches maximum steps. Terminating...")
            break
        if current_step % FLAGS.steps_per_checkpoint == 0:
            if FLAGS.loss in ['ce', 'mce']:
                perplexity = math.exp(loss) if loss < 300 else float('inf')
                mylog("global step %d learning rate %.4f step-time %.4f perplexity %.2f" % (
                    model.global_step.eval(), model.learning_rate.eval(), step_time, perplexity))
            else:
                mylog("global step %d learning rate %.4f step-time %.4f loss %.3f" % (
                    model.global_step.eval(), model.learning_rate.eval(), step_time, loss))
            mylog("  throughput %.0f events/s" % (FLAGS.batch_size / max(step_time, 1e-9)))
            if len(previous_losses) > 2 and loss > max(previous_losses[-3:]):
                model.learning_rate_decay_op()
            previous_losses.append(loss)
            step_time, loss = 0.0, 0.0
            if not FLAGS.eval:
                continue
            l_va = len(data_va)
            eval_loss, count_va = 0.0, 0
            start_time = time.time()
            for idx_s in range(0, l_va, FLAGS.batch_size):
                idx_e = idx_s + FLAGS.batch_size
                if idx_e > l_va:
                    break
                lt = data_va[idx_s:idx_e]
                user_va = [x[0] for x in lt]
                item_va_input = list(map(list, zip(*[items_dev[x[0]] for x in lt])))
                item_va = [x[1] for x in lt]
                the_loss = 'warp' if FLAGS.loss == 'mw' else FLAGS.loss
                eval_loss += model.step(sess, user_va, item_va_input, item_va, forward_only=True, loss=the_loss)
                count_va += 1
            eval_loss /= max(count_va, 1)
            eval_auc = 0.0
            step_time = (time.time() - start_time) / max(count_va, 1)
            if FLAGS.loss in ['ce', 'mce']:
                eval_ppx = math.exp(eval_loss) if eval_loss < 300 else float('inf')
                mylog("  dev: perplexity %.2f eval_auc %.4f step-time %.4f" % (eval_ppx, eval_auc, step_time))
            else:
                mylog("  dev: loss %.3f eval_auc %.4f step-time %.4f" % (eval_loss, eval_auc, step_time))
            sys.stdout.flush()
            step_time = 0.0
            if eval_loss < best_loss and not FLAGS.test:
                best_loss = eval_loss
                patience = FLAGS.patience
                model.saver.save(sess, os.path.join(FLAGS.train_dir, "best.ckpt"), global_step=0, write_meta_graph=False)
                mylog('Saving best model...')
            if FLAGS.test:
                model.saver.save(sess, os.path.join(FLAGS.train_dir, "best.ckpt"), global_step=0, write_meta_graph=False)
                mylog('Saving current model...')
            if eval_loss > best_loss:
                patience -= 1
            losses_dev.append(eval_loss)
            if patience < 0 and not FLAGS.test:
                mylog("no improvement for too long.. terminating..")
                mylog("best auc %.4f" % best_auc)
                mylog("best loss %.4f" % best_loss)
                sys.stdout.flush()
                break


def recommend(raw_data=None):
    from arecsys_b200.utils.evaluate import Evaluation
    raw_data = FLAGS.raw_data if raw_data is None else raw_data
    batch_size, topN = FLAGS.batch_size, FLAGS.top_N_items
    mylog("reading data")
    (_, items_dev, _, _, u_attributes, i_attributes, item_ind2logit_ind, logit_ind2item_ind, _, user_index,
     item_index) = get_data(raw_data, data_dir=FLAGS.data_dir)
    evaluation = Evaluation(raw_data, test=FLAGS.test)
    model = create_model(None, u_attributes, i_attributes, item_ind2logit_ind, logit_ind2item_ind, loss=FLAGS.loss)
    Uinds = evaluation.get_uinds()
    N = len(Uinds)
    mylog("N = %d" % N)
    keep = [k for k, p in enumerate(Uinds) if p in items_dev]
    uids = [evaluation.get_uids()[k] for k in keep]
    Uinds = [Uinds[k] for k in keep]
    mylog("new N = {}, (reduced from original {})".format(len(Uinds), N))
    if len(Uinds) < N:
        evaluation.set_uinds(Uinds)
    N = len(Uinds)
    rec = np.zeros((N, topN), dtype=int)
    time_start = time.time()
    for count, idx_s in enumerate(range(0, N, batch_size)):
        if (count + 1) % 100 == 0:
            mylog("idx: %d, c: %d" % (idx_s, count + 1))
        idx_e = idx_s + batch_size
        sel = list(range(idx_s, min(idx_e, N))) + [0] * max(0, idx_e - N)
        users = [Uinds[t] for t in sel]
        items_input = list(map(list, zip(*[items_dev[u] for u in users])))
        recs = model.step(None, users, items_input, forward_only=True, recommend=True, recommend_new=FLAGS.recommend_new)
        rec[idx_s:min(idx_e, N), :] = recs[:min(idx_e, N) - idx_s, :]
    mylog("Time used %.1f" % (time.time() - time_start))
    ind2id = {}
    for iid, ind in item_index.items():
        assert ind not in ind2id
        ind2id[ind] = iid
    R = {uids[i]: [ind2id[logit_ind2item_ind[v]] for v in list(rec[i, :])] for i in range(N)}
    evaluation.eval_on(R)
    scores_self, scores_ex = evaluation.get_scores()
    mylog("====evaluation scores (NDCG, RECALL, PRECISION, MAP) @ 2,5,10,20,30====")
    mylog("METRIC_FORMAT (self): {}".format(scores_self))
    mylog("METRIC_FORMAT (ex  ): {}".format(scores_ex))


def main(_=None):
    FLAGS.parse()
    if FLAGS.test:
        FLAGS.data_dir = (FLAGS.data_dir[:-1] if FLAGS.data_dir[-1] == '/' else FLAGS.data_dir) + '_test'
    if not os.path.exists(FLAGS.train_dir):
        os.makedirs(FLAGS.train_dir)
    if not FLAGS.recommend:
        logging.basicConfig(filename=os.path.join(FLAGS.train_dir, "log.txt"), level=logging.DEBUG)
        train()
    else:
        logging.basicConfig(filename=os.path.join(FLAGS.train_dir, "log.recommend.txt"), level=logging.DEBUG)
        recommend()


if __name__ == "__main__":
    main()
