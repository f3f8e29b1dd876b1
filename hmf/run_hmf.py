"""run_hmf.py — HMF runner with the reference's flag surface (hmf/run_hmf.py:18-80, 41 flags),
side effects (train_dir/log.txt, best.ckpt-0, data_dir cache) and training loop
(:127-338), on top of arecsys_b200.  `examples/run_hmf.sh` (which does `cd ../hmf; python
run_hmf.py --flags`) runs unchanged.  Unknown flags are ignored like TF-1.0's parse_known_args.
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
# datasets, paths, and preprocessing
FLAGS.DEFINE_string("dataset", "xing", ".")
FLAGS.DEFINE_string("raw_data", "../raw_data", "input data directory")
FLAGS.DEFINE_string("data_dir", "./cache0", "Cached data directory")
FLAGS.DEFINE_string("train_dir", "./tmp", "Training directory.")
FLAGS.DEFINE_boolean("test", False, "Test on test splits")
FLAGS.DEFINE_string("combine_att", 'mix', "method to combine attributes: het or mix")
FLAGS.DEFINE_boolean("use_user_feature", True, "RT")
FLAGS.DEFINE_boolean("use_item_feature", True, "RT")
FLAGS.DEFINE_integer("user_vocab_size", 150000, "User vocabulary size.")
FLAGS.DEFINE_integer("item_vocab_size", 50000, "Item vocabulary size.")
FLAGS.DEFINE_integer("item_vocab_min_thresh", 2, "filter inactive tokens.")
# tuning hypers
FLAGS.DEFINE_string("loss", 'ce', "loss function: ce, warp, (mw, mce, bpr)")
FLAGS.DEFINE_string("loss_func", 'log', "loss function: log, exp, poly")
FLAGS.DEFINE_float("loss_exp_p", 1.0005, "p in 1-p^{-x}; or in x^p")
FLAGS.DEFINE_float("learning_rate", 0.1, "Learning rate.")
FLAGS.DEFINE_float("keep_prob", 0.5, "dropout rate.")
FLAGS.DEFINE_float("learning_rate_decay_factor", 1.0, "Learning rate decays by this much.")
FLAGS.DEFINE_integer("batch_size", 64, "Batch size to use during training.")
FLAGS.DEFINE_integer("size", 20, "Size of each embedding.")
FLAGS.DEFINE_integer("patience", 20, "exit if the model can't improve for $patience evals")
FLAGS.DEFINE_integer("n_epoch", 1000, "How many epochs to train.")
FLAGS.DEFINE_integer("steps_per_checkpoint", 4000, "How many training steps to do per checkpoint.")
# to recommend
FLAGS.DEFINE_boolean("recommend", False, "Set to True for recommend items.")
FLAGS.DEFINE_string("saverec", False, "")
FLAGS.DEFINE_integer("top_N_items", 100, "number of items output")
FLAGS.DEFINE_boolean("recommend_new", False, "Set to True for recommend new items that were not used to train.")
# nonlinear
FLAGS.DEFINE_string("nonlinear", 'linear', "nonlinear activation")
FLAGS.DEFINE_integer("hidden_size", 500, "when nonlinear proj used")
FLAGS.DEFINE_integer("num_layers", 1, "Number of layers in the model.")
# algorithms with sampling
FLAGS.DEFINE_float("power", 0.5, "related to sampling rate.")
FLAGS.DEFINE_integer("n_resample", 50, "iterations before resample.")
FLAGS.DEFINE_integer("n_sampled", 1024, "sampled softmax/warp loss.")
FLAGS.DEFINE_string("sample_type", 'random', "random, sweep, permute")
FLAGS.DEFINE_float("user_sample", 1.0, "user sample rate.")
FLAGS.DEFINE_integer("seed", 0, "mini batch sampling random seed.")
#
FLAGS.DEFINE_integer("gpu", -1, "gpu card number")
FLAGS.DEFINE_boolean("profile", False, "False = no profile, True = profile")
FLAGS.DEFINE_boolean("device_log", False, "Set to True for logging device usages.")
FLAGS.DEFINE_boolean("eval", True, "Set to True for evaluation.")
FLAGS.DEFINE_boolean("use_more_train", False, "Set true if use non-appearred items to train.")
FLAGS.DEFINE_string("model_option", 'loss', "model to evaluation")
# not in the reference: cap on the number of steps (smoke runs / tests)
FLAGS.DEFINE_integer("max_steps", 0, "stop after this many steps (0 = n_epoch decides)")


# Not in the reference (single GPU): launched under torchrun (WORLD_SIZE > 1), the tables are row-sharded over the ranks
# (arecsys_b200.hmf.sharded.ShardedLatentProductModel).  Every rank runs this same script with the same seeds,
# so all of them assemble the same GLOBAL batches (--batch_size is the global batch) and draw the same negative pools;
# rank r scores rows [r * mb, (r + 1) * mb) of each batch.  Only rank 0 logs.
WORLD = int(os.environ.get('WORLD_SIZE', '1'))
RANK = int(os.environ.get('RANK', '0'))


def _dist_setup():
    if WORLD == 1:
        return
    import torch
    import torch.distributed as dist
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    if not dist.is_initialized():
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    if FLAGS.batch_size % WORLD:
        raise SystemExit('--batch_size is the GLOBAL batch: it must be a multiple of the %d ranks' % WORLD)
    if FLAGS.loss != 'mw':
        raise SystemExit('row-sharded training covers --loss mw (SURVEY 8e); got %r' % FLAGS.loss)


def _rank0_first(fn):
    """Run fn on rank 0 first (it writes the attribute-store cache), then on the other ranks (they read it)."""
    if WORLD == 1:
        return fn()
    import torch.distributed as dist
    out = fn() if RANK == 0 else None
    dist.barrier()
    if RANK != 0:
        out = fn()
    dist.barrier()
    return out


def mylog(msg):
    if RANK != 0:
        return
    print(msg)
    logging.info(msg)


def create_model(session, u_attributes=None, i_attributes=None, item_ind2logit_ind=None,
                 logit_ind2item_ind=None, loss=None, logit_size_test=None, ind_item=None):
    """run_hmf.py:96-125: build the model, restore train_dir's checkpoint if there is one."""
    from arecsys_b200.hmf import hmf_model
    loss = FLAGS.loss if loss is None else loss
    gpu = None if FLAGS.gpu == -1 else FLAGS.gpu
    n_sampled = FLAGS.n_sampled if FLAGS.loss in ['mw', 'mce'] else None
    cls = hmf_model.LatentProductModel
    if WORLD > 1:
        from arecsys_b200.hmf.sharded import ShardedLatentProductModel as cls
        gpu = int(os.environ.get('LOCAL_RANK', '0'))
    model = cls(
        FLAGS.user_vocab_size, FLAGS.item_vocab_size, FLAGS.size, FLAGS.num_layers, FLAGS.batch_size,
        FLAGS.learning_rate, FLAGS.learning_rate_decay_factor, u_attributes, i_attributes,
        item_ind2logit_ind, logit_ind2item_ind, loss_function=loss, GPU=gpu,
        logit_size_test=logit_size_test, nonlinear=FLAGS.nonlinear, dropout=FLAGS.keep_prob,
        n_sampled=n_sampled, indices_item=ind_item, top_N_items=FLAGS.top_N_items,
        hidden_size=FLAGS.hidden_size, loss_func=FLAGS.loss_func, loss_exp_p=FLAGS.loss_exp_p,
        seed=FLAGS.seed)
    os.makedirs(FLAGS.train_dir, exist_ok=True)
    ckpt = os.path.join(FLAGS.train_dir, 'checkpoint')
    if os.path.isfile(ckpt):
        name = open(ckpt).read().split('"')[1]
        path = os.path.join(FLAGS.train_dir, name)
        mylog("Reading model parameters from %s" % path)
        model.saver.restore(session, path)
    else:
        mylog("Created model with fresh parameters.")
    return model


def train():
    from arecsys_b200.attributes.input_attribute import read_data
    from arecsys_b200.utils.prepare_train import positive_items, item_frequency, sample_items
    raw_data, train_dir, data_dir = FLAGS.raw_data, FLAGS.train_dir, FLAGS.data_dir
    batch_size, steps_per_checkpoint = FLAGS.batch_size, FLAGS.steps_per_checkpoint
    loss_func, max_patience, go_test = FLAGS.loss, FLAGS.patience, FLAGS.test
    profile = FLAGS.profile
    if profile:
        steps_per_checkpoint = 30                                             # :146
    sess = None
    mylog("reading data")
    (data_tr, data_va, u_attributes, i_attributes, item_ind2logit_ind, logit_ind2item_ind, _, _) = _rank0_first(lambda: read_data(
        raw_data_dir=raw_data, data_dir=data_dir, combine_att=FLAGS.combine_att,
        logits_size_tr=FLAGS.item_vocab_size, thresh=FLAGS.item_vocab_min_thresh,
        use_user_feature=FLAGS.use_user_feature, use_item_feature=FLAGS.use_item_feature,
        test=FLAGS.test, mylog=mylog))
    mylog("train/dev size: %d/%d" % (len(data_tr), len(data_va)))
    # remove rare items from both sets (run_hmf.py:163-171)
    mylog("original train/dev size: %d/%d" % (len(data_tr), len(data_va)))
    data_tr = [p for p in data_tr if (p[1] in item_ind2logit_ind)]
    data_va = [p for p in data_va if (p[1] in item_ind2logit_ind)]
    mylog("new train/dev size: %d/%d" % (len(data_tr), len(data_va)))

    random.seed(FLAGS.seed)
    np.random.seed(FLAGS.seed)
    item_pop, p_item = item_frequency(data_tr, FLAGS.power)
    item_population = list(range(len(item_ind2logit_ind))) if FLAGS.use_more_train else item_pop
    model = create_model(sess, u_attributes, i_attributes, item_ind2logit_ind, logit_ind2item_ind,
                         loss=loss_func, ind_item=item_population)
    if loss_func in ['warp', 'mw', 'rs', 'rs-sig', 'rs-sig2', 'bbpr']:
        pos_item_list, pos_item_list_val = positive_items(data_tr, data_va)
        model.prepare_warp(pos_item_list, pos_item_list_val)

    mylog('started training')
    step_time, loss, current_step = 0.0, 0.0, 0
    repeat = 5 if loss_func.startswith('bpr') else 1
    patience = max_patience
    previous_losses, losses_dev = [], []
    best_loss = 1000000
    item_sampled, item_sampled_id2idx = None, None
    if FLAGS.sample_type == 'random':
        get_next_batch = model.get_batch
    elif FLAGS.sample_type == 'permute':
        get_next_batch = model.get_permuted_batch
    else:
        print('not implemented!')
        exit()
    train_total_size = float(len(data_tr))
    steps_per_epoch = int(1.0 * train_total_size / batch_size)
    total_steps = steps_per_epoch * FLAGS.n_epoch
    if FLAGS.max_steps:
        total_steps = min(total_steps, FLAGS.max_steps)
    mylog("Train:")
    mylog("total: {}".format(train_total_size))
    mylog("Steps_per_epoch: {}".format(steps_per_epoch))
    mylog("Total_steps:{}".format(total_steps))
    mylog("Dev:")
    mylog("total: {}".format(len(data_va)))
    mylog("\n\ntraining start!")
    torch_prof = None
    if profile:
        import torch
        torch_prof = torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CPU,
                                                        torch.profiler.ProfilerActivity.CUDA])
        torch_prof.__enter__()
    # Batches assembled on the device (utils/device_batch.py): the same interaction stream as model.get_batch /
    # get_permuted_batch under the same seeds (Python's / NumPy's generators are consumed identically), as device tensors.
    # ARX_DEVICE_BATCH=0 keeps the reference's per-example Python loop.
    device_sampler = None
    if os.environ.get('ARX_DEVICE_BATCH', '1') == '1':
        from arecsys_b200.utils.device_batch import DeviceInteractionSampler
        device_sampler = DeviceInteractionSampler(data_tr, batch_size, model.device,
                                                  'permute' if FLAGS.sample_type == 'permute' else 'random')
    while True:
        start_time = time.time()
        if device_sampler is not None:
            user_input, item_input = device_sampler.next()
            neg_item_input = None
        else:
            (user_input, item_input, neg_item_input) = get_next_batch(data_tr)
        if loss_func in ['mw', 'mce'] and current_step % FLAGS.n_resample == 0:
            item_sampled, item_sampled_id2idx = sample_items(item_population, FLAGS.n_sampled, p_item)
        else:
            item_sampled = None
        step_loss = model.step(sess, user_input, item_input, neg_item_input, item_sampled,
                               item_sampled_id2idx, loss=loss_func)
        step_time += (time.time() - start_time) / steps_per_checkpoint
        loss += step_loss / steps_per_checkpoint
        current_step += 1
        if model.global_step.eval() > total_steps:
            mylog("Training reaches maximum steps. Terminating...")
            break
        if current_step % steps_per_checkpoint == 0:
            if loss_func in ['ce', 'mce']:
                perplexity = math.exp(loss) if loss < 300 else float('inf')
                mylog("global step %d learning rate %.4f step-time %.4f perplexity %.2f" % (
                    model.global_step.eval(), model.learning_rate.eval(), step_time, perplexity))
            else:
                mylog("global step %d learning rate %.4f step-time %.4f loss %.3f" % (
                    model.global_step.eval(), model.learning_rate.eval(), step_time, loss))
            mylog("  throughput %.0f interactions/s" % (batch_size / max(step_time, 1e-9)))
            if profile:
                torch_prof.__exit__(None, None, None)
                torch_prof.export_chrome_trace('timeline.json')                # :260-266
                exit()
            # Decrease learning rate if no improvement was seen over last 3 times.
            if len(previous_losses) > 2 and loss > max(previous_losses[-3:]):
                model.learning_rate_decay_op()
            previous_losses.append(loss)
            step_time, loss = 0.0, 0.0
            if not FLAGS.eval:
                continue
            # dev loss: consecutive slices, last partial slice dropped (:286-300)
            l_va = len(data_va)
            eval_loss, count_va = 0.0, 0
            start_time = time.time()
            for idx_s in range(0, l_va, batch_size):
                idx_e = idx_s + batch_size
                if idx_e > l_va:
                    break
                lt = data_va[idx_s:idx_e]
                user_va = [x[0] for x in lt]
                item_va = [x[1] for x in lt]
                for _ in range(repeat):
                    the_loss = 'warp' if loss_func == 'mw' else loss_func
                    eval_loss += model.step(sess, user_va, item_va, None, None, None, forward_only=True,
                                            loss=the_loss)
                    count_va += 1
            eval_loss /= max(count_va, 1)
            eval_auc = 0.0
            step_time = (time.time() - start_time) / max(count_va, 1)
            if loss_func in ['ce', 'mce']:
                eval_ppx = math.exp(eval_loss) if eval_loss < 300 else float('inf')
                mylog("  dev: perplexity %.2f eval_auc(not computed) %.4f step-time %.4f" % (
                    eval_ppx, eval_auc, step_time))
            else:
                mylog("  dev: loss %.3f eval_auc(not computed) %.4f step-time %.4f" % (eval_loss, eval_auc, step_time))
            sys.stdout.flush()
            if eval_loss < best_loss and not go_test:
                best_loss = eval_loss
                patience = max_patience
                mylog('Saving best model...')
                model.saver.save(sess, os.path.join(train_dir, "best.ckpt"), global_step=0, write_meta_graph=False)
            if go_test:
                mylog('Saving best model...')
                model.saver.save(sess, os.path.join(train_dir, "best.ckpt"), global_step=0, write_meta_graph=False)
            if eval_loss > best_loss:
                patience -= 1
            losses_dev.append(eval_loss)
            step_time = 0.0
            if patience < 0 and not go_test:
                mylog("no improvement for too long.. terminating..")
                mylog("best loss %.4f" % best_loss)
                sys.stdout.flush()
                break
    return


def recommend(target_uids=[]):
    """run_hmf.py:340-409: top-N items for the given raw user ids -> {uid: [item id, ...]}."""
    from arecsys_b200.attributes.input_attribute import read_data
    batch_size, top_n = FLAGS.batch_size, FLAGS.top_N_items
    mylog("reading data")
    (_, _, u_attributes, i_attributes, item_ind2logit_ind, logit_ind2item_ind, user_index, item_index) = _rank0_first(lambda: read_data(
        raw_data_dir=FLAGS.raw_data, data_dir=FLAGS.data_dir, combine_att=FLAGS.combine_att,
        logits_size_tr=FLAGS.item_vocab_size, thresh=FLAGS.item_vocab_min_thresh,
        use_user_feature=FLAGS.use_user_feature, use_item_feature=FLAGS.use_item_feature,
        test=FLAGS.test, mylog=mylog))
    model = create_model(None, u_attributes, i_attributes, item_ind2logit_ind, logit_ind2item_ind,
                         loss=FLAGS.loss, ind_item=None)
    Uinds = [user_index[v] for v in target_uids]
    N = len(Uinds)
    mylog("%d target users to recommend" % N)
    rec = np.zeros((N, top_n), dtype=int)
    time_start = time.time()
    for count, idx_s in enumerate(range(0, N, batch_size)):
        if (count + 1) % 100 == 0:
            mylog("idx: %d, c: %d" % (idx_s, count + 1))
        idx_e = idx_s + batch_size
        if idx_e <= N:
            users = Uinds[idx_s: idx_e]
            rec[idx_s:idx_e, :] = model.step(None, users, None, None, forward_only=True, recommend=True)
        else:
            users = [Uinds[t] for t in list(range(idx_s, N)) + [0] * (idx_e - N)]
            recs = model.step(None, users, None, None, forward_only=True, recommend=True)
            rec[idx_s:N, :] = recs[:(N - idx_s), :]
    mylog("Time used %.1f" % (time.time() - time_start))
    ind2id = {}
    for iid, ind in item_index.items():
        assert ind not in ind2id
        ind2id[ind] = iid
    R = {}
    for i in range(N):
        R[target_uids[i]] = [ind2id[logit_ind2item_ind[v]] for v in list(rec[i, :])]
    return R


def compute_scores():
    """run_hmf.py:411-427."""
    from arecsys_b200.utils.evaluate import Evaluation
    evaluation = _rank0_first(lambda: Evaluation(FLAGS.raw_data, test=FLAGS.test))   # writes historical_*.csv once
    R = recommend(evaluation.get_uids())
    if RANK != 0:
        return                                                # every rank holds the same top-N lists; rank 0 scores them
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
        os.makedirs(FLAGS.train_dir, exist_ok=True)
    _dist_setup()
    if not FLAGS.recommend:
        if RANK == 0:
            print('train')
            logging.basicConfig(filename=os.path.join(FLAGS.train_dir, "log.txt"), level=logging.DEBUG)
        train()
    else:
        if RANK == 0:
            print('recommend')
            logging.basicConfig(filename=os.path.join(FLAGS.train_dir, "log.recommend.txt"), level=logging.DEBUG)
        compute_scores()
    if WORLD > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
