"""Offline ranking evaluation for `--recommend True` (reference: utils/evaluate.py:7-112,
utils/eval_metrics.py:3-74, utils/submit.py:7-61): ground-truth files res_T[_test].csv and
historical_train[_test].csv written into the raw-data directory on first use, P / R / MAP / NDCG
@ {2,5,10,20,30}, with ("self") and without ("ex") the user's training history.  Host-side post-
processing of the top-k ids the GPU returns; runs once per job.
"""
import csv
from os.path import isfile, join

import numpy as np

from .load_data import load_users, load_items, load_interactions

N_MAX = 30
_DISCOUNTS = [1.0 / np.log2(2 + n) for n in range(N_MAX)]


def _prec(x, t, Ns, cs):
    l = len(cs)
    return [0.0] * len(Ns) if l == 0 else [cs[min(n - 1, l - 1)] * 1.0 / min(n, l) for n in Ns]


def _recall(x, t, Ns, cs):
    l = len(cs)
    return [0.0] * len(Ns) if l == 0 else [cs[min(n - 1, l - 1)] * 1.0 / len(t) for n in Ns]


def _map(x, t, Ns, cs):
    l = len(x)
    if l == 0:
        return [0.0] * len(Ns)
    ap = np.cumsum([x[i] * 1.0 * cs[i] / (i + 1) for i in range(l)])
    return [ap[min(n - 1, l - 1)] / min(min(len(t), n), l) for n in Ns]


def _ndcg(x, t, Ns, cs):
    l = len(x)
    if l == 0:
        return [0.0] * len(Ns)
    n_max = min(l, N_MAX)
    dcg = np.cumsum([a * b for a, b in zip(x, _DISCOUNTS[:n_max])])
    idcg = np.cumsum(_DISCOUNTS[:len(t)][:n_max] + [0.0] * max(0, n_max - len(t)))
    return [dcg[min(n - 1, len(dcg) - 1)] / idcg[min(n - 1, len(idcg) - 1)] for n in Ns]


def metrics(X, T, Ns=(2, 5, 10, 20, 30), names=('prec', 'recall', 'map', 'ndcg')):
    """eval_metrics.py:3-31: averages over ALL users of T (users missing from X count as zero)."""
    funcs = {'prec': _prec, 'recall': _recall, 'map': _map, 'ndcg': _ndcg}
    res = {m: [0.0] * len(Ns) for m in names}
    for u, t in T.items():
        t = set(t)
        if u not in X:
            continue
        correct = [int(r in t) for r in X[u]]
        cs = np.cumsum(correct)
        for m in names:
            s = funcs[m](correct, t, list(Ns), cs)
            for i in range(len(Ns)):
                res[m][i] += s[i]
    n_users = float(len(T))
    return {m: [v / n_users for v in res[m]] for m in names}


def load_submit(sub_id, submit_dir):
    """submit.py:7-21: TSV (user_id, items) -> {user_id: [item-id strings]}."""
    out = {}
    with open(join(submit_dir, sub_id), newline='') as f:
        rd = csv.reader(f, delimiter='\t')
        next(rd)
        for row in rd:
            if not row:
                continue
            uid = int(row[0]) if row[0].lstrip('-').isdigit() else row[0]
            out[uid] = row[1].split(',') if len(row) > 1 and row[1] != '' else []
    return out


def format_submit(X, sub_id, submit_dir):
    with open(join(submit_dir, sub_id), 'w', newline='') as f:
        wr = csv.writer(f, delimiter='\t')
        wr.writerow(['user_id', 'items'])
        for k, v in X.items():
            wr.writerow([k, ','.join(str(x) for x in v) if isinstance(v, list) else v])


def combine_sub(r1, r2, opt=0, users=None):
    """submit.py:43-61: r2 with the items of r1 removed (opt=1), or r1 followed by r2 (opt=0)."""
    rec = {}
    for i in range(len(users)):
        uid = users[i, 0]
        if uid not in r1 and uid not in r2:
            continue
        seen, rec[uid] = set(), []
        for iid in r1.get(uid, []):
            if iid not in seen:
                seen.add(iid)
                if opt == 0:
                    rec[uid].append(iid)
        for iid in r2.get(uid, []):
            if iid not in seen:
                seen.add(iid)
                rec[uid].append(iid)
    return rec


class Evaluation(object):
    def __init__(self, raw_data_dir, test=False):
        res_filename = 'res_T_test.csv' if test else 'res_T.csv'
        if not isfile(join(raw_data_dir, res_filename)):
            print('eval file does not exist. creating ... ')
            self.create_eval_file(raw_data_dir)
        self.T = load_submit(res_filename, raw_data_dir)
        self.hist = load_submit('historical_train_test.csv' if test else 'historical_train.csv', raw_data_dir)
        self.Iatt, _, self.Iid2ind = load_items(raw_data_dir)
        self.Uatt, _, self.Uid2ind = load_users(raw_data_dir)
        self.Uids = self.get_uids()
        self.Uinds = [self.Uid2ind[v] for v in self.Uids]

    def get_user_n(self):
        return len(self.Uinds)

    def get_uids(self):
        return list(self.T.keys())

    def get_uinds(self):
        return self.Uinds

    def set_uinds(self, uinds):
        self.Uinds = uinds

    def eval_on(self, rec):
        self.res = rec
        for k in rec:
            rec[k] = [str(v) for v in rec[k]]
        r_ex = combine_sub(self.hist, rec, 1, users=self.Uatt)
        self.s_self = [x for sub in metrics(rec, self.T).values() for x in sub]
        self.s_ex = [x for sub in metrics(r_ex, self.T).values() for x in sub]

    def get_scores(self):
        return self.s_self, self.s_ex

    def create_eval_file(self, raw_data):
        """evaluate.py:64-112: per-user item lists, most recent first."""
        (tr, va, te), _ = load_interactions(raw_data)

        def seqs(rows, with_time=True):
            d = {}
            for r in rows:
                d.setdefault(r[0], []).append((r[1], r[2]))
            if with_time:
                return {u: [p[0] for p in sorted(v, key=lambda x: x[1], reverse=True)] for u, v in d.items()}
            return {u: [p[0] for p in v] for u, v in d.items()}
        seq_tr, seq_va, seq_te = seqs(tr), seqs(va), seqs(te, False)
        format_submit(dict(seq_tr), 'historical_train.csv', raw_data)
        format_submit(dict(seq_va), 'res_T.csv', raw_data)
        format_submit(dict(seq_te), 'res_T_test.csv', raw_data)
        both = {u: list(v) for u, v in seq_va.items()}
        for u in seq_tr:
            if u in both:
                both[u] = both[u] + seq_tr[u]
        format_submit(both, 'historical_train_test.csv', raw_data)
