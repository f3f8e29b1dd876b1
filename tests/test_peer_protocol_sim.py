"""CPU model of the peer-memory exchange protocol of hmf/exchange.py::PeerExchange (threads stand in for the ranks,
NumPy arrays for the receive blocks, threading.Barrier for arx_peer_barrier, a lock for the atomicity of red.add).

What it pins: with TWO barriers per step and the zeroing rule "a rank clears a receive block after it consumed it and
BEFORE it arrives at the next barrier" no push can land in a block that is still to be read or still to be cleared —
for any interleaving of the ranks between barriers — and moving the clear behind the barrier breaks it (negative
control with a forced schedule).  The kernels are not involved; this is the reasoning of DESIGN.md section 6 made
executable."""
import threading
import time

import numpy as np
import pytest


class Rank(object):
    def __init__(self, G, r, mb, S):
        self.G, self.r = G, r
        self.loc = np.zeros(mb)          # += by every rank's lookup            (reduce-scatter by push)
        self.sp = np.zeros(S)            # += by every rank's lookup            (all-reduce by push)
        self.ug = np.zeros(G * mb)       # stored by the owners of the rows     (all-gather by push)
        self.igp = np.zeros(S)           # += by every rank                     (all-reduce by push)


def contribution(kind, src, dst, k, n):
    """what rank `src` pushes into rank `dst`'s block `kind` at step k (distinct per (kind, src, dst, k))"""
    base = {'loc': 1.0, 'sp': 2.0, 'ug': 3.0, 'igp': 4.0}[kind]
    return base * 1000 + 100 * k + 10 * src + dst + np.arange(n) * 1e-3


def run(G, steps, mb, S, clear_after_barrier=False, delays=None, seed=0):
    ranks = [Rank(G, r, mb, S) for r in range(G)]
    lock = threading.Lock()                         # red.add is atomic per element; a lock is enough for the model
    bar = threading.Barrier(G)
    errors = []
    rng = np.random.default_rng(seed)
    jitter = rng.random((G, steps, 6)) * 2e-4

    def nap(r, k, phase):
        d = jitter[r, k, phase]
        if delays is not None:
            d += delays.get((r, phase), 0.0)
        if d > 0:
            time.sleep(d)

    def body(r):
        me = ranks[r]
        try:
            for k in range(steps):
                me.igp[:] = 0.0                                         # clear_pool_gradients: before B1
                nap(r, k, 0)
                for g in range(G):                                      # lookup launch in push mode
                    with lock:
                        ranks[g].loc += contribution('loc', r, g, k, mb)
                        ranks[g].sp += contribution('sp', r, g, k, S)
                nap(r, k, 1)
                bar.wait()                                              # B1
                want_loc = sum(contribution('loc', s, r, k, mb) for s in range(G))
                want_sp = sum(contribution('sp', s, r, k, S) for s in range(G))
                if not (np.allclose(me.loc, want_loc) and np.allclose(me.sp, want_sp)):
                    errors.append(('fwd', r, k))
                nap(r, k, 2)
                for g in range(G):                                      # backward push
                    ranks[g].ug[r * mb:(r + 1) * mb] = contribution('ug', r, g, k, mb)
                    with lock:
                        ranks[g].igp += contribution('igp', r, g, k, S)
                if not clear_after_barrier:
                    me.loc[:] = 0.0; me.sp[:] = 0.0                     # consumed: cleared BEFORE arriving at B2
                nap(r, k, 3)
                bar.wait()                                              # B2
                if clear_after_barrier:
                    nap(r, k, 4)
                    me.loc[:] = 0.0; me.sp[:] = 0.0                     # the broken variant
                want_ug = np.concatenate([contribution('ug', s, r, k, mb) for s in range(G)])
                want_igp = sum(contribution('igp', s, r, k, S) for s in range(G))
                if not (np.allclose(me.ug, want_ug) and np.allclose(me.igp, want_igp)):
                    errors.append(('bwd', r, k))
                nap(r, k, 5)
        except threading.BrokenBarrierError:
            errors.append(('barrier', r, -1))

    th = [threading.Thread(target=body, args=(r,)) for r in range(G)]
    for t in th:
        t.start()
    for t in th:
        t.join(timeout=120)
    return errors


@pytest.mark.parametrize('G,seed', [(2, 0), (4, 1), (8, 2)])
def test_two_barriers_and_clear_before_arrival_never_lose_or_mix_a_push(G, seed):
    assert run(G, 40, 16, 8, seed=seed) == []


def test_a_slow_rank_does_not_break_the_protocol():
    # rank 1 is late everywhere, rank 0 races ahead as far as the barriers allow
    assert run(2, 20, 16, 8, delays={(1, 0): 2e-3, (1, 2): 2e-3, (1, 5): 2e-3}) == []


def test_clearing_behind_the_barrier_is_caught():
    """Negative control: rank 0 clears its forward blocks AFTER B2 and is slow to do so, rank 1 is already pushing the
    next step's partial sums into them — the model must notice the lost contribution."""
    errs = run(2, 6, 16, 8, clear_after_barrier=True, delays={(0, 4): 5e-3})
    assert any(e[0] == 'fwd' for e in errs), errs
