"""tf.train.* subset: Adagrad / SGD / Adam optimizers with TF-1.0 update rules, and a Saver."""
import numpy as _np
import torch as _t

import tensorflow as _tf
from tensorflow import _ev, _op, _Operation


class _Optimizer(object):
    def __init__(self, learning_rate):
        self._lr = learning_rate
        self._slots = {}

    def compute_gradients(self, loss, var_list=None, **_kw):
        var_list = var_list or _tf.trainable_variables()
        return list(zip(_tf.gradients(loss, var_list), var_list))

    def minimize(self, loss, global_step=None, var_list=None, **_kw):
        return self.apply_gradients(self.compute_gradients(loss, var_list), global_step)

    def apply_gradients(self, grads_and_vars, global_step=None, name=None):
        gv = [(g, v) for g, v in grads_and_vars if g is not None]
        node = _Operation(None, [g for g, _ in gv], 'apply_gradients')
        opt = self

        def compute(ctx):
            lr = float(_ev(opt._lr, ctx))
            for g, v in gv:
                grad = _ev(g, ctx).detach()
                cur = v.value
                for var, val in ctx.staged:
                    if var is v:
                        cur = val
                ctx.staged.append((v, opt._update(v, cur.detach(), grad, lr, ctx)))
            if global_step is not None:
                ctx.staged.append((global_step, global_step.value + 1))
            return None
        node._compute = compute
        return node


class GradientDescentOptimizer(_Optimizer):
    def __init__(self, learning_rate, use_locking=False, name='GradientDescent'):
        _Optimizer.__init__(self, learning_rate)

    def _update(self, v, w, g, lr, ctx):
        return w - lr * g


class AdagradOptimizer(_Optimizer):
    """accum starts at initial_accumulator_value (0.1); accum += g*g; var -= lr * g / sqrt(accum).
    Rows whose gradient is zero are left untouched, as by TF's sparse apply on IndexedSlices."""

    def __init__(self, learning_rate, initial_accumulator_value=0.1, use_locking=False, name='Adagrad'):
        _Optimizer.__init__(self, learning_rate)
        self._init_acc = initial_accumulator_value

    def get_slot(self, var, name='accumulator'):
        if var not in self._slots:
            self._slots[var] = _t.full_like(var.value, self._init_acc)
        return self._slots[var]

    def _update(self, v, w, g, lr, ctx):
        acc = self.get_slot(v)
        new_acc = acc + g * g
        self._slots[v] = new_acc          # slots are private to the optimizer: commit immediately
        return w - lr * g / _t.sqrt(new_acc)


class AdamOptimizer(_Optimizer):
    def __init__(self, learning_rate=0.001, beta1=0.9, beta2=0.999, epsilon=1e-8, use_locking=False, name='Adam'):
        _Optimizer.__init__(self, learning_rate)
        self.b1, self.b2, self.eps, self.t = beta1, beta2, epsilon, 0
        self._step_seen = None

    def _update(self, v, w, g, lr, ctx):
        if self._step_seen is not ctx:
            self._step_seen = ctx
            self.t += 1
        m, s = self._slots.get(v, (_t.zeros_like(w), _t.zeros_like(w)))
        m = self.b1 * m + (1 - self.b1) * g
        s = self.b2 * s + (1 - self.b2) * g * g
        self._slots[v] = (m, s)
        lr_t = lr * (1 - self.b2 ** self.t) ** 0.5 / (1 - self.b1 ** self.t)
        return w - lr_t * m / (_t.sqrt(s) + self.eps)


class Saver(object):
    def __init__(self, var_list=None, **_kw):
        self.vars = list(var_list) if var_list is not None else _tf.global_variables()

    def save(self, sess, save_path, global_step=None, **_kw):
        _np.savez(save_path + '.npz', **{v._name.replace('/', '|'): v.value.numpy() for v in self.vars})
        return save_path

    def restore(self, sess, save_path):
        d = _np.load(save_path + '.npz')
        for v in self.vars:
            v.load(d[v._name.replace('/', '|')])


def get_checkpoint_state(*a, **k):
    return None
