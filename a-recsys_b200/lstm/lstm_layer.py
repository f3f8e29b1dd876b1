"""Single-layer LSTM over a padded batch of sequences (K8): forward and explicit backward through
libarx_b200.so.  TF-1.0 `LSTMCell(size, state_is_tuple=True)` wrapped in DropoutWrapper on the
input and on the output, statically unrolled from a zero state (lstm/seqModel.py:99-103,468,477).

  x-projection of all T steps : one tcgen05 contraction  [T*mb, d_in] x [d_in, 4H] (+ bias)
  recurrence                  : ONE persistent kernel for all T steps (arx_lstm_seq_fwd: W_h resident in shared
                                memory, h W_h on tcgen05, gates in the TMEM epilogue, h exchanged through
                                distributed shared memory) when H is 32 / 64 / 128; otherwise, and on the exact-fp32
                                parity path, per step  Z_t += h_{t-1} W_h  + the fused gate kernel
  backward                    : arx_lstm_seq_bwd (same structure, dh reduce-scattered inside the cluster) or per step
                                arx_lstm_gates_bwd and dh_{t-1} = dZ_t W_h^T; the weight, bias and input gradients
                                are three large contractions over T*mb rows
"""
import math
import os

import torch

from .. import _lib
from .._lib import call, ptr


class LSTMLayer(object):
    def __init__(self, d_in, H, device, gen=None, W=None, b=None, forget_bias=1.0):
        self.d_in, self.H, self.device = d_in, H, device
        self.forget_bias = forget_bias
        if W is None:
            lim = math.sqrt(6.0 / (d_in + H + 4 * H))
            W = (torch.rand((d_in + H, 4 * H), generator=gen) * 2 - 1) * lim
        self.W = torch.as_tensor(W, dtype=torch.float32).reshape(d_in + H, 4 * H).clone().to(device).contiguous()
        self.b = (torch.zeros(4 * H) if b is None else torch.as_tensor(b, dtype=torch.float32)).clone().to(device)
        self.dW = torch.zeros_like(self.W)
        self.db = torch.zeros_like(self.b)

    def parameters(self):
        return {'lstm_w': (self.W, self.dW), 'lstm_b': (self.b, self.db)}

    def _transposed(self):
        d_in, H = self.d_in, self.H
        WxT = torch.empty((4 * H, d_in), dtype=torch.float32, device=self.device)
        WhT = torch.empty((4 * H, H), dtype=torch.float32, device=self.device)
        rnd = 0 if _lib.exact_fp32 else 1          # tf32-nearest copies for the tensor-core path
        call('arx_transpose', self.W[:d_in].data_ptr(), d_in, 4 * H, WxT.data_ptr(), rnd)
        call('arx_transpose', self.W[d_in:].data_ptr(), H, 4 * H, WhT.data_ptr(), rnd)
        return WxT, WhT

    def forward(self, X, keep=1.0, in_mask=None, out_mask=None, out_dropout=True):
        """X [T, mb, d_in] -> outputs [T, mb, H] (after output dropout).  Masks are 0/1 tensors
        (generated when keep < 1 and none is injected)."""
        T, mb, d_in = X.shape
        H = self.H
        dev = self.device
        if keep != 1.0:
            if in_mask is None:
                in_mask = torch.floor(torch.rand_like(X) + keep)
            Xd = torch.empty_like(X)
            call('arx_scale_mask', X.data_ptr(), in_mask.data_ptr(), 1.0 / keep, X.numel(), Xd.data_ptr())
        else:
            Xd = X
        WxT, WhT = self._transposed()
        G = torch.empty((T, mb, 4 * H), dtype=torch.float32, device=dev)
        _lib.gemm(Xd.view(T * mb, d_in), WxT, G.view(T * mb, 4 * H), T * mb, 4 * H, d_in, 0, 1, self.b, b_ready=True)
        Hs = torch.zeros((T + 1, mb, H), dtype=torch.float32, device=dev)      # Hs[t] = h_{t-1}
        Cs = torch.zeros((T + 1, mb, H), dtype=torch.float32, device=dev)
        tc = not _lib.exact_fp32
        # tensor-core path: the gate kernel also emits the tf32-rounded h (the next step's A operand),
        # instead of one arx_round_tf32 launch per step
        Hr = torch.empty((2, mb, H), dtype=torch.float32, device=dev) if tc else None
        # the whole recurrence in one launch (tensor-core path, H in {32, 64, 128})
        self._seq = bool(tc and os.environ.get('ARX_LSTM_SEQ', '1') == '1' and call('arx_lstm_seq_fwd', G.data_ptr(), WhT.data_ptr(), Hs.data_ptr(), Cs.data_ptr(),
                                     T, mb, H, self.forget_bias) == 0)
        if not self._seq:
            for t in range(T):
                if t > 0:
                    _lib.gemm(Hr[t & 1] if tc else Hs[t], WhT, G[t], mb, 4 * H, H, 0, 1, None, 1.0, 1.0,
                              a_ready=tc, b_ready=True)
                call('arx_lstm_gates_fwd2', G[t].data_ptr(), Cs[t].data_ptr() if t > 0 else None,
                     Cs[t + 1].data_ptr(), Hs[t + 1].data_ptr(), Hr[(t + 1) & 1].data_ptr() if tc else None,
                     mb, H, self.forget_bias)
        out = Hs[1:]
        self._out_dropout = out_dropout
        if keep != 1.0 and out_dropout:
            if out_mask is None:
                out_mask = torch.floor(torch.rand_like(out) + keep)
            od = torch.empty_like(out)
            call('arx_scale_mask', out.data_ptr(), out_mask.data_ptr(), 1.0 / keep, out.numel(), od.data_ptr())
            out = od
        self._ctx = (Xd, G, Hs, Cs, keep, in_mask, out_mask)
        return out

    def backward(self, dOut):
        """dOut [T, mb, H] -> dX [T, mb, d_in]; fills self.dW / self.db."""
        Xd, G, Hs, Cs, keep, in_mask, out_mask = self._ctx
        T, mb, d_in = Xd.shape
        H = self.H
        dev = self.device
        if keep != 1.0 and self._out_dropout:
            dH = torch.empty_like(dOut)
            call('arx_scale_mask', dOut.data_ptr(), out_mask.data_ptr(), 1.0 / keep, dOut.numel(), dH.data_ptr())
        else:
            dH = dOut.contiguous()
        tc = not _lib.exact_fp32
        Wh = self.W[d_in:]                                   # [H, 4H] = the K-major B of dZ W_h^T
        if tc:                                               # rounded once, not once per step
            Wh = _lib.round_tf32(Wh.contiguous())
        done = False
        if getattr(self, '_seq', False):
            done = call('arx_lstm_seq_bwd', G.data_ptr(), Wh.data_ptr(), Cs.data_ptr(), dH.data_ptr(), T, mb, H) == 0
        if not done:
            dh_rec = torch.empty((mb, H), dtype=torch.float32, device=dev)
            dc = [torch.empty((mb, H), dtype=torch.float32, device=dev) for _ in range(2)]
            for t in range(T - 1, -1, -1):
                last = (t == T - 1)
                # dZ is written tf32-rounded on the tensor-core path: it only feeds contractions
                call('arx_lstm_gates_bwd2', G[t].data_ptr(), Cs[t].data_ptr() if t > 0 else None, Cs[t + 1].data_ptr(),
                     dH[t].data_ptr(), None if last else dh_rec.data_ptr(), None if last else dc[(t + 1) & 1].data_ptr(),
                     dc[t & 1].data_ptr(), mb, H, 1 if tc else 0)
                if t > 0:
                    _lib.gemm(G[t], Wh, dh_rec, mb, H, 4 * H, 0, 1, a_ready=tc, b_ready=tc)
        dZ = G.view(T * mb, 4 * H)
        dX = torch.empty((T * mb, d_in), dtype=torch.float32, device=dev)
        _lib.gemm(dZ, self.W[:d_in], dX, T * mb, d_in, 4 * H, 0, 1, a_ready=tc)
        _lib.gemm(Xd.view(T * mb, d_in), dZ, self.dW[:d_in], d_in, 4 * H, T * mb, 1, 0)
        _lib.gemm(Hs[:T].view(T * mb, H), dZ, self.dW[d_in:], H, 4 * H, T * mb, 1, 0)
        call('arx_colsum', dZ.data_ptr(), T * mb, 4 * H, 4 * H, self.db.data_ptr())
        dX = dX.view(T, mb, d_in)
        if keep != 1.0:
            dXd = torch.empty_like(dX)
            call('arx_scale_mask', dX.data_ptr(), in_mask.data_ptr(), 1.0 / keep, dX.numel(), dXd.data_ptr())
            dX = dXd
        return dX


class LSTMStack(object):
    """MultiRNNCell([DropoutWrapper(LSTMCell, input_keep_prob)] * num_layers) wrapped in DropoutWrapper(output_keep_prob)
    (lstm/seqModel.py:99-103): every layer drops its INPUT with its own mask, the stack drops its output once.
    Variables of layer l > 0 are named lstm_w_<l> / lstm_b_<l> (TF: rnn/multi_rnn_cell/cell_<l>/lstm_cell/...)."""

    def __init__(self, d_in, H, num_layers, device, gen=None, params=None, forget_bias=1.0):
        p = params or {}
        self.layers = []
        for l in range(num_layers):
            sfx = '_%d' % l if l else ''
            self.layers.append(LSTMLayer(d_in if l == 0 else H, H, device, gen, p.get('lstm_w' + sfx), p.get('lstm_b' + sfx),
                                         forget_bias))
        self.H = H

    def parameters(self):
        d = {}
        for l, layer in enumerate(self.layers):
            sfx = '_%d' % l if l else ''
            d['lstm_w' + sfx] = (layer.W, layer.dW)
            d['lstm_b' + sfx] = (layer.b, layer.db)
        return d

    def forward(self, X, keep=1.0, in_mask=None, out_mask=None, in_masks_more=()):
        L = len(self.layers)
        for l, layer in enumerate(self.layers):
            im = in_mask if l == 0 else (in_masks_more[l - 1] if l - 1 < len(in_masks_more) else None)
            X = layer.forward(X, keep, im, out_mask if l == L - 1 else None, out_dropout=(l == L - 1))
        return X

    def backward(self, dOut):
        for layer in reversed(self.layers):
            dOut = layer.backward(dOut)
        return dOut
