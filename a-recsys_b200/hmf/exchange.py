"""Collectives of the row-sharded embedding path, device-agnostic so that the host
logic is testable with the gloo backend on CPU; on B200 the backend is NCCL over NVLink 5.

With G GPUs every table is row-sharded (row t on GPU t % G).  A step over a global batch of G*mb
interactions exchanges, per phase, ONE packed buffer per collective:
  forward : reduce-scatter  [G*mb, 2d+1]  (partial pooled user | target-item vectors | target bias)
            all-reduce      [S, d+1]      (partial pooled sampled-pool vectors | bias)
  backward: all-reduce      [S, d+1]      (d sampled-pool vectors | d bias, partial over batch rows)
            all-gather      [mb, 2d+1] -> [G*mb, 2d+1]  (d user | d target vectors | d target bias)
No table gradient ever crosses a link: each GPU updates the rows it owns.
"""
import ctypes
import os

import torch
import torch.distributed as dist


class RowShardExchange(object):
    def __init__(self, group=None):
        self.group = group
        self.G = dist.get_world_size(group)
        self.r = dist.get_rank(group)
        self.nccl = dist.get_backend(group) == 'nccl'

    def reduce_scatter_rows(self, x):
        """x [G*mb, w] (this rank's partial sums for ALL rows) -> [mb, w] (full sums of this rank's rows)."""
        mb = x.shape[0] // self.G
        if self.nccl:
            out = torch.empty((mb, x.shape[1]), dtype=x.dtype, device=x.device)
            dist.reduce_scatter_tensor(out, x.contiguous(), group=self.group)
            return out
        dist.all_reduce(x, group=self.group)                     # gloo has no reduce_scatter
        return x[self.r * mb:(self.r + 1) * mb].clone()

    def all_reduce(self, x):
        dist.all_reduce(x, group=self.group)
        return x

    def all_gather_rows(self, x):
        """x [mb, w] -> [G*mb, w] in rank order."""
        x = x.contiguous()
        out = torch.empty((self.G * x.shape[0],) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
        if self.nccl:
            dist.all_gather_into_tensor(out, x, group=self.group)
        else:
            parts = list(out.chunk(self.G, 0))
            dist.all_gather(parts, x, group=self.group)
        return out


class PeerUnavailable(RuntimeError):
    """Raised on EVERY rank when any rank could not set the peer mapping up."""


class _Raw(object):
    """__cuda_array_interface__ view of a device allocation that torch does not own."""

    def __init__(self, ptr, n_float):
        self.__cuda_array_interface__ = {'shape': (n_float,), 'typestr': '<f4', 'data': (int(ptr), False), 'version': 2}


class PeerExchange(object):
    """The same four exchanges WITHOUT collectives on the dependent chain: every rank owns receive blocks in one
    arx_peer_alloc allocation that all other ranks map through CUDA IPC (NVLink 5 / NVSwitch peer memory), and the
    step's own kernels write into them (csrc/peer.cu, csrc/pool.cu push mode):

      locU [mb, d], locP [mb, d], locb [mb]   <- red.add of the partial pooled user / target vectors / target bias, from
                                                 INSIDE the lookup kernel of every rank (lookup + reduce-scatter fused; no
                                                 [G*mb, 2d+1] staging buffer)
      spP [S, d], spb [S]                     <- red.add of every rank's partial pooled pool vectors / bias, from inside
                                                 the same lookup launch (lookup + all-reduce fused)
      ug [G*mb, d]                            <- 128-bit stores of every rank's dU rows                  (all-gather by push)
      ig [S + G*mb, d], igb [S + G*mb]        <- the item-side gradient ARENA of the scatter-Adagrad kernel, assembled in
                                                 place: pool rows added by every rank (all-reduce), target rows stored by
                                                 their owners (all-gather) — no concatenation on the consumer side

    and two device-side barriers per step order the pushes before their consumers (arx_peer_barrier: flags in the same
    allocation, bounded spin).  Zeroing protocol: a rank clears loc* / sp* after it consumed them and BEFORE it arrives at
    the step's second barrier (the peers' next pushes come after that barrier), and clears the pool part of ig / igb at
    the start of a step before the first barrier.  Sums over ranks are not ordered: results can differ in the last bit
    for G > 2."""

    ALIGN = 256

    def __init__(self, group, device, mb, S, d):
        from .. import _lib
        self._lib = _lib
        self.group = group
        self.G = dist.get_world_size(group)
        self.r = dist.get_rank(group)
        self.device = device
        self.mb, self.S, self.d = mb, S, d
        G = self.G
        # loc* and sp* are adjacent: ONE memset clears them
        sizes = [('locU', mb * d), ('locP', mb * d), ('locb', mb), ('spP', S * d), ('spb', S),
                 ('ug', G * mb * d), ('ig', (S + G * mb) * d), ('igb', S + G * mb), ('flags', 64)]
        self.off = {}
        o = 0
        for name, n in sizes:
            self.off[name] = o
            o += (n * 4 + self.ALIGN - 1) // self.ALIGN * self.ALIGN
        self.nbytes = o
        self.fwd_floats = self.off['ug'] // 4                      # locU .. spb inclusive (with alignment padding)
        lib = _lib.load()
        torch.cuda.set_device(device)
        # every failure is agreed on COLLECTIVELY (a rank that raised alone would leave the others in a collective):
        # stage 1 allocate + export, stage 2 map the peers; the caller falls back to the NCCL exchange on PeerUnavailable
        base = ctypes.c_void_p()
        handle = (ctypes.c_ubyte * 64)()
        why = None
        rc = lib.arx_peer_alloc(self.nbytes, ctypes.byref(base))
        if rc != 0:
            why = 'arx_peer_alloc(%d) failed: %d' % (self.nbytes, rc)
        else:
            rc = lib.arx_peer_export(ctypes.c_void_p(base.value), handle)
            if rc != 0:
                why = 'arx_peer_export failed: %d (CUDA IPC unavailable)' % rc
        self.base = base.value
        got = [None] * G
        dist.all_gather_object(got, (why, bytes(handle)), group=group)
        if any(w is not None for w, _ in got):
            if self.base:
                lib.arx_peer_free(ctypes.c_void_p(self.base))
            raise PeerUnavailable('; '.join('rank %d: %s' % (g, w) for g, (w, _) in enumerate(got) if w))
        self.bases = []
        for g in range(G):
            if g == self.r:
                self.bases.append(self.base)
                continue
            p = ctypes.c_void_p()
            h = (ctypes.c_ubyte * 64).from_buffer_copy(got[g][1])
            rc = lib.arx_peer_open(h, ctypes.byref(p))
            if rc != 0:
                why = 'arx_peer_open(rank %d) failed: %d (no peer access between the GPUs?)' % (g, rc)
                break
            self.bases.append(p.value)
        got = [None] * G
        dist.all_gather_object(got, why, group=group)
        if any(w is not None for w in got):
            for g, b in enumerate(self.bases):
                if g != self.r:
                    lib.arx_peer_close(ctypes.c_void_p(b))
            lib.arx_peer_free(ctypes.c_void_p(self.base))
            raise PeerUnavailable('; '.join('rank %d: %s' % (g, w) for g, w in enumerate(got) if w))
        # own blocks as torch tensors, every rank's blocks as device pointer arrays
        self.t = {name: torch.as_tensor(_Raw(self.base + self.off[name], n), device=device) for name, n in sizes}
        self.fwd_all = torch.as_tensor(_Raw(self.base + self.off['locU'], self.fwd_floats), device=device)
        self.locU, self.locP, self.locb = self.t['locU'].view(mb, d), self.t['locP'].view(mb, d), self.t['locb']
        self.spP, self.spb = self.t['spP'].view(S, d), self.t['spb']
        self.ug = self.t['ug'].view(G * mb, d)
        self.ig, self.igb = self.t['ig'].view(S + G * mb, d), self.t['igb']
        self.p = {name: torch.tensor([b + self.off[name] for b in self.bases], dtype=torch.int64, device=device)
                  for name, _ in sizes}
        self.epoch = torch.zeros(1, dtype=torch.int32, device=device)
        self.err = torch.zeros(1, dtype=torch.int32, device=device)
        self.timeout_ns = int(float(os.environ.get('ARX_PEER_TIMEOUT_S', '20')) * 1e9)
        torch.cuda.synchronize(device)
        dist.barrier(group=group)          # every rank has mapped every block before anyone pushes

    # ---- pieces of the step -----------------------------------------------------------------------------------
    def push_desc(self, which):
        """kwargs['push'] of EmbeddingAttribute.pool_many: (pointer array of the receive blocks, pointer array of the bias
        vectors or None, entities per owner rank (0 = every rank gets every entity), row pitch, ranks)."""
        if which == 'user':
            return (self.p['locU'], None, self.mb, self.d, self.G)
        if which == 'item':
            return (self.p['locP'], self.p['locb'], self.mb, self.d, self.G)
        return (self.p['spP'], self.p['spb'], 0, self.d, self.G)          # 'pool'

    def barrier(self):
        self._lib.call('arx_peer_barrier', self.p['flags'].data_ptr(), self.r, self.G, self.epoch.data_ptr(), self.timeout_ns,
                       self.err.data_ptr())

    def clear_forward_blocks(self):
        self.fwd_all.zero_()

    def clear_pool_gradients(self):
        self.ig[:self.S].zero_()
        self.igb[:self.S].zero_()

    def push_gradients(self, dU0, dPt, dts, dPs, dbs):
        """The backward exchange in ONE launch: this rank's dU / dPt / dts rows stored into rows r*mb.. of every rank's
        ug / ig / igb, its partial pool gradients dPs / dbs added into the pool rows of every rank's ig / igb.
        (A vector travels as a [n/4, 4] row block.)"""
        mb, S, d, r = self.mb, self.S, self.d, self.r
        segs = (self._lib.PeerSeg * 5)()

        def seg(k, src, dst, rows, width, src_stride, dst_stride, row0, mode):
            q = segs[k]
            q.src, q.dst, q.rows, q.width, q.src_stride, q.dst_stride, q.row0, q.mode = (
                src.data_ptr(), self.p[dst].data_ptr(), rows, width, src_stride, dst_stride, row0, mode)
        seg(0, dU0, 'ug', mb, d, dU0.stride(0), d, r * mb, 0)
        seg(1, dPt, 'ig', mb, d, dPt.stride(0), d, S + r * mb, 0)
        seg(2, dts, 'igb', mb // 4, 4, 4, 4, (S + r * mb) // 4, 0)
        seg(3, dPs, 'ig', S, d, dPs.stride(0), d, 0, 1)
        seg(4, dbs, 'igb', S // 4, 4, 4, 4, 0, 1)
        self._lib.call('arx_peer_push_many', ctypes.addressof(segs), 5, self.G)

    def check(self):
        """Host-side check of the barrier watchdog (call outside the timed region)."""
        e = int(self.err.item())
        if e != 0:
            raise RuntimeError('arx_peer_barrier timed out on rank %d waiting for rank %d' % (self.r, e - 1))

    def close(self):
        lib = self._lib.load()
        torch.cuda.synchronize(self.device)
        for g, b in enumerate(self.bases):
            if g != self.r:
                lib.arx_peer_close(ctypes.c_void_p(b))
        self.t = {}
        lib.arx_peer_free(ctypes.c_void_p(self.base))
        self.bases = []
