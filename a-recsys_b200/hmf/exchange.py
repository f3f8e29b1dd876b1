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


class _Raw(object):
    """__cuda_array_interface__ view of a device allocation that torch does not own."""

    def __init__(self, ptr, n_float):
        self.__cuda_array_interface__ = {'shape': (n_float,), 'typestr': '<f4', 'data': (int(ptr), False), 'version': 2}


class PeerExchange(object):
    """The same four exchanges WITHOUT collectives on the dependent chain: every rank owns receive blocks in one
    arx_peer_alloc allocation that all other ranks map through CUDA IPC (NVLink 5 / NVSwitch peer memory), and the
    step's own kernels write into them (csrc/peer.cu, csrc/pool.cu push mode):

      loc  [mb, 2d+4]    <- red.add of partial pooled user | target vectors | target bias, from INSIDE the lookup kernel
                            of every rank (replaces reduce-scatter and its [G*mb, 2d+4] staging buffer)
      sp   [S, d+4]      <- red.add of every rank's partial pooled pool vectors | bias        (replaces all-reduce)
      back [G*mb, 2d+4]  <- 128-bit stores of every rank's gradient rows dU | dPt | dts        (replaces all-gather)
      dsp  [S, d+4]      <- red.add of every rank's partial pool gradients                    (replaces all-reduce)

    and two device-side barriers per step order the pushes before their consumers (arx_peer_barrier: flags in the same
    allocation, bounded spin).  Zeroing protocol: a rank clears loc / sp after it consumed them and BEFORE it arrives at
    the step's second barrier (the peers' next pushes come after that barrier), and clears dsp at the start of a step
    before the first barrier.  Sums over ranks are not ordered: results can differ in the last bit for G > 2."""

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
        self.W = 2 * d + 4
        self.Wp = d + 4
        sizes = [('loc', mb * self.W), ('sp', S * self.Wp), ('back', G * mb * self.W), ('dsp', S * self.Wp), ('flags', 64)]
        self.off = {}
        o = 0
        for name, n in sizes:
            self.off[name] = o
            o += (n * 4 + self.ALIGN - 1) // self.ALIGN * self.ALIGN
        self.nbytes = o
        lib = _lib.load()
        base = ctypes.c_void_p()
        torch.cuda.set_device(device)
        rc = lib.arx_peer_alloc(self.nbytes, ctypes.byref(base))
        if rc != 0:
            raise RuntimeError('arx_peer_alloc(%d) failed: %d' % (self.nbytes, rc))
        self.base = base.value
        handle = (ctypes.c_ubyte * 64)()
        rc = lib.arx_peer_export(ctypes.c_void_p(self.base), handle)
        if rc != 0:
            raise RuntimeError('arx_peer_export failed: %d (CUDA IPC unavailable)' % rc)
        handles = [None] * G
        dist.all_gather_object(handles, bytes(handle), group=group)
        self.bases = []
        for g in range(G):
            if g == self.r:
                self.bases.append(self.base)
                continue
            p = ctypes.c_void_p()
            h = (ctypes.c_ubyte * 64).from_buffer_copy(handles[g])
            rc = lib.arx_peer_open(h, ctypes.byref(p))
            if rc != 0:
                raise RuntimeError('arx_peer_open(rank %d) failed: %d (no peer access between the GPUs?)' % (g, rc))
            self.bases.append(p.value)
        # own blocks as torch tensors, every rank's blocks as device pointer arrays
        self.t = {}
        for name, n in sizes:
            self.t[name] = torch.as_tensor(_Raw(self.base + self.off[name], n), device=device)
        self.loc = self.t['loc'].view(mb, self.W)
        self.sp = self.t['sp'].view(S, self.Wp)
        self.back = self.t['back'].view(G * mb, self.W)
        self.dsp = self.t['dsp'].view(S, self.Wp)

        def ptrs(name, extra_bytes=0):
            return torch.tensor([b + self.off[name] + extra_bytes for b in self.bases], dtype=torch.int64, device=device)
        self.p_loc_user = ptrs('loc')
        self.p_loc_item = ptrs('loc', 4 * d)
        self.p_sp = ptrs('sp')
        self.p_back = ptrs('back')
        self.p_dsp = ptrs('dsp')
        self.p_flags = ptrs('flags')
        self.epoch = torch.zeros(1, dtype=torch.int32, device=device)
        self.err = torch.zeros(1, dtype=torch.int32, device=device)
        self.timeout_ns = int(float(os.environ.get('ARX_PEER_TIMEOUT_S', '20')) * 1e9)
        torch.cuda.synchronize(device)
        dist.barrier(group=group)          # every rank has mapped every block before anyone pushes

    # ---- pieces of the step -----------------------------------------------------------------------------------
    def push_desc(self, which):
        """kwargs['push'] of EmbeddingAttribute.pool_many: (device pointer array, rows per rank, row pitch, bias column)."""
        if which == 'user':
            return (self.p_loc_user, self.mb, self.W, -1)
        return (self.p_loc_item, self.mb, self.W, self.d)        # bias column 2d of the row = column d behind the item block

    def barrier(self):
        self._lib.call('arx_peer_barrier', self.p_flags.data_ptr(), self.r, self.G, self.epoch.data_ptr(), self.timeout_ns,
                       self.err.data_ptr())

    def add_to_all(self, src, which):
        """src [S, d+4] partial -> += into block `which` ('sp' | 'dsp') of every rank."""
        p = self.p_sp if which == 'sp' else self.p_dsp
        self._lib.call('arx_peer_push_rows', src.data_ptr(), src.shape[0], src.shape[1], src.stride(0), p.data_ptr(), self.Wp, 0,
                       self.G, -1, 1)

    def gather_rows(self, mine):
        """mine [mb, 2d+4] -> rows [r*mb, (r+1)*mb) of `back` on every rank (this one included)."""
        self._lib.call('arx_peer_push_rows', mine.data_ptr(), mine.shape[0], mine.shape[1], mine.stride(0), self.p_back.data_ptr(),
                       self.W, self.r * self.mb, self.G, -1, 0)

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
