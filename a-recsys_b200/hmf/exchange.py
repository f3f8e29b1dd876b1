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
