"""Multi-GPU plumbing for the hot path: one process per GPU under torchrun.

  Hessian  : calibration samples are partitioned across ranks (sample i -> rank i mod P); every rank
             accumulates H over its samples with the tcgen05 kernel, then ONE NCCL all-reduce(sum) of H
             (and of the token count) makes H bitwise identical everywhere.
  inverse  : computed on every rank from the identical H (replicated; no exchange).
  sweep    : rows of W are partitioned in contiguous slabs; ITF/AGA/feedback are row-local
             (quantizer.py reductions are all dim=1; gptq.py:186 is a right-multiplication), H^-1 is
             replicated.  With SSR the per-block column statistics are all-reduced (2*rem+1 floats)
             so every rank selects the same block; sequential / act-order need no communication.
  outputs  : each rank keeps its row slab of (alpha, mu, T); ``gather=True`` all-gathers them.

torch.distributed is plumbing only: the collectives are NCCL over NVLink, the math is libtq100.
"""

from typing import List, Optional, Tuple

import torch
import torch.distributed as dist

try:
    from . import _lib
    from .gptq import GPTQ, HessianState
    from .pipeline import LinearView  # noqa: F401
except ImportError:
    import _lib
    from gptq import GPTQ, HessianState
    from pipeline import LinearView  # noqa: F401


class ShardContext:
    def __init__(self, rank: int = 0, world: int = 1, device=None, group=None):
        self.rank, self.world, self.device, self.group = rank, world, device, group

    def my_samples(self, num_samples: int) -> List[int]:
        return list(range(self.rank, num_samples, self.world))

    def row_range(self, n: int) -> Tuple[int, int]:
        """Contiguous slab [lo, hi) of this rank; slab sizes differ by at most one 4-row group."""
        groups = (n + 3) // 4
        lo_g = (groups * self.rank) // self.world
        hi_g = (groups * (self.rank + 1)) // self.world
        return min(n, lo_g * 4), min(n, hi_g * 4)


class ShardedGPTQ:
    """GPTQ over a ShardContext.  With world == 1 this is exactly ``GPTQ``."""

    def __init__(self, layer, ctx: ShardContext, block_size: int = 128, percdamp: float = 0.01,
                 hessian: Optional[HessianState] = None):
        self.ctx = ctx
        self.layer = layer
        self.block_size, self.percdamp = block_size, percdamp
        self.rows, self.columns = layer.weight.shape
        self.inner = GPTQ(layer, block_size, percdamp, hessian=hessian)
        self._reduced = False

    def add_batch(self, inp_local: torch.Tensor):
        """Accumulate over THIS rank's calibration samples."""
        self.inner.add_batch(inp_local)
        self._reduced = False

    def _allreduce_hessian(self):
        if self.ctx.world == 1 or self._reduced:
            return
        st = self.inner.state
        dist.all_reduce(st.H, op=dist.ReduceOp.SUM, group=self.ctx.group)
        n = torch.tensor([st.nsamples], dtype=torch.int64, device=st.H.device)
        dist.all_reduce(n, op=dist.ReduceOp.SUM, group=self.ctx.group)
        st.nsamples = int(n.item())
        st._cache.clear()
        self._reduced = True

    def quantize(self, use_ssr: bool = True, aga: str = "hessian", order: Optional[str] = None,
                 max_iter: int = 100, gather: bool = False):
        if self.ctx.world == 1:
            return self.inner.quantize(use_ssr=use_ssr, aga=aga, order=order, max_iter=max_iter)
        if (order or ("ssr" if use_ssr else "sequential")) == "ssr" and not _lib.comm_ready():
            raise RuntimeError("row-sharded SSR needs the in-library NCCL communicator: call sharded.init_comm(ctx) first")
        self._allreduce_hessian()
        lo, hi = self.ctx.row_range(self.rows)
        shard = GPTQ(LinearView(self.layer.weight.data[lo:hi]), self.block_size, self.percdamp,
                     hessian=self.inner.state)
        shard.comm = self.ctx
        alpha, mu, T, perm = shard.quantize(use_ssr=use_ssr, aga=aga, order=order, max_iter=max_iter)
        self.shard = shard
        if not gather:
            return alpha, mu, T, perm
        outs = []
        for t in (alpha, mu, shard.T_int8):
            parts = [None] * self.ctx.world
            sizes = [self.ctx.__class__(r, self.ctx.world).row_range(self.rows) for r in range(self.ctx.world)]
            parts = [torch.empty((b - a,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device) for a, b in sizes]
            dist.all_gather(parts, t.contiguous(), group=self.ctx.group)
            outs.append(torch.cat(parts, dim=0))
        return outs[0], outs[1], outs[2], perm
