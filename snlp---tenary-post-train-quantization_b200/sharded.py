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
    from .pipeline import LayerDriver, LinearView, chain_cost  # noqa: F401
except ImportError:
    import _lib
    from gptq import GPTQ, HessianState
    from pipeline import LayerDriver, LinearView, chain_cost  # noqa: F401


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


def init_comm(ctx: ShardContext):
    """Create the in-library NCCL communicator (one per process): rank 0 makes the id, torch.distributed
    ships the 128 bytes, every rank joins on its own device."""
    import ctypes
    lib = _lib.load()
    if ctx.world == 1 or lib.tq_comm_ready():
        return
    buf = (ctypes.c_ubyte * 128)()
    if ctx.rank == 0:
        _lib.check(lib.tq_comm_unique_id(buf), "tq_comm_unique_id")
    t = torch.tensor(list(buf), dtype=torch.uint8, device=ctx.device)
    dist.broadcast(t, src=0, group=ctx.group)
    raw = bytes(t.cpu().tolist())
    cbuf = (ctypes.c_ubyte * 128).from_buffer_copy(raw)
    with torch.cuda.device(ctx.device):
        _lib.check(lib.tq_comm_init(cbuf, ctx.rank, ctx.world), "tq_comm_init")
    init_p2p(ctx)


def init_p2p(ctx: ShardContext, cap_floats: int = 2 * 32768 + 16):
    """Open the peer-memory mailboxes of the one-shot all-reduce (csrc/comm.cu): every rank allocates its buffer, the
    CUDA IPC handles are all-gathered, every rank maps the others'.  Falls back to NCCL (returns False) when the
    ranks cannot map each other's memory (TQ_P2P_ALLREDUCE=0 forces that)."""
    import ctypes
    import os
    lib = _lib.load()
    if ctx.world == 1 or lib.tq_comm_p2p_ready():
        return bool(lib.tq_comm_p2p_ready())
    if os.environ.get("TQ_P2P_ALLREDUCE", "1") == "0" or ctx.world > 16:
        return False
    handle = (ctypes.c_ubyte * 64)()
    with torch.cuda.device(ctx.device):
        ok = lib.tq_comm_p2p_alloc(int(cap_floats), ctx.rank, ctx.world, handle) == 0
    mine = torch.tensor(list(handle) + [1 if ok else 0], dtype=torch.uint8, device=ctx.device)
    every = [torch.empty_like(mine) for _ in range(ctx.world)]
    dist.all_gather(every, mine, group=ctx.group)
    every = [t.cpu() for t in every]
    if not all(int(t[64]) == 1 for t in every):
        return False
    raw = b"".join(bytes(t[:64].tolist()) for t in every)
    buf = (ctypes.c_ubyte * len(raw)).from_buffer_copy(raw)
    with torch.cuda.device(ctx.device):
        opened = lib.tq_comm_p2p_open(buf) == 0
    flag = torch.tensor([1 if opened else 0], dtype=torch.int32, device=ctx.device)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=ctx.group)
    if int(flag.item()) != 1:
        raise RuntimeError("peer mailboxes opened on some ranks only: set TQ_P2P_ALLREDUCE=0 to use NCCL for the exchange")
    return True


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
        shard.sweep_flags = _lib.SWEEP_ROW_SHARD
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


def _shape_of(w):
    return tuple(w.shape) if torch.is_tensor(w) else tuple(w)


def deal_linears(shapes, world: int) -> List[int]:
    """Owner rank of every linear for the linear-parallel mode: longest chain first onto the least loaded rank
    (deterministic, identical on every rank).  shapes: [(n, m)]."""
    load = [0.0] * world
    owner = [0] * len(shapes)
    for i in sorted(range(len(shapes)), key=lambda i: (-chain_cost(*shapes[i]), i)):
        r = min(range(world), key=lambda r: (load[r], r))
        owner[i] = r
        load[r] += chain_cost(*shapes[i])
    return owner


class ShardedLayer:
    """One transformer layer's linears over a ShardContext (the multi-GPU form of the per-layer loop of
    main.py:289-299).  Every rank accumulates each linear's H over ITS calibration samples (tcgen05 SYRK); then

    mode 'linears' (default): the linears are independent objects, so they are dealt to the ranks (deal_linears).
      One NCCL reduce per H onto its owner; the owner runs the whole prologue + sweep chain of its linears on its
      own streams -- no collective inside the sweep, no H^-1 exchange; results stay on the owner.  A linear whose
      chain alone exceeds split_threshold x a rank's fair share (the 11008-wide down_proj from 4 GPUs up) is split:
      its H is all-reduced, its owner inverts and broadcasts H^-1, and every rank sweeps its row slab of it.
    mode 'rows': one NCCL all-reduce per H; the damped inverses are dealt round-robin and broadcast; every rank
      sweeps its contiguous row slab of every linear; with SSR the per-block column statistics are all-reduced
      inside the C driver loop (in-library NCCL communicator).  Results are row slabs on every rank."""

    def __init__(self, ctx: ShardContext, block_size: int = 128, percdamp: float = 0.01, mode: str = "linears",
                 num_streams: int = 4, split_threshold: Optional[float] = 1.5):
        if mode not in ("linears", "rows"):
            raise ValueError("mode must be 'linears' or 'rows'")
        self.ctx, self.block_size, self.percdamp, self.mode = ctx, block_size, percdamp, mode
        self.split_threshold = split_threshold
        self._driver = LayerDriver(ctx.device, block_size, percdamp, num_streams=num_streams) if mode == "linears" else None

    def owners(self, shapes) -> List[int]:
        """Mode 'linears': owner rank per linear ((n, m) shapes in layer order).  A linear whose chain is much longer
        than a rank's fair share (chain_cost > split_threshold x total / world) is *split*: its owner (returned as
        ``-1 - rank``, i.e. negative) computes the damped inverse and broadcasts it, then every rank sweeps its row slab
        of that linear (as mode 'rows' does); W and the results of a split linear are row slabs on every rank."""
        shapes = [tuple(s) for s in shapes]
        owner = deal_linears(shapes, self.ctx.world)
        if self.ctx.world > 1 and self.split_threshold is not None:
            fair = sum(chain_cost(*s) for s in shapes) / self.ctx.world
            owner = [(-1 - o) if chain_cost(*s) > self.split_threshold * fair else o for o, s in zip(owner, shapes)]
        return owner

    def _quantize_by_linear(self, linears, use_ssr, aga, max_iter, hess_timing, order=None, phases=None):
        ctx = self.ctx
        use_ssr = (order == "ssr") if order is not None else use_ssr
        main = torch.cuda.current_stream(ctx.device)

        def mark(tag, stream=None):
            # critical-path breakdown (bench): timing events on the stream a phase ends on
            if phases is not None:
                ev = torch.cuda.Event(enable_timing=True)
                ev.record(stream if stream is not None else main)
                phases.append((tag, ev))
        mark("start")
        owner = self.owners([_shape_of(W) for _, W, _ in linears])
        split = [i for i, o in enumerate(owner) if o < 0]
        if split and use_ssr and not _lib.comm_ready():
            init_comm(ctx)                                   # row-sharded SSR all-reduces inside the C sweep loop
        # Token counts first (one tiny all-reduce while the GPU is idle: no host wait behind the Hessian kernels later).
        tokens = [X.numel() // X.shape[-1] for _, _, X in linears]
        if ctx.world > 1:
            counts = torch.tensor(tokens, dtype=torch.int64, device=ctx.device)
            dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=ctx.group)
            tokens = counts.tolist()
        # Hessians widest first, each handed to NCCL the moment its kernel is enqueued (async: the caller's stream goes on
        # with the next Hessian while the reduction of the previous one crosses NVLink).  The widest input belongs to the
        # longest chain -- from 4 GPUs up the row-split down_proj, whose inverse is the layer's critical path -- so its
        # owner can start inverting while the remaining Hessians are still being accumulated.
        if not hasattr(self, "_join_stream"):
            self._join_stream = torch.cuda.Stream(ctx.device, priority=-1)
        js = self._join_stream
        states, ready = [None] * len(linears), [None] * len(linears)
        for i in sorted(range(len(linears)), key=lambda i: (-linears[i][2].shape[-1], i)):
            _, W, X = linears[i]
            st = HessianState(X.shape[-1], ctx.device)
            if hess_timing is not None:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
            st.add_batch(X)
            if hess_timing is not None:
                e1.record()
                hess_timing.append((e0, e1, X.numel() // X.shape[-1], X.shape[-1]))
            ev = torch.cuda.Event()
            if ctx.world > 1:
                if owner[i] < 0:
                    work = dist.all_reduce(st.H, op=dist.ReduceOp.SUM, group=ctx.group, async_op=True)
                else:
                    work = dist.reduce(st.H, dst=owner[i], op=dist.ReduceOp.SUM, group=ctx.group, async_op=True)
                with torch.cuda.stream(js):
                    work.wait()                                  # js waits for NCCL's stream, the host does not
                    ev.record(js)
                st.H.record_stream(js)
            else:
                ev.record(main)
            st.nsamples = int(tokens[i])
            st._cache.clear()
            states[i], ready[i] = st, ev
        mark("hessians")
        mark("reduce", js if ctx.world > 1 else main)
        # split linears first on their owner: the inverse is the longest single dependent chain of the layer
        pending = {}
        if split:
            if not hasattr(self, "_inv_stream"):
                self._inv_stream = torch.cuda.Stream(ctx.device, priority=-3)
            for i in split:
                st, m = states[i], states[i].columns
                if -1 - owner[i] == ctx.rank:
                    self._inv_stream.wait_event(ready[i])        # this Hessian's all-reduce only, not the later Hessians
                    with torch.cuda.stream(self._inv_stream):
                        Hd, Hinv, info = st.damped_inverse(self.percdamp)
                        done = torch.cuda.Event()
                        done.record(self._inv_stream)
                        mark("split_inverse", self._inv_stream)
                else:
                    main.wait_event(ready[i])
                    Hd = st.damped(self.percdamp)
                    Hinv = torch.empty((m, m), dtype=torch.float32, device=ctx.device)
                    info = torch.zeros(1, dtype=torch.int32, device=ctx.device)
                    done = None
                pending[i] = (Hd, Hinv, info, done)
        # this rank's whole linears: chains on the side streams, asynchronous
        mine = [i for i in range(len(linears)) if owner[i] == ctx.rank]
        gs = []
        for i in mine:
            W = linears[i][1]
            if not torch.is_tensor(W):
                raise ValueError(f"{linears[i][0]}: this rank owns the linear and needs its weight, not just the shape")
            gs.append(GPTQ(LinearView(W), self.block_size, self.percdamp, hessian=states[i]))
        chain_seq = self._driver.enqueue_chains(gs, use_ssr=use_ssr, aga=aga, max_iter=max_iter, order=order,
                                                ready=[ready[i] for i in mine])
        # split linears: inverse broadcast, then every rank sweeps its row slab (statistics all-reduced per block)
        slabs = {}
        for i in split:
            name, W, _ = linears[i]
            if not torch.is_tensor(W):
                raise ValueError(f"{name}: a split linear needs (this rank's row slab of) its weight on every rank")
            st = states[i]
            Hd, Hinv, info, done = pending[i]
            if done is not None:
                main.wait_event(done)
                for t in (Hd, Hinv, info):
                    t.record_stream(main)
            src = -1 - owner[i]
            dist.broadcast(Hinv, src=src, group=ctx.group)
            dist.broadcast(info, src=src, group=ctx.group)
            mark("split_bcast")
            st._cache[float(self.percdamp)] = (Hd, Hinv, info)
            st._cache_events.pop(float(self.percdamp), None)
            lo, hi = ctx.row_range(W.shape[0])
            g = GPTQ(LinearView(W[lo:hi]), self.block_size, self.percdamp, hessian=st)
            g.sweep_flags = _lib.SWEEP_ROW_SHARD
            g.quantize(use_ssr=use_ssr, aga=aga, max_iter=max_iter, order=order)
            mark("split_sweep")
            slabs[i] = (g, (lo, hi))
        self._driver.finish_chains(gs, chain_seq)
        mark("own_chains")
        out = []
        done_whole = dict(zip(mine, gs))
        for i, (name, W, _) in enumerate(linears):
            if i in slabs:
                g, rows = slabs[i]
                out.append((name, g.alpha, g.mu, g.T_int8, g.perm, rows))
            elif i in done_whole:
                g = done_whole[i]
                out.append((name, g.alpha, g.mu, g.T_int8, g.perm, (0, g.rows)))
            else:
                out.append((name, None, None, None, None, (0, 0)))
        return out

    def quantize(self, linears, use_ssr: bool = True, aga: str = "hessian", max_iter: int = 100, hess_timing=None,
                 order: Optional[str] = None, phases=None):
        """linears: [(name, W (n, m), X_local (this rank's samples, (.., m)))].  ``order`` as in GPTQ.quantize.
        phases (mode 'linears'): optional list that receives (tag, timing event) marks of the layer's critical path --
        start, hessians, reduce, split_inverse (owner only), split_bcast, split_sweep, own_chains.
        mode 'linears': W is read only on the linear's owner (see owners()); other ranks may pass its (n, m) shape instead;
          returns [(name, alpha, mu, T_int8, perm, (0, n))] with None tensors and rows (0, 0) for linears owned elsewhere;
          a *split* linear (owners() < 0) needs W on every rank (only this rank's row slab is read) and returns row slabs.
        mode 'rows': W replicated (only this rank's row slab is read); returns
          [(name, alpha_slab, mu_slab, T_int8_slab, perm, (row_lo, row_hi))] in the input order.
        hess_timing: optional list that receives (start_event, end_event, tokens, m) per Hessian launch."""
        if self.mode == "linears":
            return self._quantize_by_linear(linears, use_ssr, aga, max_iter, hess_timing, order, phases)
        ctx = self.ctx
        sweep_order = order
        main = torch.cuda.current_stream(ctx.device)
        states = []
        for _, W, X in linears:
            st = HessianState(W.shape[1], W.device)
            if hess_timing is not None:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
            st.add_batch(X)
            if hess_timing is not None:
                e1.record()
                hess_timing.append((e0, e1, X.numel() // X.shape[-1], W.shape[1]))
            states.append(st)
        # cheapest inverses first: their sweeps then run while the widest inverse is still being computed on its owner
        order = sorted(range(len(linears)), key=lambda i: states[i].columns)
        pending = {}
        if ctx.world > 1:
            counts = torch.tensor([st.nsamples for st in states], dtype=torch.int64, device=ctx.device)
            dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=ctx.group)
            for st, c in zip(states, counts.tolist()):
                dist.all_reduce(st.H, op=dist.ReduceOp.SUM, group=ctx.group)
                st.nsamples = int(c)
                st._cache.clear()
            # inverses are dealt to the ranks and computed on a side stream of their owner
            if not hasattr(self, "_inv_stream"):
                self._inv_stream = torch.cuda.Stream(ctx.device)
            self._inv_stream.wait_stream(main)
            for slot, i in enumerate(order):
                st = states[i]
                owner = slot % ctx.world
                m = st.columns
                if owner == ctx.rank:
                    with torch.cuda.stream(self._inv_stream):
                        Hd, Hinv, info = st.damped_inverse(self.percdamp)
                        done = torch.cuda.Event()
                        done.record(self._inv_stream)
                else:
                    Hd = st.damped(self.percdamp)
                    Hinv = torch.empty((m, m), dtype=torch.float32, device=ctx.device)
                    info = torch.zeros(1, dtype=torch.int32, device=ctx.device)
                    done = None
                pending[i] = (owner, Hd, Hinv, info, done)
        out = [None] * len(linears)
        for i in order:
            name, W, _ = linears[i]
            st = states[i]
            if ctx.world > 1:
                owner, Hd, Hinv, info, done = pending[i]
                if done is not None:
                    main.wait_event(done)                       # only this inverse, not the owner's later ones
                    for t in (Hd, Hinv, info):
                        t.record_stream(main)
                dist.broadcast(Hinv, src=owner, group=ctx.group)
                dist.broadcast(info, src=owner, group=ctx.group)
                st._cache[float(self.percdamp)] = (Hd, Hinv, info)
                st._cache_events.pop(float(self.percdamp), None)
            lo, hi = ctx.row_range(W.shape[0])
            g = GPTQ(LinearView(W[lo:hi]), self.block_size, self.percdamp, hessian=st)
            if ctx.world > 1:
                g.sweep_flags = _lib.SWEEP_ROW_SHARD
            alpha, mu, _, perm = g.quantize(use_ssr=use_ssr, aga=aga, max_iter=max_iter, order=sweep_order)
            out[i] = (name, alpha, mu, g.T_int8, perm, (lo, hi))
        return out


class ShardedHostPipeline:
    """ShardedLayer fed from pinned HOST memory, one rank's view: per transformer layer this rank's calibration
    samples and its part of the weights (mode 'linears': the whole weight of the linears it owns; mode 'rows': its
    row slab of every weight) are copied host->device on a side stream, the layer is quantized
    (ShardedLayer.quantize), and the rank's alpha / mu / T (int8) / perm go back to pinned host memory on a second
    side stream.  The copies of layer l+1 are enqueued before the kernels of layer l, so they overlap."""

    def __init__(self, ctx: ShardContext, block_size: int = 128, percdamp: float = 0.01, use_ssr: bool = True,
                 aga: str = "hessian", depth: int = 2, mode: str = "linears", order: Optional[str] = None):
        self.ctx = ctx
        self.layer = ShardedLayer(ctx, block_size, percdamp, mode=mode)
        self.use_ssr, self.aga, self.order = use_ssr, aga, order
        self.depth = max(1, int(depth))
        self.copy_stream = torch.cuda.Stream(ctx.device)
        self.out_stream = torch.cuda.Stream(ctx.device)
        self.h2d_bytes = 0
        self.d2h_bytes = 0
        self._bufs = {}
        self._free_events = [None] * self.depth

    def _buf(self, key, shape, dtype):
        need = 1
        for s in shape:
            need *= s
        buf = self._bufs.get(key)
        if buf is None or buf.numel() < need or buf.dtype != dtype:
            buf = torch.empty(need, dtype=dtype, device=self.ctx.device)
            self._bufs[key] = buf
        return buf[:need].view(shape)

    def _stage(self, li, layer):
        inputs, lins = layer
        slot = li % self.depth
        with torch.cuda.stream(self.copy_stream):
            if self._free_events[slot] is not None:
                self.copy_stream.wait_event(self._free_events[slot])
            x_dev = {}
            for key, xh in inputs.items():
                xd = self._buf(("x", key, slot), tuple(xh.shape), xh.dtype)
                xd.copy_(xh, non_blocking=True)
                self.h2d_bytes += xh.numel() * xh.element_size()
                x_dev[key] = xd
            w_dev = []
            by_linear = self.layer.mode == "linears"
            owner = self.layer.owners([(n, inputs[key].shape[-1]) for _, _, n, key in lins]) if by_linear else None
            for i, (name, w_host, n, key) in enumerate(lins):
                m = inputs[key].shape[-1]
                if by_linear and owner[i] >= 0:
                    if owner[i] != self.ctx.rank:
                        w_dev.append((n, m))                               # owned elsewhere: only the shape is needed
                        continue
                    lo, hi = 0, n
                else:                                                      # mode 'rows', or a split linear
                    lo, hi = self.ctx.row_range(n)
                if w_host is None or tuple(w_host.shape) != (hi - lo, m):
                    raise ValueError(f"{name}: expected rows [{lo}, {hi}) of the weight on this rank")
                wd = self._buf(("w", i, slot), (n, m), w_host.dtype)       # mode 'rows': only the slab's rows are read
                wd[lo:hi].copy_(w_host, non_blocking=True)
                self.h2d_bytes += w_host.numel() * w_host.element_size()
                w_dev.append(wd)
            ready = torch.cuda.Event()
            ready.record(self.copy_stream)
        return x_dev, w_dev, ready

    def run_iter(self, layers):
        """layers: iterable of (inputs, linears); inputs = {key: X_host (this rank's samples, pinned, (.., m))},
        linears = [(name, W_host (pinned fp32: this rank's rows -- the whole weight if it owns the linear, None if
        another rank does, its row slab if the linear is split, see ShardedLayer.owners; its row slab in mode 'rows'),
        n (rows of the whole weight), key)].
        Yields, per layer, [{'name', 'rows': (lo, hi), and for rows held here 'alpha', 'mu', 'T' (int8), 'perm'}] in
        pinned host memory (the device->host copies may still be in flight: synchronize())."""
        compute = torch.cuda.current_stream(self.ctx.device)
        self._free_events = [None] * self.depth
        it = iter(layers)
        staged = []
        n_staged = 0

        def top_up(target):
            nonlocal n_staged
            while len(staged) < target:
                lay = next(it, None)
                if lay is None:
                    return
                staged.append((lay, self._stage(n_staged, lay)))
                n_staged += 1

        top_up(max(1, self.depth - 1))
        li = 0
        while staged:
            (inputs, lins), (x_dev, w_dev, ready) = staged.pop(0)
            compute.wait_event(ready)
            # ShardedLayer.quantize blocks the host (Cholesky status reads): stage the next layer's copies first.
            # With depth slots, depth - 1 layers are ahead; the slot a new layer lands in was freed on the host.
            top_up(self.depth - 1)
            out = self.layer.quantize([(name, wd, x_dev[key]) for (name, _, _, key), wd in zip(lins, w_dev)],
                                      use_ssr=self.use_ssr, aga=self.aga, order=self.order)
            done = torch.cuda.Event()
            done.record(compute)
            res = []
            with torch.cuda.stream(self.out_stream):
                self.out_stream.wait_event(done)
                for name, alpha, mu, T8, perm, rows in out:
                    d = {"name": name, "rows": rows}
                    if alpha is None:                                      # mode 'linears': owned by another rank
                        res.append(d)
                        continue
                    for k, t in (("alpha", alpha), ("mu", mu), ("T", T8), ("perm", perm)):
                        h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
                        h.copy_(t, non_blocking=True)
                        t.record_stream(self.out_stream)
                        self.d2h_bytes += t.numel() * t.element_size()
                        d[k] = h
                    res.append(d)
            ev = torch.cuda.Event()
            ev.record(compute)
            self._free_events[li % self.depth] = ev
            li += 1
            if not staged:
                top_up(1)
            yield res

    def synchronize(self):
        self.out_stream.synchronize()
        torch.cuda.current_stream(self.ctx.device).synchronize()
