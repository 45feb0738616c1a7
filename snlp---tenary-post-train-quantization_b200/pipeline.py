"""Host-resident driver of the hot path: weights and calibration activations start in pinned HOST
memory, results end in pinned HOST memory (what ``main.py:225-230`` returns with ``.cpu()``).

Copies ride a side stream and are double-buffered against the compute stream, so host->device
traffic of the next input overlaps Hessian + sweep of the current one.  This is the call path
``bench.py`` times as ``e2e``; the kernels are the same ones ``GPTQ`` launches.
"""

import collections
import os
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import torch

try:
    from .gptq import GPTQ, HessianState
except ImportError:
    from gptq import GPTQ, HessianState


class _Weight:
    """The two attributes GPTQ reads from an nn.Linear (gptq.py:42-47), without building a module."""

    def __init__(self, weight: torch.Tensor):
        self.weight = weight


class LinearView(_Weight):
    pass


class LayerDriver:
    """Device-resident driver for the linears of one transformer layer (the loop of main.py:289-299).

    The linears are independent once their Hessians exist, and each one's prologue + sweep is a long chain
    of small dependent kernels, so the chains of different linears are enqueued on different CUDA streams
    and overlap on the GPU; the Hessian GEMMs, which saturate the device, stay on the caller's stream."""

    def __init__(self, device, block_size: int = 128, percdamp: float = 0.01, num_streams: int = 4,
                 share_inputs: bool = False, eager_chains: Optional[bool] = None):
        self.device = torch.device(device)
        self.block_size, self.percdamp, self.share_inputs = block_size, percdamp, share_inputs
        num_streams = int(os.environ.get("TQ_CHAIN_STREAMS", num_streams))
        # the stream that gets the longest chain outranks the others: its small dependent kernels are placed first
        # whenever SMs free up (TQ_CHAIN_PRIORITY=0 gives every stream the default priority)
        # CUDA priorities: numerically lower = served first; the caller's stream (where the Hessian kernels run) is 0, the
        # lowest.  Chain streams sit above it -- their short dependent kernels take SMs as Hessian CTAs retire -- and the
        # first stream, which gets the longest chain, sits above the other chains.
        prio = os.environ.get("TQ_CHAIN_PRIORITY", "1") != "0"
        self.streams = [torch.cuda.Stream(self.device, priority=((-3 if i == 0 else -1) if prio else 0))
                        for i in range(max(1, num_streams))]
        # eager: a linear's chain starts as soon as ITS Hessian is complete and overlaps the Hessians still running;
        # otherwise (default) all chains wait for all Hessians.  Measured on one B200 (8 layers of the 7B bench): eager
        # 118.9 ms per layer with the Hessian kernels at 1 095 TFLOP/s (the chains' wide kernels take SMs from them),
        # not eager 114.8 ms with the Hessians at 1 173 TFLOP/s -- the chain phase is bound by its own GEMM launches, not
        # by when it starts.
        self.eager = (os.environ.get("TQ_EAGER_CHAINS", "0") != "0") if eager_chains is None else bool(eager_chains)

    def quantize(self, linears, use_ssr: bool = True, aga: str = "hessian", max_iter: int = 100, hess_timing=None,
                 order: Optional[str] = None):
        """linears: [(name, W (n, m) CUDA fp32, X (.., m) CUDA activations)].  Linears handed the SAME X object
        share one Hessian when share_inputs is set.  ``order`` as in GPTQ.quantize ('ssr' | 'sequential' | 'actorder';
        None follows use_ssr).  Returns [GPTQ] in order, quantized."""
        main = torch.cuda.current_stream(self.device)
        shared = {}
        gs = [None] * len(linears)
        ready = [None] * len(linears)
        # Hessian order (TQ_HESS_ORDER): 'short-first' (default) accumulates the Hessians of the short chains first, so
        # those chains run underneath the remaining -- SM-filling -- Hessian kernels and the longest chain, which starts
        # last, has the GPU almost to itself; 'long-first' starts the longest chain as early as possible instead
        sign = 1 if os.environ.get("TQ_HESS_ORDER", "short-first") == "short-first" else -1
        seq = sorted(range(len(linears)), key=lambda i: (sign * chain_cost(*linears[i][1].shape), i))
        for i in seq:
            name, W, X = linears[i]
            st = shared.get(id(X)) if self.share_inputs else None
            g = GPTQ(LinearView(W), self.block_size, self.percdamp, hessian=st[0] if st else None)
            if st is None:
                if hess_timing is not None:
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                g.add_batch(X)
                if hess_timing is not None:
                    e1.record()
                    hess_timing.append((e0, e1, X.numel() // X.shape[-1], W.shape[1]))
                ev = torch.cuda.Event()
                ev.record(main)
                if self.share_inputs:
                    shared[id(X)] = (g.state, ev)
            else:
                ev = st[1]
            gs[i], ready[i] = g, ev
        if not self.eager:
            last = torch.cuda.Event()
            last.record(main)
            ready = [last] * len(linears)
        return self.finish_chains(gs, self.enqueue_chains(gs, use_ssr=use_ssr, aga=aga, max_iter=max_iter, order=order,
                                                          ready=ready))

    def run_chains(self, gs, use_ssr: bool = True, aga: str = "hessian", max_iter: int = 100, order: Optional[str] = None):
        """Prologue + sweep of every GPTQ in ``gs`` (Hessians already accumulated on the current stream), longest
        chain first, spread over the side streams; returns ``gs`` finished."""
        return self.finish_chains(gs, self.enqueue_chains(gs, use_ssr=use_ssr, aga=aga, max_iter=max_iter, order=order))

    def enqueue_chains(self, gs, use_ssr: bool = True, aga: str = "hessian", max_iter: int = 100,
                       order: Optional[str] = None, ready=None):
        """Asynchronous half of run_chains(): returns the order to hand to finish_chains().  ``ready``: per GPTQ, the
        event after which its Hessian is complete (default: everything enqueued on the current stream so far).
        Chains go longest first onto the least-loaded stream (by estimated duration), so the longest chain has the
        first -- high-priority -- stream to itself for as long as the others have work elsewhere."""
        main = torch.cuda.current_stream(self.device)
        if ready is None:
            ev = torch.cuda.Event()
            ev.record(main)
            ready = [ev] * len(gs)
        seq = sorted(range(len(gs)), key=lambda i: (-chain_cost(gs[i].rows, gs[i].columns), i))
        load = [0.0] * len(self.streams)
        for i in seq:
            slot = min(range(len(self.streams)), key=lambda k: (load[k], k))
            load[slot] += chain_cost(gs[i].rows, gs[i].columns)
            s = self.streams[slot]
            s.wait_event(ready[i])
            with torch.cuda.stream(s):
                gs[i].enqueue(use_ssr=use_ssr, aga=aga, max_iter=max_iter, order=order)
        return seq

    def finish_chains(self, gs, order):
        main = torch.cuda.current_stream(self.device)
        for i in order:
            gs[i].finish()
        for s in self.streams:
            main.wait_stream(s)
        return gs


def chain_cost(n: int, m: int, block: int = 128) -> float:
    """Estimated duration (ms, one B200) of one linear's prologue + sweep chain for an n x m weight: the damped inverse
    is latency-bound, ~0.24 us per 1000 entries of H; a sweep block costs ~62 us of dependent small kernels plus the
    feedback read-modify-write of the remaining columns.  Fitted to profiles/r01_layer_sweep.json; used only to order
    chains and to deal linears to ranks."""
    blocks = (m + block - 1) // block
    return 0.24e-6 * m * m + blocks * (0.062 + 4.3e-9 * n * m)


class HostPipeline:
    def __init__(self, device, block_size: int = 128, percdamp: float = 0.01, use_ssr: bool = True,
                 aga: str = "hessian", share_inputs: bool = False, num_streams: int = 3, depth: int = 3,
                 order: Optional[str] = None):
        self.device = torch.device(device)
        self.block_size, self.percdamp, self.use_ssr, self.aga = block_size, percdamp, use_ssr, aga
        self.order = order                            # as in GPTQ.quantize; None follows use_ssr
        self.share_inputs = share_inputs
        self.copy_stream = torch.cuda.Stream(self.device)
        self.out_stream = torch.cuda.Stream(self.device)
        self.chain_streams = [torch.cuda.Stream(self.device) for _ in range(max(1, num_streams))]
        self.depth = max(1, int(depth))               # device-side input slots: depth - 1 groups are copied ahead
        self._free_events = [None] * self.depth
        self.h2d_bytes = 0
        self.d2h_bytes = 0
        self._slots = {}

    def _slot(self, kind: str, idx: int, shape, dtype):
        key = (kind, idx)
        need = 1
        for s in shape:
            need *= s
        buf = self._slots.get(key)
        if buf is None or buf.numel() < need or buf.dtype != dtype:
            buf = torch.empty(need, dtype=dtype, device=self.device)
            self._slots[key] = buf
        return buf[:need].view(shape)

    def _stage(self, gi: int, group):
        """Enqueue the host->device copies of one group on the copy stream (slot gi & 1); returns the device views and
        the event that marks them landed."""
        x_host, lins = group
        slot = gi % self.depth
        with torch.cuda.stream(self.copy_stream):
            if self._free_events[slot] is not None:
                self.copy_stream.wait_event(self._free_events[slot])
            x_dev = self._slot("x", slot, tuple(x_host.shape), x_host.dtype)
            x_dev.copy_(x_host, non_blocking=True)
            self.h2d_bytes += x_host.numel() * x_host.element_size()
            w_devs = []
            for li, (name, w_host) in enumerate(lins):
                w_dev = self._slot(f"w{li}", slot, tuple(w_host.shape), w_host.dtype)
                w_dev.copy_(w_host, non_blocking=True)
                self.h2d_bytes += w_host.numel() * w_host.element_size()
                w_devs.append(w_dev)
            ready = torch.cuda.Event()
            ready.record(self.copy_stream)
        return x_dev, w_devs, ready

    def run_iter(self, groups: Iterable[Tuple[torch.Tensor, List[Tuple[str, torch.Tensor]]]]):
        """Generator form of run(): yields the result dicts of one group at a time (their device->host copies may still
        be in flight on the output stream: call ``synchronize()`` or use run()).  The copies of group g+1 are enqueued
        BEFORE the compute of group g is, so they overlap it even though finishing a linear blocks the host once (the
        Cholesky status read, gptq.py:104-106); a model's layers can be streamed through one call."""
        compute = torch.cuda.current_stream(self.device)
        self._free_events = [None] * self.depth      # slot reusable once the compute stream is done with it
        it = iter(groups)
        staged = collections.deque()                 # (group, device views) of groups gi, gi+1, .. already enqueued
        n_staged = 0

        def top_up():
            # keep depth - 1 groups in flight ahead of the one being computed; the slot a new group lands in was last
            # used by a group that has been finished on the host
            nonlocal n_staged
            while len(staged) < self.depth - 1:
                grp = next(it, None)
                if grp is None:
                    return
                staged.append((grp, self._stage(n_staged, grp)))
                n_staged += 1

        top_up()
        if not staged:
            grp = next(it, None)
            if grp is not None:                      # depth == 1: no prefetch
                staged.append((grp, self._stage(n_staged, grp)))
                n_staged += 1
        gi = 0
        while staged:
            cur, (x_dev, w_devs, ready) = staged.popleft()
            compute.wait_event(ready)
            x_host, lins = cur
            shared = HessianState(x_host.shape[-1], self.device) if self.share_inputs else None
            gs = []
            for (name, _), w_dev in zip(lins, w_devs):
                g = GPTQ(LinearView(w_dev), self.block_size, self.percdamp, hessian=shared)
                if shared is None or shared.nsamples == 0:
                    g.add_batch(x_dev)
                gs.append(g)
            x_done = torch.cuda.Event()
            x_done.record(compute)
            # the chains of the group's linears are independent: spread them over the side streams
            for li, g in enumerate(gs):
                s = self.chain_streams[li % len(self.chain_streams)]
                s.wait_event(x_done)
                with torch.cuda.stream(s):
                    g.enqueue(use_ssr=self.use_ssr, aga=self.aga, order=self.order)
            # prefetch now, while this group's kernels run
            top_up()
            out_group = []
            for (name, _), g in zip(lins, gs):
                alpha, mu, _, perm = g.finish()
                done = torch.cuda.Event()
                done.record(compute)       # finish() synchronised the chain's stream; its dtype conversions ran here
                with torch.cuda.stream(self.out_stream):
                    self.out_stream.wait_event(done)
                    out = {"name": name}
                    for key, t in (("alpha", alpha), ("mu", mu), ("T", g.T_int8), ("perm", perm)):
                        h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
                        h.copy_(t, non_blocking=True)
                        t.record_stream(self.out_stream)
                        self.d2h_bytes += t.numel() * t.element_size()
                        out[key] = h
                out_group.append(out)
            for s in self.chain_streams:
                compute.wait_stream(s)
            ev = torch.cuda.Event()
            ev.record(compute)
            self._free_events[gi % self.depth] = ev
            gi += 1
            if not staged:
                top_up()
                if not staged:
                    grp = next(it, None)
                    if grp is not None:
                        staged.append((grp, self._stage(n_staged, grp)))
                        n_staged += 1
            yield out_group

    def synchronize(self):
        self.out_stream.synchronize()
        torch.cuda.current_stream(self.device).synchronize()

    def run(self, groups: Sequence[Tuple[torch.Tensor, List[Tuple[str, torch.Tensor]]]]) -> List[Dict[str, torch.Tensor]]:
        """groups: [(X_host (pinned, (Nt, m) or (B, L, m)), [(name, W_host (pinned fp32 (n, m))), ...]), ...]
        -- each group is one calibration input and the linears that read it.  Returns, per linear in
        order, {'name', 'alpha', 'mu', 'T' (int8), 'perm'} in pinned host memory."""
        results = []
        for out_group in self.run_iter(groups):
            results.extend(out_group)
        self.synchronize()
        return results
