#!/usr/bin/env python
"""Kernel timeline of ONE row-sharded SSR sweep on a single GPU: the peer mailboxes are opened with a world of one rank
(the exchange then goes through this GPU's own mailbox), so the per-block cost of the sharded step -- fold + send, gather,
select, coefficient operand, AGA vector, fit, feedback GEMM on a row slab -- can be read without cross-GPU skew.
    python scripts/sharded_sweep_probe.py [rows] [m]"""
import ctypes
import json
import os
import sys
from collections import defaultdict

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tq100  # noqa: E402
from tq100 import _lib  # noqa: E402
from tq100.pipeline import LinearView  # noqa: E402

DEV = torch.device("cuda:0")
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 512
m = int(sys.argv[2]) if len(sys.argv) > 2 else 11008
lib = _lib.load()
h = (ctypes.c_ubyte * 64)()
_lib.check(lib.tq_comm_p2p_alloc(2 * 32768 + 16, 0, 1, h), "alloc")
_lib.check(lib.tq_comm_p2p_open(h), "open")
g = torch.Generator(device=DEV).manual_seed(0)
X = torch.randn((16384, m), device=DEV, dtype=torch.float16, generator=g)
W = torch.randn((rows, m), device=DEV, generator=g) * 0.02
q0 = tq100.GPTQ(LinearView(W))
q0.add_batch(X)
q0.state.damped_inverse(0.01)


def sweep(flags):
    q = tq100.GPTQ(LinearView(W), hessian=q0.state)
    q.sweep_flags = flags
    q.quantize(use_ssr=True)


for flags, tag in ((0, "plain"), (_lib.SWEEP_ROW_SHARD, "row_shard_p2p")):
    sweep(flags)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    sweep(flags)
    e1.record()
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        sweep(flags)
        torch.cuda.synchronize()
    by = defaultdict(lambda: [0, 0.0])
    for e in prof.events():
        if e.device_type == torch.autograd.DeviceType.CUDA:
            k = e.name.split("(")[0][:48]
            by[k][0] += 1
            by[k][1] += e.time_range.end - e.time_range.start
    print(f"== {tag}: rows {rows} m {m}: sweep {e0.elapsed_time(e1):.2f} ms")
    for k, v in sorted(by.items(), key=lambda kv: -kv[1][1])[:12]:
        print(f"   {k:50s} {v[0]:5d} {v[1] / 1e3:8.2f} ms {v[1] / v[0]:8.1f} us")
