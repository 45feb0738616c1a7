#!/usr/bin/env python
"""Kernel timeline of ONE 7B-shaped layer through LayerDriver (the bench's inner loop), from torch.profiler's CUDA
activity records (CUPTI): per kernel name the summed duration, and for the whole layer the wall time, the time at least
one kernel was running, and the time two or more ran together -- what the multi-stream chains actually overlap.
A diagnostic, not a bench value (CUPTI adds a few microseconds per launch).

    python scripts/timeline.py [--shared] [--streams 4] [--order ssr] -> gpurun_out/timeline_<tag>.json
"""

import argparse
import json
import os
import sys
from collections import defaultdict

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tq100  # noqa: E402,F401
from tq100.pipeline import LayerDriver  # noqa: E402

DEV = torch.device("cuda:0")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shared", action="store_true")
    ap.add_argument("--streams", type=int, default=4)
    ap.add_argument("--order", default="ssr")
    ap.add_argument("--d", type=int, default=4096)
    ap.add_argument("--ffn", type=int, default=11008)
    ap.add_argument("--samples", type=int, default=128)
    ap.add_argument("--tag", default=None)
    ap.add_argument("--only", default=None, help="comma-separated linear names (default: all seven)")
    args = ap.parse_args()
    d, f = args.d, args.ffn
    lins = [("q_proj", d, d, "attn_in"), ("k_proj", d, d, "attn_in"), ("v_proj", d, d, "attn_in"),
            ("o_proj", d, d, "o_in"), ("gate_proj", f, d, "mlp_in"), ("up_proj", f, d, "mlp_in"),
            ("down_proj", d, f, "down_in")]
    gen = torch.Generator(device=DEV).manual_seed(1)
    acts = {}
    for key, width in (("attn_in", d), ("o_in", d), ("mlp_in", d), ("down_in", f)):
        B = torch.randn((64, width), device=DEV, generator=gen)
        x = torch.empty((args.samples, 2048, width), device=DEV, dtype=torch.float16)
        for j in range(args.samples):
            x[j] = (torch.randn((2048, width), device=DEV, generator=gen)
                    + 0.0625 * (torch.randn((2048, 64), device=DEV, generator=gen) @ B)).half()
        acts[key] = x
    ws = {name: torch.randn((n, m), device=DEV, generator=gen) * 0.02 for name, n, m, _ in lins}
    drv = LayerDriver(DEV, num_streams=args.streams, share_inputs=args.shared)
    if args.only:
        lins = [l for l in lins if l[0] in args.only.split(",")]

    def layer():
        return drv.quantize([(name, ws[name], acts[src]) for name, n, m, src in lins], use_ssr=args.order == "ssr",
                            order=args.order)

    for _ in range(2):
        layer()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    layer()
    e1.record()
    torch.cuda.synchronize()
    plain_ms = e0.elapsed_time(e1)

    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        layer()
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range.end > e.time_range.start]
    rows = []
    for e in evs:
        rows.append((e.time_range.start, e.time_range.end, e.name))
    rows.sort()
    t0, t1 = rows[0][0], max(r[1] for r in rows)
    by = defaultdict(lambda: [0, 0.0])
    for s, e, nme in rows:
        k = nme.split("(")[0][:60]
        by[k][0] += 1
        by[k][1] += (e - s)
    # sweep line: time with >= 1 and >= 2 kernels in flight
    pts = []
    for s, e, _ in rows:
        pts.append((s, 1))
        pts.append((e, -1))
    pts.sort()
    busy1 = busy2 = 0.0
    depth, last = 0, pts[0][0]
    for t, dlt in pts:
        if depth >= 1:
            busy1 += t - last
        if depth >= 2:
            busy2 += t - last
        depth += dlt
        last = t
    # the Hessian phase ends when the last hessian kernel ends
    hess_end = max((e for s, e, nme in rows if "hessian" in nme), default=t0)
    chain_rows = [(s, e, nme) for s, e, nme in rows if s >= hess_end]
    out = {"layer_ms_cuda_events_unprofiled": plain_ms, "wall_ms_profiled": (t1 - t0) / 1e3, "busy_ge1_ms": busy1 / 1e3,
           "busy_ge2_ms": busy2 / 1e3, "idle_ms": ((t1 - t0) - busy1) / 1e3,
           "hessian_phase_ms": (hess_end - t0) / 1e3, "chain_phase_ms": (t1 - hess_end) / 1e3,
           "chain_phase_kernel_sum_ms": sum(e - s for s, e, _ in chain_rows) / 1e3, "launches": len(rows),
           "kernels": {k: {"launches": v[0], "total_ms": v[1] / 1e3, "avg_us": v[1] / v[0]}
                       for k, v in sorted(by.items(), key=lambda kv: -kv[1][1])[:30]},
           "args": vars(args)}
    tag = args.tag or (("shared" if args.shared else "7h") + f"_{args.order}_s{args.streams}")
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    # per-stream view from the chrome trace (kernel records carry their stream id)
    trace = os.path.join(ROOT, "gpurun_out", f"trace_{tag}.json")
    prof.export_chrome_trace(trace)
    tr = json.load(open(trace))
    per = defaultdict(lambda: {"first_ms": None, "last_ms": 0.0, "busy_ms": 0.0, "kernels": 0, "top": defaultdict(float)})
    ks = [e for e in tr["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "dur" in e]
    base = min(e["ts"] for e in ks)
    for e in ks:
        st = per[str(e.get("args", {}).get("stream", e.get("tid")))]
        b = (e["ts"] - base) / 1e3
        st["first_ms"] = b if st["first_ms"] is None else min(st["first_ms"], b)
        st["last_ms"] = max(st["last_ms"], b + e["dur"] / 1e3)
        st["busy_ms"] += e["dur"] / 1e3
        st["kernels"] += 1
        st["top"][e["name"].split("(")[0][:40]] += e["dur"] / 1e3
    out["streams"] = {k: {"first_ms": v["first_ms"], "last_ms": v["last_ms"], "busy_ms": v["busy_ms"], "kernels": v["kernels"],
                          "top": dict(sorted(v["top"].items(), key=lambda kv: -kv[1])[:4])}
                      for k, v in sorted(per.items(), key=lambda kv: kv[1]["first_ms"])}
    os.remove(trace)
    for k, v in out["streams"].items():
        print(f"stream {k:>4s}: {v['first_ms']:8.2f} -> {v['last_ms']:8.2f} ms  busy {v['busy_ms']:7.2f}  n={v['kernels']:5d}  "
              + ", ".join(f"{n} {t:.1f}" for n, t in v["top"].items()))
    with open(os.path.join(ROOT, "gpurun_out", f"timeline_{tag}.json"), "w") as fh:
        json.dump(out, fh, indent=1)
    print(json.dumps({k: v for k, v in out.items() if k != "kernels"}))
    for k, v in list(out["kernels"].items())[:22]:
        print(f"{k:62s} {v['launches']:6d} {v['total_ms']:9.2f} ms {v['avg_us']:9.1f} us")


if __name__ == "__main__":
    main()
