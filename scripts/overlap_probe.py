#!/usr/bin/env python
"""How well do the latency-bound chains (damped inverse, column sweep) of independent linears overlap when they
are enqueued on different CUDA streams?  Prints, for k = 1, 2, 4, 7 concurrent chains, the wall time on the
device and the time per chain.  Development probe (one B200)."""

import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tq100  # noqa: E402
from tq100.pipeline import LinearView  # noqa: E402

DEV = torch.device("cuda:0")


def run_concurrent(fns, streams):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    main = torch.cuda.current_stream()
    e0.record(main)
    t0 = time.perf_counter()
    for fn, s in zip(fns, streams):
        s.wait_stream(main)
        with torch.cuda.stream(s):
            fn()
    host_ms = 1e3 * (time.perf_counter() - t0)
    for s in streams:
        main.wait_stream(s)
    e1.record(main)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1), host_ms


def main():
    n, m = int(os.environ.get("N", 4096)), int(os.environ.get("M", 4096))
    kmax = 7
    g = torch.Generator(device=DEV).manual_seed(1)
    X = torch.randn((32768, m), device=DEV, dtype=torch.float16, generator=g)
    states, gqs = [], []
    for i in range(kmax):
        st = tq100.HessianState(m, DEV)
        st.add_batch(X)
        states.append(st)
        W = torch.randn((n, m), device=DEV, generator=g) * 0.02
        gqs.append(tq100.GPTQ(LinearView(W), hessian=st))
    streams = [torch.cuda.Stream(DEV) for _ in range(kmax)]
    out = {"n": n, "m": m}

    def inv(i):
        def f():
            states[i]._cache.clear()
            states[i].damped_inverse(0.01)
        return f

    def sweep(i, ssr):
        def f():
            gqs[i].enqueue(use_ssr=ssr)
        return f

    for name, mk in (("inverse", inv), ("sweep_ssr", lambda i: sweep(i, True)), ("sweep_seq", lambda i: sweep(i, False))):
        for k in (1, 2, 4, 7):
            fns = [mk(i) for i in range(k)]
            best, host = 1e9, 0.0
            for rep in range(3):
                if name != "inverse":
                    for i in range(kmax):
                        states[i].damped_inverse(0.01)
                t, h = run_concurrent(fns, streams[:k])
                if name != "inverse":
                    for i in range(k):
                        gqs[i].finish()
                if t < best:
                    best, host = t, h
            out[f"{name}_k{k}_ms"] = best
            print(f"{name:10s} k={k}: {best:8.2f} ms total, {best / k:7.2f} ms per chain, host enqueue {host:6.2f} ms", flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", f"overlap_probe_{n}x{m}.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
