#!/usr/bin/env python
"""Bandwidth of the packed ternary layer kernels on 7B shapes: tq_tl_gemv at decode sizes, tq_tl_dequant,
tq_tl_pack.  CUDA events on the launching stream; every timed call reads a DIFFERENT layer copy and the copies total
more than twice the 126 MB L2, so the codes come from HBM.  Algorithmic bytes per gemv call: n*ceil(m/16)*4 (codes)
+ 16*n*nb (weight table) + M*m*x_bytes + 4*M*n (y).  Writes gpurun_out/tl_bench.json."""

import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tq100  # noqa: E402


def peak_gbs():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6554.6


def timed(fn, copies, iters):
    for i in range(min(len(copies), 3)):
        fn(copies[i])
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(iters):
        fn(copies[i % len(copies)])
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e-3


def main():
    dev = torch.device("cuda:0")
    peak = peak_gbs()
    out = {"peak_gbs": peak, "gemv": [], "dequant": [], "pack": []}
    gen = torch.Generator(device=dev).manual_seed(1)
    for n, m in ((4096, 4096), (11008, 4096), (4096, 11008)):
        nb = (m + 127) // 128
        code_bytes = n * ((m + 15) // 16) * 4
        ncopies = max(3, int(2.2 * 126e6 / code_bytes) + 1)
        T = torch.randint(-1, 2, (n, m), generator=gen, device=dev, dtype=torch.int8)
        alpha = 0.01 + 0.02 * torch.rand((n, nb), generator=gen, device=dev)
        mu = 0.004 * torch.randn((n, nb), generator=gen, device=dev)
        for order in ("identity", "permuted"):
            perm = torch.arange(m, device=dev) if order == "identity" else torch.randperm(m, generator=gen, device=dev)
            layers = []
            for _ in range(ncopies):
                layer = tq100.TernaryLinear(m, n, bias=False, dtype=torch.float16, device=dev)
                layer.set_quantized_params(alpha, mu, T, perm)
                layer._prepared()
                layers.append(layer)
            for M in (1, 4, 16):
                x = torch.randn((M, m), generator=gen, device=dev).half()
                sec = timed(lambda L: L(x), layers, 4 * ncopies)
                nbytes = code_bytes + 16 * n * nb + M * m * 2 + 4 * M * n
                out["gemv"].append({"n": n, "m": m, "order": order, "tokens": M, "us": sec * 1e6,
                                    "algorithmic_bytes": nbytes, "gbs": nbytes / sec / 1e9,
                                    "frac_of_hbm_peak": nbytes / sec / 1e9 / peak,
                                    "launches_per_call": (M + 3) // 4})
            sec = timed(lambda L: L._dequantize(), layers, 2 * ncopies)
            nbytes = code_bytes + 16 * n * nb + 2 * n * m
            out["dequant"].append({"n": n, "m": m, "order": order, "us": sec * 1e6, "algorithmic_bytes": nbytes,
                                   "gbs": nbytes / sec / 1e9, "frac_of_hbm_peak": nbytes / sec / 1e9 / peak})
            del layers
        layer = tq100.TernaryLinear(m, n, bias=False, dtype=torch.float16, device=dev)
        perm = torch.randperm(m, generator=gen, device=dev)
        sec = timed(lambda _: layer.set_quantized_params(alpha, mu, T, perm), [None], 10)
        out["pack"].append({"n": n, "m": m, "us_set_quantized_params": sec * 1e6,
                            "algorithmic_bytes": n * m + code_bytes})
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "tl_bench.json"), "w") as f:
        json.dump(out, f, indent=1)
    for row in out["gemv"]:
        print("gemv  {n:6d}x{m:<6d} {order:9s} M={tokens:<2d} {us:8.1f} us {gbs:8.1f} GB/s ({frac_of_hbm_peak:.2f} of peak)".format(**row))
    for row in out["dequant"]:
        print("dequant {n:6d}x{m:<6d} {order:9s} {us:8.1f} us {gbs:8.1f} GB/s ({frac_of_hbm_peak:.2f} of peak)".format(**row))


if __name__ == "__main__":
    main()
