#!/usr/bin/env python
"""Bandwidth of the packed ternary layer kernels on 7B shapes: tq_tl_gemv at decode sizes, tq_tl_dequant,
tq_tl_pack, and the many-token paths (dense weight + library GEMM vs tq_tl_gemm_tc).  CUDA events around a CUDA-graph replay of the calls; every timed call reads a DIFFERENT layer copy and the copies total
more than twice the 126 MB L2, so the codes come from HBM.  Algorithmic bytes per gemv call: n*ceil(m/16)*4 (codes)
+ 16*n*nb (weight table) + M*m*x_bytes + 4*M*n (y).  Writes gpurun_out/tl_bench.json."""

import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tq100  # noqa: E402


# TL_BENCH_PREFILL: "1" everything, "0" skip the many-token table, "only" just that table (-> tl_bench_prefill.json)
PREFILL = os.environ.get("TL_BENCH_PREFILL", "1")


def peak_gbs():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6554.6


def timed(fn, copies, iters):
    """Mean device time of fn(copy) over `iters` calls cycling through the copies.  The calls are captured into one CUDA
    graph (the library launches on the current = capturing stream) so host launch overhead is not what is measured."""
    for i in range(min(len(copies), 3)):
        fn(copies[i])
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        for i in range(iters):
            fn(copies[i % len(copies)])
    graph.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    graph.replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e-3


def main():
    dev = torch.device("cuda:0")
    peak = peak_gbs()
    out = {"peak_gbs": peak, "gemv": [], "dequant": [], "pack": [], "prefill": []}
    gen = torch.Generator(device=dev).manual_seed(1)
    for n, m in ((4096, 4096), (11008, 4096), (4096, 11008)):
        nb = (m + 127) // 128
        code_bytes = n * ((m + 15) // 16) * 4
        ncopies = 3 if PREFILL == "only" else max(3, int(2.2 * 126e6 / code_bytes) + 1)
        T = torch.randint(-1, 2, (n, m), generator=gen, device=dev, dtype=torch.int8)
        alpha = 0.01 + 0.02 * torch.rand((n, nb), generator=gen, device=dev)
        mu = 0.004 * torch.randn((n, nb), generator=gen, device=dev)
        for order in (("permuted",) if PREFILL == "only" else ("identity", "permuted")):
            perm = torch.arange(m, device=dev) if order == "identity" else torch.randperm(m, generator=gen, device=dev)
            layers = []
            for _ in range(ncopies):
                layer = tq100.TernaryLinear(m, n, bias=False, dtype=torch.float16, device=dev)
                layer.set_quantized_params(alpha, mu, T, perm)
                layer._prepared()
                layers.append(layer)
            for M in (() if PREFILL == "only" else (1, 4, 16)):
                x = torch.randn((M, m), generator=gen, device=dev).half()
                sec = timed(lambda L: L(x), layers, 4 * ncopies)
                nbytes = code_bytes + 16 * n * nb + M * m * 2 + 4 * M * n
                out["gemv"].append({"n": n, "m": m, "order": order, "tokens": M, "us": sec * 1e6,
                                    "algorithmic_bytes": nbytes, "gbs": nbytes / sec / 1e9,
                                    "frac_of_hbm_peak": nbytes / sec / 1e9 / peak,
                                    "launches_per_call": (M + 3) // 4})
            if PREFILL != "only":
                sec = timed(lambda L: L._dequantize(), layers, 2 * ncopies)
                nbytes = code_bytes + 16 * n * nb + 2 * n * m
                out["dequant"].append({"n": n, "m": m, "order": order, "us": sec * 1e6, "algorithmic_bytes": nbytes,
                                       "gbs": nbytes / sec / 1e9, "frac_of_hbm_peak": nbytes / sec / 1e9 / peak})
            if order == "permuted" and PREFILL != "0":
                for M in (64, 512, 2048):
                    x = torch.randn((M, m), generator=gen, device=dev).half()
                    row = {"n": n, "m": m, "tokens": M, "flops": 2.0 * M * n * m}
                    for mode in ("dense", "fused"):
                        for L in layers:
                            L.fused_gemm = (mode == "fused")
                            L.fused_max_tokens = 1 << 30
                        try:
                            sec = timed(lambda L: L(x), layers[:3], 6)
                            row[mode + "_us"] = sec * 1e6
                            row[mode + "_tflops"] = row["flops"] / sec / 1e12
                        except Exception as e:      # the fused kernel is new: keep the rest of the table
                            row[mode + "_error"] = str(e)[:200]
                    for L in layers:
                        L.fused_gemm = False
                    out["prefill"].append(row)
            del layers
        if PREFILL == "only":
            continue
        lib = tq100._lib.load()
        p32 = torch.randperm(m, generator=gen, device=dev).to(torch.int32)
        wpr = (m + 15) // 16
        codes = torch.empty((n, wpr), dtype=torch.int32, device=dev)

        def pack(_):
            tq100._lib.check(lib.tq_tl_pack(tq100._lib.ptr(T), n, m, tq100._lib.ptr(p32), tq100._lib.ptr(codes), wpr,
                                            tq100._lib.stream()), "tq_tl_pack")
        sec = timed(pack, [None], 10)
        nbytes = n * m + code_bytes
        out["pack"].append({"n": n, "m": m, "us": sec * 1e6, "algorithmic_bytes": nbytes, "gbs": nbytes / sec / 1e9,
                            "frac_of_hbm_peak": nbytes / sec / 1e9 / peak, "note": "T (int8, 17-45 MB) is L2-resident"})
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    name = "tl_bench_prefill.json" if PREFILL == "only" else "tl_bench.json"
    with open(os.path.join(ROOT, "gpurun_out", name), "w") as f:
        json.dump(out, f, indent=1)
    for row in out["gemv"]:
        print("gemv  {n:6d}x{m:<6d} {order:9s} M={tokens:<2d} {us:8.1f} us {gbs:8.1f} GB/s ({frac_of_hbm_peak:.2f} of peak)".format(**row))
    for row in out["prefill"]:
        print("prefill", json.dumps(row))
    for row in out["pack"]:
        print("pack  {n:6d}x{m:<6d} {us:8.1f} us {gbs:8.1f} GB/s".format(**row))
    for row in out["dequant"]:
        print("dequant {n:6d}x{m:<6d} {order:9s} {us:8.1f} us {gbs:8.1f} GB/s ({frac_of_hbm_peak:.2f} of peak)".format(**row))


if __name__ == "__main__":
    main()
