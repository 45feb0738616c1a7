#!/usr/bin/env python
"""HBM-bound kernels against the measured copy bandwidth (MEASURED_PEAKS.json hbm_gbs): algorithmic bytes / CUDA-event time.
Inputs are larger than L2 (or L2 is flushed between repetitions by rotating over several buffers)."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tq100
from tq100 import _lib
lib = _lib.load()
DEV = torch.device("cuda:0")
PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0


def timed(fns, reps=3):
    """fns: list of closures over DIFFERENT buffers (rotation defeats L2 reuse); returns best ms per call."""
    for f in fns:
        f()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for f in fns:
            f()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) / len(fns))
    return best


def main():
    out = []
    g = torch.Generator(device=DEV).manual_seed(0)
    # ---- pack / unpack: 11008 x 4096 layer codes (45 MB in, 11 MB out), 8 rotating buffers (> L2 in total)
    n, m = 11008, 4096
    Ts = [torch.randint(-1, 2, (n, m), device=DEV, dtype=torch.int8, generator=g) for _ in range(8)]
    packed = [torch.empty((n * m + 3) // 4, dtype=torch.uint8, device=DEV) for _ in range(8)]
    ms = timed([lambda i=i: _lib.check(lib.tq_pack2b(_lib.ptr(Ts[i]), n * m, _lib.ptr(packed[i]), _lib.stream()), "pack") for i in range(8)])
    out.append(("pack2b 11008x4096", 1.25 * n * m, ms))
    outs = [torch.empty((n, m), dtype=torch.int8, device=DEV) for _ in range(8)]
    ms = timed([lambda i=i: _lib.check(lib.tq_unpack2b(_lib.ptr(packed[i]), n * m, _lib.ptr(outs[i]), _lib.stream()), "unpack") for i in range(8)])
    out.append(("unpack2b 11008x4096", 1.25 * n * m, ms))
    del Ts, packed, outs
    # ---- fused ATQ block fit (contiguous 128-column block of a 4096-wide row-major W): reads 4*n*128 (strided rows), writes
    #      n*128 (T) + 8n (alpha, mu) + 8*n*128 (E hi, lo)
    for n in (4096, 11008, 28672):
        m = 4096
        Ws = [torch.randn((n, m), device=DEV, generator=g) * 0.02 for _ in range(3)]
        T = torch.empty((n, 128), dtype=torch.int8, device=DEV)
        a = torch.empty(n, device=DEV); u = torch.empty(n, device=DEV)
        E = torch.empty((n, 128), device=DEV)
        fns = []
        for i in range(3):
            for c0 in (0, 1024, 2048, 3072):
                fns.append(lambda i=i, c0=c0: _lib.check(lib.tq_atq_block(_lib.ptr(Ws[i]), m, n, None, c0, 128, None, 100, _lib.ptr(T), 128,
                                                                           _lib.ptr(a), _lib.ptr(u), 1, _lib.ptr(E), 128, None, _lib.stream()), "atq"))
        ms = timed(fns)
        out.append((f"atq_block n={n} (init+ITF+error)", n * 128 * (4 + 1 + 4) + 8 * n, ms))
        del Ws
    # ---- SSR statistics (two passes over W[:, rem]) at full width
    for n, m in ((4096, 4096), (4096, 11008)):
        Ws = [torch.randn((n, m), device=DEV, generator=g) * 0.02 for _ in range(3)]
        rem = torch.arange(m, device=DEV, dtype=torch.int32)
        chunks = lib.tq_ssr_num_chunks(n)
        rowmean = torch.empty(n, device=DEV); partials = torch.empty((chunks, 2, m), device=DEV)
        ms = timed([lambda i=i: _lib.check(lib.tq_ssr_stats(_lib.ptr(Ws[i]), m, n, _lib.ptr(rem), m, _lib.ptr(rowmean), _lib.ptr(partials), _lib.stream()), "stats") for i in range(3)])
        out.append((f"ssr_stats {n}x{m} (2 reads of W)", 2 * 4 * n * m, ms))
        del Ws
    # ---- error feedback GEMM: compulsory bytes = RMW of W[:, rem]
    for n, m in ((4096, 4096), (11008, 4096), (4096, 11008)):
        Ws = [torch.randn((n, m), device=DEV, generator=g) * 0.02 for _ in range(3)]
        E = torch.randn((n, 128), device=DEV, generator=g) * 0.01
        A = torch.randn((m, 256), device=DEV, generator=g)
        Hinv = (A @ A.T / 256 + torch.eye(m, device=DEV)).contiguous()
        ws = torch.empty(lib.tq_err_feedback_tc_workspace_floats(n, 128, m - 128), device=DEV)
        ms = timed([lambda i=i: _lib.check(lib.tq_err_feedback_tc(_lib.ptr(Ws[i]), m, n, _lib.ptr(E), 128, _lib.ptr(Hinv), m, None, 0, 128, None, 128,
                                                                    m - 128, _lib.ptr(ws), _lib.stream()), "fb") for i in range(3)])
        out.append((f"err_feedback_tc {n}x{m} first block (split + coef + GEMM)", 8 * n * (m - 128), ms))
        del Ws, Hinv, ws
    rows = []
    print(f"{'kernel':58s} {'MB':>9s} {'us':>9s} {'GB/s':>8s} {'of HBM peak':>11s}")
    for name, byts, ms in out:
        gbs = byts / ms / 1e6
        rows.append({"kernel": name, "algorithmic_bytes": byts, "us": ms * 1e3, "gbs": gbs, "frac_of_hbm_peak": gbs / PEAK})
        print(f"{name:58s} {byts / 1e6:9.1f} {ms * 1e3:9.1f} {gbs:8.0f} {gbs / PEAK:11.3f}")
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump({"hbm_peak_gbs": PEAK, "rows": rows}, open(os.path.join(ROOT, "gpurun_out", "kernel_bw.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
