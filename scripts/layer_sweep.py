#!/usr/bin/env python
"""BASELINE.json configs[4]: single-layer sweep, 4096x4096 up to 8192x28672 -- Hessian and column-sweep kernels
against the roofline, one B200.  Writes profiles/<tag>_layer_sweep.json and prints a table."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tq100  # noqa: E402
from tq100.pipeline import LinearView  # noqa: E402

DEV = torch.device("cuda:0")
PEAKS = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
TF_PEAK = PEAKS.get("bf16_tflops", 1590.0)
HBM_PEAK = PEAKS.get("hbm_gbs", 6650.0)


def ev():
    return torch.cuda.Event(enable_timing=True)


def best_of(fn, reps=3, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = ev(), ev()
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts)


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
    nt = int(os.environ.get("NT", 128 * 2048))
    shapes = [(4096, 4096), (11008, 4096), (4096, 11008), (5120, 5120), (13824, 5120), (5120, 13824),
              (8192, 8192), (28672, 8192), (8192, 28672)]
    rows = []
    for n, m in shapes:
        g = torch.Generator(device=DEV).manual_seed(n + m)
        X = torch.empty((nt, m), device=DEV, dtype=torch.float16)
        for lo in range(0, nt, 32768):
            X[lo:lo + 32768] = torch.randn((min(32768, nt - lo), m), device=DEV, generator=g).to(torch.float16)
        W = torch.randn((n, m), device=DEV, generator=g) * 0.02
        q = tq100.GPTQ(LinearView(W))
        st = q.state

        def hess():
            st.H.zero_()
            st.nsamples = 0
            st.add_batch(X)
        t_h = best_of(hess, reps=3)
        useful = nt * m * (m + 1)

        def inv():
            st._cache.clear()
            st.damped_inverse(0.01)
        t_inv = best_of(inv, reps=2)
        rec = {"n": n, "m": m, "tokens": nt,
               "hessian_ms": t_h, "hessian_tflops_useful": useful / t_h / 1e9, "hessian_frac_of_bf16_burst_peak": useful / t_h / 1e9 / TF_PEAK,
               "inverse_ms": t_inv, "inverse_tflops_fp32_equiv": m ** 3 / t_inv / 1e9}
        for name, kw in (("sweep_seq_ms", dict(use_ssr=False)), ("sweep_ssr_ms", dict(use_ssr=True))):
            t = best_of(lambda: q.quantize(**kw), reps=2)
            rec[name] = t
        # algorithmic HBM bytes of the sequential sweep: RMW of W[:, rem] per block + block reads/writes
        nb = (m + 127) // 128
        rmw = sum(8 * n * (m - 128 * (k + 1)) for k in range(nb - 1))
        rec["sweep_seq_algorithmic_gb"] = (rmw + n * m * (4 + 1 + 8)) / 1e9
        rec["sweep_seq_gbs"] = rec["sweep_seq_algorithmic_gb"] / (rec["sweep_seq_ms"] - 0.0) * 1e3
        rec["sweep_seq_frac_of_hbm_peak"] = rec["sweep_seq_gbs"] / HBM_PEAK
        rows.append(rec)
        print(json.dumps(rec), flush=True)
        del X, W, q, st
        torch.cuda.empty_cache()
    out = os.path.join(ROOT, "gpurun_out", f"{tag}_layer_sweep.json")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    json.dump({"peaks": {"bf16_tflops_burst": TF_PEAK, "hbm_gbs": HBM_PEAK}, "rows": rows}, open(out, "w"), indent=1)
    print(f"{'n x m':>14s} {'hess ms':>8s} {'TF/s':>7s} {'frac':>5s} {'inv ms':>7s} {'seq ms':>7s} {'GB/s':>6s} {'ssr ms':>7s}")
    for r in rows:
        print(f"{r['n']:>6d}x{r['m']:<7d} {r['hessian_ms']:8.2f} {r['hessian_tflops_useful']:7.0f} {r['hessian_frac_of_bf16_burst_peak']:5.2f} "
              f"{r['inverse_ms']:7.1f} {r['sweep_seq_ms']:7.1f} {r['sweep_seq_gbs']:6.0f} {r['sweep_ssr_ms']:7.1f}")


if __name__ == "__main__":
    main()
