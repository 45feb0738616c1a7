#!/usr/bin/env python
"""Per-stage CUDA-event timings of the hot path on LLaMA-2-7B linear shapes (one B200).
Development probe: prints a table and writes gpurun_out/stage_probe.json."""

import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tq100  # noqa: E402
from tq100 import _lib  # noqa: E402

DEV = torch.device("cuda:0")


def timed(fn, reps=3, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts), sum(ts) / len(ts)


def main():
    lib = _lib.load()
    shapes = [(4096, 4096), (11008, 4096), (4096, 11008)]
    nt_full = int(os.environ.get("NT", 262144))
    out = []
    for n, m in shapes:
        g = torch.Generator(device=DEV).manual_seed(n * 7 + m)
        X = torch.randn((nt_full, m), device=DEV, dtype=torch.float16, generator=g)
        W = torch.randn((n, m), device=DEV, generator=g) * 0.02
        layer = torch.nn.Linear(m, n, bias=False).to(DEV)
        layer.weight.data = W
        rec = {"n": n, "m": m, "nt": nt_full}

        # Hessian (tcgen05)
        H = torch.zeros((m, m), device=DEV)
        def hess():
            _lib.check(lib.tq_hessian_accum(_lib.ptr(H), m, _lib.ptr(X), nt_full, m, m, _lib.F16, _lib.HESS_TCGEN05,
                                            _lib.stream()), "hess")
        best, avg = timed(hess, reps=3)
        useful = nt_full * m * (m + 1)
        rec["hessian_ms"] = best
        rec["hessian_tflops_useful"] = useful / best / 1e9
        rec["hessian_tflops_dense_equiv"] = 2 * nt_full * m * m / best / 1e9

        # small-call Hessian (one 2048-token sample per call)
        Xs = X[:2048]
        def hess_small():
            _lib.check(lib.tq_hessian_accum(_lib.ptr(H), m, _lib.ptr(Xs), 2048, m, m, _lib.F16, _lib.HESS_TCGEN05,
                                            _lib.stream()), "hess")
        best_s, _ = timed(hess_small, reps=5)
        rec["hessian_2048tok_ms"] = best_s
        rec["hessian_2048tok_tflops_useful"] = 2048 * m * (m + 1) / best_s / 1e9

        # damped inverse
        st = tq100.HessianState(m, DEV)
        st.add_batch(X[:16384])
        def inv():
            st._cache.clear()
            st.damped_inverse(0.01)
        best, _ = timed(inv, reps=2)
        rec["chol_inverse_ms"] = best
        rec["chol_inverse_tflops"] = (m ** 3) / best / 1e9

        # sweep variants
        gq = tq100.GPTQ(layer, hessian=st)
        for name, kw in (("sweep_seq_ms", dict(use_ssr=False)), ("sweep_ssr_ms", dict(use_ssr=True)),
                         ("sweep_seq_noaga_ms", dict(use_ssr=False, aga="none"))):
            def sw():
                gq.quantize(**kw)
            best, _ = timed(sw, reps=2)
            rec[name] = best
        # single-block stage timings at the first block (rem = m - 128)
        import numpy as np
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import gpu_util as G
        Wd = W.clone()
        remv = torch.arange(m, device=DEV, dtype=torch.int32)
        chunks = lib.tq_ssr_num_chunks(n)
        rowmean = torch.empty(n, device=DEV)
        partials = torch.empty((chunks, 2, m), device=DEV)
        blk = torch.empty(128, device=DEV, dtype=torch.int32)
        newrem = torch.empty(m, device=DEV, dtype=torch.int32)
        sims = torch.empty(2 * m + 2, device=DEV)
        def stats():
            _lib.check(lib.tq_ssr_stats(_lib.ptr(Wd), m, n, _lib.ptr(remv), m, _lib.ptr(rowmean), _lib.ptr(partials), _lib.stream()), "stats")
        def select():
            _lib.check(lib.tq_ssr_select(_lib.ptr(partials), chunks, _lib.ptr(rowmean), n, None, _lib.ptr(remv), m, 128,
                                         _lib.ptr(blk), _lib.ptr(newrem), _lib.ptr(sims), _lib.stream()), "select")
        rec["ssr_stats_us"] = timed(stats, reps=5)[0] * 1e3
        rec["ssr_select_us"] = timed(select, reps=5)[0] * 1e3
        Hinv = st.damped_inverse(0.01)[1]
        E = torch.randn((n, 128), device=DEV) * 0.01
        ws = torch.empty(lib.tq_err_feedback_tc_workspace_floats(n, 128, m - 128), device=DEV)
        rem_g = newrem[: m - 128]
        def fb_seq():
            _lib.check(lib.tq_err_feedback_tc(_lib.ptr(Wd), m, n, _lib.ptr(E), 128, _lib.ptr(Hinv), m, None, 0, 128, None, 128,
                                              m - 128, _lib.ptr(ws), _lib.stream()), "fb")
        def fb_gather():
            _lib.check(lib.tq_err_feedback_tc(_lib.ptr(Wd), m, n, _lib.ptr(E), 128, _lib.ptr(Hinv), m, _lib.ptr(blk), 0, 128,
                                              _lib.ptr(rem_g), 0, m - 128, _lib.ptr(ws), _lib.stream()), "fb")
        def fb_ffma():
            _lib.check(lib.tq_err_feedback(_lib.ptr(Wd), m, n, _lib.ptr(E), 128, _lib.ptr(Hinv), m, None, 0, 128, None, 128,
                                           m - 128, _lib.stream()), "fb")
        rec["feedback_tc_seq_us"] = timed(fb_seq, reps=5)[0] * 1e3
        rec["feedback_tc_gather_us"] = timed(fb_gather, reps=5)[0] * 1e3
        rec["feedback_ffma_seq_us"] = timed(fb_ffma, reps=3)[0] * 1e3
        rec["feedback_rmw_bytes"] = 8 * n * (m - 128)
        rec["feedback_tc_seq_gbs"] = rec["feedback_rmw_bytes"] / rec["feedback_tc_seq_us"] / 1e3
        T8 = torch.empty((n, 128), device=DEV, dtype=torch.int8); al = torch.empty(n, device=DEV); mu_ = torch.empty(n, device=DEV)
        Eo = torch.empty((n, 128), device=DEV)
        def atq():
            _lib.check(lib.tq_atq_block(_lib.ptr(Wd), m, n, _lib.ptr(blk), 0, 128, None, 100, _lib.ptr(T8), 128, _lib.ptr(al),
                                        _lib.ptr(mu_), 1, _lib.ptr(Eo), 128, None, _lib.stream()), "atq")
        rec["atq_block_gather_us"] = timed(atq, reps=5)[0] * 1e3
        del Wd, partials, ws, E
        out.append(rec)
        print(json.dumps(rec))
        del X, W, H, st, gq
        torch.cuda.empty_cache()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "stage_probe.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
