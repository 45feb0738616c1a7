#!/usr/bin/env python
"""Wall time (CUDA events) of one 7B-shaped layer through LayerDriver, for quick A/B runs of scheduling knobs
(environment: TQ_CHAIN_STREAMS, TQ_CHAIN_PRIORITY, TQ_EAGER_CHAINS, TQ_HESS_SPARE_SMS, TQ_GX_SPARE_SMS).
    python scripts/layer_time.py [--shared] [--order ssr] [--reps 5]"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tq100  # noqa: E402,F401
from tq100.pipeline import LayerDriver  # noqa: E402

DEV = torch.device("cuda:0")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shared", action="store_true")
    ap.add_argument("--order", default="ssr")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--d", type=int, default=4096)
    ap.add_argument("--ffn", type=int, default=11008)
    ap.add_argument("--only", default=None, help="comma-separated linear names (default: all seven)")
    ap.add_argument("--no-hessian", action="store_true", help="time the chains only (Hessians accumulated beforehand)")
    args = ap.parse_args()
    d, f = args.d, args.ffn
    lins = [("q_proj", d, d, "attn_in"), ("k_proj", d, d, "attn_in"), ("v_proj", d, d, "attn_in"),
            ("o_proj", d, d, "o_in"), ("gate_proj", f, d, "mlp_in"), ("up_proj", f, d, "mlp_in"),
            ("down_proj", d, f, "down_in")]
    gen = torch.Generator(device=DEV).manual_seed(1)
    acts = {}
    for key, width in (("attn_in", d), ("o_in", d), ("mlp_in", d), ("down_in", f)):
        B = torch.randn((64, width), device=DEV, generator=gen)
        x = torch.empty((128, 2048, width), device=DEV, dtype=torch.float16)
        for j in range(128):
            x[j] = (torch.randn((2048, width), device=DEV, generator=gen)
                    + 0.0625 * (torch.randn((2048, 64), device=DEV, generator=gen) @ B)).half()
        acts[key] = x
    ws = {name: torch.randn((n, m), device=DEV, generator=gen) * 0.02 for name, n, m, _ in lins}
    drv = LayerDriver(DEV, share_inputs=args.shared)
    if args.only:
        lins = [l for l in lins if l[0] in args.only.split(",")]
    states = None
    if args.no_hessian:
        from tq100.gptq import HessianState
        states = {}
        for name, n, m, src in lins:
            states[name] = HessianState(m, DEV)
            states[name].add_batch(acts[src])

    def layer():
        if states is not None:
            from tq100.gptq import GPTQ
            from tq100.pipeline import LinearView
            gs = []
            for name, n, m, src in lins:
                states[name]._cache.clear()
                gs.append(GPTQ(LinearView(ws[name]), hessian=states[name]))
            return drv.run_chains(gs, use_ssr=args.order == "ssr", order=args.order)
        return drv.quantize([(name, ws[name], acts[src]) for name, n, m, src in lins], use_ssr=args.order == "ssr",
                            order=args.order)

    for _ in range(2):
        layer()
    torch.cuda.synchronize()
    ts = []
    for _ in range(args.reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        layer()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    knobs = {k: os.environ.get(k) for k in ("TQ_CHAIN_STREAMS", "TQ_CHAIN_PRIORITY", "TQ_EAGER_CHAINS", "TQ_HESS_SPARE_SMS",
                                           "TQ_GX_SPARE_SMS", "CUDA_DEVICE_MAX_CONNECTIONS") if os.environ.get(k) is not None}
    print(json.dumps({"layer_ms_min": min(ts), "layer_ms_median": sorted(ts)[len(ts) // 2], "shared": args.shared, "only": args.only,
                      "chains_only": args.no_hessian, "order": args.order, "knobs": knobs}), flush=True)


if __name__ == "__main__":
    main()
