#!/usr/bin/env python
"""The reference against ITSELF at model level: ``PT2LLMQuantizer.quantize()`` of the unmodified reference on the toy
model, run with all host threads and again with ONE thread (same inputs, same code; only the reduction order inside
oneMKL changes).  The spread between the two runs is the floor any other implementation is judged against at model
level (the block sweep amplifies 1e-6 input differences into code flips at thresholds, and layer 1
sees activations that went through layer 0's quantised weights).  Stored per layer index in
``tests/golden/model_toy_floor.json``; ``tests/parity.py`` derives its model-level tolerances from it.

    python tests/golden/make_golden_model_floor.py          (build container only)
"""

import contextlib
import io
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import parity  # noqa: E402
import toy_model  # noqa: E402
from make_golden_model import load_main  # noqa: E402


def run(M, use_ssr, threads):
    import torch
    torch.set_num_threads(threads)
    model = toy_model.build()
    toks = toy_model.samples()
    pq = M.PT2LLMQuantizer(model, None, model_type="llama", use_ssr=use_ssr, device="cpu")
    pq.get_calibration_data = lambda: toks
    with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
        params = pq.quantize()
    return {n: {k: v.numpy() for k, v in p.items()} for n, p in params.items()}


def main():
    M = load_main()
    out = {"how": "unmodified reference quantize() on tests/toy_model.py, all host threads vs 1 thread", "threads": os.cpu_count()}
    for use_ssr in (False, True):
        a, b = run(M, use_ssr, os.cpu_count()), run(M, use_ssr, 1)
        tag = "ssr" if use_ssr else "seq"
        floor = {}
        for name in a:
            layer = name.split(".")[0]
            if use_ssr and layer != "layer_0":       # with SSR only layer 0 of the reference's run is the algorithm's answer
                continue
            f = floor.setdefault(layer, {"code_agreement_min": 1.0, "alpha_rel_err_max": 0.0, "mu_err_rel_alpha_max": 0.0,
                                         "membership_equal": True})
            f["membership_equal"] &= bool(parity.same_block_membership(a[name]["perm"], b[name]["perm"]))
            f["code_agreement_min"] = min(f["code_agreement_min"], parity.code_agreement(a[name]["T"], b[name]["T"]))
            if np.array_equal(a[name]["perm"], b[name]["perm"]):
                mask = parity.block_pairs_agree(a[name]["T"], b[name]["T"], b[name]["perm"], 128)
                ra = np.asarray(b[name]["alpha"], dtype=np.float64)
                mask &= np.isfinite(ra) & (np.abs(ra) < 1e3)
                f["alpha_rel_err_max"] = max(f["alpha_rel_err_max"], parity.scale_rel_err(a[name]["alpha"], b[name]["alpha"], mask))
                f["mu_err_rel_alpha_max"] = max(f["mu_err_rel_alpha_max"],
                                                parity.scale_rel_err(a[name]["mu"], b[name]["mu"], mask, floor=np.abs(ra)))
        out[tag] = floor
        print(tag, json.dumps(floor))
    with open(os.path.join(HERE, "model_toy_floor.json"), "w") as fh:
        json.dump(out, fh, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
