#!/usr/bin/env python
"""Which stage of the CUDA path carries the numerical noise that tips rows across ITF threshold ties?

One bench-shape linear (default 4096 x 11008, sequential order) is quantised several times with ONE stage at a time
replaced by an fp32 library stand-in (diagnostic only -- nothing here is a product path):

    base     tcgen05 Hessian + split-TF32 Cholesky inverse + split-TF32 feedback (the product)
    H        Hessian = X'X by an fp32 library GEMM (TF32 off), rest as the product
    H+inv    + inverse by torch.linalg.cholesky / cholesky_inverse (cuSOLVER fp32) injected into the state's cache
    H+inv+fb + CUDA-core fp32 feedback (TQ_SWEEP_FFMA_FEEDBACK)
    inv      product Hessian, library inverse

and each result is adjudicated against oracle/torch_port.py run on the GPU in fp64 (the common yardstick).  Output:
rows that left the oracle's trajectory, disagreeing (row, block) pairs, code agreement -> gpurun_out/noise_probe.json.

    python scripts/noise_probe.py [n m]
"""

import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import parity  # noqa: E402
import tq100  # noqa: E402
from oracle import torch_port  # noqa: E402
from tq100 import _lib  # noqa: E402
from tq100.pipeline import LinearView  # noqa: E402

DEV = "cuda:0"
NT = 16384


def main():
    n, m = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (4096, 11008)
    order = sys.argv[3] if len(sys.argv) > 3 else "sequential"
    torch.backends.cuda.matmul.allow_tf32 = False
    g = torch.Generator(device=DEV).manual_seed(4000 + m)
    r, lam = 64, 0.5
    B = torch.randn((r, m), device=DEV, generator=g)
    X = (torch.randn((NT, m), device=DEV, generator=g)
         + (lam / r ** 0.5) * (torch.randn((NT, r), device=DEV, generator=g) @ B)).to(torch.float16)
    gw = torch.Generator(device=DEV).manual_seed(17 * n + m)
    W = torch.randn((n, m), device=DEV, generator=gw) * 0.02
    Xd = X.double()
    H64 = Xd.T @ Xd
    H32 = X.float().T @ X.float()
    del Xd
    use_ssr = order == "ssr"
    a, u, T, p, aids = torch_port.quantize_layer(W.double(), H64, NT, 128, 0.01, use_ssr=use_ssr, return_margin=True)
    ref = dict(alpha=a.cpu().numpy(), mu=u.cpu().numpy(), T=T.cpu().numpy(), perm=p.cpu().numpy())
    aids = {k: (v.cpu().numpy() if torch.is_tensor(v) else v) for k, v in aids.items()}
    out = {"shape": [n, m], "order": order, "tokens": NT, "yardstick": "oracle/torch_port.py on the GPU in fp64"}

    def adjudicate(tag, alpha, mu, T_, perm):
        got = dict(alpha=alpha.cpu().numpy(), mu=mu.cpu().numpy(), T=T_.cpu().numpy(), perm=perm.cpu().numpy())
        rep = parity.adjudicate(got, ref, aids)
        keep = {k: rep.get(k) for k in ("code_agreement", "rows_diverged", "pairs_disagreeing", "pairs",
                                         "leading_blocks_same_membership", "alpha_rel_err_max")}
        tie = rep.get("tie_margin_first_divergence")
        if tie:
            keep["tie_margin_median"], keep["tie_margin_max"] = tie["median"], tie["max"]
        out[tag] = keep
        print(tag, json.dumps(keep), flush=True)

    # the fp32 oracle (library everything) against the fp64 one: the reference's own floor
    a, u, T, p = torch_port.quantize_layer(W, H32, NT, 128, 0.01, use_ssr=use_ssr)
    adjudicate("oracle_fp32", a, u, T, p)

    def product(tag, lib_H, lib_inv, ffma):
        q = tq100.GPTQ(LinearView(W))
        if lib_H:
            q.H = H32.clone()
            q.nsamples = NT
        else:
            q.add_batch(X.reshape(NT // 2048, 2048, m))
        if lib_inv:
            Hd = q.state.damped(q.percdamp)
            Hinv = torch.cholesky_inverse(torch.linalg.cholesky(Hd))
            q.state._cache[float(q.percdamp)] = (Hd, Hinv.contiguous(), torch.zeros(1, dtype=torch.int32, device=DEV))
        if ffma:
            q.sweep_flags = _lib.SWEEP_FFMA_FEEDBACK
        alpha, mu, T_, perm = q.quantize(use_ssr=use_ssr)
        if not lib_inv:
            Hd, Hinv, _ = q.state.damped_inverse(q.percdamp)
        v = torch.randn((m, 8), dtype=torch.float64, device=DEV)
        out.setdefault("inverse_residual", {})[tag] = float((Hd.double() @ (Hinv.double() @ v) - v).abs().max())
        adjudicate(tag, alpha, mu, T_, perm)

    product("base", False, False, False)
    product("H", True, False, False)
    product("H+inv", True, True, False)
    product("H+inv+fb", True, True, True)
    product("inv", False, True, False)
    product("fb", False, False, True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", f"noise_probe_{n}x{m}_{order}.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
