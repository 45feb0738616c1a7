#!/usr/bin/env python
"""Golden fixture for the Cholesky-failure route (gptq.py:101-106): RUN THE UNMODIFIED REFERENCE on a Hessian that is
symmetric but NOT positive definite, so ``torch.linalg.cholesky`` raises and the reference falls back to
``torch.linalg.pinv``.  Build container only (``/root/reference`` does not exist on the GPU box):

    python tests/golden/make_golden_pinv.py

H = Q diag(d) Q' with d in [0.5, 2] except ONE eigenvalue of -1: well conditioned (pinv(H) = H^-1, no rank decision
involved) and indefinite, with diag(H^-1) still positive so the feedback's clamp (gptq.py:177) stays inert.  The
reference's outputs on seeded inputs are stored; none of its source is."""

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
import synth  # noqa: E402
from make_golden import load_reference  # noqa: E402


def indefinite_hessian(m, seed):
    rng = np.random.default_rng(seed)
    Q, _ = np.linalg.qr(rng.standard_normal((m, m)))
    d = rng.uniform(0.5, 2.0, size=m)
    d[m // 3] = -1.0
    H = (Q * d) @ Q.T
    return np.ascontiguousarray(((H + H.T) / 2).astype(np.float32))


def main():
    import torch
    import torch.nn as nn
    torch.set_num_threads(os.cpu_count())
    G = load_reference()["gptq"]
    n, m = 96, 384
    W = synth.make_weight(n, m, seed=901)
    H = indefinite_hessian(m, seed=902)
    out = {"W": W, "H": H, "checksum": synth.checksum(W, H)}
    for tag, use_ssr in (("seq", False), ("ssr", True)):
        layer = nn.Linear(m, n, bias=False)
        layer.weight.data = torch.from_numpy(W.copy())
        g = G.GPTQ(layer, block_size=128, percdamp=0.01)
        g.H = torch.from_numpy(H.copy())
        g.nsamples = 1
        try:
            torch.linalg.cholesky(g.H / 1 + 0.01 * torch.diag(g.H).mean() * torch.eye(m))
            raise SystemExit("the damped matrix is positive definite: the fixture would not exercise the pinv route")
        except RuntimeError:
            pass
        alpha, mu, T, perm = g.quantize(use_ssr=use_ssr)
        out.update({f"{tag}_alpha": alpha.numpy(), f"{tag}_mu": mu.numpy(), f"{tag}_T": T.numpy().astype(np.int8),
                    f"{tag}_perm": perm.numpy()})
        print(tag, "alpha range", float(alpha.min()), float(alpha.max()), "finite", bool(torch.isfinite(alpha).all()))
    np.savez_compressed(os.path.join(HERE, "gptq_pinv.npz"), **out)


if __name__ == "__main__":
    main()
