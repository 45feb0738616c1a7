#!/usr/bin/env python
"""Golden fixture for the ternary inference layer: RUN THE UNMODIFIED REFERENCE ``TernaryLinear``
(``/root/reference/model.py:17-127``) on the quantised parameters the reference's ``GPTQ.quantize`` produced
(the committed ``gptq_small_{seq,ssr}.npz`` fixtures) and store its outputs.

    python tests/golden/make_golden_tl.py          (build container only)

``model.py`` imports transformers at module level; only the class is needed, so its ``ClassDef`` is compiled out
of the reference file at run time (nothing is copied into the repo).
"""

import ast
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("TQ_REFERENCE_DIR", "/root/reference")


def load_ternary_linear():
    import torch
    import torch.nn as nn
    from typing import Optional
    tree = ast.parse(open(os.path.join(REF, "model.py")).read())
    ns = {"torch": torch, "nn": nn, "Optional": Optional}
    for node in tree.body:
        if isinstance(node, ast.ClassDef) and node.name == "TernaryLinear":
            exec(compile(ast.Module(body=[node], type_ignores=[]), "model.py", "exec"), ns)
    return ns["TernaryLinear"]


def main():
    import torch
    TL = load_ternary_linear()
    out = {}
    rng = np.random.default_rng(501)
    for tag in ("seq", "ssr"):
        g = np.load(os.path.join(HERE, f"gptq_small_{tag}.npz"))
        alpha, mu, T, perm = g["alpha"], g["mu"], g["T"], g["perm"]
        n, m = T.shape
        x = rng.standard_normal((5, m)).astype(np.float32)
        bias = (rng.standard_normal(n) * 0.1).astype(np.float32)
        out[f"x_{tag}"], out[f"bias_{tag}"] = x, bias
        for dt_name, dt in (("float32", torch.float32), ("float16", torch.float16)):
            for with_bias in (False, True):
                layer = TL(m, n, block_size=128, bias=with_bias, dtype=dt)
                layer.set_quantized_params(torch.from_numpy(alpha), torch.from_numpy(mu), torch.from_numpy(T),
                                           torch.from_numpy(perm),
                                           torch.from_numpy(bias) if with_bias else None)
                y = layer(torch.from_numpy(x).to(dt))
                key = f"y_{tag}_{dt_name}_{'bias' if with_bias else 'nobias'}"
                out[key] = y.float().numpy()
                if not with_bias:
                    out[f"W_{tag}_{dt_name}"] = layer._dequantize().float().numpy()
                    out[f"footprint_{tag}_{dt_name}"] = np.int64(layer.memory_footprint())
    np.savez_compressed(os.path.join(HERE, "ternary_linear.npz"), **out)
    print("wrote ternary_linear.npz:", sorted(out))


if __name__ == "__main__":
    main()
