#!/usr/bin/env python
"""Golden fixture for the model-level driver: RUN THE UNMODIFIED REFERENCE ``PT2LLMQuantizer.quantize()``
(``/root/reference/main.py:232-311``) on the seeded toy model of tests/toy_model.py (CPU, fp32) and store every quantised
parameter set and the overwritten weights.

    python tests/golden/make_golden_model.py          (build container only)

``main.py`` imports ``model`` / ``utils`` (transformers, datasets): they are stubbed except for the two pure-Python
walkers it really calls (``get_llm_layers``, ``find_linear_layers``, executed from the reference's own file), and
``get_calibration_data`` is replaced by the seeded token tensors.  Nothing of the reference's source is copied.
"""

import ast
import contextlib
import importlib.util
import io
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import toy_model  # noqa: E402
from make_golden import load_reference, REF  # noqa: E402


def reference_walkers():
    import torch
    import torch.nn as nn
    from typing import Dict, List
    tree = ast.parse(open(os.path.join(REF, "model.py")).read())
    ns = {"torch": torch, "nn": nn, "Dict": Dict, "List": List}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in ("get_llm_layers", "find_linear_layers"):
            exec(compile(ast.Module(body=[node], type_ignores=[]), "model.py", "exec"), ns)
    return ns["get_llm_layers"], ns["find_linear_layers"]


def load_main():
    ref = load_reference()
    get_llm_layers, find_linear_layers = reference_walkers()
    stub_model = types.ModuleType("model")
    for nm in ("load_model_for_quantization", "get_model_type", "compute_model_size", "TernaryLinear",
               "replace_linear_with_ternary"):
        setattr(stub_model, nm, None)
    stub_model.get_llm_layers, stub_model.find_linear_layers = get_llm_layers, find_linear_layers
    stub_utils = types.ModuleType("utils")
    for nm in ("set_seed", "get_calibration_data", "evaluate_perplexity", "save_quantized_model",
               "compute_bits_per_weight", "get_wikitext2", "get_c4", "get_ptb", "load_quantized_model",
               "pack_ternary", "unpack_ternary"):
        setattr(stub_utils, nm, (lambda *a, **k: None))
    saved = {k: sys.modules.get(k) for k in ("model", "utils", "quantizer", "reorder", "gptq", "main")}
    try:
        sys.modules["model"], sys.modules["utils"] = stub_model, stub_utils
        sys.modules["quantizer"], sys.modules["reorder"], sys.modules["gptq"] = ref["quantizer"], ref["reorder"], ref["gptq"]
        spec = importlib.util.spec_from_file_location("main", os.path.join(REF, "main.py"))
        M = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(M)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return M


def main():
    import torch
    torch.set_num_threads(os.cpu_count())
    M = load_main()
    for use_ssr in (False, True):
        model = toy_model.build()
        toks = toy_model.samples()
        pq = M.PT2LLMQuantizer(model, None, model_type="llama", use_ssr=use_ssr, device="cpu")
        pq.get_calibration_data = lambda: toks
        with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
            params = pq.quantize()
        out = {"forward_calls": np.int64(model.forward_calls), "num_samples": np.int64(len(toks))}
        for name, p in params.items():
            # With SSR the reference overwrites layer 0 with a wrongly dequantised weight (main.py:313-335), so
            # only layer 0 of that run is the algorithm's answer; later layers are kept for use_ssr=False only.
            if use_ssr and not name.startswith("layer_0."):
                continue
            for k, v in p.items():
                out[f"{name}/{k}"] = v.numpy()
        # one overwritten weight (the rest follow from the parameters) and the quantised model's logits on sample 0
        out["layer_0.self_attn.q_proj/W_after"] = model.model.layers[0].self_attn.q_proj.weight.detach().numpy().astype(np.float32)
        out["logits_after"] = model(toks[0]).detach().numpy().astype(np.float32)
        np.savez_compressed(os.path.join(HERE, f"model_toy_{'ssr' if use_ssr else 'seq'}.npz"), **out)
        print("wrote", f"model_toy_{'ssr' if use_ssr else 'seq'}.npz", len(params), "linears, model forwards:", model.forward_calls)


if __name__ == "__main__":
    main()
