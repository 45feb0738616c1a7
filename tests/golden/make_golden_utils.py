#!/usr/bin/env python
"""Golden values for the calibration / evaluation helpers: RUN THE UNMODIFIED REFERENCE ``get_calibration_data`` and
``evaluate_perplexity`` (``/root/reference/utils.py:24-72``, ``:127-186``) with ``load_dataset`` replaced by a seeded
in-memory corpus (no network) and the toy tokenizer / model of tests/toy_model.py.

    python tests/golden/make_golden_utils.py          (build container only)

Only the two function definitions (and set_seed) are compiled out of the reference file; nothing is copied."""

import ast
import os
import random
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import toy_model  # noqa: E402

REF = os.environ.get("TQ_REFERENCE_DIR", "/root/reference")


def main():
    import torch
    import torch.nn as nn
    from typing import Any, Dict, List, Optional, Tuple
    text = toy_model.corpus()

    def load_dataset(name, config=None, split=None, **kw):
        return {"text": text.split("q")}                        # the reference joins the pieces with blank lines

    ns = {"torch": torch, "nn": nn, "np": np, "random": random, "List": List, "Dict": Dict, "Optional": Optional,
          "Tuple": Tuple, "Any": Any, "load_dataset": load_dataset}
    tree = ast.parse(open(os.path.join(REF, "utils.py")).read())
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in ("set_seed", "get_calibration_data", "evaluate_perplexity"):
            exec(compile(ast.Module(body=[node], type_ignores=[]), "utils.py", "exec"), ns)
    tok = toy_model.CharTokenizer()
    samples = ns["get_calibration_data"](tok, num_samples=12, seq_len=40, seed=42)
    model = toy_model.ToyCausalLM(toy_model.build())
    joined = "\n\n".join(text.split("q"))
    ppl = {}
    for seq_len in (50, 173, 100000):
        ppl[seq_len] = ns["evaluate_perplexity"](model, tok, seq_len=seq_len, device=torch.device("cpu"))
    np.savez_compressed(os.path.join(HERE, "utils_calibration.npz"), samples=torch.cat(samples, 0).numpy(),
                        n_tokens=np.int64(len(tok(joined)["input_ids"][0])),
                        ppl_seq=np.array(sorted(ppl), dtype=np.int64), ppl=np.array([ppl[k] for k in sorted(ppl)]))
    print("wrote utils_calibration.npz", {k: round(v, 4) for k, v in ppl.items()})


if __name__ == "__main__":
    main()
