#!/usr/bin/env python
"""Generate the golden fixtures by RUNNING THE UNMODIFIED REFERENCE (CPU, fp32).

Run in the build container only (``/root/reference`` does not exist on the GPU box):

    python tests/golden/make_golden.py

The reference modules are flat top-level names (``gptq.py:17-18`` does
``from quantizer import ...``) that collide with this repo's own modules, so they are
loaded through an isolated loader and removed from ``sys.modules`` afterwards.  Nothing
of the reference's source is copied; only its outputs on seeded inputs are stored.
"""

import importlib.util
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import synth  # noqa: E402

REF = os.environ.get("TQ_REFERENCE_DIR", "/root/reference")


def load_reference(names=("quantizer", "reorder", "gptq")):
    sys.dont_write_bytecode = True
    saved = {k: sys.modules.get(k) for k in names}
    mods = {}
    try:
        for name in names:
            spec = importlib.util.spec_from_file_location(name, os.path.join(REF, name + ".py"))
            mod = importlib.util.module_from_spec(spec)
            sys.modules[name] = mod
            spec.loader.exec_module(mod)
            mods[name] = mod
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return mods


def main():
    import torch
    import torch.nn as nn

    torch.set_num_threads(os.cpu_count())
    ref = load_reference()
    Q, R, G = ref["quantizer"], ref["reorder"], ref["gptq"]
    atq = Q.AsymmetricTernaryQuantizer(max_iter=100)
    t = torch.from_numpy

    # ---------------------------------------------------------------- A. ATQ stages
    W = synth.make_weight(64, 128, seed=101, row_offset=0.01)
    X = np.random.default_rng(102).standard_normal((16, 128), dtype=np.float32)
    a0, u0, T0 = atq.ternary_init(t(W))
    ag, ug = atq.build_optimal_grid(t(W), T0)
    Tr = atq.flexible_round(t(W), ag, ug)
    a1, u1, T1 = atq.iterative_ternary_fitting(t(W), a0, u0, T0)
    a2, u2 = atq.activation_aware_grid_alignment(t(W), T1, t(X))
    aq, uq, Tq = atq.quantize(t(W), t(X))
    an, un, Tn = atq.quantize(t(W), None)
    np.savez_compressed(
        os.path.join(HERE, "atq_stages.npz"),
        W=W, X=X, checksum=synth.checksum(W, X),
        init_alpha=a0.numpy(), init_mu=u0.numpy(), init_T=T0.numpy().astype(np.int8),
        grid_alpha=ag.numpy(), grid_mu=ug.numpy(), round_T=Tr.numpy().astype(np.int8),
        itf_alpha=a1.numpy(), itf_mu=u1.numpy(), itf_T=T1.numpy().astype(np.int8),
        aga_alpha=a2.numpy(), aga_mu=u2.numpy(),
        q_alpha=aq.numpy(), q_mu=uq.numpy(), q_T=Tq.numpy().astype(np.int8),
        qn_alpha=an.numpy(), qn_mu=un.numpy(), qn_T=Tn.numpy().astype(np.int8),
    )

    # heavy-tailed rows + degenerate rows: constant row, zero row, two-valued row
    rng = np.random.default_rng(103)
    Wd = (rng.standard_t(3, size=(40, 128)).astype(np.float32) * np.float32(0.02))
    Wd[3, :] = 0.03
    Wd[7, :] = 0.0
    Wd[11, :64] = 0.05
    Wd[11, 64:] = -0.05
    ad, ud, Td = atq.quantize(t(Wd), None)
    np.savez_compressed(os.path.join(HERE, "atq_degenerate.npz"), W=Wd,
                        alpha=ad.numpy(), mu=ud.numpy(), T=Td.numpy().astype(np.int8))

    # blocks of other widths (ragged last block / other block sizes)
    out = {}
    for b in (8, 40, 64, 72, 200, 256):
        Wb = synth.make_weight(24, b, seed=110 + b, row_offset=0.005)
        Xb = np.random.default_rng(120 + b).standard_normal((b, b), dtype=np.float32)
        Xb = (Xb @ Xb.T / b + np.eye(b, dtype=np.float32)).astype(np.float32)
        ab, ub, Tb = atq.quantize(t(Wb), t(Xb).unsqueeze(0))
        out[f"W{b}"], out[f"X{b}"] = Wb, Xb
        out[f"alpha{b}"], out[f"mu{b}"], out[f"T{b}"] = ab.numpy(), ub.numpy(), Tb.numpy().astype(np.int8)
    np.savez_compressed(os.path.join(HERE, "atq_widths.npz"), **out)

    # ---------------------------------------------------------------- B. SSR selection
    Ws = synth.make_weight(32, 300, seed=201)
    Ws[:, ::7] += 0.01  # a family of columns sharing a direction
    remaining = np.setdiff1d(np.arange(300), np.arange(5, 300, 11)).astype(np.int64)
    sim = R.compute_column_similarity_to_mean(t(Ws), t(remaining))
    blk, new_rem = R.select_next_block_ssr(t(Ws), t(remaining), 128)
    small = remaining[:100]
    blk2, new_rem2 = R.select_next_block_ssr(t(Ws), t(small), 128)
    np.savez_compressed(os.path.join(HERE, "ssr_select.npz"), W=Ws, remaining=remaining,
                        sim=sim.numpy(), block=blk.numpy(), new_remaining=new_rem.numpy(),
                        small=small, block_small=blk2.numpy(), new_remaining_small=new_rem2.numpy())

    # ---------------------------------------------------------------- C/D. GPTQ class + main.py
    def run_gptq(n, m, samples, seq, seed, use_ssr, lam):
        Wg = synth.make_weight(n, m, seed=seed)
        Xg = synth.make_activations(samples, seq, m, seed=seed + 1000, lam=lam)
        layer = nn.Linear(m, n, bias=False)
        layer.weight.data = t(Wg).clone()
        g = G.GPTQ(layer, block_size=128, percdamp=0.01)
        for i in range(samples):
            if i % 2 == 0:
                g.add_batch(t(Xg[i:i + 1]).float())          # 3-D (1, L, m)
            else:
                g.add_batch(t(Xg[i]).float())                # 2-D (L, m)
        alpha, mu, T, perm = g.quantize(use_ssr=use_ssr)
        Wq = g.get_quantized_weight()
        return Wg, Xg, g, alpha, mu, T, perm, Wq

    for tag, (n, m, samples, seq, seed, lam) in {
        "small": (48, 320, 4, 160, 301, 0.5),       # ragged: blocks 128,128,64; inputs stored
        "mid": (256, 512, 8, 256, 302, 0.5),        # inputs regenerated from the seed
    }.items():
        for use_ssr in (False, True):
            Wg, Xg, g, alpha, mu, T, perm, Wq = run_gptq(n, m, samples, seq, seed, use_ssr, lam)
            rec = dict(n=n, m=m, samples=samples, seq=seq, seed=seed, lam=lam,
                       checksum=synth.checksum(Wg, Xg),
                       nsamples=g.nsamples,
                       alpha=alpha.numpy(), mu=mu.numpy(), T=T.numpy().astype(np.int8),
                       perm=perm.numpy(), Wq=Wq.numpy().astype(np.float32))
            if tag == "small":
                rec.update(W=Wg, X=Xg, H=g.H.numpy())
            np.savez_compressed(os.path.join(HERE, f"gptq_{tag}_{'ssr' if use_ssr else 'seq'}.npz"), **rec)

    # main.py's inline loop (second oracle, AGA on raw activations).  main.py cannot be imported
    # here without its model/dataset dependencies pulling in network-facing packages, so the
    # method is taken from the class by importing main with a stubbed ``model``/``utils``.
    import types
    stub_model = types.ModuleType("model")
    for nm in ("load_model_for_quantization", "get_llm_layers", "find_linear_layers",
               "get_model_type", "compute_model_size", "TernaryLinear", "replace_linear_with_ternary"):
        setattr(stub_model, nm, None)
    stub_utils = types.ModuleType("utils")
    for nm in ("set_seed", "get_calibration_data", "evaluate_perplexity", "save_quantized_model",
               "compute_bits_per_weight", "get_wikitext2", "get_c4", "get_ptb", "load_quantized_model",
               "pack_ternary", "unpack_ternary"):
        setattr(stub_utils, nm, (lambda *a, **k: None))
    saved = {k: sys.modules.get(k) for k in ("model", "utils", "quantizer", "reorder", "gptq", "main")}
    try:
        sys.modules["model"], sys.modules["utils"] = stub_model, stub_utils
        sys.modules["quantizer"], sys.modules["reorder"], sys.modules["gptq"] = Q, R, G
        spec = importlib.util.spec_from_file_location("main", os.path.join(REF, "main.py"))
        M = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(M)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    import contextlib
    import io
    for use_ssr in (False, True):
        n, m, samples, seq, seed, lam = 48, 320, 4, 160, 301, 0.5
        Wg = synth.make_weight(n, m, seed=seed)
        Xg = synth.make_activations(samples, seq, m, seed=seed + 1000, lam=lam)
        layer = nn.Linear(m, n, bias=False)
        layer.weight.data = t(Wg).clone()
        pq = M.PT2LLMQuantizer(None, None, use_ssr=use_ssr, device="cpu")
        with contextlib.redirect_stdout(io.StringIO()):
            res = pq.quantize_layer(layer, "l", t(Xg).float())
        np.savez_compressed(os.path.join(HERE, f"main_small_{'ssr' if use_ssr else 'seq'}.npz"),
                            checksum=synth.checksum(Wg, Xg), seed=seed,
                            alpha=res["alpha"].numpy(), mu=res["mu"].numpy(),
                            T=res["T"].numpy(), perm=res["perm"].numpy())

    # ---------------------------------------------------------------- E. pack / unpack
    # utils.py imports ``datasets``; the two codec functions only need torch, so they are
    # executed from the reference file by extracting just those two defs at run time.
    import ast
    src = open(os.path.join(REF, "utils.py")).read()
    tree = ast.parse(src)
    from typing import Tuple
    ns = {"torch": torch, "Tuple": Tuple}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in ("pack_ternary", "unpack_ternary"):
            exec(compile(ast.Module(body=[node], type_ignores=[]), "utils.py", "exec"), ns)
    rng = np.random.default_rng(401)
    packs = {}
    for i, shape in enumerate([(7, 13), (4, 128), (1, 1), (3, 5, 2), (33, 257)]):
        Tp = rng.integers(-1, 2, size=shape).astype(np.float32)
        packed, oshape = ns["pack_ternary"](t(Tp))
        un = ns["unpack_ternary"](packed, oshape)
        packs[f"T{i}"], packs[f"packed{i}"], packs[f"unpacked{i}"] = Tp.astype(np.int8), packed.numpy(), un.numpy()
    np.savez_compressed(os.path.join(HERE, "pack.npz"), **packs)

    # ---------------------------------------------------------------- F. examples.py known answers
    torch.manual_seed(42)
    We = torch.randn(256, 512)
    Xe = torch.randn(32, 512)
    a0, u0, T0 = atq.ternary_init(We)
    e_init = Q.compute_quantization_error(We, atq.dequantize(a0, u0, T0))
    a1, u1, T1 = atq.iterative_ternary_fitting(We, a0, u0, T0)
    e_itf = Q.compute_quantization_error(We, atq.dequantize(a1, u1, T1))
    a2, u2 = atq.activation_aware_grid_alignment(We, T1, Xe)
    o_before = Q.compute_output_error(We, atq.dequantize(a1, u1, T1), Xe)
    o_after = Q.compute_output_error(We, atq.dequantize(a2, u2, T1), Xe)
    np.savez_compressed(os.path.join(HERE, "examples_known.npz"),
                        e_init=e_init, e_itf=e_itf, o_before=o_before, o_after=o_after)
    print("example 1:", e_init, e_itf, o_before, o_after)
    print("fixtures written to", HERE)
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f"  {f}: {os.path.getsize(os.path.join(HERE, f))} bytes")


if __name__ == "__main__":
    main()
