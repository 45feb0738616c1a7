#!/usr/bin/env python
"""Headline benchmark: LLaMA-2-7B-shaped ternary GPTQ (PT2-LLM) wall-time on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the hot path over the whole LLaMA-2-7B-shaped model: for each of the 32
transformer layers and each of its 7 linears, GPTQ.add_batch (Hessian over 128 x 2048 calibration
tokens) + GPTQ.quantize (damped Cholesky inverse, SSR column sweep with ITF/AGA grid fit and error
feedback) -- BASELINE.json configs[1].  Synthetic data: random-init weights N(0, 0.02^2) per linear;
one layer's four distinct calibration inputs (attention in, o_proj in, MLP in, down_proj in: 12.2 GB
fp16, far larger than L2) are reused for the 32 layers because 32 x 12.2 GB does not fit in HBM.

Prints ONE JSON line (see the keys at the bottom).  `value` times the device-resident path with CUDA
events; `e2e` drives the same public API from pinned HOST buffers (weights and activations copied
host->device, results copied device->host, every step, inside the timed region).
"""

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

LLAMA2_7B = dict(name="LLaMA-2-7B", d=4096, ffn=11008, layers=32)
LLAMA2_13B = dict(name="LLaMA-2-13B", d=5120, ffn=13824, layers=40)
SAMPLES, SEQ = 128, 2048
# (linear name, out_features n, in_features m, which of the layer's inputs it reads)
def linears(cfg):
    d, f = cfg["d"], cfg["ffn"]
    return [("q_proj", d, d, "attn_in"), ("k_proj", d, d, "attn_in"), ("v_proj", d, d, "attn_in"),
            ("o_proj", d, d, "o_in"), ("gate_proj", f, d, "mlp_in"), ("up_proj", f, d, "mlp_in"),
            ("down_proj", d, f, "down_in")]


def hessian_useful_flops(nt, m):
    return nt * m * (m + 1)          # SYRK, upper triangle; dense-equivalent is 2*nt*m*m


# ------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.gpu_index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 8:
                continue
            try:
                sm.append(float(parts[1]))
                mx = float(parts[2])
            except ValueError:
                continue
            for nm, v in zip(names, parts[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        # median over the samples taken under load (above 60 % of the peak sample filters idle gaps)
        loaded = [x for x in sm if sm and x >= 0.6 * sm[-1]]
        med = loaded[len(loaded) // 2] if loaded else None
        return {"sm_mhz": med, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------ reference arm
def find_reference_dir():
    """The unmodified reference, if this machine has it: $TQ_REFERENCE_DIR, baseline/_ref, /root/reference.
    It is a flat directory of Python scripts, not an installable package, and does not exist on the GPU box; there the
    oracle's port (oracle/torch_port.py, pinned to the reference's outputs by tests/test_oracle_golden.py) stands in."""
    if os.environ.get("TQ_BENCH_FORCE_PORT"):            # A/B of the port against the reference on the same cores
        return None
    for d in (os.environ.get("TQ_REFERENCE_DIR"), os.path.join(ROOT, "baseline", "_ref"), "/root/reference"):
        if d and os.path.isfile(os.path.join(d, "gptq.py")) and os.path.isfile(os.path.join(d, "quantizer.py")):
            return d
    return None


def load_reference_gptq(ref_dir):
    """Import the reference's gptq module under an isolated loader: its modules are flat top-level names (gptq.py:17-18
    does `from quantizer import ...`) that collide with this repo's mirror; nothing is written into the reference dir."""
    import importlib.util
    sys.dont_write_bytecode = True
    names = ("quantizer", "reorder", "gptq")
    saved = {k: sys.modules.get(k) for k in names}
    mods = {}
    try:
        for name in names:
            spec = importlib.util.spec_from_file_location(name, os.path.join(ref_dir, name + ".py"))
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
    return mods["gptq"]


class ReferenceRunner:
    """The reference's implementation of the path for ONE linear: GPTQ.add_batch per calibration sample (gptq.py:59-76)
    and GPTQ.quantize (gptq.py:78-199), fp32, on `device` ('cpu': all host threads; 'cuda': the reference's own default
    device, main.py:368, TF32 off).  kind 'reference' = the unmodified reference's class; 'port' = oracle/torch_port.py,
    the same ATen call sequence."""

    def __init__(self, device="cpu", threads=None, order="ssr"):
        import torch
        self.torch = torch
        self.device = torch.device(device)
        self.order = order
        self.cores = threads or os.cpu_count()
        if self.device.type == "cpu":
            torch.set_num_threads(self.cores)
        else:
            torch.backends.cuda.matmul.allow_tf32 = False
            torch.backends.cudnn.allow_tf32 = False
        self.ref_dir = find_reference_dir()
        self.kind = "reference" if self.ref_dir else "port"
        if self.ref_dir:
            self.G = load_reference_gptq(self.ref_dir)
        else:
            from oracle import torch_port
            self.port = torch_port

    def _sync(self):
        if self.device.type == "cuda":
            self.torch.cuda.synchronize(self.device)

    def run(self, W, samples):
        """W (n, m) fp32 on the device; samples: iterable of (L, m) activations (any float dtype, on the device).
        Returns (hessian_seconds, quantize_seconds, outputs dict)."""
        torch = self.torch
        n, m = W.shape
        use_ssr = self.order == "ssr"
        static_perm = None
        if self.kind == "reference" and self.order != "actorder":
            layer = torch.nn.Linear(m, n, bias=False, device=self.device)
            layer.weight.data = W
            g = self.G.GPTQ(layer, block_size=128, percdamp=0.01)
            self._sync()
            t0 = time.perf_counter()
            for x in samples:
                g.add_batch(x.float())
            self._sync()
            t_h = time.perf_counter() - t0
            t0 = time.perf_counter()
            alpha, mu, T, perm = g.quantize(use_ssr=use_ssr)
            self._sync()
            t_q = time.perf_counter() - t0
        else:
            from oracle import torch_port
            H = torch.zeros((m, m), device=self.device)
            ns = 0
            self._sync()
            t0 = time.perf_counter()
            for x in samples:
                ns += torch_port.hessian_add(H, x.float())
            self._sync()
            t_h = time.perf_counter() - t0
            t0 = time.perf_counter()
            if self.order == "actorder":
                # the act-order extension's oracle: the reference's sequential sweep in descending-diag(H) order
                static_perm = torch.argsort(torch.diagonal(H), descending=True, stable=True)
            alpha, mu, T, perm = torch_port.quantize_layer(W, H, ns, 128, 0.01, use_ssr=use_ssr, static_perm=static_perm)
            self._sync()
            t_q = time.perf_counter() - t0
        return t_h, t_q, dict(alpha=alpha, mu=mu, T=T, perm=perm)


def synth_linear_cpu(torch, n, m, tokens, seed):
    """W ~ N(0, 0.02^2); activations Z + 0.5 F B / 8 rounded once to fp16, as (tokens / SEQ) samples."""
    g = torch.Generator().manual_seed(seed)
    W = torch.randn((n, m), generator=g) * 0.02
    Bm = torch.randn((64, m), generator=g)
    X = torch.randn((tokens, m), generator=g) + (0.5 / 8.0) * (torch.randn((tokens, 64), generator=g) @ Bm)
    X = X.half().float()
    return W, [X[i:i + SEQ] for i in range(0, tokens, SEQ)]


def cpu_layer_sample(cfg, order, steps=1, budget_s=None, hess_tokens=16384, threads=None, log=None):
    """The reference arm's bounded sample: ONE transformer layer's 7 linears on the host cores.
    Per linear: add_batch over `hess_tokens` of the 262144 calibration tokens (cost exactly linear in tokens,
    gptq.py:75) and the FULL quantize().  `steps` passes rotate over the linears -- step s measures linear s mod 7 (with
    fewer than 7 steps, several per step) -- so every linear is measured at least once; repeat measurements stop when
    `budget_s` of wall clock is used up.  Model estimate = layers x sum over linears of (mean add_batch x tokens factor +
    mean quantize)."""
    import torch
    runner = ReferenceRunner("cpu", threads=threads, order=order)
    lins = linears(cfg)
    k = max(1, steps)
    plan = [[j for j in range(len(lins)) if j % k == s] for s in range(k)] if k < len(lins) else \
           [[s % len(lins)] for s in range(k)]
    t_h = {j: [] for j in range(len(lins))}
    t_q = {j: [] for j in range(len(lins))}
    t_start = time.perf_counter()
    steps_measured = 0
    for s, js in enumerate(plan):
        covered = all(t_q[j] for j in range(len(lins)))
        if covered and budget_s is not None and time.perf_counter() - t_start > budget_s:
            break
        for j in js:
            name, n, m, _ = lins[j]
            W, samples = synth_linear_cpu(torch, n, m, hess_tokens, seed=1000 + j)
            th, tq, _ = runner.run(W, samples)
            t_h[j].append(th)
            t_q[j].append(tq)
            if log:
                log(f"[reference arm] step {s} {name} {n}x{m}: add_batch({hess_tokens} tokens) {th:.2f} s, quantize {tq:.2f} s")
        steps_measured += 1
    tok_factor = SAMPLES * SEQ / hess_tokens
    mean = lambda v: sum(v) / len(v)
    per_layer_h = sum(mean(t_h[j]) for j in t_h if t_h[j]) * tok_factor
    per_layer_q = sum(mean(t_q[j]) for j in t_q if t_q[j])
    measured_layer = sum(mean(t_h[j]) + mean(t_q[j]) for j in t_h if t_h[j])
    total = cfg["layers"] * (per_layer_h + per_layer_q)
    detail = {lins[j][0]: {"shape": [lins[j][1], lins[j][2]], "add_batch_s": mean(t_h[j]), "quantize_s": mean(t_q[j]),
                           "measurements": len(t_q[j])} for j in t_h if t_h[j]}
    src = runner.ref_dir if runner.kind == "reference" else "oracle/torch_port.py"
    return {"value": total, "unit": "s", "cores": runner.cores, "kind": runner.kind, "source": src,
            "sample": (f"{'unmodified reference gptq.GPTQ' if runner.kind == 'reference' else 'oracle/torch_port.py'} "
                       f"(torch CPU fp32, {runner.cores} threads) on ONE transformer layer's 7 linears: add_batch over "
                       f"{hess_tokens} of {SAMPLES * SEQ} tokens per linear + the full quantize(order={order}); "
                       f"measured {measured_layer:.1f} s per layer pass; estimate = {cfg['layers']} layers x "
                       f"(add_batch x {tok_factor:.0f} + quantize)"),
            "extrapolated": True,
            "extrapolation": {"measured_layer_s": measured_layer, "tokens_factor": tok_factor, "layers_factor": cfg["layers"],
                              "per_layer_add_batch_s": per_layer_h, "per_layer_quantize_s": per_layer_q,
                              "overall_factor": total / measured_layer if measured_layer > 0 else None,
                              "steps_measured": steps_measured, "wall_s": time.perf_counter() - t_start},
            "per_linear": detail}


def run_reference(args, rank):
    """`--impl reference`: rank 0 alone runs; the other ranks exit 0 without work."""
    if rank != 0:
        return
    cfg, order = bench_config(args)
    import torch
    if args.warmup > 0:                      # thread pool + BLAS warm-up on a small layer (not part of any measurement)
        W, samples = synth_linear_cpu(torch, 256, 512, 2048, seed=1)
        ReferenceRunner("cpu", order=order).run(W, samples)
    budget = float(os.environ.get("TQ_REF_BUDGET_S", 200))
    r = cpu_layer_sample(cfg, order, steps=args.steps, budget_s=budget,
                         log=(lambda m: print(m, file=sys.stderr, flush=True)))
    v = r["value"]
    line = {"impl": "reference", "metric": METRIC[args.config], "value": v, "unit": "s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": v * 1e3,
            "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": workload_config(args, reference=True),
            "extrapolated": True,
            "extrapolation_note": ("value is an ESTIMATE of the whole-model time from one measured transformer layer: "
                                   f"add_batch measured on 1/{r['extrapolation']['tokens_factor']:.0f} of the tokens (linear in "
                                   f"tokens) and every layer being identical ({cfg['layers']} layers); measured wall "
                                   f"{r['extrapolation']['wall_s']:.0f} s; steps x ms_per_step therefore exceeds the run time"),
            "cpu_baseline": r,
            "e2e": {"value": v, "unit": "s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def bench_config(args):
    """(model shape dict, sweep order) of --config."""
    if args.config == "13b-actorder":
        return dict(LLAMA2_13B, layers=args.layers or LLAMA2_13B["layers"]), "actorder"
    return dict(LLAMA2_7B, layers=args.layers or LLAMA2_7B["layers"]), "ssr"


METRIC = {"7b-ssr": "LLaMA-2-7B ternary PTQ wall-time", "13b-actorder": "LLaMA-2-13B ternary PTQ wall-time",
          "layer-sweep": "single-layer sweep: Hessian + column-sweep kernels vs roofline"}


def workload_config(args, reference=False):
    cfg, order = bench_config(args)
    d, f, L = cfg["d"], cfg["ffn"], cfg["layers"]
    act_gb = SAMPLES * SEQ * (3 * d + f) * 2 / 1e9
    w_gb = L * (4 * d * d + 3 * d * f) * 4 / 1e9
    what = {"ssr": "SSR column reordering", "actorder": "act-order (descending diag(H)) column order"}[order]
    return {"workload": f"{cfg['name']}-shaped ternary GPTQ with {what} "
                        f"({L} layers x {{q,k,v,o {d}x{d}; gate,up {f}x{d}; down {d}x{f}}}, "
                        "128x2048 calibration tokens per linear, block 128, percdamp 0.01, ITF + AGA(hessian))",
            "baseline_config": "configs[1]" if args.config == "7b-ssr" else "configs[3]", "order": order, "aga": "hessian",
            "block_size": 128,
            "activations": f"fp16, one layer's 4 distinct inputs ({act_gb:.1f} GB) reused for all {L} layers",
            "cache": f"inputs ({act_gb:.1f} GB activations + {w_gb:.1f} GB weights) far exceed the 126 MB L2; no explicit flush",
            "hessians_per_layer": 7, "streams": args.streams, "parallelism": "1 GPU" if args.gpus == 1 else (
                f"{args.gpus} GPUs: Hessian sample-sharded, NCCL reduce of each H onto the rank that owns the linear, "
                "whole linears dealt to ranks (no collective inside their sweeps); a linear whose chain exceeds 1.5x a "
                "rank's fair share (down_proj from 4 GPUs up) is split by rows after its owner broadcasts H^-1"
                if getattr(args, "shard_mode", "linears") == "linears" else
                f"{args.gpus} GPUs: Hessian sample-sharded + NCCL allreduce, H^-1 dealt + broadcast, sweep row-sharded "
                "(SSR statistics all-reduced per block)")}


# ------------------------------------------------------------------------------------ parity of the timed run
def _np(t):
    return t.detach().cpu().numpy()


def parity_summary(got, ref, aids=None):
    """SURVEY 8c adjudication (tests/parity.py) of one linear, reduced to the fields a bench line can carry."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import parity
    rep = parity.adjudicate({k: _np(v) for k, v in got.items()}, {k: _np(v) for k, v in ref.items()}, aids)
    keep = {k: rep.get(k) for k in ("n", "m", "code_agreement", "perm_equal", "blocks", "leading_blocks_same_membership",
                                     "first_block_same_membership", "pairs", "pairs_disagreeing", "rows_diverged",
                                     "alpha_rel_err_max", "mu_err_rel_alpha_max")}
    lead = rep["leading_blocks_same_membership"]
    if 0 < lead < rep["blocks"]:
        import numpy as np
        cols = _np(ref["perm"])[:lead * 128]
        keep["code_agreement_leading_blocks"] = float((_np(got["T"])[:, cols] == _np(ref["T"])[:, cols]).mean())
    return keep


def recon_error(torch, W, alpha, mu, T, perm, H, block=128):
    """||(W - Wq) X|| / ||W X|| through H = X'X, fp64, on the device (gptq.py:201-230 dequantisation on both sides)."""
    Wq = torch.empty_like(W, dtype=torch.float64)
    perm = perm.long()
    for k in range(alpha.shape[1]):
        cols = perm[k * block:(k + 1) * block]
        Wq[:, cols] = alpha[:, k:k + 1].double() * T[:, cols].double() + mu[:, k:k + 1].double()
    D = W.double() - Wq
    Hd = H.double()
    return float(torch.sqrt((D @ Hd * D).sum() / ((W.double() @ Hd) * W.double()).sum()))


# ------------------------------------------------------------------------------------ same-box reference + parity
def reference_cuda_leg(torch, args, cfg, order, lins, w0, acts, driver, dev):
    runner = ReferenceRunner(dev, order=order)
    outs, detail = {}, {}
    t_h_all = t_q_all = 0.0
    for name, n, m, src in lins:
        x = acts[src]
        th, tq, out = runner.run(w0[name].clone(), (x[i] for i in range(x.shape[0])))
        outs[name] = out
        detail[name] = {"shape": [n, m], "add_batch_s": th, "quantize_s": tq}
        t_h_all += th
        t_q_all += tq
    layer_s = t_h_all + t_q_all
    hess_flops = sum(2.0 * SAMPLES * SEQ * m * m for _, _, m, _ in lins)
    ref = {"value": cfg["layers"] * layer_s, "unit": "s", "device": torch.cuda.get_device_name(dev), "kind": runner.kind,
           "measured_layer_s": layer_s, "layers_factor": cfg["layers"], "extrapolated": True,
           "add_batch_s_per_layer": t_h_all, "quantize_s_per_layer": t_q_all,
           "add_batch_tflops_dense_fp32": hess_flops / t_h_all / 1e12,
           "note": (f"{'unmodified reference gptq.GPTQ' if runner.kind == 'reference' else 'oracle/torch_port.py'} with "
                    "device='cuda' (main.py:368), fp32, TF32 off: one transformer layer's 7 linears, all "
                    f"{SAMPLES * SEQ} tokens, order={order}; value = measured layer x {cfg['layers']} layers"),
           "per_linear": detail}
    if args.no_parity:
        return ref, None
    # the product on the same layer-0 inputs (one untimed pass of the bench's own inner loop)
    gs = driver.quantize([(name, w0[name], acts[src]) for name, n, m, src in lins], use_ssr=order == "ssr", order=order)
    torch.cuda.synchronize()
    par = {"against": ref["note"].split(":")[0], "tolerances": "north_star: codes >= 0.999, scales 1e-4, recon 1e-3 (relative)",
           "linears": {}}
    worst_code, worst_recon = 1.0, 0.0
    for (name, n, m, src), g in zip(lins, gs):
        o = outs[name]
        got = dict(alpha=g.alpha, mu=g.mu, T=g.T_int8, perm=g.perm)
        rep = parity_summary(got, o)
        H = g.H
        e_got = recon_error(torch, w0[name], g.alpha, g.mu, g.T_int8, g.perm, H)
        e_ref = recon_error(torch, w0[name], o["alpha"], o["mu"], o["T"], o["perm"], H)
        rep["recon_got"], rep["recon_ref"] = e_got, e_ref
        rep["recon_rel_diff"] = abs(e_got - e_ref) / e_ref
        par["linears"][name] = rep
        code = rep.get("code_agreement_leading_blocks", rep["code_agreement"])
        worst_code, worst_recon = min(worst_code, code), max(worst_recon, rep["recon_rel_diff"])
    par["min_code_agreement_on_comparable_blocks"] = worst_code
    par["max_recon_rel_diff"] = worst_recon
    par["ok"] = bool(worst_code >= 0.999 and worst_recon <= 1e-3)
    return ref, par


def sharded_parity_leg(torch, dist, tq100, args, ctx, order, lins, w_last, acts, sharded_out, sharded_layer):
    """N > 1: the sharded run's results of the last timed layer against the plain single-GPU path on the same inputs.
    For one whole-owner linear (o_proj) and for down_proj (split by rows from 4 GPUs up) the local Hessians are reduced
    exactly as the sharded path reduces them (NCCL reduce onto the owner / all-reduce for a split linear), the ranks
    that hold results quantise the linear with the single-GPU GPTQ and compare THEIR part of the sharded output: sweep
    order, codes (over the leading blocks whose membership agrees -- one swapped SSR top-k boundary legitimately
    de-correlates what follows), scales on rows whose codes agree.  Minima / maxima over the ranks that compared."""
    from tq100.pipeline import LinearView
    from tq100.gptq import GPTQ, HessianState
    dev = ctx.device
    owners = sharded_layer.owners([(a_, b_) for _, a_, b_, _ in lins]) if sharded_layer.mode == "linears" else [-1] * len(lins)
    rep = {"against": "single-GPU GPTQ.quantize on the same NCCL-reduced Hessian and weights", "linears": {}}
    for name in ("o_proj", "down_proj"):
        i = [l[0] for l in lins].index(name)
        _, n, m, src = lins[i]
        st = HessianState(m, dev)
        st.add_batch(acts[src])
        H = st.full()
        split = owners[i] < 0
        if split:
            dist.all_reduce(H)
        else:
            dist.reduce(H, dst=owners[i])
        cnt = torch.tensor([st.nsamples], dtype=torch.int64, device=dev)
        dist.all_reduce(cnt)
        st.nsamples = int(cnt.item())
        st._cache.clear()
        _, a, u, T8, perm, (lo, hi) = sharded_out[i]
        # code agreement, leading blocks, blocks, perm equal, alpha err, rows  (neutral element where this rank holds nothing)
        lo_stats = torch.tensor([1.0, 1e9, 1e9, 1.0], dtype=torch.float64, device=dev)       # reduced with MIN
        hi_stats = torch.tensor([0.0, 0.0], dtype=torch.float64, device=dev)                 # reduced with MAX / SUM
        if a is not None and hi > lo:
            g = GPTQ(LinearView(w_last[name]), 128, 0.01, hessian=st)
            g.quantize(use_ssr=order == "ssr", order=order)
            nb = (m + 127) // 128
            ps, pg = perm.long(), g.perm.long()
            lead = 0
            for k in range(nb):
                if not torch.equal(torch.sort(ps[k * 128:(k + 1) * 128]).values, torch.sort(pg[k * 128:(k + 1) * 128]).values):
                    break
                lead += 1
            cols = pg[:lead * 128]
            same = (T8[:, cols] == g.T_int8[lo:hi][:, cols]) if lead else torch.zeros((1, 1), dtype=torch.bool, device=dev)
            clean = same.all(dim=1) if lead else torch.zeros(1, dtype=torch.bool, device=dev)
            ga = g.alpha[lo:hi, :lead].float()
            aerr = ((a[:, :lead].float() - ga).abs() / ga.abs().clamp(min=1e-12))[clean] if lead else torch.zeros(0, device=dev)
            lo_stats = torch.tensor([float(same.float().mean()), float(lead), float(nb), 1.0 if torch.equal(ps, pg) else 0.0],
                                    dtype=torch.float64, device=dev)
            hi_stats = torch.tensor([float(aerr.max()) if aerr.numel() else 0.0, float(hi - lo)], dtype=torch.float64, device=dev)
        dist.all_reduce(lo_stats, op=dist.ReduceOp.MIN)
        mx = hi_stats.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = hi_stats.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        rep["linears"][name] = {"shape": [n, m], "split_by_rows": bool(split),
                                "min_code_agreement_on_leading_blocks": float(lo_stats[0]),
                                "leading_blocks_same_membership": int(lo_stats[1]), "blocks": int(lo_stats[2]),
                                "sweep_order_equal_on_every_rank": bool(lo_stats[3] == 1.0),
                                "max_alpha_rel_err_on_rows_with_equal_codes": float(mx[0]), "rows_compared": int(sm[1])}
    rep["ok"] = all(v["min_code_agreement_on_leading_blocks"] >= 0.999 and v["leading_blocks_same_membership"] >= 1
                    for v in rep["linears"].values())
    return rep


# ------------------------------------------------------------------------------------ configs[4]: single-layer sweep
def layer_sweep_line(args):
    """BASELINE configs[4]: single linears 4096x4096 ... 8192x28672, each stage timed alone (best of 3, CUDA events) and
    put against its roofline: Hessian vs the measured bf16 peak (useful SYRK flops), damped inverse (m^3 flops), the
    column sweep in sequential and SSR order vs the measured HBM copy peak on the algorithmic bytes of SURVEY 8d
    (read-modify-write of the remaining columns per block + the block's own reads and writes)."""
    import torch
    import tq100
    from tq100.pipeline import LinearView
    dev = torch.device("cuda:0")
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    tf_peak, hbm_peak = peaks.get("bf16_tflops", 1590.0), peaks.get("hbm_gbs", 6650.0)
    nt = SAMPLES * SEQ
    shapes = [(4096, 4096), (11008, 4096), (4096, 11008), (5120, 5120), (13824, 5120), (5120, 13824),
              (8192, 8192), (28672, 8192), (8192, 28672)]
    if args.layers:
        shapes = shapes[:args.layers]

    def best_of(fn, reps):
        fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        return min(ts)

    rows = []
    launches0 = tq100._lib.launch_count()
    for n, m in shapes:
        g = torch.Generator(device=dev).manual_seed(n + m)
        X = torch.empty((nt, m), device=dev, dtype=torch.float16)
        for lo in range(0, nt, 32768):
            X[lo:lo + 32768] = torch.randn((min(32768, nt - lo), m), device=dev, generator=g).to(torch.float16)
        W = torch.randn((n, m), device=dev, generator=g) * 0.02
        q = tq100.GPTQ(LinearView(W))
        st = q.state

        def hess():
            st.H.zero_()
            st.nsamples = 0
            st.add_batch(X)
        t_h = best_of(hess, max(1, args.steps))

        def inv():
            st._cache.clear()
            st.damped_inverse(0.01)
        t_inv = best_of(inv, max(1, args.steps))
        nb = (m + 127) // 128
        rmw = sum(8 * n * (m - 128 * (k + 1)) for k in range(nb - 1))              # feedback RMW of W[:, rem] per block
        blk = n * m * (4 + 1 + 8)                                                  # block read, codes, E (hi, lo)
        rec = {"n": n, "m": m, "tokens": nt, "hessian_ms": t_h,
               "hessian_tflops_useful": hessian_useful_flops(nt, m) / t_h / 1e9,
               "hessian_frac_of_bf16_burst_peak": hessian_useful_flops(nt, m) / t_h / 1e9 / tf_peak,
               "inverse_ms": t_inv, "inverse_tflops_fp32_equiv": m ** 3 / t_inv / 1e9}
        for key, kw, extra in (("sweep_seq", dict(use_ssr=False), 0), ("sweep_ssr", dict(use_ssr=True), 2 * 4 * n * m)):
            t = best_of(lambda: q.quantize(**kw), max(1, args.steps)) - t_inv * 0.0
            # quantize() reuses the cached inverse: the time is the sweep alone.  SSR adds the first block's two
            # statistics passes over W; later blocks take their statistics from the feedback epilogue (no extra bytes)
            gb = (rmw + blk + extra) / 1e9
            rec[key + "_ms"] = t
            rec[key + "_algorithmic_gb"] = gb
            rec[key + "_gbs"] = gb / t * 1e3
            rec[key + "_frac_of_hbm_peak"] = gb / t * 1e3 / hbm_peak
        rows.append(rec)
        print(json.dumps(rec), file=sys.stderr, flush=True)
        del X, W, q, st
        torch.cuda.empty_cache()
    line = {"metric": METRIC["layer-sweep"], "value": sum(r["hessian_ms"] + r["inverse_ms"] + r["sweep_ssr_ms"] for r in rows) / 1e3,
            "unit": "s", "n_gpus": 1, "steps": args.steps, "warmup": 1, "higher_is_better": False, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "single-layer sweep, one linear at a time: Hessian over 128x2048 fp16 tokens, damped inverse, "
                                   "column sweep (sequential and SSR), block 128", "baseline_config": "configs[4]",
                       "cache": "each stage's inputs exceed L2 except the 4096-wide inverses; no explicit flush"},
            "peaks": {"bf16_tflops_burst": tf_peak, "hbm_gbs": hbm_peak, "source": "MEASURED_PEAKS.json" if peaks else "fallback"},
            "gpu_launches": int(tq100._lib.launch_count() - launches0), "rows": rows,
            "value_note": "sum over the shapes of Hessian + inverse + SSR sweep, each timed alone"}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------ N1 leg: whole-model driver at 7B width
def whole_model_leg(torch, tq100, dev, layers=2, samples=SAMPLES, seq=SEQ):
    """SURVEY 8f N1 at full width: PT2LLMQuantizer.quantize() -- the reference's main.py:232-311 entry point -- on a
    llama-shaped stack of `layers` transformer layers at LLaMA-2-7B width (fp16 weights, random init), fed 128 x 2048
    synthetic calibration tokens.  Everything the per-linear bench leaves out is inside: one capture forward per sample,
    two layer forwards per layer and sample, Hessians streamed from forward hooks (one tq_hessian_accum per sample and
    distinct input), shared Hessians for q/k/v and gate/up, AGA on the activation Gram (main.py:177-180), the
    dequantised weights written back, every result copied to the host.  Activations never leave the GPU, so this is
    the compute-bound form of the end-to-end story (the `e2e` leg streams 12 GB of stored activations per layer over
    PCIe instead)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import toy_model
    torch.manual_seed(0)
    lm = toy_model.ToyLM(vocab=512, d=LLAMA2_7B["d"], ffn=LLAMA2_7B["ffn"], layers=layers).to(dev)
    with torch.no_grad():
        for name, p in lm.named_parameters():
            if p.dim() == 2:
                p.normal_(0.0, 0.02 if "embed" not in name else 1.0)
    lm = lm.half().eval()
    gen = torch.Generator().manual_seed(3)
    toks = [torch.randint(0, 512, (1, seq), generator=gen) for _ in range(samples)]
    w0 = {k: v.clone() for k, v in lm.state_dict().items()}
    dt = None
    for attempt in range(2):                    # the first pass warms cuBLAS / the allocator; the second is reported
        lm.load_state_dict(w0)
        pq = tq100.PT2LLMQuantizer(lm, None, model_type="llama", use_ssr=True, device=dev)
        before = tq100._lib.launch_count()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        params = pq.quantize(toks)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    return {"value": dt, "unit": "s", "layers": layers, "per_layer_s": dt / layers,
            "whole_model_estimate_s": dt / layers * LLAMA2_7B["layers"], "linears": len(params),
            "layer_forwards": int(pq.layer_forwards), "gpu_launches": int(tq100._lib.launch_count() - before),
            "note": f"PT2LLMQuantizer.quantize on {layers} llama-shaped layers at 7B width, {samples} x {seq} tokens, "
                    "fp16 model, SSR, hook-streamed Hessians (4 per layer), results on the host; wall clock; the estimate "
                    "multiplies the per-layer time by 32 (every layer has the same shapes); second of two passes.  About half of the "
                    "time is the model's own fp16 forwards (2 per layer and sample), which the per-linear legs do not contain"}


# ------------------------------------------------------------------------------------ N2 leg
def packed_layer_leg(torch, tq100, dev):
    n = m = 4096
    gen = torch.Generator(device=dev).manual_seed(7)
    T = torch.randint(-1, 2, (n, m), generator=gen, device=dev, dtype=torch.int8)
    alpha = 0.01 + 0.02 * torch.rand((n, m // 128), generator=gen, device=dev)
    mu = 0.004 * torch.randn((n, m // 128), generator=gen, device=dev)
    perm = torch.randperm(m, generator=gen, device=dev)
    code_bytes = n * (m // 16) * 4
    copies = int(2.2 * 126e6 / code_bytes) + 1
    layers = []
    for _ in range(copies):
        layer = tq100.TernaryLinear(m, n, bias=False, dtype=torch.float16, device=dev)
        layer.set_quantized_params(alpha, mu, T, perm)
        layer._prepared()             # derived buffers exist before any call is captured into a CUDA graph
        layers.append(layer)

    def timed(fn, items, iters):
        for it in items[:3]:
            fn(it)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            for i in range(iters):
                fn(items[i % len(items)])
        graph.replay()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        graph.replay()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / iters * 1e-3

    x1 = torch.randn((1, m), generator=gen, device=dev).half()
    t_dec = timed(lambda L: L(x1), layers, 2 * copies)
    dec_bytes = code_bytes + 16 * n * (m // 128) + 2 * m + 4 * n
    xs = torch.randn((512, m), generator=gen, device=dev).half()
    out = {"layer": "4096x4096 fp16, block 128, random permutation",
           "decode_1_token": {"us": t_dec * 1e6, "algorithmic_bytes": dec_bytes, "gbs": dec_bytes / t_dec / 1e9,
                              "kernel": "tl_gemv_kernel", "copies": copies}}
    for mode in ("fused", "dense"):
        for L in layers[:3]:
            L.fused_gemm = mode == "fused"
        t = timed(lambda L: L(xs), layers[:3], 6)
        out[f"tokens_512_{mode}"] = {"us": t * 1e6, "tflops": 2.0 * 512 * n * m / t / 1e12,
                                     "kernel": "tl_gemm_tc_kernel" if mode == "fused" else "tl_expand_kernel + library GEMM"}
    out["note"] = "SURVEY 8f N2 (TernaryLinear on 2-bit codes); secondary measurement, not part of `value`"
    return out


# ------------------------------------------------------------------------------------ GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="7b-ssr", choices=["7b-ssr", "13b-actorder", "layer-sweep"],
                    help="7b-ssr = BASELINE configs[1] (the driver's line); 13b-actorder = configs[3]; layer-sweep = configs[4]")
    ap.add_argument("--layers", type=int, default=0, help="(debug) fewer layers; the line says so")
    ap.add_argument("--streams", type=int, default=4, help="CUDA streams the per-linear prologue+sweep chains are spread over")
    ap.add_argument("--shard-mode", default="linears", choices=["linears", "rows"],
                    help="N > 1: deal whole linears to ranks (default) or row-shard every linear (SSR statistics all-reduced)")
    ap.add_argument("--split-threshold", type=float, default=1.5,
                    help="N > 1, --shard-mode linears: a linear whose chain exceeds this multiple of a rank's fair share is "
                         "split by rows (default 1.5: down_proj from 4 GPUs up)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-shared", action="store_true", help="skip the secondary shared-Hessian (N1) measurement")
    ap.add_argument("--no-packed", action="store_true", help="skip the secondary packed-layer (N2) measurement")
    ap.add_argument("--no-whole-model", action="store_true", help="skip the secondary whole-model (N1) measurement")
    ap.add_argument("--no-ref-cuda", action="store_true", help="skip the same-box reference-on-CUDA leg (and the parity it feeds)")
    ap.add_argument("--no-parity", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if args.config == "layer-sweep":
        if rank == 0:
            layer_sweep_line(args)
        return

    import torch
    import torch.distributed as dist
    import tq100
    from tq100 import _lib

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    from tq100 import sharded as par
    from tq100.pipeline import HostPipeline, LinearView
    ctx = par.ShardContext(rank, world, dev)

    cfg, order = bench_config(args)
    full_layers = (LLAMA2_13B if args.config == "13b-actorder" else LLAMA2_7B)["layers"]
    use_ssr = order == "ssr"
    lins = linears(cfg)
    nt = SAMPLES * SEQ

    # ---- synthetic inputs, resident in HBM before the timed region -----------------------------
    gen = torch.Generator(device=dev).manual_seed(1234)
    lam, r = 0.5, 64
    acts = {}
    my_samples = ctx.my_samples(SAMPLES)                       # calibration samples of this rank
    for ki, (key, width) in enumerate((("attn_in", cfg["d"]), ("o_in", cfg["d"]), ("mlp_in", cfg["d"]),
                                       ("down_in", cfg["ffn"]))):
        B = torch.randn((r, width), device=dev, generator=gen)
        x = torch.empty((len(my_samples), SEQ, width), device=dev, dtype=torch.float16)
        for j, s in enumerate(my_samples):
            g2 = torch.Generator(device=dev).manual_seed(100000 + 1000 * ki + s)
            z = torch.randn((SEQ, width), device=dev, generator=g2)
            f = torch.randn((SEQ, r), device=dev, generator=g2)
            x[j] = (z + (lam / r ** 0.5) * (f @ B)).to(torch.float16)
        acts[key] = x
    weights = []
    for li in range(cfg["layers"]):
        ws = {}
        for wi, (name, n, m, _) in enumerate(lins):
            gw = torch.Generator(device=dev).manual_seed(1000 * li + wi)          # identical on every rank
            ws[name] = torch.randn((n, m), device=dev, generator=gw) * 0.02
        weights.append(ws)
    torch.cuda.synchronize()

    hess_events = []
    results_keep = {}

    if world > 1:
        par.init_comm(ctx)
    sharded_layer = par.ShardedLayer(ctx, block_size=128, percdamp=0.01, mode=args.shard_mode, num_streams=args.streams,
                                     split_threshold=args.split_threshold)
    from tq100.pipeline import LayerDriver
    driver = LayerDriver(dev, block_size=128, percdamp=0.01, num_streams=args.streams)

    def quantize_model(record=False):
        for li in range(cfg["layers"]):
            if world > 1:
                # per layer: local Hessians -> NCCL all-reduce -> inverses dealt to ranks + broadcast -> row-slab sweeps
                ph = [] if (record and li == cfg["layers"] - 1) else None
                results_keep["last"] = sharded_layer.quantize(
                    [(name, weights[li][name], acts[src]) for name, n, m, src in lins], use_ssr=use_ssr, order=order,
                    hess_timing=hess_events if record else None, phases=ph)
                if ph is not None:
                    results_keep["phases"] = ph
                continue
            if os.environ.get("BENCH_DEBUG"):
                torch.cuda.synchronize()
                t_dbg = time.perf_counter()
            results_keep["last"] = driver.quantize(
                [(name, weights[li][name], acts[src]) for name, n, m, src in lins], use_ssr=use_ssr, order=order,
                hess_timing=hess_events if record else None)
            if os.environ.get("BENCH_DEBUG"):
                torch.cuda.synchronize()
                print(f"[debug] layer {li}: {1e3 * (time.perf_counter() - t_dbg):.1f} ms, allocated "
                      f"{torch.cuda.memory_allocated() / 1e9:.1f} GB, reserved {torch.cuda.memory_reserved() / 1e9:.1f} GB, "
                      f"mallocs {torch.cuda.memory_stats().get('num_device_alloc', -1)}, "
                      f"frees {torch.cuda.memory_stats().get('num_device_free', -1)}", file=sys.stderr, flush=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        quantize_model()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0 and not os.environ.get("BENCH_NO_SAMPLER"):
        sampler.start()
    launches0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        quantize_model(record=True)
    e1.record()
    barrier()
    launches = _lib.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_per_step = ms / args.steps
    value = ms_per_step / 1e3

    # N > 1: critical-path breakdown of the last timed layer on every rank (ms since the layer's start), gathered on rank 0
    breakdown = None
    if world > 1 and results_keep.get("phases"):
        ph = results_keep["phases"]
        mine = {tag: ph[0][1].elapsed_time(ev) for tag, ev in ph[1:]}
        tags = ["hessians", "reduce", "split_inverse", "split_bcast", "split_sweep", "own_chains"]
        row = torch.tensor([mine.get(tg, -1.0) for tg in tags], device=dev)
        rows = [torch.empty_like(row) for _ in range(world)]
        dist.all_gather(rows, row)
        breakdown = {"unit": "ms since the layer's first kernel, last timed layer, per rank", "tags": tags,
                     "ranks": [[round(float(x), 3) for x in r.tolist()] for r in rows],
                     "note": "hessians: local SYRKs done; reduce: NCCL reduce / all-reduce of every H done; split_inverse: the "
                             "owner's damped inverse of a row-split linear done (-1 on the other ranks); split_bcast: its H^-1 "
                             "received; split_sweep: this rank's row slab of it swept; own_chains: this rank's whole linears done"}

    # roofline of the dominant kernel (hessian_tc_kernel), live CUDA-event timing per launch
    flops = sum(hessian_useful_flops(t_, m_) for _, _, t_, m_ in hess_events)
    h_ms = sum(a.elapsed_time(b) for a, b, _, _ in hess_events)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = peaks.get("bf16_tflops_sustained", 1400.0)
    achieved = flops / (h_ms * 1e-3) / 1e12 if h_ms > 0 else 0.0
    # DRAM traffic per launch from the committed ncu --set full capture (profiles/r01c_ncu_full_hessian_tc.txt:
    # dram__bytes_read.sum + dram__bytes_write.sum), averaged over the launches of one layer like `achieved`
    NCU_TRAFFIC = {4096: 2.25e9, 11008: 46.4e9} if world == 1 else {}   # bytes per launch at Nt = 262144, by m
    known = [NCU_TRAFFIC[m_] for _, _, t_, m_ in hess_events if m_ in NCU_TRAFFIC and t_ == SAMPLES * SEQ]
    traffic = sum(known) / len(known) if known and len(known) == len(hess_events) else None
    roofline = {"kernel": "hessian_tc_pair_kernel", "bound": "tensor", "achieved": achieved, "peak": peak,
                "unit": "TFLOP/s", "frac": achieved / peak, "traffic": traffic,
                "traffic_note": "mean DRAM bytes per launch of the 2-SM kernel (ncu, dram__bytes_read.sum + dram__bytes_write.sum, "
                                "profiles/r02_hessian_traffic_ncu.txt: single-pass captures): 2.25 GB at m=4096 = 1.02x the algorithmic "
                                "2.21 GB (one L2-resident tile group, X read once), 46.4 GB at m=11008 (7.4x of 6.3 GB): H does not "
                                "fit L2, X is re-read once per 3584-wide tile group and the reduce-adds spill (13 GB written).  The "
                                "multi-pass --set full capture of the same kernel (profiles/r02_ncu_full_hessian_tc.txt) reads 3.37 / "
                                "58.1 GB.  The kernel is tensor-bound (tensor pipe 78 %, DRAM at 12-27 % of peak)" if traffic else None,
                "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step)"
                               if peaks else "fallback 1.4 PFLOP/s sustained",
                "flops_counted": "useful SYRK flops Nt*m*(m+1) per launch (dense-equivalent 2*Nt*m^2 is 2x)",
                "share_of_step": h_ms / ms if ms > 0 else None, "launches": len(hess_events)}

    # ---- SURVEY 8(f) N1, reported beside the headline: linears that read the same activations (q,k,v; gate,up) share one
    # Hessian and one inverse -- 4 instead of 7 per layer, outputs unchanged.  The headline `value` above keeps the
    # reference's 7 (main.py:289-299 computes one per linear).
    n1 = None
    if world == 1 and not args.no_shared:
        shared_driver = LayerDriver(dev, block_size=128, percdamp=0.01, num_streams=args.streams, share_inputs=True)

        def quantize_model_shared():
            for li in range(cfg["layers"]):
                results_keep["last"] = shared_driver.quantize(
                    [(name, weights[li][name], acts[src]) for name, n, m, src in lins], use_ssr=use_ssr, order=order)

        quantize_model_shared()
        torch.cuda.synchronize()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for _ in range(args.steps):
            quantize_model_shared()
        s1.record()
        torch.cuda.synchronize()
        n1_ms = s0.elapsed_time(s1) / args.steps
        n1 = {"value": n1_ms / 1e3, "unit": "s", "ms_per_step": n1_ms, "hessians_per_layer": 4,
              "note": "same workload with LayerDriver(share_inputs=True): one Hessian + inverse per distinct calibration "
                      "input (SURVEY 8f N1); identical outputs; not the headline"}

    # ---- e2e: same API from pinned host buffers, copies inside the timed region ------------------
    e2e = None
    if not args.no_e2e and world == 1:
        # one layer's weights and the four calibration inputs in pinned host memory, streamed per layer
        host_acts = {k: v.reshape(-1, v.shape[-1]).cpu().pin_memory() for k, v in acts.items()}
        host_w = {name: weights[0][name].cpu().pin_memory() for name, _, _, _ in lins}
        groups_one_layer = [(host_acts[k], [(name, host_w[name]) for name, _, _, src in lins if src == k])
                            for k in ("attn_in", "o_in", "mlp_in", "down_in")]
        pipe = HostPipeline(dev, block_size=128, percdamp=0.01, use_ssr=use_ssr, aga="hessian", order=order)
        pipe.run(groups_one_layer)                                 # warm-up (allocations, pinned outputs)
        torch.cuda.synchronize()
        pipe.h2d_bytes = pipe.d2h_bytes = 0
        t0 = time.perf_counter()
        # the whole model streams through one call: the copies of the next input overlap this one's kernels, across
        # layer boundaries too; results are consumed (here: dropped) per group, as a checkpoint writer would
        res = None
        for res in pipe.run_iter(groups_one_layer * cfg["layers"]):
            pass
        pipe.synchronize()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        e2e = {"value": dt, "unit": "s", "h2d_bytes_per_step": int(pipe.h2d_bytes),
               "d2h_bytes_per_step": int(pipe.d2h_bytes),
               "note": "HostPipeline.run per layer: each of the layer's 4 distinct calibration inputs and 7 weights "
                       "copied from pinned host memory on a side stream (double-buffered: the next input's copy is "
                       "enqueued before this input's kernels), alpha/mu/T(int8)/perm copied back to pinned host memory; "
                       "wall clock around all layers"}
        del host_acts, host_w, pipe, res
    elif not args.no_e2e:
        # N > 1: every rank streams ITS calibration samples and ITS row slab of the weights from pinned host memory,
        # runs the sharded layer (all-reduce of H, dealt inverses, row-slab sweeps), and copies its slabs of
        # alpha / mu / T (int8) and perm back to pinned host memory.
        host_acts = {k: v.cpu().pin_memory() for k, v in acts.items()}                    # already this rank's samples
        if args.shard_mode == "rows":
            slabs = {name: ctx.row_range(n) for name, n, m, _ in lins}
        else:
            own = sharded_layer.owners([(n, m) for _, n, m, _ in lins])
            slabs = {name: (ctx.row_range(n) if own[i] < 0 else (0, n) if own[i] == rank else None)
                     for i, (name, n, m, _) in enumerate(lins)}
        host_w = {name: (weights[0][name][slabs[name][0]:slabs[name][1]].cpu().pin_memory() if slabs[name] else None)
                  for name, _, _, _ in lins}
        one_layer = (host_acts, [(name, host_w[name], n, src) for name, n, m, src in lins])
        spipe = par.ShardedHostPipeline(ctx, block_size=128, percdamp=0.01, use_ssr=use_ssr, aga="hessian",
                                        mode=args.shard_mode, order=order)
        for keep in spipe.run_iter([one_layer]):                   # warm-up (device slots, pinned outputs)
            pass
        spipe.synchronize()
        barrier()
        spipe.h2d_bytes = spipe.d2h_bytes = 0
        t0 = time.perf_counter()
        for keep in spipe.run_iter([one_layer] * cfg["layers"]):
            pass
        spipe.synchronize()
        barrier()
        dt = time.perf_counter() - t0
        e2e = {"value": dt, "unit": "s", "h2d_bytes_per_step": int(spipe.h2d_bytes), "d2h_bytes_per_step": int(spipe.d2h_bytes),
               "note": "ShardedHostPipeline per rank: its calibration samples and its part of the weights (--shard-mode) copied from "
                       "pinned host memory (next layer's copies enqueued before this layer's kernels), "
                       "ShardedLayer.quantize, its alpha/mu/T(int8)/perm copied back; bytes are those of rank 0; "
                       "wall clock around all layers with barriers, max over ranks"}
        del host_acts, host_w, spipe, keep
    if e2e is not None and world > 1:
        t = torch.tensor([e2e["value"]], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e["value"] = float(t.item())
    # ---- same-box bar: the reference's own code path with device='cuda' (its default
    # device, main.py:368), fp32, TF32 off, on ONE transformer layer with all 262144 tokens -- cuBLAS sgemm, cuSOLVER
    # potrf/potri and the reference's per-iteration host syncs -- times the layer count.  Its outputs on layer 0 are also
    # the yardstick of `parity`: the timed run's own layer-0 results against them, linear by linear.
    ref_cuda = parity = None
    if world == 1 and not args.no_ref_cuda:
        try:
            ref_cuda, parity = reference_cuda_leg(torch, args, cfg, order, lins, weights[0], acts, driver, dev)
        except Exception as exc:
            ref_cuda = {"error": f"{type(exc).__name__}: {exc}"[:300]}
    elif world > 1 and not args.no_parity:
        try:
            parity = sharded_parity_leg(torch, dist, tq100, args, ctx, order, lins, weights[cfg["layers"] - 1], acts,
                                        results_keep["last"], sharded_layer)
        except Exception as exc:
            parity = {"error": f"{type(exc).__name__}: {exc}"[:300]}

    # ---- SURVEY 8(f) N2, reported beside the headline: the packed inference layer on the codes (TernaryLinear):
    # one 4096x4096 fp16 layer with a random permutation; decode call (1 token, tq_tl_gemv) over layer copies that
    # total > 2x L2, and a 512-token call through tq_tl_gemm_tc vs dense weight + library GEMM.  Device-timed replay
    # of a CUDA graph of the calls; never part of `value`;
    # last, so that nothing it does can disturb the measurements above.
    n2 = None
    if world == 1 and not args.no_packed:
        try:
            n2 = packed_layer_leg(torch, tq100, dev)
        except Exception as exc:                                   # secondary measurement: never lose the headline line
            n2 = {"error": f"{type(exc).__name__}: {exc}"[:300]}

    n1_model = None
    if world == 1 and not args.no_whole_model and args.config == "7b-ssr":
        try:
            n1_model = whole_model_leg(torch, tq100, dev)
        except Exception as exc:
            n1_model = {"error": f"{type(exc).__name__}: {exc}"[:300]}

    if rank == 0:
        line = {"metric": METRIC[args.config], "value": value, "unit": "s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
                "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": workload_config(args), "clocks": clocks,
                "gpu_launches": int(launches), "roofline": roofline, "e2e": e2e}
        if breakdown is not None:
            line["critical_path"] = breakdown
        if cfg["layers"] != full_layers:
            line["config"]["workload"] += f" [DEBUG: only {cfg['layers']} of {full_layers} layers]"
        if n1 is not None:
            line["n1_shared_inputs"] = n1
        if n2 is not None:
            line["n2_packed_layer"] = n2
        if n1_model is not None:
            line["n1_whole_model"] = n1_model
        line["dtype_note"] = "fp32 arithmetic; fp16 activations enter the tensor cores exactly, fp32 accumulate"
        if ref_cuda is not None:
            line["reference_cuda"] = ref_cuda
        if parity is not None:
            line["parity"] = parity
        if not args.no_cpu_baseline and world == 1:
            # rank 0 at N = 1 only (at N > 1 the other ranks would spin in the final barrier on the same host cores);
            # one pass over the layer's 7 linears -- the `--impl reference` arm is the repeated measurement
            free = torch.cuda.mem_get_info(dev)[0]
            del acts, weights
            torch.cuda.empty_cache()
            line["cpu_baseline"] = cpu_layer_sample(cfg, order, steps=1)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
