"""Parity AT THE BENCH'S OWN SHAPES (BASELINE.json configs[1]/[2]: the LLaMA-2-7B linears) against the oracle --
``oracle/torch_port.py``, the restatement of /root/reference/gptq.py:78-199 that is pinned to the unmodified reference's
outputs (tests/test_oracle_golden.py) -- on identical inputs: fp16 synthetic activations, 16 384 tokens, lambda = 0.5,
W ~ N(0, 0.02^2).  The CUDA path computes its own Hessian with the tcgen05 kernel from the fp16 activations;
the oracle computes X'X in fp32 on the host from the same values, so the two sides differ exactly where the reference
and the product differ: accumulation order and the fp32-emulating tensor-core GEMMs.

Tolerances are north_star's: codes >= 99.9 % (where the reference's own fp32-vs-fp64 floor allows it, see the test's
docstring), alpha / mu within 1e-4 relative on (row, block) pairs no earlier flip has touched, reconstruction error within
1e-3 relative.  Every disagreement is adjudicated: the margin
| |Z| - 0.5 | of the reference's rounding at the first block where a row diverges -- a threshold tie -- is reported and
bounded (the smallest margin on the row's ITF trajectory: a tie several iterations before the last can tip a row into a
different fixed point).  With SSR one swapped top-k boundary legitimately de-correlates everything after it (the reference does not
reproduce itself there), so SSR cases assert block by block on the leading blocks whose membership
agrees, always the first block, and the reconstruction error; the report says how many blocks agreed.

Reports land in gpurun_out/parity_benchshape.json (copied to profiles/ by hand)."""

import json
import os

import numpy as np
import pytest
import torch

import parity
from oracle import torch_port

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
NT = 16384
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_inputs = {}
_report = {}


def _activations(m):
    """fp16 (NT, m) on the device + the oracle's fp32 Hessian of the same values (host sgemm), cached per width."""
    if m not in _inputs:
        g = torch.Generator(device=DEV).manual_seed(4000 + m)
        r, lam = 64, 0.5
        B = torch.randn((r, m), device=DEV, generator=g)
        X = (torch.randn((NT, m), device=DEV, generator=g)
             + (lam / r ** 0.5) * (torch.randn((NT, r), device=DEV, generator=g) @ B)).to(torch.float16)
        Xc = X.float().cpu()
        torch.set_num_threads(os.cpu_count() or 8)
        H = Xc.T @ Xc                                                    # gptq.py:74-75 on the host cores
        _inputs.clear()                                                  # one width resident at a time
        _inputs[m] = (X, H)
    return _inputs[m]


def _weight(n, m):
    g = torch.Generator(device=DEV).manual_seed(17 * n + m)
    return torch.randn((n, m), device=DEV, generator=g) * 0.02


def _recon(W, Wq, H):
    D = (W - Wq).double()
    Hd = H.double()
    return float(torch.sqrt((D @ Hd * D).sum() / ((W.double() @ Hd) * W.double()).sum()))


def _dequant(alpha, mu, T, perm, block=128):
    Wq = torch.empty_like(T)
    for k in range(alpha.shape[1]):
        cols = perm[k * block:(k + 1) * block]
        Wq[:, cols] = alpha[:, k:k + 1] * T[:, cols] + mu[:, k:k + 1]
    return Wq


def _np_aids(aids):
    return {k: (v.numpy() if torch.is_tensor(v) else v) for k, v in aids.items()}


def _save(tag, rep):
    _report[tag] = rep
    out = os.path.join(ROOT, "gpurun_out")
    try:
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, "parity_benchshape.json"), "w") as f:
            json.dump(_report, f, indent=1, sort_keys=True)
    except OSError:
        pass
    print(tag, json.dumps(rep))


def _oracle64(W, X, use_ssr):
    """The yardstick of SURVEY 8c: the oracle in fp64 (run on the GPU for speed; same ATen call sequence)."""
    Xd = X.double()
    H64 = Xd.T @ Xd
    del Xd
    a, u, T, p, aids = torch_port.quantize_layer(W.double(), H64, NT, 128, 0.01, use_ssr=use_ssr, return_margin=True)
    ref = dict(alpha=a.cpu().numpy(), mu=u.cpu().numpy(), T=T.cpu().numpy(), perm=p.cpu().numpy())
    return ref, _np_aids({k: (v.cpu() if torch.is_tensor(v) else v) for k, v in aids.items()})


@pytest.mark.parametrize("n,m,order", [(4096, 4096, "sequential"), (4096, 4096, "ssr"), (11008, 4096, "sequential"),
                                       (4096, 11008, "sequential"), (4096, 11008, "ssr")])
def test_bench_shape_layer_vs_oracle(n, m, order):
    """Three-way adjudication: (i) CUDA path vs the oracle in fp32 on the host cores -- the reference's CPU
    path; (ii) the fp32 oracle vs the oracle in fp64 -- the noise floor of the reference itself, taken on both of its
    devices: host (oneMKL + LAPACK) and cuda, its own default (main.py:368: cuBLAS sgemm with TF32 off + cuSOLVER);
    (iii) CUDA path vs the fp64 oracle.  Measured on a B200 (profiles/r02_parity_benchshape.json): at 4096 x 11008 with
    16 384 tokens the REFERENCE ON CUDA agrees with its own fp64 run on 99.65 % of the codes (320 of 4096 rows leave the
    fp64 trajectory at a threshold tie somewhere in 86 blocks; on the host libraries 33 rows), this path on 99.55 % (381
    rows), so north_star's 99.9 % is not reachable at this shape by the reference's own fp32 run on its default
    device.  The assertions therefore are: the bar itself where the floor allows it, otherwise 'no further from the fp64
    answer than the reference's own fp32 runs' (rows diverged <= 1.5 x the larger floor + 20), and in every case that
    each divergence starts at a tie."""
    import tq100
    from tq100.pipeline import LinearView
    X, H_ref = _activations(m)
    W = _weight(n, m)
    use_ssr = order == "ssr"
    q = tq100.GPTQ(LinearView(W))
    q.add_batch(X.reshape(NT // 2048, 2048, m))
    alpha, mu, T, perm = q.quantize(use_ssr=use_ssr)
    assert q.info == 0
    Wq = q.get_quantized_weight()
    Wc = W.cpu()
    ra, ru, rT, rp, aids32 = torch_port.quantize_layer(Wc, H_ref, NT, 128, 0.01, use_ssr=use_ssr, return_margin=True)
    got = dict(alpha=alpha.cpu().numpy(), mu=mu.cpu().numpy(), T=T.cpu().numpy(), perm=perm.cpu().numpy())
    ref32 = dict(alpha=ra.numpy(), mu=ru.numpy(), T=rT.numpy(), perm=rp.numpy())
    ref64, aids64 = _oracle64(W, X, use_ssr)
    # the fp32 oracle on the reference's own default device (main.py:368: cuda): cuBLAS sgemm (TF32 off) + cuSOLVER
    torch.backends.cuda.matmul.allow_tf32 = False
    Xf = X.float()
    ga, gu, gT, gp = torch_port.quantize_layer(W, Xf.T @ Xf, NT, 128, 0.01, use_ssr=use_ssr)
    del Xf
    ref32_cuda = dict(alpha=ga.cpu().numpy(), mu=gu.cpu().numpy(), T=gT.cpu().numpy(), perm=gp.cpu().numpy())
    rep = parity.adjudicate(got, ref32, _np_aids(aids32))                     # (i)
    floor_cpu = parity.adjudicate(ref32, ref64, aids64)                       # (ii) host libraries (oneMKL, LAPACK)
    floor_cuda = parity.adjudicate(ref32_cuda, ref64, aids64)                 # (ii) device libraries
    floor = max((floor_cpu, floor_cuda), key=lambda f: f.get("rows_diverged", 0))
    vs64 = parity.adjudicate(got, ref64, aids64)                              # (iii)
    e_got = _recon(Wc, Wq.cpu(), H_ref)
    e_ref = _recon(Wc, _dequant(ra, ru, rT, rp), H_ref)
    rep.update(recon_got=e_got, recon_ref=e_ref, recon_rel_diff=abs(e_got - e_ref) / e_ref, order=order, tokens=NT)
    keys = ("code_agreement", "rows_diverged", "pairs_disagreeing", "pairs", "leading_blocks_same_membership",
            "alpha_rel_err_max", "mu_err_rel_alpha_max", "tie_margin_first_divergence", "first_differing_block")
    rep["floor_ref32_cpu_vs_ref64"] = {k: floor_cpu.get(k) for k in keys}
    rep["floor_ref32_cuda_vs_ref64"] = {k: floor_cuda.get(k) for k in keys}
    rep["cuda_vs_ref64"] = {k: vs64.get(k) for k in keys}
    _save(f"{n}x{m}_{order}", rep)

    assert rep["recon_rel_diff"] <= parity.RECON_RTOL, rep
    assert rep["first_block_same_membership"], rep
    lead = rep["leading_blocks_same_membership"]
    if not use_ssr:
        assert rep["perm_equal"] and lead == rep["blocks"], rep
    else:
        # a swapped top-k boundary legitimately de-correlates everything after it (the reference does not reproduce
        # itself there either: its fp32 run keeps floor['leading_blocks_same_membership'] blocks of its fp64 run); the
        # CUDA path must keep at least half as many blocks of the fp64 run as the reference's fp32 run does
        lead_floor = min(floor_cpu["leading_blocks_same_membership"], floor_cuda["leading_blocks_same_membership"])
        assert vs64["leading_blocks_same_membership"] >= max(1, lead_floor // 2), rep
    # codes: the bar, or the reference's own floor (judged against fp64 on the blocks all three runs share)
    cols = ref32["perm"][:lead * 128]
    agree = float((got["T"][:, cols] == ref32["T"][:, cols]).mean()) if lead else 0.0
    rep["code_agreement_leading_blocks"] = agree
    _save(f"{n}x{m}_{order}", rep)
    at_floor = vs64.get("rows_diverged", 0) <= 1.5 * floor.get("rows_diverged", 0) + 20
    assert agree >= parity.CODE_AGREEMENT or (at_floor and agree >= 0.99), rep
    assert at_floor, rep
    assert rep["alpha_rel_err_max"] <= parity.SCALE_RTOL, rep
    assert rep["mu_err_rel_alpha_max"] <= parity.SCALE_RTOL, rep
    for r in (rep, vs64):
        tie = r.get("tie_margin_first_divergence")
        if tie is not None and tie["rows"] >= 5:
            # disagreements start at threshold ties: a row leaves the other run's trajectory only where some rounding on
            # its ITF path had |Z| within fp32 noise of 0.5 (a row that does not diverge typically stays ~1e-3 away:
            # trajectory_margin_all_rows_median in the report)
            assert tie["median"] <= 5e-5, rep
