"""Hessian throughput and trace error vs the token-chunk cap (env TQ_HESS_KC_MAX)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tq100
from tq100 import _lib
DEV = torch.device("cuda:0")
lib = _lib.load()
nt = 262144
for m in (4096, 11008):
    g = torch.Generator(device=DEV).manual_seed(m)
    X = torch.randn((nt, m), device=DEV, dtype=torch.float16, generator=g)
    H = torch.zeros((m, m), device=DEV)
    def run():
        _lib.check(lib.tq_hessian_accum(_lib.ptr(H), m, _lib.ptr(X), nt, m, m, _lib.F16, _lib.HESS_TCGEN05, _lib.stream()), "h")
    run(); torch.cuda.synchronize()
    ts = []
    for _ in range(3):
        H.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    tr = torch.diagonal(H).double().sum().item()
    ref = 0.0
    for lo in range(0, nt, 32768):
        ref += (X[lo:lo + 32768].double() ** 2).sum().item()
    # off-diagonal check on a 256-column slab against fp64
    cols = slice(128, 384)
    Href = X[:, :128].double().T @ X[:, cols].double()
    off = ((H[:128, cols].double() - Href).abs().max() / Href.abs().max()).item()
    best = min(ts)
    print(f"KC_MAX={os.environ.get('TQ_HESS_KC_MAX', '2048')} m={m}: {best:.2f} ms, {nt * m * (m + 1) / best / 1e9:.0f} TFLOP/s useful, "
          f"trace rel err {abs(tr - ref) / ref:.2e}, off-diagonal block max err / max {off:.2e}", flush=True)
    del X, H
