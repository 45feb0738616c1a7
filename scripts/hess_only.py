"""One tcgen05 Hessian launch per width (ncu target): M env = comma-separated widths, NT tokens."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tq100
from tq100 import _lib
DEV = torch.device("cuda:0")
lib = _lib.load()
nt = int(os.environ.get("NT", 262144))
for m in [int(x) for x in os.environ.get("M", "4096,11008").split(",")]:
    X = torch.randn((nt, m), device=DEV, dtype=torch.float16)
    H = torch.zeros((m, m), device=DEV)
    _lib.check(lib.tq_hessian_accum(_lib.ptr(H), m, _lib.ptr(X), nt, m, m, _lib.F16, _lib.HESS_TCGEN05, _lib.stream()), "h")
    torch.cuda.synchronize()
    del X, H
print("done")
