import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tq100
from tq100 import _lib
lib = _lib.load(); DEV = torch.device("cuda:0")
n = m = 4096
g = torch.Generator(device=DEV).manual_seed(0)
W = torch.randn((n, m), device=DEV, generator=g) * 0.02
E = torch.randn((n, 128), device=DEV, generator=g) * 0.01
A = torch.randn((m, m // 2), device=DEV, generator=g)
Hinv = (A @ A.T / m + torch.eye(m, device=DEV)).contiguous()
rem = m - 128
ws = torch.empty(lib.tq_err_feedback_tc_workspace_floats(n, 128, rem), device=DEV)
def run():
    _lib.check(lib.tq_err_feedback_tc(_lib.ptr(W), m, n, _lib.ptr(E), 128, _lib.ptr(Hinv), m, None, 0, 128, None, 128, rem, _lib.ptr(ws), _lib.stream()), "fb")
for _ in range(3): run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): run()
e1.record(); torch.cuda.synchronize()
print("TQ_GX_DEBUG=%s: %.1f us per call" % (os.environ.get("TQ_GX_DEBUG", "0"), e0.elapsed_time(e1) / 20 * 1e3))
