import os, sys, ctypes, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import gpu_util as G
from tq100 import _lib
lib = _lib.load()
m = 1024
rng = np.random.default_rng(0)
X = rng.standard_normal((2048, m)).astype(np.float32)
H = G.dev(X.T @ X)
buf = torch.zeros(64, dtype=torch.int64, device="cuda:0")
lib.tq_debug_set_diag_prof.argtypes = [ctypes.c_void_p]
lib.tq_debug_set_diag_prof(ctypes.c_void_p(buf.data_ptr()))
for _ in range(2):
    G.finalize_and_invert(H, 2048)
t = buf.cpu().numpy()
t = t[t > 0]
names = ["load"] + sum([[f"factor+invert{d}", f"rows_below+Vrow{d}", f"update{d}"] for d in range(4)], []) + ["store"]
d = np.diff(t)
for nme, c in zip(names, d):
    print(f"{nme:18s} {c:8d} cycles")
print("total", t[-1] - t[0], "cycles")
