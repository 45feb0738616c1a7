"""One linear through the public API (for ncu launch lists): N, M, NT, SSR env vars."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tq100
from tq100.pipeline import LinearView
dev = torch.device("cuda:0")
n, m, nt = int(os.environ.get("N", 4096)), int(os.environ.get("M", 4096)), int(os.environ.get("NT", 32768))
g = torch.Generator(device=dev).manual_seed(0)
X = torch.randn((nt, m), device=dev, dtype=torch.float16, generator=g)
W = torch.randn((n, m), device=dev, generator=g) * 0.02
for rep in range(int(os.environ.get("REPS", 2))):
    q = tq100.GPTQ(LinearView(W))
    q.add_batch(X)
    q.quantize(use_ssr=bool(int(os.environ.get("SSR", 1))))
torch.cuda.synchronize()
print("done")
