import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tq100
from tq100.pipeline import LayerDriver
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
shapes = [("a", 384, 640), ("b", 640, 384), ("c", 300, 520)]
NT = 2048
Xs = {m: torch.randn((NT, m), device=dev, dtype=torch.float16, generator=g) for m in (640, 384, 520)}
Ws = {n_: torch.randn((n, m), device=dev, generator=g) * 0.02 for n_, n, m in shapes}
drv = LayerDriver(dev, num_streams=int(os.environ.get("STREAMS", 1)))
keep = None
for rep in range(2):
    for ssr in (True, False):
        gs = drv.quantize([(n_, Ws[n_], Xs[m]) for n_, n, m in shapes], use_ssr=ssr)
        torch.cuda.synchronize()
        print("rep", rep, "ssr", ssr, "info", [x.info for x in gs], flush=True)
        keep = gs
print("done")
