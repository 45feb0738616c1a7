import os, sys, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tq100
from tq100.pipeline import LayerDriver
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
shapes = [("q", 4096, 4096), ("k", 4096, 4096), ("v", 4096, 4096), ("o", 4096, 4096), ("g", 11008, 4096), ("u", 11008, 4096), ("d", 4096, 11008)]
NT = int(os.environ.get('NT', 32768))
Xs = {4096: torch.randn((NT, 4096), device=dev, dtype=torch.float16, generator=g), 11008: torch.randn((NT, 11008), device=dev, dtype=torch.float16, generator=g)}
Ws = {n_: torch.randn((n, m), device=dev, generator=g) * 0.02 for n_, n, m in shapes}
for ns in (int(os.environ.get("STREAMS", 2)),):
    drv = LayerDriver(dev, num_streams=ns)
    for rep in range(5):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        os.environ['TQ_DRIVER_DEBUG'] = ''
        gs = drv.quantize([(n_, Ws[n_], Xs[m]) for n_, n, m in shapes], use_ssr=True, hess_timing=([] if os.environ.get('EVT') else None))
        torch.cuda.synchronize(); t1 = time.perf_counter()
        print(f"streams={ns} rep={rep}: layer {1e3*(t1-t0):.1f} ms, info={[x.info for x in gs]}, mem={torch.cuda.memory_allocated()/1e9:.1f} GB reserved={torch.cuda.memory_reserved()/1e9:.1f} GB", flush=True)
        if os.environ.get('KEEP'):
            keep = gs
        del gs
