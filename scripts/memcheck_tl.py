"""compute-sanitizer target for the packed-layer kernels (ragged shapes, every dtype / tile width):
    compute-sanitizer --tool memcheck python scripts/memcheck_tl.py
    compute-sanitizer --tool racecheck python scripts/memcheck_tl.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tq100  # noqa: E402

dev = torch.device("cuda:0")
gen = torch.Generator(device=dev).manual_seed(0)
for n, m, block in ((70, 1000, 64), (200, 328, 128), (33, 4224, 128)):
    nb = (m + block - 1) // block
    T = torch.randint(-1, 2, (n, m), generator=gen, device=dev, dtype=torch.int8)
    alpha = 0.01 + 0.02 * torch.rand((n, nb), generator=gen, device=dev)
    mu = 0.004 * torch.randn((n, nb), generator=gen, device=dev)
    perm = torch.randperm(m, generator=gen, device=dev)
    for dtype in (torch.float16, torch.bfloat16, torch.float32):
        layer = tq100.TernaryLinear(m, n, block_size=block, bias=True, dtype=dtype, device=dev)
        layer.set_quantized_params(alpha, mu, T, perm, torch.zeros(n, device=dev))
        assert torch.equal(layer.T, T)
        for tokens in (1, 3, 5, 16, 40, 300):
            x = torch.randn((tokens, m), generator=gen, device=dev).to(dtype)
            for width in ("128", "256", "512"):
                os.environ["TQ_TL_GEMM_BN"] = width
                y = layer(x)
                if tokens <= 16:
                    break
        layer.fused_gemm = False
        layer(torch.randn((40, m), generator=gen, device=dev).to(dtype))
    torch.cuda.synchronize()
    print("ok", n, m, block, flush=True)
print("done")
