#!/usr/bin/env python
"""One 4096 x 4096 fp16 TernaryLinear (SSR-like permutation): decode calls with 1 and 4 tokens, then a 512-token call on
the dense path and (TL_FUSED=1) on tq_tl_gemm_tc.  The ncu target for the packed-layer kernels (scripts/gpu_trip_tl.sh)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tq100  # noqa: E402

dev = torch.device("cuda:0")
gen = torch.Generator(device=dev).manual_seed(1)
n = m = 4096
T = torch.randint(-1, 2, (n, m), generator=gen, device=dev, dtype=torch.int8)
alpha = 0.01 + 0.02 * torch.rand((n, 32), generator=gen, device=dev)
mu = 0.004 * torch.randn((n, 32), generator=gen, device=dev)
perm = torch.randperm(m, generator=gen, device=dev)
layer = tq100.TernaryLinear(m, n, bias=False, dtype=torch.float16, device=dev)
layer.set_quantized_params(alpha, mu, T, perm)
for rep in range(2):
    for M in (1, 4):
        layer(torch.randn((M, m), generator=gen, device=dev).half())
x = torch.randn((512, m), generator=gen, device=dev).half()
layer.fused_gemm = False
layer(x)
if os.environ.get("TL_FUSED") == "1":
    layer.fused_gemm = True
    layer(x)
    layer(x)
torch.cuda.synchronize()
print("ok", tq100._lib.launch_count(), "launches")
