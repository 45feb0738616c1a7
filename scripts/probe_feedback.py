#!/usr/bin/env python
"""Tiny driver for ncu: a few error-feedback steps (tensor-core and CUDA-core) on one 4096 x 4096 layer."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tq100
from tq100 import _lib
lib = _lib.load()
DEV = torch.device("cuda:0")
n = int(os.environ.get("N", 4096)); m = int(os.environ.get("M", 4096))
gather = int(os.environ.get("GATHER", 0))
g = torch.Generator(device=DEV).manual_seed(0)
W = torch.randn((n, m), device=DEV, generator=g) * 0.02
E = torch.randn((n, 128), device=DEV, generator=g) * 0.01
A = torch.randn((m, m // 2), device=DEV, generator=g)
Hinv = (A @ A.T / m + torch.eye(m, device=DEV)).contiguous()
rem = m - 128
ws = torch.empty(lib.tq_err_feedback_tc_workspace_floats(n, 128, rem), device=DEV)
perm = torch.randperm(m, device=DEV, generator=g)
blk = perm[:128].to(torch.int32).contiguous()
remi = torch.sort(perm[128:]).values.to(torch.int32).contiguous()
for it in range(3):
    if gather:
        _lib.check(lib.tq_err_feedback_tc(_lib.ptr(W), m, n, _lib.ptr(E), 128, _lib.ptr(Hinv), m, _lib.ptr(blk), 0, 128,
                                          _lib.ptr(remi), 0, rem, _lib.ptr(ws), _lib.stream()), "fb")
    else:
        _lib.check(lib.tq_err_feedback_tc(_lib.ptr(W), m, n, _lib.ptr(E), 128, _lib.ptr(Hinv), m, None, 0, 128, None, 128,
                                          rem, _lib.ptr(ws), _lib.stream()), "fb")
torch.cuda.synchronize()
print("done")
