"""Thin helpers for the -m gpu tests: raw C-ABI calls on torch CUDA tensors."""

import numpy as np
import torch

import tq100
from tq100 import _lib

DEV = "cuda:0"


def lib():
    return _lib.load()


def dev(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a)).to(DEV)
    return t if dtype is None else t.to(dtype)


def i32(a):
    return torch.from_numpy(np.ascontiguousarray(np.asarray(a, dtype=np.int32))).to(DEV)


def hessian(X, m, path, H=None):
    """X: torch CUDA tensor (Nt, m) f32/f16/bf16 -> H (m, m) fp32 full symmetric."""
    L = lib()
    if H is None:
        H = torch.zeros((m, m), dtype=torch.float32, device=DEV)
    _lib.check(L.tq_hessian_accum(_lib.ptr(H), m, _lib.ptr(X), X.shape[0], m, X.stride(0),
                                  _lib.dtype_code(X.dtype), path, _lib.stream()), "tq_hessian_accum")
    _lib.check(L.tq_symmetrize(_lib.ptr(H), m, m, _lib.stream()), "tq_symmetrize")
    return H


def finalize_and_invert(Hraw, nsamples, percdamp=0.01):
    L = lib()
    m = Hraw.shape[0]
    Hd = torch.empty_like(Hraw)
    Hinv = torch.empty_like(Hraw)
    work = torch.empty(L.tq_chol_workspace_floats(m), dtype=torch.float32, device=DEV)
    scratch = torch.empty(8, dtype=torch.float32, device=DEV)
    info = torch.zeros(1, dtype=torch.int32, device=DEV)
    _lib.check(L.tq_hessian_finalize(_lib.ptr(Hd), _lib.ptr(Hraw), m, float(nsamples), float(percdamp),
                                     _lib.ptr(scratch), _lib.stream()), "tq_hessian_finalize")
    _lib.check(L.tq_chol_inverse(_lib.ptr(Hinv), _lib.ptr(Hd), m, _lib.ptr(work), _lib.ptr(info), _lib.stream()),
               "tq_chol_inverse")
    torch.cuda.synchronize()
    return Hd, Hinv, int(info.item())


def atq_block(W, blk_idx=None, col0=0, b=None, s1d=None, max_iter=100, want_E=True):
    """W torch CUDA (n, ldw) fp32.  Returns alpha, mu, T(int8), E, iters as numpy."""
    L = lib()
    n = W.shape[0]
    if b is None:
        b = len(blk_idx) if blk_idx is not None else W.shape[1] - col0
    T = torch.zeros((n, b), dtype=torch.int8, device=DEV)
    a = torch.empty(n, dtype=torch.float32, device=DEV)
    u = torch.empty(n, dtype=torch.float32, device=DEV)
    E = torch.zeros((n, b), dtype=torch.float32, device=DEV) if want_E else None
    iters = torch.zeros(n, dtype=torch.int32, device=DEV)
    _lib.check(L.tq_atq_block(_lib.ptr(W), W.stride(0), n, _lib.ptr(blk_idx), col0, b, _lib.ptr(s1d), max_iter,
                              _lib.ptr(T), b, _lib.ptr(a), _lib.ptr(u), 1, _lib.ptr(E), b, _lib.ptr(iters),
                              _lib.stream()), "tq_atq_block")
    torch.cuda.synchronize()
    return (a.cpu().numpy().reshape(-1, 1), u.cpu().numpy().reshape(-1, 1), T.cpu().numpy(),
            None if E is None else E.cpu().numpy(), iters.cpu().numpy())


def aga_vector(H, blk_idx, col0, b, mode):
    L = lib()
    s1d = torch.empty(b + 1, dtype=torch.float32, device=DEV)
    _lib.check(L.tq_aga_vector(_lib.ptr(H), H.stride(0), _lib.ptr(blk_idx), col0, b, mode, _lib.ptr(s1d),
                               _lib.stream()), "tq_aga_vector")
    return s1d


def err_feedback(W, E, Hinv, blk_idx, blk0, b, rem_idx, rem0, rem):
    L = lib()
    _lib.check(L.tq_err_feedback(_lib.ptr(W), W.stride(0), W.shape[0], _lib.ptr(E), E.stride(0), _lib.ptr(Hinv),
                                 Hinv.stride(0), _lib.ptr(blk_idx), blk0, b, _lib.ptr(rem_idx), rem0, rem,
                                 _lib.stream()), "tq_err_feedback")
    torch.cuda.synchronize()
    return W


def err_feedback_tc(W, E, Hinv, blk_idx, blk0, b, rem_idx, rem0, rem):
    L = lib()
    n = W.shape[0]
    ws = torch.empty(L.tq_err_feedback_tc_workspace_floats(n, b, rem), dtype=torch.float32, device=DEV)
    _lib.check(L.tq_err_feedback_tc(_lib.ptr(W), W.stride(0), n, _lib.ptr(E), E.stride(0), _lib.ptr(Hinv),
                                    Hinv.stride(0), _lib.ptr(blk_idx), blk0, b, _lib.ptr(rem_idx), rem0, rem,
                                    _lib.ptr(ws), _lib.stream()), "tq_err_feedback_tc")
    torch.cuda.synchronize()
    return W
