"""Asymmetric Ternary Quantizer (ATQ): ternary init, ITF, AGA -- B200 mirror of the reference's
``quantizer.py`` (class ``AsymmetricTernaryQuantizer``, ``/root/reference/quantizer.py:16-293``).

Same names, argument meaning and shapes as the reference (``alpha``/``mu`` are ``(n, 1)``, ``T`` is
``(n, b)`` in W's dtype); underneath, every method is one launch of the warp-per-row CUDA kernels
in ``csrc/atq.cu`` through the C ABI (``include/tq100.h``).  Tensors must live on a CUDA device;
there is no CPU path.
"""

from typing import Optional, Tuple

import torch

try:
    from . import _lib
except ImportError:  # flat import (package directory on sys.path, like the reference's modules)
    import _lib

MAX_BLOCK = 512  # a row's block is held in registers by one warp


def _prep_w(W: torch.Tensor):
    _lib.require_cuda(W, "W")
    if W.dim() != 2:
        raise ValueError(f"W must be 2-D (n, b), got {tuple(W.shape)}")
    if W.shape[1] > MAX_BLOCK:
        raise ValueError(f"block of {W.shape[1]} columns: the ATQ kernels hold a row's block in registers "
                         f"and take at most {MAX_BLOCK} columns")
    Wf = W.detach()
    if Wf.dtype != torch.float32:
        Wf = Wf.float()
    if Wf.stride(1) != 1:
        Wf = Wf.contiguous()
    return Wf


def _codes(T: torch.Tensor) -> torch.Tensor:
    return T.detach().to(torch.int8).contiguous()


def _vec(a: torch.Tensor, n: int) -> torch.Tensor:
    return a.detach().float().reshape(n).contiguous()


class AsymmetricTernaryQuantizer:
    """quantizer.py:16-293."""

    def __init__(self, max_iter: int = 100):
        self.max_iter = max_iter

    # -- one op of the stage kernel ------------------------------------------------------------
    def _stage(self, op, W, T_in=None, alpha_in=None, mu_in=None, s1d=None, want_T=False, want_am=True):
        lib = _lib.load()
        Wf = _prep_w(W)
        n, b = Wf.shape
        T_out = torch.empty((n, b), dtype=torch.int8, device=Wf.device) if want_T else None
        a_out = torch.empty(n, dtype=torch.float32, device=Wf.device) if want_am else None
        u_out = torch.empty(n, dtype=torch.float32, device=Wf.device) if want_am else None
        with torch.cuda.device(Wf.device):
            _lib.check(lib.tq_atq_stage(op, _lib.ptr(Wf), Wf.stride(0), n, b, _lib.ptr(T_in), _lib.ptr(alpha_in),
                                        _lib.ptr(mu_in), _lib.ptr(s1d), int(self.max_iter), _lib.ptr(T_out),
                                        _lib.ptr(a_out), _lib.ptr(u_out), _lib.stream()), "tq_atq_stage")
        return a_out, u_out, T_out

    @staticmethod
    def _shape_out(W, alpha, mu, T):
        out = []
        if alpha is not None:
            out += [alpha.reshape(-1, 1).to(W.dtype), mu.reshape(-1, 1).to(W.dtype)]
        if T is not None:
            out.append(T.to(W.dtype))
        return tuple(out)

    def ternary_init(self, W: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """quantizer.py:32-69."""
        a, u, T = self._stage(_lib.OP_INIT, W, want_T=True)
        return self._shape_out(W, a, u, T)

    def build_optimal_grid(self, W: torch.Tensor, T: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """quantizer.py:71-108."""
        a, u, _ = self._stage(_lib.OP_GRID, W, T_in=_codes(T))
        return self._shape_out(W, a, u, None)

    def flexible_round(self, W: torch.Tensor, alpha: torch.Tensor, mu: torch.Tensor) -> torch.Tensor:
        """quantizer.py:110-134."""
        n = W.shape[0]
        _, _, T = self._stage(_lib.OP_ROUND, W, alpha_in=_vec(alpha, n), mu_in=_vec(mu, n), want_T=True,
                              want_am=False)
        return T.to(W.dtype)

    def iterative_ternary_fitting(self, W, alpha, mu, T):
        """quantizer.py:136-175 (per-row stop test, see csrc/atq.cu)."""
        n = W.shape[0]
        a, u, To = self._stage(_lib.OP_ITF, W, T_in=_codes(T), alpha_in=_vec(alpha, n), mu_in=_vec(mu, n),
                               want_T=True)
        return self._shape_out(W, a, u, To)

    def _s1d_from_activations(self, X: torch.Tensor, b: int) -> torch.Tensor:
        """s1 = (X'X) 1 and d = 1'(X'X)1 (quantizer.py:207-218): Gram by the Hessian kernel, then the
        AGA-vector kernel."""
        lib = _lib.load()
        _lib.require_cuda(X, "X")
        X2 = X.detach().reshape(-1, b)
        if X2.dtype not in (torch.float32, torch.float16, torch.bfloat16):
            X2 = X2.float()
        X2 = X2.contiguous()
        S = torch.zeros((b, b), dtype=torch.float32, device=X2.device)
        s1d = torch.empty(b + 1, dtype=torch.float32, device=X2.device)
        with torch.cuda.device(X2.device):
            st = _lib.stream()
            _lib.check(lib.tq_hessian_accum(_lib.ptr(S), b, _lib.ptr(X2), X2.shape[0], b, X2.stride(0),
                                            _lib.dtype_code(X2.dtype), _lib.HESS_FFMA, st), "tq_hessian_accum")
            _lib.check(lib.tq_symmetrize(_lib.ptr(S), b, b, st), "tq_symmetrize")
            _lib.check(lib.tq_aga_vector(_lib.ptr(S), b, None, 0, b, _lib.AGA_ACTIVATIONS, _lib.ptr(s1d), st),
                       "tq_aga_vector")
        return s1d

    def activation_aware_grid_alignment(self, W, T, X):
        """quantizer.py:177-248."""
        s1d = self._s1d_from_activations(X, W.shape[1])
        a, u, _ = self._stage(_lib.OP_AGA, W, T_in=_codes(T), s1d=s1d)
        return self._shape_out(W, a, u, None)

    def quantize(self, W: torch.Tensor, X: Optional[torch.Tensor] = None):
        """quantizer.py:250-277: init -> ITF -> AGA, one fused kernel launch."""
        lib = _lib.load()
        Wf = _prep_w(W)
        n, b = Wf.shape
        s1d = self._s1d_from_activations(X, b) if X is not None else None
        T = torch.empty((n, b), dtype=torch.int8, device=Wf.device)
        a = torch.empty(n, dtype=torch.float32, device=Wf.device)
        u = torch.empty(n, dtype=torch.float32, device=Wf.device)
        with torch.cuda.device(Wf.device):
            _lib.check(lib.tq_atq_block(_lib.ptr(Wf), Wf.stride(0), n, None, 0, b, _lib.ptr(s1d), int(self.max_iter),
                                        _lib.ptr(T), b, _lib.ptr(a), _lib.ptr(u), 1, None, 0, None, _lib.stream()),
                       "tq_atq_block")
        return self._shape_out(W, a, u, T)

    def dequantize(self, alpha, mu, T):
        """quantizer.py:279-293."""
        return alpha * T + mu


def compute_quantization_error(W: torch.Tensor, W_c: torch.Tensor) -> float:
    """quantizer.py:296-298."""
    return ((W - W_c) ** 2).sum().item()


def compute_output_error(W: torch.Tensor, W_c: torch.Tensor, X: torch.Tensor) -> float:
    """quantizer.py:301-306."""
    if X.dim() == 3:
        X = X.reshape(-1, X.shape[-1])
    diff = (W - W_c) @ X.T
    return (diff ** 2).sum().item()
