"""Oracle: asymmetric ternary quantizer (ATQ) -- numpy restatement.  TEST INFRASTRUCTURE.

Follows ``/root/reference/quantizer.py`` (class ``AsymmetricTernaryQuantizer``):
ternary_init :32-69, build_optimal_grid :71-108, flexible_round :110-134,
iterative_ternary_fitting :136-175, activation_aware_grid_alignment :177-248,
quantize :250-277, dequantize :279-293, error metrics :296-306.

All functions work on 2-D arrays ``W (n, b)`` and keep the reference's shapes:
``alpha, mu`` are ``(n, 1)``, ``T`` is ``(n, b)`` in W's dtype.  ``dtype`` is taken
from ``W`` (float32 = the reference's working precision, float64 = adjudication).
"""

import numpy as np

_TINY = 1e-8  # the reference's clamp floor everywhere (quantizer.py:66,100,125,240)


def _rowsum(a):
    return a.sum(axis=1, keepdims=True)


def ternary_init(W):
    """quantizer.py:32-69.  mu = row mean; threshold 0.75*mean|W-mu| (strict);
    alpha = sum(T*(W-mu)) / max(sum|T|, 1e-8)."""
    W = np.asarray(W)
    mu = W.mean(axis=1, keepdims=True, dtype=W.dtype)
    Wc = W - mu
    delta = W.dtype.type(0.75) * np.abs(Wc).mean(axis=1, keepdims=True, dtype=W.dtype)
    T = np.zeros_like(W)
    T[Wc > delta] = 1
    T[Wc < -delta] = -1
    num = _rowsum(T * Wc)
    den = np.maximum(_rowsum(np.abs(T)), W.dtype.type(_TINY))
    return num / den, mu, T


def build_optimal_grid(W, T):
    """quantizer.py:71-108 (Eq. 9): closed-form least-squares (alpha, mu) for fixed T."""
    W = np.asarray(W)
    m = W.dtype.type(W.shape[1])
    wt = _rowsum(W * T)
    ts = _rowsum(T)
    ws = _rowsum(W)
    t2 = _rowsum(T * T)
    den = np.maximum(m * t2 - ts * ts, W.dtype.type(_TINY))
    alpha = (m * wt - ts * ws) / den
    mu = (t2 * ws - ts * wt) / den
    return alpha, mu


def flexible_round(W, alpha, mu):
    """quantizer.py:110-134 (Eq. 10): nearest of {-1,0,1} to (W-mu)/max(alpha,1e-8);
    boundaries at +-0.5 are strict, so exact ties go to 0."""
    W = np.asarray(W)
    Z = (W - mu) / np.maximum(alpha, W.dtype.type(_TINY))
    T = np.zeros_like(W)
    T[Z > 0.5] = 1
    T[Z < -0.5] = -1
    return T


def iterative_ternary_fitting(W, alpha, mu, T, max_iter=100, return_iters=False):
    """quantizer.py:136-175.  The stop test is GLOBAL over the block matrix
    (``torch.equal(T, T_prev)``, :164) with ``T_prev`` starting at zeros; the returned
    (alpha, mu) belong to the T of the previous round (= T at convergence)."""
    T_prev = np.zeros_like(T)
    iters = 0
    for _ in range(max_iter):
        if np.array_equal(T, T_prev):
            break
        T_prev = T.copy()
        alpha, mu = build_optimal_grid(W, T)
        T = flexible_round(W, alpha, mu)
        iters += 1
    if return_iters:
        return alpha, mu, T, iters
    return alpha, mu, T


def aga_from_gram(W, T, S):
    """quantizer.py:215-248 given the Gram matrix S (b, b) already formed (:207).
    Only s1 = S @ 1 and d = 1' S 1 enter the result."""
    W = np.asarray(W)
    dt = W.dtype
    S = np.asarray(S, dtype=dt)
    ones = np.ones((S.shape[0], 1), dtype=dt)
    s1 = S @ ones
    d = dt.type((ones.T @ s1).item())
    v = T @ s1
    ws1 = W @ s1
    wts1 = (W * T) @ s1
    t2s1 = (T * T) @ s1
    den = np.maximum(d * t2s1 - v * v, dt.type(_TINY))
    alpha = (d * wts1 - v * ws1) / den
    mu = (t2s1 * ws1 - v * wts1) / den
    return alpha, mu


def activation_aware_grid_alignment(W, T, X):
    """quantizer.py:177-248.  X is (rows, b) or (B, L, b); S = X' X (:207)."""
    W = np.asarray(W)
    X = np.asarray(X, dtype=W.dtype)
    if X.ndim == 3:
        X = X.reshape(-1, W.shape[1])
    return aga_from_gram(W, T, X.T @ X)


def atq_quantize(W, X=None, max_iter=100, gram=None):
    """quantizer.py:250-277: init -> ITF -> (AGA if X given).  ``gram`` lets a caller
    hand over S = X'X directly (used for the ``main.py:177-180`` variant where
    S = (X'X)[blk, blk] is a sub-block of the raw Hessian)."""
    W = np.asarray(W)
    alpha, mu, T = ternary_init(W)
    alpha, mu, T = iterative_ternary_fitting(W, alpha, mu, T, max_iter=max_iter)
    if gram is not None:
        alpha, mu = aga_from_gram(W, T, gram)
    elif X is not None:
        alpha, mu = activation_aware_grid_alignment(W, T, X)
    return alpha, mu, T


def dequantize(alpha, mu, T):
    """quantizer.py:279-293."""
    return alpha * T + mu


def compute_quantization_error(W, Wc):
    """quantizer.py:296-298."""
    return float(((W - Wc) ** 2).sum())


def compute_output_error(W, Wc, X):
    """quantizer.py:301-306."""
    X = np.asarray(X)
    if X.ndim == 3:
        X = X.reshape(-1, X.shape[-1])
    diff = (W - Wc) @ X.T
    return float((diff ** 2).sum())
