"""Oracle: GPTQ layer quantizer (Hessian, damped inverse, block column sweep) -- numpy
restatement.  TEST INFRASTRUCTURE.

Follows ``/root/reference/gptq.py`` (class ``GPTQ``): add_batch :59-76, quantize
:78-199, get_quantized_weight :201-230; and the second, inline implementation
``/root/reference/main.py:102-230`` (``PT2LLMQuantizer.quantize_layer``), which
differs only in what it hands to AGA (main.py:177-180 raw activations of the block
columns vs gptq.py:147-150 the damped Hessian sub-block).

Deliberate deviations from the reference, both outside its working envelope:
  * gptq.py:162-170 raises UnboundLocalError when the layer has a single block
    (SURVEY Q3); here the last block simply has no columns left to update.
  * ``order='actorder'`` is an extension (the reference has none): a static
    permutation by descending diag(H), then the sequential sweep over that order.
"""

import numpy as np
from scipy.linalg import lapack

from .atq import atq_quantize
from .ssr import select_next_block_ssr

_TINY = 1e-8


def hessian_add_batch(H, nsamples, inp):
    """gptq.py:59-76: flatten to (Nt, m); H += X' X (no factor 2, no running
    average); nsamples += Nt.  Returns the new (H, nsamples)."""
    X = np.asarray(inp)
    if X.ndim == 3:
        X = X.reshape(-1, X.shape[-1])
    X = X.astype(H.dtype, copy=False)
    H += X.T @ X
    return H, nsamples + X.shape[0]


def damped_inverse(Hraw, nsamples, percdamp=0.01):
    """gptq.py:94-106: H = Hraw / nsamples; H.diag += percdamp * mean(diag H);
    Hinv = potri(potrf(H)) (full symmetric inverse); pinv if Cholesky fails.
    Returns (H_damped, Hinv)."""
    dt = Hraw.dtype
    H = Hraw / dt.type(nsamples)
    damp = dt.type(percdamp) * np.diag(H).mean(dtype=dt)
    H[np.diag_indices_from(H)] += damp
    potrf, potri = (lapack.spotrf, lapack.spotri) if dt == np.float32 else (lapack.dpotrf, lapack.dpotri)
    L, info = potrf(H, lower=1, clean=1)
    if info == 0:
        inv, info2 = potri(L, lower=1)
        if info2 == 0:
            inv = np.tril(inv) + np.tril(inv, -1).T
            return H, np.ascontiguousarray(inv, dtype=dt)
    # torch.linalg.pinv's default cutoff (gptq.py:106): singular values below max(m, n) * eps * sigma_max are dropped
    return H, np.linalg.pinv(H, rcond=max(H.shape) * np.finfo(dt).eps).astype(dt)


def quantize_layer(W, Hraw, nsamples, block_size=128, percdamp=0.01, order="ssr",
                   aga="hessian", X=None, max_iter=100, trace=None):
    """gptq.py:78-199 (aga='hessian') / main.py:102-230 (aga='activations').

    W (n, m) weights; Hraw (m, m) accumulated X'X; nsamples token count.
    order: 'ssr' (use_ssr=True), 'sequential' (use_ssr=False), 'actorder' (extension).
    aga:   'hessian'      AGA input is the damped, normalised H[blk, blk] (gptq.py:147-150)
           'activations'  AGA Gram is X[:, blk]' X[:, blk] (main.py:177-180); taken from
                          ``X`` if given, else from Hraw[blk, blk] (same matrix)
           'none'         skip AGA (quantizer.py:274 with X=None)
    Returns alpha (n, nb), mu (n, nb), T (n, m) in ORIGINAL column positions
    (gptq.py:155), perm (m,) int64 (gptq.py:191).
    """
    W = np.array(W, copy=True)
    dt = W.dtype
    n, m = W.shape
    Hraw = np.asarray(Hraw, dtype=dt)
    H, Hinv = damped_inverse(Hraw, nsamples, percdamp)
    hinv_diag = np.maximum(np.diag(Hinv), dt.type(_TINY))

    T_full = np.zeros_like(W)
    alphas, mus, perm = [], [], []

    if order == "ssr":
        remaining = np.arange(m, dtype=np.int64)
        static = None
    elif order == "sequential":
        static = np.arange(m, dtype=np.int64)
    elif order == "actorder":
        static = np.argsort(-np.diag(H), kind="stable").astype(np.int64)
    else:
        raise ValueError(order)

    done = 0
    while done < m:
        if static is None:
            blk, remaining = select_next_block_ssr(W, remaining, block_size)   # gptq.py:130
            rem = remaining
        else:
            blk = static[done:done + block_size]                                 # gptq.py:136-137
            rem = static[done + block_size:]                                     # gptq.py:167
        perm.extend(blk.tolist())
        Wb = W[:, blk]                                                           # gptq.py:142

        if aga == "hessian":
            a, u, Tb = atq_quantize(Wb, X=H[np.ix_(blk, blk)], max_iter=max_iter)   # gptq.py:147-150
        elif aga == "activations":
            if X is not None:
                a, u, Tb = atq_quantize(Wb, X=np.asarray(X, dtype=dt).reshape(-1, m)[:, blk],
                                        max_iter=max_iter)                       # main.py:177-180
            else:
                a, u, Tb = atq_quantize(Wb, gram=Hraw[np.ix_(blk, blk)], max_iter=max_iter)
        elif aga == "none":
            a, u, Tb = atq_quantize(Wb, max_iter=max_iter)
        else:
            raise ValueError(aga)

        alphas.append(a)
        mus.append(u)
        T_full[:, blk] = Tb                                                      # gptq.py:155
        E = Wb - (a * Tb + u)                                                    # gptq.py:158-159
        if trace is not None:
            trace.setdefault("blocks", []).append(blk.copy())
            trace.setdefault("E", []).append(E.copy())
        if rem.shape[0] > 0:                                                     # gptq.py:170
            C = Hinv[np.ix_(blk, rem)] / hinv_diag[blk][:, None]                 # gptq.py:173-181
            W[:, rem] -= E @ C                                                   # gptq.py:186
        done += blk.shape[0]

    alpha = np.concatenate(alphas, axis=1)
    mu = np.concatenate(mus, axis=1)
    if trace is not None:
        trace["H"] = H
        trace["Hinv"] = Hinv
        trace["W_final"] = W
    return alpha, mu, T_full, np.asarray(perm, dtype=np.int64)


def quantize_layer_main(W, X, block_size=128, percdamp=0.01, use_ssr=True, max_iter=100):
    """main.py:102-230 end to end: H = X'X / Nt from one GEMM over all tokens
    (:128-129), AGA on raw block activations (:177-180), T as int8 (:144,:185)."""
    W = np.asarray(W)
    X = np.asarray(X, dtype=W.dtype)
    if X.ndim == 3:
        X = X.reshape(-1, X.shape[-1])
    Hraw = X.T @ X
    alpha, mu, T, perm = quantize_layer(W, Hraw, X.shape[0], block_size, percdamp,
                                        "ssr" if use_ssr else "sequential",
                                        aga="activations", X=X, max_iter=max_iter)
    return {"alpha": alpha, "mu": mu, "T": T.astype(np.int8), "perm": perm}


def get_quantized_weight(alpha, mu, T, perm, block_size=128):
    """gptq.py:201-230: Wq[:, perm[blk_k]] = alpha_k * T[:, perm[blk_k]] + mu_k."""
    n, m = T.shape
    Wq = np.zeros((n, m), dtype=alpha.dtype)
    for k in range(alpha.shape[1]):
        cols = perm[k * block_size:min((k + 1) * block_size, m)]
        Wq[:, cols] = alpha[:, k:k + 1] * T[:, cols] + mu[:, k:k + 1]
    return Wq


def reconstruction_error(W, Wq, Hraw):
    """||W X' - Wq X'||_F / ||W X'||_F evaluated through the Gram matrix:
    ||D X'||^2 = trace(D H D') (north_star's layer reconstruction error; fp64)."""
    D = (np.asarray(W, dtype=np.float64) - np.asarray(Wq, dtype=np.float64))
    H = np.asarray(Hraw, dtype=np.float64)
    Wd = np.asarray(W, dtype=np.float64)
    num = np.einsum("ij,jk,ik->", D, H, D)
    den = np.einsum("ij,jk,ik->", Wd, H, Wd)
    return float(np.sqrt(max(num, 0.0) / max(den, 1e-300)))
