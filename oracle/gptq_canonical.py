"""Oracle for the canonical-GPTQ exactness options -- TEST INFRASTRUCTURE, numpy only.

The reference's feedback (gptq.py:173-186) divides rows of the STATIC inverse Hessian by its diagonal: it is
not the OBS-exact compensation and it never updates H^-1 for the columns already fixed.  Canonical GPTQ (Frantar et al.)
does both through the Cholesky factor of H^-1.  The options restated here, derived from first principles rather than from
the Cholesky shortcut so that the product's formulation is checked against the definition:

  feedback='block_exact'   after a block B is fixed with error E_B, the remaining columns R take the OBS-optimal update
                           dW_R = -E_B (Hc_BB)^-1 Hc_BR  where Hc is the inverse Hessian OF THE CURRENT REMAINING PROBLEM,
                           which then shrinks by the Schur complement  Hc <- Hc_RR - Hc_RB (Hc_BB)^-1 Hc_BR.
                           (= canonical GPTQ at block granularity: the ATQ grid fit of gptq.py:147-159 is joint over the
                           block, so there is no column-by-column rounding inside it.)
  dead_columns=True        input features that are identically zero in the calibration data (diag(H) == 0) get H_jj = 1
                           and their weights are set to 0 before quantisation (canonical GPTQ; the reference has no such
                           handling).

Only static sweep orders ('sequential', 'actorder'): the exact recursion needs the order up front.
"""

import numpy as np

from .atq import atq_quantize
from .gptq import _TINY, damped_inverse


def apply_dead_columns(W, Hraw, value=1.0):
    """Returns (W', Hraw', dead mask): canonical GPTQ's `dead = diag(H) == 0; H[dead, dead] = 1; W[:, dead] = 0`.
    ``value``: what the dead diagonal entries of the RAW accumulator become (the caller passes nsamples so that they are 1
    after the reference's H /= nsamples, gptq.py:94)."""
    W = np.array(W, copy=True)
    Hraw = np.array(Hraw, copy=True)
    dead = np.diag(Hraw) == 0
    if dead.any():
        idx = np.nonzero(dead)[0]
        Hraw[idx, idx] = value
        W[:, idx] = 0
    return W, Hraw, dead


def quantize_layer_exact(W, Hraw, nsamples, block_size=128, percdamp=0.01, order="sequential", aga="hessian",
                         feedback="block_exact", dead_columns=False, max_iter=100):
    if order not in ("sequential", "actorder"):
        raise ValueError("the exact feedback needs a static sweep order")
    W = np.array(W, copy=True)
    dt = W.dtype
    n, m = W.shape
    Hraw = np.asarray(Hraw, dtype=dt)
    if dead_columns:
        W, Hraw, _ = apply_dead_columns(W, Hraw, value=nsamples)
    H, Hinv = damped_inverse(Hraw, nsamples, percdamp)
    static = np.arange(m, dtype=np.int64) if order == "sequential" else np.argsort(-np.diag(H), kind="stable").astype(np.int64)
    Hc = np.array(Hinv[np.ix_(static, static)], dtype=np.float64)       # inverse Hessian of the remaining problem, sweep order
    hinv_diag = np.maximum(np.diag(Hinv), dt.type(_TINY))
    T_full = np.zeros_like(W)
    alphas, mus = [], []
    done = 0
    while done < m:
        blk = static[done:done + block_size]
        rem = static[done + block_size:]
        Wb = W[:, blk]
        if aga == "hessian":
            a, u, Tb = atq_quantize(Wb, X=H[np.ix_(blk, blk)], max_iter=max_iter)
        elif aga == "activations":
            a, u, Tb = atq_quantize(Wb, gram=Hraw[np.ix_(blk, blk)], max_iter=max_iter)
        else:
            a, u, Tb = atq_quantize(Wb, max_iter=max_iter)
        alphas.append(a)
        mus.append(u)
        T_full[:, blk] = Tb
        E = Wb - (a * Tb + u)
        b = blk.shape[0]
        if rem.shape[0] > 0:
            if feedback == "block_exact":
                C = np.linalg.solve(Hc[:b, :b], Hc[:b, b:])                     # (Hc_BB)^-1 Hc_BR
                W[:, rem] -= (E.astype(np.float64) @ C).astype(dt)
                Hc = Hc[b:, b:] - Hc[b:, :b] @ C                                # Schur complement
            else:
                C = Hinv[np.ix_(blk, rem)] / hinv_diag[blk][:, None]            # the reference's formula
                W[:, rem] -= E @ C
        done += b
    return np.concatenate(alphas, axis=1), np.concatenate(mus, axis=1), T_full, static
