"""Oracle, timing variant: the same algorithm restated on torch CPU tensors so that every stage runs on
all host cores through the library the reference itself uses (PyTorch CPU: oneMKL GEMM, LAPACK
potrf/potri, multi-threaded elementwise kernels).  TEST / BASELINE INFRASTRUCTURE -- used only by
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` arm and by ``tests/`` (which pin it to the same
golden fixtures as the numpy oracle).  Follows gptq.py:59-199, quantizer.py:32-277, reorder.py:36-143.
"""

import torch

_TINY = 1e-8


def _grid(W, T):
    b = W.shape[1]
    wt, ts, ws, t2 = (W * T).sum(1, True), T.sum(1, True), W.sum(1, True), (T * T).sum(1, True)
    den = (b * t2 - ts * ts).clamp(min=_TINY)                       # quantizer.py:99-100
    return (b * wt - ts * ws) / den, (t2 * ws - ts * wt) / den      # quantizer.py:103,106


def _round(W, alpha, mu):
    Z = (W - mu) / alpha.clamp(min=_TINY)                           # quantizer.py:125-126
    return (Z > 0.5).to(W.dtype) - (Z < -0.5).to(W.dtype)           # quantizer.py:130-132


def atq_block(W, gram_in=None, gram=None, max_iter=100, margin_out=None):
    """init (quantizer.py:49-67) -> ITF with the global stop test (:160-175) -> AGA (:207-246).
    gram_in: matrix handed to AGA as 'X' (S = X'X); gram: S itself.
    margin_out (adjudication aid): a list that receives (final, trajectory): ``final`` (n, b) is
    | |Z| - 0.5 | of the rounding that produced the returned T; ``trajectory`` (n,) is the smallest such margin the row
    met at ANY rounding on the way (the init threshold | |Wc| / (2 delta) - 0.5 | included) -- ITF can be tipped into a
    different fixed point by a tie several iterations before the last one."""
    mu = W.mean(1, keepdim=True)
    Wc = W - mu
    delta = 0.75 * Wc.abs().mean(1, keepdim=True)
    T = (Wc > delta).to(W.dtype) - (Wc < -delta).to(W.dtype)
    alpha = (T * Wc).sum(1, True) / T.abs().sum(1, True).clamp(min=_TINY)
    traj = None
    if margin_out is not None:
        traj = (0.5 * Wc.abs() / delta.clamp(min=_TINY) - 0.5).abs().amin(1)
    T_prev = torch.zeros_like(T)
    for _ in range(max_iter):
        if torch.equal(T, T_prev):
            break
        T_prev = T
        alpha, mu = _grid(W, T)
        T = _round(W, alpha, mu)
        if traj is not None:
            traj = torch.minimum(traj, (((W - mu) / alpha.clamp(min=_TINY)).abs() - 0.5).abs().amin(1))
    if margin_out is not None:
        margin_out.append((((W - mu) / alpha.clamp(min=_TINY)).abs() - 0.5).abs())
        margin_out.append(traj)
    S = gram if gram is not None else (gram_in.T @ gram_in if gram_in is not None else None)
    if S is not None:
        s1 = S.sum(1, keepdim=True)                                 # S @ 1
        d = float(s1.sum())
        v, ws1, wts1, t2s1 = T @ s1, W @ s1, (W * T) @ s1, (T * T) @ s1
        den = (d * t2s1 - v * v).clamp(min=_TINY)
        alpha, mu = (d * wts1 - v * ws1) / den, (t2s1 * ws1 - v * wts1) / den
    return alpha, mu, T


def ssr_select(W, remaining, block, gap_out=None):
    """gap_out (adjudication aid): a list that receives the similarity gap between the last column taken
    and the best column left out -- a top-k boundary closer than the fp32 noise of the similarities may legitimately
    fall the other way."""
    if remaining.numel() <= block:                                  # reorder.py:125-126
        if gap_out is not None:
            gap_out.append(float("inf"))
        return remaining, remaining[:0]
    Wr = W[:, remaining]
    wm = Wr.mean(1, keepdim=True)
    sim = ((Wr / Wr.norm(dim=0, keepdim=True).clamp(min=_TINY)).T @ (wm / wm.norm().clamp(min=_TINY))).squeeze(1)
    top = torch.topk(sim, block).indices                            # reorder.py:133
    if gap_out is not None:
        two = torch.topk(sim, block + 1).values
        gap_out.append(float(two[block - 1] - two[block]))
    keep = torch.ones(remaining.numel(), dtype=torch.bool, device=remaining.device)
    keep[top] = False
    return remaining[top], remaining[keep]


def hessian_add(H, X):
    X = X.reshape(-1, X.shape[-1]).to(H.dtype)                      # gptq.py:68-75
    H += X.T @ X
    return X.shape[0]


def damped_inverse(Hraw, nsamples, percdamp=0.01):
    H = Hraw / nsamples                                             # gptq.py:94-103
    H.diagonal().add_(percdamp * torch.diag(H).mean())
    try:
        return H, torch.cholesky_inverse(torch.linalg.cholesky(H))
    except RuntimeError:                                            # gptq.py:104-106
        return H, torch.linalg.pinv(H)


def quantize_layer(W, Hraw, nsamples, block=128, percdamp=0.01, use_ssr=True, aga="hessian", max_iter=100,
                   static_perm=None, return_margin=False):
    """gptq.py:78-199 on W's device and dtype (CPU fp32 = the reference's CPU path; a CUDA tensor runs the same ATen
    calls on the GPU, the reference's own default device, main.py:368).  static_perm: a fixed sweep order (the
    act-order extension: the reference with use_ssr=False on pre-permuted columns).  return_margin adds a
    fifth output, the adjudication aids of SURVEY 8c: {'final': | |Z| - 0.5 | per code in original column positions,
    'trajectory': (n, nb) smallest margin each row met at any rounding of the block, 'topk_gap': per block, the
    similarity gap at the SSR top-k boundary}."""
    W = W.clone()
    n, m = W.shape
    dev = W.device
    H, Hinv = damped_inverse(Hraw, nsamples, percdamp)
    dinv = torch.diag(Hinv).clamp(min=_TINY)
    T_full = torch.zeros_like(W)
    margin = torch.zeros_like(W) if return_margin else None
    traj, gaps = [], ([] if return_margin else None)
    alphas, mus, perm = [], [], []
    remaining = torch.arange(m, device=dev)
    done = 0
    while done < m:
        if use_ssr:
            blk, remaining = ssr_select(W, remaining, block, gap_out=gaps)
            rem = remaining
        elif static_perm is not None:
            hi = min(done + block, m)
            blk, rem = static_perm[done:hi], static_perm[hi:]
        else:
            hi = min(done + block, m)
            blk, rem = torch.arange(done, hi, device=dev), torch.arange(hi, m, device=dev)
        perm.append(blk)
        Wb = W[:, blk]
        mo = [] if return_margin else None
        if aga == "hessian":
            a, u, Tb = atq_block(Wb, gram_in=H[blk][:, blk], max_iter=max_iter, margin_out=mo)
        elif aga == "activations":
            a, u, Tb = atq_block(Wb, gram=Hraw[blk][:, blk], max_iter=max_iter, margin_out=mo)
        else:
            a, u, Tb = atq_block(Wb, max_iter=max_iter, margin_out=mo)
        if return_margin:
            margin[:, blk] = mo[0]
            traj.append(mo[1])
        alphas.append(a)
        mus.append(u)
        T_full[:, blk] = Tb
        if rem.numel() > 0:
            E = Wb - (a * Tb + u)
            W[:, rem] -= E @ (Hinv[blk][:, rem] / dinv[blk][:, None])   # gptq.py:173-186
        done += blk.numel()
    out = (torch.cat(alphas, 1), torch.cat(mus, 1), T_full, torch.cat(perm))
    if return_margin:
        return out + ({"final": margin, "trajectory": torch.stack(traj, 1), "topk_gap": gaps if use_ssr else None},)
    return out
