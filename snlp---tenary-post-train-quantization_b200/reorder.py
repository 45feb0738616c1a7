"""Structural Similarity-based Reordering (SSR), dynamic block selection -- B200 mirror of the
reference's ``reorder.py`` (``/root/reference/reorder.py:36-61``, ``:107-143``).

``select_next_block_ssr`` keeps the reference's signature; underneath it is the two-stage CUDA
path of ``csrc/ssr.cu`` (column statistics over W[:, remaining], then a single-CTA radix select).
Ties between equal similarities go to the lower position in ``remaining_indices`` (``torch.topk``
leaves that unspecified, reorder.py:133).
"""

from typing import List, Tuple

import torch

try:
    from . import _lib
except ImportError:
    import _lib


def _prep(W, indices):
    _lib.require_cuda(W, "W")
    Wf = W.detach()
    if Wf.dtype != torch.float32:
        Wf = Wf.float()
    if Wf.stride(1) != 1:
        Wf = Wf.contiguous()
    idx = indices.detach().to(device=Wf.device, dtype=torch.int32).contiguous()
    return Wf, idx


def _stats(Wf, idx):
    lib = _lib.load()
    n, rem = Wf.shape[0], idx.numel()
    chunks = lib.tq_ssr_num_chunks(n)
    rowmean = torch.empty(n, dtype=torch.float32, device=Wf.device)
    partials = torch.empty((chunks, 2, rem), dtype=torch.float32, device=Wf.device)
    with torch.cuda.device(Wf.device):
        _lib.check(lib.tq_ssr_stats(_lib.ptr(Wf), Wf.stride(0), n, _lib.ptr(idx), rem, _lib.ptr(rowmean),
                                    _lib.ptr(partials), _lib.stream()), "tq_ssr_stats")
    return rowmean, partials, chunks


def _select(Wf, idx, block_size):
    lib = _lib.load()
    n, rem = Wf.shape[0], idx.numel()
    rowmean, partials, chunks = _stats(Wf, idx)
    k = min(block_size, rem - 1) if rem > 1 else 0
    blk = torch.empty(max(k, 1), dtype=torch.int32, device=Wf.device)
    new_rem = torch.empty(max(rem - k, 1), dtype=torch.int32, device=Wf.device)
    sims = torch.empty(2 * rem + 2, dtype=torch.float32, device=Wf.device)
    with torch.cuda.device(Wf.device):
        _lib.check(lib.tq_ssr_select(_lib.ptr(partials), chunks, _lib.ptr(rowmean), n, None, _lib.ptr(idx), rem, k,
                                     _lib.ptr(blk), _lib.ptr(new_rem), _lib.ptr(sims), _lib.stream()),
                   "tq_ssr_select")
    return blk[:k], new_rem[:rem - k], sims[:rem]


def compute_column_similarity_to_mean(W: torch.Tensor, indices: torch.Tensor) -> torch.Tensor:
    """reorder.py:36-61: cosine similarity of each remaining column to the mean remaining column."""
    Wf, idx = _prep(W, indices)
    if idx.numel() == 1:
        # one column is its own mean: similarity = ||c||^2 / max(||c||, 1e-8)^2 (reorder.py:55-59); the selection
        # kernel needs at least two remaining columns, and no caller on the sweep path asks for this case
        nrm = Wf[:, idx.long()].norm().clamp(min=1e-8)
        return ((Wf[:, idx.long()] / nrm) ** 2).sum().to(W.dtype)
    _, _, sims = _select(Wf, idx, 1)
    return sims.to(W.dtype).squeeze()


def select_next_block_ssr(W: torch.Tensor, remaining_indices: torch.Tensor,
                          block_size: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """reorder.py:107-143.  Returns (block_indices, new_remaining), same dtype/device as the input list."""
    if len(remaining_indices) <= block_size:                       # reorder.py:125-126
        return remaining_indices, torch.tensor([], dtype=remaining_indices.dtype,
                                               device=remaining_indices.device)
    Wf, idx = _prep(W, remaining_indices)
    blk, new_rem, _ = _select(Wf, idx, block_size)
    return (blk.to(dtype=remaining_indices.dtype, device=remaining_indices.device),
            new_rem.to(dtype=remaining_indices.dtype, device=remaining_indices.device))


class SSRReorderer:
    """reorder.py:146-189, dynamic mode (the only mode the quantize path uses, reorder.py:166-173).
    The static O(m^2) greedy pre-ordering (reorder.py:64-104) is outside the hot path."""

    def __init__(self, W: torch.Tensor, block_size: int = 128, use_dynamic: bool = True):
        if not use_dynamic:
            raise NotImplementedError("static SSR pre-ordering (reorder.py:64-104) is not part of the GPTQ hot path")
        self.block_size = block_size
        self.use_dynamic = use_dynamic
        self.n, self.m = W.shape
        self.perm = torch.arange(self.m, device=W.device)
        self.inv_perm = self.perm.clone()

    def get_permutation(self):
        return self.perm, self.inv_perm

    def reorder_weights(self, W):
        return W[:, self.perm]

    def reorder_activations(self, X):
        return X[..., self.perm]

    def restore_order(self, W):
        return W[:, self.inv_perm]


def apply_permutation(W: torch.Tensor, perm: torch.Tensor) -> torch.Tensor:
    """reorder.py:192-194."""
    return W[:, perm]


def apply_permutation_to_input(X: torch.Tensor, perm: torch.Tensor) -> torch.Tensor:
    """reorder.py:197-204."""
    if X.dim() == 2:
        return X[:, perm]
    if X.dim() == 3:
        return X[:, :, perm]
    raise ValueError(f"Unexpected input dimension: {X.dim()}")
