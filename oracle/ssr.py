"""Oracle: SSR dynamic block selection -- numpy restatement.  TEST INFRASTRUCTURE.

Follows ``/root/reference/reorder.py``: compute_column_similarity_to_mean :36-61,
select_next_block_ssr :107-143.

Tie rule: ``torch.topk`` leaves the order of equal similarities unspecified
(reorder.py:133); the oracle (and the CUDA path) break ties towards the LOWER
position in ``remaining`` so the result is deterministic.
"""

import numpy as np

_TINY = 1e-8


def column_similarity_to_mean(W, indices):
    """reorder.py:36-61: cosine similarity of every remaining column to the mean of
    the remaining columns, with the 1e-8 norm clamps of :55-56."""
    W = np.asarray(W)
    dt = W.dtype
    Wr = W[:, np.asarray(indices, dtype=np.int64)]
    w_mean = Wr.mean(axis=1, keepdims=True, dtype=dt)
    w_mean_n = w_mean / np.maximum(np.sqrt((w_mean * w_mean).sum(dtype=dt)), dt.type(_TINY))
    col_norm = np.sqrt((Wr * Wr).sum(axis=0, keepdims=True, dtype=dt))
    Wr_n = Wr / np.maximum(col_norm, dt.type(_TINY))
    return (Wr_n.T @ w_mean_n).reshape(-1)


def select_next_block_ssr(W, remaining, block_size):
    """reorder.py:107-143.  Returns (block_indices, new_remaining), int64.
    If no more than ``block_size`` columns remain they are all returned in their
    current (ascending) order (:125-126); otherwise the block is the top-k by
    similarity in DESCENDING similarity order (:133-136) and the rest keep their order
    (:139-141)."""
    remaining = np.asarray(remaining, dtype=np.int64)
    if remaining.shape[0] <= block_size:
        return remaining, np.zeros((0,), dtype=np.int64)
    sim = column_similarity_to_mean(W, remaining)
    order = np.argsort(-sim, kind="stable")          # descending, ties -> lower position
    top = order[:block_size]
    keep = np.ones(remaining.shape[0], dtype=bool)
    keep[top] = False
    return remaining[top], remaining[keep]
