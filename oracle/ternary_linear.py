"""Oracle: the ternary inference layer and its packed format -- numpy restatement.  TEST INFRASTRUCTURE.

Follows ``/root/reference/model.py``: ``TernaryLinear.forward`` :75-95, ``_dequantize`` :97-110,
``set_quantized_params`` :58-73, ``memory_footprint`` :112-127; the correct dequantisation is
``/root/reference/gptq.py:201-230`` (``oracle.get_quantized_weight``).

Two forwards are restated:
  * ``forward_reference``  -- literally what model.py:84-90 computes: gather x by perm, dequantise T block by block
    over POSITIONAL column ranges (although gptq.py:155 stores T in original positions), un-permute W's columns.
    Equal to the intended layer only when perm is the identity.
  * ``forward``            -- F.linear(x, Wq) with Wq = get_quantized_weight (gptq.py:201-230): what the product
    computes.  The dequantised weight is rounded to the layer dtype exactly like ``alpha * T + mu`` evaluated in
    that dtype (model.py:106-108): the product alpha*T is exact, the sum rounds once.

Packed layer format ("TL2", the product's ``tq_tl_pack``): see ``pack_layer``.
"""

import numpy as np

_NP = {"float32": np.float32, "float16": np.float16}


def _round_bf16(a):
    """round-to-nearest-even float32 -> bfloat16, returned as float32."""
    a = np.ascontiguousarray(a, dtype=np.float32)
    u = a.view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
    return u.astype(np.uint32).view(np.float32).reshape(a.shape)


def round_to(a, dtype):
    """value of `a` after a cast to the layer dtype ('float32' | 'float16' | 'bfloat16'), as float64."""
    a = np.asarray(a, dtype=np.float64)
    if dtype == "bfloat16":
        # float64 -> float32 -> bfloat16 double rounding cannot occur for sums of two bf16 values (<= 24 bits apart
        # are exact in float32); callers only pass such sums or bf16-exact values
        return _round_bf16(a.astype(np.float32)).astype(np.float64)
    return a.astype(_NP[dtype]).astype(np.float64)


def dequantized_weight(alpha, mu, T, perm, block_size=128, dtype="float32"):
    """gptq.py:201-230 evaluated in the layer dtype (model.py:106-108): alpha, mu are first stored in `dtype`
    (model.py:46-47, :67-68 copy_), then Wq[:, perm[blk_k]] = fl(alpha_k * T + mu_k).  float64 array of
    dtype-representable values."""
    T = np.asarray(T)
    n, m = T.shape
    a = round_to(alpha, dtype)
    u = round_to(mu, dtype)
    perm = np.asarray(perm, dtype=np.int64)
    Wq = np.zeros((n, m), dtype=np.float64)
    for k in range(a.shape[1]):
        cols = perm[k * block_size:min((k + 1) * block_size, m)]
        Wq[:, cols] = round_to(a[:, k:k + 1] * T[:, cols].astype(np.float64) + u[:, k:k + 1], dtype)
    return Wq


def forward(x, alpha, mu, T, perm, bias=None, block_size=128, dtype="float32"):
    """The intended layer: y = x Wq' (+ bias), accumulated in float64 (the product accumulates in fp32;
    tests state the tolerance)."""
    Wq = dequantized_weight(alpha, mu, T, perm, block_size, dtype)
    y = np.asarray(x, dtype=np.float64) @ Wq.T
    if bias is not None:
        y = y + np.asarray(bias, dtype=np.float64)
    return y


def forward_reference(x, alpha, mu, T, perm, bias=None, block_size=128, dtype="float32"):
    """model.py:75-95 literally (float64 accumulation)."""
    T = np.asarray(T)
    n, m = T.shape
    perm = np.asarray(perm, dtype=np.int64)
    inv_perm = np.argsort(perm, kind="stable")                               # model.py:70
    a = round_to(alpha, dtype)
    u = round_to(mu, dtype)
    x_perm = np.asarray(x, dtype=np.float64)[..., perm]                      # model.py:84
    W = np.zeros((n, m), dtype=np.float64)                                   # model.py:97-110
    for b in range(a.shape[1]):
        s, e = b * block_size, min((b + 1) * block_size, m)
        W[:, s:e] = round_to(a[:, b:b + 1] * T[:, s:e].astype(np.float64) + u[:, b:b + 1], dtype)
    y = x_perm @ W[:, inv_perm].T                                            # model.py:90
    if bias is not None:
        y = y + np.asarray(bias, dtype=np.float64)                           # model.py:92-93
    return y


def pack_layer(T, perm):
    """TL2 codes: uint32 [n, ceil(m/16)]; word w of row r holds sweep positions 16w..16w+15 as two bit planes:
    bit j = 1 iff T[r, perm[16w+j]] = +1, bit 16+j = 1 iff it is -1 (neither: 0); positions >= m hold 0.
    Bit-exact bar."""
    T = np.asarray(T).astype(np.int64)
    n, m = T.shape
    perm = np.asarray(perm, dtype=np.int64)
    wpr = (m + 15) // 16
    Tp = np.zeros((n, wpr * 16), dtype=np.int64)
    Tp[:, :m] = T[:, perm]
    Tp = Tp.reshape(n, wpr, 16)
    words = np.zeros((n, wpr), dtype=np.uint32)
    for j in range(16):
        words |= (Tp[:, :, j] == 1).astype(np.uint32) << np.uint32(j)
        words |= (Tp[:, :, j] == -1).astype(np.uint32) << np.uint32(16 + j)
    return words


def unpack_layer(words, m, perm):
    """inverse of pack_layer: int8 T [n, m] in original column positions."""
    words = np.asarray(words, dtype=np.uint32)
    n, wpr = words.shape
    perm = np.asarray(perm, dtype=np.int64)
    c = np.empty((n, wpr, 16), dtype=np.int8)
    for j in range(16):
        plus = ((words >> np.uint32(j)) & np.uint32(1)).astype(np.int8)
        minus = ((words >> np.uint32(16 + j)) & np.uint32(1)).astype(np.int8)
        c[:, :, j] = plus - minus
    Tp = c.reshape(n, wpr * 16)[:, :m]
    T = np.empty((n, m), dtype=np.int8)
    T[:, perm] = Tp
    return T


def weight_table(alpha, mu, dtype="float32"):
    """[n, nb, 4] = (fl(mu - alpha), mu, fl(alpha + mu), 0) in the layer dtype, as float32 (the product's tq_tl_wtab)."""
    a = round_to(alpha, dtype)
    u = round_to(mu, dtype)
    out = np.zeros(a.shape + (4,), dtype=np.float32)
    out[..., 0] = round_to(-a + u, dtype)
    out[..., 1] = u
    out[..., 2] = round_to(a + u, dtype)
    return out


def memory_footprint_reference(n, m, nb, has_bias):
    """model.py:112-127: int8 T, fp16 alpha/mu, int64 perm (inv_perm is not counted there), fp16 bias."""
    return n * m + 2 * n * nb * 2 + m * 8 + (n * 2 if has_bias else 0)
