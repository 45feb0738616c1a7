"""Oracle: 2-bit ternary codec -- numpy restatement.  TEST INFRASTRUCTURE.

Follows ``/root/reference/utils.py``: pack_ternary :189-219, unpack_ternary :222-248.
Integer/byte work: the parity bar is bit-exact.
"""

import numpy as np


def pack_ternary(T):
    """utils.py:189-219: code = T+1 in {0,1,2}; flat row-major; zero-pad to a
    multiple of 4; byte = c0 | c1<<2 | c2<<4 | c3<<6.  Returns (uint8[ceil(N/4)], shape)."""
    T = np.asarray(T)
    shape = tuple(T.shape)
    flat = (T + 1).astype(np.uint8).reshape(-1)
    pad = (-flat.shape[0]) % 4
    if pad:
        flat = np.concatenate([flat, np.zeros(pad, dtype=np.uint8)])
    q = flat.reshape(-1, 4)
    packed = q[:, 0] | (q[:, 1] << 2) | (q[:, 2] << 4) | (q[:, 3] << 6)
    return packed.astype(np.uint8), shape


def unpack_ternary(packed, shape):
    """utils.py:222-248: inverse of pack_ternary; int8 output in {-1,0,1}."""
    packed = np.asarray(packed, dtype=np.uint8)
    out = np.empty(packed.shape[0] * 4, dtype=np.int8)
    for k in range(4):
        out[k::4] = ((packed >> (2 * k)) & 3).astype(np.int8)
    out -= 1
    total = int(np.prod(shape)) if len(shape) else 1
    return out[:total].reshape(shape)
