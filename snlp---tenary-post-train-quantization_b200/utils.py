"""2-bit ternary codec -- B200 mirror of the reference's ``utils.pack_ternary`` /
``utils.unpack_ternary`` (``/root/reference/utils.py:189-248``): code = T+1 in {0,1,2}, four codes
per byte (c0 | c1<<2 | c2<<4 | c3<<6) over the flat row-major tensor, zero padded."""

from typing import Tuple

import torch

try:
    from . import _lib
except ImportError:
    import _lib


def pack_ternary(T: torch.Tensor):
    """utils.py:189-219.  Returns (packed uint8[ceil(N/4)], orig_shape)."""
    lib = _lib.load()
    _lib.require_cuda(T, "T")
    orig_shape = T.shape
    flat = T.detach().contiguous().reshape(-1)
    count = flat.numel()
    packed = torch.empty((count + 3) // 4, dtype=torch.uint8, device=T.device)
    with torch.cuda.device(T.device):
        if flat.dtype == torch.int8:
            _lib.check(lib.tq_pack2b(_lib.ptr(flat), count, _lib.ptr(packed), _lib.stream()), "tq_pack2b")
        else:
            flat = flat.float() if flat.dtype != torch.float32 else flat
            _lib.check(lib.tq_pack2b_f32(_lib.ptr(flat), count, _lib.ptr(packed), _lib.stream()), "tq_pack2b_f32")
    return packed, orig_shape


def unpack_ternary(packed: torch.Tensor, orig_shape: Tuple[int, ...]) -> torch.Tensor:
    """utils.py:222-248.  int8 output in {-1, 0, 1}."""
    lib = _lib.load()
    _lib.require_cuda(packed, "packed")
    total = 1
    for d in orig_shape:
        total *= d
    if packed.numel() * 4 < total:
        raise ValueError(f"packed buffer holds {packed.numel() * 4} codes, shape {tuple(orig_shape)} needs {total}")
    out = torch.empty(total, dtype=torch.int8, device=packed.device)
    p = packed.detach().contiguous()
    with torch.cuda.device(packed.device):
        _lib.check(lib.tq_unpack2b(_lib.ptr(p), total, _lib.ptr(out), _lib.stream()), "tq_unpack2b")
    return out.reshape(orig_shape)
