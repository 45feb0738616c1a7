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


# ---------------------------------------------------------------------- model-level helpers (host only)
def set_seed(seed: int):
    """utils.py:15-21."""
    import random
    import numpy as np
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)


def compute_bits_per_weight(model: torch.nn.Module, include_scales: bool = True) -> float:
    """utils.py:251-285: 1.58 bits per ternary weight + 16 bits per alpha/mu entry over every ternary layer;
    16.0 for a model without one.  (The codes are stored at 2 bits per weight; ``stored_bits_per_weight`` reports
    what the buffers really occupy.)"""
    weights = 0
    bits = 0.0
    for module in model.modules():
        if hasattr(module, "codes") and hasattr(module, "alpha") and hasattr(module, "mu"):
            count = module.out_features * module.in_features
            weights += count
            bits += count * 1.58
            if include_scales:
                bits += (module.alpha.numel() + module.mu.numel()) * 16
    return bits / weights if weights else 16.0


def stored_bits_per_weight(model: torch.nn.Module) -> float:
    """Resident bits per weight of the ternary layers (2-bit codes, scales, int32 permutation, bias)."""
    weights = 0
    nbytes = 0
    for module in model.modules():
        if hasattr(module, "codes") and hasattr(module, "memory_footprint"):
            weights += module.out_features * module.in_features
            nbytes += module.memory_footprint()
    return 8.0 * nbytes / weights if weights else 16.0


def save_quantized_model(model: torch.nn.Module, save_path: str, quantized_params):
    """utils.py:288-296: one file holding the state dict and the per-layer parameter dicts."""
    torch.save({"model_state_dict": model.state_dict(), "quantized_params": quantized_params}, save_path)


def load_quantized_model(model: torch.nn.Module, load_path: str):
    """utils.py:299-304."""
    checkpoint = torch.load(load_path, map_location="cpu")
    model.load_state_dict(checkpoint["model_state_dict"])
    return model, checkpoint.get("quantized_params", {})
