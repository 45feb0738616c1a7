"""CPU-side checks of the boundary: the C-ABI library loads, exports every symbol include/tq100.h
declares, and the Python mirror refuses to run without CUDA tensors (no CPU fallback)."""

import os
import re

import pytest
import torch

import tq100
from tq100 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "tq100.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(tq_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    handle = _lib.load()
    syms = _header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(handle, s), f"libtq100.so does not export {s}"
    assert sorted(_lib.EXPORTS) == syms, "ctypes signature table and header disagree"
    assert handle.tq_abi_version() == 1


def test_size_queries_need_no_gpu():
    handle = _lib.load()
    assert handle.tq_chol_workspace_floats(4096) == 3 * 4096 * 4096 + 32 * 128 * 128   # L^-1, (hi, lo) splits, diagonal inverses
    assert handle.tq_ssr_num_chunks(4096) == 32
    assert handle.tq_sweep_workspace_bytes(4096, 4096, 128) > 4096 * 4096


def test_bad_arguments_return_status_and_message():
    handle = _lib.load()
    rc = handle.tq_pack2b(None, 16, None, None)
    assert rc == -1
    assert b"tq_pack2b" in handle.tq_last_error_string()


def test_packed_layer_argument_checks_need_no_gpu():
    """tq_tl_* validate before touching CUDA: NULL pointers -> TQ_E_BADARG, a block size that lets a code word
    straddle two scale blocks -> TQ_E_UNSUPPORTED, fp32 activations for the tensor-core GEMM -> TQ_E_UNSUPPORTED."""
    import ctypes
    handle = _lib.load()
    assert handle.tq_tl_words_per_row(4096) == 256 and handle.tq_tl_words_per_row(17) == 2
    assert handle.tq_tl_pack(None, 4, 16, None, None, 1, None) == -1
    assert b"tq_tl_pack" in handle.tq_last_error_string()
    buf = (ctypes.c_float * 64)()                       # host memory: only its address is inspected before the status
    p = ctypes.cast(buf, ctypes.c_void_p)
    assert handle.tq_tl_gemv(p, 1, p, 4, 16, 24, p, 0, 16, 1, None, None, p, 4, None) == -2
    assert b"multiple of 16" in handle.tq_last_error_string()
    assert handle.tq_tl_gemv(p, 0, p, 4, 16, 16, p, 0, 16, 1, None, None, p, 4, None) == -1     # wpr < ceil(m/16)
    assert handle.tq_tl_gemm_tc(p, 1, p, 4, 16, 16, p, 0, 16, 1, None, None, None, p, 4, None) == -2
    assert b"f16 or bf16" in handle.tq_last_error_string()
    assert handle.tq_tl_dequant(p, 1, p, 4, 16, 20, None, None, p, 0, 16, None) == -2
    assert handle.tq_tl_wtab(p, p, 4, 1, 9, p, None) == -1
    assert b"unknown dtype" in handle.tq_last_error_string()


def test_no_cpu_fallback():
    q = tq100.AsymmetricTernaryQuantizer()
    with pytest.raises(RuntimeError, match="CUDA"):
        q.quantize(torch.randn(4, 128))
    with pytest.raises(RuntimeError, match="CUDA"):
        tq100.pack_ternary(torch.zeros(8))
    with pytest.raises(RuntimeError, match="CUDA"):
        tq100.GPTQ(torch.nn.Linear(128, 8, bias=False))
    with pytest.raises(RuntimeError, match="CUDA"):
        tq100.select_next_block_ssr(torch.randn(4, 300), torch.arange(300), 128)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "snlp---tenary-post-train-quantization_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f"{f} references the oracle"


def test_reference_api_surface():
    g = tq100.gptq
    assert g.GPTQ.fasterquant is g.GPTQ.quantize
    for name in ("add_batch", "quantize", "get_quantized_weight"):
        assert callable(getattr(g.GPTQ, name))
    for name in ("ternary_init", "build_optimal_grid", "flexible_round", "iterative_ternary_fitting",
                 "activation_aware_grid_alignment", "quantize", "dequantize"):
        assert callable(getattr(tq100.AsymmetricTernaryQuantizer, name))
    with pytest.raises(ValueError, match="not prepared"):
        g.GPTQQuantizer(None).quantize_layer("missing")
