"""SURVEY 8f N2, many-token path: tq_tl_gemm_tc (tcgen05 GEMM that expands the 2-bit codes in shared memory) against
the numpy oracle and against the dense-weight path.  Needs a B200: -m gpu.

Tolerance: the kernel multiplies exactly the reference's 16-bit weights (wtab), accumulates in fp32 in TMEM and rounds
the output once to the layer dtype: |err| <= eps_out * |y| + 2e-5 * sum_p |w_p x_p| with eps_out = 2^-10 (fp16) or
2^-7 (bf16)."""

import numpy as np
import pytest
import torch

from oracle import ternary_linear as otl

pytestmark = pytest.mark.gpu

TDT = {"float16": torch.float16, "bfloat16": torch.bfloat16}
EPS = {"float16": 2.0 ** -10, "bfloat16": 2.0 ** -7}


@pytest.fixture(scope="module")
def G():
    import gpu_util
    return gpu_util


def _layer(G, n, m, block, dtype, identity, bias, seed):
    import tq100
    rng = np.random.default_rng(seed)
    nb = (m + block - 1) // block
    T = rng.integers(-1, 2, size=(n, m)).astype(np.int8)
    perm = np.arange(m) if identity else rng.permutation(m)
    alpha = (0.01 + 0.02 * rng.random((n, nb))).astype(np.float32)
    mu = (0.004 * rng.standard_normal((n, nb))).astype(np.float32)
    b = (0.1 * rng.standard_normal(n)).astype(np.float32) if bias else None
    layer = tq100.TernaryLinear(m, n, block_size=block, bias=bias, dtype=TDT[dtype], device=G.DEV)
    layer.set_quantized_params(G.dev(alpha), G.dev(mu), G.dev(T), G.dev(perm.astype(np.int64)), None if b is None else G.dev(b))
    layer.fused_gemm = True
    layer.fused_max_tokens = 1 << 30
    return layer, (alpha, mu, T, perm, b)


@pytest.mark.parametrize("dtype", ["float16", "bfloat16"])
@pytest.mark.parametrize("n,m,block,M,identity,bias", [
    (256, 512, 128, 300, True, False),       # two token tiles, exact K slabs
    (200, 1000, 128, 33, False, True),       # ragged rows, K tail (1000 = 15 slabs + 40), ragged tokens, gather
    (130, 328, 64, 257, False, False),       # block 64, m % 16 != 0 (pad codes), second token tile holds one token
    (384, 2048, 128, 17, True, True),        # smallest many-token call
    (256, 512, 128, 200, False, True),       # 256-token tile (200 tokens are not padded to 512)
])
def test_tl_gemm_tc_vs_oracle(G, n, m, block, M, identity, bias, dtype):
    layer, (alpha, mu, T, perm, b) = _layer(G, n, m, block, dtype, identity, bias, seed=n + m + M)
    rng = np.random.default_rng(5)
    x = torch.from_numpy(rng.standard_normal((M, m)).astype(np.float32)).to(G.DEV).to(TDT[dtype])
    y = layer(x)
    torch.cuda.synchronize()
    assert y.dtype == TDT[dtype] and tuple(y.shape) == (M, n)
    xr = x.float().cpu().numpy().astype(np.float64)
    br = None if b is None else layer.bias.float().cpu().numpy().astype(np.float64)
    want = otl.forward(xr, alpha, mu, T, perm, br, block, dtype)
    Wq = otl.dequantized_weight(alpha, mu, T, perm, block, dtype)
    bound = EPS[dtype] * np.abs(want) + 2e-5 * (np.abs(xr) @ np.abs(Wq).T) + 1e-9
    got = y.float().cpu().numpy().astype(np.float64)
    assert np.all(np.isfinite(got))
    bad = np.abs(got - want) > bound
    assert not bad.any(), (int(bad.sum()), float(np.abs(got - want).max()))


def test_tl_gemm_tc_persistent_tiles_match_dense_path(G):
    """11008 x 4096 (86 row tiles) x 512 tokens = 172 tiles on 148 CTAs: every CTA role wraps its pipeline and both
    TMEM accumulators; compared with the dense-weight path of the same layer (fp32 GEMM on the dequantised weight)."""
    layer, _ = _layer(G, 11008, 4096, 128, "float16", False, True, seed=77)
    gen = torch.Generator(device=G.DEV).manual_seed(9)
    x = torch.randn((512, 4096), generator=gen, device=G.DEV).half()
    y = layer(x)
    W = layer._dequantize().float()
    want = x.float() @ W.T + layer.bias.float()
    bound = EPS["float16"] * want.abs() + 2e-5 * (x.float().abs() @ W.abs().T)
    assert torch.all((y.float() - want).abs() <= bound)
    layer.fused_gemm = False
    y_dense = layer(x)
    assert (y.float() - y_dense.float()).abs().max().item() <= 4 * EPS["float16"] * want.abs().max().item()


@pytest.mark.parametrize("width", [128, 256, 512])
def test_tl_gemm_tc_every_tile_width_vs_oracle(G, width, monkeypatch):
    """The same ragged call (200 rows, K tail, 300 tokens, permuted, bias) forced through each tile width."""
    monkeypatch.setenv("TQ_TL_GEMM_BN", str(width))
    test_tl_gemm_tc_vs_oracle(G, 200, 1000, 128, 300, False, True, "float16")


@pytest.mark.parametrize("tokens,dtype", [(256, "float16"), (1024, "bfloat16"), (128, "float16")])
def test_tl_gemm_tc_more_tiles_than_sms(G, tokens, dtype):
    """20480 x 1024: 160 row tiles > 148 SMs, so some CTAs run two tiles whatever the tile width the host picks
    (256 tokens -> 256-wide tiles, 1024 tokens -> 512-wide tiles with the single accumulator reused, 128 tokens ->
    128-wide tiles)."""
    layer, _ = _layer(G, 20480, 1024, 128, dtype, False, False, seed=31)
    gen = torch.Generator(device=G.DEV).manual_seed(tokens)
    x = torch.randn((tokens, 1024), generator=gen, device=G.DEV).to(TDT[dtype])
    y = layer(x)
    W = layer._dequantize().float()
    want = x.float() @ W.T
    bound = EPS[dtype] * want.abs() + 2e-5 * (x.float().abs() @ W.abs().T)
    assert torch.all((y.float() - want).abs() <= bound)


def test_tl_gemm_tc_rejects_fp32(G):
    from tq100 import _lib
    L = G.lib()
    z = torch.zeros(64, device=G.DEV)
    rc = L.tq_tl_gemm_tc(_lib.ptr(z), 1, _lib.ptr(z), 4, 16, 16, _lib.ptr(z), 0, 16, 1, None, None, None, _lib.ptr(z), 4,
                         _lib.stream())
    assert rc == -2 and b"f16 or bf16" in L.tq_last_error_string()
