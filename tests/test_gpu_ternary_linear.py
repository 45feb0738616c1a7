"""SURVEY 8f N2 on the GPU: the packed ternary layer (tq_tl_* through the C ABI and through the TernaryLinear
mirror) against the numpy oracle and the outputs of the reference's own TernaryLinear.  Needs a B200: -m gpu.

Bars: codes, unpacked T and the dequantised weight are bit-exact; y = x Wq' is accumulated in fp32 by the kernel and
in float64 by the oracle -- tolerance 1e-5 of sum_p |w_p x_p| per output (fp32 inputs), fp16 outputs additionally
rounded to 11 bits (3e-3 of the output scale, the same tolerance the CPU tests use against the reference)."""

import os

import numpy as np
import pytest
import torch
import torch.nn as nn

from oracle import ternary_linear as otl

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TDT = {"float32": torch.float32, "float16": torch.float16, "bfloat16": torch.bfloat16}


@pytest.fixture(scope="module")
def G():
    import gpu_util
    return gpu_util


def _random_layer(n, m, block, seed, identity=False):
    rng = np.random.default_rng(seed)
    nb = (m + block - 1) // block
    T = rng.integers(-1, 2, size=(n, m)).astype(np.int8)
    perm = np.arange(m) if identity else rng.permutation(m)
    alpha = (0.01 + 0.02 * rng.random((n, nb))).astype(np.float32)
    mu = (0.004 * rng.standard_normal((n, nb))).astype(np.float32)
    return T, perm.astype(np.int64), alpha, mu


def _pack(G, T, perm):
    from tq100 import _lib
    L = G.lib()
    n, m = T.shape
    wpr = L.tq_tl_words_per_row(m)
    codes = torch.zeros((n, wpr), dtype=torch.int32, device=G.DEV)
    Td = G.dev(T)
    pd = None if perm is None else G.i32(perm)
    _lib.check(L.tq_tl_pack(_lib.ptr(Td), n, m, _lib.ptr(pd), _lib.ptr(codes), wpr, _lib.stream()), "tq_tl_pack")
    torch.cuda.synchronize()
    return codes, pd


def _wtab(G, alpha, mu, dtype):
    from tq100 import _lib
    L = G.lib()
    n, nb = alpha.shape
    wtab = torch.empty((n, nb, 4), dtype=torch.float32, device=G.DEV)
    a, u = G.dev(alpha), G.dev(mu)          # named: a temporary would be freed (and its block reused) before the launch
    _lib.check(L.tq_tl_wtab(_lib.ptr(a), _lib.ptr(u), n, nb, _lib.dtype_code(TDT[dtype]),
                            _lib.ptr(wtab), _lib.stream()), "tq_tl_wtab")
    torch.cuda.synchronize()
    return wtab


@pytest.mark.parametrize("n,m", [(5, 16), (3, 37), (9, 320), (64, 4096), (7, 1001)])
def test_tl_pack_unpack_bit_exact(G, n, m):
    from tq100 import _lib
    L = G.lib()
    assert L.tq_tl_words_per_row(m) == (m + 15) // 16
    T, perm, _, _ = _random_layer(n, m, 128, seed=n * 1000 + m)
    for p in (perm, None):
        codes, pd = _pack(G, T, p)
        want = otl.pack_layer(T, np.arange(m) if p is None else p)
        assert np.array_equal(codes.cpu().numpy().view(np.uint32), want)
        inv = None if p is None else G.i32(np.argsort(p))
        for ip in ((None,) if p is None else (None, inv)):          # scatter by perm / gather by inv_perm
            out = torch.full((n, m), 7, dtype=torch.int8, device=G.DEV)
            _lib.check(L.tq_tl_unpack(_lib.ptr(codes), codes.shape[1], n, m, _lib.ptr(pd), _lib.ptr(ip), _lib.ptr(out),
                                      _lib.stream()), "tq_tl_unpack")
            assert np.array_equal(out.cpu().numpy(), T)


@pytest.mark.parametrize("dtype", ["float32", "float16", "bfloat16"])
def test_tl_wtab_and_dequant_bit_exact(G, dtype):
    from tq100 import _lib
    L = G.lib()
    n, m, block = 37, 1000, 64
    T, perm, alpha, mu = _random_layer(n, m, block, seed=5)
    wtab = _wtab(G, alpha, mu, dtype)
    assert np.array_equal(wtab.cpu().numpy(), otl.weight_table(alpha, mu, dtype))
    codes, pd = _pack(G, T, perm)
    want = otl.dequantized_weight(alpha, mu, T, perm, block, dtype)
    inv = G.i32(np.argsort(perm))
    for ip in (None, inv):                                          # scatter by perm / gather by inv_perm
        W = torch.zeros((n, m), dtype=TDT[dtype], device=G.DEV)
        _lib.check(L.tq_tl_dequant(_lib.ptr(codes), codes.shape[1], _lib.ptr(wtab), n, m, block, _lib.ptr(pd), _lib.ptr(ip),
                                   _lib.ptr(W), _lib.dtype_code(TDT[dtype]), m, _lib.stream()), "tq_tl_dequant")
        assert np.array_equal(W.float().cpu().numpy().astype(np.float64), want)


@pytest.mark.parametrize("xdtype", ["float32", "float16", "bfloat16"])
@pytest.mark.parametrize("n,m,block,identity", [(48, 320, 128, False), (33, 2500, 64, False), (130, 4224, 128, True),
                                                 (17, 100, 128, False), (70, 11008, 128, False)])
def test_tl_gemv_vs_oracle(G, n, m, block, identity, xdtype):
    from tq100 import _lib
    L = G.lib()
    T, perm, alpha, mu = _random_layer(n, m, block, seed=n + m, identity=identity)
    wdtype = "float16" if xdtype == "float16" else "float32"
    wtab = _wtab(G, alpha, mu, wdtype)
    codes, pd = _pack(G, T, None if identity else perm)
    rng = np.random.default_rng(99)
    bias = (0.1 * rng.standard_normal(n)).astype(np.float32)
    Wq = otl.dequantized_weight(alpha, mu, T, perm, block, wdtype)
    for M in (1, 2, 3, 4, 5, 9):
        x = torch.from_numpy(rng.standard_normal((M, m)).astype(np.float32)).to(G.DEV).to(TDT[xdtype])
        xr = x.float().cpu().numpy().astype(np.float64)
        for b in (None, bias):
            y = torch.full((M, n), float("nan"), dtype=torch.float32, device=G.DEV)
            bd = None if b is None else G.dev(b)
            _lib.check(L.tq_tl_gemv(_lib.ptr(codes), codes.shape[1], _lib.ptr(wtab), n, m, block, _lib.ptr(x),
                                    _lib.dtype_code(x.dtype), m, M, _lib.ptr(pd), _lib.ptr(bd),
                                    _lib.ptr(y), n, _lib.stream()), "tq_tl_gemv")
            torch.cuda.synchronize()
            want = otl.forward(xr, alpha, mu, T, perm, b, block, wdtype)
            bound = 1e-5 * (np.abs(xr) @ np.abs(Wq).T + (0 if b is None else np.abs(b))) + 1e-12
            got = y.cpu().numpy().astype(np.float64)
            assert np.all(np.isfinite(got))
            assert np.all(np.abs(got - want) <= bound), (M, b is not None, np.abs(got - want).max())


def test_tl_rejects_bad_block(G):
    from tq100 import _lib
    L = G.lib()
    T, perm, alpha, mu = _random_layer(4, 200, 100, seed=1)
    wtab = _wtab(G, alpha, mu, "float32")
    codes, pd = _pack(G, T, perm)
    x = torch.zeros((1, 200), device=G.DEV)
    y = torch.zeros((1, 4), device=G.DEV)
    rc = L.tq_tl_gemv(_lib.ptr(codes), codes.shape[1], _lib.ptr(wtab), 4, 200, 100, _lib.ptr(x), 0, 200, 1, _lib.ptr(pd),
                      None, _lib.ptr(y), 4, _lib.stream())
    assert rc == -2 and b"multiple of 16" in L.tq_last_error_string()


# ---------------------------------------------------------------------- the layer mirror
@pytest.mark.parametrize("dtype", ["float32", "float16"])
@pytest.mark.parametrize("tag", ["seq", "ssr"])
def test_ternary_linear_layer_vs_oracle_and_reference(G, tag, dtype):
    import tq100
    g = np.load(os.path.join(GOLD, f"gptq_small_{tag}.npz"))
    t = np.load(os.path.join(GOLD, "ternary_linear.npz"))
    n, m = g["T"].shape
    x = t[f"x_{tag}"]
    for with_bias in (False, True):
        layer = tq100.TernaryLinear(m, n, block_size=128, bias=with_bias, dtype=TDT[dtype], device=G.DEV)
        layer.set_quantized_params(G.dev(g["alpha"]), G.dev(g["mu"]), G.dev(g["T"]), G.dev(g["perm"]),
                                   G.dev(t[f"bias_{tag}"]) if with_bias else None)
        assert np.array_equal(layer.T.cpu().numpy(), g["T"])
        assert np.array_equal(layer.perm.cpu().numpy(), g["perm"])
        Wd = layer._dequantize()
        assert Wd.dtype == TDT[dtype]
        want_W = otl.dequantized_weight(g["alpha"], g["mu"], g["T"], g["perm"], 128, dtype)
        assert np.array_equal(Wd.float().cpu().numpy().astype(np.float64), want_W)
        xin = G.dev(x).to(TDT[dtype])
        xr = xin.float().cpu().numpy().astype(np.float64)
        bias = layer.bias.float().cpu().numpy().astype(np.float64) if with_bias else None
        want = otl.forward(xr, g["alpha"], g["mu"], g["T"], g["perm"], bias, 128, dtype)
        scale = np.abs(want).max()
        tol = (3e-3 if dtype == "float16" else 1e-5) * scale
        y = layer(xin)
        assert y.dtype == TDT[dtype] and tuple(y.shape) == (x.shape[0], n)
        assert np.abs(y.float().cpu().numpy() - want).max() <= tol
        # many-token path (dense weight + library GEMM) and a 3-D input
        xl = torch.cat([xin] * 8, 0).reshape(2, -1, m)
        yl = layer(xl)
        assert tuple(yl.shape) == (2, xl.shape[1], n)
        assert np.abs(yl.float().cpu().numpy().reshape(-1, n)[: x.shape[0]] - want).max() <= tol
        if tag == "seq":
            # identity permutation: the reference's own TernaryLinear output is the bar (model.py:75-95)
            ref = t[f"y_seq_{dtype}_{'bias' if with_bias else 'nobias'}"]
            assert np.abs(y.float().cpu().numpy() - ref).max() <= (6e-3 if dtype == "float16" else 1e-5) * scale
            if not with_bias:
                assert np.array_equal(Wd.float().cpu().numpy(), t[f"W_seq_{dtype}"])
        # state dict round trip into a fresh layer
        clone = tq100.TernaryLinear(m, n, block_size=128, bias=with_bias, dtype=TDT[dtype], device=G.DEV)
        clone.load_state_dict(layer.state_dict())
        assert torch.equal(clone(xin), y)


def test_quantize_then_replace_linear(G):
    """GPTQ.quantize (SSR) -> replace_linear_with_ternary -> the model's output is x @ get_quantized_weight()'."""
    import synth
    import tq100
    n, m = 96, 384
    W = synth.make_weight(n, m, seed=21)
    X = synth.make_activations(4, 128, m, seed=22, lam=0.5)
    model = nn.Sequential(nn.Linear(m, n, bias=True)).to(G.DEV)
    model[0].weight.data = G.dev(W)
    g = tq100.GPTQ(model[0], block_size=128, percdamp=0.01)
    g.add_batch(G.dev(X))
    alpha, mu, T, perm = g.quantize(use_ssr=True)
    Wq = g.get_quantized_weight()
    assert not torch.equal(perm, torch.arange(m, device=perm.device))
    tq100.replace_linear_with_ternary(model, {"0": {"alpha": alpha, "mu": mu, "T": T, "perm": perm}}, block_size=128)
    assert isinstance(model[0], tq100.TernaryLinear) and model[0].alpha.dtype == torch.float32
    x = torch.randn(3, m, device=G.DEV)
    want = x.double() @ Wq.double().T + model[0].bias.double()
    got = model(x).double()
    assert (got - want).abs().max().item() <= 1e-5 * want.abs().max().item()
    assert abs(tq100.compute_bits_per_weight(model) - (1.58 + 16 * 2 * 3 / 384)) < 1e-9
    assert tq100.stored_bits_per_weight(model) < 3.0


def test_full_size_gemv_matches_dense_path(G):
    """7B down_proj shape (4096 x 11008): decode kernel vs dense weight + fp32 library GEMM; codes round trip."""
    import tq100
    n, m = 4096, 11008
    gen = torch.Generator(device=G.DEV).manual_seed(3)
    T = torch.randint(-1, 2, (n, m), generator=gen, device=G.DEV, dtype=torch.int8)
    perm = torch.randperm(m, generator=gen, device=G.DEV)
    alpha = 0.01 + 0.02 * torch.rand((n, 86), generator=gen, device=G.DEV)
    mu = 0.004 * torch.randn((n, 86), generator=gen, device=G.DEV)
    layer = tq100.TernaryLinear(m, n, bias=False, dtype=torch.float32, device=G.DEV)
    layer.set_quantized_params(alpha, mu, T, perm)
    assert torch.equal(layer.T, T)
    x = torch.randn((4, m), generator=gen, device=G.DEV)
    y = layer(x)
    W = layer._dequantize().double()
    want = x.double() @ W.T
    bound = 1e-5 * (x.double().abs() @ W.abs().T)
    assert torch.all((y.double() - want).abs() <= bound)
