"""SURVEY 8f N2 on CPU: the ternary-layer oracle against outputs of the unmodified reference ``TernaryLinear``
(tests/golden/ternary_linear.npz, made by tests/golden/make_golden_tl.py), the packed-layer format, and the host
side of the mirror (no compute: this box has no GPU)."""

import os

import numpy as np
import pytest
import torch
import torch.nn as nn

from oracle import ternary_linear as otl

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
# y is a sum of 320 products of O(1) inputs and O(0.02) weights; fp32 accumulation order differs between MKL and
# the float64 oracle, fp16 outputs are rounded to 11 bits (and the bias add rounds once more)
TOL = {"float32": 2e-6, "float16": 3e-3}


def _case(tag):
    g = np.load(os.path.join(GOLD, f"gptq_small_{tag}.npz"))
    t = np.load(os.path.join(GOLD, "ternary_linear.npz"))
    return g, t


@pytest.mark.parametrize("dtype", ["float32", "float16"])
@pytest.mark.parametrize("tag", ["seq", "ssr"])
def test_forward_reference_restatement_matches_reference(tag, dtype):
    g, t = _case(tag)
    x = t[f"x_{tag}"]
    xin = x.astype(np.float16).astype(np.float64) if dtype == "float16" else x
    for with_bias in (False, True):
        want = t[f"y_{tag}_{dtype}_{'bias' if with_bias else 'nobias'}"].astype(np.float64)
        bias = t[f"bias_{tag}"] if with_bias else None
        if bias is not None and dtype == "float16":
            bias = bias.astype(np.float16).astype(np.float64)
        got = otl.forward_reference(xin, g["alpha"], g["mu"], g["T"], g["perm"], bias, 128, dtype)
        scale = np.abs(want).max()
        assert np.abs(got - want).max() <= TOL[dtype] * scale * (2 if with_bias else 1), (tag, dtype, with_bias)


@pytest.mark.parametrize("dtype", ["float32", "float16"])
def test_identity_permutation_reference_equals_intended_layer(dtype):
    """use_ssr=False: perm is the identity, the reference's forward IS F.linear(x, get_quantized_weight())."""
    g, t = _case("seq")
    assert np.array_equal(g["perm"], np.arange(g["perm"].shape[0]))
    Wq = otl.dequantized_weight(g["alpha"], g["mu"], g["T"], g["perm"], 128, dtype)
    assert np.array_equal(Wq.astype(np.float32), t[f"W_seq_{dtype}"])          # bit-exact dequantised weight
    x = t["x_seq"]
    a = otl.forward(x, g["alpha"], g["mu"], g["T"], g["perm"], None, 128, dtype)
    b = otl.forward_reference(x, g["alpha"], g["mu"], g["T"], g["perm"], None, 128, dtype)
    np.testing.assert_allclose(a, b, rtol=0, atol=1e-12)     # float64 BLAS summation order only
    if dtype == "float32":
        np.testing.assert_allclose(Wq.astype(np.float32), g["Wq"], rtol=0, atol=0)


def test_ssr_permutation_reference_forward_is_not_the_quantised_layer():
    """SURVEY Q11, pinned: with an SSR permutation the reference's forward differs from x @ Wq' where Wq is the
    reference's own GPTQ.get_quantized_weight(); the intended-layer oracle reproduces x @ Wq'."""
    g, t = _case("ssr")
    x = t["x_ssr"].astype(np.float64)
    want = x @ g["Wq"].astype(np.float64).T
    got = otl.forward(x, g["alpha"], g["mu"], g["T"], g["perm"], None, 128, "float32")
    assert np.abs(got - want).max() <= 1e-6 * np.abs(want).max()
    ref = t["y_ssr_float32_nobias"].astype(np.float64)
    assert np.abs(ref - want).max() > 0.05 * np.abs(want).max()


def test_pack_layer_roundtrip_and_planes():
    rng = np.random.default_rng(7)
    for n, m in ((5, 16), (3, 37), (9, 320), (4, 128)):
        T = rng.integers(-1, 2, size=(n, m)).astype(np.int8)
        perm = rng.permutation(m)
        words = otl.pack_layer(T, perm)
        assert words.shape == (n, (m + 15) // 16) and words.dtype == np.uint32
        assert np.array_equal(otl.unpack_layer(words, m, perm), T)
        assert not np.any(words & (words >> np.uint32(16)) & np.uint32(0xFFFF))     # a position is never both +1 and -1
        if m % 16:       # padding positions are T = 0 in both planes
            keep = np.uint32((1 << (m % 16)) - 1)
            assert np.all((words[:, -1] & np.uint32(0xFFFF) & ~keep) == 0)
            assert np.all(((words[:, -1] >> np.uint32(16)) & ~keep) == 0)
    T = np.array([[1, -1, 0, 1] + [0] * 12], dtype=np.int8)
    assert otl.pack_layer(T, np.arange(16))[0, 0] == (0b1001 | (0b0010 << 16))


def test_weight_table_values():
    a = np.array([[0.0123, 1.5]], dtype=np.float32)
    u = np.array([[-0.004, 0.25]], dtype=np.float32)
    for dtype in ("float32", "float16", "bfloat16"):
        w = otl.weight_table(a, u, dtype)
        ar, ur = otl.round_to(a, dtype), otl.round_to(u, dtype)
        np.testing.assert_array_equal(w[..., 1], ur.astype(np.float32))
        np.testing.assert_array_equal(w[..., 0], otl.round_to(ur - ar, dtype).astype(np.float32))
        np.testing.assert_array_equal(w[..., 2], otl.round_to(ur + ar, dtype).astype(np.float32))
    # bf16 rounding helper: round to nearest even on the 16 dropped bits
    v = np.array([1.0 + 2.0 ** -8, 1.0 + 3 * 2.0 ** -9, 1.0 + 2.0 ** -7], dtype=np.float32)
    np.testing.assert_array_equal(otl._round_bf16(v), np.array([1.0, 1.0 + 2.0 ** -7, 1.0 + 2.0 ** -7], dtype=np.float32))


def test_reference_footprint_formula():
    _, t = _case("seq")
    assert int(t["footprint_seq_float16"]) == otl.memory_footprint_reference(48, 320, 3, False)


# ---------------------------------------------------------------------- host side of the mirror
def test_ternary_linear_host_surface():
    import tq100
    layer = tq100.TernaryLinear(320, 48, block_size=128, bias=True, dtype=torch.float16)
    assert layer.codes.shape == (48, 20) and layer.codes.dtype == torch.int32
    assert layer.alpha.shape == (48, 3) and layer.mu.dtype == torch.float16
    assert layer.perm.dtype == torch.long and torch.equal(layer.inv_perm, torch.arange(320))
    assert set(layer.state_dict()) == {"codes", "alpha", "mu", "perm", "inv_perm", "bias"}
    # 2-bit codes: 48*20*4 B, fp16 alpha/mu, int32 perm, fp16 bias
    assert layer.memory_footprint() == 48 * 20 * 4 + 2 * 48 * 3 * 2 + 320 * 4 + 48 * 2
    assert layer.memory_footprint() < otl.memory_footprint_reference(48, 320, 3, True) / 2
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        layer(torch.zeros(2, 320))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        layer.set_quantized_params(torch.ones(48, 3), torch.zeros(48, 3), torch.zeros(48, 320), torch.arange(320))
    with pytest.raises(ValueError):
        tq100.TernaryLinear(320, 48, block_size=100)


def test_model_walkers():
    import tq100

    class Block(nn.Module):
        def __init__(self):
            super().__init__()
            self.attn = nn.ModuleDict({"q_proj": nn.Linear(8, 8), "o_proj": nn.Linear(8, 8)})
            self.fc = nn.Linear(8, 4)
            self.norm = nn.LayerNorm(8)

    class Inner(nn.Module):
        def __init__(self):
            super().__init__()
            self.layers = nn.ModuleList([Block(), Block()])

    class LM(nn.Module):
        def __init__(self):
            super().__init__()
            self.model = Inner()

    lm = LM()
    assert len(tq100.get_llm_layers(lm, "llama2")) == 2
    with pytest.raises(ValueError, match="Unknown model type"):
        tq100.get_llm_layers(lm, "mamba")
    with pytest.raises(AttributeError):
        tq100.get_llm_layers(lm, "opt")
    assert sorted(tq100.find_linear_layers(lm.model.layers[0])) == ["attn.o_proj", "attn.q_proj", "fc"]
    assert len(tq100.get_model_layers(lm)) == 6
    for name, want in (("meta-llama/Llama-2-7b-hf", "llama2"), ("facebook/opt-125m", "opt"), ("google/gemma-3-1b", "gemma3"),
                       ("Qwen/Qwen3-8B", "qwen3"), ("bigscience/bloom-560m", "bloom"), ("mistral", "llama")):
        assert tq100.get_model_type(name) == want
    assert tq100.compute_bits_per_weight(lm) == 16.0
    assert tq100.compute_compression_ratio(4.0, 1.0) == 4.0
    assert tq100.compute_model_size(lm) > 0


# ---------------------------------------------------------------------- checkpoint interchange with the reference
def test_reference_checkpoint_layout_loads_and_exports(tmp_path):
    """The reference's TernaryLinear keeps an int8 buffer ``T`` in original positions (model.py:43) and its
    save_quantized_model / load_quantized_model (utils.py:288-304) round-trip the state dict.  A state dict in that
    layout must load into this layer (packed on the way in, bit-identical to the oracle's pack_layer), and
    ``export_reference_T`` must write it back out."""
    import pickle
    import tq100
    from tq100 import model as tmodel
    rng = np.random.default_rng(5)
    n, m = 24, 200                                   # ragged: 200 = 12 words + 8 positions
    T = rng.integers(-1, 2, size=(n, m)).astype(np.int8)
    perm = rng.permutation(m)
    words = otl.pack_layer(T, perm)
    got = tmodel.codes_from_T(torch.from_numpy(T), torch.from_numpy(perm))
    np.testing.assert_array_equal(got.numpy().view(np.uint32), words)
    np.testing.assert_array_equal(tmodel.T_from_codes(got, torch.from_numpy(perm), m).numpy(), T)

    ref_state = {"T": torch.from_numpy(T), "alpha": torch.rand(n, 2).half(), "mu": torch.rand(n, 2).half(),
                 "perm": torch.from_numpy(perm)}                       # the reference's buffers, bias=False
    layer = tq100.TernaryLinear(m, n, block_size=128, bias=False, dtype=torch.float16)
    missing = layer.load_state_dict(ref_state, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    np.testing.assert_array_equal(layer.codes.numpy().view(np.uint32), words)
    assert torch.equal(layer.inv_perm, torch.argsort(torch.from_numpy(perm)))
    assert torch.equal(layer.alpha, ref_state["alpha"])

    # through the one-file checkpoint of utils.save_quantized_model / load_quantized_model
    holder = nn.Sequential(layer)
    path = str(tmp_path / "q.pt")
    tq100.TernaryLinear.export_reference_T = True
    try:
        tq100.save_quantized_model(holder, path, {"0": {"perm": torch.from_numpy(perm)}})
        saved = torch.load(path, map_location="cpu")["model_state_dict"]
        assert torch.equal(saved["0.T"], torch.from_numpy(T)) and saved["0.T"].dtype == torch.int8
    finally:
        tq100.TernaryLinear.export_reference_T = False
    fresh = nn.Sequential(tq100.TernaryLinear(m, n, block_size=128, bias=False, dtype=torch.float16))
    _, params = tq100.load_quantized_model(fresh, path)
    assert torch.equal(fresh[0].codes, layer.codes) and "0" in params
    assert "0.T" not in fresh.state_dict()

    # whole-module pickle (torch.save(model)) works: no lambda hooks on the layer
    clone = pickle.loads(pickle.dumps(layer))
    assert torch.equal(clone.codes, layer.codes)
