"""SURVEY 8f N4 on the device the hot path runs on: the calibration / perplexity helpers driving a CUDA model, pinned to the
same outputs of the unmodified reference functions as the CPU test (tests/golden/utils_calibration.npz), and the whole
flow they exist for -- sample windows -> PT2LLMQuantizer.quantize (CUDA kernels) -> TernaryLinear layers -> perplexity of
the quantised model -- end to end on the GPU."""

import os

import numpy as np
import pytest
import torch

import toy_model

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "utils_calibration.npz")


def _joined():
    return "\n\n".join(toy_model.corpus().split("q"))


def test_perplexity_on_cuda_matches_reference_windows():
    import tq100
    g = np.load(GOLD)
    tok = toy_model.CharTokenizer()
    model = toy_model.ToyCausalLM(toy_model.build().to(DEV))
    for seq_len, want in zip(g["ppl_seq"], g["ppl"]):
        got = tq100.evaluate_perplexity(model, tok, seq_len=int(seq_len), text=_joined(), device=DEV)
        assert abs(got - want) <= 2e-4 * want, (seq_len, got, want)          # fp32 on another device: looser than the CPU test


def test_prepare_calibration_inputs_on_cuda():
    import tq100
    lm = toy_model.build().to(DEV)
    acts = tq100.prepare_calibration_inputs(lm, toy_model.samples(num=3, seq=8), DEV)
    assert len(acts) == 2 * 7 + 1 and tuple(acts["model.layers.0.self_attn.q_proj"].shape) == (3, 8, 256)
    ref = tq100.prepare_calibration_inputs(toy_model.build(), toy_model.samples(num=3, seq=8), torch.device("cpu"))
    for k in ref:
        assert torch.allclose(acts[k].cpu(), ref[k], rtol=1e-4, atol=1e-5), k


def test_quantise_then_perplexity_end_to_end_on_cuda():
    import tq100
    tok = toy_model.CharTokenizer()
    lm = toy_model.build().to(DEV)
    ppl_fp = tq100.evaluate_perplexity(toy_model.ToyCausalLM(lm), tok, seq_len=50, text=_joined(), device=DEV)
    # calibration windows: the toy model's seeded token samples (the same ones the reference's own quantize() run of
    # tests/golden/make_golden_model.py saw; windows of the repetitive toy corpus drive the reference's AGA into its
    # clamped-denominator blow-up, SURVEY Q9, which is not what this test is about)
    samples = toy_model.samples()
    before = tq100._lib.launch_count()
    pq = tq100.PT2LLMQuantizer(lm, tok, model_type="llama", use_ssr=False, device=DEV)
    params = pq.quantize(samples)
    assert tq100._lib.launch_count() > before and len(params) == 14
    ppl_q = tq100.evaluate_perplexity(toy_model.ToyCausalLM(lm), tok, seq_len=50, text=_joined(), device=DEV)
    assert np.isfinite(ppl_q) and ppl_q >= 0.9 * ppl_fp          # ternary weights cannot beat the fp32 model by a margin
    assert ppl_q < 50 * ppl_fp                                   # ... and the model still works
