"""calibration sampling and perplexity helpers against outputs of the unmodified reference functions
(tests/golden/utils_calibration.npz, made by tests/golden/make_golden_utils.py with an in-memory corpus).  Host-side
torch code: runs wherever the model runs, so the CPU is fine here."""

import os

import numpy as np
import pytest
import torch

import toy_model

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "utils_calibration.npz")


def _joined():
    return "\n\n".join(toy_model.corpus().split("q"))


def test_calibration_windows_match_reference_sampling():
    import tq100
    g = np.load(GOLD)
    tok = toy_model.CharTokenizer()
    got = tq100.get_calibration_data(tok, num_samples=12, seq_len=40, seed=42, text=_joined())
    assert len(got) == 12 and all(tuple(s.shape) == (1, 40) for s in got)
    assert np.array_equal(torch.cat(got, 0).numpy(), g["samples"])
    ids = tok(_joined())["input_ids"][0]
    assert len(ids) == int(g["n_tokens"])
    again = tq100.get_calibration_data(None, num_samples=12, seq_len=40, seed=42, tokens=ids)
    assert np.array_equal(torch.cat(again, 0).numpy(), g["samples"])
    with pytest.raises(ValueError, match="seq_len"):
        tq100.get_calibration_data(None, num_samples=1, seq_len=40, tokens=ids[:41])


def test_perplexity_matches_reference_windows():
    import tq100
    g = np.load(GOLD)
    tok = toy_model.CharTokenizer()
    model = toy_model.ToyCausalLM(toy_model.build())
    for seq_len, want in zip(g["ppl_seq"], g["ppl"]):
        got = tq100.evaluate_perplexity(model, tok, seq_len=int(seq_len), text=_joined())
        assert abs(got - want) <= 1e-5 * want, (seq_len, got, want)
    ids = tok(_joined())["input_ids"]
    assert abs(tq100.evaluate_perplexity(model, None, seq_len=50, input_ids=ids) - g["ppl"][0]) <= 1e-5 * g["ppl"][0]


def test_prepare_calibration_inputs_collects_every_linear():
    import tq100
    lm = toy_model.build()
    acts = tq100.prepare_calibration_inputs(lm, toy_model.samples(num=3, seq=8), torch.device("cpu"))
    assert len(acts) == 2 * 7 + 1 and tuple(acts["model.layers.0.self_attn.q_proj"].shape) == (3, 8, 256)
    assert tuple(acts["model.layers.1.mlp.down_proj"].shape) == (3, 8, 384)
    assert torch.equal(acts["model.layers.0.self_attn.q_proj"], acts["model.layers.0.self_attn.v_proj"])
