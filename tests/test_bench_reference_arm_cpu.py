"""Host logic of bench.py's reference arm (no GPU): which implementation it picks, how the steps rotate over a layer's
seven linears, and what the line says about its own extrapolation.  Run on a tiny llama-shaped configuration."""

import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

TINY = dict(name="tiny", d=256, ffn=384, layers=3)


@pytest.mark.parametrize("steps", [1, 3, 9])
def test_layer_sample_covers_every_linear_and_states_its_factors(steps, monkeypatch):
    monkeypatch.setenv("TQ_BENCH_FORCE_PORT", "1")              # the oracle's port: what the GPU box runs
    r = bench.cpu_layer_sample(TINY, "ssr", steps=steps, hess_tokens=2048, threads=2)
    assert r["kind"] == "port" and r["source"] == "oracle/torch_port.py" and r["cores"] == 2
    assert set(r["per_linear"]) == {"q_proj", "k_proj", "v_proj", "o_proj", "gate_proj", "up_proj", "down_proj"}
    assert all(v["measurements"] >= 1 for v in r["per_linear"].values())
    ex = r["extrapolation"]
    assert r["extrapolated"] is True and ex["tokens_factor"] == bench.SAMPLES * bench.SEQ / 2048 and ex["layers_factor"] == 3
    want = 3 * (ex["per_layer_add_batch_s"] + ex["per_layer_quantize_s"])
    assert abs(r["value"] - want) <= 1e-9 * want
    assert ex["steps_measured"] == steps
    if steps >= 7:                                              # one linear per step, round robin
        assert r["per_linear"]["q_proj"]["measurements"] == 2 and r["per_linear"]["v_proj"]["measurements"] == 1


def test_budget_stops_repeats_but_not_the_first_pass(monkeypatch):
    monkeypatch.setenv("TQ_BENCH_FORCE_PORT", "1")
    r = bench.cpu_layer_sample(TINY, "sequential", steps=40, budget_s=0.0, hess_tokens=2048, threads=2)
    assert r["extrapolation"]["steps_measured"] == 7            # every linear once, then the budget ends the repeats
    assert all(v["measurements"] == 1 for v in r["per_linear"].values())


@pytest.mark.skipif(not os.path.isfile("/root/reference/gptq.py"), reason="the reference exists in the build container only")
def test_the_unmodified_reference_is_used_when_present(monkeypatch):
    monkeypatch.delenv("TQ_BENCH_FORCE_PORT", raising=False)
    assert bench.find_reference_dir() is not None
    r = bench.cpu_layer_sample(TINY, "ssr", steps=1, hess_tokens=2048, threads=2)
    assert r["kind"] == "reference" and os.path.isdir(r["source"])
    # the isolated loader leaves this repo's own modules importable under their flat names
    assert "gptq" not in sys.modules or not str(getattr(sys.modules["gptq"], "__file__", "")).startswith("/root/reference")


def test_act_order_arm_uses_the_derived_oracle(monkeypatch):
    monkeypatch.delenv("TQ_BENCH_FORCE_PORT", raising=False)
    r = bench.cpu_layer_sample(dict(TINY, layers=1), "actorder", steps=1, hess_tokens=2048, threads=2)
    assert set(r["per_linear"]) and r["value"] > 0             # runs through torch_port's static_perm path on either kind
