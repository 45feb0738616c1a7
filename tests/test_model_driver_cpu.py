"""SURVEY 8f N1 on CPU: the host logic of the model-level driver (tq100.PT2LLMQuantizer) against outputs of the
unmodified reference ``PT2LLMQuantizer.quantize()`` on the seeded toy model (tests/golden/model_toy_*.npz, made by
tests/golden/make_golden_model.py).

The driver's two seams (how a per-linear quantiser is made, how a layer's quantisers are run) are filled with the numpy
ORACLE here -- a test double for the CUDA path, legal only in tests -- so everything else is the product's own code:
capture of the first layer's inputs, one accumulate forward + one advance forward per layer, streaming hooks, shared
Hessians for linears fed the same tensor, parameter naming, weight overwrite.  The GPU twin of this test
(tests/test_gpu_whole_model.py) runs the same comparison with the real kernels."""

import os

import numpy as np
import pytest
import torch

import oracle
import parity
import toy_model

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


class _State:
    def __init__(self, m):
        self.H = np.zeros((m, m), dtype=np.float32)
        self.nsamples = 0
        self.add_calls = 0


class OracleGPTQ:
    """Stands where gptq.GPTQ stands in the driver: add_batch accumulates X'X (gptq.py:59-76), run() is the oracle's
    restatement of main.py:102-230."""
    made = []

    def __init__(self, layer, hessian):
        self.layer = layer
        self.rows, self.columns = layer.weight.shape
        self.device = layer.weight.device
        self.state = hessian if hessian is not None else _State(self.columns)
        self.shared = hessian is not None
        OracleGPTQ.made.append(self)

    def add_batch(self, x):
        self.state.H, self.state.nsamples = oracle.hessian_add_batch(self.state.H, self.state.nsamples, x.float().numpy())
        self.state.add_calls += 1

    def run(self, use_ssr):
        W = self.layer.weight.detach().float().numpy()
        a, u, T, perm = oracle.quantize_layer(W, self.state.H, self.state.nsamples, 128, 0.01,
                                              "ssr" if use_ssr else "sequential", aga="activations")
        self.alpha, self.mu = torch.from_numpy(a), torch.from_numpy(u)
        self.T_int8, self.perm = torch.from_numpy(T.astype(np.int8)), torch.from_numpy(perm)

    def get_quantized_weight(self):
        return torch.from_numpy(oracle.get_quantized_weight(self.alpha.numpy(), self.mu.numpy(),
                                                            self.T_int8.numpy().astype(np.float32), self.perm.numpy()))


def _run(use_ssr, share_inputs=True):
    import tq100
    OracleGPTQ.made = []
    model = toy_model.build()
    pq = tq100.PT2LLMQuantizer(model, None, model_type="llama", use_ssr=use_ssr, device="cpu", share_inputs=share_inputs,
                               gptq_factory=OracleGPTQ, chain_runner=lambda gs, ssr: [g.run(ssr) for g in gs])
    params = pq.quantize(toy_model.samples())
    return model, pq, params


@pytest.mark.parametrize("use_ssr", [False, True])
def test_driver_matches_reference_quantize(use_ssr):
    gold = np.load(os.path.join(GOLD, f"model_toy_{'ssr' if use_ssr else 'seq'}.npz"))
    model, pq, params = _run(use_ssr)
    assert sorted(params) == sorted({f"layer_{i}.{n}" for i in range(2) for n in
                                     ("self_attn.q_proj", "self_attn.k_proj", "self_attn.v_proj", "self_attn.o_proj",
                                      "mlp.gate_proj", "mlp.up_proj", "mlp.down_proj")})
    compared = 0
    for name, p in params.items():
        if f"{name}/T" not in gold.files:
            continue                      # SSR: only layer 0 of the reference run is meaningful
        assert p["T"].dtype == torch.int8 and p["perm"].dtype == torch.int64 and p["alpha"].dtype == torch.float32
        got = {k: v.numpy() for k, v in p.items()}
        ref = {k: gold[f"{name}/{k}"] for k in ("alpha", "mu", "T", "perm")}
        parity.assert_model_level_parity(name, got, ref)
        compared += 1
    assert compared == (7 if use_ssr else 14)
    # layer forwards: one accumulate + one advance per layer and sample, except no advance after the last layer;
    # model forwards: one (stopped at layer 0) per sample -- the reference ran 2 * 16 full forwards (recorded before the logits forward)
    assert pq.layer_forwards == 16 * (2 * 2 - 1)
    assert model.forward_calls == 16 and int(gold["forward_calls"]) == 2 * 16
    if not use_ssr:
        # identity permutation: the reference's overwrite is the correct dequantisation, so the quantised models agree
        W0 = model.model.layers[0].self_attn.q_proj.weight.detach().numpy()
        Wr = gold["layer_0.self_attn.q_proj/W_after"]
        assert (np.abs(W0 - Wr) > 1e-4 * np.abs(Wr).max()).mean() <= 1e-3   # a flipped code moves its weight by alpha
        logits = model(toy_model.samples()[0]).detach().numpy()
        ref = gold["logits_after"]
        # layer 1's ~0.3 % flipped codes (the reference's own run-to-run floor, see parity.assert_model_level_parity) move the
        # logits by a few per cent of their norm
        rel = np.linalg.norm(logits - ref) / np.linalg.norm(ref)
        assert rel <= 0.1, rel           # the reference against itself (8 vs 1 MKL threads): 0.028; quantisation itself: 0.32


def test_shared_hessians_and_streaming():
    _, _, _ = _run(False, share_inputs=True)
    by_layer = OracleGPTQ.made[:7]
    names = ["q", "k", "v", "o", "gate", "up", "down"]
    st = dict(zip(names, by_layer))
    assert st["k"].state is st["q"].state and st["v"].state is st["q"].state        # same tensor -> one Hessian
    assert st["up"].state is st["gate"].state
    assert st["o"].state is not st["q"].state and st["down"].state is not st["gate"].state
    assert st["q"].state.add_calls == 16 and st["q"].state.nsamples == 16 * 64      # streamed per sample, counted once
    _, _, _ = _run(False, share_inputs=False)
    assert len({id(g.state) for g in OracleGPTQ.made[:7]}) == 7


def test_shared_and_unshared_runs_agree():
    _, _, a = _run(True, share_inputs=True)
    _, _, b = _run(True, share_inputs=False)
    for name in a:
        for k in ("T", "perm"):
            assert torch.equal(a[name][k], b[name][k]), (name, k)
        for k in ("alpha", "mu"):
            # equal_nan: one row of layer_1.mlp.down_proj hits the reference's AGA denominator blow-up (SURVEY Q9,
            # quantizer.py:239-240 reproduced literally) and carries inf/nan scales on both sides
            assert torch.allclose(a[name][k], b[name][k], rtol=0, atol=0, equal_nan=True), (name, k)


def test_dequantize_weight_and_calibration_stub():
    import tq100
    gold = np.load(os.path.join(GOLD, "model_toy_ssr.npz"))
    pq = tq100.PT2LLMQuantizer(toy_model.build(), None, device="cpu")
    name = "layer_0.mlp.down_proj"
    p = {k: torch.from_numpy(gold[f"{name}/{k}"]) for k in ("alpha", "mu", "T", "perm")}
    W = pq._dequantize_weight(p).numpy()
    want = oracle.get_quantized_weight(gold[f"{name}/alpha"], gold[f"{name}/mu"], gold[f"{name}/T"].astype(np.float32),
                                       gold[f"{name}/perm"])
    np.testing.assert_array_equal(W, want)
    with pytest.raises(RuntimeError, match="network"):
        pq.quantize()


def test_replace_with_ternary_routes_each_layer_its_own_parameters(monkeypatch):
    """Host routing only (the layer swap itself runs on the GPU: tests/test_gpu_ternary_linear.py)."""
    import tq100
    calls = []
    monkeypatch.setattr(tq100.main, "_replace_linear_with_ternary",
                        lambda module, params, block_size: calls.append((module, params, block_size)))
    model = toy_model.build()
    pq = tq100.PT2LLMQuantizer(model, None, device="cpu")
    z = torch.zeros(1)
    pq.quantized_params = {"layer_0.self_attn.q_proj": {"alpha": z, "mu": z, "T": z, "perm": z},
                           "layer_1.mlp.down_proj": {"alpha": z, "mu": z, "T": z, "perm": z},
                           "layer_1.mlp.up_proj": {"alpha": z, "mu": z, "T": z, "perm": z}}
    assert pq.replace_with_ternary(dtype=torch.float16) is model
    assert [c[0] for c in calls] == [model.model.layers[0], model.model.layers[1]] and all(c[2] == 128 for c in calls)
    assert sorted(calls[1][1]) == ["mlp.down_proj", "mlp.up_proj"] and sorted(calls[0][1]) == ["self_attn.q_proj"]
    assert calls[0][1]["self_attn.q_proj"]["alpha"].dtype == torch.float16
    assert pq.quantized_params["layer_0.self_attn.q_proj"]["alpha"].dtype == torch.float32       # stored copy untouched

