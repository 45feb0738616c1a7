"""The model-level driver on a real Hugging Face decoder (tiny random LlamaForCausalLM, CPU): the captured layer inputs
must not carry a live KV cache.  Every transformer layer is run twice on the captured kwargs (original weights, then
quantised weights); with ``use_cache=True`` the second pass would attend to the first pass's keys and values and every
layer after the first would be calibrated on wrong activations -- the reference runs the whole model with a fresh cache
each time (main.py:278-282), so it has no such problem.  The toy model of the other driver tests takes no kwargs and
cannot see this."""

import pytest
import torch

transformers = pytest.importorskip("transformers")


def _tiny_llama():
    cfg = transformers.LlamaConfig(vocab_size=97, hidden_size=64, intermediate_size=96, num_hidden_layers=2,
                                   num_attention_heads=4, num_key_value_heads=4, max_position_embeddings=64,
                                   use_cache=True)
    torch.manual_seed(3)
    return transformers.LlamaForCausalLM(cfg).eval()


class _Passthrough:
    """A quantiser that changes nothing: the 'quantised' model equals the original, so the driver's layer-by-layer
    advance must reproduce a plain whole-model forward exactly."""

    def __init__(self, layer, hessian):
        self.layer, self.state = layer, hessian if hessian is not None else object()
        self.rows, self.columns = layer.weight.shape
        self.device = layer.weight.device
        self.calls = 0

    def add_batch(self, x):
        self.calls += 1

    def run(self):
        n, m = self.rows, self.columns
        self.alpha = self.mu = torch.zeros(n, (m + 127) // 128)
        self.T_int8, self.perm = torch.zeros(n, m, dtype=torch.int8), torch.arange(m)

    def get_quantized_weight(self):
        return self.layer.weight.detach().clone()


def test_layer_inputs_carry_no_kv_cache_and_advance_matches_whole_model():
    import tq100
    model = _tiny_llama()
    samples = [torch.randint(0, 97, (1, 16), generator=torch.Generator().manual_seed(i)) for i in range(3)]
    with torch.no_grad():
        want = [model(s, use_cache=False, output_hidden_states=True).hidden_states for s in samples]
    pq = tq100.PT2LLMQuantizer(model, None, model_type="llama", device="cpu", gptq_factory=_Passthrough,
                               chain_runner=lambda gs, ssr: [g.run() for g in gs])
    layers = tq100.get_llm_layers(model, "llama")
    captured = pq._capture_first_layer_inputs(layers[0], samples)
    for args, kwargs in captured:
        for key in ("past_key_value", "past_key_values", "layer_past"):
            assert kwargs.get(key) is None, key
        assert not kwargs.get("use_cache", False)
    # two passes per layer over the same kwargs, as quantize() does: the second must not see the first
    with torch.no_grad():
        for li, layer in enumerate(layers):
            first = [pq._layer_forward(layer, a, k) for a, k in captured]
            captured = [pq._advance(layer, a, k) for a, k in captured]
            for si, ((a, _), f) in enumerate(zip(captured, first)):
                assert torch.equal(a[0], f), (li, si)
                # (transformers appends the last hidden state after the final norm)
                got = model.model.norm(a[0]) if li + 1 == len(layers) else a[0]
                torch.testing.assert_close(got, want[si][li + 1], rtol=1e-5, atol=1e-6)
    # and the whole driver runs on it
    params = pq.quantize(samples)
    assert len(params) == 2 * 7
