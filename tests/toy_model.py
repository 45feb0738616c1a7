"""A small llama-shaped causal LM for the model-level driver tests: ``model.model.layers[i]`` with q/k/v/o and
gate/up/down linears (the structure ``get_llm_layers(..., 'llama')`` and ``find_linear_layers`` walk).  Weights are drawn
from a seeded CPU generator, so the golden generator (which runs the unmodified reference on it) and the tests build the
same model."""

import math

import torch
import torch.nn as nn


class ToyAttention(nn.Module):
    def __init__(self, d):
        super().__init__()
        self.q_proj = nn.Linear(d, d, bias=False)
        self.k_proj = nn.Linear(d, d, bias=False)
        self.v_proj = nn.Linear(d, d, bias=False)
        self.o_proj = nn.Linear(d, d, bias=False)

    def forward(self, x):
        q, k, v = self.q_proj(x), self.k_proj(x), self.v_proj(x)
        L = x.shape[-2]
        att = (q @ k.transpose(-1, -2)) / math.sqrt(x.shape[-1])
        mask = torch.ones(L, L, dtype=torch.bool, device=x.device).tril()
        att = att.masked_fill(~mask, float("-inf")).softmax(-1)
        return self.o_proj(att @ v)


class ToyMLP(nn.Module):
    def __init__(self, d, ffn):
        super().__init__()
        self.gate_proj = nn.Linear(d, ffn, bias=False)
        self.up_proj = nn.Linear(d, ffn, bias=False)
        self.down_proj = nn.Linear(ffn, d, bias=False)

    def forward(self, x):
        return self.down_proj(torch.nn.functional.silu(self.gate_proj(x)) * self.up_proj(x))


class ToyBlock(nn.Module):
    def __init__(self, d, ffn):
        super().__init__()
        self.input_layernorm = nn.LayerNorm(d)
        self.self_attn = ToyAttention(d)
        self.post_attention_layernorm = nn.LayerNorm(d)
        self.mlp = ToyMLP(d, ffn)

    def forward(self, x):
        x = x + self.self_attn(self.input_layernorm(x))
        return x + self.mlp(self.post_attention_layernorm(x))


class ToyInner(nn.Module):
    def __init__(self, vocab, d, ffn, layers):
        super().__init__()
        self.embed_tokens = nn.Embedding(vocab, d)
        self.layers = nn.ModuleList([ToyBlock(d, ffn) for _ in range(layers)])
        self.norm = nn.LayerNorm(d)


class ToyLM(nn.Module):
    def __init__(self, vocab=64, d=256, ffn=384, layers=2):
        super().__init__()
        self.model = ToyInner(vocab, d, ffn, layers)
        self.lm_head = nn.Linear(d, vocab, bias=False)
        self.forward_calls = 0

    def forward(self, input_ids):
        self.forward_calls += 1
        x = self.model.embed_tokens(input_ids)
        for layer in self.model.layers:
            x = layer(x)
        return self.lm_head(self.model.norm(x))


def build(seed=7, **kw):
    gen = torch.Generator().manual_seed(seed)
    model = ToyLM(**kw)
    with torch.no_grad():
        for name, p in model.named_parameters():
            if p.dim() == 2 and "embed" not in name:
                p.copy_(torch.randn(p.shape, generator=gen) * 0.05)
            elif "embed" in name:
                p.copy_(torch.randn(p.shape, generator=gen))
            elif name.endswith("weight"):
                p.fill_(1.0)
            else:
                p.zero_()
    return model.eval()


def samples(num=16, seq=64, vocab=64, seed=11):
    gen = torch.Generator().manual_seed(seed)
    return [torch.randint(0, vocab, (1, seq), generator=gen) for _ in range(num)]


class CausalLMOutput:
    def __init__(self, loss, logits):
        self.loss, self.logits = loss, logits


class ToyCausalLM(nn.Module):
    """HF-style call signature over ToyLM: model(input_ids, labels=...) -> object with .loss (labels shifted by one,
    -100 ignored, mean over the rest), which is what utils.evaluate_perplexity relies on (utils.py:176-177)."""

    def __init__(self, lm):
        super().__init__()
        self.lm = lm

    def forward(self, input_ids, labels=None):
        logits = self.lm(input_ids)
        loss = None
        if labels is not None:
            loss = torch.nn.functional.cross_entropy(logits[:, :-1].reshape(-1, logits.shape[-1]), labels[:, 1:].reshape(-1),
                                                     ignore_index=-100)
        return CausalLMOutput(loss, logits)


class CharTokenizer:
    """tokenizer(text, return_tensors='pt')['input_ids'] -> (1, len) ids: byte value modulo the vocabulary."""

    def __init__(self, vocab=64):
        self.vocab = vocab

    def __call__(self, text, return_tensors="pt"):
        return {"input_ids": torch.tensor([[b % self.vocab for b in text.encode()]], dtype=torch.long)}


def corpus(n_chars=3000, seed=5):
    gen = torch.Generator().manual_seed(seed)
    return "".join(chr(97 + int(v)) for v in torch.randint(0, 26, (n_chars,), generator=gen))
