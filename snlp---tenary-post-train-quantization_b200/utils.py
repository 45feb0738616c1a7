"""2-bit ternary codec -- B200 mirror of the reference's ``utils.pack_ternary`` /
``utils.unpack_ternary`` (``/root/reference/utils.py:189-248``): code = T+1 in {0,1,2}, four codes
per byte (c0 | c1<<2 | c2<<4 | c3<<6) over the flat row-major tensor, zero padded."""

from typing import Tuple

import torch

try:
    from . import _lib
except ImportError:
    import _lib


def pack_ternary(T: torch.Tensor):
    """utils.py:189-219.  Returns (packed uint8[ceil(N/4)], orig_shape)."""
    lib = _lib.load()
    _lib.require_cuda(T, "T")
    orig_shape = T.shape
    flat = T.detach().contiguous().reshape(-1)
    count = flat.numel()
    packed = torch.empty((count + 3) // 4, dtype=torch.uint8, device=T.device)
    with torch.cuda.device(T.device):
        if flat.dtype == torch.int8:
            _lib.check(lib.tq_pack2b(_lib.ptr(flat), count, _lib.ptr(packed), _lib.stream()), "tq_pack2b")
        else:
            flat = flat.float() if flat.dtype != torch.float32 else flat
            _lib.check(lib.tq_pack2b_f32(_lib.ptr(flat), count, _lib.ptr(packed), _lib.stream()), "tq_pack2b_f32")
    return packed, orig_shape


def unpack_ternary(packed: torch.Tensor, orig_shape: Tuple[int, ...]) -> torch.Tensor:
    """utils.py:222-248.  int8 output in {-1, 0, 1}."""
    lib = _lib.load()
    _lib.require_cuda(packed, "packed")
    total = 1
    for d in orig_shape:
        total *= d
    if packed.numel() * 4 < total:
        raise ValueError(f"packed buffer holds {packed.numel() * 4} codes, shape {tuple(orig_shape)} needs {total}")
    out = torch.empty(total, dtype=torch.int8, device=packed.device)
    p = packed.detach().contiguous()
    with torch.cuda.device(packed.device):
        _lib.check(lib.tq_unpack2b(_lib.ptr(p), total, _lib.ptr(out), _lib.stream()), "tq_unpack2b")
    return out.reshape(orig_shape)


# ---------------------------------------------------------------------- model-level helpers (host only)
def set_seed(seed: int):
    """utils.py:15-21."""
    import random
    import numpy as np
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)


def compute_bits_per_weight(model: torch.nn.Module, include_scales: bool = True) -> float:
    """utils.py:251-285: 1.58 bits per ternary weight + 16 bits per alpha/mu entry over every ternary layer;
    16.0 for a model without one.  (The codes are stored at 2 bits per weight; ``stored_bits_per_weight`` reports
    what the buffers really occupy.)"""
    weights = 0
    bits = 0.0
    for module in model.modules():
        if hasattr(module, "codes") and hasattr(module, "alpha") and hasattr(module, "mu"):
            count = module.out_features * module.in_features
            weights += count
            bits += count * 1.58
            if include_scales:
                bits += (module.alpha.numel() + module.mu.numel()) * 16
    return bits / weights if weights else 16.0


def stored_bits_per_weight(model: torch.nn.Module) -> float:
    """Resident bits per weight of the ternary layers (2-bit codes, scales, int32 permutation, bias)."""
    weights = 0
    nbytes = 0
    for module in model.modules():
        if hasattr(module, "codes") and hasattr(module, "memory_footprint"):
            weights += module.out_features * module.in_features
            nbytes += module.memory_footprint()
    return 8.0 * nbytes / weights if weights else 16.0


def save_quantized_model(model: torch.nn.Module, save_path: str, quantized_params):
    """utils.py:288-296: one file holding the state dict and the per-layer parameter dicts."""
    torch.save({"model_state_dict": model.state_dict(), "quantized_params": quantized_params}, save_path)


def load_quantized_model(model: torch.nn.Module, load_path: str):
    """utils.py:299-304."""
    checkpoint = torch.load(load_path, map_location="cpu")
    model.load_state_dict(checkpoint["model_state_dict"])
    return model, checkpoint.get("quantized_params", {})


# ---------------------------------------------------------------------- calibration / evaluation plumbing
def _load_text(dataset_name: str, dataset_config: str, split_kind: str, num_samples: int = 128) -> str:
    """The text the reference concatenates (utils.py:45-62 for calibration, :156-164 for evaluation).  Needs the
    `datasets` package and its cache / the network; callers without either pass `text=` or `tokens=` instead."""
    try:
        from datasets import load_dataset
    except ImportError as exc:                                          # pragma: no cover
        raise RuntimeError("the `datasets` package is not installed: pass text= or tokens=") from exc
    train = split_kind == "train"
    if dataset_name == "wikitext":
        return "\n\n".join(load_dataset(dataset_name, dataset_config, split="train" if train else "test")["text"])
    if dataset_name == "c4":
        stream = load_dataset("allenai/c4", "en", split="train" if train else "validation", streaming=True)
        return "\n\n".join(item["text"] for item in stream.take(num_samples * 10 if train else 1000))
    if dataset_name == "ptb" and train:
        return "\n\n".join(load_dataset("ptb_text_only", "penn_treebank", split="train")["sentence"])
    raise ValueError(f"Unknown dataset: {dataset_name}")


def get_calibration_data(tokenizer, dataset_name: str = "wikitext", dataset_config: str = "wikitext-2-raw-v1",
                         num_samples: int = 128, seq_len: int = 2048, seed: int = 42, text: str = None,
                         tokens: torch.Tensor = None):
    """utils.py:24-72: `num_samples` windows of `seq_len` tokens at seeded random offsets of the tokenised corpus, each
    (1, seq_len).  `text` / `tokens` (1-D ids) replace the dataset download; the sampling is the reference's
    (set_seed, then random.randint(0, len - seq_len - 1) per sample)."""
    import random
    set_seed(seed)
    if tokens is None:
        if text is None:
            text = _load_text(dataset_name, dataset_config, "train", num_samples)
        tokens = tokenizer(text, return_tensors="pt")["input_ids"][0]
    if len(tokens) < seq_len + 2:
        raise ValueError(f"corpus has {len(tokens)} tokens, need more than seq_len + 1 = {seq_len + 1}")
    samples = []
    for _ in range(num_samples):
        start = random.randint(0, len(tokens) - seq_len - 1)
        samples.append(tokens[start:start + seq_len].unsqueeze(0))
    return samples


def prepare_calibration_inputs(model: torch.nn.Module, samples, device):
    """utils.py:75-124: inputs of every nn.Linear over the samples, concatenated on the CPU.  Kept for API parity; the
    quantisation path streams activations into GPTQ.add_batch instead of storing them (main.PT2LLMQuantizer)."""
    activations, hooks = {}, []

    def make_hook(name):
        def hook(module, inp, out):
            x = inp[0] if isinstance(inp, tuple) else inp
            activations.setdefault(name, []).append(x.detach().cpu())
        return hook

    for name, module in model.named_modules():
        if isinstance(module, torch.nn.Linear):
            hooks.append(module.register_forward_hook(make_hook(name)))
    model.eval()
    try:
        with torch.no_grad():
            for sample in samples:
                model(sample.to(device))
    finally:
        for h in hooks:
            h.remove()
    return {name: torch.cat(parts, dim=0) for name, parts in activations.items()}


@torch.no_grad()
def evaluate_perplexity(model: torch.nn.Module, tokenizer, dataset_name: str = "wikitext",
                        dataset_config: str = "wikitext-2-raw-v1", seq_len: int = 2048, device=None, text: str = None,
                        input_ids: torch.Tensor = None) -> float:
    """utils.py:127-186: exp(mean NLL) over consecutive windows of `seq_len` tokens; the model is called as
    model(input_chunk, labels=target_ids) and must return an object with `.loss` (HF causal LM).  `text` /
    `input_ids` ((1, N) ids) replace the dataset download."""
    if device is None:
        device = next(model.parameters()).device
    if input_ids is None:
        if text is None:
            text = _load_text(dataset_name, dataset_config, "eval")
        input_ids = tokenizer(text, return_tensors="pt")["input_ids"]
    input_ids = input_ids.to(device)
    total = input_ids.size(1)
    seq_len = min(seq_len, total)
    nlls, prev_end = [], 0
    for begin in range(0, total, seq_len):
        end = min(begin + seq_len, total)
        trg_len = end - prev_end
        chunk = input_ids[:, begin:end]
        target = chunk.clone()
        target[:, :-trg_len] = -100
        nlls.append(model(chunk, labels=target).loss * trg_len)
        prev_end = end
        if end >= total:
            break
    return torch.exp(torch.stack(nlls).sum() / prev_end).item()
