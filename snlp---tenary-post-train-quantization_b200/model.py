"""Ternary inference layer on 2-bit codes -- B200 mirror of the reference's ``model.py`` layer wrappers
(``/root/reference/model.py``: ``TernaryLinear`` :17-127, ``get_model_layers`` :130-136, ``get_llm_layers``
:139-159, ``find_linear_layers`` :162-171, ``replace_linear_with_ternary`` :174-225, ``get_model_type`` :268-290,
``compute_model_size`` :293-303).  SURVEY 8f N2: the consumer of the packed codes.

``TernaryLinear`` keeps the reference's constructor, ``set_quantized_params(alpha, mu, T, perm, bias)``,
``forward``, ``_dequantize`` and ``memory_footprint``, but

  * stores the codes as 2 bits per weight (a +1 and a -1 bit plane per 16 positions) in sweep order (``codes``, the TL2
    layout of include/tq100.h) instead of
    an int8 matrix (1 byte per weight, model.py:43); ``T`` is a read-only property that unpacks them;
  * computes ``F.linear(x, Wq)`` with ``Wq = GPTQ.get_quantized_weight()`` (gptq.py:201-230).  The reference's
    forward (model.py:84-90) gathers the input by ``perm`` AND un-permutes a weight whose T is already stored in
    original positions (gptq.py:155), which is only right for the identity permutation (pinned in
    tests/test_ternary_linear_cpu.py against the reference's own outputs).  For ``use_ssr=False`` results the two
    agree to rounding.

Decode-sized inputs (<= ``gemv_max_tokens`` rows) go through ``tq_tl_gemv`` (reads 0.25 B per weight).  Up to
``fused_max_tokens`` rows, fp16/bf16 layers run ``tq_tl_gemm_tc`` -- one tcgen05 GEMM that expands the codes to the
layer's 16-bit dtype in shared memory, no dense weight in HBM; beyond that (and for fp32 layers) the layer
dequantises to a dense weight in its dtype (``tq_tl_dequant``) and calls the library GEMM, like the reference's
``_dequantize`` + ``F.linear``.  CUDA only: no CPU fallback.

HF model loading (``load_model_for_quantization``, model.py:228-265) is out of scope (no network).
"""

from typing import Dict, List, Optional

import torch
import torch.nn as nn

try:
    from . import _lib
except ImportError:
    import _lib


def codes_from_T(T: torch.Tensor, perm: torch.Tensor) -> torch.Tensor:
    """Checkpoint interchange (host glue, any device): the reference's int8 ``T`` buffer (model.py:43; original column
    positions) -> the TL2 code words of include/tq100.h: word w of a row holds sweep positions 16w..16w+15, bit j set
    where T = +1, bit 16+j where T = -1.  Bit-identical to ``tq_tl_pack`` (tests/test_gpu_ternary_linear.py)."""
    n, m = T.shape
    Ts = T.to(torch.int64).index_select(1, perm.to(device=T.device, dtype=torch.long))
    pad = (-m) % 16
    if pad:
        Ts = torch.nn.functional.pad(Ts, (0, pad))
    Ts = Ts.view(n, -1, 16)
    w = torch.ones(16, dtype=torch.int64, device=T.device) << torch.arange(16, device=T.device)
    word = ((Ts == 1).to(torch.int64) * w).sum(-1) | (((Ts == -1).to(torch.int64) * w).sum(-1) << 16)
    return torch.where(word >= 2 ** 31, word - 2 ** 32, word).to(torch.int32)


def T_from_codes(codes: torch.Tensor, perm: torch.Tensor, m: int) -> torch.Tensor:
    """Inverse of codes_from_T: int8 (n, m) in original column positions."""
    n = codes.shape[0]
    word = codes.to(torch.int64) & 0xFFFFFFFF
    j = torch.arange(16, device=codes.device)
    plus = (word.unsqueeze(-1) >> j) & 1
    minus = (word.unsqueeze(-1) >> (j + 16)) & 1
    Ts = (plus - minus).view(n, -1)[:, :m].to(torch.int8)
    T = torch.empty_like(Ts)
    T[:, perm.to(device=codes.device, dtype=torch.long)] = Ts
    return T


def _after_load(module, incompatible):
    module._invalidate()


class TernaryLinear(nn.Module):
    """model.py:17-127 on packed codes.

    Checkpoints: the state dict holds ``codes`` (2-bit planes) where the reference's holds an int8 ``T``
    (model.py:43).  Loading accepts either: a reference-produced ``<prefix>T`` is packed on the way in
    (``_load_from_state_dict``).  ``TernaryLinear.export_reference_T = True`` makes ``state_dict()`` also carry
    ``<prefix>T`` so the reference's ``load_quantized_model`` (utils.py:299-304, ``strict=False`` for the extra
    ``codes`` / ``inv_perm`` keys) can read checkpoints written here."""

    export_reference_T = False

    gemv_max_tokens = 16
    # many-token path of fp16/bf16 layers: tq_tl_gemm_tc up to fused_max_tokens rows, dense weight + library GEMM beyond
    # (measured crossover on B200, profiles/r01d_tl_bench_prefill.json); fused_gemm = False forces the dense path
    fused_gemm = True
    fused_max_tokens = 1024

    def __init__(self, in_features: int, out_features: int, block_size: int = 128, bias: bool = True,
                 dtype: torch.dtype = torch.float16, device=None):
        super().__init__()
        if dtype not in (torch.float32, torch.float16, torch.bfloat16):
            raise ValueError(f"unsupported layer dtype {dtype}")
        if block_size <= 0 or block_size % 16 != 0:
            raise ValueError("TernaryLinear: block_size must be a positive multiple of 16 (a 32-bit code word must "
                             "not straddle two scale blocks)")
        self.in_features = in_features
        self.out_features = out_features
        self.block_size = block_size
        num_blocks = (in_features + block_size - 1) // block_size
        wpr = (in_features + 15) // 16
        # both bit planes empty = T 0 everywhere, like the reference's zero-initialised T (model.py:43)
        self.register_buffer("codes", torch.zeros((out_features, wpr), dtype=torch.int32, device=device))
        self.register_buffer("alpha", torch.ones(out_features, num_blocks, dtype=dtype, device=device))
        self.register_buffer("mu", torch.zeros(out_features, num_blocks, dtype=dtype, device=device))
        self.register_buffer("perm", torch.arange(in_features, dtype=torch.long, device=device))
        self.register_buffer("inv_perm", torch.arange(in_features, dtype=torch.long, device=device))
        if bias:
            self.register_buffer("bias", torch.zeros(out_features, dtype=dtype, device=device))
        else:
            self.bias = None
        self._derived = None          # (wtab f32 [n, nb, 4], perm / inv_perm int32 or None, bias f32 or None), rebuilt lazily
        self.register_load_state_dict_post_hook(_after_load)      # a module-level function: the layer stays picklable
        self._register_state_dict_hook(TernaryLinear._export_T)

    @staticmethod
    def _export_T(module, state_dict, prefix, local_metadata):
        if module.export_reference_T:
            state_dict[prefix + "T"] = T_from_codes(module.codes, module.perm, module.in_features)
        return state_dict

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs):
        """Accept the reference's checkpoint layout (int8 ``T`` in original positions + ``perm``, model.py:43-47) next to
        this layer's own (``codes``): T is packed into code words before the regular loading runs."""
        tkey, ckey = prefix + "T", prefix + "codes"
        if tkey in state_dict:
            T = state_dict.pop(tkey)
            if ckey not in state_dict:
                perm = state_dict.get(prefix + "perm", self.perm)
                if tuple(T.shape) != (self.out_features, self.in_features):
                    error_msgs.append(f"size mismatch for {tkey}: {tuple(T.shape)} vs {(self.out_features, self.in_features)}")
                else:
                    state_dict[ckey] = codes_from_T(T, perm)
                    if prefix + "inv_perm" not in state_dict:
                        state_dict[prefix + "inv_perm"] = torch.argsort(perm.long())
        super()._load_from_state_dict(state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs)

    # ------------------------------------------------------------------ parameters
    def _invalidate(self):
        self._derived = None

    def _apply(self, fn, *args, **kwargs):
        self._invalidate()
        return super()._apply(fn, *args, **kwargs)

    def set_quantized_params(self, alpha: torch.Tensor, mu: torch.Tensor, T: torch.Tensor, perm: torch.Tensor,
                             bias: Optional[torch.Tensor] = None):
        """model.py:58-73.  T is the (n, m) ternary matrix in ORIGINAL column positions as GPTQ.quantize returns it
        (gptq.py:155), any dtype; it is packed on the device the layer lives on."""
        lib = _lib.load()
        dev = self.codes.device
        _lib.require_cuda(self.codes, "TernaryLinear buffers")
        n, m = self.out_features, self.in_features
        if tuple(T.shape) != (n, m) or perm.numel() != m:
            raise ValueError(f"expected T {(n, m)} and perm ({m},), got {tuple(T.shape)} and {tuple(perm.shape)}")
        self.alpha.copy_(alpha)
        self.mu.copy_(mu)
        self.perm.copy_(perm)
        self.inv_perm.copy_(torch.argsort(self.perm))
        T8 = T.detach().to(device=dev).to(torch.int8).contiguous()
        p32 = self.perm.to(torch.int32).contiguous()
        with torch.cuda.device(dev):
            _lib.check(lib.tq_tl_pack(_lib.ptr(T8), n, m, _lib.ptr(p32), _lib.ptr(self.codes), self.codes.shape[1],
                                      _lib.stream()), "tq_tl_pack")
        if bias is not None and self.bias is not None:
            self.bias.copy_(bias)
        self._invalidate()
        self._prepared()              # build the weight table now: the first forward is then free of host syncs (graph capture)

    def _prepared(self):
        if self._derived is not None:
            return self._derived
        lib = _lib.load()
        _lib.require_cuda(self.codes, "TernaryLinear buffers")
        dev = self.codes.device
        n, nb = self.alpha.shape
        wtab = torch.empty((n, nb, 4), dtype=torch.float32, device=dev)
        a32 = self.alpha.float().contiguous()
        u32 = self.mu.float().contiguous()
        with torch.cuda.device(dev):
            _lib.check(lib.tq_tl_wtab(_lib.ptr(a32), _lib.ptr(u32), n, nb, _lib.dtype_code(self.alpha.dtype),
                                      _lib.ptr(wtab), _lib.stream()), "tq_tl_wtab")
        identity = bool(torch.equal(self.perm, torch.arange(self.in_features, device=dev)))
        perm32 = None if identity else self.perm.to(torch.int32).contiguous()
        inv32 = None if identity else torch.argsort(self.perm).to(torch.int32).contiguous()
        bias32 = None if self.bias is None else self.bias.float().contiguous()
        self._derived = (wtab, perm32, bias32, inv32)
        return self._derived

    @property
    def T(self) -> torch.Tensor:
        """int8 (n, m) ternary matrix in original column positions (the reference's buffer, model.py:43)."""
        lib = _lib.load()
        _lib.require_cuda(self.codes, "TernaryLinear buffers")
        _, perm32, _, inv32 = self._prepared()
        n, m = self.out_features, self.in_features
        out = torch.empty((n, m), dtype=torch.int8, device=self.codes.device)
        with torch.cuda.device(self.codes.device):
            _lib.check(lib.tq_tl_unpack(_lib.ptr(self.codes), self.codes.shape[1], n, m, _lib.ptr(perm32), _lib.ptr(inv32),
                                        _lib.ptr(out), _lib.stream()), "tq_tl_unpack")
        return out

    # ------------------------------------------------------------------ forward
    def _dequantize(self) -> torch.Tensor:
        """Dense weight (n, m) in the layer dtype, original column positions (gptq.py:201-230 semantics)."""
        lib = _lib.load()
        wtab, perm32, _, inv32 = self._prepared()
        n, m = self.out_features, self.in_features
        W = torch.empty((n, m), dtype=self.alpha.dtype, device=self.codes.device)
        with torch.cuda.device(self.codes.device):
            _lib.check(lib.tq_tl_dequant(_lib.ptr(self.codes), self.codes.shape[1], _lib.ptr(wtab), n, m, self.block_size,
                                         _lib.ptr(perm32), _lib.ptr(inv32), _lib.ptr(W), _lib.dtype_code(W.dtype), m,
                                         _lib.stream()), "tq_tl_dequant")
        return W

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """model.py:75-95."""
        lib = _lib.load()
        _lib.require_cuda(x, "x")
        if x.shape[-1] != self.in_features:
            raise ValueError(f"expected (..., {self.in_features}) input, got {tuple(x.shape)}")
        if x.device != self.codes.device:
            raise RuntimeError(f"input is on {x.device} but the layer's codes are on {self.codes.device}")
        lead = x.shape[:-1]
        x2 = x.reshape(-1, self.in_features)
        tokens = x2.shape[0]
        dtype = self.alpha.dtype
        if tokens == 0:
            return torch.empty((*lead, self.out_features), dtype=dtype, device=x.device)
        if tokens > self.gemv_max_tokens:
            if (self.fused_gemm and tokens <= self.fused_max_tokens and dtype != torch.float32
                    and self.in_features % 8 == 0):
                return self._forward_fused(x2.to(dtype)).reshape(*lead, self.out_features)
            out = torch.nn.functional.linear(x2.to(dtype), self._dequantize(), self.bias)
            return out.reshape(*lead, self.out_features)
        wtab, perm32, bias32, _ = self._prepared()
        if x2.dtype not in (torch.float32, torch.float16, torch.bfloat16):
            x2 = x2.to(dtype)
        if x2.stride(-1) != 1 or (tokens > 1 and x2.stride(0) < self.in_features):
            x2 = x2.contiguous()
        y = torch.empty((tokens, self.out_features), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            _lib.check(lib.tq_tl_gemv(_lib.ptr(self.codes), self.codes.shape[1], _lib.ptr(wtab), self.out_features,
                                      self.in_features, self.block_size, _lib.ptr(x2), _lib.dtype_code(x2.dtype),
                                      x2.stride(0) if tokens > 1 else self.in_features, tokens, _lib.ptr(perm32),
                                      _lib.ptr(bias32), _lib.ptr(y), self.out_features, _lib.stream()), "tq_tl_gemv")
        return y.to(dtype).reshape(*lead, self.out_features)

    def _forward_fused(self, x2: torch.Tensor) -> torch.Tensor:
        """Many tokens, 16-bit layer: one tcgen05 GEMM that expands the codes in shared memory (tq_tl_gemm_tc)."""
        lib = _lib.load()
        wtab, perm32, bias32, _ = self._prepared()
        if x2.stride(-1) != 1 or (x2.stride(0) * 2) % 16 != 0 or x2.data_ptr() % 16 != 0:
            x2 = x2.contiguous()
        tokens = x2.shape[0]
        y = torch.empty((tokens, self.out_features), dtype=x2.dtype, device=x2.device)
        work = None if perm32 is None else torch.empty((tokens, self.in_features), dtype=x2.dtype, device=x2.device)
        with torch.cuda.device(x2.device):
            _lib.check(lib.tq_tl_gemm_tc(_lib.ptr(self.codes), self.codes.shape[1], _lib.ptr(wtab), self.out_features,
                                         self.in_features, self.block_size, _lib.ptr(x2), _lib.dtype_code(x2.dtype),
                                         x2.stride(0), tokens, _lib.ptr(perm32), _lib.ptr(work), _lib.ptr(bias32),
                                         _lib.ptr(y), self.out_features, _lib.stream()), "tq_tl_gemm_tc")
        return y

    def memory_footprint(self) -> int:
        """Bytes this layer keeps resident (model.py:112-127 counts an int8 T and an int64 perm; here the codes are
        2 bits per weight and the kernels read perm as int32)."""
        codes = self.codes.numel() * 4
        scales = (self.alpha.numel() + self.mu.numel()) * self.alpha.element_size()
        perm = self.perm.numel() * 4
        bias = self.bias.numel() * self.bias.element_size() if self.bias is not None else 0
        return codes + scales + perm + bias

    def extra_repr(self) -> str:
        return (f"in_features={self.in_features}, out_features={self.out_features}, block_size={self.block_size}, "
                f"bias={self.bias is not None}, dtype={self.alpha.dtype}, codes=2-bit")


# ---------------------------------------------------------------------- model walkers (pure Python)
def get_model_layers(model: nn.Module) -> Dict[str, nn.Linear]:
    """model.py:130-136: every nn.Linear by qualified name."""
    return {name: mod for name, mod in model.named_modules() if isinstance(mod, nn.Linear)}


_LAYER_PATHS = {
    "llama": ("model.layers",), "llama2": ("model.layers",), "llama3": ("model.layers",),
    "qwen": ("model.layers",), "qwen3": ("model.layers",),
    "gemma": ("language_model.layers", "model.layers"), "gemma3": ("language_model.layers", "model.layers"),
    "opt": ("model.decoder.layers",),
    "bloom": ("transformer.h",),
}


def get_llm_layers(model: nn.Module, model_type: str = "llama") -> List[nn.Module]:
    """model.py:139-159: the list of transformer blocks of a HF causal LM."""
    if model_type not in _LAYER_PATHS:
        raise ValueError(f"Unknown model type: {model_type}")
    for path in _LAYER_PATHS[model_type]:
        obj = model
        try:
            for part in path.split("."):
                obj = getattr(obj, part)
        except AttributeError:
            continue
        return obj
    raise AttributeError(f"Cannot find layers in {model_type} model")


def find_linear_layers(module: nn.Module, prefix: str = "") -> Dict[str, nn.Linear]:
    """model.py:162-171: nn.Linear children, recursively, keyed by dotted path below `module`."""
    found = {}
    for name, child in module.named_children():
        path = f"{prefix}.{name}" if prefix else name
        if isinstance(child, nn.Linear):
            found[path] = child
        else:
            found.update(find_linear_layers(child, path))
    return found


def replace_linear_with_ternary(model: nn.Module, quantized_params: Dict[str, Dict[str, torch.Tensor]],
                                block_size: int = 128) -> nn.Module:
    """model.py:174-225: swap every named nn.Linear for a TernaryLinear holding its quantised parameters."""
    for name, params in quantized_params.items():
        *parents, leaf = name.split(".")
        parent = model
        for part in parents:
            parent = getattr(parent, part)
        old = getattr(parent, leaf)
        new = TernaryLinear(old.in_features, old.out_features, block_size=block_size, bias=old.bias is not None,
                            dtype=params["alpha"].dtype, device=old.weight.device)
        new.set_quantized_params(params["alpha"], params["mu"], params["T"], params["perm"],
                                 old.bias.data if old.bias is not None else None)
        setattr(parent, leaf, new)
        del old
    return model


_TYPE_KEYS = (("gemma-3", "gemma3"), ("gemma3", "gemma3"), ("gemma", "gemma"), ("llama-3", "llama3"),
              ("llama3", "llama3"), ("llama-2", "llama2"), ("llama2", "llama2"), ("llama", "llama"),
              ("qwen3", "qwen3"), ("qwen", "qwen"), ("opt", "opt"), ("bloom", "bloom"))


def get_model_type(model_name: str) -> str:
    """model.py:268-290: first matching family key in the lower-cased name; 'llama' when nothing matches."""
    low = model_name.lower()
    for key, family in _TYPE_KEYS:
        if key in low:
            return family
    return "llama"


def compute_model_size(model: nn.Module, quantized: bool = False) -> float:
    """model.py:293-303: parameters + buffers in GiB."""
    total = sum(p.numel() * p.element_size() for p in model.parameters())
    total += sum(b.numel() * b.element_size() for b in model.buffers())
    return total / (1024 ** 3)


def compute_compression_ratio(original_size: float, quantized_size: float) -> float:
    """model.py:306-308."""
    return original_size / quantized_size
