"""Model-level driver -- B200 mirror of the reference's ``PT2LLMQuantizer`` (``/root/reference/main.py:40-335``).
SURVEY 8f N1: the caller of the hot path.

Same constructor, ``quantize_layer(layer, layer_name, calibration_activations)`` (main.py:102-230) and
``quantize()`` (main.py:232-311) returning ``{f"layer_{i}.{name}": {'alpha','mu','T' (int8),'perm'}}`` on the CPU, with
every quantised linear's weight overwritten by its dequantised value (main.py:298-299).  What changes is how the
calibration activations reach the hot path:

  * the reference hooks the linears of transformer layer i and re-runs the WHOLE model over every calibration sample for
    every i (main.py:278-282: O(L^2) layer forwards), keeping every hooked input and ``torch.cat``-ing them
    (main.py:293: 128 x 2048 x m values per linear);
  * here the inputs of layer 0 are captured once (a forward pre-hook that stops the forward), and each layer is run twice
    on its cached inputs: once with its original weights while forward hooks stream every linear's input straight into
    ``GPTQ.add_batch`` (the tcgen05 Hessian; nothing is concatenated), once with the quantised weights to produce the
    next layer's inputs.  In exact arithmetic this is the reference's computation: layers before i already hold their
    quantised weights in both schemes, and all linears of layer i see activations produced by layer i's original
    weights.  Linears that are handed the same tensor (q/k/v; gate/up) share one Hessian and one inverse
    (``share_inputs``; the Hessian of identical inputs is identical), their prologue + sweep chains overlap on CUDA
    streams (``pipeline.LayerDriver``).
  * ``_dequantize_weight`` applies ``GPTQ.get_quantized_weight`` semantics (gptq.py:201-230).  The reference's
    ``main.py:313-335`` indexes ``T`` as if it were stored in sweep order although main.py:185 stores it in original
    positions (with SSR its reconstruction error is 1.37 instead of 0.32); for ``use_ssr=False`` the two agree.

Under ``torchrun`` pass ``shard=sharded.ShardContext(rank, world, device)``: the calibration samples are dealt to the ranks
(each rank runs the forwards of its own), the Hessians are all-reduced per layer, the linears are dealt to the ranks by
chain cost and their results broadcast, so every rank ends with the same quantised model and parameters.

Dataset loading (``utils.get_calibration_data``) needs the `datasets` cache or the network; without them pass the token
tensors to ``quantize(calibration_samples=...)``.
"""

import inspect
import time
from typing import Callable, Dict, List, Optional

import torch
import torch.nn as nn

try:
    from . import gptq as _gptq
    from .model import find_linear_layers, get_llm_layers, replace_linear_with_ternary as _replace_linear_with_ternary
    from .pipeline import LayerDriver
    from .utils import set_seed, get_calibration_data as _get_calibration_data
except ImportError:
    import gptq as _gptq
    from model import find_linear_layers, get_llm_layers, replace_linear_with_ternary as _replace_linear_with_ternary
    from pipeline import LayerDriver
    from utils import set_seed, get_calibration_data as _get_calibration_data


class _StopForward(Exception):
    pass


# decoder-layer keyword arguments that carry (or switch on) a KV cache in Hugging Face models
_CACHE_KWARGS = ("past_key_value", "past_key_values", "use_cache", "layer_past")


class PT2LLMQuantizer:
    """main.py:40-335."""

    def __init__(self, model: nn.Module, tokenizer, model_type: str = "llama", block_size: int = 128,
                 num_calibration_samples: int = 128, seq_len: int = 2048, use_ssr: bool = True, percdamp: float = 0.01,
                 seed: int = 42, device: str = "cuda", share_inputs: bool = True, num_streams: int = 4,
                 gptq_factory: Optional[Callable] = None, chain_runner: Optional[Callable] = None, shard=None):
        self.model = model
        self.tokenizer = tokenizer
        self.model_type = model_type
        self.block_size = block_size
        self.num_calibration_samples = num_calibration_samples
        self.seq_len = seq_len
        self.use_ssr = use_ssr
        self.percdamp = percdamp
        self.seed = seed
        self.device = torch.device(device)
        self.share_inputs = share_inputs
        self.num_streams = num_streams
        # seams for host-logic tests: how a per-linear quantiser is made and how a layer's quantisers are run
        self._gptq_factory = gptq_factory or (lambda layer, hessian: _gptq.GPTQ(layer, block_size=self.block_size,
                                                                                 percdamp=self.percdamp, hessian=hessian))
        self._chain_runner = chain_runner
        # multi-GPU: a sharded.ShardContext under torchrun.  Every rank holds the model, runs the forwards of
        # ITS calibration samples and accumulates local Hessians; per layer the Hessians are all-reduced, the linears
        # are dealt to the ranks by chain cost, each owner quantises its linears and broadcasts their results.
        self.shard = shard if shard is not None and getattr(shard, "world", 1) > 1 else None
        set_seed(seed)
        self.quantized_params: Dict[str, Dict[str, torch.Tensor]] = {}
        self.layer_forwards = 0              # transformer-layer forwards executed by quantize() (2 per layer and sample)

    # ------------------------------------------------------------------ main.py:90-100
    def get_calibration_data(self) -> List[torch.Tensor]:
        """main.py:90-100: wikitext-2 windows through utils.get_calibration_data.  Without the `datasets` cache or the
        network this raises; pass the (1, seq_len) token tensors to quantize(calibration_samples=...) instead."""
        if self.tokenizer is None:
            raise RuntimeError("get_calibration_data() needs a tokenizer (and the `datasets` cache or the network); pass the "
                               "token tensors to quantize(calibration_samples=...)")
        try:
            return _get_calibration_data(self.tokenizer, dataset_name="wikitext", dataset_config="wikitext-2-raw-v1",
                                         num_samples=self.num_calibration_samples, seq_len=self.seq_len, seed=self.seed)
        except Exception as exc:
            raise RuntimeError("get_calibration_data() could not load wikitext-2 (it needs the `datasets` cache or the "
                               "network); pass the token tensors to quantize(calibration_samples=...)") from exc

    # ------------------------------------------------------------------ main.py:102-230
    def quantize_layer(self, layer: nn.Linear, layer_name: str, calibration_activations: torch.Tensor):
        return _gptq.quantize_layer(layer, layer_name, calibration_activations, block_size=self.block_size,
                                    percdamp=self.percdamp, use_ssr=self.use_ssr, device=self.device)

    # ------------------------------------------------------------------ main.py:232-311
    @torch.no_grad()
    def quantize(self, calibration_samples: Optional[List[torch.Tensor]] = None) -> Dict[str, Dict[str, torch.Tensor]]:
        start = time.time()
        if calibration_samples is None:
            calibration_samples = self.get_calibration_data()
        layers = get_llm_layers(self.model, self.model_type)
        self.model.eval()
        if self.shard is not None:
            if len(calibration_samples) < self.shard.world:
                # a rank without samples would skip the forwards (and the collectives that follow them) and hang the rest
                raise ValueError(f"{len(calibration_samples)} calibration samples cannot be dealt to {self.shard.world} ranks")
            calibration_samples = [calibration_samples[i] for i in self.shard.my_samples(len(calibration_samples))]
        inputs = self._capture_first_layer_inputs(layers[0], calibration_samples)
        for layer_idx, layer in enumerate(layers):
            linear_layers = find_linear_layers(layer)
            quantizers = self._accumulate(layer, linear_layers, inputs)
            names = [n for n in linear_layers if n in quantizers]
            owners = self._reduce_and_deal(quantizers, names)
            rank = self.shard.rank if self.shard is not None else 0
            self._run_chains([quantizers[n] for n, o in zip(names, owners) if o == rank])
            for name, owner in zip(names, owners):
                g, linear = quantizers[name], linear_layers[name]
                if owner == rank:
                    res = (g.alpha.float(), g.mu.float(), g.T_int8, g.perm,
                           g.get_quantized_weight().to(linear.weight.dtype))                 # main.py:298-299
                else:
                    res = None
                alpha, mu, T8, perm, W_quant = self._publish(res, owner, g, linear)
                self.quantized_params[f"layer_{layer_idx}.{name}"] = {"alpha": alpha.cpu(), "mu": mu.cpu(), "T": T8.cpu(),
                                                                      "perm": perm.cpu()}    # main.py:225-230
                linear.weight.data = W_quant.to(linear.weight.device, linear.weight.dtype)
            del quantizers
            if layer_idx + 1 < len(layers):
                inputs = [self._advance(layer, args, kwargs) for args, kwargs in inputs]     # with the quantised weights
        self.elapsed = time.time() - start
        return self.quantized_params

    def _capture_first_layer_inputs(self, first_layer: nn.Module, samples):
        """One model forward per sample, stopped at the first transformer layer: [(args, kwargs)] of that layer."""
        captured = []

        def pre_hook(module, args, kwargs):
            # Every layer is later run TWICE on these kwargs (original weights, then quantised weights).  A live KV cache
            # among them would be appended to by the first pass and attended to by the second, so the cache and the flag
            # that enables it are dropped: the layers run cache-free, like a fresh whole-model forward of the reference
            # (main.py:278-282).
            kept = {}
            for k, v in kwargs.items():
                if k in _CACHE_KWARGS:
                    kept[k] = False if k == "use_cache" else None
                else:
                    kept[k] = v.detach() if torch.is_tensor(v) else v
            captured.append((tuple(a.detach() if torch.is_tensor(a) else a for a in args), kept))
            raise _StopForward()

        # ask the model not to build a KV cache at all when its forward knows the flag (Hugging Face causal LMs do)
        try:
            params = inspect.signature(self.model.forward).parameters
            takes_flag = "use_cache" in params or any(p.kind is inspect.Parameter.VAR_KEYWORD for p in params.values())
        except (TypeError, ValueError):
            takes_flag = False
        no_cache = {"use_cache": False} if takes_flag else {}
        handle = first_layer.register_forward_pre_hook(pre_hook, with_kwargs=True)
        try:
            for sample in samples:
                try:
                    self.model(sample.to(self.device), **no_cache)
                except _StopForward:
                    pass
        finally:
            handle.remove()
        if len(captured) != len(samples):
            raise RuntimeError("the first transformer layer was not reached for every calibration sample")
        return captured

    def _layer_forward(self, layer, args, kwargs):
        self.layer_forwards += 1
        out = layer(*args, **kwargs)
        return out[0] if isinstance(out, (tuple, list)) else out

    def _advance(self, layer, args, kwargs):
        return ((self._layer_forward(layer, args, kwargs),) + tuple(args[1:]), kwargs)

    def _accumulate(self, layer, linear_layers, inputs):
        """Run `layer` over the cached inputs with its ORIGINAL weights; forward hooks stream each linear's input into its
        quantiser's add_batch (main.py:262-275 without the list + cat).  Linears that receive the same tensor share one
        Hessian state: the hook of the first of them accumulates, the others only attach."""
        quantizers = {}
        owner_of_tensor = {}
        alive = []            # inputs seen in the current forward, kept alive so that an address cannot be reused within it

        def make_hook(name, linear):
            def hook(module, inp, out):
                x = inp[0] if isinstance(inp, tuple) else inp
                x = x.detach()
                alive.append(x)
                key = (x.data_ptr(), tuple(x.shape), tuple(x.stride()), x.dtype, x._version)
                first = owner_of_tensor.setdefault(key, name)      # first linear of this forward handed this tensor
                if name not in quantizers:
                    shared = quantizers[first].state if self.share_inputs and first != name else None
                    quantizers[name] = self._gptq_factory(linear, shared)
                    quantizers[name]._owner = first if shared is not None else name
                g = quantizers[name]
                if g._owner == name:
                    g.add_batch(x)
                elif first != g._owner:
                    raise RuntimeError(f"{name} shared a Hessian with {g._owner} but now receives a different tensor")
            return hook

        handles = [linear.register_forward_hook(make_hook(name, linear)) for name, linear in linear_layers.items()]
        try:
            for args, kwargs in inputs:
                owner_of_tensor.clear()          # tensor identity is only meaningful within one forward
                alive.clear()
                self._layer_forward(layer, args, kwargs)
        finally:
            for h in handles:
                h.remove()
        return quantizers

    def _reduce_and_deal(self, quantizers, names):
        """Single process: every linear is rank 0's.  Sharded: all-reduce every distinct Hessian and its token count
        (the NCCL all-reduce of H over NVLink of SURVEY 8e; identical on every rank afterwards), then deal the linears
        to the ranks by estimated chain duration (sharded.deal_linears: deterministic, same answer on every rank)."""
        if self.shard is None:
            return [0] * len(names)
        import torch.distributed as dist
        try:
            from .sharded import deal_linears
        except ImportError:
            from sharded import deal_linears
        seen = set()
        for name in names:
            st = quantizers[name].state
            if id(st) in seen:
                continue
            seen.add(id(st))
            H = st.H if torch.is_tensor(st.H) else torch.from_numpy(st.H)       # (numpy only in the host-logic tests)
            dist.all_reduce(H, op=dist.ReduceOp.SUM, group=self.shard.group)
            count = torch.tensor([st.nsamples], dtype=torch.int64, device=H.device)
            dist.all_reduce(count, op=dist.ReduceOp.SUM, group=self.shard.group)
            st.nsamples = int(count.item())
            if hasattr(st, "_cache"):
                st._cache.clear()
        return deal_linears([(quantizers[n].rows, quantizers[n].columns) for n in names], self.shard.world)

    def _publish(self, res, owner, g, linear):
        """Owner -> everyone: (alpha, mu, T int8, perm, dequantised weight) of one linear."""
        if self.shard is None:
            return res
        import torch.distributed as dist
        n, m = g.rows, g.columns
        nb = (m + self.block_size - 1) // self.block_size
        dev = linear.weight.device
        if res is None:
            res = (torch.empty((n, nb), dtype=torch.float32, device=dev), torch.empty((n, nb), dtype=torch.float32, device=dev),
                   torch.empty((n, m), dtype=torch.int8, device=dev), torch.empty(m, dtype=torch.int64, device=dev),
                   torch.empty((n, m), dtype=linear.weight.dtype, device=dev))
        out = []
        for t in res:
            t = t.to(dev).contiguous()
            dist.broadcast(t, src=owner, group=self.shard.group)
            out.append(t)
        return tuple(out)

    def _run_chains(self, gs):
        if not gs:
            return gs
        if self._chain_runner is not None:
            return self._chain_runner(gs, self.use_ssr)
        driver = LayerDriver(gs[0].device, block_size=self.block_size, percdamp=self.percdamp, num_streams=self.num_streams)
        # AGA on the raw-activation Gram, as main.py:177-180 feeds it
        return driver.run_chains(gs, use_ssr=self.use_ssr, aga="activations")

    def replace_with_ternary(self, dtype: Optional[torch.dtype] = None) -> nn.Module:
        """After quantize(): swap every quantised nn.Linear of the model for a TernaryLinear on its 2-bit codes
        (model.replace_linear_with_ternary, model.py:174-225).  The keys of ``quantized_params`` are the reference's
        ``layer_{i}.{name}`` (main.py:290), so each transformer layer is handed its own sub-dictionary.  ``dtype`` casts
        alpha / mu (and with them the layer) -- the reference's default layer dtype is fp16 (model.py:34)."""
        layers = get_llm_layers(self.model, self.model_type)
        per_layer: Dict[int, Dict[str, Dict[str, torch.Tensor]]] = {}
        for key, params in self.quantized_params.items():
            head, name = key.split(".", 1)
            if dtype is not None:
                params = dict(params, alpha=params["alpha"].to(dtype), mu=params["mu"].to(dtype))
            per_layer.setdefault(int(head[len("layer_"):]), {})[name] = params
        for idx, group in per_layer.items():
            _replace_linear_with_ternary(layers[idx], group, block_size=self.block_size)
        return self.model

    # ------------------------------------------------------------------ main.py:313-335
    def _dequantize_weight(self, params: Dict[str, torch.Tensor]) -> torch.Tensor:
        """Wq[:, perm[blk_k]] = alpha_k * T[:, perm[blk_k]] + mu_k (gptq.py:201-230); see the module docstring for how
        this differs from the reference's main.py:313-335 under SSR."""
        T = params["T"].float()
        alpha, mu, perm = params["alpha"], params["mu"], params["perm"].long()
        n, m = T.shape
        W = torch.zeros(n, m, dtype=alpha.dtype)
        for b in range(alpha.shape[1]):
            cols = perm[b * self.block_size:min((b + 1) * self.block_size, m)]
            W[:, cols] = alpha[:, b:b + 1] * T[:, cols].to(alpha.dtype) + mu[:, b:b + 1]
        return W
