"""GPTQ layer quantizer with ATQ + SSR -- B200 mirror of the reference's ``gptq.py``
(``/root/reference/gptq.py``: class ``GPTQ`` :21-230, ``GPTQQuantizer`` :233-272) and of the inline
loop ``main.py`` really runs (``PT2LLMQuantizer.quantize_layer``, ``/root/reference/main.py:102-230``).

Public surface is the reference's: ``GPTQ(layer, block_size, percdamp)``, ``add_batch(inp)``,
``quantize(use_ssr) -> (alpha, mu, T, perm)``, ``get_quantized_weight()``; ``fasterquant`` is an alias
of ``quantize`` (BASELINE north_star's name for it).  Underneath:

  add_batch  -> tq_hessian_accum   TMA-fed tcgen05/TMEM SYRK, fp32 accumulate (fp16/bf16 activations);
                                   fp32 activations take the CUDA-core fp32 SYRK
  quantize   -> tq_hessian_finalize (scale + damp), tq_chol_inverse (potrf + potri), tq_sweep_layer
                (per block: SSR select, AGA vector, fused ATQ fit + error, error-feedback GEMM),
                all enqueued on the current stream with one host sync at the end (the Cholesky status)

Everything runs on the layer's CUDA device; there is no CPU fallback.
"""

from typing import Dict, Optional, Tuple

import torch
import torch.nn as nn

try:
    from . import _lib
    from .quantizer import AsymmetricTernaryQuantizer  # noqa: F401  (re-exported like the reference)
    from .reorder import SSRReorderer, select_next_block_ssr, compute_column_similarity_to_mean  # noqa: F401
except ImportError:
    import _lib
    from quantizer import AsymmetricTernaryQuantizer  # noqa: F401
    from reorder import SSRReorderer, select_next_block_ssr, compute_column_similarity_to_mean  # noqa: F401

_AGA = {"none": _lib.AGA_NONE, "hessian": _lib.AGA_HESSIAN, "activations": _lib.AGA_ACTIVATIONS}


class HessianState:
    """Accumulated H = sum X'X for one layer INPUT, shareable between the linears that read the same
    activations (q/k/v, gate/up) -- SURVEY 8f N1.  Holds the raw accumulator, the token count and a
    cache of the damped matrix and its inverse per ``percdamp``."""

    def __init__(self, columns: int, device):
        self.columns = columns
        self.device = torch.device(device)
        self.H = torch.zeros((columns, columns), device=self.device, dtype=torch.float32)
        self.nsamples = 0
        self.hessian_path = _lib.HESS_AUTO
        self._upper_only = False      # the tcgen05 path fills only tiles touching the upper triangle
        self._cache = {}
        self._cache_events = {}
        self._full_event = None       # (event, stream) of the last symmetrisation, for readers on other streams

    def add_batch(self, inp: torch.Tensor):
        lib = _lib.load()
        _lib.require_cuda(inp, "inp")
        if inp.dim() == 3:
            inp = inp.reshape(-1, inp.shape[-1])               # gptq.py:68-69
        if inp.dim() != 2 or inp.shape[1] != self.columns:
            raise ValueError(f"add_batch: expected (..., {self.columns}) activations, got {tuple(inp.shape)}")
        x = inp.detach()
        if x.dtype not in (torch.float32, torch.float16, torch.bfloat16):
            x = x.float()
        if x.stride(1) != 1 or (x.stride(0) * x.element_size()) % 16 != 0 or x.data_ptr() % 16 != 0:
            x = x.contiguous()
        nt = x.shape[0]
        with torch.cuda.device(self.device):
            _lib.check(lib.tq_hessian_accum(_lib.ptr(self.H), self.columns, _lib.ptr(x), nt, self.columns,
                                            x.stride(0), _lib.dtype_code(x.dtype), self.hessian_path,
                                            _lib.stream()), "tq_hessian_accum")
        x.record_stream(torch.cuda.current_stream(self.device))
        self._upper_only = True
        self._full_event = None
        self._cache.clear()
        self.nsamples += nt                                     # gptq.py:76

    def full(self) -> torch.Tensor:
        """H as a full symmetric matrix (mirrors the upper triangle on first read).  Linears sharing this state may read
        it from other streams than the one the mirroring kernel ran on: those wait for its event."""
        cur = torch.cuda.current_stream(self.device)
        if self._upper_only:
            lib = _lib.load()
            with torch.cuda.device(self.device):
                _lib.check(lib.tq_symmetrize(_lib.ptr(self.H), self.columns, self.columns, _lib.stream()),
                           "tq_symmetrize")
            self._upper_only = False
            ev = torch.cuda.Event()
            ev.record(cur)
            self._full_event = (ev, cur)
        elif self._full_event is not None and self._full_event[1] != cur:
            cur.wait_event(self._full_event[0])
        return self.H

    def damped(self, percdamp: float) -> torch.Tensor:
        """gptq.py:94-98 only: Hd = H/nsamples + percdamp*mean(diag)*I (no inverse)."""
        if self.nsamples <= 0:
            raise RuntimeError("quantize() called before add_batch(): the Hessian is empty")
        lib = _lib.load()
        m, dev = self.columns, self.device
        Hd = torch.empty((m, m), dtype=torch.float32, device=dev)
        scratch = torch.empty(8, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(lib.tq_hessian_finalize(_lib.ptr(Hd), _lib.ptr(self.H), m, float(self.nsamples), float(percdamp),
                                               _lib.ptr(scratch), _lib.stream()), "tq_hessian_finalize")
        scratch.record_stream(torch.cuda.current_stream(dev))
        return Hd

    def damped_inverse(self, percdamp: float):
        """gptq.py:94-106: (Hd, Hinv, info) with Hd = H/nsamples + percdamp*mean(diag)*I and
        Hinv = cholesky_inverse(cholesky(Hd)); ``info`` is a device int (0 = ok)."""
        key = float(percdamp)
        if key in self._cache:
            # linears sharing this Hessian may sweep on other streams than the one the inverse was enqueued on
            ev = self._cache_events.get(key)
            if ev is not None and ev[1] != torch.cuda.current_stream(self.device):
                torch.cuda.current_stream(self.device).wait_event(ev[0])
            return self._cache[key]
        if self.nsamples <= 0:
            raise RuntimeError("quantize() called before add_batch(): the Hessian is empty")
        lib = _lib.load()
        m = self.columns
        dev = self.device
        Hd = torch.empty((m, m), dtype=torch.float32, device=dev)
        Hinv = torch.empty((m, m), dtype=torch.float32, device=dev)
        work = torch.empty(lib.tq_chol_workspace_floats(m), dtype=torch.float32, device=dev)
        scratch = torch.empty(8, dtype=torch.float32, device=dev)
        info = torch.zeros(1, dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            st = _lib.stream()
            _lib.check(lib.tq_hessian_finalize(_lib.ptr(Hd), _lib.ptr(self.H), m, float(self.nsamples), key,
                                               _lib.ptr(scratch), st), "tq_hessian_finalize")
            _lib.check(lib.tq_chol_inverse(_lib.ptr(Hinv), _lib.ptr(Hd), m, _lib.ptr(work), _lib.ptr(info), st),
                       "tq_chol_inverse")
        for t in (work, scratch):
            t.record_stream(torch.cuda.current_stream(dev))
        self._cache[key] = (Hd, Hinv, info)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(dev))
        self._cache_events[key] = (ev, torch.cuda.current_stream(dev))
        return self._cache[key]


def _exact_feedback_matrix(Hinv: torch.Tensor, static_perm: Optional[torch.Tensor], block: int) -> torch.Tensor:
    """Coefficient source of the OBS-exact block feedback, in SWEEP order.

    With U the upper Cholesky factor of H^-1 in sweep order (H^-1 = U'U), the inverse Hessian of the problem that remains
    after the first k blocks is U_RR' U_RR, and the exact update of the remaining columns R after block B is
    dW_R = -E_B (Hc_BB)^-1 Hc_BR = -E_B U_BB^-1 U_BR.  The sweep's coefficient kernel forms G[blk, rem] / G[blk, blk]
    (gptq.py:173-181), so the matrix handed to it is G with unit diagonal and G[B, R] = U_BB^-1 U_BR per block: one
    Cholesky factorisation and one triangular solve per block (library calls on the GPU, prologue only; default off)."""
    m = Hinv.shape[0]
    Hp = Hinv
    if static_perm is not None:
        p64 = static_perm.long()
        Hp = Hinv.index_select(0, p64).index_select(1, p64)
    U = torch.linalg.cholesky(Hp.double(), upper=True)
    G = torch.eye(m, dtype=torch.float64, device=Hinv.device)
    for k0 in range(0, m, block):
        k1 = min(k0 + block, m)
        if k1 < m:
            G[k0:k1, k1:] = torch.linalg.solve_triangular(U[k0:k1, k0:k1], U[k0:k1, k1:], upper=True)
    return G.float().contiguous()


class GPTQ:
    """gptq.py:21-230."""

    def __init__(self, layer: nn.Linear, block_size: int = 128, percdamp: float = 0.01,
                 hessian: Optional[HessianState] = None):
        self.layer = layer
        self.block_size = block_size
        self.percdamp = percdamp
        self.device = layer.weight.device
        self.dtype = layer.weight.dtype
        _lib.require_cuda(layer.weight, "layer.weight")
        self.rows, self.columns = layer.weight.shape
        if not (1 <= block_size <= 512):
            raise ValueError("block_size must be in [1, 512]")
        self.state = hessian if hessian is not None else HessianState(self.columns, self.device)
        if self.state.columns != self.columns:
            raise ValueError("shared HessianState has a different width than this layer's in_features")
        self.alpha = None
        self.mu = None
        self.T = None
        self.perm = None
        self.T_int8 = None
        self.info = None
        self.sweep_flags = 0          # _lib.SWEEP_ROW_SHARD when this object sweeps one row slab of the layer

    # the reference exposes H and nsamples as plain attributes (gptq.py:50-51)
    @property
    def H(self) -> torch.Tensor:
        return self.state.full()

    @H.setter
    def H(self, value: torch.Tensor):
        self.state.H = value.to(device=self.device, dtype=torch.float32).contiguous()
        self.state._upper_only = False
        self.state._cache.clear()

    @property
    def nsamples(self) -> int:
        return self.state.nsamples

    @nsamples.setter
    def nsamples(self, value: int):
        self.state.nsamples = value
        self.state._cache.clear()

    def add_batch(self, inp: torch.Tensor):
        """gptq.py:59-76: H += X'X over (batch, seq, in) or (tokens, in) activations."""
        self.state.add_batch(inp)

    @torch.no_grad()
    def quantize(self, use_ssr: bool = True, aga: str = "hessian", order: Optional[str] = None,
                 max_iter: int = 100, feedback: str = "reference",
                 dead_columns: bool = False) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
        """gptq.py:78-199.  New optional arguments keep the reference's behaviour by default:
        aga   'hessian' (gptq.py:147-150) | 'activations' (main.py:177-180) | 'none'
        order None -> 'ssr' if use_ssr else 'sequential'; 'actorder' = descending diag(H) (extension).
        Canonical-GPTQ exactness options (they sit where the reference has gptq.py:158-186):
        feedback 'reference' = rows of the static H^-1 over its diagonal (gptq.py:173-181) |
                 'block_exact' = the OBS-exact block update through the Cholesky factor of H^-1 (static orders only)
        dead_columns  True: inputs that are identically zero in the calibration data get H_jj = 1 and weight 0."""
        self.enqueue(use_ssr=use_ssr, aga=aga, order=order, max_iter=max_iter, feedback=feedback, dead_columns=dead_columns)
        return self.finish()

    @torch.no_grad()
    def enqueue(self, use_ssr: bool = True, aga: str = "hessian", order: Optional[str] = None, max_iter: int = 100,
                feedback: str = "reference", dead_columns: bool = False):
        """Asynchronous half of quantize(): enqueue prologue (scale + damp, Cholesky inverse) and the whole
        column sweep on the CURRENT stream and return without touching the host again.  finish() must follow.
        A layer driver enqueues several linears on different streams before finishing any of them."""
        lib = _lib.load()
        if aga not in _AGA:
            raise ValueError(f"aga must be one of {sorted(_AGA)}")
        if order is None:
            order = "ssr" if use_ssr else "sequential"
        if order not in ("ssr", "sequential", "actorder"):
            raise ValueError("order must be 'ssr', 'sequential' or 'actorder'")
        if feedback not in ("reference", "block_exact"):
            raise ValueError("feedback must be 'reference' or 'block_exact'")
        if feedback == "block_exact" and order == "ssr":
            raise ValueError("feedback='block_exact' needs a static sweep order ('sequential' or 'actorder'): the exact "
                             "recursion eliminates columns in an order known up front")
        n, m, b = self.rows, self.columns, self.block_size
        dev = self.device
        nb = (m + b - 1) // b

        state, dead = self.state, None
        if dead_columns:
            # canonical GPTQ: dead = diag(H) == 0; H[dead, dead] = 1; W[:, dead] = 0.  Applied to a private copy of the
            # accumulator (the state may be shared with other linears) and without a host round trip.
            Hfull = self.state.full()
            dead = torch.diagonal(Hfull) == 0
            state = HessianState(m, dev)
            state.H = Hfull.clone()
            torch.diagonal(state.H).masked_fill_(dead, float(self.state.nsamples))      # = 1 after the 1 / nsamples scaling
            state.nsamples = self.state.nsamples
        if aga == "activations":
            # the AGA Gram reads H[blk, blk] on both sides of the diagonal: mirror the tcgen05 SYRK's upper tiles BEFORE the
            # inverse is enqueued (its cache event then also covers the mirroring for linears sharing this Hessian)
            state.full()
        Hd, Hinv, info = state.damped_inverse(self.percdamp)
        static_perm = None
        order_code = {"ssr": _lib.ORDER_SSR, "sequential": _lib.ORDER_SEQUENTIAL, "actorder": _lib.ORDER_STATIC}[order]
        if order == "actorder":
            static_perm = torch.argsort(torch.diagonal(Hd), descending=True, stable=True).to(torch.int32).contiguous()

        def run(hinv):
            W = self.layer.weight.data.detach().to(torch.float32).clone().contiguous()     # gptq.py:91
            if dead is not None:
                W.masked_fill_(dead.unsqueeze(0), 0.0)
            hd, hraw, code, sperm = Hd, state.H, order_code, static_perm
            if feedback == "block_exact":
                hinv = _exact_feedback_matrix(hinv, static_perm, b)          # in sweep order when static_perm is given
            if order == "actorder":
                # A fixed sweep order is the sequential sweep of the problem with its columns permuted up front (the
                # extension's own definition).  Doing the permutation physically -- one gather of W and of the
                # Hessian matrices -- keeps every block step on contiguous columns: the feedback GEMM then takes its TMA
                # epilogue instead of a 4-byte-granular gather through an arbitrary permutation (measured at 5120 x 13824:
                # 2.4 ms per block step gathered, vs ~0.1 ms contiguous).  Same values, same operation order: same bits.
                p64 = static_perm.long()
                W = W.index_select(1, p64).contiguous()
                if feedback != "block_exact":                                # (the exact matrix is built in sweep order)
                    hinv = hinv.index_select(0, p64).index_select(1, p64).contiguous()
                if aga == "hessian":
                    hd = Hd.index_select(0, p64).index_select(1, p64).contiguous()
                elif aga == "activations":
                    hraw = state.full().index_select(0, p64).index_select(1, p64).contiguous()
                code, sperm = _lib.ORDER_SEQUENTIAL, None
            T8 = torch.empty((n, m), dtype=torch.int8, device=dev)
            alpha = torch.empty((n, nb), dtype=torch.float32, device=dev)
            mu = torch.empty((n, nb), dtype=torch.float32, device=dev)
            perm = torch.empty(m, dtype=torch.int32, device=dev)
            ws_bytes = lib.tq_sweep_workspace_bytes(n, m, b)
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
            with torch.cuda.device(dev):
                _lib.check(lib.tq_sweep_layer(_lib.ptr(W), W.stride(0), n, m, _lib.ptr(hd), _lib.ptr(hraw),
                                              _lib.ptr(hinv), b, code, _AGA[aga], int(max_iter),
                                              _lib.ptr(sperm), _lib.ptr(T8), _lib.ptr(alpha), _lib.ptr(mu),
                                              _lib.ptr(perm), _lib.ptr(ws), ws_bytes, int(self.sweep_flags), _lib.stream()),
                           "tq_sweep_layer")
            for t in (W, ws, hinv, hd, hraw):
                if t is not None:
                    t.record_stream(torch.cuda.current_stream(dev))
            if order == "actorder":
                # codes back to ORIGINAL column positions (gptq.py:155); block k of alpha / mu belongs to perm[128k : 128k+128]
                T8 = torch.empty_like(T8).index_copy_(1, static_perm.long(), T8)
                perm = static_perm
            return alpha, mu, T8, perm

        self._sweep_state = state
        self._pending = (run, Hd, info, run(Hinv), torch.cuda.current_stream(dev))

    @torch.no_grad()
    def finish(self) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
        """Synchronous half of quantize(): read the Cholesky status (the one host sync of the layer), take the
        reference's pseudo-inverse route if the factorisation failed (gptq.py:104-106), publish the results."""
        if getattr(self, "_pending", None) is None:
            raise RuntimeError("finish() without enqueue()")
        run, Hd, info, (alpha, mu, T8, perm), stream = self._pending
        self._pending = None
        # everything was enqueued on `stream`, which need not be the caller's current stream: wait for it before
        # touching any result from the host or from another stream
        stream.synchronize()
        self.info = int(info.item())
        if self.info != 0:
            with torch.cuda.stream(stream):
                Hinv = torch.linalg.pinv(Hd)          # library SVD, as in the reference
                st = getattr(self, "_sweep_state", None) or self.state
                st._cache[float(self.percdamp)] = (Hd, Hinv, torch.zeros_like(info))
                ev = torch.cuda.Event()
                ev.record(stream)
                st._cache_events[float(self.percdamp)] = (ev, stream)
                alpha, mu, T8, perm = run(Hinv)
            stream.synchronize()
        self.alpha = alpha.to(self.dtype)
        self.mu = mu.to(self.dtype)
        self.T_int8 = T8
        self.T = T8.to(self.dtype)                     # gptq.py:109: T has W's dtype
        self.perm = perm.to(torch.int64)               # gptq.py:191
        return self.alpha, self.mu, self.T, self.perm

    fasterquant = quantize

    def get_quantized_weight(self) -> torch.Tensor:
        """gptq.py:201-230."""
        if self.T is None:
            raise RuntimeError("Must call quantize() first")
        lib = _lib.load()
        n, m, b = self.rows, self.columns, self.block_size
        nb = (m + b - 1) // b
        Wq = torch.empty((n, m), dtype=torch.float32, device=self.device)
        alpha = self.alpha.float().contiguous()
        mu = self.mu.float().contiguous()
        perm = self.perm.to(torch.int32).contiguous()
        with torch.cuda.device(self.device):
            _lib.check(lib.tq_dequant(_lib.ptr(alpha), _lib.ptr(mu), nb, _lib.ptr(self.T_int8), n, m, _lib.ptr(perm), b,
                                      _lib.ptr(Wq), _lib.stream()), "tq_dequant")
        return Wq.to(self.dtype)

    def packed_codes(self):
        """2-bit codes of T in original column positions (utils.pack_ternary layout) + per-block scales."""
        try:
            from .utils import pack_ternary
        except ImportError:
            from utils import pack_ternary
        if self.T_int8 is None:
            raise RuntimeError("Must call quantize() first")
        packed, shape = pack_ternary(self.T_int8)
        return {"packed": packed, "shape": tuple(shape), "alpha": self.alpha, "mu": self.mu, "perm": self.perm}


class GPTQQuantizer:
    """gptq.py:233-272."""

    def __init__(self, model: nn.Module, block_size: int = 128, percdamp: float = 0.01, use_ssr: bool = True):
        self.model = model
        self.block_size = block_size
        self.percdamp = percdamp
        self.use_ssr = use_ssr
        self.quantizers: Dict[str, GPTQ] = {}
        self.calibration_data = []

    def add_calibration_data(self, data: torch.Tensor):
        self.calibration_data.append(data)

    def prepare_quantizer(self, name: str, layer: nn.Linear):
        self.quantizers[name] = GPTQ(layer, self.block_size, self.percdamp)

    def quantize_layer(self, name: str) -> Dict[str, torch.Tensor]:
        if name not in self.quantizers:
            raise ValueError(f"Layer {name} not prepared for quantization")
        alpha, mu, T, perm = self.quantizers[name].quantize(use_ssr=self.use_ssr)
        return {"alpha": alpha, "mu": mu, "T": T, "perm": perm}


@torch.no_grad()
def quantize_layer(layer: nn.Linear, layer_name: str, calibration_activations: torch.Tensor,
                   block_size: int = 128, percdamp: float = 0.01, use_ssr: bool = True,
                   device=None) -> Dict[str, torch.Tensor]:
    """Drop-in body for ``PT2LLMQuantizer.quantize_layer`` (main.py:102-230): same inputs, same CPU
    dict {'alpha','mu','T' (int8),'perm'}; AGA is fed the raw-activation Gram (main.py:177-180)."""
    dev = torch.device(device) if device is not None else layer.weight.device
    if layer.weight.device != dev:
        layer = layer.to(dev)
    g = GPTQ(layer, block_size=block_size, percdamp=percdamp)
    g.add_batch(calibration_activations.to(dev))
    alpha, mu, _, perm = g.quantize(use_ssr=use_ssr, aga="activations")
    return {"alpha": alpha.float().cpu(), "mu": mu.float().cpu(), "T": g.T_int8.cpu(), "perm": perm.cpu()}
