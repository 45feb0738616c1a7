"""B200-native ternary GPTQ hot path (PT2-LLM): the reference's gptq / quantizer / reorder / pack API, its ternary
inference layer (model.py) and its whole-model driver (main.py) over hand-written sm_100a CUDA kernels behind a C ABI
(include/tq100.h, libtq100.so)."""

import os as _os

# The chains of a layer's linears run on many CUDA streams (seven chains, each with two helper streams inside the
# inverse).  With the default of 8 hardware work queues, streams alias onto the same queue and serialise falsely (measured:
# 135 -> 126 ms per 7B layer with 32).  Read by the driver when the CUDA context is created, so it only takes effect if this
# package is imported before the first CUDA call; an explicit user setting wins.
_os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

from . import _lib  # noqa: E402,F401
from .quantizer import (  # noqa: F401
    AsymmetricTernaryQuantizer,
    compute_quantization_error,
    compute_output_error,
)
from .reorder import (  # noqa: F401
    SSRReorderer,
    compute_column_similarity_to_mean,
    select_next_block_ssr,
    apply_permutation,
    apply_permutation_to_input,
)
from .gptq import GPTQ, GPTQQuantizer, HessianState, quantize_layer  # noqa: F401
from .utils import (  # noqa: F401
    pack_ternary,
    unpack_ternary,
    compute_bits_per_weight,
    stored_bits_per_weight,
    save_quantized_model,
    load_quantized_model,
    set_seed,
    get_calibration_data,
    prepare_calibration_inputs,
    evaluate_perplexity,
)
from .model import (  # noqa: F401
    TernaryLinear,
    get_model_layers,
    get_llm_layers,
    find_linear_layers,
    replace_linear_with_ternary,
    get_model_type,
    compute_model_size,
    compute_compression_ratio,
)

from .main import PT2LLMQuantizer  # noqa: F401

__version__ = "0.1.0"
