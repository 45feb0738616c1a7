"""B200-native ternary GPTQ hot path (PT2-LLM): the reference's gptq / quantizer / reorder / pack API
over hand-written sm_100a CUDA kernels behind a C ABI (include/tq100.h, libtq100.so)."""

from . import _lib  # noqa: F401
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
from .utils import pack_ternary, unpack_ternary  # noqa: F401

__version__ = "0.1.0"
