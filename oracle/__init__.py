"""CPU oracle for the ternary GPTQ hot path -- TEST INFRASTRUCTURE, NOT PRODUCT.

This package restates, in plain numpy (+ scipy's LAPACK bindings for potrf/potri),
the algorithm of the reference's hot path (``/root/reference/gptq.py``,
``quantizer.py``, ``reorder.py``, ``utils.py:189-248``).  Every function cites the
reference file:line it follows.  The arithmetic of the reference lives in a
third-party dependency -- PyTorch (``requirements.txt:1`` ``torch>=2.0.0``,
unpinned; 2.11.0+cu128 in this image, CPU math = oneMKL + LAPACK) -- so the
restatement is of the published algorithm of each ATen call, anchored on the
reference's own call sites.

Pinning: the reference ships no tests and no golden vectors, so the
oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF, produced in the build
container by ``tests/golden/make_golden.py`` (imports the unmodified reference
from ``/root/reference`` through an isolated loader) and committed as small
``.npz`` fixtures under ``tests/golden/``; plus the values ``examples.py``
prints.  ``tests/test_oracle_golden.py`` is the pin.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package, and only as the checker /
the CPU arm.  The product path (the package next door) never imports it and has
no CPU fallback.
"""

from .atq import (  # noqa: F401
    ternary_init,
    build_optimal_grid,
    flexible_round,
    iterative_ternary_fitting,
    aga_from_gram,
    activation_aware_grid_alignment,
    atq_quantize,
    dequantize,
    compute_quantization_error,
    compute_output_error,
)
from .ssr import (  # noqa: F401
    column_similarity_to_mean,
    select_next_block_ssr,
)
from .gptq import (  # noqa: F401
    hessian_add_batch,
    damped_inverse,
    quantize_layer,
    quantize_layer_main,
    get_quantized_weight,
    reconstruction_error,
)
from .pack import pack_ternary, unpack_ternary  # noqa: F401
from . import ternary_linear  # noqa: F401  (model.py:17-127)
