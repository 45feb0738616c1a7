"""Import alias: the package directory is named after the reference repo
(``snlp---tenary-post-train-quantization_b200``), which is not a Python identifier.  ``import tq100``
loads that directory as the package ``tq100`` (``tq100.gptq``, ``tq100.quantizer``, ...)."""

import importlib.util
import os
import sys

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "snlp---tenary-post-train-quantization_b200")
_spec = importlib.util.spec_from_file_location(__name__, os.path.join(_DIR, "__init__.py"),
                                               submodule_search_locations=[_DIR])
_mod = importlib.util.module_from_spec(_spec)
sys.modules[__name__] = _mod
_spec.loader.exec_module(_mod)
