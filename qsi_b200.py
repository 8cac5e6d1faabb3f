"""Import shim: the product package lives in `quantization-sparsity-interplay_b200/` (a directory name Python cannot
import because of the hyphens).  `import qsi_b200` loads that directory as the package `qsi_b200`, so
`qsi_b200.bfp_ops` is the drop-in for the reference's `transformers.bfp.bfp_ops`."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "quantization-sparsity-interplay_b200")
_spec = importlib.util.spec_from_file_location("qsi_b200", os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["qsi_b200"] = _mod
_spec.loader.exec_module(_mod)
