"""B200-native (sm_100a) implementation of the block-floating-point + N:M sparsity hot path of
parsa-epfl/quantization-sparsity-interplay (src/transformers/bfp/bfp_ops.py).

Layout:
  csrc/            hand-written CUDA kernels + the C ABI (include/bfp_b200.h) -> libbfp_b200.so
  _lib.py          ctypes binding of the C ABI (no torch types cross the boundary)
  bfp_ops.py       host-side mirror of the reference module: same names, arguments and error behaviour
  mx_layers.py     host-side mirror of the reference's mx_layers.py (MXLinear / MXConv2d / MXMatmul over the MX formats)
  dist.py          multi-GPU partitioning of the compression pass / column-parallel linear

There is no CPU fallback and no dependency on oracle/: every compute entry point fails loudly when the CUDA library or
a B200 is missing.
"""
from . import _lib  # noqa: F401

__all__ = ["_lib", "bfp_ops", "mx_layers", "install_as_reference_module"]


def install_as_reference_module():
    """Registers this package's bfp_ops as `transformers.bfp.bfp_ops`, the name the reference's patched OPT / LLaMA /
    ViT model files import (modeling_opt.py:42, modeling_llama.py:65, modeling_vit.py:41)."""
    import sys
    import types
    from . import bfp_ops, mx_layers
    pkg = sys.modules.get("transformers.bfp")
    if pkg is None:
        pkg = types.ModuleType("transformers.bfp")
        pkg.__path__ = []
        sys.modules["transformers.bfp"] = pkg
    pkg.bfp_ops = bfp_ops
    sys.modules["transformers.bfp.bfp_ops"] = bfp_ops
    pkg.mx_layers = mx_layers                       # modeling_opt.py:44, modeling_llama.py:67, modeling_vit.py:43
    sys.modules["transformers.bfp.mx_layers"] = mx_layers
    return bfp_ops
