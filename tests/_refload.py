"""Loads the reference's bfp_ops.py (pure torch) as a stand-alone module, for differential tests and for
generating golden fixtures.  Looks in /root/reference (build container) and baseline/_ref (git-ignored copy that
travels to the GPU box).  Returns None when neither exists -- callers must then skip."""
import importlib.util
import os
import sys
import types

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_CANDIDATES = ["/root/reference/src/transformers/bfp", os.path.join(_REPO, "baseline", "_ref")]
_cached = False
_ref = None


def load_reference():
    global _cached, _ref
    if _cached:
        return _ref
    _cached = True
    for root in _CANDIDATES:
        if os.path.exists(os.path.join(root, "bfp_ops.py")) and os.path.exists(os.path.join(root, "int_ops.py")):
            pkg = types.ModuleType("refbfp")
            pkg.__path__ = [root]
            sys.modules["refbfp"] = pkg
            mods = {}
            for name in ("int_ops", "bfp_ops"):
                spec = importlib.util.spec_from_file_location(f"refbfp.{name}", os.path.join(root, f"{name}.py"))
                m = importlib.util.module_from_spec(spec)
                sys.modules[f"refbfp.{name}"] = m
                spec.loader.exec_module(m)
                mods[name] = m
            _ref = mods["bfp_ops"]
            return _ref
    return None


def ref_args(ref, **kw):
    base = dict(num_format="bfp", sparsity_num_format="bfp", epsilon=1e-8, weight_mant_bits=15, in_sparsity=False,
                grad_sparsity=False, sparsity_frac=0.5, N=2, M=4, sparsity_mode="structured", device="cpu",
                rounding_mode="determ", w_sparsity=True, first="s", mant_bits=7, block_size=64)
    base.update(kw)
    return ref.unpack_bfp_args(base)
