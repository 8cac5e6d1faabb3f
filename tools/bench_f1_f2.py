"""Throughput of the next-row formats: unstructured magnitude sparsity (f1) and the INT per-channel quantiser (f2), vs the
reference's torch-CUDA implementation on the same GPU when its sources are present."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from qsi_b200 import bfp_ops as ours, _lib
from _refload import load_reference
ref = load_reference()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n * 1e3
base = dict(num_format="bfp", rounding_mode="determ", epsilon=1e-8, mant_bits=7, block_size=64, w_sparsity=True, N=2, M=4, first="s", sparsity_frac=0.5, device="cuda")
for shape in [(4096, 4096), (4096, 11008)]:
    w = torch.randn(*shape, device="cuda") * 0.02; nb = w.numel() * 8
    for name, kw in (("unstructured 50% + HBFP8 (s->q)", dict(base, sparsity_num_format="bfp", sparsity_mode="unstructured")),
                     ("unstructured 50% only", dict(base, sparsity_num_format="fp32", sparsity_mode="unstructured")),
                     ("INT8 per-channel + 2:4", dict(base, sparsity_num_format="int", sparsity_mode="structured")),
                     ("INT8 per-channel only", dict(base, sparsity_num_format="int", sparsity_mode="structured", w_sparsity=False))):
        a = ours.unpack_bfp_args(dict(kw))
        us = t(lambda: ours.float_to_bfp_blocked(w, **a, identifier="w"))
        line = f"{shape} {name}: {us:.1f} us = {nb/us/1e3:.0f} GB/s (8 B/elt algorithmic)"
        if ref is not None:
            ar = ref.unpack_bfp_args(dict(kw))
            y = ours.float_to_bfp_blocked(w, **a, identifier="w"); yr = ref.float_to_bfp_blocked(w, **ar, identifier="w")
            usr = t(lambda: ref.float_to_bfp_blocked(w, **ar, identifier="w"), 3)
            line += f" | reference torch-CUDA {usr/1e3:.2f} ms (x{usr/us:.0f}), equal {torch.equal(y, yr)}"
        print(line, flush=True)
