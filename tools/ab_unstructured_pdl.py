"""A/B of programmatic dependent launch on the unstructured-sparsity pipeline (9 launches per call): option pdl = 1 / 0,
alternated, best and median of 5 rounds of 20 calls."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from qsi_b200 import bfp_ops as ours, _lib
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n * 1e3
for shape in [(4096, 4096), (4096, 11008), (8192, 22016)]:
    for dt in (torch.float32, torch.bfloat16):
        w = (torch.randn(*shape, device="cuda") * 0.02).to(dt)
        res = {0: [], 1: []}
        for rnd in range(5):
            for pdl in (1, 0):
                _lib.set_option("pdl", pdl)
                res[pdl].append(t(lambda: ours._unstructured_sparsity(w, "cuda", 0.5)))
        _lib.set_option("pdl", 1)
        f = lambda v: f"best {min(v):.1f} median {sorted(v)[len(v)//2]:.1f} us"
        print(f"{shape} {str(dt)[6:]}: pdl=1 {f(res[1])} | pdl=0 {f(res[0])}", flush=True)
