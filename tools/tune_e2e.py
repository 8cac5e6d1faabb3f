"""bench.py's e2e leg (36 float_to_bfp_blocked calls on pinned CPU tensors per step) under different host-pipeline chunk schedules."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from qsi_b200 import _lib, bfp_ops
os.environ["BFP_TIE_RULE"] = "cuda"
SHAPES = [(4096, 4096), (4096, 11008)]
host_in = {s: (torch.randn(*s, generator=torch.Generator().manual_seed(7)) * 0.02).pin_memory() for s in SHAPES}
args = bfp_ops.unpack_bfp_args(dict(num_format="bfp", sparsity_num_format="bfp", rounding_mode="determ", epsilon=1e-8, w_sparsity=True, N=2, M=4,
                                    sparsity_mode="structured", device="cuda"))
cfgs = [(m, b, o) for m in (3, 5, 7) for b in (16, 32, 64) for o in ("s", "q")]
bytes_step = sum(s[0] * s[1] * 8 for s in SHAPES) * len(cfgs)


def step():
    last = None
    for (m, b, o) in cfgs:
        for s in SHAPES:
            last = bfp_ops.float_to_bfp_blocked(host_in[s], **dict(args, mant_bits=m, block_size=b, first=o), identifier="w")
    return last


for mx, mn in ((8, 8), (16, 1), (16, 16), (32, 32), (4, 4), (8, 8)):
    _lib.set_option("host_chunk_bytes", mx << 20); _lib.set_option("host_chunk_min_bytes", mn << 20)
    step(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(3): y = step()
    dt = (time.perf_counter() - t0) / 3
    print(f"chunk max {mx} min {mn} MiB: {dt*1e3:.1f} ms/step  {bytes_step/dt/1e9:.1f} GB/s", flush=True)
