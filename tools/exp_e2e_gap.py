"""Where does the host-buffer path lose bandwidth?  Same 36-call sweep as bench.py's e2e leg, measured as
  raw      bfp_quantize_host into PREALLOCATED pinned outputs (no Python allocation in the loop)
  api      bfp_ops.float_to_bfp_blocked(pinned CPU tensor) (allocates a pinned output per call)
with and without binding the process to the GPU's NUMA node first (argument `numa`), and the plain cudaMemcpy ceilings."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qsi_b200 import dist as qd
cpus = qd.bind_to_gpu_numa(0) if "numa" in sys.argv else None
import torch
from qsi_b200 import _lib, bfp_ops
print("numa-bound cpus:", None if cpus is None else (len(cpus), cpus[:4], "..."), "affinity now:", len(os.sched_getaffinity(0)), flush=True)
os.environ["BFP_TIE_RULE"] = "cuda"
SHAPES = [(4096, 4096), (4096, 11008)]
host_in = {s: (torch.randn(*s, generator=torch.Generator().manual_seed(7)) * 0.02).pin_memory() for s in SHAPES}
host_out = {s: torch.empty(*s).pin_memory() for s in SHAPES}
args = bfp_ops.unpack_bfp_args(dict(num_format="bfp", sparsity_num_format="bfp", rounding_mode="determ", epsilon=1e-8, w_sparsity=True, N=2, M=4,
                                    sparsity_mode="structured", device="cuda"))
cfgs = [(m, b, o) for m in (3, 5, 7) for b in (16, 32, 64) for o in ("s", "q")]
bytes_step = sum(s[0] * s[1] * 8 for s in SHAPES) * len(cfgs)
L = _lib.lib()

# plain copy ceilings
d = {s: torch.empty(*s, device="cuda") for s in SHAPES}
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
for s in SHAPES:
    for _ in range(3):
        d[s].copy_(host_in[s], non_blocking=True); host_out[s].copy_(d[s], non_blocking=True)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(10):
        with torch.cuda.stream(s1): d[s].copy_(host_in[s], non_blocking=True)
        with torch.cuda.stream(s2): host_out[s].copy_(d[s], non_blocking=True)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 10
    print(f"{s} concurrent h2d+d2h: {2 * s[0] * s[1] * 4 / dt / 1e9:.1f} GB/s", flush=True)


def raw_step():
    for (m, b, o) in cfgs:
        for s in SHAPES:
            _lib.check(L.bfp_quantize_host(host_in[s].data_ptr(), host_out[s].data_ptr(), s[0], s[1], 0, 0, b, m, 1e-8, 0, 0, 0, 2, 4,
                                           1 if o == "s" else 2, 0))


def api_step():
    last = None
    for (m, b, o) in cfgs:
        for s in SHAPES:
            last = bfp_ops.float_to_bfp_blocked(host_in[s], **dict(args, mant_bits=m, block_size=b, first=o), identifier="w")
    return last


for name, fn in (("raw", raw_step), ("api", api_step), ("raw", raw_step), ("api", api_step)):
    for _ in range(4): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(4): fn()
    dt = (time.perf_counter() - t0) / 4
    print(f"{name}: {dt * 1e3:.1f} ms/step  {bytes_step / dt / 1e9:.1f} GB/s", flush=True)
