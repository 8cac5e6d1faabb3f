"""Which part of bench.py's process state depresses the host-buffer (e2e) leg?  Runs the 36-call sweep through the public API after
enabling one suspect at a time: numa (NVML + sched_setaffinity), nccl (process group of one), sampler (10 Hz nvidia-smi started
and stopped), bigmem (26 GB of device tensors), sweep (a second of kernel launches first)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
flags = set(sys.argv[1:])
from qsi_b200 import dist as qd
if "numa" in flags: qd.bind_to_gpu_numa(0)
import torch
from qsi_b200 import _lib, bfp_ops
torch.cuda.set_device(0)
if "nccl" in flags:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29533")
    os.environ.setdefault("RANK", "0"); os.environ.setdefault("WORLD_SIZE", "1")
    qd.init("nccl")
    qd.barrier(torch.device("cuda", 0))
if "sampler" in flags:
    import bench
    s = bench.ClockSampler(0); s.start(); time.sleep(1.0); s.stop(0, time.time(), "x")
keep = None
if "bigmem" in flags:
    keep = [torch.empty(1 << 30, device="cuda") for _ in range(6)]
if "sweep" in flags:
    x = torch.randn(4096, 11008, device="cuda")
    a = bfp_ops.unpack_bfp_args(dict(num_format="bfp", sparsity_num_format="bfp", rounding_mode="determ", epsilon=1e-8, w_sparsity=True, N=2, M=4, sparsity_mode="structured", device="cuda", mant_bits=7, block_size=64))
    t0 = time.time()
    while time.time() - t0 < 1.0:
        bfp_ops.float_to_bfp_blocked(x, **a, identifier="w")
    torch.cuda.synchronize()
os.environ["BFP_TIE_RULE"] = "cuda"
SHAPES = [(4096, 4096), (4096, 11008)]
host_in = [(torch.randn(*s, generator=torch.Generator().manual_seed(7)) * 0.02).pin_memory() for s in SHAPES]
args = bfp_ops.unpack_bfp_args(dict(num_format="bfp", sparsity_num_format="bfp", rounding_mode="determ", epsilon=1e-8, w_sparsity=True, N=2, M=4,
                                    sparsity_mode="structured", device="cuda"))
cfgs = [(m, b, o) for m in (3, 5, 7) for b in (16, 32, 64) for o in ("s", "q")]
calls = [(w, m, b, o) for (m, b, o) in cfgs for w in host_in]
bytes_step = sum(w.numel() * 8 for (w, _, _, _) in calls)
def step():
    last = None
    for (w, m, b, o) in calls:
        last = bfp_ops.float_to_bfp_blocked(w, **dict(args, mant_bits=m, block_size=b, first=o), identifier="w")
    return last
for _ in range(8): step()
torch.cuda.synchronize(); t0 = time.perf_counter()
if "drop" in flags:
    for _ in range(5): step()
else:
    for _ in range(5): y = step()
dt = (time.perf_counter() - t0) / 5
print(sorted(flags), f"{dt * 1e3:.1f} ms/step  {bytes_step / dt / 1e9:.1f} GB/s", flush=True)
# per-call times
import statistics
ts = {}
for _ in range(3):
    for (w, m, b, o) in calls:
        t1 = time.perf_counter()
        r = bfp_ops.float_to_bfp_blocked(w, **dict(args, mant_bits=m, block_size=b, first=o), identifier="w")
        ts.setdefault(tuple(w.shape), []).append((time.perf_counter() - t1) * 1e3)
for k, v in ts.items():
    v.sort(); print("   per call", k, f"min {v[0]:.2f} med {statistics.median(v):.2f} p90 {v[int(len(v) * 0.9)]:.2f} max {v[-1]:.2f} ms -> med {k[0] * k[1] * 8 / statistics.median(v) / 1e6:.1f} GB/s", flush=True)
