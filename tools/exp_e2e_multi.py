"""Why does the host-buffer (e2e) leg not scale with the number of GPUs?  Run under torchrun with N ranks.  Phases:
  solo   each rank in turn runs the N > 1 leg of bench.py (7 half-tensors of one LLaMA-65B layer through the public API) while the others idle
  all    every rank runs it at the same time
  raw    plain cudaMemcpyAsync H2D + D2H of the same bytes on two streams, every rank at once (the link / host ceiling, no library)
Prints per-rank GB/s per phase and per-call times.
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/exp_e2e_multi.py"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qsi_b200 import dist as qd  # noqa: E402

rank, local_rank, world = qd.env_world()
if "nobind" not in sys.argv:
    qd.bind_to_gpu_numa(local_rank)
import torch  # noqa: E402
from qsi_b200 import bfp_ops  # noqa: E402

torch.cuda.set_device(local_rank)
dev = torch.device("cuda", local_rank)
qd.init("nccl")
os.environ["BFP_TIE_RULE"] = "cuda"
shapes = qd.LAYER_SHAPES["llama-65b"]
gen = torch.Generator().manual_seed(70 + rank)
host_in = [(torch.randn(n // 2, k, generator=gen) * 0.02).pin_memory() for n, k in shapes]
args = bfp_ops.unpack_bfp_args(dict(num_format="bfp", sparsity_num_format="bfp", rounding_mode="determ", epsilon=1e-8, w_sparsity=True, N=2, M=4,
                                    sparsity_mode="structured", device="cuda", mant_bits=7, block_size=64, first="s"))
nbytes = sum(w.numel() * 8 for w in host_in)


def step(times=None):
    last = None
    for w in host_in:
        t0 = time.perf_counter()
        last = bfp_ops.float_to_bfp_blocked(w, **args, identifier="w")
        if times is not None:
            times.append((tuple(w.shape), (time.perf_counter() - t0) * 1e3))
    return last


y = None
for _ in range(3):
    y = step()
torch.cuda.synchronize()


def run(tag, active):
    global y
    qd.barrier(dev)
    times = []
    t0 = time.perf_counter()
    if active:
        for _ in range(4):
            y = step(times)
    dt = time.perf_counter() - t0
    qd.barrier(dev)
    if active:
        per = {}
        for s, t in times:
            per.setdefault(s, []).append(t)
        desc = "  ".join(f"{s}: min {min(v):.1f} max {max(v):.1f} ms ({s[0] * s[1] * 8 / min(v) / 1e6:.0f} GB/s best)" for s, v in per.items())
        print(f"[{tag}] rank {rank}: {4 * nbytes / dt / 1e9:.1f} GB/s   {desc}", flush=True)


for r in range(world):
    run(f"solo{r}", rank == r)
run("all", True)
run("all2", True)

# raw copies: same bytes, two streams, no kernels
dbuf = [torch.empty_like(w, device=dev) for w in host_in]
hout = [torch.empty_like(w).pin_memory() for w in host_in]
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
for phase in ("raw_warm", "raw_all"):
    qd.barrier(dev)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(4):
        for w, d, h in zip(host_in, dbuf, hout):
            with torch.cuda.stream(s1):
                d.copy_(w, non_blocking=True)
            with torch.cuda.stream(s2):
                h.copy_(d, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    qd.barrier(dev)
    print(f"[{phase}] rank {rank}: {4 * nbytes / dt / 1e9:.1f} GB/s (H2D + D2H concurrently, bytes of both directions)", flush=True)
for r in range(world):
    qd.barrier(dev)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    if rank == r:
        for _ in range(4):
            for w, d, h in zip(host_in, dbuf, hout):
                with torch.cuda.stream(s1):
                    d.copy_(w, non_blocking=True)
                with torch.cuda.stream(s2):
                    h.copy_(d, non_blocking=True)
        torch.cuda.synchronize()
        print(f"[raw_solo{r}] rank {rank}: {4 * nbytes / (time.perf_counter() - t0) / 1e9:.1f} GB/s", flush=True)
    qd.barrier(dev)
qd.shutdown()
