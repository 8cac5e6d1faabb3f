"""torchrun --nproc-per-node G tools/check_column_parallel.py : column-parallel BFP linear (65B shapes) on G GPUs.
Checks bit-equality with the single-GPU result on rank 0 and times local GEMM vs all-gather (max over ranks)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from qsi_b200 import dist as qd, bfp_ops as ops
rank, local_rank, world = qd.init("nccl")
torch.cuda.set_device(local_rank); dev = torch.device("cuda", local_rank)
kw = dict(num_format="bfp", sparsity_num_format="bfp", rounding_mode="determ", epsilon=1e-8, mant_bits=7, block_size=64,
          w_sparsity=True, N=2, M=4, first="s", sparsity_mode="structured", device="cuda")
T = 4096
res = []
SHAPES = (("65b q_proj", (8192, 8192)), ("65b up_proj", (22016, 8192)), ("65b down_proj", (8192, 22016)))
DT = torch.float32
for flag, dt in (("--fp16", torch.float16), ("--bf16", torch.bfloat16)):
    if flag in sys.argv:
        sys.argv.remove(flag); DT = dt
if "--small" in sys.argv:                      # quick functional check (tests/test_dist_gpu.py)
    sys.argv.remove("--small"); T = 520
    SHAPES = (("small even", (1024, 512)), ("small uneven shards", (1097, 640)))
for name, (N, K) in SHAPES:
    g = torch.Generator(device=dev).manual_seed(5)
    w = (torch.randn(N, K, device=dev, generator=g) * 0.02).to(DT)
    x = torch.randn(T, K, device=dev, generator=g).to(DT)
    cp = qd.ColumnParallelBFPLinear(K, N, bias=False, **dict(kw)).to(dev).to(DT).eval().load_full(w)
    with torch.no_grad():
        y = cp(x)
        full = ops.BFPLinear(K, N, bias=False, **dict(kw)).to(dev).to(DT).eval()
        full.weight.copy_(w)
        y_ref = full(x)
        ok = torch.equal(y, y_ref)
        # timing: whole forward, and the local part only
        def timed(fn, n=10):
            for _ in range(3): fn()
            torch.cuda.synchronize(); qd.barrier(dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True); e0.record()
            for _ in range(n): fn()
            e1.record(); torch.cuda.synchronize()
            return qd.max_over_ranks(e0.elapsed_time(e1) / n, dev)
        fused = cp._path == 'fused'
        ms_fwd = timed(lambda: cp(x)); ms_alias = timed(lambda: cp(x, alias_output=True)); ms_local = timed(lambda: cp.local(x)); ms_full = timed(lambda: full(x))
        os.environ["BFP_COLUMN_PARALLEL"] = "nccl"
        y_nccl = cp(x); ok_nccl = torch.equal(y_nccl, y_ref)
        ms_nccl = timed(lambda: cp(x))
        os.environ["BFP_COLUMN_PARALLEL"] = "fused"
    row = dict(layer=name, dtype=str(DT)[6:], N=N, K=K, T=T, world=world, fused_path=bool(fused), fused_failed=cp._fused_failed, bit_equal_to_single_gpu=ok,
               nccl_path_bit_equal=ok_nccl, fwd_ms=ms_fwd, fwd_alias_ms=ms_alias, nccl_fwd_ms=ms_nccl, local_ms=ms_local, single_gpu_ms=ms_full,
               tops=2.0 * T * N * K / ms_alias / 1e9, speedup_vs_1gpu=ms_full / ms_alias, fused_vs_nccl=ms_nccl / ms_alias)
    res.append(row)
    if rank == 0: print(json.dumps(row), flush=True)
    del w, x, cp, full, y, y_ref
if rank == 0 and len(sys.argv) > 1: json.dump(res, open(sys.argv[1], "w"), indent=1)
dist.destroy_process_group()
