# Survey-time behaviour probe of the REFERENCE torch implementation on a B200 (not product code).
# Measures, for SURVEY.md section 8: torch.topk tie-breaking on CUDA vs CPU, exponent (log2) boundary behaviour on
# CUDA vs CPU, CPU-vs-CUDA agreement of the full reference path, reference timings on the GPU, int8/HBM sanity peaks.
# Usage (build session):  mkdir -p baseline/_ref && cp /root/reference/src/transformers/bfp/{bfp_ops,int_ops}.py baseline/_ref/
#                         gpurun --gpus 1 --timeout 900 -- python3 baseline/probe_ref_gpu.py   -> gpurun_out/probe_ref_gpu.json
# Local dry run of sections 1-3 without a GPU:  PROBE_DEV=cpu PROBE_STOP_AFTER_3=1 python3 baseline/probe_ref_gpu.py
import importlib.util, sys, types, time, json, itertools, os, torch
out = {}
def save():
    os.makedirs("gpurun_out", exist_ok=True); json.dump(out, open("gpurun_out/probe_ref_gpu.json", "w"), indent=1, default=str)
root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
pkg = types.ModuleType("refbfp"); pkg.__path__ = [root]; sys.modules["refbfp"] = pkg
def load(name):
    spec = importlib.util.spec_from_file_location(f"refbfp.{name}", f"{root}/{name}.py")
    m = importlib.util.module_from_spec(spec); sys.modules[f"refbfp.{name}"] = m; spec.loader.exec_module(m); return m
load("int_ops"); ops = load("bfp_ops")
dev = os.environ.get("PROBE_DEV", "cuda")
if dev == "cpu":
    torch.cuda.synchronize = lambda *a, **k: None
out["env"] = dict(torch=torch.__version__, cuda=torch.version.cuda, gpu=torch.cuda.get_device_name(0) if dev != "cpu" else "cpu",
                  ngpu=torch.cuda.device_count(), cpu_threads=torch.get_num_threads(),
                  tf32_matmul=torch.backends.cuda.matmul.allow_tf32, cc=torch.cuda.get_device_capability(0) if dev != "cpu" else None)
print(out["env"], flush=True)

# 1. topk tie table, M=4,k=2
pats = sorted({tuple(sorted(set(p)).index(v) for v in p) for p in itertools.product(range(4), repeat=4)})
t = torch.tensor(pats, dtype=torch.float32)
tab = {}
for reps in (1, 60000):
    for dt in (torch.float32, torch.float16, torch.bfloat16):
        x = t.to(dt).repeat(reps, 1).to(dev)
        _, idx = torch.topk(x.abs(), k=2, dim=1, largest=False)
        idx = idx.cpu()
        first = [tuple(sorted(i)) for i in idx[:len(pats)].tolist()]
        last = [tuple(sorted(i)) for i in idx[-len(pats):].tolist()]
        tab[f"{reps}_{dt}"] = first
        tab[f"{reps}_{dt}_last_equal"] = (first == last)
print({k:v for k,v in tab.items() if k.endswith("_last_equal")})
keys = [k for k in tab if not k.endswith("_last_equal")]
out["tie_cuda_consistent_across_dtype_and_rows"] = all(tab[k] == tab[keys[0]] for k in keys)
cuda_tab = tab[keys[0]]
_, idxc = torch.topk(t.abs(), k=2, dim=1, largest=False)
cpu_tab = [tuple(sorted(i)) for i in idxc.tolist()]
low = [tuple(sorted(sorted(range(4), key=lambda j: (p[j], j))[:2])) for p in pats]
high = [tuple(sorted(sorted(range(4), key=lambda j: (p[j], -j))[:2])) for p in pats]
out["tie_cuda_eq_lowest_index_rule"] = cuda_tab == low
out["tie_cuda_eq_highest_index_rule"] = cuda_tab == high
out["tie_cuda_eq_cpu"] = cuda_tab == cpu_tab
out["tie_table_M4_k2"] = [dict(pattern=p, cuda=c, cpu=u, lowest=l) for p, c, u, l in zip(pats, cuda_tab, cpu_tab, low)
                          if sorted(p)[1] == sorted(p)[2]]
print("tie: cuda==lowest", out["tie_cuda_eq_lowest_index_rule"], "cuda==highest", out["tie_cuda_eq_highest_index_rule"],
      "cuda==cpu", out["tie_cuda_eq_cpu"], "consistent", out["tie_cuda_consistent_across_dtype_and_rows"], flush=True)
for r in out["tie_table_M4_k2"]: print(r)
# other (N,M): 1:4, 3:4 (k=3,1), 4:8 (k=4), 2:8, 1:2
def rule_check(M, k, vals):
    P = list(itertools.product(vals, repeat=M)); x = torch.tensor(P, dtype=torch.float32, device=dev)
    _, idx = torch.topk(x.abs(), k=k, dim=1, largest=False); idx = idx.cpu().tolist()
    lo = sum(set(i) == set(sorted(range(M), key=lambda j: (p[j], j))[:k]) for p, i in zip(P, idx))
    hi = sum(set(i) == set(sorted(range(M), key=lambda j: (p[j], -j))[:k]) for p, i in zip(P, idx))
    return dict(M=M, k=k, n=len(P), lowest=lo, highest=hi)
out["tie_other"] = [rule_check(4, 1, [0, 1, 2]), rule_check(4, 3, [0, 1, 2]), rule_check(8, 4, [0, 1]), rule_check(8, 6, [0, 1]), rule_check(2, 1, [0, 1, 2]), rule_check(16, 8, [0, 1])]
print(out["tie_other"], flush=True)

save()
# 2. exponent: CUDA vs CPU of reference get_exponent
rows = []
for k in range(-30, 31):
    x = torch.tensor([[2.0 ** k]])
    for u in range(0, 9):
        ec = ops.get_exponent(x, 1e-8).item(); eg = ops.get_exponent(x.to(dev), 1e-8).item()
        rows.append((k, u, ec, eg))
        x = torch.nextafter(x, torch.tensor([[float("inf")]]))
out["exp_boundary_mismatch"] = [r for r in rows if r[2] != r[3]]
out["exp_boundary_nonideal_cuda"] = [r for r in rows if r[3] != (r[0] + (1 if r[1] > 0 else 0))][:80]
print("exp boundary cpu!=cuda:", out["exp_boundary_mismatch"], flush=True)
print("exp boundary cuda non-ideal (k,ulps,cpu,cuda):", out["exp_boundary_nonideal_cuda"], flush=True)
g = torch.Generator().manual_seed(1)
x = (torch.randn(1 << 24, 1, generator=g).abs() * torch.exp2(torch.randint(-20, 20, (1 << 24, 1), generator=g).float()))
ec = ops.get_exponent(x, 1e-8); eg = ops.get_exponent(x.to(dev), 1e-8).cpu()
out["exp_random_mismatch"] = int((ec != eg).sum()); print("exp random 16.7M mismatches:", out["exp_random_mismatch"], flush=True)
xh = x.half(); xh = xh[(xh > 0).squeeze() & torch.isfinite(xh).squeeze()]
ech = ops.get_exponent(xh, 1e-8); egh = ops.get_exponent(xh.to(dev), 1e-8).cpu()
out["exp_random_mismatch_fp16"] = int((ech != egh).sum()); print("exp fp16 mismatches:", out["exp_random_mismatch_fp16"], "of", xh.numel(), flush=True)
z = torch.zeros(1, 8, device=dev)
out["zero_block"] = {str(dt): ops._convert_blocked_float_to_bfp(z.to(dt), 7, 1e-8, "determ", dev).float().cpu().tolist() for dt in (torch.float32, torch.float16, torch.bfloat16)}
print("zero block:", out["zero_block"], flush=True)

save()
# 3. full path CPU vs CUDA
base = dict(num_format="bfp", sparsity_num_format="bfp", epsilon=1e-8, weight_mant_bits=15, in_sparsity=False, grad_sparsity=False,
            sparsity_frac=0.5, N=2, M=4, sparsity_mode="structured")
def run(w, device, ident="w", **kw):
    a = ops.unpack_bfp_args(dict(base, device=device, **kw)); return ops.float_to_bfp_blocked(w, **a, identifier=ident)
torch.manual_seed(0); w = torch.randn(4096, 4096) * 0.02; wg = w.to(dev)
cmp = {}
for mb in (3, 5, 7):
    for bs in (16, 64):
        for first in ("s", "q"):
            kw = dict(mant_bits=mb, block_size=bs, rounding_mode="determ", w_sparsity=True, first=first)
            a = run(w, "cpu", **kw); b = run(wg, dev, **kw).cpu()
            neq = int((a != b).sum()); bits = int((a.view(torch.int32) != b.view(torch.int32)).sum())
            cmp[f"mb{mb}_b{bs}_{first}"] = dict(value_mismatch=neq, bit_mismatch=bits, zeros_cpu=float((a == 0).float().mean()), zeros_cuda=float((b == 0).float().mean()))
out["cpu_vs_cuda_fullpath"] = cmp; print(json.dumps(cmp, indent=0), flush=True)

save()
# 4. timings of the reference path on the GPU
if os.environ.get("PROBE_STOP_AFTER_3"): save(); print("STOP"); sys.exit(0)
def bench(fn, iters=5):
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(iters):
        torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    return min(ts)
tim = {}
for shape in ((4096, 4096), (4096, 11008), (11008, 4096)):
    W = torch.randn(*shape, device=dev)
    cases = {"q_only_determ": dict(mant_bits=7, block_size=64, rounding_mode="determ", w_sparsity=False, first="s"),
             "q_only_stoc": dict(mant_bits=7, block_size=64, rounding_mode="stoc", w_sparsity=False, first="s"),
             "s_only_2:4": dict(mant_bits=7, block_size=64, rounding_mode="determ", w_sparsity=True, first="s", sparsity_num_format="fp32"),
             "s_then_q": dict(mant_bits=7, block_size=64, rounding_mode="determ", w_sparsity=True, first="s"),
             "q_then_s": dict(mant_bits=7, block_size=64, rounding_mode="determ", w_sparsity=True, first="q"),
             "s_then_q_hbfp4_b16": dict(mant_bits=3, block_size=16, rounding_mode="determ", w_sparsity=True, first="s")}
    for name, kw in cases.items():
        dt = bench(lambda: run(W, dev, **kw)); tim[f"{shape}_{name}"] = dict(ms=dt * 1e3, GBps_8B_per_elt=W.numel() * 8 / dt / 1e9)
        print(shape, name, f"{dt*1e3:.2f} ms {W.numel()*8/dt/1e9:.1f} GB/s", flush=True)
    for dt_ in (torch.float16, torch.bfloat16):
        Wh = W.to(dt_); kw = cases["s_then_q"]; d = bench(lambda: run(Wh, dev, **kw))
        tim[f"{shape}_s_then_q_{dt_}"] = dict(ms=d * 1e3, GBps_4B_per_elt=W.numel() * 4 / d / 1e9); print(shape, dt_, f"{d*1e3:.2f} ms", flush=True)
W = torch.randn(4096, 4096, device=dev)
kwu = dict(mant_bits=7, block_size=64, rounding_mode="determ", w_sparsity=True, first="s", sparsity_mode="unstructured")
d = bench(lambda: run(W, dev, **kwu), iters=3); tim["(4096,4096)_unstructured_s_then_q"] = dict(ms=d * 1e3); print("unstructured", d * 1e3, "ms", flush=True)
out["ref_gpu_timing"] = tim
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    run(W, dev, mant_bits=7, block_size=64, rounding_mode="determ", w_sparsity=True, first="s"); torch.cuda.synchronize()
tb = prof.key_averages().table(sort_by="cuda_time_total", row_limit=18, max_name_column_width=60); print(tb, flush=True); out["ref_gpu_profile"] = tb

save()
# 5. BFP linear forward as the reference runs it (quantize x, quantize w, fp32 F.linear) and GEMM-only
lin = {}
torch.backends.cuda.matmul.allow_tf32 = False
for (M_, N_, K_) in ((4096, 4096, 4096), (4096, 11008, 4096), (4096, 4096, 11008), (4096, 5120, 5120), (4096, 13824, 5120), (4096, 5120, 13824)):
    X = torch.randn(M_, K_, device=dev); Wt = torch.randn(N_, K_, device=dev) * 0.02
    args = dict(base, device=dev, mant_bits=7, block_size=64, rounding_mode="determ", w_sparsity=True, first="s")
    L = ops.BFPLinear(K_, N_, bias=False, **dict(args)).to(dev)
    with torch.no_grad():
        d_all = bench(lambda: L(X), iters=3)
        Xq = run(X, dev, ident="in", **{k: args[k] for k in ("mant_bits", "block_size", "rounding_mode", "w_sparsity", "first")}); Wq = run(Wt, dev, **{k: args[k] for k in ("mant_bits", "block_size", "rounding_mode", "w_sparsity", "first")})
        d_mm = bench(lambda: torch.nn.functional.linear(Xq, Wq), iters=5)
        torch.backends.cuda.matmul.allow_tf32 = True; d_tf = bench(lambda: torch.nn.functional.linear(Xq, Wq), iters=5); torch.backends.cuda.matmul.allow_tf32 = False
        d_bf = bench(lambda: torch.nn.functional.linear(Xq.bfloat16(), Wq.bfloat16()), iters=5)
    fl = 2 * M_ * N_ * K_
    lin[f"{M_}x{N_}x{K_}"] = dict(bfplinear_ms=d_all * 1e3, fp32_gemm_ms=d_mm * 1e3, fp32_TFLOPS=fl / d_mm / 1e12, tf32_TFLOPS=fl / d_tf / 1e12, bf16_incl_cast_ms=d_bf * 1e3)
    print(M_, N_, K_, lin[f"{M_}x{N_}x{K_}"], flush=True)
out["ref_linear"] = lin
save()
# 6. int8 library peak (denominator sanity): torch._int_mm 8192^3
try:
    A = torch.randint(-127, 127, (8192, 8192), dtype=torch.int8, device=dev); B = torch.randint(-127, 127, (8192, 8192), dtype=torch.int8, device=dev).t()
    d = bench(lambda: torch._int_mm(A, B), iters=10); out["int8_int_mm_8192_TOPS"] = 2 * 8192 ** 3 / d / 1e12; print("int8 _int_mm TOPS", out["int8_int_mm_8192_TOPS"], flush=True)
except Exception as e:
    out["int8_int_mm_error"] = repr(e); print("int_mm failed", e, flush=True)
# HBM copy sanity
a = torch.empty(1 << 29, dtype=torch.float32, device=dev); b = torch.empty_like(a)
d = bench(lambda: b.copy_(a), iters=10); out["copy_fp32_2GiB_GBps"] = a.numel() * 8 / d / 1e9; print("copy GB/s", out["copy_fp32_2GiB_GBps"], flush=True)
os.makedirs("gpurun_out", exist_ok=True); json.dump(out, open("gpurun_out/probe_ref_gpu.json", "w"), indent=1, default=str)
print("DONE")
