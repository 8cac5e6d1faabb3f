#!/usr/bin/env python
"""bench.py -- headline benchmark of the BFP + N:M hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

N = 1 (BASELINE.json configs[1]): LLaMA-7B weight shapes (4096x4096, 4096x11008), fp32, quantise+sparsify sweep over
BFP4/6/8 (mant_bits 3/5/7) x block 16/32/64 x both orderings, 2:4, round-to-nearest.  One "step" = one pass of the
sweep = 36 fused-kernel launches.  Bytes are ALGORITHMIC: numel x (sizeof(in) + sizeof(out)) = 8 B/element.

N > 1 (BASELINE.json configs[4]), STRONG scaling: one "step" = one whole-model compression pass over LLaMA-65B (80 layers x 7
weight tensors = 64.76 G elements, HBFP8 block 64, 2:4, sparsify -> quantise, fp32 -> fp32), sharded by tensor -- layer l ->
rank l mod N, no data-path collective; value = 64.76 G x 8 B / max-over-ranks time.  The same pass is also timed at N = 1 and
reported as `compress_65b` in every line, so the 1 -> 8 curve of ONE workload can be read from that key.  The column-parallel
BFP linear of the same model (q / up / down_proj at 4096 tokens; fused all-gather-in-epilogue vs NCCL; bit-equality against
the single-GPU result checked on rank 0) is reported as `column_parallel`.

  value     whole-job GB/s with inputs resident in HBM (CUDA events on the launching stream, max over ranks)
  e2e       the same operator through the public API with HOST (pinned) buffers: H2D + kernel + D2H inside the timed region
  roofline  the stream kernel against the measured HBM copy peak (MEASURED_PEAKS.json); `traffic` is measured in the run by an
            ncu child process (dram bytes of one launch) when ncu is available, else null
  quant_modes   the reference's real operating modes (stochastic rounding, fp16 / bf16 tensors) on the same shapes
  unstructured  global magnitude pruning fused with the quantiser (the other sparsity mode of the reference's scripts)
  models        BASELINE configs[0] / [3]: OPT-125M and ViT-B/16 drop-ins, logits vs the unmodified reference on the same GPU, both timed
  config3_llama13b_layer  BASELINE configs[2]: the seven BFPLinear forwards of a LLaMA-2-13B layer at 4096 tokens (HBFP8, HBFP4, reference)
  gemm      BFP GEMM TOPS at the LLaMA-7B shapes, burst and sustained (>= 2 s), with its own roofline block
  cpu_baseline  the CPU implementation (reference if baseline/_ref is present, else the oracle port) on a bounded sample
"""
import argparse
import json
import os
import shutil
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "BFP+N:M quantize GB/s (% HBM peak); BFP GEMM TOPS at LLaMA-7B shapes"
_REAL_STDOUT = None            # the process's original stdout once main() has redirected fd 1 to stderr


def emit(line):
    out = _REAL_STDOUT if _REAL_STDOUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()
SHAPES = [(4096, 4096), (4096, 11008)]
MANTS = [3, 5, 7]
BLOCKS = [16, 32, 64]
ORDERS = ["s", "q"]            # first='s' (sparsify->quantise) / 'q' (quantise->sparsify)
N_, M_ = 2, 4
MODEL5 = "llama-65b"           # BASELINE.json configs[4]


def sweep_configs():
    return [(m, b, o) for m in MANTS for b in BLOCKS for o in ORDERS]


def step_bytes(shapes=SHAPES, bytes_per_elt=8):
    return sum(r * k for r, k in shapes) * bytes_per_elt * len(sweep_configs())


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    d = {"hbm_gbs": 6650.0, "bf16_tflops": None, "bf16_tflops_sustained": None, "source": "fallback (B200_PROFILING.md)"}
    if os.path.exists(p):
        j = json.load(open(p))
        d.update(hbm_gbs=float(j["hbm_gbs"]), bf16_tflops=j.get("bf16_tflops"), bf16_tflops_sustained=j.get("bf16_tflops_sustained"),
                 source="measured (MEASURED_PEAKS.json)")
    return d


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the GPU legs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None
        t_wait = time.time()
        while self.proc is not None and not self.rows and time.time() - t_wait < 4.0:   # nvidia-smi takes ~1 s to print its first row
            time.sleep(0.05)

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1, window):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for ts, line in self.rows:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                if t0 <= ts <= t1 + 0.15:
                    sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            if t0 <= ts <= t1 + 0.15:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm), "window": window}


# ---------------------------------------------------------------------------------------------------------------
# CPU arm: the reference's own implementation (or the oracle port) on the host cores
# ---------------------------------------------------------------------------------------------------------------
def cpu_arm(budget_s, threads, workload):
    """Times one bounded sample of `workload` ("sweep7b" or "pass65b") on the CPU.
    Returns (GB/s, kind, cores, sample description, seconds)."""
    import torch
    from _refload import load_reference, ref_args
    torch.set_num_threads(threads)
    ref = load_reference()
    g = torch.Generator().manual_seed(0)
    if ref is not None:
        kind = "reference"

        def run(w, m, b, o):
            return ref.float_to_bfp_blocked(w, **ref_args(ref, mant_bits=m, block_size=b, first=o), identifier="w")
    else:
        kind = "port"
        from oracle import bfp_oracle as O
        O.set_num_threads(threads)

        def run(w, m, b, o):
            return O.float_to_bfp_blocked(w.numpy(), m, b, "sq" if o == "s" else "qs", tie_rule="cpu")[0]
    if workload == "pass65b":
        from qsi_b200 import dist as qd
        shapes, cfgs = qd.LAYER_SHAPES[MODEL5], [(7, 64, "s")]
        what = "the 7 weight tensors of one LLaMA-65B layer (HBFP8 block 64, 2:4, s->q, fp32, nearest)"
    else:
        shapes, cfgs = SHAPES, sweep_configs()
        what = "the 18-config sweep on 4096x4096 and 4096x11008 (fp32, 2:4, nearest)"
    # calibrate on a small slice, then size the row sample so the whole sample fits the budget
    probe = torch.randn(256, 4096, generator=g) * 0.02
    run(probe, 7, 64, "s")
    t0 = time.perf_counter()
    run(probe, 7, 64, "s"); run(probe, 3, 16, "q")
    per_elt = (time.perf_counter() - t0) / (2 * probe.numel())
    total_elts = sum(r * k for r, k in shapes) * len(cfgs)
    # size the sample (rows of each shape, whole passes when one pass is cheap) for ~0.7 x budget of CPU work; the small
    # probe over-estimates the per-element cost, so a sample that came out under half the budget is re-sized once from its
    # own timing and re-measured
    for attempt in range(2):
        frac = min(1.0, 0.7 * budget_s / max(per_elt * total_elts, 1e-9))
        rows = [max(8, int(r * frac) // 8 * 8) for r, _ in shapes]
        ws = [torch.randn(rs, k, generator=g) * 0.02 for rs, (_, k) in zip(rows, shapes)]
        est = per_elt * sum(w.numel() for w in ws) * len(cfgs)
        passes = max(1, min(4, int(0.7 * budget_s / max(est, 1e-9))))
        t0 = time.perf_counter()
        nbytes = 0
        for _ in range(passes):
            for (m, b, o) in cfgs:
                for w in ws:
                    run(w, m, b, o)
                    nbytes += w.numel() * 8
        dt = time.perf_counter() - t0
        if dt >= 0.4 * budget_s or (frac >= 1.0 and passes >= 4):
            break
        per_elt = dt / (nbytes / 8)
    sample = f"{passes} pass(es) of {what}, first {rows} rows of each tensor, {nbytes / 1e9:.2f} GB algorithmic"
    return nbytes / dt / 1e9, kind, threads, sample, dt


def config_for(world):
    if world == 1:
        return {"workload": "llama7b_quant_sparsify_sweep", "shapes": SHAPES, "mant_bits": MANTS, "block": BLOCKS,
                "orders": ["s->q", "q->s"], "nm": "2:4", "rounding": "nearest"}
    return {"workload": "llama65b_compression_pass_sharded_by_layer", "model": MODEL5, "layers": 80, "tensors": 560,
            "elements": 64760053760, "mant_bits": 7, "block": 64, "order": "s->q", "nm": "2:4", "rounding": "nearest"}


def run_reference_arm(a, rank, world):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    budget = max(2.0, min(20.0, 150.0 / max(1, a.steps + a.warmup)))
    workload = "sweep7b" if world == 1 else "pass65b"
    for _ in range(a.warmup):
        cpu_arm(budget, threads, workload)
    vals, secs, info = [], [], None
    for _ in range(a.steps):
        v, kind, cores, sample, dt = cpu_arm(budget, threads, workload)
        vals.append(v); secs.append(dt); info = (kind, cores, sample)
    value = sum(vals) / len(vals)
    line = {"impl": "reference", "device": "cpu (the reference's torch-CPU path on the host cores; no GPU is used by this arm)", "metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": a.gpus, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": 1e3 * sum(secs) / len(secs), "higher_is_better": True,
            "scaling": "weak" if world == 1 else "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_for(world), "device": "host CPU cores (the reference arm of this tier is the reference's CPU path; no GPU is used)",
            "cpu_baseline": {"value": value, "unit": "GB/s", "cores": info[1], "kind": info[0], "sample": info[2]},
            "e2e": {"value": value, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


# ---------------------------------------------------------------------------------------------------------------
# legs of our arm
# ---------------------------------------------------------------------------------------------------------------
def timed_region(torch, qd, dev, fn, steps, warmup):
    """`warmup` untimed + `steps` timed calls of fn(i), barrier + synchronize on both sides, CUDA events on the current stream,
    max over ranks.  Returns (total ms, launches counted by the library)."""
    from qsi_b200 import _lib
    for i in range(warmup):
        fn(i)
    torch.cuda.synchronize()
    qd.barrier(dev)
    torch.cuda.synchronize()
    n0 = _lib.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for i in range(steps):
        fn(i)
    ev1.record()
    torch.cuda.synchronize()
    qd.barrier(dev)
    torch.cuda.synchronize()
    return qd.max_over_ranks(ev0.elapsed_time(ev1), dev), _lib.launch_count() - n0


def make_sweep(torch, dev, rank):
    from qsi_b200 import _lib
    L = _lib.lib()
    cfgs = sweep_configs()
    # inputs resident in HBM; rotate over several distinct buffers per shape so that nothing is re-read from L2
    # (L2 = 126 MB; between two uses of a buffer the sweep touches >= 600 MB of other data)
    g = torch.Generator(device=dev).manual_seed(1000 + rank)
    n_rot = {SHAPES[0]: 4, SHAPES[1]: 2}
    ins = {s: [torch.randn(*s, device=dev, generator=g) * 0.02 for _ in range(n_rot[s])] for s in SHAPES}
    outs = {s: [torch.empty(*s, device=dev) for _ in range(2)] for s in SHAPES}
    stream = torch.cuda.current_stream().cuda_stream

    def device_step(i):
        for ci, (m, b, o) in enumerate(cfgs):
            order = _lib.ORDER_SPARSIFY_QUANT if o == "s" else _lib.ORDER_QUANT_SPARSIFY
            for s in SHAPES:
                x = ins[s][(i * len(cfgs) + ci) % n_rot[s]]
                y = outs[s][ci % 2]
                rc = L.bfp_quantize(x.data_ptr(), y.data_ptr(), s[0], s[1], _lib.DT_F32, _lib.DT_F32, b, m, 1e-8,
                                    _lib.ROUND_NEAREST, 0, 0, N_, M_, order, _lib.TIE_TORCH_CUDA, stream)
                if rc:
                    _lib.check(rc)
    return device_step, (ins, outs)


def make_compress_pass(torch, qd, dev, rank, world):
    """One whole-model compression pass of MODEL5 on this rank's layers (layer l -> rank l % world), raw C-ABI calls into
    preallocated outputs.  Up to 4 resident layer sets (3.2 GB in + 3.2 GB out each) rotate over the rank's layers so that
    consecutive layers touch different memory (>> L2)."""
    from qsi_b200 import _lib
    L = _lib.lib()
    shapes = qd.LAYER_SHAPES[MODEL5]
    my_layers = sorted({t[0] for t in qd.shard_by_layer(qd.model_tensors(MODEL5), rank, world)})
    g = torch.Generator(device=dev).manual_seed(77 + rank)
    nsets = max(1, min(4, len(my_layers)))
    sets = [[torch.randn(n, k, device=dev, generator=g) * 0.02 for n, k in shapes] for _ in range(nsets)]
    outs = [[torch.empty_like(w) for w in ws] for ws in sets]
    stream = torch.cuda.current_stream().cuda_stream

    def one_pass(i):
        for j in range(len(my_layers)):
            s = (i * len(my_layers) + j) % nsets
            for w, o in zip(sets[s], outs[s]):
                rc = L.bfp_quantize(w.data_ptr(), o.data_ptr(), w.shape[0], w.shape[1], _lib.DT_F32, _lib.DT_F32, 64, 7, 1e-8,
                                    _lib.ROUND_NEAREST, 0, 0, N_, M_, _lib.ORDER_SPARSIFY_QUANT, _lib.TIE_TORCH_CUDA, stream)
                if rc:
                    _lib.check(rc)
    elems_rank = len(my_layers) * sum(n * k for n, k in shapes)
    return one_pass, elems_rank, len(my_layers) * len(shapes), (sets, outs)


def leg_column_parallel(torch, qd, dev, rank, world):
    """Column-parallel BFPLinear of the 65B model at T = 4096: fused (all-gather in the GEMM epilogue over peer
    memory) vs NCCL all-gather, both checked bit for bit against the single-GPU BFPLinear on rank 0."""
    from qsi_b200 import bfp_ops as ops
    kw = dict(num_format="bfp", sparsity_num_format="bfp", rounding_mode="determ", epsilon=1e-8, mant_bits=7, block_size=64,
              w_sparsity=True, N=N_, M=M_, first="s", sparsity_mode="structured", device="cuda")
    T = 4096
    rows = []

    def timed(fn, n=10):
        for _ in range(3):
            fn()
        torch.cuda.synchronize(); qd.barrier(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record(); torch.cuda.synchronize()
        return qd.max_over_ranks(e0.elapsed_time(e1) / n, dev)

    for dt_name, DT in (("float32", torch.float32), ("bfloat16", torch.bfloat16)):
        mods = {}
        for name, (N, K) in (("q_proj", (8192, 8192)), ("up_proj", (22016, 8192)), ("down_proj", (8192, 22016))):
            g = torch.Generator(device=dev).manual_seed(5)              # same full weight / input on every rank
            w = (torch.randn(N, K, device=dev, generator=g) * 0.02).to(DT)
            x = torch.randn(T, K, device=dev, generator=g).to(DT)
            cp = qd.ColumnParallelBFPLinear(K, N, bias=False, **dict(kw)).to(dev).to(DT).eval().load_full(w)
            with torch.no_grad():
                os.environ["BFP_COLUMN_PARALLEL"] = "fused"
                y = cp(x)
                fused = cp._path == "fused"
                ok = ok_nccl = None
                if rank == 0:                                           # the single-GPU result: no collective inside
                    full = ops.BFPLinear(K, N, bias=False, **dict(kw)).to(dev).to(DT).eval()
                    full.weight.copy_(w)
                    y_ref = full(x)
                    ok = bool(torch.equal(y, y_ref))
                ms_fused = timed(lambda: cp(x, alias_output=True))
                ms_local = timed(lambda: cp.local(x))
                ms_full = 0.0
                if rank == 0:                                           # local timing only: the other ranks are not in here
                    for _ in range(3):
                        full(x)
                    torch.cuda.synchronize()
                    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    f0.record()
                    for _ in range(10):
                        full(x)
                    f1.record(); torch.cuda.synchronize()
                    ms_full = f0.elapsed_time(f1) / 10
                    del full
                ms_full = qd.max_over_ranks(ms_full, dev)
                # NCCL path: a second module (the path decision is per module and collective)
                os.environ["BFP_COLUMN_PARALLEL"] = "nccl"
                cpn = qd.ColumnParallelBFPLinear(K, N, bias=False, **dict(kw)).to(dev).to(DT).eval().load_full(w)
                y_n = cpn(x)
                if rank == 0:
                    ok_nccl = bool(torch.equal(y_n, y_ref))
                ms_nccl = timed(lambda: cpn(x))
                os.environ["BFP_COLUMN_PARALLEL"] = "fused"
            es = 4 if DT == torch.float32 else 2
            ingress = (world - 1) / world * T * N * es
            rows.append({"proj": name, "dtype": dt_name, "N": N, "K": K, "T": T, "path": cp._path, "fused_ruled_out": cp._fused_failed,
                         "bit_equal_to_single_gpu": ok, "nccl_bit_equal_to_single_gpu": ok_nccl,
                         "fused_ms": ms_fused, "nccl_ms": ms_nccl, "local_gemm_ms": ms_local, "single_gpu_ms": ms_full,
                         "fused_tops": 2.0 * T * N * K / ms_fused / 1e9, "speedup_vs_single_gpu": ms_full / ms_fused,
                         "fused_vs_nccl": ms_nccl / ms_fused,
                         "nvlink_ingress_GBps": ingress / (ms_fused * 1e-3) / 1e9, "nvlink_ingress_frac_of_900": ingress / (ms_fused * 1e-3) / 1e9 / 900.0})
            if name == "q_proj":
                mods = (cp, x, w, K, N)
            del cpn, y, y_n
            if name != "q_proj":
                del cp, w, x
        # q / k / v as one unit: one barrier pair, GEMMs back to back, the peer stores of one projection under the next GEMM
        cp, x, w, K, N = mods
        with torch.no_grad():
            sib = [cp] + [qd.ColumnParallelBFPLinear(K, N, bias=False, **dict(kw)).to(dev).to(DT).eval().load_full(w) for _ in range(2)]
            ys = qd.column_parallel_group_forward(sib, x)
            same = bool(all(torch.equal(ys[0], yy) for yy in ys[1:]))
            ms_group = timed(lambda: qd.column_parallel_group_forward(sib, x))
            ms_each = timed(lambda: [m(x, alias_output=True) for m in sib])
        es = 4 if DT == torch.float32 else 2
        rows.append({"proj": "q+k+v as a group", "dtype": dt_name, "N": N, "K": K, "T": T, "path": cp._path, "outputs_identical": same,
                     "group_ms": ms_group, "one_by_one_ms": ms_each, "per_proj_ms": ms_group / 3,
                     "nvlink_ingress_GBps": 3 * (world - 1) / world * T * N * es / (ms_group * 1e-3) / 1e9})
        del sib, ys, cp, x, w, mods
        torch.cuda.empty_cache()
    return rows


def leg_quant_modes(torch, dev, peak):
    """The reference's real operating modes (every script sets rounding_mode 'stoc', LLaMA runs in fp16): HBFP8 block 64, 2:4,
    s->q on both shapes; input buffers rotate (> L2).  GB/s = numel x (sizeof(in) + sizeof(out)) / device time."""
    from qsi_b200 import _lib
    L = _lib.lib()
    stream = torch.cuda.current_stream().cuda_stream
    dts = {"f32": (torch.float32, _lib.DT_F32), "bf16": (torch.bfloat16, _lib.DT_BF16), "f16": (torch.float16, _lib.DT_F16)}
    out = []
    for shape in SHAPES:
        for dname, (tdt, cdt) in dts.items():
            xs = [(torch.randn(*shape, device=dev) * 0.02).to(tdt) for _ in range(4)]
            for rounding, rname in ((_lib.ROUND_NEAREST, "nearest"), (_lib.ROUND_STOCHASTIC, "stochastic")):
                odt_t, odt_c = (torch.float32, _lib.DT_F32) if rounding == _lib.ROUND_STOCHASTIC else (tdt, cdt)
                ys = [torch.empty(*shape, device=dev, dtype=odt_t) for _ in range(2)]

                def call(i):
                    rc = L.bfp_quantize(xs[i % 4].data_ptr(), ys[i % 2].data_ptr(), shape[0], shape[1], cdt, odt_c, 64, 7, 1e-8, rounding,
                                        1234, i, N_, M_, _lib.ORDER_SPARSIFY_QUANT, _lib.TIE_TORCH_CUDA, stream)
                    if rc:
                        _lib.check(rc)
                for i in range(5):
                    call(i)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                iters = 40
                for i in range(iters):
                    call(i)
                e1.record(); torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / iters
                bpe = xs[0].element_size() + ys[0].element_size()
                gbs = shape[0] * shape[1] * bpe / (ms * 1e-3) / 1e9
                out.append({"shape": list(shape), "dtype": dname, "rounding": rname, "bytes_per_element": bpe, "us": ms * 1e3,
                            "GBps": gbs, "frac_of_hbm_peak": gbs / peak})
                del ys
            del xs
    return out


def leg_unstructured(torch, dev, peak):
    """Global (unstructured) magnitude pruning at 50 % fused with HBFP8 block 64 -- the sparsity mode of four of the reference's
    seven LM scripts (bfp_ops.py:61-71 + :46-59) -- through the C ABI: the two-read pipeline (csrc/bfp_unstructured_fused.cu).
    GB/s = numel x (sizeof(in) + sizeof(out)) / device time; the kernels move 2 reads + 1 write, so 2/3 of the HBM peak is the
    ceiling of this figure for equal in/out widths."""
    from qsi_b200 import _lib
    L = _lib.lib()
    stream = torch.cuda.current_stream().cuda_stream
    dts = {"f32": (torch.float32, _lib.DT_F32), "bf16": (torch.bfloat16, _lib.DT_BF16)}
    out = []
    for shape in SHAPES:
        n = shape[0] * shape[1]
        for dname, (tdt, cdt) in dts.items():
            xs = [(torch.randn(*shape, device=dev) * 0.02).to(tdt) for _ in range(4)]
            nbytes = L.bfp_unstructured_quantize_workspace_bytes(n, cdt)
            ws = torch.empty(nbytes // 8 + 1, dtype=torch.int64, device=dev)
            for order, oname in ((_lib.ORDER_SPARSIFY_QUANT, "s->q"), (_lib.ORDER_QUANT_SPARSIFY, "q->s")):
                ys = [torch.empty(*shape, device=dev, dtype=tdt) for _ in range(2)]

                def call(i):
                    rc = L.bfp_unstructured_quantize(xs[i % 4].data_ptr(), ys[i % 2].data_ptr(), shape[0], shape[1], cdt, cdt, n // 2, order,
                                                     64, 7, 1e-8, _lib.ROUND_NEAREST, 0, 0, ws.data_ptr(), nbytes, stream)
                    if rc:
                        _lib.check(rc)
                for i in range(5):
                    call(i)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                iters = 30
                for i in range(iters):
                    call(i)
                e1.record(); torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / iters
                bpe = 2 * xs[0].element_size()
                gbs = n * bpe / (ms * 1e-3) / 1e9
                out.append({"shape": list(shape), "dtype": dname, "order": oname, "sparsity_frac": 0.5, "bytes_per_element": bpe, "us": ms * 1e3,
                            "GBps": gbs, "frac_of_hbm_peak": gbs / peak, "frac_of_two_read_ceiling": gbs / (peak * 2.0 / 3.0), "launches_per_call": 4})
                del ys
            del xs, ws
    return out


def leg_mx_formats(torch, dev, peak):
    """SURVEY section 8 row f4 (last item): the MX formats behind the reference's mx_layers.py.  The fused quantiser against the HBM
    roofline (fake-quant and the block-scaled operand form) and MXLinear forwards at the LLaMA-7B shapes (4096 tokens) on the tensor
    cores: fp8_e4m3 / fp4_e2m1 on tcgen05.mma.kind::mxf8f6f4.block_scale (an MX block IS the hardware's block), int8 on the bf16 kinds."""
    from qsi_b200 import mx_layers as mx
    g = torch.Generator(device=dev).manual_seed(31)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def timed(fn, iters=10):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(iters):
            fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters

    out = {"block": 32, "bfloat": 16, "scale_bits": 8, "quantiser": [], "linear": []}
    shape = (4096, 11008)
    bufs = [torch.randn(*shape, device=dev, generator=g) for _ in range(4)]
    n = shape[0] * shape[1]
    for fmt in ("fp8_e4m3", "int8"):
        i = [0]

        def fq():
            i[0] += 1
            return mx._mx_quantize_last(bufs[i[0] % 4], mx.ELEM_FORMATS[fmt], 32, 8, 16, False)
        us = timed(fq) * 1e3
        out["quantiser"].append({"shape": list(shape), "dtype": "f32", "format": fmt, "out": "fake-quant", "bytes_per_element": 8, "us": us,
                                 "GBps": n * 8 / us / 1e3, "frac_of_hbm_peak": n * 8 / us / 1e3 / peak})
    sp = mx.finalize_mx_specs(mx.apply_mx_specs(dict(block_size=32, bfloat=16, scale_bits=8, w_elem_format="fp8_e4m3", a_elem_format="fp8_e4m3")))
    i = [0]

    def pk():
        i[0] += 1
        return mx._pack_block_scaled(bufs[i[0] % 4], mx.ELEM_FORMATS["fp8_e4m3"], 128, sp, 16)
    us = timed(pk) * 1e3
    bpe = 4 + 1 + 1.0 / 32
    out["quantiser"].append({"shape": list(shape), "dtype": "f32", "format": "fp8_e4m3", "out": "E4M3 bytes + UE8M0 scale atoms", "bytes_per_element": bpe,
                             "us": us, "GBps": n * bpe / us / 1e3, "frac_of_hbm_peak": n * bpe / us / 1e3 / peak})
    del bufs
    T = 4096
    for fmt in ("fp8_e4m3", "fp4_e2m1", "int8"):
        ms_sum, ops = 0.0, 0.0
        for (N, K) in ((4096, 4096), (11008, 4096), (4096, 11008)):
            x = torch.randn(T, K, device=dev, generator=g)
            lin = mx.MXLinear(K, N, bias=False, mx_specs=dict(block_size=32, bfloat=16, scale_bits=8, w_elem_format=fmt, a_elem_format=fmt)).to(dev).eval()
            with torch.no_grad():
                ms_sum += timed(lambda: lin(x))
            ops += 2.0 * T * N * K
            del lin, x
        out["linear"].append({"format": fmt, "tokens": T, "shapes": "llama-7b", "ms_three_forwards": ms_sum, "tflops": ops / ms_sum / 1e9,
                              "includes": "activation quantise + pack, GEMM, bfloat output rounding (weight pack cached)",
                              "kind": "tcgen05.mma.kind::mxf8f6f4.block_scale" if fmt != "int8" else "tcgen05.mma.kind::f16 on exact-bf16 MX values"})
    return out


def leg_models(torch, dev):
    """BASELINE.json configs[0] and configs[3] at the model level, on this GPU: stock `transformers` OPT-125M (8 x 512 tokens, HBFP8 + 2:4)
    and ViT-B/16 (batch 256, BFP6 + 2:4), random init, every block nn.Linear swapped for BFPLinear (and the ViT patch embedding for
    BFPConv2d) -- the substitution the reference's patched model files make (tools/model_dropin.py).  One forward each with this repo's
    bfp_ops and with the unmodified reference's (baseline/_ref) on the same GPU: logits compared, both timed."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import model_dropin as md
    from qsi_b200 import bfp_ops
    from _refload import load_reference
    ref = load_reference()
    out = {}
    for name, kind, m, seq in (("config1_opt125m_8x512_hbfp8_2to4", "opt", 7, 512), ("config4_vit_b16_b256_bfp6_2to4", "vit", 5, 0)):
        kw = dict(num_format="bfp", sparsity_num_format="bfp", rounding_mode="determ", epsilon=1e-8, mant_bits=m, weight_mant_bits=15,
                  block_size=64, w_sparsity=True, N=N_, M=M_, first="s", sparsity_mode="structured", sparsity_frac=0.5, device="cuda")
        row, logits = {}, {}
        for tag, impl in (("ours", bfp_ops), ("reference_same_gpu", ref)):
            if impl is None:
                continue
            model, cfg = md.build(kind, 0)
            n = md.swap(model, impl, kw, md.OPT_TARGETS if kind == "opt" else None)
            model = model.to(dev)
            g = torch.Generator().manual_seed(1)
            inp = (dict(input_ids=torch.randint(0, cfg.vocab_size, (8, seq), generator=g).to(dev)) if kind == "opt"
                   else dict(pixel_values=torch.randn(256, 3, 224, 224, generator=g).to(dev)))
            ours = impl is bfp_ops
            with torch.no_grad():
                for _ in range(3 if ours else 1):
                    y = model(**inp).logits
                torch.cuda.synchronize()
                iters = 5 if ours else 2
                t0 = time.perf_counter()
                for _ in range(iters):
                    y = model(**inp).logits
                torch.cuda.synchronize()
                dt = (time.perf_counter() - t0) / iters
            logits[tag] = y.float()
            row[tag] = {"forward_ms": dt * 1e3, "swapped_modules": n}
            del model
        if "reference_same_gpu" in logits:
            d = logits["ours"] - logits["reference_same_gpu"]
            row["logits_rel_err_vs_reference"] = float(d.norm() / logits["reference_same_gpu"].norm())
            row["speedup_vs_reference_same_gpu"] = row["reference_same_gpu"]["forward_ms"] / row["ours"]["forward_ms"]
        out[name] = row
    return out


def leg_config3(torch, dev):
    """BASELINE.json configs[2]: LLaMA-2-13B BFP linear forward, 4096 tokens per step, on-the-fly BFP activations x 2:4-sparse BFP
    weights: the seven BFPLinear forwards of one decoder layer through the public module API (activation quantise + GEMM, packed
    weight cached), HBFP8 and HBFP4, and the unmodified reference's BFPLinear on the same GPU (one pass) when baseline/_ref is present."""
    from qsi_b200 import bfp_ops, dist as qd2
    from _refload import load_reference
    shapes, T = qd2.LAYER_SHAPES["llama-13b"], 4096
    flop = sum(2.0 * T * n * k for n, k in shapes)
    g = torch.Generator(device=dev).manual_seed(13)
    ws = [torch.randn(n, k, device=dev, generator=g) * 0.02 for n, k in shapes]
    xs = {k: torch.randn(T, k, device=dev, generator=g) for k in sorted({k for _, k in shapes})}
    out = {"model": "llama-13b", "tokens": T, "layer_shapes_N_K": [list(sh) for sh in shapes], "tflop_per_layer": flop / 1e12}
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    keep = None
    for tag, impl, m in (("hbfp8", bfp_ops, 7), ("hbfp4", bfp_ops, 3), ("reference_hbfp8_same_gpu", load_reference(), 7)):
        if impl is None:
            continue
        kw = dict(num_format="bfp", sparsity_num_format="bfp", rounding_mode="determ", epsilon=1e-8, mant_bits=m, weight_mant_bits=15, block_size=64,
                  w_sparsity=True, N=N_, M=M_, first="s", sparsity_mode="structured", sparsity_frac=0.5, device="cuda")
        lins = []
        for w in ws:
            lin = impl.BFPLinear(w.shape[1], w.shape[0], bias=False, **dict(kw)).to(dev)
            lin.weight = torch.nn.Parameter(w, requires_grad=False)
            lins.append(lin)
        ours = impl is bfp_ops
        with torch.no_grad():
            for _ in range(2 if ours else 1):
                ys = [lin(xs[lin.in_features]) for lin in lins]
            torch.cuda.synchronize()
            iters = 10 if ours else 1
            e0.record()
            for _ in range(iters):
                ys = [lin(xs[lin.in_features]) for lin in lins]
            e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        out[tag] = {"ms_per_layer": ms, "tflops": flop / ms / 1e9}
        if ours:
            out[tag]["weight_kinds"] = sorted({lin._packed_w[0][0] for lin in lins})
        if tag == "hbfp8":
            keep = [y.clone() for y in ys]
        elif tag.startswith("reference") and keep is not None:
            out["hbfp8_max_rel_err_vs_reference"] = max(float((a_ - b_).norm() / b_.norm()) for a_, b_ in zip(keep, ys))
            out["hbfp8_speedup_vs_reference_same_gpu"] = ms / out["hbfp8"]["ms_per_layer"]
        del lins, ys
    return out


def leg_gemm(torch, dev, g, peaks):
    """Secondary metric of BASELINE.json: BFP GEMM TOPS at the LLaMA-7B shapes (T = 4096 tokens, HBFP8 B=64, 2:4 s->q weights):
    2:4-sparse exact-bf16 kind (what BFPLinear runs), dense exact-bf16 kind, int8 + per-block rescale kind; burst (10 launches)
    and sustained (the three shapes cycled for >= 2 s)."""
    from qsi_b200 import _lib, bfp_ops
    L = _lib.lib()
    stream = torch.cuda.current_stream().cuda_stream
    gargs = bfp_ops.unpack_bfp_args(dict(num_format="bfp", sparsity_num_format="bfp", rounding_mode="determ", epsilon=1e-8,
                                         mant_bits=7, block_size=64, w_sparsity=True, N=N_, M=M_, first="s",
                                         sparsity_mode="structured", device="cuda"))

    def timed(fn, iters=10):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        for _ in range(iters):
            fn()
        g1.record()
        torch.cuda.synchronize()
        return g0.elapsed_time(g1) / iters

    margs = dict(gargs, mant_bits=3)                       # HBFP4: the block-scaled FP8-class kind (csrc/bfp_gemm_mx.cu)
    per_shape, calls, ops_total = [], {"sp": [], "bf16": [], "i8": [], "linear_fwd": [], "mx_hbfp4": [], "linear_fwd_hbfp4": []}, 0.0
    ms_sum = {k: 0.0 for k in calls}
    keep = []
    for (T, Nn, Kk) in ((4096, 4096, 4096), (4096, 11008, 4096), (4096, 4096, 11008)):
        xg = torch.randn(T, Kk, device=dev, generator=g)
        wg = torch.randn(Nn, Kk, device=dev, generator=g) * 0.02
        xp, wp = bfp_ops.pack_bfp(xg, identifier="in", **gargs), bfp_ops.pack_bfp(wg, identifier="w", **gargs)
        xb, wb = bfp_ops.pack_bfp_bf16(xg, identifier="in", **gargs), bfp_ops.pack_bfp_bf16(wg, identifier="w", **gargs)
        ws = bfp_ops.compress_2to4_bf16(wb)
        og = torch.empty(T, Nn, device=dev)
        xm = bfp_ops.pack_activation_mx(xg, margs)
        wm = bfp_ops.pack_bfp_mx(wg, 240, fold=True, identifier="w", **margs)
        keep.append((xg, xp, wp, xb, wb, ws, og, xm, wm))
        f = {
            "i8": (lambda xp=xp, wp=wp, og=og, T=T, Nn=Nn, Kk=Kk: _lib.check(L.bfp_gemm_i8(
                xp.mant.data_ptr(), xp.scale_t.data_ptr(), wp.mant.data_ptr(), wp.scale_t.data_ptr(), None, og.data_ptr(), T, Nn, Kk, 64, stream))),
            "bf16": (lambda xb=xb, wb=wb, og=og, T=T, Nn=Nn, Kk=Kk: _lib.check(L.bfp_gemm_bf16(
                xb.data_ptr(), wb.data_ptr(), None, og.data_ptr(), T, Nn, Kk, stream))),
            "sp": (lambda xb=xb, ws=ws, og=og, T=T, Nn=Nn, Kk=Kk: _lib.check(L.bfp_gemm_bf16_sp(
                xb.data_ptr(), ws.comp.data_ptr(), ws.meta.data_ptr(), None, og.data_ptr(), T, Nn, Kk, stream))),
            # the whole BFPLinear forward a caller sees: quantise x on the fly + contraction (weight pack cached)
            "linear_fwd": (lambda xg=xg, ws=ws: bfp_ops.bfp_linear_bf16_sp(bfp_ops.pack_bfp_bf16(xg, identifier="in", **gargs), ws)),
            # HBFP4 (2:4 s->q weights, zeros kept): tcgen05.mma.kind::mxf8f6f4.block_scale on E4M3 mantissas + UE8M0 block scales
            "mx_hbfp4": (lambda xm=xm, wm=wm, og=og, T=T, Nn=Nn, Kk=Kk: _lib.check(L.bfp_gemm_mx(
                xm.vals.data_ptr(), xm.sf.data_ptr(), wm.vals.data_ptr(), wm.sf.data_ptr(), wm.tile_rows, 1, None, og.data_ptr(), T, Nn, Kk, stream))),
            "linear_fwd_hbfp4": (lambda xg=xg, wm=wm: bfp_ops.bfp_linear_mx(bfp_ops.pack_activation_mx(xg, margs), wm)),
        }
        nops = 2.0 * T * Nn * Kk
        row = {"T": T, "N": Nn, "K": Kk}
        for k, fn in f.items():
            ms = timed(fn)
            row[f"{k}_ms"], row[f"{k}_tops"] = ms, nops / ms / 1e9
            ms_sum[k] += ms
            calls[k].append(fn)
        per_shape.append(row)
        ops_total += nops
    burst = {k: ops_total / v / 1e9 for k, v in ms_sum.items()}
    # sustained: cycle the three shapes back to back for >= 2 s per kind (the power-limited steady state)
    sustained = {}
    for k in ("sp", "bf16", "i8", "mx_hbfp4"):
        cyc_ms = ms_sum[k]
        reps = max(3, int(2000.0 / max(cyc_ms, 1e-3)) + 1)
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        g0.record()
        for _ in range(reps):
            for fn in calls[k]:
                fn()
        g1.record(); torch.cuda.synchronize()
        sec = g0.elapsed_time(g1) * 1e-3
        sustained[k] = {"tops": reps * ops_total / sec / 1e12, "seconds": sec, "launches": reps * 3}
    bf16_burst = peaks.get("bf16_tflops") or 1658.0
    bf16_sus = peaks.get("bf16_tflops_sustained") or 1391.6
    int_mm = 2988.0      # torch._int_mm 8192^3 on this pool's B200 (profiles/r01_probe_ref_gpu.log), the measured library int8 figure
    sp_s, bf_s, i8_s, mx_s = sustained["sp"]["tops"], sustained["bf16"]["tops"], sustained["i8"]["tops"], sustained["mx_hbfp4"]["tops"]
    return {
        "tops": burst["sp"], "unit": "TOPS (2*T*N*K ops, dense-equivalent; the 2:4 kernel executes half)",
        "kernel": "bfp_gemm_bf16_sp_kernel (tcgen05.mma.sp.cta_group::2.kind::f16, 2:4-compressed exact-bf16 BFP weight)",
        "burst_tops": burst, "sustained": sustained, "block": 64, "mant_bits": 7, "tokens": 4096, "per_shape": per_shape,
        "roofline": {
            "bound": "tensor", "unit": "TFLOP/s", "timing": "sustained (>= 2 s, three LLaMA-7B shapes cycled)",
            "sparse_bf16_kind": {"achieved": sp_s, "peak": 2.0 * bf16_sus, "frac": sp_s / (2.0 * bf16_sus),
                                 "peak_source": "2 x MEASURED_PEAKS.json bf16_tflops_sustained (2:4 sparse MMA doubles the dense bf16 rate)",
                                 "frac_of_nominal_sparse_bf16_4500": sp_s / 4500.0, "frac_of_measured_int_mm_2988": sp_s / int_mm},
            "dense_bf16_kind": {"achieved": bf_s, "peak": bf16_sus, "frac": bf_s / bf16_sus,
                                "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained", "burst_frac_of_bf16_burst_peak": burst["bf16"] / bf16_burst},
            "mx_hbfp4_kind": {"achieved": mx_s, "peak": 2.0 * bf16_sus, "frac": mx_s / (2.0 * bf16_sus),
                              "peak_source": "2 x MEASURED_PEAKS.json bf16_tflops_sustained (the FP8-class kinds run at twice the bf16 rate)",
                              "frac_of_nominal_fp8_4500": mx_s / 4500.0, "frac_of_measured_int_mm_2988": mx_s / int_mm,
                              "kernel": "bfp_gemm_mx_kernel<240, 2> (tcgen05.mma.cta_group::2.kind::mxf8f6f4.block_scale, dense; mant_bits 3)"},
            "int8_kind": {"achieved": i8_s, "peak": int_mm, "frac": i8_s / int_mm, "peak_source": "measured torch._int_mm 8192^3 (cuBLASLt int8)",
                          "frac_of_nominal_int8_4500": i8_s / 4500.0,
                          "note": "kind::i8 MMA + per-block fp32 rescale on the CUDA cores: epilogue-bound by construction (DESIGN.md section 4); never the default kind"},
        },
    }


def measure_traffic(timeout_s=150):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, averaged over the sweep's launch mix (both
    shapes, HBFP8 B=64, s->q), measured now by an ncu child process of this script (--traffic-child).
    Returns (bytes or None, provenance)."""
    ncu = shutil.which("ncu") or ("/usr/local/cuda/bin/ncu" if os.path.exists("/usr/local/cuda/bin/ncu") else None)
    if ncu is None:
        return None, "ncu not found"
    cmd = [ncu, "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum", "--clock-control", "none", "-k", "regex:quant_stream_kernel",
           "-s", "4", "-c", "4", "--csv", sys.executable, os.path.abspath(__file__), "--traffic-child"]
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout_s, cwd=ROOT)
    except Exception as e:                                  # noqa: BLE001
        return None, f"ncu child failed: {e!r}"[:200]
    vals = {"dram__bytes_read.sum": [], "dram__bytes_write.sum": []}
    import csv
    for row in csv.reader(r.stdout.splitlines()):
        for name in vals:
            if name in row:
                try:
                    i = row.index(name)
                    unit, v = row[i + 1], float(row[i + 2].replace(",", ""))
                    mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
                    vals[name].append(v * mult)
                except (ValueError, IndexError):
                    pass
    if not vals["dram__bytes_read.sum"] or not vals["dram__bytes_write.sum"]:
        return None, ("ncu produced no dram counters (rc %d): " % r.returncode + (r.stderr or r.stdout)[-160:].replace("\n", " "))
    rd = sum(vals["dram__bytes_read.sum"]) / len(vals["dram__bytes_read.sum"])
    wr = sum(vals["dram__bytes_write.sum"]) / len(vals["dram__bytes_write.sum"])
    return rd + wr, (f"ncu child in this run: mean of {len(vals['dram__bytes_read.sum'])} launches alternating 4096x4096 and 4096x11008 fp32 (the sweep's "
                     f"launch mix), read {rd / 1e6:.1f} MB + write {wr / 1e6:.1f} MB per launch")


def traffic_child():
    """Launches the two shapes of the sweep alternately (m = 7, B = 64, s->q); the parent captures 2 launches of each."""
    import torch
    from qsi_b200 import _lib
    L = _lib.lib()
    dev = torch.device("cuda", 0)
    xs = {s: [torch.randn(*s, device=dev) * 0.02 for _ in range(2)] for s in SHAPES}
    ys = {s: torch.empty(*s, device=dev) for s in SHAPES}
    st = torch.cuda.current_stream().cuda_stream
    for i in range(6):
        for s in SHAPES:
            _lib.check(L.bfp_quantize(xs[s][i % 2].data_ptr(), ys[s].data_ptr(), s[0], s[1], _lib.DT_F32, _lib.DT_F32, 64, 7, 1e-8, _lib.ROUND_NEAREST,
                                      0, 0, N_, M_, _lib.ORDER_SPARSIFY_QUANT, _lib.TIE_TORCH_CUDA, st))
    torch.cuda.synchronize()


def leg_reference_same_gpu(torch, dev):
    """The unmodified reference (baseline/_ref) on torch-CUDA on this same GPU, one configuration of the sweep: the like-for-like
    baseline next to the CPU one.  None when baseline/_ref is absent."""
    from _refload import load_reference, ref_args
    ref = load_reference()
    if ref is None:
        return None
    w = torch.randn(4096, 4096, device=dev) * 0.02
    args = ref_args(ref, mant_bits=7, block_size=64, first="s", device="cuda")
    for _ in range(2):
        ref.float_to_bfp_blocked(w, **args, identifier="w")
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        ref.float_to_bfp_blocked(w, **args, identifier="w")
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    return {"what": "reference float_to_bfp_blocked on torch-CUDA, 4096x4096 fp32, HBFP8 block 64, 2:4 s->q (~25 eager kernels + a host-built mask)",
            "ms": ms, "GBps": 4096 * 4096 * 8 / (ms * 1e-3) / 1e9}


# ---------------------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the quant_modes / gemm / column_parallel / traffic legs")
    ap.add_argument("--e2e-steps", type=int, default=0, help="0 = min(steps, 5)")
    ap.add_argument("--traffic-child", action="store_true", help=argparse.SUPPRESS)
    ap.add_argument("--cpu-baseline-child", action="store_true", help=argparse.SUPPRESS)
    a = ap.parse_args()
    global _REAL_STDOUT
    if not (a.traffic_child or a.cpu_baseline_child):
        # stdout carries exactly ONE line, the JSON record: everything else that writes to fd 1 (NCCL's version banner, library chatter)
        # is sent to stderr for the duration of the run
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)
    if a.traffic_child:
        return traffic_child()
    if a.cpu_baseline_child:
        v, kind, cores, sample, dt = cpu_arm(15.0, os.cpu_count() or 1, "sweep7b")
        print(json.dumps({"value": v, "unit": "GB/s", "cores": cores, "kind": kind, "sample": sample, "seconds": round(dt, 2)}), flush=True)
        return None
    a.warmup = max(a.warmup, 3) if a.impl == "ours" else a.warmup

    from qsi_b200 import dist as qd
    rank, local_rank, world = qd.env_world()
    orig_affinity = os.sched_getaffinity(0) if hasattr(os, "sched_getaffinity") else None
    numa_cpus = qd.bind_to_gpu_numa(local_rank) if a.impl == "ours" else None    # before the first pinned allocation
    import torch
    if a.impl == "reference":
        return run_reference_arm(a, rank, world)

    from qsi_b200 import _lib, bfp_ops
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    qd.init("nccl")
    peaks = measured_peaks()
    peak, peak_src = peaks["hbm_gbs"], peaks["source"] + " hbm_gbs"
    cfgs = sweep_configs()

    sampler = ClockSampler(local_rank)
    sampler.start()
    t_gpu0 = time.time()

    # ---- headline ---------------------------------------------------------------------------------------------
    one_pass, elems_rank, launches_pass, keep_pass = make_compress_pass(torch, qd, dev, rank, world)
    total_elems = qd.sum_over_ranks(elems_rank, dev)
    if world == 1:
        device_step, keep_sweep = make_sweep(torch, dev, rank)
        ms_total, launches = timed_region(torch, qd, dev, device_step, a.steps, a.warmup)
        ms_step = ms_total / a.steps
        bytes_step = step_bytes()
        value = bytes_step / (ms_step * 1e-3) / 1e9
        bytes_launch = bytes_step / (len(cfgs) * len(SHAPES))
        del keep_sweep
        # the 65B pass as an extra leg (the N = 1 point of the strong-scaling curve)
        ms_pass_total, _ = timed_region(torch, qd, dev, one_pass, 3, 1)
        ms_pass = ms_pass_total / 3
    else:
        ms_total, launches = timed_region(torch, qd, dev, one_pass, a.steps, a.warmup)
        ms_step = ms_pass = ms_total / a.steps
        bytes_step = int(total_elems) * 8
        value = bytes_step / (ms_step * 1e-3) / 1e9
        bytes_launch = elems_rank * 8 / launches_pass
    compress = {"model": MODEL5, "elements": int(total_elems), "ms": ms_pass, "GBps": total_elems * 8 / (ms_pass * 1e-3) / 1e9,
                "per_gpu_GBps": total_elems * 8 / (ms_pass * 1e-3) / 1e9 / world, "frac_of_hbm_peak_per_gpu": total_elems * 8 / (ms_pass * 1e-3) / 1e9 / world / peak,
                "sharding": "layer l -> rank l % N, no collective", "launches_per_pass_rank0": launches_pass}
    del keep_pass, one_pass
    torch.cuda.empty_cache()

    # roofline of the dominant kernel (quant_stream_kernel): algorithmic bytes per launch / average launch duration
    avg_launch_s = (ms_total * 1e-3) / launches
    achieved = bytes_launch / avg_launch_s / 1e9

    extras = {}
    if not a.no_extras:
        if world > 1:
            try:
                extras["column_parallel"] = leg_column_parallel(torch, qd, dev, rank, world)
            except Exception as e:          # noqa: BLE001  the headline must survive; a failure here is reported, not hidden
                extras["column_parallel"] = {"error": repr(e)[:400]}
        else:
            try:
                extras["quant_modes"] = leg_quant_modes(torch, dev, peak)
            except Exception as e:          # noqa: BLE001
                extras["quant_modes"] = {"error": repr(e)[:300]}
            try:
                extras["unstructured"] = leg_unstructured(torch, dev, peak)
            except Exception as e:
                extras["unstructured"] = {"error": repr(e)[:300]}
            try:
                extras["mx_formats"] = leg_mx_formats(torch, dev, peak)
            except Exception as e:          # noqa: BLE001
                extras["mx_formats"] = {"error": repr(e)[:300]}
            try:
                extras["models"] = leg_models(torch, dev)
            except Exception as e:          # noqa: BLE001
                extras["models"] = {"error": repr(e)[:300]}
            try:
                extras["config3_llama13b_layer"] = leg_config3(torch, dev)
            except Exception as e:          # noqa: BLE001
                extras["config3_llama13b_layer"] = {"error": repr(e)[:300]}
            try:
                extras["gemm"] = leg_gemm(torch, dev, torch.Generator(device=dev).manual_seed(2000), peaks)
            except Exception as e:          # noqa: BLE001
                extras["gemm"] = {"error": repr(e)[:300]}

    # the clock sampler covers the device-timed legs; it is stopped before the host-buffer leg because a 10 Hz nvidia-smi query
    # takes driver locks that the copy submissions of that leg wait on (55 vs 80 GB/s measured)
    t_gpu1 = time.time()
    clocks = sampler.stop(t_gpu0, t_gpu1, "headline + compress_65b + extras legs (device-timed)")

    # ---- e2e: the operator through the public API on pinned HOST tensors (H2D + kernel + D2H per call, inside the timing) ----
    e2e_steps = a.e2e_steps or min(a.steps, 5)
    args = bfp_ops.unpack_bfp_args(dict(num_format="bfp", sparsity_num_format="bfp", rounding_mode="determ", epsilon=1e-8,
                                        w_sparsity=True, N=N_, M=M_, sparsity_mode="structured", device="cuda"))
    os.environ["BFP_TIE_RULE"] = "cuda"
    if world == 1:
        host_in = [(torch.randn(*s, generator=torch.Generator().manual_seed(7)) * 0.02).pin_memory() for s in SHAPES]
        e2e_calls = [(w, m, b, o) for (m, b, o) in cfgs for w in host_in]
        e2e_sample = "the whole 36-call sweep per step"
    else:
        # a bounded sample of the pass: this rank's first layer (layer `rank`), first half of the rows of each of its 7 tensors
        shapes5 = qd.LAYER_SHAPES[MODEL5]
        gen = torch.Generator().manual_seed(70 + rank)
        host_in = [(torch.randn(n // 2, k, generator=gen) * 0.02).pin_memory() for n, k in shapes5]
        e2e_calls = [(w, 7, 64, "s") for w in host_in]
        e2e_sample = f"per step every rank compresses the first half of the rows of the 7 tensors of ONE layer (layer = rank) from pinned host memory: {world} layers / 2 per step"
    e2e_bytes_rank = sum(w.numel() * 8 for (w, _, _, _) in e2e_calls)

    def e2e_step():
        last = None
        for (w, m, b, o) in e2e_calls:
            last = bfp_ops.float_to_bfp_blocked(w, **dict(args, mant_bits=m, block_size=b, first=o), identifier="w")
        return last

    y = None
    for _ in range(8 if world == 1 else 6):                # warm-up: staging buffers + torch's pinned-host block cache.  The result is
        y = e2e_step()                                     # HELD exactly as in the timed loop: holding one output alive needs one more
                                                           # cached pinned block (a one-time ~0.3 s cudaHostAlloc that round 1 timed)
    torch.cuda.synchronize()
    qd.barrier(dev)
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        y = e2e_step()
    chk = float(y[0, 0])                                   # result is already on the host; touch it
    e2e_s = qd.max_over_ranks(time.perf_counter() - t0, dev)
    e2e_bytes = qd.sum_over_ranks(e2e_bytes_rank, dev)
    e2e_value = e2e_bytes * e2e_steps / e2e_s / 1e9
    del host_in, e2e_calls, y

    if rank != 0:
        return
    traffic, traffic_src = (None, "skipped (--no-extras)") if a.no_extras else ((None, "not measured at N > 1 (ncu profiles one process)") if world > 1 else measure_traffic())
    cpu = ref_gpu = None
    if world == 1 and not a.no_cpu_baseline:
        # in a child process started with the ORIGINAL CPU affinity: this process (and every thread pool it has created) is
        # pinned to the GPU's NUMA node, the CPU baseline must see all host cores
        if orig_affinity is not None:
            os.sched_setaffinity(0, orig_affinity)
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--cpu-baseline-child"], capture_output=True, text=True, timeout=300, cwd=ROOT)
            cpu = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])
        except Exception as e:              # noqa: BLE001
            cpu = {"error": repr(e)[:200]}
        try:
            ref_gpu = leg_reference_same_gpu(torch, dev)
        except Exception as e:              # noqa: BLE001
            ref_gpu = {"error": repr(e)[:200]}
    cfg = config_for(world)                # identical in both arms; the descriptive extras live in config_detail
    cfg_detail = {"launches_per_step": int(launches // a.steps), "bytes_per_step": int(bytes_step),
                  "parallelism": f"tensor-sharded x{world}, no collective on the data path",
                  "l2": "inputs rotate over buffers larger than the 126 MB L2 (sweep: 4x64MB + 2x180MB; pass: up to 4 resident layers of 3.2 GB); outputs alternate"}
    line = {
        "metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak" if world == 1 else "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": cfg, "config_detail": cfg_detail,
        "hbm_peak_pct": 100.0 * (value / world) / peak,
        "roofline": {"bound": "hbm", "kernel": "quant_stream_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "frac_of_nominal_hbm3e_8000": achieved / 8000.0, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": bytes_launch},
        "e2e": {"value": e2e_value, "unit": "GB/s", "h2d_bytes_per_step": int(e2e_bytes // 2), "d2h_bytes_per_step": int(e2e_bytes // 2),
                "steps": e2e_steps, "api": "bfp_ops.float_to_bfp_blocked(pinned CPU tensor) -> bfp_quantize_host", "sample": e2e_sample,
                "pcie_frac_of_measured_97GBps_bidirectional": e2e_value / world / 97.0, "numa_bound_cpus": len(numa_cpus) if numa_cpus else None,
                "check": chk},
        "compress_65b": compress,
        "cpu_baseline": cpu,
        "reference_same_gpu": ref_gpu,
        "gpu_launches": int(launches),
        "clocks": clocks,
    }
    line.update(extras)
    emit(line)


if __name__ == "__main__":
    try:
        main()
    finally:
        if "torch.distributed" in sys.modules:
            from qsi_b200 import dist as _qd
            _qd.shutdown()
