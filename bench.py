#!/usr/bin/env python
"""bench.py -- headline benchmark of the BFP + N:M hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

Workload (BASELINE.json configs[1]): LLaMA-7B weight shapes (4096x4096, 4096x11008), fp32, quantise+sparsify sweep over
BFP4/6/8 (mant_bits 3/5/7) x block 16/32/64 x both orderings, 2:4, round-to-nearest.  One "step" = one pass of the
sweep = 36 fused-kernel launches.  Bytes are ALGORITHMIC: numel x (sizeof(in) + sizeof(out)) = 8 B/element.

  value     whole-job GB/s with inputs resident in HBM (CUDA events on the launching stream, max over ranks)
  e2e       the same sweep through the public API with HOST (pinned) buffers: H2D + kernel + D2H inside the timed region
  roofline  the stream kernel against the measured HBM copy peak (MEASURED_PEAKS.json)
  cpu_baseline  the CPU implementation (reference if baseline/_ref is present, else the oracle port) on a bounded sample

Multi-GPU: every rank runs the same per-GPU sweep on its own tensors (tensor-sharded compression pass: no data-path
collective), value = units of all ranks / max-over-ranks time  ->  "scaling": "weak".
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "BFP+N:M quantize GB/s (% HBM peak); BFP GEMM TOPS at LLaMA-7B shapes"
SHAPES = [(4096, 4096), (4096, 11008)]
MANTS = [3, 5, 7]
BLOCKS = [16, 32, 64]
ORDERS = ["s", "q"]            # first='s' (sparsify->quantise) / 'q' (quantise->sparsify)
N_, M_ = 2, 4


def sweep_configs():
    return [(m, b, o) for m in MANTS for b in BLOCKS for o in ORDERS]


def step_bytes(shapes=SHAPES, bytes_per_elt=8):
    return sum(r * k for r, k in shapes) * bytes_per_elt * len(sweep_configs())


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


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

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
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
                "samples": len(sm), "window": "warm-up + timed sweep + GEMM leg"}


# ---------------------------------------------------------------------------------------------------------------
# CPU arm: the reference's own implementation (or the oracle port) on the host cores
# ---------------------------------------------------------------------------------------------------------------
def cpu_arm(budget_s, threads):
    """Times one bounded sample of the sweep on the CPU.  Returns (GB/s, kind, cores, sample description, seconds)."""
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
    # calibrate on a small slice, then size the row sample so the whole sweep fits the budget
    probe = torch.randn(256, 4096, generator=g) * 0.02
    run(probe, 7, 64, "s")
    t0 = time.perf_counter()
    run(probe, 7, 64, "s"); run(probe, 3, 16, "q")
    per_elt = (time.perf_counter() - t0) / (2 * probe.numel())
    total_elts = sum(r * k for r, k in SHAPES) * len(sweep_configs())
    # size the sample (rows of each shape, whole passes when one pass is cheap) for ~0.7 x budget of CPU work; the small
    # probe over-estimates the per-element cost, so a sample that came out under half the budget is re-sized once from its
    # own timing and re-measured
    for attempt in range(2):
        frac = min(1.0, 0.7 * budget_s / max(per_elt * total_elts, 1e-9))
        rows = [max(8, int(r * frac) // 8 * 8) for r, _ in SHAPES]
        ws = [torch.randn(rs, k, generator=g) * 0.02 for rs, (_, k) in zip(rows, SHAPES)]
        est = per_elt * sum(w.numel() for w in ws) * len(sweep_configs())
        passes = max(1, min(4, int(0.7 * budget_s / max(est, 1e-9))))
        t0 = time.perf_counter()
        nbytes = 0
        for _ in range(passes):
            for (m, b, o) in sweep_configs():
                for w in ws:
                    run(w, m, b, o)
                    nbytes += w.numel() * 8
        dt = time.perf_counter() - t0
        if dt >= 0.4 * budget_s or (frac >= 1.0 and passes >= 4):
            break
        per_elt = dt / (nbytes / 8)
    sample = (f"{passes} pass(es) of the 18-config sweep on the first {rows[0]} rows of 4096x4096 and {rows[1]} rows of 4096x11008 "
              f"(fp32, 2:4, nearest), {nbytes / 1e9:.2f} GB algorithmic")
    return nbytes / dt / 1e9, kind, threads, sample, dt


def run_reference_arm(a, rank, world):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    budget = max(2.0, min(20.0, 150.0 / max(1, a.steps + a.warmup)))
    for _ in range(a.warmup):
        cpu_arm(budget, threads)
    vals, secs, info = [], [], None
    for _ in range(a.steps):
        v, kind, cores, sample, dt = cpu_arm(budget, threads)
        vals.append(v); secs.append(dt); info = (kind, cores, sample)
    value = sum(vals) / len(vals)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": a.gpus, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": 1e3 * sum(secs) / len(secs), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "llama7b_quant_sparsify_sweep", "shapes": SHAPES, "mant_bits": MANTS, "block": BLOCKS,
                       "orders": ["s->q", "q->s"], "nm": "2:4", "rounding": "nearest"},
            "cpu_baseline": {"value": value, "unit": "GB/s", "cores": info[1], "kind": info[0], "sample": info[2]},
            "e2e": {"value": value, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


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
    ap.add_argument("--e2e-steps", type=int, default=0, help="0 = min(steps, 5)")
    a = ap.parse_args()
    a.warmup = max(a.warmup, 3) if a.impl == "ours" else a.warmup

    import torch
    from qsi_b200 import dist as qd
    rank, local_rank, world = qd.env_world()
    if a.impl == "reference":
        return run_reference_arm(a, rank, world)

    from qsi_b200 import _lib, bfp_ops
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    qd.init("nccl")
    L = _lib.lib()
    peak, peak_src = measured_peaks()
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

    sampler = ClockSampler(local_rank)
    sampler.start()
    t_wait = time.time()
    while sampler.proc is not None and not sampler.rows and time.time() - t_wait < 4.0:   # nvidia-smi takes ~1 s to print its first row
        time.sleep(0.05)
    t_gpu0 = time.time()
    for i in range(a.warmup):
        device_step(i)
    torch.cuda.synchronize()
    qd.barrier(dev)
    torch.cuda.synchronize()
    n0 = _lib.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for i in range(a.steps):
        device_step(i)
    ev1.record()
    torch.cuda.synchronize()
    qd.barrier(dev)
    torch.cuda.synchronize()
    launches = _lib.launch_count() - n0
    ms_total = qd.max_over_ranks(ev0.elapsed_time(ev1), dev)
    ms_step = ms_total / a.steps
    bytes_step = step_bytes()
    value = world * bytes_step / (ms_step * 1e-3) / 1e9

    # roofline of the dominant kernel (quant_stream_kernel): algorithmic bytes per launch / average launch duration
    avg_launch_s = (ms_total * 1e-3) / launches
    achieved = (bytes_step / (len(cfgs) * len(SHAPES))) / avg_launch_s / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tp):
        traffic = json.load(open(tp)).get("dram_bytes_per_launch")

    # secondary metric of BASELINE.json: BFP GEMM TOPS at the LLaMA-7B shapes (T = 4096 tokens, HBFP8 B=64, 2:4 s->q weights),
    # for both tensor-core kinds: exact-bf16 operands (default of BFPLinear) and int8 mantissas + per-block rescale
    gemm = None
    try:
        gargs = bfp_ops.unpack_bfp_args(dict(num_format="bfp", sparsity_num_format="bfp", rounding_mode="determ", epsilon=1e-8,
                                             mant_bits=7, block_size=64, w_sparsity=True, N=N_, M=M_, first="s",
                                             sparsity_mode="structured", device="cuda"))
        per_shape, ops_total, ms_sum = [], 0.0, {"sp": 0.0, "bf16": 0.0, "i8": 0.0, "linear_fwd": 0.0}

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

        for (T, Nn, Kk) in ((4096, 4096, 4096), (4096, 11008, 4096), (4096, 4096, 11008)):
            xg = torch.randn(T, Kk, device=dev, generator=g)
            wg = torch.randn(Nn, Kk, device=dev, generator=g) * 0.02
            xp, wp = bfp_ops.pack_bfp(xg, identifier="in", **gargs), bfp_ops.pack_bfp(wg, identifier="w", **gargs)
            xb, wb = bfp_ops.pack_bfp_bf16(xg, identifier="in", **gargs), bfp_ops.pack_bfp_bf16(wg, identifier="w", **gargs)
            og = torch.empty(T, Nn, device=dev)
            ms_i8 = timed(lambda: _lib.check(L.bfp_gemm_i8(xp.mant.data_ptr(), xp.scale_t.data_ptr(), wp.mant.data_ptr(),
                                                           wp.scale_t.data_ptr(), None, og.data_ptr(), T, Nn, Kk, 64, stream)))
            ms_bf = timed(lambda: _lib.check(L.bfp_gemm_bf16(xb.data_ptr(), wb.data_ptr(), None, og.data_ptr(), T, Nn, Kk, stream)))
            # 2:4-compressed weight on the structured-sparse tensor-core path (what BFPLinear runs for these arguments)
            ws = bfp_ops.compress_2to4_bf16(wb)
            ms_sp = timed(lambda: _lib.check(L.bfp_gemm_bf16_sp(xb.data_ptr(), ws.comp.data_ptr(), ws.meta.data_ptr(), None, og.data_ptr(),
                                                                T, Nn, Kk, stream)))
            # the whole BFPLinear forward a caller sees: quantise x on the fly + contraction (weight pack cached)
            ms_fwd = timed(lambda: bfp_ops.bfp_linear_bf16_sp(bfp_ops.pack_bfp_bf16(xg, identifier="in", **gargs), ws))
            nops = 2.0 * T * Nn * Kk
            per_shape.append({"T": T, "N": Nn, "K": Kk, "sp_ms": ms_sp, "sp_tops": nops / ms_sp / 1e9, "bf16_ms": ms_bf,
                              "bf16_tops": nops / ms_bf / 1e9, "i8_ms": ms_i8, "i8_tops": nops / ms_i8 / 1e9, "linear_fwd_ms": ms_fwd,
                              "linear_fwd_tops": nops / ms_fwd / 1e9})
            ops_total += nops
            ms_sum["sp"] += ms_sp; ms_sum["bf16"] += ms_bf; ms_sum["i8"] += ms_i8; ms_sum["linear_fwd"] += ms_fwd
            del xg, wg, xp, wp, xb, wb, og, ws
        tops = {k: ops_total / v / 1e9 for k, v in ms_sum.items()}
        gemm = {"tops": tops["sp"], "unit": "TOPS (2*T*N*K ops, dense-equivalent; the 2:4 kernel executes half)",
                "kernel": "bfp_gemm_bf16_sp_kernel (tcgen05.mma.sp.cta_group::2.kind::f16, 2:4-compressed exact-bf16 BFP weight)",
                "frac_of_nominal_int8_4500": tops["sp"] / 4500.0, "frac_of_nominal_sparse_bf16_4500": tops["sp"] / 4500.0,
                "dense_bf16_kernel_tops": tops["bf16"], "dense_bf16_kernel": "bfp_gemm_bf16_kernel (tcgen05.mma.kind::f16, dense)",
                "dense_bf16_frac_of_measured_bf16_peak": tops["bf16"] / 1658.0,
                "i8_kernel_tops": tops["i8"], "i8_kernel": "bfp_gemm_i8_kernel (tcgen05.mma.kind::i8 + per-block fp32 rescale)",
                "linear_forward_tops": tops["linear_fwd"], "block": 64, "mant_bits": 7, "tokens": 4096, "per_shape": per_shape}
    except Exception as e:          # the headline metric must survive a GEMM problem; report it instead of hiding it
        gemm = {"error": repr(e)[:300]}

    # the clock sampler covers the device-timed legs (quantiser sweep, GEMM); it is stopped before the host-buffer leg because
    # a 10 Hz nvidia-smi query takes driver locks that the copy submissions of that leg wait on (55 vs 80 GB/s measured)
    t_gpu1 = time.time()
    clocks = sampler.stop(t_gpu0, t_gpu1)
    # e2e: the same sweep through the public API on pinned HOST tensors (H2D + kernel + D2H per call, inside the timing)
    e2e_steps = a.e2e_steps or min(a.steps, 5)
    host_in = {s: (torch.randn(*s, generator=torch.Generator().manual_seed(7)) * 0.02).pin_memory() for s in SHAPES}
    args = bfp_ops.unpack_bfp_args(dict(num_format="bfp", sparsity_num_format="bfp", rounding_mode="determ", epsilon=1e-8,
                                        w_sparsity=True, N=N_, M=M_, sparsity_mode="structured", device="cuda"))

    def e2e_step():
        last = None
        for (m, b, o) in cfgs:
            for s in SHAPES:
                last = bfp_ops.float_to_bfp_blocked(host_in[s], **dict(args, mant_bits=m, block_size=b, first=o), identifier="w")
        return last

    os.environ["BFP_TIE_RULE"] = "cuda"
    for _ in range(8):                                     # warm-up: staging buffers + torch's pinned-host block cache; measured:
        e2e_step()                                         # the rate climbs from ~58 to ~80 GB/s over the first 5-6 passes
    torch.cuda.synchronize()
    qd.barrier(dev)
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        y = e2e_step()
    chk = float(y[0, 0])                                   # result is already on the host; touch it
    e2e_s = qd.max_over_ranks(time.perf_counter() - t0, dev)
    e2e_value = world * bytes_step * e2e_steps / e2e_s / 1e9
    h2d = bytes_step // 2
    d2h = bytes_step // 2

    if rank != 0:
        return
    cpu = None
    if world == 1 and not a.no_cpu_baseline:
        v, kind, cores, sample, dt = cpu_arm(15.0, os.cpu_count() or 1)
        cpu = {"value": v, "unit": "GB/s", "cores": cores, "kind": kind, "sample": sample, "seconds": round(dt, 2)}
    line = {
        "metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": "llama7b_quant_sparsify_sweep", "shapes": SHAPES, "mant_bits": MANTS, "block": BLOCKS,
                   "orders": ["s->q", "q->s"], "nm": "2:4", "rounding": "nearest", "launches_per_step": len(cfgs) * len(SHAPES),
                   "bytes_per_step": bytes_step, "parallelism": f"tensor-sharded x{world}, no collective",
                   "l2": "inputs rotate over 4x64MB + 2x180MB buffers (> 126 MB L2); outputs alternate"},
        "hbm_peak_pct": 100.0 * (value / world) / peak,
        "roofline": {"bound": "hbm", "kernel": "quant_stream_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": bytes_step / (len(cfgs) * len(SHAPES))},
        "e2e": {"value": e2e_value, "unit": "GB/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": e2e_steps, "api": "bfp_ops.float_to_bfp_blocked(pinned CPU tensor) -> bfp_quantize_host", "check": chk},
        "cpu_baseline": cpu,
        "gemm": gemm,
        "gpu_launches": int(launches),
        "clocks": clocks,
    }
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    try:
        main()
    finally:
        if "torch.distributed" in sys.modules:
            from qsi_b200 import dist as _qd
            _qd.shutdown()
