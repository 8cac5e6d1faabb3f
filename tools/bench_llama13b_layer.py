"""BASELINE config 3: LLaMA-2-13B BFP linear forward, 4096 tokens per step, on-the-fly BFP activations x 2:4-sparse BFP weights,
one B200.  Times the seven BFPLinear forwards of one decoder layer (HBFP8, block 64, 2:4 s->q weights) through the public
module API (activation quantise + GEMM, packed weight cached) and, when the reference sources are present, the reference's
BFPLinear on the same GPU.   python tools/bench_llama13b_layer.py [--model llama-13b] [--out file.json]"""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from qsi_b200 import bfp_ops as ours, dist as qd, _lib
from _refload import load_reference
ap = argparse.ArgumentParser(); ap.add_argument("--model", default="llama-13b"); ap.add_argument("--tokens", type=int, default=4096)
ap.add_argument("--iters", type=int, default=10); ap.add_argument("--out", default=""); ap.add_argument("--rounding", default="determ", choices=["determ", "stoc"])
a = ap.parse_args()
kw = dict(num_format="bfp", sparsity_num_format="bfp", rounding_mode=a.rounding, epsilon=1e-8, mant_bits=7, weight_mant_bits=15, block_size=64,
          w_sparsity=True, N=2, M=4, first="s", sparsity_mode="structured", sparsity_frac=0.5, device="cuda")
shapes = qd.LAYER_SHAPES[a.model]; T = a.tokens
flop = sum(2.0 * T * n * k for n, k in shapes)
res = {"model": a.model, "tokens": T, "layer_shapes_N_K": shapes, "tflop_per_layer": flop / 1e12, "format": "HBFP8 B=64, 2:4 s->q weights, " + ("nearest" if a.rounding == "determ" else "stochastic rounding (weights re-quantised every forward, like the reference)")}
torch.manual_seed(0)
ws = [torch.randn(n, k, device="cuda") * 0.02 for n, k in shapes]
xs = {k: torch.randn(T, k, device="cuda") for k in sorted({k for _, k in shapes})}
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for tag, impl in (("ours", ours), ("reference", load_reference())):
    if impl is None: continue
    lins = []
    for w in ws:
        l = impl.BFPLinear(w.shape[1], w.shape[0], bias=False, **dict(kw)).cuda(); l.weight = torch.nn.Parameter(w, requires_grad=False); lins.append(l)
    with torch.no_grad():
        def layer():
            return [l(xs[l.in_features]) for l in lins]
        n0 = _lib.launch_count()
        ys = layer(); ys = layer(); torch.cuda.synchronize()
        iters = a.iters if tag == "ours" else 2
        e0.record()
        for _ in range(iters): ys = layer()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
    res[tag] = {"ms_per_layer": ms, "tflops": flop / ms / 1e9, "ms_40_layers": ms * qd.NUM_LAYERS.get(a.model, 40)}
    if tag == "ours" and a.rounding == "determ":
        # the same seven forwards captured once in a CUDA graph and replayed: removes the Python / launch overhead
        with torch.no_grad():
            s_ = torch.cuda.Stream(); s_.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s_):
                layer(); layer()
            torch.cuda.current_stream().wait_stream(s_)
            gph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gph):
                ys_g = layer()
            gph.replay(); torch.cuda.synchronize()
            e0.record()
            for _ in range(iters): gph.replay()
            e1.record(); torch.cuda.synchronize()
            msg = e0.elapsed_time(e1) / iters
        res[tag]["cuda_graph_ms_per_layer"] = msg; res[tag]["cuda_graph_tflops"] = flop / msg / 1e9
        res[tag]["cuda_graph_equal_to_eager"] = all(torch.equal(a_, b_) for a_, b_ in zip(ys_g, ys))
    if tag == "ours": res[tag]["kernel_launches_per_layer"] = (_lib.launch_count() - n0) / (iters + 2); keep = [y.clone() for y in ys]
    elif a.rounding == "determ": res["rel_err_vs_reference"] = max(float((a_ - b_).norm() / b_.norm()) for a_, b_ in zip(keep, ys))
    print(tag, res[tag], flush=True)
    del lins, ys
if "reference" in res: res["speedup_vs_reference_same_gpu"] = res["reference"]["ms_per_layer"] / res["ours"]["ms_per_layer"]
print(json.dumps(res))
if a.out: json.dump(res, open(a.out, "w"), indent=1)
