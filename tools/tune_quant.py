"""Per-configuration throughput table of the fused quantiser (device-resident inputs, CUDA events, rotating buffers).
    python tools/tune_quant.py [--ctas 4,6,8] [--dtypes f32,bf16] [--quick]"""
import argparse, itertools, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from qsi_b200 import _lib

ap = argparse.ArgumentParser()
ap.add_argument("--ctas", default="8")
ap.add_argument("--dtypes", default="f32")
ap.add_argument("--quick", action="store_true")
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--out", default="")
ap.add_argument("--pdl", default="1")
ap.add_argument("--tma", default="0")
a = ap.parse_args()
L = _lib.lib()
dev = torch.device("cuda", 0)
TD = {"f32": (torch.float32, _lib.DT_F32), "bf16": (torch.bfloat16, _lib.DT_BF16), "f16": (torch.float16, _lib.DT_F16)}
ORD = {"q": _lib.ORDER_QUANT_ONLY, "sq": _lib.ORDER_SPARSIFY_QUANT, "qs": _lib.ORDER_QUANT_SPARSIFY, "s": _lib.ORDER_SPARSIFY_ONLY}
shapes = [(4096, 4096), (4096, 11008)]
cfgs = [(7, 64, "sq"), (7, 64, "qs"), (3, 16, "sq"), (3, 16, "qs"), (5, 32, "sq"), (7, 64, "q"), (7, 64, "s")]
if a.quick:
    cfgs = [(7, 64, "sq"), (3, 16, "qs")]
stream = torch.cuda.current_stream().cuda_stream
res = []
for dtn in a.dtypes.split(","):
    tdt, cdt = TD[dtn]
    for shape in shapes:
        nbuf = max(2, int(600e6 // (shape[0] * shape[1] * tdt.itemsize)) + 1)
        xs = [(torch.randn(*shape, device=dev) * 0.02).to(tdt) for _ in range(nbuf)]
        ys = [torch.empty_like(xs[0]) for _ in range(2)]
        for ctas, pdl, tma in itertools.product([int(c) for c in a.ctas.split(",")], [int(c) for c in a.pdl.split(",")], [int(c) for c in a.tma.split(",")]):
            _lib.set_option("stream_ctas_per_sm", ctas)
            _lib.set_option("pdl", pdl)
            _lib.set_option("quant_tma", tma)
            for (m, b, o), rnd in itertools.product(cfgs, (0, 1)):
                if rnd and (o == "s" or a.quick and o != "sq"):
                    continue
                odt = _lib.DT_F32 if rnd else cdt
                yo = [torch.empty(*shape, device=dev) for _ in range(2)] if (rnd and dtn != "f32") else ys
                def run(i):
                    _lib.check(L.bfp_quantize(xs[i % nbuf].data_ptr(), yo[i % 2].data_ptr(), shape[0], shape[1], cdt, odt, b, m, 1e-8, rnd, 1, i, 2, 4, ORD[o], 0, stream))
                for i in range(3):
                    run(i)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for i in range(a.iters):
                    run(i)
                e1.record(); torch.cuda.synchronize()
                us = e0.elapsed_time(e1) * 1e3 / a.iters
                nbytes = shape[0] * shape[1] * (tdt.itemsize + (4 if rnd else tdt.itemsize))
                gbs = nbytes / us / 1e3
                res.append(dict(dtype=dtn, shape=shape, ctas=ctas, pdl=pdl, tma=tma, m=m, B=b, order=o, stoc=rnd, us=round(us, 2), GBps=round(gbs, 1)))
                print(f"{dtn:5s} {str(shape):14s} ctas={ctas:2d} pdl={pdl} tma={tma} m={m} B={b:3d} {o:2s} {'stoc' if rnd else 'near'}  {us:8.2f} us  {gbs:8.1f} GB/s", flush=True)
        del xs, ys
# reference points: torch copy of the same sizes
for shape in shapes:
    x = [torch.randn(*shape, device=dev) for _ in range(6)]; y = torch.empty_like(x[0])
    for i in range(3): y.copy_(x[i])
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True); e0.record()
    for i in range(12): y.copy_(x[i % 6])
    e1.record(); torch.cuda.synchronize(); us = e0.elapsed_time(e1) * 1e3 / 12
    print(f"torch copy_ {shape}: {us:.2f} us {x[0].numel()*8/us/1e3:.1f} GB/s", flush=True)
    res.append(dict(kind="torch_copy", shape=shape, us=us, GBps=x[0].numel() * 8 / us / 1e3))
if a.out:
    json.dump(res, open(a.out, "w"), indent=1)
