"""Unstructured (global magnitude) sparsity + HBFP quantiser: the fused two-read pipeline against the stand-alone kernels
(multi-pass radix select + streaming quantiser), LLaMA-7B weight shapes, per dtype / order / input kind.  Algorithmic bytes:
numel x (sizeof(in) + sizeof(out)).  Inputs rotate over buffers larger than the L2."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from qsi_b200 import bfp_ops as ours, _lib

e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
def t(fn, bufs, n=20):
    for i in range(3): fn(bufs[i % len(bufs)])
    torch.cuda.synchronize(); e0.record()
    for i in range(n): fn(bufs[i % len(bufs)])
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n * 1e3

ORD = {"s": _lib.ORDER_SPARSIFY_ONLY, "sq": _lib.ORDER_SPARSIFY_QUANT, "qs": _lib.ORDER_QUANT_SPARSIFY}
def fused(x, order, m, B, rounding="determ"):
    return ours._unstructured_fused(x, 0.5, ORD[order], block_size=B, mant_bits=m, epsilon=1e-8, rounding_mode=rounding)
def composed(x, order, m, B, rounding="determ"):
    os.environ["BFP_UNSTRUCTURED_FUSED"] = "0"
    try:
        q = lambda z: ours._fused(z, _lib.ORDER_QUANT_ONLY, block_size=B, mant_bits=m, epsilon=1e-8, rounding_mode=rounding)
        if order == "s": return ours._unstructured_sparsity(x, "cuda", 0.5)
        if order == "sq": return q(ours._unstructured_sparsity(x, "cuda", 0.5))
        return ours._unstructured_sparsity(q(x), "cuda", 0.5)
    finally:
        os.environ.pop("BFP_UNSTRUCTURED_FUSED", None)

rows = []
DT = {"f32": torch.float32, "bf16": torch.bfloat16, "f16": torch.float16}
for shape in [(4096, 4096), (4096, 11008)]:
    for dt in ("f32", "bf16", "f16"):
        nbuf = 4 if shape[1] == 4096 else 2
        base = [(torch.randn(*shape, device="cuda") * 0.02).to(DT[dt]) for _ in range(nbuf)]
        kinds = {"randn": base}
        if dt == "f32":
            kinds["relu"] = [torch.relu(b) for b in base]
        for kind, bufs in kinds.items():
            for order, m, B, rounding in (("s", 7, 64, "determ"), ("sq", 7, 64, "determ"), ("qs", 7, 64, "determ"), ("qs", 3, 16, "determ"), ("sq", 7, 64, "stoc")):
                if kind == "relu" and order not in ("s", "sq"): continue
                if rounding == "stoc" and dt == "f16": continue
                out_b = 4 if (rounding == "stoc" and order != "s") else base[0].element_size()
                nb = base[0].numel() * (base[0].element_size() + out_b)
                us_f = t(lambda x: fused(x, order, m, B, rounding), bufs)
                us_c = t(lambda x: composed(x, order, m, B, rounding), bufs, n=5)
                same = bool(torch.equal(fused(bufs[0], order, m, B, "determ"), composed(bufs[0], order, m, B, "determ")))
                r = dict(shape=list(shape), dtype=dt, input=kind, order=order, mant_bits=m, block=B, rounding=rounding, fused_us=round(us_f, 1),
                         fused_GBps=round(nb / us_f / 1e3), composed_us=round(us_c, 1), composed_GBps=round(nb / us_c / 1e3), speedup=round(us_c / us_f, 2),
                         equal=same)
                rows.append(r); print(json.dumps(r), flush=True)
if len(sys.argv) > 1:
    json.dump(rows, open(sys.argv[1], "w"), indent=1)
