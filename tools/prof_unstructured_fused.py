"""Per-phase device times of the fused unstructured pipeline (BFP_UNSTRUCTURED_TIMING=1 makes the library print them).
usage: python tools/prof_unstructured_fused.py [dtype order rows cols [relu]]"""
import os, sys
os.environ["BFP_UNSTRUCTURED_TIMING"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from qsi_b200 import bfp_ops as ours, _lib
DT = {"f32": torch.float32, "bf16": torch.bfloat16, "f16": torch.float16}
ORD = {"s": _lib.ORDER_SPARSIFY_ONLY, "sq": _lib.ORDER_SPARSIFY_QUANT, "qs": _lib.ORDER_QUANT_SPARSIFY}
cfgs = [("f32", "s", 4096, 4096, ""), ("f32", "sq", 4096, 11008, ""), ("f32", "qs", 4096, 11008, ""), ("f32", "sq", 4096, 11008, "relu"),
        ("bf16", "sq", 4096, 11008, ""), ("f16", "qs", 4096, 11008, "")]
if len(sys.argv) > 4:
    cfgs = [(sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), sys.argv[5] if len(sys.argv) > 5 else "")]
for dt, order, r, c, kind in cfgs:
    x = (torch.randn(r, c, device="cuda") * 0.02).to(DT[dt])
    if kind == "relu":
        x = torch.relu(x)
    print(f"--- {dt} {order} {r}x{c} {kind}", file=sys.stderr, flush=True)
    for _ in range(4):
        ours._unstructured_fused(x, 0.5, ORD[order], block_size=64, mant_bits=7, epsilon=1e-8)
    torch.cuda.synchronize()
