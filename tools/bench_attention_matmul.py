"""F_matmul_bfp at GPT-2 attention shapes (modeling_gpt2.py:205-207, :295; B = 8, 12 heads, S = 1024, D = 64): one batched
tcgen05 launch per product vs the reference on the same GPU."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from qsi_b200 import bfp_ops as ours
from _refload import load_reference
ref = load_reference()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n
kw = dict(num_format="bfp", sparsity_num_format="bfp", rounding_mode="determ", epsilon=1e-8, mant_bits=7, block_size=64, device="cuda")
res = []
for B, H, S, D in [(8, 12, 1024, 64), (8, 12, 512, 64), (4, 32, 2048, 128)]:
    torch.manual_seed(0)
    q, k, v = (torch.randn(B, H, S, D, device="cuda") for _ in range(3))
    p = torch.softmax(torch.randn(B, H, S, S, device="cuda"), dim=-1)
    mm = ours.F_matmul_bfp(**dict(kw))
    row = {"B": B, "H": H, "S": S, "D": D}
    with torch.no_grad():
        row["qk_ms"] = t(lambda: mm(q, k.transpose(-1, -2))); row["pv_ms"] = t(lambda: mm(p, v))
        row["qk_tflops"] = 2 * B * H * S * S * D / row["qk_ms"] / 1e9; row["pv_tflops"] = 2 * B * H * S * S * D / row["pv_ms"] / 1e9
        if ref is not None:
            rm = ref.F_matmul_bfp(**dict(kw))
            row["ref_qk_ms"] = t(lambda: rm(q, k.transpose(-1, -2)), 3); row["ref_pv_ms"] = t(lambda: rm(p, v), 3)
            y, yr = mm(q, k.transpose(-1, -2)), rm(q, k.transpose(-1, -2))
            row["qk_rel_diff"] = float((y - yr).norm() / yr.norm())
            y, yr = mm(p, v), rm(p, v)
            row["pv_rel_diff"] = float((y - yr).norm() / yr.norm())
    print(json.dumps(row), flush=True); res.append(row)
if len(sys.argv) > 1: json.dump(res, open(sys.argv[1], "w"), indent=1)
