"""BFPLinear forward + backward (row f3) at LLaMA-7B shapes, fp32 modules, T = 4096 tokens: tensor-core autograd Function vs
this repo's fake-quant path under torch autograd vs the reference on the same GPU."""
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
res = []
for rounding in ("determ", "stoc"):
    kw = dict(num_format="bfp", sparsity_num_format="bfp", rounding_mode=rounding, epsilon=1e-8, mant_bits=7, block_size=64, w_sparsity=True,
              N=2, M=4, first="s", sparsity_mode="structured", device="cuda")
    for N, K in [(4096, 4096), (11008, 4096), (4096, 11008)]:
        torch.manual_seed(0)
        lin = ours.BFPLinear(K, N, bias=True, **dict(kw)).cuda()
        x = torch.randn(4096, K, device="cuda", requires_grad=True)
        gy = torch.randn(4096, N, device="cuda")
        def step(m):
            x.grad = None; m.weight.grad = None
            m(x).backward(gy)
        row = {"rounding": rounding, "N": N, "K": K, "T": 4096}
        row["tc_ms"] = t(lambda: step(lin)); row["tflops"] = 3 * 2 * 4096 * N * K / row["tc_ms"] / 1e9
        gx = x.grad.clone(); gw = lin.weight.grad.clone()
        os.environ["BFP_TRAIN_PATH"] = "fakequant"; row["fakequant_autograd_ms"] = t(lambda: step(lin), 3); os.environ["BFP_TRAIN_PATH"] = "tc"
        if rounding == "determ":
            row["grad_x_rel_diff"] = float((gx - x.grad).norm() / x.grad.norm()); row["grad_w_rel_diff"] = float((gw - lin.weight.grad).norm() / lin.weight.grad.norm())
        if ref is not None:
            rl = ref.BFPLinear(K, N, bias=True, **dict(kw)).cuda(); rl.weight, rl.bias = lin.weight, lin.bias
            row["reference_ms"] = t(lambda: step(rl), 2); row["speedup_vs_reference"] = row["reference_ms"] / row["tc_ms"]
        print(json.dumps(row), flush=True); res.append(row)
if len(sys.argv) > 1: json.dump(res, open(sys.argv[1], "w"), indent=1)
