"""BFPConv2d: im2col + BFP GEMM on the tensor cores vs fused quantiser + library convolution vs the reference, for a patch
embedding (kernel == stride) and ResNet-style 3x3 / 1x1 / 7x7 convolutions."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from qsi_b200 import bfp_ops as ours
from _refload import load_reference
ref = load_reference()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
def t(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize(); e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n
kw = dict(num_format="bfp", sparsity_num_format="bfp", rounding_mode="determ", epsilon=1e-8, mant_bits=5, block_size=64, device="cuda")
res = []
for name, (B, C, H, O, k, s, p) in {"vit patch 16x16/16": (256, 3, 224, 768, 16, 16, 0), "resnet 3x3/1": (128, 64, 56, 64, 3, 1, 1),
                                    "resnet 1x1/1": (128, 256, 56, 64, 1, 1, 0), "resnet 3x3/2": (128, 128, 56, 128, 3, 2, 1),
                                    "resnet stem 7x7/2": (128, 3, 224, 64, 7, 2, 3)}.items():
    torch.manual_seed(0)
    conv = ours.BFPConv2d(C, O, k, stride=s, padding=p, bias=True, **dict(kw)).cuda()
    x = torch.randn(B, C, H, H, device="cuda")
    row = {"conv": name, "x": [B, C, H, H], "out_channels": O}
    with torch.no_grad():
        os.environ["BFP_CONV_IM2COL_MAX_EXPANSION"] = "1e9"; row["im2col_tc_ms"] = t(lambda: conv(x)); y_tc = conv(x)
        os.environ["BFP_CONV_IM2COL_MAX_EXPANSION"] = "-1"; row["fused_quant_cudnn_ms"] = t(lambda: conv(x)); y_lib = conv(x)
        del os.environ["BFP_CONV_IM2COL_MAX_EXPANSION"]; row["default_ms"] = t(lambda: conv(x))
        row["rel_diff_tc_vs_lib"] = float((y_tc - y_lib).norm() / y_lib.norm())
        if ref is not None:
            rc = ref.BFPConv2d(C, O, k, stride=s, padding=p, bias=True, **dict(kw)).cuda(); rc.weight, rc.bias = conv.weight, conv.bias
            row["reference_ms"] = t(lambda: rc(x), 2); row["rel_diff_vs_reference"] = float((conv(x) - rc(x)).norm() / rc(x).norm())
    print(json.dumps(row), flush=True); res.append(row)
if len(sys.argv) > 1: json.dump(res, open(sys.argv[1], "w"), indent=1)
