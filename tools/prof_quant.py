"""A handful of launches of one quantiser configuration, for ncu.  python tools/prof_quant.py [dtype m B order rows K [stoc]]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from qsi_b200 import _lib
dtn = sys.argv[1] if len(sys.argv) > 1 else "f32"
m = int(sys.argv[2]) if len(sys.argv) > 2 else 7
B = int(sys.argv[3]) if len(sys.argv) > 3 else 64
o = sys.argv[4] if len(sys.argv) > 4 else "sq"
rows = int(sys.argv[5]) if len(sys.argv) > 5 else 4096
K = int(sys.argv[6]) if len(sys.argv) > 6 else 11008
stoc = 1 if (len(sys.argv) > 7 and sys.argv[7] == "stoc") else 0
TD = {"f32": (torch.float32, _lib.DT_F32), "bf16": (torch.bfloat16, _lib.DT_BF16), "f16": (torch.float16, _lib.DT_F16)}
ORD = {"q": 0, "sq": 1, "qs": 2, "s": 3}
tdt, cdt = TD[dtn]
L = _lib.lib()
xs = [(torch.randn(rows, K, device="cuda") * 0.02).to(tdt) for _ in range(4)]
y = torch.empty(rows, K, device="cuda", dtype=torch.float32 if stoc else tdt)
st = torch.cuda.current_stream().cuda_stream
for i in range(6):
    _lib.check(L.bfp_quantize(xs[i % 4].data_ptr(), y.data_ptr(), rows, K, cdt, _lib.DT_F32 if stoc else cdt, B, m, 1e-8, stoc, 1234, 16 * i, 2, 4, ORD[o], 0, st))
torch.cuda.synchronize()
print("ok", float(y.float().abs().sum()))
