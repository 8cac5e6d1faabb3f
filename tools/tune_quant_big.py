"""Throughput of the fused quantiser vs tensor size (fp32, HBFP8 B=64, 2:4 s->q), raw C-ABI calls with preallocated outputs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from qsi_b200 import _lib
L = _lib.lib(); st = torch.cuda.current_stream().cuda_stream
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for shape in [(4096, 11008), (8192, 8192), (22016, 8192), (8192, 22016), (36864, 9216)]:
    n = 3 if shape[0] * shape[1] < 2e8 else 2
    xs = [torch.randn(*shape, device="cuda") * 0.02 for _ in range(n)]; ys = [torch.empty_like(xs[0]) for _ in range(2)]
    def run(i): _lib.check(L.bfp_quantize(xs[i % n].data_ptr(), ys[i % 2].data_ptr(), shape[0], shape[1], 0, 0, 64, 7, 1e-8, 0, 0, 0, 2, 4, 1, 0, st))
    for i in range(3): run(i)
    torch.cuda.synchronize(); e0.record()
    for i in range(10): run(i)
    e1.record(); torch.cuda.synchronize(); us = e0.elapsed_time(e1) * 100
    y = torch.empty_like(xs[0]); 
    for i in range(3): y.copy_(xs[i % n])
    torch.cuda.synchronize(); e0.record()
    for i in range(10): y.copy_(xs[i % n])
    e1.record(); torch.cuda.synchronize(); usc = e0.elapsed_time(e1) * 100
    b = shape[0] * shape[1] * 8
    print(f"{shape}: quantise {us:.1f} us = {b/us/1e3:.0f} GB/s | torch copy_ {usc:.1f} us = {b/usc/1e3:.0f} GB/s", flush=True)
    del xs, ys, y
