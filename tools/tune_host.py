"""End-to-end (host buffers) throughput of bfp_quantize_host vs pipeline chunk size.  python tools/tune_host.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from qsi_b200 import _lib
L = _lib.lib()
for shape in [(4096, 4096), (4096, 11008)]:
    x = (torch.randn(*shape) * 0.02).pin_memory(); y = torch.empty_like(x).pin_memory()
    nbytes = x.numel() * 8
    # plain copies for reference: H2D alone, D2H alone, both at once
    d = torch.empty(*shape, device="cuda"); d2 = torch.empty(*shape, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    def both():
        with torch.cuda.stream(s1): d.copy_(x, non_blocking=True)
        with torch.cuda.stream(s2): y.copy_(d2, non_blocking=True)
        torch.cuda.synchronize()
    for name, fn in (("h2d", lambda: (d.copy_(x, non_blocking=True), torch.cuda.synchronize())), ("d2h", lambda: (y.copy_(d2, non_blocking=True), torch.cuda.synchronize())), ("h2d+d2h", both)):
        fn(); t0 = time.perf_counter()
        for _ in range(5): fn()
        dt = (time.perf_counter() - t0) / 5
        print(f"{shape} {name}: {dt*1e3:.2f} ms  {(nbytes if name == 'h2d+d2h' else nbytes / 2) / dt / 1e9:.1f} GB/s", flush=True)
    for mb, mn in ((8, 8), (16, 16), (8, 1), (16, 1), (16, 2), (32, 1), (32, 2), (16, 0.5)):
        _lib.set_option("host_chunk_bytes", mb << 20); _lib.set_option("host_chunk_min_bytes", int(mn * (1 << 20)))
        f = lambda: _lib.check(L.bfp_quantize_host(x.data_ptr(), y.data_ptr(), shape[0], shape[1], 0, 0, 64, 7, 1e-8, 0, 0, 0, 2, 4, 1, 0))
        f(); f(); t0 = time.perf_counter()
        for _ in range(8): f()
        dt = (time.perf_counter() - t0) / 8
        print(f"{shape} chunk max {mb:2d} min {mn} MiB: {dt*1e3:.2f} ms  {nbytes/dt/1e9:.1f} GB/s (in+out bytes)", flush=True)
