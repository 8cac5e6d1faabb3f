"""Python-level overhead of the host path: float_to_bfp_blocked(pinned CPU tensor) vs the bare C-ABI call."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from qsi_b200 import _lib, bfp_ops
L = _lib.lib()
args = bfp_ops.unpack_bfp_args(dict(num_format="bfp", sparsity_num_format="bfp", rounding_mode="determ", epsilon=1e-8, mant_bits=7, block_size=64,
                                    w_sparsity=True, N=2, M=4, first="s", sparsity_mode="structured", device="cuda"))
os.environ["BFP_TIE_RULE"] = "cuda"
for shape in [(4096, 4096), (4096, 11008)]:
    x = (torch.randn(*shape) * 0.02).pin_memory(); y = torch.empty_like(x).pin_memory(); nbytes = x.numel() * 8
    f = lambda: _lib.check(L.bfp_quantize_host(x.data_ptr(), y.data_ptr(), shape[0], shape[1], 0, 0, 64, 7, 1e-8, 0, 0, 0, 2, 4, 1, 0))
    g = lambda: bfp_ops.float_to_bfp_blocked(x, **args, identifier="w")
    def alloc():
        return torch.empty(x.shape, dtype=x.dtype, pin_memory=True)
    for name, fn in (("C ABI", f), ("python API", g), ("pinned alloc only", alloc)):
        fn(); fn(); t0 = time.perf_counter()
        for _ in range(8): r = fn()
        dt = (time.perf_counter() - t0) / 8
        print(f"{shape} {name}: {dt*1e3:.3f} ms  {nbytes/dt/1e9:.1f} GB/s", flush=True)
    import cProfile, pstats
    pr = cProfile.Profile(); pr.enable()
    for _ in range(5): g()
    pr.disable(); pstats.Stats(pr).sort_stats("cumtime").print_stats(8)
