import os, sys, time, cProfile, pstats
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from qsi_b200 import bfp_ops, _lib
args = bfp_ops.unpack_bfp_args(dict(num_format="bfp", sparsity_num_format="bfp", rounding_mode="determ", epsilon=1e-8, mant_bits=7, block_size=64,
                                    w_sparsity=True, N=2, M=4, first="s", sparsity_mode="structured", device="cuda"))
w = torch.randn(8192, 8192, device="cuda")
for _ in range(3): y = bfp_ops.float_to_bfp_blocked(w, **args, identifier="w")
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(200): y = bfp_ops.float_to_bfp_blocked(w, **args, identifier="w")
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"host time per call {1e6*(t1-t0)/200:.1f} us; incl. drain {1e6*(t2-t0)/200:.1f} us")
small = torch.randn(64, 64, device="cuda")
t0 = time.perf_counter()
for _ in range(2000): y = bfp_ops.float_to_bfp_blocked(small, **args, identifier="w")
torch.cuda.synchronize(); print(f"small tensor: {1e6*(time.perf_counter()-t0)/2000:.1f} us per call")
pr = cProfile.Profile(); pr.enable()
for _ in range(2000): y = bfp_ops.float_to_bfp_blocked(small, **args, identifier="w")
pr.disable(); pstats.Stats(pr).sort_stats("tottime").print_stats(10)
