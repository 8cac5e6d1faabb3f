"""Where does the HOST time of a small-model forward go?  cProfile of OPT-125M (8 x 512 tokens) with every linear swapped."""
import cProfile, pstats, os, sys, io
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch
import model_dropin as md
from qsi_b200 import bfp_ops
kw = dict(num_format="bfp", sparsity_num_format="bfp", rounding_mode="determ", epsilon=1e-8, mant_bits=7, weight_mant_bits=15, block_size=64,
          w_sparsity=True, N=2, M=4, first="s", sparsity_mode="structured", sparsity_frac=0.5, device="cuda")
model, cfg = md.build("opt", 0)
md.swap(model, bfp_ops, kw, md.OPT_TARGETS)
model = model.cuda()
inp = dict(input_ids=torch.randint(0, cfg.vocab_size, (8, 512)).cuda())
with torch.no_grad():
    for _ in range(3): model(**inp)
    torch.cuda.synchronize()
    pr = cProfile.Profile(); pr.enable()
    for _ in range(5): model(**inp)
    torch.cuda.synchronize()
    pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(45); print(s.getvalue()[:9000])
