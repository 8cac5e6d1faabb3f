"""Model-level MX drop-in: stock OPT-125M (OPTConfig() defaults, 8 x 512 tokens) with the six block linears of every layer swapped for
MXLinear (fp8_e4m3, block 32, bfloat16, 2:4-pruned weights: the substitution modeling_opt.py:165-169,328-330 makes), against the same
model built from the library-style emulation in torch ops (tests/test_mx_gpu._EmulatedMXLinear) on the same GPU.
    python tools/model_dropin_mx.py [out.json]"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import qsi_b200  # noqa: E402,F401
from qsi_b200 import mx_layers as mx  # noqa: E402
from oracle import mx_oracle as O  # noqa: E402  (format table for the emulation)
import test_mx_gpu as T  # noqa: E402

FMT, BLOCK = "fp8_e4m3", 32
TARGETS = ("q_proj", "k_proj", "v_proj", "out_proj", "fc1", "fc2")


def build(make):
    import transformers
    torch.manual_seed(0)
    cfg = transformers.OPTConfig()
    model = transformers.OPTForCausalLM(cfg).eval()
    n = 0
    for parent in list(model.modules()):
        for name, ch in list(parent.named_children()):
            if isinstance(ch, torch.nn.Linear) and name in TARGETS:
                setattr(parent, name, make(ch))
                n += 1
    return model.cuda(), cfg, n


def ours(ch):
    new = mx.MXLinear(ch.in_features, ch.out_features, bias=ch.bias is not None, mx_specs=dict(T.SPEC, w_elem_format=FMT, a_elem_format=FMT),
                      sparsity=True, device="cuda", sparsity_mode="structured", N=2, M=4)
    new.weight, new.bias = ch.weight, ch.bias
    return new.eval()


def timed(model, ids, iters):
    with torch.no_grad():
        y = model(input_ids=ids).logits
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(iters):
            y = model(input_ids=ids).logits
        torch.cuda.synchronize()
    return (time.perf_counter() - t0) / iters * 1e3, y.float()


def main():
    m1, cfg, n = build(ours)
    ids = torch.randint(0, cfg.vocab_size, (8, 512), generator=torch.Generator().manual_seed(1)).cuda()
    with torch.no_grad():
        m1(input_ids=ids)                                           # first forward prunes the weights (mx_layers.py:49-56)
    ms1, y1 = timed(m1, ids, 3)
    pruned = {k: v.detach().clone() for k, v in m1.state_dict().items()}
    del m1
    m2, _, _ = build(lambda ch: T._EmulatedMXLinear(ch, FMT, BLOCK, O))
    m2.load_state_dict(pruned)
    ms2, y2 = timed(m2, ids, 1)
    # the emulation against ITSELF with another summation order (the contraction accumulated in fp64 instead of the library's fp32 GEMM): how
    # far two faithful evaluations of the same model drift apart through twelve layers of quantisers
    class _Emu64(T._EmulatedMXLinear):
        def forward(self, x):
            K = x.shape[-1]
            qx = T._torch_quantize_mx(self.rb(x).view(-1, K), self.fmt, self.block, self.O).view(x.shape)
            qw = T._torch_quantize_mx(self.rb(self.weight.detach()), self.fmt, self.block, self.O)
            y = self.rb(torch.nn.functional.linear(qx.double(), qw.double()).float())
            return self.rb(y + self.rb(self.bias.detach())) if self.bias is not None else y
    del m2
    m3, _, _ = build(lambda ch: _Emu64(ch, FMT, BLOCK, O))
    m3.load_state_dict(pruned)
    with torch.no_grad():
        y3 = m3(input_ids=ids).logits.float()
    out = {"model": "OPT-125M (OPTConfig defaults), 8 x 512 tokens", "swapped_modules": n, "format": FMT, "block": BLOCK, "bfloat": 16, "weights": "2:4 pruned",
           "ours_forward_ms": ms1, "emulation_same_gpu_forward_ms": ms2, "speedup": ms2 / ms1,
           "logits_rel_diff": float((y1 - y2).norm() / y2.norm()),
           "emulation_vs_emulation_fp64_accumulation_rel_diff": float((y3 - y2).norm() / y2.norm()),
           "ours_vs_emulation_fp64_accumulation_rel_diff": float((y1 - y3).norm() / y3.norm())}
    s = json.dumps(out, indent=1)
    print(s)
    if len(sys.argv) > 1:
        open(sys.argv[1], "w").write(s)


if __name__ == "__main__":
    main()
