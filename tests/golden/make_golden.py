"""Generates the golden fixtures for the BFP + N:M hot path by RUNNING THE REFERENCE.

The reference (src/transformers/bfp/bfp_ops.py) has no golden vectors of its own, so parity is
pinned by recording the reference's outputs on fixed inputs:

    python tests/golden/make_golden.py --device cpu     # in the build container (reads /root/reference)
    python tests/golden/make_golden.py --device cuda    # on a B200 box (reads the git-ignored baseline/_ref copy)

writes tests/golden/ref_<device>.npz (cuda: gpurun_out/ref_cuda.npz, copied into tests/golden/ afterwards).
torch.topk's tie-breaking differs between torch-CPU and torch-CUDA, hence one file per device;
everything else is identical between the two (checked by tests/test_oracle.py).

Stochastic rounding: torch.rand inside the reference is wrapped so the uniforms it drew are recorded next to the
output; the oracle and the CUDA kernel are then checked on those same uniforms.
bf16 arrays are stored as uint16 bit patterns (key suffix '__bf16').
"""
import argparse
import itertools
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from _refload import load_reference, ref_args  # noqa: E402

DTYPES = {"f32": torch.float32, "f16": torch.float16, "bf16": torch.bfloat16}


def to_np(t):
    t = t.detach().cpu().contiguous()
    if t.dtype == torch.bfloat16:
        return t.view(torch.int16).numpy().view(np.uint16), "bf16"
    return t.numpy(), {torch.float32: "f32", torch.float16: "f16"}[t.dtype]


def put(store, key, t):
    a, dt = to_np(t)
    store[f"{key}__{dt}"] = a


def make_inputs():
    g = torch.Generator().manual_seed(1234)
    cases = {}
    cases["w_8x256"] = torch.randn(8, 256, generator=g) * 0.02
    cases["x_6x128_outl"] = torch.randn(6, 128, generator=g)
    cases["x_6x128_outl"][2, 17] = 37.5
    cases["ragged_3x5x200"] = torch.randn(3, 5, 200, generator=g)
    cases["conv_6x3x16x16"] = torch.randn(6, 3, 16, 16, generator=g) * 0.1
    cases["k_lt_b_4x10"] = torch.randn(4, 10, generator=g)
    cases["one_block_1x64"] = torch.randn(1, 64, generator=g)
    ties = torch.randint(-3, 4, (16, 64), generator=g).float()
    cases["ties_16x64"] = ties
    z = torch.randn(4, 128, generator=g)
    z[1] = 0.0
    z[2, :64] = 0.0
    z[3, 5] = -0.0
    cases["zeros_4x128"] = z
    p2 = torch.tensor([[2.0 ** k for k in range(-12, 20)]]) * torch.tensor([[1.0], [-1.0], [1.0 + 2 ** -20], [1.0 - 2 ** -20]])
    cases["pow2_4x32"] = p2
    cases["tiny_2x64"] = torch.randn(2, 64, generator=g) * 1e-9
    cases["huge_2x64"] = torch.randn(2, 64, generator=g) * 1e30
    cases["kat_sat"] = torch.tensor([[1.0, 0.99, -1.0, 0.5]])
    cases["kat_even"] = torch.tensor([[1.0, 0.0625, 0.1875, 0.3125, 0.4375, -0.0625, -0.1875, 0.99]])
    cases["kat_grid"] = torch.tensor([[128, .5, 1.5, 2.5, -.5, -1.5]])
    cases["kat_nm"] = torch.tensor([[-1.0, 2.0, -3.0, 4.0]])
    cases["kat_negzero"] = torch.tensor([[-0.3, 0.3, 100.0, -100.0]])
    return cases


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--device", default="cpu")
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    dev = a.device
    ref = load_reference()
    if ref is None:
        sys.exit("reference not found (need /root/reference or baseline/_ref)")
    store = {}
    meta = []

    # 1. exponent boundaries: 2^k (+u ulp), bfp_ops.py:29-33
    for dtn, dtype in DTYPES.items():
        ks = range(-30, 31) if dtn != "f16" else range(-14, 16)
        vals = []
        for k in ks:
            x = torch.tensor(2.0 ** k, dtype=dtype)
            for _ in range(9):
                vals.append(x.clone())
                x = torch.nextafter(x, torch.tensor(float("inf"), dtype=dtype))
        x = torch.stack(vals).reshape(-1, 1)
        e = ref.get_exponent(x.to(dev), 1e-8)
        put(store, f"expb_in_{dtn}", x)
        store[f"expb_out_{dtn}"] = e.float().cpu().numpy()

    # 2. quantiser / sparsifier / both orderings on the input zoo
    inputs = make_inputs()
    for name, t32 in inputs.items():
        for dtn, dtype in DTYPES.items():
            if dtn != "f32" and name.startswith(("huge", "tiny", "pow2")) and dtn == "f16":
                pass  # fp16 overflow/underflow cases are kept on purpose: NaN/Inf are part of the semantics
            t = t32.to(dtype)
            put(store, f"in_{name}_{dtn}", t)
            td = t.to(dev)
            combos = [(3, 16), (5, 32), (7, 64), (7, 16), (15, 64)] if dtn == "f32" else [(3, 16), (7, 64), (5, 32)]
            for (m, B) in combos:
                y = ref._no_sparsity_float_to_bfp(td, B, m, 1e-8, "determ", dev)
                put(store, f"q_{name}_{dtn}_m{m}_b{B}", y)
                for first, tag in (("s", "sq"), ("q", "qs")):
                    args = ref_args(ref, mant_bits=m, block_size=B, first=first, device=dev)
                    y = ref.float_to_bfp_blocked(td, **args, identifier="w")
                    put(store, f"{tag}_{name}_{dtn}_m{m}_b{B}_2:4", y)
            for (N, M) in ((2, 4), (1, 4), (3, 4), (4, 8), (1, 2), (2, 8), (8, 16), (3, 5)):
                y = ref._structured_N_M_sparsity(td, dev, N, M)
                put(store, f"s_{name}_{dtn}_{N}:{M}", y)
    # other N:M through the full entry point (q->s ties at other group sizes)
    for (N, M) in ((1, 4), (4, 8), (2, 8)):
        for dtn in ("f32", "bf16"):
            t = inputs["w_8x256"].to(DTYPES[dtn]).to(dev)
            for first, tag in (("s", "sq"), ("q", "qs")):
                args = ref_args(ref, mant_bits=3, block_size=32, first=first, N=N, M=M, device=dev)
                put(store, f"{tag}_w_8x256_{dtn}_m3_b32_{N}:{M}", ref.float_to_bfp_blocked(t, **args, identifier="w"))

    # 3. all weak orderings of 4 magnitudes -> 2:4 / 1:4 / 3:4 masks (topk tie table)
    pats = sorted({tuple(sorted(set(p)).index(v) for v in p) for p in itertools.product(range(4), repeat=4)})
    tp = torch.tensor(pats, dtype=torch.float32) + 1.0
    tp = tp * torch.tensor([1.0, -1.0, 1.0, -1.0])
    put(store, "tie_in", tp)
    for (N, M) in ((2, 4), (1, 4), (3, 4)):
        put(store, f"tie_out_{N}:{M}", ref._structured_N_M_sparsity(tp.to(dev), dev, N, M))

    # 4. stochastic rounding with recorded uniforms (bfp_ops.py:20-23)
    real_rand = torch.rand
    for dtn, dtype in DTYPES.items():
        for (m, B) in ((3, 16), (7, 64)):
            t = (inputs["w_8x256"] * 50).to(dtype)
            drawn = []

            def rec(*args, **kw):
                u = real_rand(*args, **kw)
                drawn.append(u)
                return u
            ref.torch.rand = rec
            try:
                torch.manual_seed(7)
                y = ref._no_sparsity_float_to_bfp(t.to(dev), B, m, 1e-8, "stoc", dev)
            finally:
                ref.torch.rand = real_rand
            put(store, f"stoc_in_{dtn}_m{m}_b{B}", t)
            store[f"stoc_u_{dtn}_m{m}_b{B}"] = drawn[0].float().cpu().numpy().reshape(t.shape)
            put(store, f"stoc_out_{dtn}_m{m}_b{B}", y)

    # 5. BFPLinear forward (bfp_ops.py:270-287): x [2,12,128], W [48,128], HBFP8/64 + 2:4 s->q
    g = torch.Generator().manual_seed(99)
    x = torch.randn(2, 12, 128, generator=g)
    for first in ("s", "q"):
        args = ref_args(ref, mant_bits=7, block_size=64, first=first, device=dev)
        lin = ref.BFPLinear(128, 48, bias=True, **dict(args))
        with torch.no_grad():
            lin.weight.copy_(torch.randn(48, 128, generator=g) * 0.05)
            lin.bias.copy_(torch.randn(48, generator=g))
        lin = lin.to(dev)
        with torch.no_grad():
            y = lin(x.to(dev))
        put(store, f"lin_{first}_x", x)
        put(store, f"lin_{first}_w", lin.weight.detach())
        put(store, f"lin_{first}_b", lin.bias.detach())
        put(store, f"lin_{first}_y", y)

    # 6. the 'int' number format (bfp_ops.py:111-120 -> int_ops.Quantizer): weights and 2-D / 3-D / 4-D activations
    g = torch.Generator().manual_seed(321)
    int_in = {"w2d": torch.randn(12, 96, generator=g) * 0.3, "a3d": torch.randn(2, 9, 40, generator=g), "a4d": torch.randn(2, 5, 6, 8, generator=g)}
    int_in["w2d"][1] = int_in["w2d"][1].abs()
    int_in["w2d"][2] = 0.0
    for name, t32 in int_in.items():
        for dtn, dtype in DTYPES.items():
            t = t32.to(dtype)
            put(store, f"int_in_{name}_{dtn}", t)
            for ident, bits in itertools.product(("w", "in"), (8, 4)):
                args = ref_args(ref, sparsity_num_format="int", mant_bits=bits, w_sparsity=False, device=dev)
                put(store, f"int_out_{name}_{dtn}_{ident}_b{bits}", ref.float_to_bfp_blocked(t.to(dev), **args, identifier=ident))

    # 7. unstructured sparsity (bfp_ops.py:61-71).  Tie-free inputs for both devices; tie-heavy input only where the
    #    tie order is a usable contract (torch-CUDA: index order).
    g = torch.Generator().manual_seed(55)
    un = {"randn": torch.randn(50, 333, generator=g)}
    if dev != "cpu":
        un["ties"] = torch.randint(-3, 4, (64, 257), generator=g).float()
    for name, t32 in un.items():
        for dtn in ("f32", "bf16"):
            t = t32.to(DTYPES[dtn]) if name == "ties" else (t32 + torch.arange(t32.numel()).reshape(t32.shape) * 1e-3).to(torch.float32)
            if dtn == "bf16" and name != "ties":
                continue
            put(store, f"un_in_{name}_{dtn}", t)
            for frac in (0.5, 0.13, 0.9):
                put(store, f"un_out_{name}_{dtn}_{frac}", ref._unstructured_sparsity(t.to(dev), dev, frac))

    store["__meta__"] = np.array([f"device={dev}", f"torch={torch.__version__}"])
    out = a.out or os.path.join(HERE, f"ref_{dev}.npz")
    os.makedirs(os.path.dirname(os.path.abspath(out)), exist_ok=True)
    np.savez_compressed(out, **store)
    print(f"wrote {out}: {len(store)} arrays, {os.path.getsize(out) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
