"""MX (OCP Microscaling) path on the B200: the fused quantiser against the HBM roofline and the MX linear on the tensor cores, next to
the library-style emulation (the eager torch op sequence microxcaling runs, restated in tests/test_mx_gpu.py) on the same GPU.
    python tools/bench_mx.py [out.json]"""
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
from oracle import mx_oracle as O  # noqa: E402  (format table only)


def timed(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def emulation(A, fmt, block):
    import test_mx_gpu as T
    bits = A.view(torch.int32)
    A = ((bits + 0x8000) & ~0xFFFF).view(torch.float32)          # bfloat16, half away (magnitude bits; sign bit is untouched by the add)
    return T._torch_quantize_mx(A, fmt, block, O)


def main():
    dev = torch.device("cuda", 0)
    peak = 6461.8
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:      # noqa: BLE001
        pass
    out = {"hbm_peak_GBps": peak, "quantiser": [], "linear": []}
    g = torch.Generator(device=dev).manual_seed(0)
    for shape in ((4096, 4096), (4096, 11008)):
        bufs32 = [torch.randn(*shape, device=dev, generator=g) for _ in range(4)]
        for dt in (torch.float32, torch.bfloat16):
            bufs = [b.to(dt) for b in bufs32]
            esz = bufs[0].element_size()
            for fmt in ("fp8_e4m3", "fp4_e2m1", "int8"):
                i = [0]

                def fq():
                    i[0] += 1
                    return mx._mx_quantize_last(bufs[i[0] % 4], mx.ELEM_FORMATS[fmt], 32, 8, 16, False)
                us = timed(fq) * 1e3
                n = shape[0] * shape[1]
                row = {"shape": list(shape), "dtype": str(dt).split(".")[-1], "format": fmt, "block": 32, "out": "fake-quant", "us": us,
                       "bytes_per_element": 2 * esz, "GBps": n * 2 * esz / us / 1e3, "frac_of_hbm_peak": n * 2 * esz / us / 1e3 / peak}
                out["quantiser"].append(row)
                if fmt != "int8":
                    sp = mx.finalize_mx_specs(mx.apply_mx_specs(dict(block_size=32, bfloat=16, scale_bits=8, w_elem_format=fmt, a_elem_format=fmt)))

                    def pk():
                        i[0] += 1
                        return mx._pack_block_scaled(bufs[i[0] % 4], mx.ELEM_FORMATS[fmt], 128, sp, 16)
                    us = timed(pk) * 1e3
                    bpe = esz + 1 + 1.0 / 32
                    out["quantiser"].append({"shape": list(shape), "dtype": str(dt).split(".")[-1], "format": fmt, "block": 32, "out": "E4M3 + UE8M0 atoms", "us": us,
                                             "bytes_per_element": bpe, "GBps": n * bpe / us / 1e3, "frac_of_hbm_peak": n * bpe / us / 1e3 / peak})
            if dt == torch.float32 and shape == (4096, 4096):
                us = timed(lambda: emulation(bufs[0], "fp8_e4m3", 32), iters=3, warm=1) * 1e3
                out["quantiser"].append({"shape": list(shape), "dtype": "float32", "format": "fp8_e4m3", "block": 32, "out": "library-style emulation (eager torch ops, same GPU)",
                                         "us": us, "GBps": shape[0] * shape[1] * 8 / us / 1e3})
        del bufs32
    T = 4096
    for (N, K) in ((4096, 4096), (11008, 4096), (4096, 11008)):
        x = torch.randn(T, K, device=dev, generator=g)
        for fmt, sparse in (("fp8_e4m3", False), ("fp4_e2m1", False), ("int8", False), ("int8", True)):
            lin = mx.MXLinear(K, N, bias=False, mx_specs=dict(block_size=32, bfloat=16, scale_bits=8, w_elem_format=fmt, a_elem_format=fmt),
                              sparsity=sparse, device="cuda", sparsity_mode="structured", N=2, M=4).to(dev).eval()
            with torch.no_grad():
                ms = timed(lambda: lin(x), iters=10)
                row = {"T": T, "N": N, "K": K, "format": fmt, "weight_2to4": sparse, "ms": ms, "tflops": 2.0 * T * N * K / ms / 1e9,
                       "kind": "block-scaled mxf8f6f4" if mx._format_id(fmt) in mx._E4M3_SUBSET else ("sparse bf16" if sparse else "dense bf16")}
                if (N, K) == (4096, 4096) and not sparse:
                    w = lin.weight.detach()

                    def emu():
                        y = torch.nn.functional.linear(emulation(x, fmt, 32), emulation(w, fmt, 32))
                        return ((y.view(torch.int32) + 0x8000) & ~0xFFFF).view(torch.float32)
                    row["emulation_same_gpu_ms"] = timed(emu, iters=2, warm=1)
                    yo, ye = lin(x), emu()
                    row["rel_diff_vs_emulation"] = float((yo - ye).norm() / ye.norm())
                    row["mismatch_fraction_vs_emulation"] = float((yo != ye).float().mean())
                out["linear"].append(row)
            del lin
    s = json.dumps(out, indent=1)
    print(s)
    if len(sys.argv) > 1:
        open(sys.argv[1], "w").write(s)


if __name__ == "__main__":
    t0 = time.time()
    main()
    print(f"# {time.time() - t0:.1f} s", file=sys.stderr)
