"""Model-level drop-in check (BASELINE.json configs[0] and [3]): stock `transformers` OPT-125M / ViT-B/16 with every
block nn.Linear replaced by BFPLinear (and the ViT patch-embedding conv by BFPConv2d), exactly the substitution the
reference's patched modeling_opt.py:162-176,325-335 / modeling_vit.py:168-215 make.  Runs the model once with this
repo's bfp_ops and, when the reference sources are present, once with the reference's, and compares logits.

    python tools/model_dropin.py [opt|vit|llama] [--layers L] [--batch B] [--seq S] [--dtype fp32|fp16|bf16] [--out file.json]
(llama: a LLaMA-architecture decoder at reduced width -- hidden 2048, 4 layers by default -- with all seven projections of
every layer swapped, run in fp16 like the reference's LLaMA scripts)
"""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from qsi_b200 import bfp_ops as ours
from _refload import load_reference

OPT_TARGETS = ("q_proj", "k_proj", "v_proj", "out_proj", "fc1", "fc2")
LLAMA_TARGETS = ("q_proj", "k_proj", "v_proj", "o_proj", "gate_proj", "up_proj", "down_proj")


def swap(model, impl, kw, targets=None, conv_name="projection"):
    n = 0
    for parent in list(model.modules()):
        for name, ch in list(parent.named_children()):
            if isinstance(ch, torch.nn.Linear) and (targets is None or name in targets):
                new = impl.BFPLinear(ch.in_features, ch.out_features, bias=ch.bias is not None, **dict(kw))
                new.weight, new.bias = ch.weight, ch.bias
                new.train(ch.training)
                setattr(parent, name, new); n += 1
            elif isinstance(ch, torch.nn.Conv2d) and name == conv_name:
                new = impl.BFPConv2d(ch.in_channels, ch.out_channels, ch.kernel_size, ch.stride, ch.padding, ch.dilation, ch.groups,
                                     bias=ch.bias is not None, **dict(kw))
                new.weight, new.bias = ch.weight, ch.bias
                new.train(ch.training)
                setattr(parent, name, new); n += 1
    return n


def build(kind, layers):
    import transformers
    torch.manual_seed(0)
    if kind == "opt":
        cfg = transformers.OPTConfig()
        if layers: cfg.num_hidden_layers = layers
        return transformers.OPTForCausalLM(cfg).eval(), cfg
    if kind == "llama":
        cfg = transformers.LlamaConfig(hidden_size=2048, intermediate_size=5504, num_hidden_layers=layers or 4, num_attention_heads=16,
                                       num_key_value_heads=16, vocab_size=32000, max_position_embeddings=2048)
        return transformers.LlamaForCausalLM(cfg).eval(), cfg
    cfg = transformers.ViTConfig()
    if layers: cfg.num_hidden_layers = layers
    return transformers.ViTForImageClassification(cfg).eval(), cfg


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("kind", nargs="?", default="opt", choices=["opt", "vit", "llama"])
    ap.add_argument("--dtype", default="", choices=["", "fp32", "fp16", "bf16"])
    ap.add_argument("--layers", type=int, default=0); ap.add_argument("--batch", type=int, default=0); ap.add_argument("--seq", type=int, default=512)
    ap.add_argument("--mant", type=int, default=0); ap.add_argument("--out", default="")
    ap.add_argument("--rounding", default="determ", choices=["determ", "stoc"]); ap.add_argument("--format", default="bfp", choices=["bfp", "int"]); ap.add_argument("--mode", default="structured", choices=["structured", "unstructured"]); ap.add_argument("--first", default="s", choices=["s", "q"])
    a = ap.parse_args()
    dev = "cuda"
    m = a.mant or (5 if a.kind == "vit" else 7)            # config 0: HBFP8, config 3: BFP6
    dtn = a.dtype or ("fp16" if a.kind == "llama" else "fp32")
    tdt = {"fp32": torch.float32, "fp16": torch.float16, "bf16": torch.bfloat16}[dtn]
    kw = dict(num_format="bfp", sparsity_num_format=a.format, rounding_mode=a.rounding, epsilon=1e-8, mant_bits=m, weight_mant_bits=15,
              block_size=64, w_sparsity=True, N=2, M=4, first=a.first, sparsity_mode=a.mode, sparsity_frac=0.5, device=dev)
    res = {"model": a.kind, "rounding": a.rounding, "format": a.format, "mant_bits": m, "block": 64, "sparsity": ("2:4" if a.mode == "structured" else "unstructured 50%") + (" s->q" if a.first == "s" else " q->s"), "dtype": dtn}
    outs = {}
    ref = load_reference()
    for tag, impl in (("ours", ours), ("reference", ref)):
        if impl is None:
            continue
        model, cfg = build(a.kind, a.layers)
        n = swap(model, impl, kw, OPT_TARGETS if a.kind == "opt" else (LLAMA_TARGETS if a.kind == "llama" else None))
        model = model.to(dev).to(tdt)
        g = torch.Generator().manual_seed(1)
        if a.kind in ("opt", "llama"):
            B = a.batch or (8 if a.kind == "opt" else 4)
            inp = dict(input_ids=torch.randint(0, cfg.vocab_size, (B, a.seq), generator=g).to(dev))
        else:
            B = a.batch or 256
            inp = dict(pixel_values=torch.randn(B, 3, 224, 224, generator=g).to(dev).to(tdt))
        with torch.no_grad():
            for _ in range(3):                                  # weight packs, allocator, clocks
                y = model(**inp).logits
            torch.cuda.synchronize(); t0 = time.perf_counter()
            for _ in range(5):
                y = model(**inp).logits
            torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
        outs[tag] = y.float()
        res[tag] = {"swapped_modules": n, "forward_s": dt, "logits_shape": list(y.shape), "finite": bool(torch.isfinite(y).all())}
        print(tag, res[tag], flush=True)
        if a.rounding == "stoc":
            # stochastic rounding: two forwards of the SAME implementation differ by their draws; that spread is the yardstick
            with torch.no_grad():
                y2 = model(**inp).logits.float()
            res[tag]["rel_diff_between_two_of_its_own_forwards"] = float((y2 - outs[tag]).norm() / outs[tag].norm())
            print(tag, "self spread:", res[tag]["rel_diff_between_two_of_its_own_forwards"], flush=True)
        del model
    if "reference" in outs:
        d = (outs["ours"] - outs["reference"])
        res["rel_err_vs_reference"] = float(d.norm() / outs["reference"].norm())
        res["max_abs_err"] = float(d.abs().max())
        res["speedup_vs_reference_on_same_gpu"] = res["reference"]["forward_s"] / res["ours"]["forward_s"]
        print("rel err vs reference:", res["rel_err_vs_reference"], "speed-up:", res["speedup_vs_reference_on_same_gpu"], flush=True)
    if a.out:
        json.dump(res, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
