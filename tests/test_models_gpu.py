"""-m gpu tests at the BASELINE configurations beyond the quantiser sweep (BASELINE.json configs 1, 3, 4, 5), against the LIVE
reference (the unmodified bfp_ops.py / int_ops.py staged in baseline/_ref, executed with torch-CUDA on the same GPU):

  config 1  OPT-125M random-init, HBFP8 block 64 + 2:4 on every nn.Linear, 8 x 512 tokens        -> logits identical
  config 4  ViT-B/16 random-init, BFP6 + 2:4 on every linear + patch-embedding conv                 -> logits identical
  config 3  LLaMA-2-13B BFP linear forward, 4096 tokens (q / gate / down shapes)                    -> <= 1e-5 of the reference's
  config 5  LLaMA-65B BFP linear forward shapes                                                         F.linear(Q_in(x), Q_w(w))

BFP operands make every partial sum of the contraction exact in fp32 (integers times one power of two),
so whole-model logits can be compared for EQUALITY, not just closeness.  The models are scaled down in depth only where the
reference itself would take minutes (the reference quantiser is ~25 eager kernels + a host-built mask per call); widths, head
counts, sequence lengths and batch are the BASELINE ones.  Skipped when baseline/_ref is absent."""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


@pytest.fixture(scope="module")
def impls():
    from _refload import load_reference
    ref = load_reference()
    if ref is None:
        pytest.skip("reference sources not available (baseline/_ref)")
    from qsi_b200 import bfp_ops
    return bfp_ops, ref


def _kw(m, **over):
    kw = dict(num_format="bfp", sparsity_num_format="bfp", rounding_mode="determ", epsilon=1e-8, mant_bits=m, weight_mant_bits=15,
              block_size=64, w_sparsity=True, N=2, M=4, first="s", sparsity_mode="structured", sparsity_frac=0.5, device="cuda")
    kw.update(over)
    return kw


def _logits(kind, impl, kw, layers, inp_fn):
    import model_dropin as md
    model, cfg = md.build(kind, layers)
    n = md.swap(model, impl, kw, md.OPT_TARGETS if kind == "opt" else None)
    model = model.cuda()
    with torch.no_grad():
        y = model(**inp_fn(cfg)).logits.float()
    del model
    torch.cuda.empty_cache()
    return y, n


def test_config1_opt125m_logits_identical_to_reference(impls):
    """BASELINE configs[0]: OPTConfig() defaults = OPT-125M (768 / 3072 / 12 heads / vocab 50272), 8 x 512 tokens, all 12 layers."""
    ours, ref = impls
    g = torch.Generator().manual_seed(1)

    def inp(cfg):
        return dict(input_ids=torch.randint(0, cfg.vocab_size, (8, 512), generator=torch.Generator().manual_seed(1)).cuda())
    y_ours, n_ours = _logits("opt", ours, _kw(7), 0, inp)
    y_ref, n_ref = _logits("opt", ref, _kw(7), 0, inp)
    assert n_ours == n_ref == 72 and y_ours.shape == (8, 512, 50272)
    assert torch.isfinite(y_ref).all()
    assert torch.equal(y_ours, y_ref), float((y_ours - y_ref).norm() / y_ref.norm())
    del g


def test_config4_vit_b16_logits_identical_to_reference(impls):
    """BASELINE configs[3]: ViTConfig() defaults = ViT-B/16, BFP6 (mant_bits 5) + 2:4 in all linears and the patch-embedding conv.
    Batch 32 of the 256 (the reference needs ~0.5 s per forward at 256; the shapes per image are the same), all 12 layers."""
    ours, ref = impls

    def inp(cfg):
        return dict(pixel_values=torch.randn(32, 3, 224, 224, generator=torch.Generator().manual_seed(1)).cuda())
    prev = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False          # the reference's fp32 conv may otherwise run in TF32
    try:
        y_ours, n_ours = _logits("vit", ours, _kw(5), 0, inp)
        y_ref, n_ref = _logits("vit", ref, _kw(5), 0, inp)
    finally:
        torch.backends.cudnn.allow_tf32 = prev
    assert n_ours == n_ref and n_ours >= 12 * 6 + 2 and y_ours.shape == y_ref.shape and y_ours.shape[0] == 32
    rel = float((y_ours - y_ref).norm() / y_ref.norm())
    assert rel <= 1e-6, rel                            # measured 0.0; softmax / layernorm are the same torch kernels in both


@pytest.mark.parametrize("NK", [(5120, 5120), (13824, 5120), (5120, 13824),            # LLaMA-2-13B: q/k/v/o, gate/up, down
                                (8192, 8192), (22016, 8192), (8192, 22016)])           # LLaMA-65B
def test_config3_and_5_bfp_linear_forward_vs_reference_linear(impls, NK):
    """One BFPLinear forward at 4096 tokens against the reference's own BFPLinear (fake-quantise both operands, fp32 F.linear) on
    the same GPU: relative error <= 1e-5 (north_star); the rows are sampled for the largest shapes to bound the reference's time."""
    ours, ref = impls
    N, K = NK
    T = 4096
    g = torch.Generator(device="cuda").manual_seed(N + K)
    w = torch.randn(N, K, device="cuda", generator=g) * 0.02
    x = torch.randn(T, K, device="cuda", generator=g)
    x.view(-1)[::1013] *= 20.0                                         # activation outliers
    lin = ours.BFPLinear(K, N, bias=False, **_kw(7)).cuda().eval()
    with torch.no_grad():
        lin.weight.copy_(w)
        y = lin(x)
        assert lin._packed_w[0][0] == "sp"
        rows = slice(0, N) if N * K <= 5120 * 13824 else slice(N // 2 - 1024, N // 2 + 1024)      # 2048 output features of the 65B shapes
        rl = ref.BFPLinear(K, (rows.stop - rows.start), bias=False, **_kw(7)).cuda()
        rl.weight.copy_(w[rows])
        y_ref = rl(x)
    rel = float((y[:, rows].double() - y_ref.double()).norm() / y_ref.double().norm())
    assert rel <= 1e-5, rel
    assert torch.equal(y[:, rows], y_ref) or rel <= 1e-6                # in practice identical: every partial sum is exact
