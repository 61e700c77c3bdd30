"""GPU: StyleGAN2 generator kernels (fused_bias_act / upfirdn2d replacements, modulation glue) against their torch
restatements, and the generator engine against the fixture produced by the reference Generator."""
import os

import pytest
import torch

from gen_adversarial_b200 import ops, synth
from gen_adversarial_b200._lib import ACT_NONE, ACT_LRELU_SQRT2
from gen_adversarial_b200.stylegan_engine import StyleGan2Engine
from tests import emu_ops

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "stylegan2_gen32.pt")


def test_pixelnorm_demod_scale_lerp():
    g = torch.Generator().manual_seed(0)
    z = torch.randn(37, 512, generator=g)
    assert (ops.pixelnorm(z.to(DEV), torch.float32).cpu() - emu_ops.pixelnorm(z, torch.float32)).abs().max().item() <= 1e-5
    s, wsq = torch.randn(3, 96, generator=g), torch.rand(40, 96, generator=g)
    assert (ops.style_demod(s.to(DEV), wsq.to(DEV)).cpu() - emu_ops.style_demod(s, wsq)).abs().max().item() <= 1e-5
    x = torch.randn(3, 8, 8, 96, generator=g)
    assert (ops.channel_scale(x.to(DEV), s.to(DEV), torch.float32).cpu() - emu_ops.channel_scale(x, s, torch.float32)).abs().max().item() <= 1e-6
    codes, styles, a = torch.randn(3, 6, 512, generator=g), torch.randn(3, 6, 512, generator=g), torch.rand(6, generator=g)
    assert (ops.latent_lerp(codes.to(DEV), styles.to(DEV), a.to(DEV)).cpu() - emu_ops.latent_lerp(codes, styles, a)).abs().max().item() <= 1e-6


@pytest.mark.parametrize("phases", [False, True])
def test_styled_bias_act(phases):
    """fused demod + noise + bias + leaky_relu*sqrt(2) (= fused_bias_act_kernel.cu + NoiseInjection)"""
    g = torch.Generator().manual_seed(1)
    n, h, w, c = 2, 16, 16, 32
    y = torch.randn((n, h // 2, w // 2, 4 * c) if phases else (n, h, w, c), generator=g)
    demod, noise, bias = torch.rand(n, c, generator=g) + 0.5, torch.randn(h, w, generator=g), torch.randn(c, generator=g)
    skip = torch.randn(n, h, w, c, generator=g)
    for act, d, nz, sk in ((ACT_LRELU_SQRT2, demod, noise, None), (ACT_NONE, None, None, skip)):
        ref = emu_ops.styled_bias_act(y, phases, d, nz, 0.37, bias, act, sk, torch.float32)
        got = ops.styled_bias_act(y.to(DEV), phases, d.to(DEV) if d is not None else None, nz.to(DEV) if nz is not None else None,
                                  0.37, bias.to(DEV), act, sk.to(DEV) if sk is not None else None, torch.float32)
        assert (got.cpu() - ref).abs().max().item() <= 1e-5
    # consumers' modulation applied in the same pass (two scaled copies) and the half-resolution RGB skip up-sampled on the fly
    sa, sb = torch.rand(n, c, generator=g) + 0.5, torch.rand(n, c, generator=g) + 0.5
    ra, rb = emu_ops.styled_bias_act(y, phases, demod, noise, 0.37, bias, ACT_LRELU_SQRT2, None, torch.float32, scale_a=sa, scale_b=sb)
    ga, gb = ops.styled_bias_act(y.to(DEV), phases, demod.to(DEV), noise.to(DEV), 0.37, bias.to(DEV), ACT_LRELU_SQRT2, None, torch.float32,
                                 scale_a=sa.to(DEV), scale_b=sb.to(DEV))
    assert (ga.cpu() - ra).abs().max().item() <= 1e-5 and (gb.cpu() - rb).abs().max().item() <= 1e-5
    if not phases:
        k = torch.tensor([1.0, 3.0, 3.0, 1.0]); k = k[None] * k[:, None]; k = k / k.sum() * 4
        low = torch.randn(n, h // 2, w // 2, c, generator=g)
        ref = emu_ops.styled_bias_act(y, False, None, None, 0.0, bias, ACT_NONE, low, torch.float32, skip_up_kernel=k)
        got = ops.styled_bias_act(y.to(DEV), False, None, None, 0.0, bias.to(DEV), ACT_NONE, low.to(DEV), torch.float32, skip_up_kernel=k.to(DEV))
        assert (got.cpu() - ref).abs().max().item() <= 1e-5


@pytest.mark.parametrize("ydt", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("phases", [False, True])
@pytest.mark.parametrize("n,h,w,c", [(3, 16, 16, 32), (2, 32, 64, 8), (5, 8, 8, 256), (1, 64, 32, 64)])
def test_styled_bias_act_vector_path(phases, ydt, n, h, w, c):
    """the generator's big layers take the 8-channel / 4-items-per-thread kernel (bf16 outputs, lrelu * sqrt(2), power-of-two H, W, C / 8):
    same arithmetic in the same order as the general kernel -- compared with the torch restatement at one bf16 ulp of the rounded outputs --
    for plain and phase-interleaved inputs (bf16 / fp32 conv outputs), one or two modulated copies, ragged grid tails"""
    g = torch.Generator().manual_seed(n + h + c)
    y = torch.randn((n, h // 2, w // 2, 4 * c) if phases else (n, h, w, c), generator=g).to(ydt)
    demod, noise, bias = torch.rand(n, c, generator=g) + 0.5, torch.randn(h, w, generator=g), torch.randn(c, generator=g)
    sa, sb = torch.rand(n, c, generator=g) + 0.5, torch.rand(n, c, generator=g) + 0.5
    yd, dd, nd, bd, sad, sbd = (t.to(DEV) for t in (y, demod, noise, bias, sa, sb))
    for kw_ref, kw_got in (({}, {}), ({"scale_a": sa}, {"scale_a": sad}), ({"scale_a": sa, "scale_b": sb}, {"scale_a": sad, "scale_b": sbd}),
                           ({"scale_b": sb, "want_out": False}, {"scale_b": sbd, "want_out": False})):
        ref = emu_ops.styled_bias_act(y.float(), phases, demod, noise, 0.37, bias, ACT_LRELU_SQRT2, None, torch.float32, **kw_ref)
        got = ops.styled_bias_act(yd, phases, dd, nd, 0.37, bd, ACT_LRELU_SQRT2, None, torch.bfloat16, **kw_got)
        ref = ref if isinstance(ref, tuple) else (ref,)
        got = got if isinstance(got, tuple) else (got,)
        for r_, g_ in zip(ref, got):
            if r_ is None:
                assert g_ is None
                continue
            assert g_.dtype == torch.bfloat16 and g_.shape == (n, h, w, c)
            err = (g_.float().cpu() - r_).abs()
            assert (err <= 8e-3 * r_.abs().clamp_min(1.0)).all(), err.max().item()


@pytest.mark.parametrize("with_skip", [False, True])
@pytest.mark.parametrize("n,h,w,cin", [(3, 16, 16, 32), (2, 8, 32, 64), (1, 4, 4, 512), (5, 32, 16, 8), (2, 64, 64, 40)])
def test_torgb_fused(n, h, w, cin, with_skip):
    """ToRGB (generator.py:271-292) in one pass: 1x1 conv to the padded RGB channels + bias + Upsample(skip), against the torch restatement
    (bf16 inputs and weights, fp32 accumulation: only the summation order differs)"""
    from gen_adversarial_b200 import ops as _ops
    g = torch.Generator().manual_seed(n + h + cin)
    x = torch.randn(n, h, w, cin, generator=g).to(torch.bfloat16)
    L = _ops.ConvLayer(1, 1, 1, 0, cin, 4, name="to_rgb")
    w4 = torch.zeros(4, cin)
    w4[:3] = torch.randn(3, cin, generator=g) / cin ** 0.5
    L.w_tc = w4.to(torch.bfloat16)
    bias = torch.tensor([0.1, -0.2, 0.3, 0.0])
    k = torch.tensor([1.0, 3.0, 3.0, 1.0]); k = k[None] * k[:, None]; k = k / k.sum() * 4
    skip = torch.randn(n, h // 2, w // 2, 4, generator=g) if with_skip else None
    ref = emu_ops.torgb_fused(x, L, bias, skip, k if with_skip else None)
    LD = _ops.ConvLayer(1, 1, 1, 0, cin, 4, name="to_rgb")
    LD.w_tc = L.w_tc.to(DEV)
    got = _ops.torgb_fused(x.to(DEV), LD, bias.to(DEV), skip.to(DEV) if with_skip else None, k.to(DEV) if with_skip else None)
    torch.cuda.synchronize()
    assert got.shape == (n, h, w, 4) and got.dtype == torch.float32
    assert (got.cpu() - ref).abs().max().item() <= 2e-5 * max(1.0, ref.abs().max().item())
    assert (got[..., 3].cpu() - (ref[..., 3])).abs().max().item() <= 1e-6          # the padding channel stays bias + skip


@pytest.mark.parametrize("up,down,pad", [(1, 1, (1, 1)), (2, 1, (2, 1)), (1, 2, (2, 2)), (1, 1, (2, 1))])
def test_upfirdn2d_matches_reference_semantics(up, down, pad):
    g = torch.Generator().manual_seed(2)
    x = torch.randn(2, 9, 9, 4, generator=g)
    k = torch.tensor([1.0, 3.0, 3.0, 1.0])
    k = (k[None] * k[:, None]); k = k / k.sum() * (up ** 2)
    ref = emu_ops.upfirdn2d(x, k, up, down, pad, torch.float32)
    got = ops.upfirdn2d(x.to(DEV), k.to(DEV), up, down, pad, torch.float32)
    assert got.shape == ref.shape and (got.cpu() - ref).abs().max().item() <= 1e-5


@pytest.mark.parametrize("mode,tol", [("fp32", 2e-4), ("bf16", 6e-2)])
def test_generator_matches_reference_fixture(mode, tol):
    """`Generator([latent], input_is_latent=True, randomize_noise=False)` + `Generator.style` of the reference (fixture)"""
    g = torch.load(GOLDEN, weights_only=True)
    sd = synth.make_stylegan2_state_dict(g["size"], seed=g["seed"])
    eng = StyleGan2Engine(sd, g["size"], DEV, mode)
    img = eng.decode(g["latent"].to(DEV), pool=1).cpu()
    scale = g["image"].abs().max().item()
    err = (img - g["image"]).abs().max().item()
    w = eng.mapping(g["z"].reshape(-1, 512).to(DEV)).reshape(g["z"].shape).cpu()
    werr = (w - g["w"]).abs().max().item()
    print(f"[{mode}] generator@32: image max-abs err {err:.3e} (range {scale:.2f}); mapping err {werr:.3e} (range {g['w'].abs().max():.2f}); "
          f"kernels {ops.launch_count(True)}")
    assert err <= tol * scale
    assert werr <= tol * max(1.0, g["w"].abs().max().item())
    pooled = eng.decode(g["latent"].to(DEV), pool=2).cpu()
    assert (pooled - g["image_pool2"]).abs().max().item() <= tol * scale
