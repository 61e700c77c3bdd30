"""CPU: StyleGAN2 generator host logic (weight folds, modulation-as-scaling, up-conv + blur as 4 phase convs, batched
mapping MLP, code mixing) through the torch emulation of the kernels, against the UNMODIFIED reference Generator
(skipped where /root/reference is absent) and against the committed fixture produced by it."""
import os

import pytest
import torch

from gen_adversarial_b200 import synth, stylegan_engine
from oracle import ref_import
from tests import emu_ops

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "stylegan2_gen32.pt")


@pytest.fixture()
def emu(monkeypatch):
    monkeypatch.setattr(stylegan_engine, "ops", emu_ops)
    return emu_ops


def _inputs(n_latent, b=2, seed=0):
    g = torch.Generator().manual_seed(seed)
    latent = torch.randn(b, n_latent, 512, generator=g) * 0.7
    z = torch.randn(n_latent, b, 512, generator=g)
    return latent, z


@pytest.mark.skipif(not ref_import.reference_available(), reason="reference tree not mounted")
@pytest.mark.parametrize("mode,tol", [("fp32", 2e-4), ("bf16", 6e-2)])
def test_generator_matches_reference(emu, mode, tol):
    import importlib
    ref_import.install()
    gen = importlib.import_module("src.mlvgms_autoencoders.StyleGan_E4E.stylegan2.generator")
    sd = synth.make_stylegan2_state_dict(32, seed=2)
    G = gen.Generator(32, 512, 8, channel_multiplier=2).eval()
    G.load_state_dict(sd, strict=True)
    latent, z = _inputs(G.n_latent)
    with torch.no_grad():
        img_ref, _ = G([latent], input_is_latent=True, randomize_noise=False)
        w_ref = torch.stack([G.style(n) for n in z], dim=0)                       # models.py:120
    eng = stylegan_engine.StyleGan2Engine(sd, 32, "cpu", mode, _host_logic_test=True)
    img = eng.decode(latent, pool=1)
    w = eng.mapping(z.reshape(-1, 512)).reshape(z.shape)
    scale = img_ref.abs().max().item()
    assert (img - img_ref).abs().max().item() <= tol * scale, ((img - img_ref).abs().max().item(), scale)
    assert (w - w_ref).abs().max().item() <= tol * max(1.0, w_ref.abs().max().item())
    # per-level mixing (models.py:117-127)
    alphas = torch.linspace(0, 1, G.n_latent)
    mixed_ref = ((1 - alphas.view(-1, 1, 1)) * latent.permute(1, 0, 2) + alphas.view(-1, 1, 1) * w_ref).permute(1, 0, 2)
    mixed = eng.mix_codes(latent, z, alphas)
    assert (mixed - mixed_ref).abs().max().item() <= tol * max(1.0, mixed_ref.abs().max().item())


def test_generator_matches_reference_fixture(emu):
    if not os.path.exists(GOLDEN):
        pytest.skip("fixture not generated")
    g = torch.load(GOLDEN, weights_only=True)
    sd = synth.make_stylegan2_state_dict(32, seed=g["seed"])
    eng = stylegan_engine.StyleGan2Engine(sd, 32, "cpu", "fp32", _host_logic_test=True)
    img = eng.decode(g["latent"], pool=2)
    scale = g["image_pool2"].abs().max().item()
    assert (img - g["image_pool2"]).abs().max().item() <= 2e-4 * scale


def test_superpixel_weights_are_the_same_convolution():
    """stylegan_engine.superpixel_weights: a 3x3 conv on [h][w][c] equals the transformed conv on the [h][w/2][2c] view (exact in fp64)"""
    import torch
    import torch.nn.functional as F
    from gen_adversarial_b200.stylegan_engine import superpixel_weights
    g = torch.Generator().manual_seed(0)
    for cin, cout, h, w in ((4, 6, 5, 8), (32, 32, 6, 12), (3, 2, 4, 2)):
        wt = torch.randn(cout, cin, 3, 3, generator=g, dtype=torch.float64)
        x = torch.randn(2, h, w, cin, generator=g, dtype=torch.float64)                      # NHWC
        ref = F.conv2d(x.permute(0, 3, 1, 2), wt, padding=1).permute(0, 2, 3, 1)              # [n, h, w, cout]
        xs = x.reshape(2, h, w // 2, 2 * cin)
        got = F.conv2d(xs.permute(0, 3, 1, 2), superpixel_weights(wt), padding=1).permute(0, 2, 3, 1).reshape(2, h, w, cout)
        assert (got - ref).abs().max().item() <= 1e-12
