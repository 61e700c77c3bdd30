"""GPU: ablation defenses (reference ablations/models.py) and the alpha-search objective on the fused preprocessing kernel."""
import math

import pytest
import torch
import torch.nn.functional as F

from gen_adversarial_b200 import synth
from gen_adversarial_b200.alpha_schedules import AlphaEvaluator, get_cosine_alphas, get_linear_alphas
from gen_adversarial_b200.defenses.ablations.models import GaussianBlurDefenseModel, GaussianNoiseDefenseModel

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


class _Probe(torch.nn.Module):
    def forward(self, x):
        return x.flatten(1)[:, :10]


def _kornia_blur(x, k):
    t = torch.arange(k, dtype=torch.float64) - k // 2
    g = torch.exp(-(t ** 2) / 2.0)
    g = (g / g.sum()).to(torch.float32)
    c = x.shape[1]
    xp = F.pad(x, (k // 2, k // 2, k // 2, k // 2), mode="reflect")
    y = F.conv2d(xp, g.view(1, 1, 1, k).repeat(c, 1, 1, 1), groups=c)
    return F.conv2d(y, g.view(1, 1, k, 1).repeat(c, 1, 1, 1), groups=c)


def test_ablation_models_match_reference_semantics():
    x = synth.synthetic_batch(5, (3, 64, 64), seed=4)[0]
    blur = GaussianBlurDefenseModel(_Probe())
    k = int(2 ** (math.sqrt(64) // 2) - 1)
    got = blur.purify(x.to(DEV)).cpu()
    assert got.shape == x.shape and (got - _kornia_blur(x, k)).abs().max().item() <= 1e-5
    assert blur(x.to(DEV)).shape == (5, 10)
    noise = torch.randn(x.shape, generator=torch.Generator().manual_seed(1))
    nd = GaussianNoiseDefenseModel(_Probe(), eps=0.5)
    nd.set_explicit_noise(noise)
    ref = (x + noise * (0.5 / noise.flatten(1).norm(dim=1).view(-1, 1, 1, 1))).clamp(0.0, 1.0)
    assert (nd.purify(x.to(DEV)).cpu() - ref).abs().max().item() <= 1e-6
    nd.set_explicit_noise(None)
    a, b = nd.purify(x.to(DEV)), nd.purify(x.to(DEV))
    assert (a - b).abs().max().item() > 0                                     # fresh noise per call
    d = (a.cpu() - x).flatten(1).norm(dim=1)
    assert ((d - 0.5).abs() <= 0.05).all()                                    # L2 norm eps (clamping only shortens it)
    with pytest.raises(RuntimeError):
        nd.purify(x)                                                          # no CPU path


def test_alpha_schedules_and_objective():
    assert get_linear_alphas(4) == [0.25, 0.5, 0.75, 1.0]
    assert abs(get_cosine_alphas(4)[1] - 0.5) < 1e-12 and get_cosine_alphas(4)[-1] == 1.0
    from gen_adversarial_b200.nvae_spec import NvaeSpec, tiny_config
    from gen_adversarial_b200.defenses.ours.models import NVAEDefenseModel, CelebaIdentityClassifier
    cfg, res = tiny_config(initial_channels=16, groups=2, scales=2, latent=4), (3, 32, 32)
    spec = NvaeSpec(cfg, res)
    clf = CelebaIdentityClassifier({"state_dict": synth.make_vgg11_state_dict(10, seed=3, device=DEV)}, DEV, mode="fp32", n_classes=10, image_size=32)
    dm = NVAEDefenseModel(clf, synth.make_nvae_checkpoint(cfg, res, seed=3), [0.0] * spec.n_latents, 0.7, 0.0, False, DEV, mode="fp32")
    x = synth.synthetic_batch(6, res, seed=2)[0].to(DEV)
    y = dm(x).argmax(1)                                                       # alpha = 0: labels the reconstruction keeps
    ev = AlphaEvaluator(dm, [(x, y)], alpha_attenuation=0.7, eot_steps=4, images_per_call=4)
    acc0 = ev.objective_function(torch.zeros(spec.n_latents))
    assert acc0 == 1.0                                                        # alpha 0: no resampling, EoT replicas agree
    acc1 = ev.objective_function(get_cosine_alphas(spec.n_latents))
    assert 0.0 <= acc1 <= 1.0 and dm.interpolation_alphas[-1] == pytest.approx(0.7)


def test_ablation_models_are_differentiable_like_the_reference():
    """white-box attacks on the ablation configs take autograd.grad(loss, [x]) through purify + classifier
    (reference ablations/models.py:21-60 is plain torch/kornia and differentiable): gradient parity against torch autograd
    through the same arithmetic."""
    g = torch.Generator().manual_seed(3)
    x = synth.synthetic_batch(3, (3, 64, 64), seed=9)[0]
    w = torch.randn(3, 3, 64, 64, generator=g)
    k = int(2 ** (math.sqrt(64) // 2) - 1)

    class _Lin(torch.nn.Module):      # any differentiable torch classifier works behind the ablation purifiers
        def forward(self, t):
            return (t * w.to(t.device)).flatten(1).sum(1, keepdim=True)

    # blur
    xr = x.clone().requires_grad_(True)
    g_ref, = torch.autograd.grad(_Lin()(_kornia_blur(xr, k)).sum(), [xr])
    xd = x.to(DEV).requires_grad_(True)
    out = GaussianBlurDefenseModel(_Lin())(xd)
    assert out.requires_grad
    g_got, = torch.autograd.grad(out.sum(), [xd])
    assert (g_got.cpu() - g_ref).abs().max().item() <= 1e-5 * max(1.0, g_ref.abs().max().item())
    # noise + clamp: the gradient is the clamp mask
    noise = torch.randn(x.shape, generator=g)
    xr = x.clone().requires_grad_(True)
    ref = (xr + noise * (3.0 / noise.flatten(1).norm(dim=1).view(-1, 1, 1, 1))).clamp(0.0, 1.0)
    g_ref, = torch.autograd.grad(_Lin()(ref).sum(), [xr])
    nd = GaussianNoiseDefenseModel(_Lin(), eps=3.0)
    nd.set_explicit_noise(noise)
    xd = x.to(DEV).requires_grad_(True)
    g_got, = torch.autograd.grad(nd(xd).sum(), [xd])
    assert 0.0 < (g_ref == 0).float().mean().item() < 1.0                     # some pixels clamp, most do not
    assert (g_got.cpu() - g_ref).abs().max().item() <= 1e-6
    # no-grad calls stay on the plain path
    with torch.no_grad():
        assert not nd(x.to(DEV)).requires_grad
