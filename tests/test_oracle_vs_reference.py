"""CPU, build container only: the oracle restatement against the UNMODIFIED reference imported from
/root/reference (skipped on the GPU box, where the tree does not exist)."""
import math
import os

import pytest
import torch

from oracle import ref_import, nvae_ref
from gen_adversarial_b200 import synth
from gen_adversarial_b200.nvae_spec import NvaeSpec, NVAE_C32_CONFIG, NVAE_C32_RESOLUTION, tiny_config

pytestmark = pytest.mark.skipif(not ref_import.reference_available(), reason="reference tree not mounted")


class _MeanClassifier:
    def set_device(self, d):
        pass

    def __call__(self, x):
        return x.mean(dim=(2, 3))


@pytest.mark.parametrize("cfg,res", [(tiny_config(), (3, 32, 32)), (NVAE_C32_CONFIG, NVAE_C32_RESOLUTION)])
def test_state_dict_layout_matches_reference(cfg, res):
    ae = ref_import.ref_nvae_module().AutoEncoder(cfg, res)
    ref_sd = ae.state_dict()
    mine = synth.make_nvae_state_dict(cfg, res, seed=0)
    assert set(ref_sd) == set(mine)
    for k in ref_sd:
        assert tuple(ref_sd[k].shape) == tuple(mine[k].shape), k
    ae.load_state_dict(mine, strict=True)      # loading_utils.py:63


def test_restatement_matches_reference_call(tmp_path):
    cfg, res = tiny_config(initial_channels=8, groups=3, scales=3, latent=6), (3, 64, 64)
    spec = NvaeSpec(cfg, res)
    ckpt = synth.make_nvae_checkpoint(cfg, res, seed=5)
    path = os.path.join(tmp_path, "nvae.pt")
    torch.save(ckpt, path)
    mm = ref_import.ref_models()
    n = spec.n_latents
    alphas = [0.5 * (1 - math.cos(math.pi * i / n)) for i in range(1, n + 1)]
    dm = mm.NVAEDefenseModel(_MeanClassifier(), path, alphas, 0.7, 2.0, True, "cpu")
    x, _ = synth.synthetic_batch(2, res, seed=1)
    noises = synth.synthetic_noise(spec, 2, seed=2)
    with torch.no_grad(), ref_import.ExplicitNoise(noises):
        _, pur_ref = dm(x, preds_only=False)
    with torch.no_grad():
        _, pur = nvae_ref.defense_call(ckpt["state_dict_temp=0.6"], spec, None, x, [a * 0.7 for a in alphas],
                                       noises, 2.0, True)
    assert (pur - pur_ref).abs().max().item() <= 1e-5


def test_restatement_gradient_matches_reference(tmp_path):
    """input-gradient of the oracle == autograd through the reference (attack path, untargeted.py:146)."""
    cfg, res = tiny_config(), (3, 32, 32)
    spec = NvaeSpec(cfg, res)
    ckpt = synth.make_nvae_checkpoint(cfg, res, seed=6)
    path = os.path.join(tmp_path, "nvae.pt")
    torch.save(ckpt, path)
    mm = ref_import.ref_models()
    n = spec.n_latents
    alphas = [i / n for i in range(1, n + 1)]
    dm = mm.NVAEDefenseModel(_MeanClassifier(), path, alphas, 1.0, 1.0, True, "cpu")
    x, _ = synth.synthetic_batch(2, res, seed=3)
    noises = synth.synthetic_noise(spec, 2, seed=4)
    w = torch.randn(2, 3, 32, 32)
    xr = x.clone().requires_grad_(True)
    with ref_import.ExplicitNoise(noises):
        _, pur_ref = dm(xr, preds_only=False)
    g_ref, = torch.autograd.grad((pur_ref * w).sum(), [xr])
    xo = x.clone().requires_grad_(True)
    _, pur = nvae_ref.defense_call(ckpt["state_dict_temp=0.6"], spec, None, xo, alphas, noises, 1.0, True)
    g, = torch.autograd.grad((pur * w).sum(), [xo])
    assert (g - g_ref).abs().max().item() <= 1e-5 * max(1.0, g_ref.abs().max().item())
