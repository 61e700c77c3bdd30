"""Concrete defense models -- drop-in for /root/reference/src/defenses/ours/models.py on the CUDA path.

Implemented here: `CelebaIdentityClassifier` (VGG11, models.py:40-58) and `NVAEDefenseModel` (models.py:135-274),
i.e. everything BASELINE configs 1, 2 and 5 touch.  Constructor signatures are the reference's (positional use at
src/experiments/load_defense.py:134-140); `mode` is a keyword-only extension selecting the bf16 tensor-core path or
the exact fp32 path.  Checkpoints are read in the reference's on-disk formats (loading_utils.py:20-26,51-66); a
checkpoint dict may be passed in place of a path (used with synthetic weights).
"""
from __future__ import annotations

import torch
from torch import nn

from ... import ops
from ...nvae_engine import NvaeEngine
from ...nvae_spec import NvaeSpec
from ...vgg_engine import Vgg11Engine
from .abstract_models import BaseClassificationModel, MLVGMDefenseModel, default_mode


def _load_ckpt(path_or_dict):
    if isinstance(path_or_dict, dict):
        return path_or_dict
    return torch.load(path_or_dict, map_location="cpu")


class CelebaIdentityClassifier(BaseClassificationModel, torch.nn.Module):

    def __init__(self, model_path: str, device: str, *, mode: str = None, n_classes: int = 100, image_size: int = 64):
        """
        Wrapper for the CelebA-64 Identities VGG-11 custom model.
        """
        mean = (0.5, 0.5, 0.5)
        std = (0.5, 0.5, 0.5)
        self._mode = mode or default_mode()
        self._image_size = image_size
        super().__init__(model_path, device, mean, std)

    def load_classifier(self, model_path: str, device: str):
        ckpt = _load_ckpt(model_path)           # loading_utils.py:20-26: ckpt['state_dict']
        return Vgg11Engine(ckpt["state_dict"], device, self._mode, in_hw=self._image_size)


class NVAEDefenseModel(MLVGMDefenseModel, torch.nn.Module):

    def __init__(self, classifier: BaseClassificationModel, autoencoder_path: str,
                 interpolation_alphas: tuple, alpha_attenuation: float = 1.0, initial_noise_eps: float = 0.0,
                 apply_gaussian_blur: bool = False, device: str = 'cpu', temperature: float = 0.6, *, mode: str = None):
        """
        Defense model using an NVAE.
        :param temperature: temperature for sampling.
        """
        self.temperature = temperature
        self._mode = mode or default_mode()
        # no need for preprocessing, since it is done directly in NVAE forward pass.
        super().__init__(classifier, autoencoder_path, interpolation_alphas, alpha_attenuation, initial_noise_eps,
                         apply_gaussian_blur, device)
        if len(self.interpolation_alphas) != self.autoencoder.spec.n_latents:
            raise ValueError(f"{len(self.interpolation_alphas)} interpolation alphas for "
                             f"{self.autoencoder.spec.n_latents} latent levels")

    def load_autoencoder(self, model_path: str, device: str):
        ckpt = _load_ckpt(model_path)           # loading_utils.py:51-66
        config = ckpt["configuration"]
        spec = NvaeSpec(config["autoencoder"], config["resolution"])
        return NvaeEngine(ckpt[f"state_dict_temp={self.temperature}"], spec, device, self._mode, self.temperature)

    # ------------------------------------------------------------------
    def _noise_args(self):
        if self._explicit_noise is not None:
            return self._explicit_noise[0], list(self._explicit_noise[1:])
        return None, None

    def purify(self, batch: torch.Tensor) -> torch.Tensor:
        """
        MLVGM encoding procedure to extract the codes.
        :param batch: pre-processed images of shape (B, C, H, W) in [0, 1].
        :return: post_precessed purified reconstructions (B, C, H, W)
        """
        eng = self.autoencoder
        x = ops.nchw_to_nhwc(batch.detach().to(torch.float32), eng.adt, 2.0, -1.0)      # (x-0.5)/0.5, models.py:170
        _, eps_levels = self._noise_args()
        purified, _ = eng.purify(x, self._alphas_device(), eps_levels, self._next_seed(), self.sample_offset)
        return purified

    def _forward_cuda(self, batch: torch.Tensor, tape=None):
        """fused `__call__` body (abstract_models.py:161-193): blur -> noise -> normalise -> purify -> classify."""
        eng = self.autoencoder
        noise0, eps_levels = self._noise_args()
        seed = self._next_seed()
        x, pre = ops.preprocess(batch.to(torch.float32), noise0, float(self.eps), bool(self.blur_input), eng.adt,
                                seed=seed, sample0=self.sample_offset, normalize=True, save_pre=tape is not None,
                                taps_cache=self._taps_cache)
        if tape is not None:
            tape.pre = pre
        purified, cls_in = eng.purify(x, self._alphas_device(), eps_levels, seed, self.sample_offset,
                                      cls_dtype=self.classifier.classifier.adt, tape=tape)
        preds = self.classifier.classifier.forward(cls_in, tape=tape) if tape is not None \
            else self.classifier.classifier.forward(cls_in)
        return preds, purified
