"""Ablation defenses -- drop-in for /root/reference/src/defenses/ablations/models.py:13-66 on the CUDA path.

Both are the preprocessing stage of the purification call used alone: they run on the same fused kernel
(`ga_preprocess_fwd`: separable reflect-border Gaussian blur / L2-normalised Gaussian noise / clamp, one pass over the image)
and hand the result to the base classifier.  Same class names, constructor arguments and `purify` / `forward` methods as the
reference; noise comes from the in-kernel Philox stream (or from `set_explicit_noise` for parity runs).  `purify` / `forward` are
differentiable w.r.t. the input (white-box attacks on the ablation configs): autograd.py `_PreprocessFn`.
"""
from __future__ import annotations

import torch
from torch import nn

from ... import ops  # noqa: F401
from ...autograd import preprocess_apply


class _AblationBase(nn.Module):
    def __init__(self, base_classifier: nn.Module):
        super().__init__()
        self.base_classifier = base_classifier
        self.noise_seed = None          # None: a fresh seed per call (the reference draws fresh noise on every call, models.py:24)
        self.sample_offset = 0          # global index of sample 0 (data-parallel shards)
        self._explicit_noise = None
        self._taps_cache = {}

    def set_explicit_noise(self, noise):
        """parity hook: the next calls consume this N(0,1) tensor (B,C,H,W) instead of the Philox stream; None switches back"""
        self._explicit_noise = None if noise is None else noise.to(torch.float32).contiguous()

    def _seed(self) -> int:
        return int(self.noise_seed) if self.noise_seed is not None else int(torch.empty((), dtype=torch.int64).random_().item())

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.base_classifier(self.purify(x))


class GaussianNoiseDefenseModel(_AblationBase):

    def __init__(self, base_classifier: nn.Module, eps: float = 0.5):
        super().__init__(base_classifier)
        self.eps = eps

    def purify(self, x: torch.Tensor) -> torch.Tensor:
        """x + N(0,1) noise scaled to L2 norm eps per sample, clamped to [0, 1] (ablations/models.py:21-34)"""
        if not x.is_cuda:
            raise RuntimeError("GaussianNoiseDefenseModel.purify: CUDA tensor expected (there is no CPU path)")
        noise = self._explicit_noise.to(x.device) if self._explicit_noise is not None else None
        return preprocess_apply(x, noise, float(self.eps), False, self._seed(), self.sample_offset)


class GaussianBlurDefenseModel(_AblationBase):

    def __init__(self, base_classifier: nn.Module):
        super().__init__(base_classifier)

    def purify(self, x: torch.Tensor) -> torch.Tensor:
        """Gaussian blur, sigma 1, kernel 2^(sqrt(h)//2) - 1, reflect border (ablations/models.py:48-60)"""
        if not x.is_cuda:
            raise RuntimeError("GaussianBlurDefenseModel.purify: CUDA tensor expected (there is no CPU path)")
        return preprocess_apply(x, None, 0.0, True, 0, 0, self._taps_cache)
