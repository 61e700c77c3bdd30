"""Alpha schedules and the alpha-search objective on the CUDA path.

Mirrors /root/reference/src/experiments/alpha_learning/common_utils.py:15-22 (`get_linear_alphas`, `get_cosine_alphas`) and the
objective of `AlphaEvaluator.objective_function` (:81-103): accuracy of the EoT-averaged defended classifier on a fixed set of
(adversarial) images for a candidate alpha vector.  The reference walks the set one image at a time (batch 1 x EoT 32 with a
host sync per image); here a whole batch of images goes through the batched EoT wrapper and the only device->host read is the
final count.
"""
from __future__ import annotations

import math
from typing import Iterable, Sequence, Tuple

import torch

from .defenses.wrappers import EoTWrapper


def get_linear_alphas(n: int) -> list:
    return [i / n for i in range(1, n + 1)]


def get_cosine_alphas(n: int) -> list:
    return [0.5 * (1 - math.cos(math.pi * (i / n))) for i in range(1, n + 1)]


class AlphaEvaluator:
    """objective of the alpha search: `defense_model.interpolation_alphas = alphas * attenuation`, then accuracy over the set"""

    def __init__(self, defense_model, batches: Iterable[Tuple[torch.Tensor, torch.Tensor]], alpha_attenuation: float = 1.0,
                 eot_steps: int = 32, images_per_call: int = 16):
        self.alpha_attenuation = float(alpha_attenuation)
        self.eot_steps = int(eot_steps)
        self.images_per_call = int(images_per_call)
        self.defense_model = EoTWrapper(defense_model, self.eot_steps).eval()
        self.batches = list(batches)

    @torch.no_grad()
    def objective_function(self, alphas: Sequence[float]) -> float:
        alphas = alphas.detach().cpu().tolist() if isinstance(alphas, torch.Tensor) else list(alphas)
        self.defense_model.model.interpolation_alphas = [a * self.alpha_attenuation for a in alphas]      # common_utils.py:88
        correct = None
        total = 0
        for x, y in self.batches:
            for i in range(0, x.shape[0], self.images_per_call):
                xb, yb = x[i:i + self.images_per_call], y[i:i + self.images_per_call]
                preds = self.defense_model(xb).argmax(dim=1)
                c = (preds == yb.to(preds.device)).sum()
                correct = c if correct is None else correct + c
                total += xb.shape[0]
        return float(correct.item()) / max(total, 1) if correct is not None else 0.0
