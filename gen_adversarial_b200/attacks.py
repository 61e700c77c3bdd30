"""PGD-Linf attack on the CUDA path (BASELINE config 5).

The reference has NO PGD-Linf attack class (its attacks are L2: src/attacks/untargeted.py); the update rule restated
here is the only L-inf PGD step in the reference tree, `src/defenses/competitors/trades/modules.py:43-45`:
    x_adv <- x_adv + a * sign(grad);  x_adv <- min(max(x_adv, x - eps), x + eps);  x_adv <- clamp(x_adv, 0, 1)
Documented choices (parity for this config is unpinned): cross-entropy loss on the defended classifier,
eps = 8/255, 50 steps (BASELINE), step a = 2/255, start at the clean image (`random_start` adds 0.001*N(0,1) as
trades/modules.py:35 does).  The whole inner update is ONE fused kernel (ga_pgd_linf_step); loss + dlogits + accuracy
counting is another (ga_softmax_xent); the gradient comes from the dgrad-only backward sweep of the engines.
"""
from __future__ import annotations

import torch

from . import ops


class PGDLinf:
    def __init__(self, eps: float = 8 / 255, step: float = 2 / 255, steps: int = 50, random_start: bool = False):
        self.eps, self.step, self.steps, self.random_start = float(eps), float(step), int(steps), random_start

    def __call__(self, images: torch.Tensor, labels: torch.Tensor, net, noise_schedule=None):
        """images (B,3,H,W) in [0,1], labels (B,), net: defense model.  noise_schedule: optional list (one entry per
        step, plus one for the final evaluation) of explicit-noise lists for parity runs.
        -> (success (B,) bool, linf (B,) fp32, adversarial images (B,3,H,W))"""
        x = images.detach().to(torch.float32).contiguous()
        x_adv = x.clone()
        if self.random_start:
            x_adv = (x_adv + 0.001 * torch.randn_like(x_adv)).contiguous()
        fused = hasattr(net, "loss_input_grad")
        if fused and noise_schedule is None and getattr(net, "use_cuda_graph", False) and x.is_cuda:
            # launch-rate bound at attack batch sizes (~1500 launches per iteration): replay one captured iteration `steps` times
            from .graphs import GraphedPGD, GraphedForward
            key = ("pgd", tuple(x.shape), self.step, self.eps)
            g = net._graphs.get(key)
            if g is None or g.key[3] != GraphedForward.make_key(net, x):
                g = net._graphs[key] = GraphedPGD(net, x, labels, self.step, self.eps)
            x_adv = g.run(net, x, labels, self.steps, x_adv if self.random_start else None).clone()
            steps_left = 0
        else:
            steps_left = self.steps
        for i in range(steps_left):
            if noise_schedule is not None:
                net.set_explicit_noise(noise_schedule[i])
            if fused:
                _, grad, _ = net.loss_input_grad(x_adv, labels)
            else:   # any other differentiable torch module (e.g. the EoT wrapper): torch autograd drives our backward
                xa = x_adv.clone().requires_grad_(True)
                loss = torch.nn.functional.cross_entropy(net(xa), labels)
                grad, = torch.autograd.grad(loss, [xa])
                grad = grad.contiguous()
            ops.pgd_linf_step_(x_adv, grad, x, self.step, self.eps)
        if noise_schedule is not None:
            net.set_explicit_noise(noise_schedule[self.steps])
        with torch.no_grad():
            preds = net(x_adv)
        success = preds.argmax(dim=1) != labels
        linf = (x_adv - x).abs().flatten(1).max(dim=1).values
        return success, linf, x_adv
