"""Batched attack drivers on the CUDA path: PGD-Linf (BASELINE config 5) and the batched forms of the reference's APGD / FGSM.

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


def _loss_and_grad(net, x_adv: torch.Tensor, labels: torch.Tensor, criterion):
    """per-image loss (B,) and d sum(loss) / d x_adv (B,C,H,W).  Defense models of this package with a cross-entropy criterion go through the
    fused primitive (no autograd engine); anything else (EoT wrapper, torch classifiers, DLR loss) through torch.autograd.grad on net(x)."""
    if criterion is None and hasattr(net, "loss_input_grad"):
        loss, grad, _ = net.loss_input_grad(x_adv, labels)
        return loss, grad * float(x_adv.shape[0])          # the primitive differentiates the MEAN loss; APGD sums per-image losses
    xa = x_adv.detach().clone().requires_grad_(True)
    with torch.enable_grad():
        logits = net(xa)
        loss = torch.nn.functional.cross_entropy(logits, labels, reduction="none") if criterion is None else criterion(logits, labels)
        grad, = torch.autograd.grad(loss.sum(), [xa])
    return loss.detach(), grad.detach().contiguous()


def dlr_loss(logits: torch.Tensor, labels: torch.Tensor, division_eps: float = 1e-12) -> torch.Tensor:
    """Difference-of-Logits-Ratio loss of APGD-DLR, batched restatement of src/attacks/untargeted.py:87-125 -> (B,)"""
    if logits.shape[1] < 4:
        raise AttributeError('APGD_DLR is undefined for problems with less than 4 classes!')
    srt, idx = logits.sort(dim=1)
    failed = idx[:, -1] == labels
    correct = logits.gather(1, labels.view(-1, 1)).squeeze(1)
    highest_wrong = torch.where(failed, srt[:, -2], srt[:, -1])
    normalizer = torch.where(srt[:, -3] != correct, srt[:, -3], srt[:, -4])
    return -(correct - highest_wrong) / (srt[:, -1] - normalizer + division_eps)


class APGDL2:
    """Batched APGD-CE / APGD-DLR, L2-bounded, untargeted: the algorithm of the reference's `APGDAttack`
    (/root/reference/src/attacks/untargeted.py:37-243) with every image of the batch carrying its own state (step size, best loss / iterate /
    gradient, step-size-reduction flags) in device tensors -- no `.item()` sync inside the loop -- and the whole per-iteration update
    (two projections on the L2 ball, momentum, clamps, per-image norms) in ONE kernel (ga_apgd_l2_step).  The checkpoint schedule
    (after how many iterations the step size is reconsidered) depends only on n_iter, so it is shared by the batch.
    Same constructor arguments as the reference class; __call__ takes a batch and returns per-image results."""

    def __init__(self, n_iter: int, rho: float, max_bound: float, ce_loss: bool = True):
        self.n_iter, self.rho, self.max_bound = int(n_iter), float(rho), float(max_bound)
        self.criterion = None if ce_loss else dlr_loss
        self.initial_step_size_iters = max(int(0.22 * n_iter), 1)
        self.min_step_size_iters = max(int(0.06 * n_iter), 1)
        self.step_size_decr = max(int(0.03 * n_iter), 1)

    def __call__(self, images: torch.Tensor, labels: torch.Tensor, net, initial_noise: torch.Tensor = None):
        """images (B,3,H,W) in [0,1], labels (B,).  initial_noise: optional N(0,1) tensor of the images' shape (parity runs; the reference
        draws it with torch.randn_like).  -> (success (B,) bool, l2 bound (B,) fp32, adversarial images (B,3,H,W))"""
        x = images.detach().to(torch.float32).contiguous()
        b = x.shape[0]
        noise = torch.randn_like(x) if initial_noise is None else initial_noise.to(x.device, torch.float32).contiguous()
        x_adv = ops.l2_ball_start(x, noise, self.max_bound)
        x_adv_old = x_adv.clone()
        loss, grad = _loss_and_grad(net, x_adv, labels, self.criterion)
        step_size = torch.full((b,), 2.0 * self.max_bound, device=x.device, dtype=torch.float32)
        loss_steps = torch.zeros((self.n_iter, b), device=x.device, dtype=torch.float32)
        reduced_last = torch.ones((b,), device=x.device, dtype=torch.bool)
        best_loss, prev_best = loss.clone(), loss.clone()
        x_best, grad_best = x_adv.clone(), grad.clone()
        counter, iters = 0, self.initial_step_size_iters
        for i in range(self.n_iter):
            ops.apgd_l2_step_(x_adv, x_adv_old, grad, x, step_size, 0.75 if i > 0 else 1.0, self.max_bound)
            loss, grad = _loss_and_grad(net, x_adv, labels, self.criterion)
            loss_steps[i] = loss
            better = (loss > best_loss).view(-1, 1, 1, 1)
            best_loss = torch.maximum(best_loss, loss)
            x_best = torch.where(better, x_adv, x_best)
            grad_best = torch.where(better, grad, grad_best)
            counter += 1
            if counter == iters:
                prev = loss_steps[i - (counter - 1): i + 1]                      # untargeted.py:68-85, per image
                incr = (prev[1:] > prev[:-1]).sum(dim=0)
                not_increasing = incr < counter * self.rho
                reduce = not_increasing | ((prev_best >= best_loss) & ~reduced_last)
                reduced_last = reduce
                prev_best = best_loss.clone()
                step_size = torch.where(reduce, step_size * 0.5, step_size)
                r4 = reduce.view(-1, 1, 1, 1)
                x_adv = torch.where(r4, x_best, x_adv).contiguous()              # restart from the best iterate
                grad = torch.where(r4, grad_best, grad).contiguous()
                counter = 0
                iters = max(iters - self.step_size_decr, self.min_step_size_iters)
        with torch.no_grad():
            success = net(x_adv).argmax(dim=1) != labels
        bound = (x_adv - x).flatten(1).norm(dim=1)
        return success, bound, x_adv


class FGSML2:
    """Batched FGSM with an L2-normalised sign step, the algorithm of the reference's `FGSM` (src/attacks/untargeted.py:708-750): images the
    network already misclassifies are returned unchanged (success, bound 0); the others move by l2_bound along sign(grad CE) / ||sign||."""

    def __init__(self, l2_bound: float):
        self.l2_bound = float(l2_bound)

    def __call__(self, images: torch.Tensor, labels: torch.Tensor, net):
        x = images.detach().to(torch.float32).contiguous()
        if hasattr(net, "loss_input_grad"):
            _, grad, pred = net.loss_input_grad(x, labels)
            wrong = pred.to(torch.int64) != labels
        else:
            xa = x.clone().requires_grad_(True)
            with torch.enable_grad():
                logits = net(xa)
                grad, = torch.autograd.grad(torch.nn.functional.cross_entropy(logits, labels), [xa])
            wrong = logits.argmax(dim=1) != labels
        x_adv = ops.fgsm_l2_step(x, grad.contiguous(), self.l2_bound)
        with torch.no_grad():
            flipped = net(x_adv).argmax(dim=1) != labels
        success = wrong | flipped
        bound = torch.where(wrong, torch.zeros_like(flipped, dtype=torch.float32), torch.full_like(flipped, self.l2_bound, dtype=torch.float32))
        x_adv = torch.where(wrong.view(-1, 1, 1, 1), x, x_adv)
        return success, bound, x_adv
