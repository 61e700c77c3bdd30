"""CUDA-graph replay of the purification call and of the PGD inner loop.

The NVAE forward is ~550 kernel launches and one PGD iteration (forward + input-gradient + step) ~1500; at attack batch sizes the
kernels are short and the Python launch rate bounds the step.  Both are captured once per (shape, configuration) and replayed.
Everything the captured kernels read that changes between calls lives in device memory: the input batch (static buffer), the alphas
(`MLVGMDefenseModel._alphas_device`), and the Philox seed salt (`ga_seed_salt_*`): the by-value seeds are frozen into the graph, the
first node of every graph bumps the salt, so each replay draws fresh noise, as the reference does on every call
(src/defenses/ours/abstract_models.py:132, NVAE/modules/distributions.py:43).
"""
from __future__ import annotations

import torch

from . import _lib, ops

_SALT = {}
REPLAYED_LAUNCHES = [0]          # kernel nodes executed through graph replays (ops.launch_count() only sees eager launches)


def enable_seed_salt(device) -> torch.Tensor:
    """register the per-process salt buffer (one process drives one GPU)"""
    device = torch.device(device)
    if device not in _SALT:
        t = torch.zeros(1, dtype=torch.int64, device=device)
        _lib.check(_lib.lib().ga_seed_salt_set(t.data_ptr()), "seed_salt_set")
        _SALT.clear()
        _SALT[device] = t
    return _SALT[device]


def bump_seed_salt():
    _lib.check(_lib.lib().ga_seed_salt_bump(ops.stream()), "seed_salt_bump")


def _capture(fn):
    """warm-up on a side stream (lazy weight preparation, allocator), then capture `fn` -> (graph, outputs)"""
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    n0 = ops.launch_count()
    with torch.cuda.graph(g):
        out = fn()
    return g, out, ops.launch_count() - n0


class GraphedForward:
    """`model._forward_cuda(batch)` for one batch shape: copy-in, replay, read the static outputs (overwritten by the next replay)"""

    def __init__(self, model, example: torch.Tensor):
        enable_seed_salt(example.device)
        model._alphas_device()                                       # host -> device refresh happens outside the graph
        self.key = self.make_key(model, example)
        self.x = example.detach().to(torch.float32).clone()

        def body():
            bump_seed_salt()
            with torch.no_grad():
                return model._forward_parts(self.x)

        self.graph, (self.preds, self.purified), self.nodes = _capture(body)

    @staticmethod
    def make_key(model, batch):
        # everything a captured launch takes BY VALUE is part of the key (eps, blur, sample offset, the seed, the generator chunking)
        return (tuple(batch.shape), float(model.eps), bool(model.blur_input), int(model.sample_offset), len(model.interpolation_alphas),
                model.noise_seed, getattr(model, "max_chunk", None), int(getattr(model, "streams", 1)))

    def __call__(self, model, batch: torch.Tensor):
        model._alphas_device()
        self.x.copy_(batch, non_blocking=True)
        self.graph.replay()
        REPLAYED_LAUNCHES[0] += self.nodes
        return self.preds, self.purified


class GraphedPGD:
    """one PGD-Linf iteration (loss + input gradient through purifier and classifier + fused step) as a graph, replayed `steps` times"""

    def __init__(self, net, images: torch.Tensor, labels: torch.Tensor, step: float, eps: float):
        enable_seed_salt(images.device)
        net._alphas_device()
        self.key = (tuple(images.shape), float(step), float(eps), GraphedForward.make_key(net, images))
        self.x = images.detach().to(torch.float32).clone()
        self.x_adv = self.x.clone()
        self.labels = labels.to(torch.int64).clone()          # the loss kernel reads int64 labels

        def body():
            bump_seed_salt()
            _, grad, _ = net.loss_input_grad(self.x_adv, self.labels)
            ops.pgd_linf_step_(self.x_adv, grad, self.x, step, eps)

        x0 = self.x_adv.clone()
        self.graph, _, self.nodes = _capture(body)
        self.x_adv.copy_(x0)                                          # warm-up and capture stepped the buffer

    def run(self, net, images, labels, steps: int, x_start=None):
        net._alphas_device()
        self.x.copy_(images, non_blocking=True)
        self.x_adv.copy_(images if x_start is None else x_start, non_blocking=True)
        self.labels.copy_(labels, non_blocking=True)
        for _ in range(steps):
            self.graph.replay()
        REPLAYED_LAUNCHES[0] += self.nodes * steps
        return self.x_adv
