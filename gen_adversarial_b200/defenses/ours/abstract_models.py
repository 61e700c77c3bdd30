"""Abstract defense-module API -- same names, constructor arguments, attributes and error behaviour as
/root/reference/src/defenses/ours/abstract_models.py (BaseClassificationModel :13-62, MLVGMDefenseModel :65-193),
with the compute routed through libga_b200.so (no torch/kornia arithmetic on the hot path, no CPU fallback).
"""
from __future__ import annotations

import os
from abc import ABC, abstractmethod
from typing import Union

import torch
from torch import nn

from ... import ops


def default_mode() -> str:
    """compute mode of the CUDA path: "bf16" (tcgen05 tensor cores, 1e-2 tolerance) or "fp32" (exact SIMT path)."""
    m = os.environ.get("GA_B200_MODE", "bf16")
    if m not in ("bf16", "fp32"):
        raise ValueError(f"GA_B200_MODE must be bf16 or fp32, got {m}")
    return m


class BaseClassificationModel(ABC):

    def __init__(self, model_path: str, device: str, mean: tuple = None, std: tuple = None):
        """
        Model containing only a base (pre-trained) classifier, with no additional defenses.

        :param model_path: absolute path to the pretrained model (or an already loaded checkpoint dict).
        :param device: cuda device to load the model on
        :param mean: optional param for Normalization.
        :param std: optional param for Normalization.
        """
        super().__init__()

        if (mean is not None and std is None) or (mean is None and std is not None):
            raise ValueError("to apply Normalization, please specify both mean and std.")

        self.mean = torch.tensor(mean, device=device) if mean is not None else None
        self.std = torch.tensor(std, device=device) if std is not None else None
        self.preprocess = self.mean is not None
        if self.preprocess and not (all(m == mean[0] for m in mean) and all(s == std[0] for s in std)):
            raise NotImplementedError("the CUDA path supports a scalar mean/std (the reference only uses 0.5/0.5)")
        self._norm_scale = 1.0 / float(std[0]) if self.preprocess else 1.0
        self._norm_shift = -float(mean[0]) / float(std[0]) if self.preprocess else 0.0

        self.classifier = self.load_classifier(model_path, device)

    @abstractmethod
    def load_classifier(self, model_path: str, device: str):
        """custom method to load the pretrained classifier -> engine object with `.forward(x_nhwc)`."""
        pass

    def set_device(self, device: str):
        """the engines are bound to their device at load time; moving is only allowed to the same device."""
        if torch.device(device) != self.classifier.device:
            if torch.device(device).type == "cuda" and torch.device(device).index in (None, self.classifier.device.index):
                return
            raise RuntimeError(f"classifier was loaded on {self.classifier.device}; reload it to use {device}")

    def classify_nhwc(self, x_nhwc: torch.Tensor) -> torch.Tensor:
        """normalised NHWC activations -> logits (used by the fused defense path)."""
        return self.classifier.forward(x_nhwc)

    def __call__(self, batch: torch.Tensor) -> torch.Tensor:
        """
        :param batch: image tensor of shape (B C H W)
        :return un-normalized predictions of shape (B N_CLASSES)
        """
        if batch.requires_grad and torch.is_grad_enabled():
            from ...autograd import classifier_apply
            return classifier_apply(self, batch)
        x = ops.nchw_to_nhwc(batch.detach().to(torch.float32), self.classifier.adt, self._norm_scale, self._norm_shift)
        return self.classifier.forward(x)


class MLVGMDefenseModel(ABC):

    def __init__(self, classifier: BaseClassificationModel, autoencoder_path: str,
                 interpolation_alphas: tuple, alpha_attenuation: float = 1.0,
                 initial_noise_eps: float = 0.0, apply_gaussian_blur: bool = False, device: str = 'cpu',
                 mean: tuple = None, std: tuple = None):
        """
        Model composed of HL-Autoencoder + CNN (see the reference docstring, abstract_models.py:71-87).
        """
        super().__init__()

        self.eps = initial_noise_eps
        self.blur_input = apply_gaussian_blur

        self.device = device

        self.classifier = classifier
        self.classifier.set_device(device)

        if (mean is not None and std is None) or (mean is None and std is not None):
            raise ValueError("to apply Normalization/Denormalization, please specify both mean and std.")

        self.mean = torch.tensor(mean, device=device) if mean is not None else None
        self.std = torch.tensor(std, device=device) if std is not None else None
        self.preprocess = self.mean is not None
        self.postprocess = self.mean is not None

        self.interpolation_alphas = [a * alpha_attenuation for a in interpolation_alphas]
        self.autoencoder = self.load_autoencoder(autoencoder_path, device)

        # ---- CUDA-path state (not part of the reference API)
        self._alpha_key = None
        self._alpha_dev = None
        self._alpha_pinned = None
        self._explicit_noise = None
        self._taps_cache = {}
        self.use_cuda_graph = False     # enable_cuda_graph(): replay a captured graph of the whole call (no-grad calls, Philox noise)
        self._graphs = {}
        self.noise_seed = None          # None: a fresh seed is drawn from torch's CPU generator on every call
        self.sample_offset = 0          # global index of sample 0 (data-parallel shards keep results G-independent)
        self.streams = 1                # set_streams(k): no-grad calls split the batch over k CUDA streams
        self._side_streams = []
        self._stream_warm = set()

    @abstractmethod
    def load_autoencoder(self, model_path: str, device: str):
        pass

    @abstractmethod
    def purify(self, batch: torch.Tensor) -> torch.Tensor:
        pass

    # ------------------------------------------------------------------ CUDA-path helpers
    def set_explicit_noise(self, noises):
        """Parity/test hook: the next call consumes these N(0,1) tensors (reference draw order)
        instead of the in-kernel Philox stream.  Pass None to go back to Philox."""
        self._explicit_noise = None if noises is None else [t.to(self.device, torch.float32).contiguous() for t in noises]

    def enable_cuda_graph(self, on: bool = True):
        """Replay the whole `__call__` (and, through `PGDLinf`, the attack iteration) as a CUDA graph: one graph per (batch shape, eps,
        blur, sample_offset); alphas and noise stay run-time inputs (device memory / seed salt).  Outputs are static buffers that the
        next call overwrites.  Calls that need gradients or explicit noise run eagerly."""
        self.use_cuda_graph = bool(on)
        if not on:
            self._graphs.clear()
        return self

    def set_streams(self, k: int = 2):
        """Run no-grad calls as k part-batches on k CUDA streams: the HBM-bound kernels of one part (SE / residual, channel sums, latent
        mixing) fill the SMs' idle memory pipes while the tensor-core / FMA-bound kernels of another part run, and kernel tails overlap.
        Results do not change: every kernel keys its Philox stream on the GLOBAL sample index (`sample_offset` + index inside the part)
        and no kernel reduces across images.  Measured on B200, NVAE ids batch 512: +5% img/s at k = 2."""
        self.streams = max(1, int(k))
        self._graphs.clear()
        return self

    def _run_parts(self, tag: str, fn, batch: torch.Tensor):
        """fn(part, lo, hi) for `streams` contiguous parts [lo, hi) of the batch, each on its own side stream, joined back into the caller's
        stream -> list of fn's results, or None when the call should not be split (one stream, tiny batch, explicit noise)."""
        k, n = self.streams, batch.shape[0]
        if k <= 1 or n < 2 * k or not batch.is_cuda or self._explicit_noise is not None:
            return None
        bounds = [(n * i) // k for i in range(k + 1)]
        sizes = (tag,) + tuple(bounds[i + 1] - bounds[i] for i in range(k))
        cur = torch.cuda.current_stream()
        self._alphas_device()                                   # host -> device refresh on the caller's stream, before the fork
        seed_was, off_was = self.noise_seed, self.sample_offset
        if seed_was is None:
            self.noise_seed = self._next_seed()                 # one draw per call, as the unsplit call makes for the whole batch
        # the first call with these part sizes runs the parts back to back on the caller's stream: lazily built state (prepared weights,
        # dgrad layers, blur taps, the broadcast prior) is then created in stream order and only read afterwards
        warm = sizes in self._stream_warm
        self._stream_warm.add(sizes)
        while len(self._side_streams) < k:
            self._side_streams.append(torch.cuda.Stream(device=batch.device))
        outs = []
        try:
            for i in range(k):
                part = batch[bounds[i]:bounds[i + 1]]
                self.sample_offset = off_was + bounds[i]
                if not warm:
                    outs.append(fn(part, bounds[i], bounds[i + 1]))
                    continue
                s = self._side_streams[i]
                s.wait_stream(cur)
                with torch.cuda.stream(s):
                    outs.append(fn(part, bounds[i], bounds[i + 1]))
            if warm:
                for s in self._side_streams[:k]:
                    cur.wait_stream(s)
        finally:
            self.noise_seed, self.sample_offset = seed_was, off_was
        return outs

    def _forward_parts(self, batch: torch.Tensor):
        """`_forward_cuda` of the whole batch, or of `streams` contiguous parts on side streams."""
        outs = self._run_parts("fwd", lambda part, lo, hi: self._forward_cuda(part), batch)
        if outs is None:
            return self._forward_cuda(batch)
        return torch.cat([o[0] for o in outs], dim=0), torch.cat([o[1] for o in outs], dim=0)

    def _alphas_device(self) -> torch.Tensor:
        """`interpolation_alphas` is a plain list that callers reassign between calls (common_utils.py:88); the
        kernels read alphas from device memory, refreshed here only when the list changed."""
        key = tuple(float(a) for a in self.interpolation_alphas)
        if key != self._alpha_key:
            if self._alpha_pinned is None or self._alpha_pinned.numel() != len(key):
                # one (pinned, device) buffer pair PER LENGTH, never freed: captured graphs are keyed by the length and keep the device
                # pointer of the buffer that existed at capture time
                if not hasattr(self, "_alpha_bufs"):
                    self._alpha_bufs = {}
                if len(key) not in self._alpha_bufs:
                    self._alpha_bufs[len(key)] = (torch.empty(len(key), dtype=torch.float32).pin_memory(),
                                                  torch.empty(len(key), dtype=torch.float32, device=self.device))
                elif torch.cuda.is_available():
                    torch.cuda.current_stream().synchronize()
                self._alpha_pinned, self._alpha_dev = self._alpha_bufs[len(key)]
            elif torch.cuda.is_available():
                torch.cuda.current_stream().synchronize()      # the previous async copy must have read the pinned buffer
            self._alpha_pinned.copy_(torch.tensor(key, dtype=torch.float32))
            self._alpha_dev.copy_(self._alpha_pinned, non_blocking=True)
            self._alpha_key = key
        return self._alpha_dev

    def _next_seed(self) -> int:
        if self.noise_seed is not None:
            return int(self.noise_seed)
        return int(torch.empty((), dtype=torch.int64).random_().item())

    # ------------------------------------------------------------------ reference API
    def add_gaussian_noise(self, x: torch.Tensor) -> torch.Tensor:
        """abstract_models.py:129-143 (N(0,1) noise scaled to L2 norm eps per sample, clamp to [0,1])."""
        from ...autograd import preprocess_apply
        noise = self._explicit_noise[0] if self._explicit_noise is not None else None
        return preprocess_apply(x, noise, float(self.eps), False, self._next_seed(), self.sample_offset)

    def apply_gaussian_blur(self, x: torch.Tensor) -> torch.Tensor:
        """abstract_models.py:145-159."""
        if not self.blur_input:
            return x
        from ...autograd import preprocess_apply
        # NOTE: the fused kernel also clamps to [0,1]; a blur of values in [0,1] stays in [0,1]
        return preprocess_apply(x, None, 0.0, True, 0, 0, self._taps_cache)

    def __call__(self, batch: torch.Tensor, preds_only: bool = True) \
            -> Union[torch.Tensor, [torch.Tensor, torch.Tensor]]:
        """
        :param batch: image tensor of shape (B C H W)
        :return if preds only: un-normalized predictions (B, N_CLASSES) computed on the purified images,
                else: (predictions, purified images (B, C, H, W))
        """
        if batch.shape[0] == 0:
            # empty batch: the reference's torch modules return empty tensors; the kernels are never launched on zero-sized buffers
            preds = torch.zeros((0, self.classifier.classifier.n_classes), device=batch.device, dtype=torch.float32)
            purified = torch.zeros((0,) + tuple(batch.shape[1:]), device=batch.device, dtype=torch.float32)
        elif batch.requires_grad and torch.is_grad_enabled():
            from ...autograd import defense_apply
            preds, purified = defense_apply(self, batch)
        elif self.use_cuda_graph and self._explicit_noise is None and batch.is_cuda:
            from ...graphs import GraphedForward
            key = GraphedForward.make_key(self, batch)
            g = self._graphs.get(key)
            if g is None:
                g = self._graphs[key] = GraphedForward(self, batch)
            preds, purified = g(self, batch)
        else:
            preds, purified = self._forward_parts(batch.detach())
        if preds_only:
            return preds
        return preds, purified

    @abstractmethod
    def _forward_cuda(self, batch: torch.Tensor, tape=None):
        pass

    def loss_input_grad(self, batch: torch.Tensor, labels: torch.Tensor, counter: torch.Tensor = None):
        """Fused attack primitive: cross-entropy loss of the defended classifier and its gradient w.r.t. `batch`
        (what `torch.autograd.grad(F.cross_entropy(net(x), y), [x])` returns, untargeted.py:146,201) without going through
        torch's autograd engine.  -> (loss[n], grad (n,C,H,W) fp32, pred[n] int32).  `counter`: optional uint64 device
        scalar accumulating argmax == label."""
        n = batch.shape[0]
        outs = self._run_parts("grad", lambda part, lo, hi: self._loss_input_grad_one(part, labels[lo:hi], counter, (hi - lo) / n), batch.detach())
        if outs is None:
            return self._loss_input_grad_one(batch, labels, counter, 1.0)
        return tuple(torch.cat([o[j] for o in outs], dim=0) for j in range(3))

    def _loss_input_grad_one(self, batch, labels, counter, weight: float):
        """`weight` = this part's share of the batch: the loss kernel differentiates the mean over ITS rows, the caller wants the mean over
        the whole batch (exact for the power-of-two shares of an even split)"""
        from ...autograd import Tape
        tape = Tape()
        preds, _ = self._forward_cuda(batch.detach(), tape=tape)
        loss, dlogits, pred = ops.softmax_xent(preds, labels, want_grad=True, counter=counter)
        if weight != 1.0:
            dlogits.mul_(weight)
        g_cls = self.classifier.classifier.backward(tape.vgg, dlogits)
        g_x = self.autoencoder.backward(tape.nvae, None, g_cls)
        gx = ops.preprocess_bwd(g_x, tape.pre, bool(self.blur_input), normalize=True, taps_cache=self._taps_cache)
        return loss, gx, pred
