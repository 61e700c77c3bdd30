"""Load-time weight preparation for the CUDA path (host side, done once per checkpoint).

Ingests the reference's own state_dict layout (weight-norm `parametrizations.weight.original0/1`, eval-mode
SyncBatchNorm running statistics; /root/reference/src/defenses/loading_utils.py:51-66) and produces folded,
re-laid-out device tensors.  All folds are algebraically exact and are done in fp64:

  * weight-norm          W = g * v / ||v||_2                      (architecture.py:75,89,122,125,193,213)
  * eval BN              y = a*x + b,  a = gamma/sqrt(var+eps), b = beta - mean*a
  * encoder cell         BN2 folded into conv1; BN1 stays as the (a1,b1)+SiLU pre-activation because the
                         reference zero-pads AFTER BN+SiLU (architecture.py:119-126)
  * decoder cell         BN0,BN1 folded into the 1x1 expand, BN2 into the depthwise 5x5, BN3 into the 1x1
                         project (architecture.py:164-173); the nearest x2 up-sampling of the up cells commutes
                         with BN0 + 1x1 conv, and the bilinear x2 of the skip commutes with its 1x1 conv, so both
                         1x1 convs run at the LOW resolution (4x fewer FLOPs, exact)
  * dead outputs         the encoder samplers' log-sigma half is never read by purify (models.py:206,249) and is
                         dropped from the weights
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch

from .ops import ConvLayer
from ._lib import PRE_NONE, PRE_ELU, PRE_SILU, PRE_AFFINE_SILU, ACT_NONE, ACT_SILU, ACT_ELU, ACT_RELU

BN_EPS = 1e-5


def round_up(x: int, m: int) -> int:
    return (x + m - 1) // m * m


class Folder:
    def __init__(self, sd: Dict[str, torch.Tensor], device, want_tc: bool):
        self.sd = sd
        self.device = device
        self.want_tc = want_tc

    def f64(self, key):
        return self.sd[key].detach().to("cpu", torch.float64)

    def wn(self, prefix) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
        g = self.f64(f"{prefix}.parametrizations.weight.original0")
        v = self.f64(f"{prefix}.parametrizations.weight.original1")
        w = g * v / v.flatten(1).norm(dim=1).view(-1, 1, 1, 1)
        b = self.f64(f"{prefix}.bias") if f"{prefix}.bias" in self.sd else None
        return w, b

    def bn(self, prefix) -> Tuple[torch.Tensor, torch.Tensor]:
        a = self.f64(f"{prefix}.weight") / torch.sqrt(self.f64(f"{prefix}.running_var") + BN_EPS)
        b = self.f64(f"{prefix}.bias") - self.f64(f"{prefix}.running_mean") * a
        return a, b

    def dev32(self, t: Optional[torch.Tensor]):
        return None if t is None else t.to(torch.float32).contiguous().to(self.device)

    def se(self, prefix):
        return tuple(self.dev32(self.f64(f"{prefix}.{k}")) for k in
                     ("linear_1.weight", "linear_1.bias", "linear_2.weight", "linear_2.bias"))

    def conv(self, w: torch.Tensor, b: Optional[torch.Tensor], stride=1, pad=0, pre_op=PRE_NONE, pre_affine=None,
             post_act=ACT_NONE, name="", w2: Optional[torch.Tensor] = None, simt: bool = True, up: int = 1,
             tf32: bool = False) -> ConvLayer:
        """w: [cout, cin, kh, kw] fp64 (already folded). w2: optional [cout, cin2] weights of a second 1x1 source."""
        cout, cin, kh, kw = w.shape
        L = ConvLayer(kh, kw, stride, pad, cin, cout, pre_op=pre_op, post_act=post_act, name=name, up=up)
        if simt:
            L.w_simt = self.dev32(w.permute(2, 3, 1, 0).reshape(kh * kw * cin, cout))
        if self.want_tc and stride in (1, 2) and up == 1 and cin % 8 == 0 and ((kh == 1 and pad == 0) or (kh == 3 and pad == 1)):
            wk = w.permute(0, 2, 3, 1).reshape(cout, kh * kw * cin)
            if w2 is not None:
                wk = torch.cat([wk, w2], dim=1)
                L.cin2 = w2.shape[1]
            L.w_tc = wk.to(torch.bfloat16).contiguous().to(self.device)
            if tf32 and w2 is None and cin % 4 == 0:
                w32 = wk.to(torch.float32).contiguous()
                bits = (w32.view(torch.int32) + 0x1000) & ~0x1FFF                # round to nearest TF32 (the MMA truncates: biased otherwise)
                L.w_tf32 = bits.view(torch.float32).contiguous().to(self.device)
        L.bias = self.dev32(b)
        if pre_affine is not None:
            L.pre_scale, L.pre_shift = self.dev32(pre_affine[0]), self.dev32(pre_affine[1])
        return L


def transpose_for_dgrad(f: Folder, L_w: torch.Tensor, stride: int, pad: int, name: str) -> ConvLayer:
    """dgrad of conv(w [cout,cin,kh,kw], stride, pad) as a convolution over grad_out:
    weights flipped in space and transposed in channels; stride-2 becomes input dilation (`up`)."""
    cout, cin, kh, kw = L_w.shape
    wt = L_w.flip(2, 3).permute(1, 0, 2, 3).contiguous()     # [cin, cout, kh, kw]
    return f.conv(wt, None, stride=1, pad=kh - 1 - pad, name=name + ".dgrad", up=stride)
