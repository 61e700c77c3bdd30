"""VGG11-bn + 4-layer head (the `ids` classifier) on the hand-written CUDA kernels.

Replaces `Vgg.forward` (/root/reference/src/classifier/model.py:31-50; body = torchvision `vgg11_bn`) for the
purification path: conv3x3+BN+ReLU (BN folded) x8, 2x2 max-pools, AdaptiveAvgPool2d(7) + flatten +
Linear(25088,25088,bias=False) + BatchNorm1d + ReLU + Linear(25088,n_classes).

Exact load-time rewrites:
  * eval BN2d folded into the preceding conv, BN1d into the first head linear;
  * AdaptiveAvgPool2d(7) + NCHW flatten folded into the first head linear: for a (h x w) feature map the pooled
    vector is a fixed linear map P (7 x h) (x) P (7 x w) of the map, so W_eff[o, (p,q,c)] = sum_ij W[o, c*49+i*7+j]
    P[i,p] P[j,q].  At 64x64 inputs (2x2 map) this turns the 25088x25088 GEMM into 25088x2048 (12x fewer FLOPs
    and weight bytes); roofline fractions in bench.py still use the unfolded FLOP count.
"""
from __future__ import annotations

from typing import Dict

import torch

from . import ops
from ._lib import ACT_NONE, ACT_RELU
from .fold import Folder, BN_EPS
from .synth import vgg11_feature_layout


def adaptive_pool_matrix(out: int, inp: int) -> torch.Tensor:
    """P[i, p] of torch's AdaptiveAvgPool (window [floor(i*inp/out), ceil((i+1)*inp/out)) )."""
    P = torch.zeros((out, inp), dtype=torch.float64)
    for i in range(out):
        s = (i * inp) // out
        e = -((-(i + 1) * inp) // out)
        P[i, s:e] = 1.0 / (e - s)
    return P


class Vgg11Engine:
    def __init__(self, state_dict: Dict[str, torch.Tensor], device, mode: str = "fp32", in_hw: int = 64,
                 _host_logic_test: bool = False):
        self.device = torch.device(device)
        if self.device.type != "cuda" and not _host_logic_test:   # (tests/emu_ops.py drives the host logic on CPU)
            raise RuntimeError("Vgg11Engine runs only on CUDA devices: there is no CPU fallback")
        self.mode = mode
        self.bf16 = mode == "bf16"
        self.adt = torch.bfloat16 if self.bf16 else torch.float32
        self._has_dgrad = False
        sd = {(k[len("model."):] if k.startswith("model.") else k): v for k, v in state_dict.items()}
        f = Folder(sd, self.device, want_tc=self.bf16)
        self.layers = []
        hw = in_hw
        for lay in vgg11_feature_layout():
            if lay[0] == "pool":
                self.layers.append(("pool", None))
                hw //= 2
                continue
            _, idx, cin, cout = lay
            w = f.f64(f"features.{idx}.weight")
            b = f.f64(f"features.{idx}.bias")
            a, sh = f.bn(f"features.{idx + 1}")
            self.layers.append(("conv", f.conv(w * a.view(-1, 1, 1, 1), b * a + sh, pad=1, post_act=ACT_RELU,
                                               name=f"vgg.features.{idx}")))
        if hw < 1:
            raise ValueError("input too small for VGG11")
        self.feat_hw = hw
        self.feat_c = 512
        # ---- head: fold avgpool(7) + flatten + BN1d into linear 0 (chunked, on the device, fp32)
        w0 = sd["classifier.0.weight"]
        d_out, d_in = w0.shape
        assert d_in == 512 * 49
        P = adaptive_pool_matrix(7, hw).to(torch.float32).to(self.device)
        a1 = (sd["classifier.1.weight"].double() / torch.sqrt(sd["classifier.1.running_var"].double() + BN_EPS))
        b1 = sd["classifier.1.bias"].double() - sd["classifier.1.running_mean"].double() * a1
        a1 = a1.to(torch.float32).to(self.device)
        k_eff = hw * hw * 512
        w_eff = torch.empty((d_out, k_eff), dtype=torch.float32, device=self.device)
        step = 1024
        for r0 in range(0, d_out, step):
            chunk = w0[r0:r0 + step].to(self.device, torch.float32).view(-1, 512, 7, 7)
            eff = torch.einsum("ocij,ip,jq->opqc", chunk, P, P)            # NHWC flatten order (p, q, c)
            w_eff[r0:r0 + step] = eff.reshape(eff.shape[0], -1) * a1[r0:r0 + step, None]
        self.fc0 = ops.ConvLayer(1, 1, 1, 0, k_eff, d_out, post_act=ACT_RELU, name="vgg.classifier.0")
        self.fc0.bias = b1.to(torch.float32).to(self.device)
        if self.bf16:
            self.fc0.w_tc = w_eff.to(torch.bfloat16).contiguous()
        else:
            self.fc0.w_simt = w_eff.t().contiguous()
        del w_eff
        w3 = sd["classifier.3.weight"].to(self.device, torch.float32)
        self.n_classes = w3.shape[0]
        self.fc1 = ops.ConvLayer(1, 1, 1, 0, d_out, self.n_classes, name="vgg.classifier.3")
        self.fc1.bias = sd["classifier.3.bias"].to(self.device, torch.float32).contiguous()
        self.fc1.w_simt = w3.t().contiguous()
        if self.bf16:
            self.fc1.w_tc = w3.to(torch.bfloat16).contiguous()

    def _conv(self, x, L, out_f32=False, mul=None, mul_mode=0):
        if self.bf16 and ops.conv2d_tc_supported(x, L):
            ob, of = ops.conv2d_tc(x, L, want_bf16=not out_f32, want_f32=out_f32, mul=mul, mul_mode=mul_mode)
            return of if out_f32 else ob
        return ops.conv2d_simt(x, L, torch.float32 if (out_f32 or not self.bf16) else torch.bfloat16, mul=mul, mul_mode=mul_mode)

    def forward(self, x_nhwc: torch.Tensor, tape=None) -> torch.Tensor:
        """x_nhwc: (N,H,W,3) already normalised with mean=std=0.5 (abstract_models.py:59-60). -> logits fp32 (N, classes)
        tape: None or a list receiving the activations the input-gradient needs (ReLU outputs)."""
        x = x_nhwc
        if tape is not None and not self._has_dgrad:
            self._build_dgrad()
        for i, (kind, L) in enumerate(self.layers):
            if kind == "pool":
                y = ops.maxpool2x2(x, self.adt)
                if tape is not None:
                    tape.append(("vgg_pool", x))
                x = y
            else:
                # convs feeding a max-pool write fp32: the arg-max (and the gradient routing of the backward pass) is then
                # decided on fp32 values -- bf16 feature maps tie in ~10% of the 2x2 windows and misroute the gradient
                before_pool = i + 1 < len(self.layers) and self.layers[i + 1][0] == "pool"
                x = self._conv(x, L, out_f32=before_pool)
                if tape is not None:
                    tape.append(("vgg_conv", L, x))
        n = x.shape[0]
        assert x.shape[1] == self.feat_hw and x.shape[2] == self.feat_hw, "input resolution differs from the one folded at load"
        feat_shape = x.shape
        x = x.reshape(n, 1, 1, -1)
        h = self._conv(x, self.fc0)
        logits = self._conv(h, self.fc1, out_f32=True)
        if tape is not None:
            tape.append(("vgg_head", h, feat_shape))
        return logits.reshape(n, self.n_classes)

    # ------------------------------------------------------------------ backward (input gradient only)
    def _build_dgrad(self):
        """transposed / flipped weights for the input-gradient pass (built on first use: attacks only)."""
        def dg_conv(L):
            w = L.w_simt.view(L.kh, L.kw, L.cin, L.cout)                       # fp32 [kh,kw,cin,cout]
            wt = w.flip(0, 1).permute(0, 1, 3, 2).contiguous()                 # taps flipped, [kh,kw,cout,cin]
            D = ops.ConvLayer(L.kh, L.kw, 1, L.kh - 1 - L.pad, L.cout, L.cin, name=L.name + ".dgrad")
            D.w_simt = wt.reshape(L.kh * L.kw * L.cout, L.cin).contiguous()
            if self.bf16 and L.cout % 8 == 0:
                D.w_tc = wt.permute(3, 0, 1, 2).reshape(L.cin, L.kh * L.kw * L.cout).to(torch.bfloat16).contiguous()
            return D

        self.dgrad = {id(L): dg_conv(L) for kind, L in self.layers if kind == "conv"}
        # head: y = x W^T  ->  g_x = g_y W
        def dg_fc(L):
            D = ops.ConvLayer(1, 1, 1, 0, L.cout, L.cin, name=L.name + ".dgrad")
            if self.bf16:
                D.w_tc = L.w_tc.t().contiguous()                                # [cin, cout]
                if L.w_simt is not None:
                    D.w_simt = L.w_simt.t().contiguous()
            else:
                D.w_simt = L.w_simt.t().contiguous()                            # [cout, cin]
            return D
        self.fc0_d, self.fc1_d = dg_fc(self.fc0), dg_fc(self.fc1)
        self._has_dgrad = True

    def backward(self, tape, g_logits: torch.Tensor) -> torch.Tensor:
        """g_logits (N, classes) fp32 -> d loss / d x_nhwc (N,H,W,3) in fp32.  `tape` holds only "vgg_*" records."""
        n = g_logits.shape[0]
        g = g_logits.contiguous().reshape(n, 1, 1, -1)
        pending_relu = None                      # ReLU output whose mask still has to be applied to g
        for rec in reversed(tape):
            kind = rec[0]
            if kind == "vgg_head":
                _, h, feat_shape = rec
                if self.bf16 and g.dtype != torch.bfloat16:
                    g = ops.cast(g, torch.bfloat16)
                g = self._conv(g, self.fc1_d, mul=h, mul_mode=1)                # times ReLU mask of h
                g = self._conv(g, self.fc0_d).reshape(feat_shape)
            elif kind == "vgg_pool":
                _, x_in = rec
                g = ops.maxpool2x2_bwd(x_in, g, True, self.adt)                  # + ReLU mask of the conv before the pool
                pending_relu = None
            elif kind == "vgg_conv":
                _, L, y = rec
                if pending_relu is not None:                                     # conv directly followed by another conv
                    raise RuntimeError("internal: unresolved ReLU mask")
                D = self.dgrad[id(L)]
                # the input of this conv is either the network input, a pool output, or the previous conv's ReLU output;
                # in the last case that ReLU's mask is applied here as the epilogue multiplier
                prev = self._prev_relu_output(tape, rec)
                is_first = L.cin == 3
                if prev is not None:
                    g = self._conv(g, D, mul=prev, mul_mode=1)
                else:
                    g = self._conv(g, D, out_f32=is_first)
            else:
                raise RuntimeError(f"unknown tape record {kind}")
        return g if g.dtype == torch.float32 else ops.cast(g, torch.float32)

    @staticmethod
    def _prev_relu_output(tape, rec):
        """the ReLU output feeding `rec`'s conv directly (None if its input is a pool output or the network input)."""
        i = next(k for k, r in enumerate(tape) if r is rec)
        if i == 0:
            return None
        p = tape[i - 1]
        return p[2] if p[0] == "vgg_conv" else None
