"""Synthetic (random-init) checkpoints in the reference's own on-disk formats.

No weights ship with the reference (HF links only, /root/reference/README.md:31-68) and there is
no network, so every test / benchmark uses seeded random-init weights of the named architectures
(SURVEY.md section 8c "Weights", 8d "Synthetic inputs").  The dictionaries written here are
ingested unchanged by the reference loaders (`src/defenses/loading_utils.py:10-81`), which is
checked in tests/test_oracle_vs_reference.py.

Random-init hazard: with PyTorch default init the NVAE output is ~constant, so BN
running statistics / affine parameters, SE and conv biases are all randomised.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict, Tuple

import torch

from .nvae_spec import NvaeSpec, NVAE_C32_CONFIG, NVAE_C32_RESOLUTION

VGG11_CFG = [64, "M", 128, "M", 256, 256, "M", 512, 512, "M", 512, 512, "M"]


def _randn(shape, g, std=1.0, device="cpu"):
    return torch.randn(shape, generator=g, device=device, dtype=torch.float32) * std


def _rand(shape, g, lo, hi, device="cpu"):
    return torch.rand(shape, generator=g, device=device, dtype=torch.float32) * (hi - lo) + lo


# Weight-norm gains per module family, chosen so that activations stay O(1) through the 24-level
# decoder (a decoder whose prior log-sigma saturates the soft clamp at +5 amplifies rounding noise
# by 1e4 and makes every tolerance meaningless -- a trained checkpoint does not behave like that).
_WN_GAINS = (
    ("preprocessing_block.init_conv", (1.6, 2.4)),
    ("dec_sampler", (0.25, 0.45)),
    ("enc_sampler", (0.5, 0.8)),
    ("decoder_combiners", (0.75, 0.95)),
    ("encoder_combiners", (0.5, 0.8)),
    ("skip_connection", (0.9, 1.1)),
    ("to_logits", (3.0, 4.0)),
)


def _wn_gain_range(key: str):
    for sub, rng in _WN_GAINS:
        if sub in key:
            return rng
    return (0.8, 1.4)


def make_nvae_state_dict(cfg: dict = None, resolution: Tuple[int, int, int] = None, seed: int = 0) -> "OrderedDict[str, torch.Tensor]":
    cfg = NVAE_C32_CONFIG if cfg is None else cfg
    resolution = NVAE_C32_RESOLUTION if resolution is None else resolution
    spec = NvaeSpec(cfg, resolution)
    g = torch.Generator(device="cpu").manual_seed(seed)
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for key, shape in spec.state_dict_shapes().items():
        leaf = key.rsplit(".", 1)[-1]
        if key == "const_prior":
            t = _rand(shape, g, 0.0, 1.0)
        elif leaf == "num_batches_tracked":
            t = torch.tensor(100, dtype=torch.long)
        elif leaf == "running_mean":
            t = _randn(shape, g, 0.1)
        elif leaf == "running_var":
            t = _rand(shape, g, 0.5, 1.5)
        elif leaf == "original0":                       # weight-norm magnitude g
            lo, hi = _wn_gain_range(key)
            t = _rand(shape, g, lo, hi)
        elif leaf == "original1":                       # weight-norm direction v
            t = _randn(shape, g, 1.0)
        elif leaf == "weight" and len(shape) == 1:      # BN affine scale
            t = _rand(shape, g, 0.5, 1.5)
        elif leaf == "bias":
            t = _randn(shape, g, 0.1)
        elif leaf == "weight" and len(shape) == 4:      # plain conv (decoder cells)
            fan_in = shape[1] * shape[2] * shape[3]
            t = _randn(shape, g, 1.3 / math.sqrt(fan_in))
        elif leaf == "weight" and len(shape) == 2:      # SE linears
            t = _randn(shape, g, 1.5 / math.sqrt(shape[1]))
        else:
            raise KeyError(f"unhandled key {key} {shape}")
        sd[key] = t
    return sd


def make_nvae_checkpoint(cfg: dict = None, resolution=None, seed: int = 0, temperature: float = 0.6) -> dict:
    """Format of loading_utils.py:57-64."""
    cfg = NVAE_C32_CONFIG if cfg is None else cfg
    resolution = NVAE_C32_RESOLUTION if resolution is None else resolution
    return {"configuration": {"autoencoder": dict(cfg), "resolution": tuple(resolution)},
            f"state_dict_temp={temperature}": make_nvae_state_dict(cfg, resolution, seed)}


def vgg11_feature_layout():
    """-> list of ('conv', idx, cin, cout) / ('pool',) following torchvision vgg11_bn `features`."""
    layers, idx, cin = [], 0, 3
    for v in VGG11_CFG:
        if v == "M":
            layers.append(("pool", idx)); idx += 1
        else:
            layers.append(("conv", idx, cin, v)); idx += 3          # conv, bn, relu
            cin = v
    return layers


def calibrate_vgg11_bn(sd, x_calib: torch.Tensor):
    """Set every BN's running statistics to the batch statistics of a calibration batch, the way a
    trained checkpoint would have them.  Without this a random-init VGG predicts one class for every
    input and the accuracy counters of the parity tests would be trivially equal.  (Weight synthesis
    only -- plain torch ops on whatever device the tensors live on; not part of the product path.)"""
    import torch.nn.functional as F
    x = (x_calib.to(sd["model.features.0.weight"].device) - 0.5) / 0.5
    for lay in vgg11_feature_layout():
        if lay[0] == "pool":
            x = F.max_pool2d(x, 2)
            continue
        _, idx, cin, cout = lay
        y = F.conv2d(x, sd[f"model.features.{idx}.weight"], sd[f"model.features.{idx}.bias"], padding=1)
        m, v = y.mean(dim=(0, 2, 3)), y.var(dim=(0, 2, 3), unbiased=False)
        sd[f"model.features.{idx + 1}.running_mean"] = m
        sd[f"model.features.{idx + 1}.running_var"] = v
        x = F.relu(F.batch_norm(y, m, v, sd[f"model.features.{idx + 1}.weight"],
                                sd[f"model.features.{idx + 1}.bias"], False, 0.0, 1e-5))
    x = F.adaptive_avg_pool2d(x, 7).flatten(1)
    y = x @ sd["model.classifier.0.weight"].t()
    sd["model.classifier.1.running_mean"] = y.mean(0)
    sd["model.classifier.1.running_var"] = y.var(0, unbiased=False)
    return sd


def make_vgg11_state_dict(n_classes: int = 100, seed: int = 1, device: str = "cpu",
                          head_dim: int = 25088, calibrate: bool = True) -> "OrderedDict[str, torch.Tensor]":
    """State dict of /root/reference/src/classifier/model.py:31-50 (`Vgg`, keys prefixed `model.`)."""
    g = torch.Generator(device=device).manual_seed(seed)
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()

    def bn(prefix, c):
        sd[f"{prefix}.weight"] = _rand((c,), g, 0.5, 1.5, device)
        sd[f"{prefix}.bias"] = _randn((c,), g, 0.1, device)
        sd[f"{prefix}.running_mean"] = _randn((c,), g, 0.1, device)
        sd[f"{prefix}.running_var"] = _rand((c,), g, 0.5, 1.5, device)
        sd[f"{prefix}.num_batches_tracked"] = torch.tensor(100, dtype=torch.long, device=device)

    for lay in vgg11_feature_layout():
        if lay[0] != "conv":
            continue
        _, idx, cin, cout = lay
        sd[f"model.features.{idx}.weight"] = _randn((cout, cin, 3, 3), g, math.sqrt(2.0 / (cin * 9)), device)
        sd[f"model.features.{idx}.bias"] = _randn((cout,), g, 0.05, device)
        bn(f"model.features.{idx + 1}", cout)
    sd["model.classifier.0.weight"] = _randn((head_dim, head_dim), g, math.sqrt(2.0 / head_dim), device)
    bn("model.classifier.1", head_dim)
    sd["model.classifier.3.weight"] = _randn((n_classes, head_dim), g, math.sqrt(1.0 / head_dim), device)
    sd["model.classifier.3.bias"] = _randn((n_classes,), g, 0.05, device)
    if calibrate:
        xc = torch.rand((32, 3, 64, 64), generator=g, device=device, dtype=torch.float32)
        calibrate_vgg11_bn(sd, xc)
    return sd


def make_vgg11_checkpoint(n_classes: int = 100, seed: int = 1, device: str = "cpu") -> dict:
    """Format of loading_utils.py:20-26."""
    return {"state_dict": make_vgg11_state_dict(n_classes, seed, device)}


def synthetic_batch(batch: int, resolution=NVAE_C32_RESOLUTION, n_classes: int = 100, seed: int = 42):
    """x = rand(B,3,H,W) in [0,1] with seed 42, labels randint(n_classes)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    c, h, w = resolution
    x = torch.rand((batch, c, h, w), generator=g, dtype=torch.float32)
    y = torch.randint(0, n_classes, (batch,), generator=g)
    return x, y


def synthetic_noise(spec: NvaeSpec, batch: int, seed: int = 7):
    """Explicit N(0,1) tensors in the reference's draw order (SURVEY 8c "RNG order")."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    return [torch.randn(s, generator=g, dtype=torch.float32) for s in spec.noise_shapes(batch)]


# ----------------------------------------------------------------------------------------------- StyleGAN2 generator
STYLEGAN_CHANNELS = lambda cm: {4: 512, 8: 512, 16: 512, 32: 512, 64: 256 * cm, 128: 128 * cm, 256: 64 * cm, 512: 32 * cm, 1024: 16 * cm}


def make_stylegan2_state_dict(size: int = 32, style_dim: int = 512, n_mlp: int = 8, channel_multiplier: int = 2, seed: int = 2,
                              lr_mlp: float = 0.01) -> "OrderedDict[str, torch.Tensor]":
    """State dict of the reference `Generator` (/root/reference/src/mlvgms_autoencoders/StyleGan_E4E/stylegan2/
    generator.py:294-385): mapping MLP, constant input, StyledConv / ToRGB stacks, fixed noise buffers, blur kernels."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    ch = STYLEGAN_CHANNELS(channel_multiplier)
    log_size = int(math.log2(size))
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for i in range(1, n_mlp + 1):
        sd[f"style.{i}.weight"] = _randn((style_dim, style_dim), g) / lr_mlp
        sd[f"style.{i}.bias"] = _randn((style_dim,), g, 0.1 / lr_mlp)
    sd["input.input"] = _randn((1, ch[4], 4, 4), g)
    blur = torch.tensor([1.0, 3.0, 3.0, 1.0])
    blur2 = blur[None, :] * blur[:, None]
    blur2 = blur2 / blur2.sum()

    def styled(prefix, cin, cout, k, up):
        sd[f"{prefix}.conv.weight"] = _randn((1, cout, cin, k, k), g)
        if up:
            sd[f"{prefix}.conv.blur.kernel"] = blur2 * 4.0
        sd[f"{prefix}.conv.modulation.weight"] = _randn((cin, style_dim), g)
        sd[f"{prefix}.conv.modulation.bias"] = torch.ones(cin) + _randn((cin,), g, 0.1)

    def conv_block(prefix, cin, cout, up):
        styled(prefix, cin, cout, 3, up)
        sd[f"{prefix}.noise.weight"] = _randn((1,), g, 0.3)
        sd[f"{prefix}.activate.bias"] = _randn((cout,), g, 0.1)

    def to_rgb(prefix, cin, up):
        sd[f"{prefix}.bias"] = _randn((1, 3, 1, 1), g, 0.1)
        if up:
            sd[f"{prefix}.upsample.kernel"] = blur2 * 4.0
        styled(prefix, cin, 3, 1, False)

    conv_block("conv1", ch[4], ch[4], False)
    to_rgb("to_rgb1", ch[4], False)
    cin = ch[4]
    for j, i in enumerate(range(3, log_size + 1)):
        cout = ch[2 ** i]
        conv_block(f"convs.{2 * j}", cin, cout, True)
        conv_block(f"convs.{2 * j + 1}", cout, cout, False)
        to_rgb(f"to_rgbs.{j}", cout, True)
        cin = cout
    num_layers = (log_size - 2) * 2 + 1
    for layer_idx in range(num_layers):
        res = (layer_idx + 5) // 2
        sd[f"noises.noise_{layer_idx}"] = _randn((1, 1, 2 ** res, 2 ** res), g)
    return sd
