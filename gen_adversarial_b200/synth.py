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
        elif leaf == "mask":                            # MaskedConv2d buffer: fixed by the module's (mirror, zero_diag), architecture.py:17-28
            t = NvaeSpec.nf_mask(shape, mirror=".cell2." in key, zero_diag=key.endswith("layers.0.mask"))
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


def synthetic_stylegan_inputs(batch: int, res: int, n_codes: int, seed: int = 42):
    """x in [0,1] (B,3,res,res) + the two explicit N(0,1) draws of a StyleGAN defense call in the reference's order
    (SURVEY 8c): input noise (B,3,res,res), style noise (n_codes,B,512)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    x = torch.rand((batch, 3, res, res), generator=g, dtype=torch.float32)
    noises = [torch.randn((batch, 3, res, res), generator=g, dtype=torch.float32),
              torch.randn((n_codes, batch, 512), generator=g, dtype=torch.float32)]
    return x, noises


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

    def styled(prefix, cin, cout, k, up, gain=1.0):
        sd[f"{prefix}.conv.weight"] = _randn((1, cout, cin, k, k), g, gain)
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
        styled(prefix, cin, 3, 1, False, gain=0.1)     # the RGB skip sums log2(size)-1 such outputs: keeps the image inside ~[-1, 1]

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


# ----------------------------------------------------------------------------------------------- IR-SE50 encoders (E4E / Style-Transformer)
IRSE50_BLOCKS = [(64, 64, 3), (64, 128, 4), (128, 256, 14), (256, 512, 3)]          # encoding/helpers.py:30-37


def _bn_entries(sd, prefix, c, g):
    sd[f"{prefix}.weight"] = _rand((c,), g, 0.7, 1.3)
    sd[f"{prefix}.bias"] = _randn((c,), g, 0.1)
    sd[f"{prefix}.running_mean"] = _randn((c,), g, 0.1)
    sd[f"{prefix}.running_var"] = _rand((c,), g, 0.7, 1.3)
    sd[f"{prefix}.num_batches_tracked"] = torch.tensor(100, dtype=torch.long)


def _irse50_backbone_entries(sd, g, in_channels: int = 3):
    """`input_layer` + 24 `bottleneck_IR_SE` units (encoding/encoder.py:72-83, helpers.py:98-120).  Residual branches are
    scaled down (gain 0.5) so the un-normalised residual stream stays O(1) through 24 units, like a trained network's."""
    sd["input_layer.0.weight"] = _randn((64, in_channels, 3, 3), g, math.sqrt(2.0 / (in_channels * 9)))
    _bn_entries(sd, "input_layer.1", 64, g)
    sd["input_layer.2.weight"] = _rand((64,), g, 0.1, 0.4)
    i = 0
    for cin0, depth, n in IRSE50_BLOCKS:
        for j in range(n):
            cin = cin0 if j == 0 else depth
            p = f"body.{i}"
            if cin != depth:
                sd[f"{p}.shortcut_layer.0.weight"] = _randn((depth, cin, 1, 1), g, math.sqrt(1.0 / cin))
                _bn_entries(sd, f"{p}.shortcut_layer.1", depth, g)
            _bn_entries(sd, f"{p}.res_layer.0", cin, g)
            sd[f"{p}.res_layer.1.weight"] = _randn((depth, cin, 3, 3), g, math.sqrt(2.0 / (cin * 9)))
            sd[f"{p}.res_layer.2.weight"] = _rand((depth,), g, 0.1, 0.4)
            sd[f"{p}.res_layer.3.weight"] = _randn((depth, depth, 3, 3), g, 0.5 * math.sqrt(1.0 / (depth * 9)))
            _bn_entries(sd, f"{p}.res_layer.4", depth, g)
            sd[f"{p}.res_layer.5.fc1.weight"] = _randn((depth // 16, depth, 1, 1), g, 1.5 / math.sqrt(depth))
            sd[f"{p}.res_layer.5.fc2.weight"] = _randn((depth, depth // 16, 1, 1), g, 1.5 / math.sqrt(depth // 16))
            i += 1
    for name, cin in (("latlayer1", 256), ("latlayer2", 128)):
        sd[f"{name}.weight"] = _randn((512, cin, 1, 1), g, math.sqrt(1.0 / cin))
        sd[f"{name}.bias"] = _randn((512,), g, 0.1)


def make_e4e_encoder_state_dict(stylegan_size: int = 1024, seed: int = 4) -> "OrderedDict[str, torch.Tensor]":
    """State dict of `Encoder4Editing(50, 'ir_se', opts)` (StyleGan_E4E/encoding/encoder.py:57-108)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    _irse50_backbone_entries(sd, g)
    n_styles = 2 * int(math.log2(stylegan_size)) - 2
    for i in range(n_styles):
        spatial = 16 if i < 3 else (32 if i < 7 else 64)
        for j in range(int(math.log2(spatial))):
            sd[f"styles.{i}.convs.{2 * j}.weight"] = _randn((512, 512, 3, 3), g, math.sqrt(2.0 / (512 * 9)))
            sd[f"styles.{i}.convs.{2 * j}.bias"] = _randn((512,), g, 0.1)
        sd[f"styles.{i}.linear.weight"] = _randn((512, 512), g, 0.5 if i else 1.0)
        sd[f"styles.{i}.linear.bias"] = _randn((512,), g, 0.1)
    return sd


def make_trans_encoder_state_dict(seed: int = 5) -> "OrderedDict[str, torch.Tensor]":
    """State dict of `GradualStyleEncoder(50, 'ir_se', opts)` (StyleGan_Trans/models/encoders/style_transformer_encoders.py:10-40)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    _irse50_backbone_entries(sd, g)
    d, ff = 512, 1024
    for n in ("coarse", "medium", "fine"):
        p = f"transformerlayer_{n}"
        for att in ("self_attn", "multihead_attn"):
            sd[f"{p}.{att}.in_proj_weight"] = _randn((3 * d, d), g, math.sqrt(1.0 / d))
            sd[f"{p}.{att}.in_proj_bias"] = _randn((3 * d,), g, 0.1)
            sd[f"{p}.{att}.out_proj.weight"] = _randn((d, d), g, math.sqrt(1.0 / d))
            sd[f"{p}.{att}.out_proj.bias"] = _randn((d,), g, 0.1)
        sd[f"{p}.linear1.weight"] = _randn((ff, d), g, math.sqrt(2.0 / d))
        sd[f"{p}.linear1.bias"] = _randn((ff,), g, 0.1)
        sd[f"{p}.linear2.weight"] = _randn((d, ff), g, math.sqrt(1.0 / ff))
        sd[f"{p}.linear2.bias"] = _randn((d,), g, 0.1)
        for k in ("1", "2", "3"):
            sd[f"{p}.norm{k}.weight"] = _rand((d,), g, 0.7, 1.3)
            sd[f"{p}.norm{k}.bias"] = _randn((d,), g, 0.1)
    sd["z"] = _randn((1, 16, d), g)
    return sd


def make_e4e_checkpoint(stylegan_size: int = 1024, seed: int = 4) -> dict:
    """Format read by `load_E4EStyleGan` / `pSp.load_weights` (loading_utils.py:38-49, psp.py:39-45,117-127)."""
    n_styles = 2 * int(math.log2(stylegan_size)) - 2
    g = torch.Generator(device="cpu").manual_seed(seed + 1000)
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for k, v in make_e4e_encoder_state_dict(stylegan_size, seed).items():
        sd["encoder." + k] = v
    for k, v in make_stylegan2_state_dict(stylegan_size, seed=seed + 1).items():
        sd["decoder." + k] = v
    return {"opts": {"stylegan_size": stylegan_size, "encoder_type": "Encoder4Editing", "start_from_latent_avg": True},
            "state_dict": sd, "latent_avg": _randn((n_styles, 512), g, 0.3)}


def make_trans_checkpoint(output_size: int = 512, seed: int = 5) -> dict:
    """Format read by `load_TranStyleGan` / `StyleTransformer.load_weights` (loading_utils.py:69-81, style_transformer.py:30-36):
    keys carry the `.module` prefix of the DataParallel-trained checkpoint."""
    g = torch.Generator(device="cpu").manual_seed(seed + 1000)
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for k, v in make_trans_encoder_state_dict(seed).items():
        sd["encoder.module." + k] = v
    for k, v in make_stylegan2_state_dict(output_size, seed=seed + 1).items():
        sd["decoder.module." + k] = v
    return {"opts": {"output_size": output_size, "input_nc": 3, "start_from_latent_avg": True, "learn_in_w": False, "device": "cpu"},
            "state_dict": sd, "latent_avg": _randn((16, 512), g, 0.3)}


# ----------------------------------------------------------------------------------------------- ResNet-50 / ResNeXt-50 classifiers
def make_resnet_state_dict(n_classes: int, groups: int = 1, width_per_group: int = 64, seed: int = 6, calib_hw: int = 64,
                           calibrate: bool = True) -> "OrderedDict[str, torch.Tensor]":
    """State dict of `ResNet` / `ResNext` (/root/reference/src/classifier/model.py:10-28,53-70: torchvision resnet50 /
    resnext50_32x4d bodies, keys prefixed `model.`, 4-layer head in `model.fc`)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()

    def bn(prefix, c):
        _bn_entries(sd, prefix, c, g)

    sd["model.conv1.weight"] = _randn((64, 3, 7, 7), g, math.sqrt(2.0 / (3 * 49)))
    bn("model.bn1", 64)
    inplanes = 64
    for li, (planes, n) in enumerate(zip((64, 128, 256, 512), (3, 4, 6, 3)), start=1):
        width = int(planes * (width_per_group / 64.0)) * groups
        for bi in range(n):
            p = f"model.layer{li}.{bi}"
            sd[f"{p}.conv1.weight"] = _randn((width, inplanes, 1, 1), g, math.sqrt(2.0 / inplanes))
            bn(f"{p}.bn1", width)
            sd[f"{p}.conv2.weight"] = _randn((width, width // groups, 3, 3), g, math.sqrt(2.0 / (width // groups * 9)))
            bn(f"{p}.bn2", width)
            sd[f"{p}.conv3.weight"] = _randn((planes * 4, width, 1, 1), g, 0.5 * math.sqrt(1.0 / width))
            bn(f"{p}.bn3", planes * 4)
            if bi == 0:
                sd[f"{p}.downsample.0.weight"] = _randn((planes * 4, inplanes, 1, 1), g, math.sqrt(1.0 / inplanes))
                bn(f"{p}.downsample.1", planes * 4)
            inplanes = planes * 4
    sd["model.fc.0.weight"] = _randn((2048, 2048), g, math.sqrt(2.0 / 2048))
    bn("model.fc.1", 2048)
    sd["model.fc.3.weight"] = _randn((n_classes, 2048), g, math.sqrt(1.0 / 2048))
    sd["model.fc.3.bias"] = _randn((n_classes,), g, 0.05)
    if calibrate:
        # smooth calibration images (bilinearly up-sampled 8x8 noise): closer to purified reconstructions than white noise
        low = torch.rand((16, 3, 8, 8), generator=g, dtype=torch.float32)
        xc = torch.nn.functional.interpolate(low, size=(calib_hw, calib_hw), mode="bilinear", align_corners=False)
        calibrate_resnet_bn(sd, groups, width_per_group, xc)
    return sd


def calibrate_resnet_bn(sd, groups: int, width_per_group: int, x_calib: torch.Tensor):
    """Weight synthesis only (not the product path): set every BatchNorm's running statistics to the batch statistics of a
    calibration batch, as a trained checkpoint would have them -- otherwise a random-init network predicts one class for
    every input and equal accuracy counters would prove nothing.  Functional torch ops on the state dict."""
    import torch.nn.functional as F

    def bn_cal(x, prefix):
        dims = (0, 2, 3) if x.dim() == 4 else (0,)
        m, v = x.mean(dim=dims), x.var(dim=dims, unbiased=False)
        sd[f"{prefix}.running_mean"], sd[f"{prefix}.running_var"] = m, v
        return F.batch_norm(x, m, v, sd[f"{prefix}.weight"], sd[f"{prefix}.bias"], False, 0.0, 1e-5)

    x = (x_calib - 0.5) / 0.5
    x = F.relu(bn_cal(F.conv2d(x, sd["model.conv1.weight"], stride=2, padding=3), "model.bn1"))
    x = F.max_pool2d(x, 3, 2, 1)
    for li, n in enumerate((3, 4, 6, 3), start=1):
        for bi in range(n):
            p = f"model.layer{li}.{bi}"
            stride = 2 if (li > 1 and bi == 0) else 1
            h = F.relu(bn_cal(F.conv2d(x, sd[f"{p}.conv1.weight"]), f"{p}.bn1"))
            h = F.relu(bn_cal(F.conv2d(h, sd[f"{p}.conv2.weight"], stride=stride, padding=1, groups=groups), f"{p}.bn2"))
            h = bn_cal(F.conv2d(h, sd[f"{p}.conv3.weight"]), f"{p}.bn3")
            idt = x
            if f"{p}.downsample.0.weight" in sd:
                idt = bn_cal(F.conv2d(x, sd[f"{p}.downsample.0.weight"], stride=stride), f"{p}.downsample.1")
            x = F.relu(h + idt)
    x = x.mean(dim=(2, 3))
    bn_cal(x @ sd["model.fc.0.weight"].t(), "model.fc.1")
    return sd


def make_resnet50_checkpoint(n_classes: int = 2, seed: int = 6) -> dict:
    """`load_ResNet50` format (loading_utils.py:10-17): gender classifier."""
    return {"state_dict": make_resnet_state_dict(n_classes, 1, 64, seed)}


def make_resnext50_checkpoint(n_classes: int = 4, seed: int = 7) -> dict:
    """`load_ResNext50` format (loading_utils.py:29-36): cars classifier (resnext50_32x4d)."""
    return {"state_dict": make_resnet_state_dict(n_classes, 32, 4, seed)}
