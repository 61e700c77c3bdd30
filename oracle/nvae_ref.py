"""TEST INFRASTRUCTURE ONLY -- CPU restatement (plain torch fp32/fp64 ops) of the reference's
NVAE purification path.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this; the product (gen_adversarial_b200/) never does.

Parity status: PINNED against the unmodified reference run in the build container
(tests/test_oracle_vs_reference.py imports /root/reference through oracle/ref_import.py and
compares on seeded weights/inputs/noise), and against the committed fixtures in tests/golden/
that were produced by the reference itself (oracle/make_golden.py).  The reference has no
tests or golden vectors of its own.

Every function cites the reference lines it restates (paths relative to /root/reference).
The restatement works directly on the reference's `state_dict` (weight-norm g/v pairs, BN
running statistics) and does NOT fold anything, so that it checks the folds done by the product.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F

from gen_adversarial_b200.nvae_spec import NvaeSpec, EncCell, DecCell

BN_EPS = 1e-5   # SyncBatchNorm(eps=1e-5), architecture.py:120-126,165-173


# ----------------------------------------------------------------------------- pre-processing
def gaussian_blur_ksize(h: int) -> int:
    """src/defenses/ours/abstract_models.py:150-156: k = int(2**(sqrt(h)//2) - 1)."""
    return int(2 ** (math.sqrt(h) // 2) - 1)


def gaussian_blur(x: torch.Tensor) -> torch.Tensor:
    """abstract_models.py:145-159 + kornia.filters.gaussian_blur2d(k, sigma=(1,1)) (reflect border,
    separable, kernel exp(-t^2/2)/sum, t = arange(k) - k//2)."""
    b, c, h, w = x.shape
    k = gaussian_blur_ksize(h)
    t = torch.arange(k, dtype=x.dtype) - k // 2
    g = torch.exp(-(t ** 2) / 2.0)
    g = g / g.sum()
    p = k // 2
    xp = F.pad(x, (p, p, p, p), mode="reflect")
    xp = F.conv2d(xp, g.view(1, 1, 1, k).repeat(c, 1, 1, 1), groups=c)
    xp = F.conv2d(xp, g.view(1, 1, k, 1).repeat(c, 1, 1, 1), groups=c)
    return xp


def add_gaussian_noise(x: torch.Tensor, noise: torch.Tensor, eps: float) -> torch.Tensor:
    """abstract_models.py:129-143 with the N(0,1) draw supplied explicitly."""
    norm = noise.reshape(noise.shape[0], -1).norm(dim=1).view(-1, 1, 1, 1)
    return (x + noise * (eps / norm)).clamp(0.0, 1.0)


def preprocess(x, noise, eps: float, blur: bool):
    """abstract_models.py:173-175 (blur then noise)."""
    if blur:
        x = gaussian_blur(x)
    return add_gaussian_noise(x, noise, eps)


# ----------------------------------------------------------------------------- NVAE building blocks
def soft_clamp5(x):
    """NVAE/modules/distributions.py:20-29."""
    return 5.0 * torch.tanh(x / 5.0)


class _SD:
    def __init__(self, sd: Dict[str, torch.Tensor], dtype):
        self.sd = {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in sd.items()}

    def wn(self, prefix):
        """torch weight_norm parametrisation: W = g * v / ||v|| (norm over dims 1,2,3), architecture.py:75,89,122."""
        g = self.sd[f"{prefix}.parametrizations.weight.original0"]
        v = self.sd[f"{prefix}.parametrizations.weight.original1"]
        n = v.flatten(1).norm(dim=1).view(-1, 1, 1, 1)
        return g * v / n, self.sd.get(f"{prefix}.bias")

    def bn(self, x, prefix):
        """eval-mode SyncBatchNorm == F.batch_norm with running stats."""
        s = self.sd
        return F.batch_norm(x, s[f"{prefix}.running_mean"], s[f"{prefix}.running_var"],
                            s[f"{prefix}.weight"], s[f"{prefix}.bias"], False, 0.0, BN_EPS)

    def se(self, x, prefix):
        """architecture.py:52-61."""
        s = self.sd
        m = x.mean(dim=(2, 3))
        h = F.relu(F.linear(m, s[f"{prefix}.linear_1.weight"], s[f"{prefix}.linear_1.bias"]))
        gte = torch.sigmoid(F.linear(h, s[f"{prefix}.linear_2.weight"], s[f"{prefix}.linear_2.bias"]))
        return x * gte[:, :, None, None]


def enc_cell(S: _SD, x, cell: EncCell):
    """ResidualCellEncoder.forward, architecture.py:96-136 (+ SkipDown :64-82)."""
    p = cell.prefix
    stride = 2 if cell.down else 1
    r = F.silu(S.bn(x, f"{p}.residual.0"))
    w, b = S.wn(f"{p}.residual.2")
    r = F.conv2d(r, w, b, stride=stride, padding=1)
    r = F.silu(S.bn(r, f"{p}.residual.3"))
    w, b = S.wn(f"{p}.residual.5")
    r = F.conv2d(r, w, b, stride=1, padding=1)
    r = S.se(r, f"{p}.residual.6")
    if cell.down:
        w, b = S.wn(f"{p}.skip_connection.conv")
        x = F.conv2d(F.silu(x), w, b, stride=2)
    return x + 0.1 * r


def dec_cell(S: _SD, x, cell: DecCell):
    """ResidualCellDecoder.forward, architecture.py:139-186 (+ SkipUp :85-93)."""
    p, o = cell.prefix, cell.off
    r = x
    if cell.up:
        r = F.interpolate(r, scale_factor=2, mode="nearest")
    r = S.bn(r, f"{p}.residual.{0 + o}")
    r = F.conv2d(r, S.sd[f"{p}.residual.{1 + o}.weight"])
    r = F.silu(S.bn(r, f"{p}.residual.{2 + o}"))
    r = F.conv2d(r, S.sd[f"{p}.residual.{4 + o}.weight"], padding=2, groups=cell.hidden)
    r = F.silu(S.bn(r, f"{p}.residual.{5 + o}"))
    r = F.conv2d(r, S.sd[f"{p}.residual.{7 + o}.weight"])
    r = S.bn(r, f"{p}.residual.{8 + o}")
    r = S.se(r, f"{p}.residual.{9 + o}")
    if cell.up:
        x = F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=True)
        w, b = S.wn(f"{p}.skip_connection.conv")
        x = F.conv2d(x, w, b)
    return x + 0.1 * r


def disc_mix_logistic_mean(logits: torch.Tensor, n_mix: int) -> torch.Tensor:
    """DiscMixLogistic.__init__ + .mean(), distributions.py:103-129,231-254.  (B, n+9n, H, W) -> (B,3,H,W) in [-1,1]."""
    b, _, h, w = logits.shape
    pi = torch.softmax(logits[:, :n_mix], dim=1)                         # (B,n,H,W)
    rest = logits[:, n_mix:].reshape(b, n_mix, 9, h, w)                  # 'b (n c) h w', c = 9
    means, coef = rest[:, :, 0:3], torch.tanh(rest[:, :, 6:9])           # chunk(3, dim=2): m, s, k
    mu = (means * pi[:, :, None]).sum(1)                                 # (B,3,H,W)
    kk = (coef * pi[:, :, None]).sum(1)
    r = mu[:, 0].clamp(-1, 1)
    g = (mu[:, 1] + kk[:, 0] * r).clamp(-1, 1)
    bl = (mu[:, 2] + kk[:, 1] * r + kk[:, 2] * g).clamp(-1, 1)
    return torch.stack([r, g, bl], dim=1)


def nf_cells(S: _SD, z, cells):
    """NFBlock / NFCell.forward, architecture.py:221-253: z - conv1x1(ELU(dw5x5(ELU(conv3x3(z))))) with every weight multiplied by the
    module's mask buffer (MaskedConv2d.forward :30-34), cell after cell.  (The mask of the 1x1 conv is all zero by construction --
    (1*1)//2 = 0 taps survive -- so each cell only subtracts that conv's bias; the restatement keeps the full arithmetic.)"""
    for q, _ in cells:
        sd = S.sd
        h = F.conv2d(z, sd[f"{q}.0.weight"] * sd[f"{q}.0.mask"], sd[f"{q}.0.bias"], padding=1)
        h = F.elu(h)
        h = F.conv2d(h, sd[f"{q}.2.weight"] * sd[f"{q}.2.mask"], sd[f"{q}.2.bias"], padding=2, groups=h.shape[1])
        h = F.elu(h)
        h = F.conv2d(h, sd[f"{q}.4.weight"] * sd[f"{q}.4.mask"], sd[f"{q}.4.bias"])
        z = z - h
    return z


def nvae_purify(sd: Dict[str, torch.Tensor], spec: NvaeSpec, batch: torch.Tensor, alphas: Sequence[float],
                eps_levels: List[torch.Tensor], temperature: float = 0.6, dtype=torch.float32,
                taps: Optional[dict] = None) -> torch.Tensor:
    """NVAEDefenseModel.purify, src/defenses/ours/models.py:160-274.

    `alphas` are the already-attenuated interpolation alphas (abstract_models.py:107);
    `eps_levels[i]` is the N(0,1) draw of latent level i (distributions.py:43).
    `taps`, if given, receives named intermediate tensors (for kernel-level parity debugging)."""
    S = _SD(sd, dtype)
    x = batch.to(dtype)
    b = x.shape[0]
    x = (x - 0.5) / 0.5                                                   # models.py:170; NVAE/model.py:32
    w, bias = S.wn("preprocessing_block.init_conv")
    x = F.conv2d(x, w, bias, padding=1)
    if taps is not None:
        taps["init_conv"] = x
    for cell in spec.pre_cells:
        x = enc_cell(S, x, cell)
    if taps is not None:
        taps["pre"] = x
    stash = {}
    for sc in spec.enc_scales:                                            # models.py:176-192
        s = sc["s"]
        for g, cells in enumerate(sc["groups"]):
            for cell in cells:
                x = enc_cell(S, x, cell)
            if not (s == 0 and g == 0):
                stash[(s, g)] = x
        if sc["down"] is not None:
            x = enc_cell(S, x, sc["down"])
    w, bias = S.wn("encoder_0.1")                                         # models.py:195; model.py:184-187
    x = F.elu(F.conv2d(F.elu(x), w, bias))
    if taps is not None:
        taps["enc0"] = x
    w, bias = S.wn("enc_sampler.sampler_0:0")                             # models.py:198-206
    mu_q = F.conv2d(x, w, bias, padding=1)[:, :spec.z]
    a0 = float(alphas[0])
    z = (1 - a0) * soft_clamp5(mu_q) + a0 * (eps_levels[0].to(dtype) * temperature)   # prior N(0,1)*temp
    if taps is not None:
        taps["z0"] = z
    if spec.use_nf:                                                       # models.py:209-210
        z = nf_cells(S, z, spec.nf_cells_of(0, 0))
    x = S.sd["const_prior"].expand(b, -1, -1, -1)                         # models.py:215
    w, bias = S.wn("decoder_combiners.combiner_0:0.conv")
    x = F.conv2d(torch.cat([x, z], dim=1), w, bias)                       # architecture.py:215-218
    idx = 1
    for s in range(spec.num_scales):                                      # models.py:222-263
        for lvl in spec.levels:
            if lvl.s != s or (lvl.s == 0 and lvl.g == 0):
                continue
            for cell in lvl.cells:
                x = dec_cell(S, x, cell)
            w, bias = S.wn(f"encoder_combiners.combiner_{lvl.s}:{lvl.g}.conv")
            comb = stash[(lvl.s, lvl.g)] + F.conv2d(x, w, bias)           # architecture.py:195-202
            w, bias = S.wn(f"enc_sampler.sampler_{lvl.s}:{lvl.g}")
            mu_q = F.conv2d(comb, w, bias, padding=1)[:, :spec.z]
            w, bias = S.wn(f"dec_sampler.sampler_{lvl.s}:{lvl.g}.1")
            pp = F.conv2d(F.elu(x), w, bias)
            mu_p, ls_p = pp[:, :spec.z], pp[:, spec.z:]
            a = float(alphas[idx])
            enc_mu = soft_clamp5(mu_p + mu_q)                             # Normal(mu_p+mu_q, .).mu
            dec_sample = soft_clamp5(mu_p) + eps_levels[idx].to(dtype) * (temperature * torch.exp(soft_clamp5(ls_p)))
            z = (1 - a) * enc_mu + a * dec_sample                         # models.py:246-250
            if taps is not None:
                taps[f"z{idx}"] = z
            if spec.use_nf:                                               # models.py:253-254
                z = nf_cells(S, z, spec.nf_cells_of(lvl.s, lvl.g))
            w, bias = S.wn(f"decoder_combiners.combiner_{lvl.s}:{lvl.g}.conv")
            x = F.conv2d(torch.cat([x, z], dim=1), w, bias)
            idx += 1
        if s in spec.up_cells:
            x = dec_cell(S, x, spec.up_cells[s])
    if taps is not None:
        taps["dec_out"] = x
    for cell in spec.post_cells:
        x = dec_cell(S, x, cell)
    if taps is not None:
        taps["post"] = x
    w, bias = S.wn("to_logits.1")                                         # model.py:310-313
    logits = F.conv2d(F.elu(x), w, bias, padding=1)
    if taps is not None:
        taps["logits"] = logits
    rec = disc_mix_logistic_mean(logits, spec.num_mixtures)
    return rec * 0.5 + 0.5                                                # denormalization, models.py:274


# ----------------------------------------------------------------------------- classifier (torchvision body)
def build_vgg11(sd: Dict[str, torch.Tensor], n_classes: int = 100, dtype=torch.float32):
    """src/classifier/model.py:31-50 -- torchvision vgg11_bn with the 4-layer head."""
    from torchvision.models import vgg11_bn
    import torch.nn as nn

    class Vgg(nn.Module):
        def __init__(self):
            super().__init__()
            self.model = vgg11_bn(weights=None)
            d = self.model.classifier[0].weight.shape[1]
            self.model.classifier = nn.Sequential(nn.Linear(d, d, bias=False), nn.BatchNorm1d(d),
                                                  nn.ReLU(inplace=True), nn.Linear(d, n_classes))

        def forward(self, x):
            return self.model(x)

    with torch.device("meta"):
        m = Vgg()
    m.load_state_dict(sd, assign=True)
    return m.to(dtype).eval()


def classify(model, purified: torch.Tensor) -> torch.Tensor:
    """BaseClassificationModel.__call__, abstract_models.py:53-62: normalize(0.5, 0.5) then the net."""
    return model((purified - 0.5) / 0.5)


def defense_call(nvae_sd, spec, vgg_model, batch, alphas, noises, eps: float, blur: bool,
                 temperature: float = 0.6, dtype=torch.float32, taps=None):
    """MLVGMDefenseModel.__call__, abstract_models.py:161-193 for the NVAE model.
    noises[0] is the input-noise draw, noises[1:] the per-level draws."""
    x = preprocess(batch.to(dtype), noises[0].to(dtype), eps, blur)
    if taps is not None:
        taps["preprocessed"] = x
    purified = nvae_purify(nvae_sd, spec, x, alphas, noises[1:], temperature, dtype, taps)
    logits = classify(vgg_model, purified) if vgg_model is not None else None
    return logits, purified


# ----------------------------------------------------------------------------- PGD-Linf (config 5; parity unpinned)
def pgd_linf_step(x_adv, grad, x_nat, step: float, eps: float):
    """Update rule of src/defenses/competitors/trades/modules.py:43-45 (the only L-inf PGD step in
    the reference; there is no PGD attack class in src/attacks -> parity unpinned)."""
    x_adv = x_adv + step * torch.sign(grad)
    x_adv = torch.min(torch.max(x_adv, x_nat - eps), x_nat + eps)
    return x_adv.clamp(0.0, 1.0)


def pgd_linf_attack(logits_fn, x, y, eps: float, step: float, steps: int):
    """PGD-Linf with the update of trades/modules.py:43-45, cross-entropy loss, start at x (documented choices --
    the reference has no PGD attack; parity unpinned).  logits_fn(x_adv, step_index) -> logits (differentiable)."""
    x_adv = x.clone()
    for i in range(steps):
        xa = x_adv.clone().requires_grad_(True)
        loss = F.cross_entropy(logits_fn(xa, i), y)
        grad, = torch.autograd.grad(loss, [xa])
        x_adv = pgd_linf_step(x_adv, grad, x, step, eps).detach()
    return x_adv
