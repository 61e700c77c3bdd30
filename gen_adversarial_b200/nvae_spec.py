"""Structure of the NVAE autoencoder used by the purification path.

The reference keeps the architecture hyper-parameters inside the checkpoint
(`/root/reference/src/defenses/loading_utils.py:57-62`); only "3 scales x 8 groups = 24 latents"
is pinned by the YAML configs.  This module is the ONE place where the synthetic "C32"
configuration lives, and where the module tree of
`/root/reference/src/mlvgms_autoencoders/NVAE/model.py:16-321` is enumerated as plain data
(names, channel counts, strides) so that
  * `synth.py` can write random-init checkpoints with the reference's exact state_dict keys,
  * `nvae_engine.py` can fold weights and run the CUDA path,
  * `oracle/nvae_ref.py` can restate the forward on the CPU.
"""
from __future__ import annotations

from collections import OrderedDict
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

# SURVEY.md section 8d -- declared synthetic configuration ("C32").  Printed in every report.
NVAE_C32_CONFIG = {
    "initial_channels": 32,
    "num_pre-post_process_blocks": 1,
    "num_pre-post_process_cells": 2,
    "num_logistic_mixtures": 10,
    "num_scales": 3,
    "num_groups_per_scale": 8,
    "is_adaptive": False,
    "min_groups_per_scale": 1,
    "num_cells_per_group": 2,
    "num_latent_per_group": 20,
    "num_nf_cells": None,
}
NVAE_C32_RESOLUTION = (3, 64, 64)


def tiny_config(initial_channels: int = 8, groups: int = 2, scales: int = 2, latent: int = 4) -> dict:
    """A small architecture of the same family (used for fast tests and the golden fixtures)."""
    cfg = dict(NVAE_C32_CONFIG)
    cfg.update({"initial_channels": initial_channels, "num_groups_per_scale": groups,
                "num_scales": scales, "num_latent_per_group": latent})
    return cfg


@dataclass
class EncCell:           # ResidualCellEncoder, architecture.py:96-136
    prefix: str
    cin: int
    cout: int
    down: bool


@dataclass
class DecCell:           # ResidualCellDecoder, architecture.py:139-186
    prefix: str
    cin: int
    cout: int
    up: bool
    hidden_mul: int

    @property
    def hidden(self) -> int:
        return self.cin * self.hidden_mul

    @property
    def off(self) -> int:   # index shift of residual.* when the up-sampling layer is present
        return 1 if self.up else 0


@dataclass
class Level:             # one latent group (s, g)
    s: int
    g: int
    channels: int
    res: int             # spatial size of the level
    cells: List[DecCell] = field(default_factory=list)


class NvaeSpec:
    def __init__(self, cfg: dict, resolution: Tuple[int, int, int]):
        self.cfg = dict(cfg)
        self.resolution = tuple(resolution)
        self.img_channels, self.image_resolution, _ = self.resolution
        c0 = cfg["initial_channels"]
        self.base_channels = c0
        self.n_blocks = cfg["num_pre-post_process_blocks"]
        self.n_cells_block = cfg["num_pre-post_process_cells"]
        self.num_mixtures = cfg["num_logistic_mixtures"]
        self.num_scales = cfg["num_scales"]
        if cfg["is_adaptive"]:
            gps = [max(cfg["min_groups_per_scale"], cfg["num_groups_per_scale"] // (2 ** i))
                   for i in range(self.num_scales)]
        else:
            gps = [cfg["num_groups_per_scale"]] * self.num_scales
        gps.reverse()                                   # model.py:46-52
        self.groups_per_scale = gps
        self.num_cells_per_group = cfg["num_cells_per_group"]
        self.z = cfg["num_latent_per_group"]
        self.num_nf_cells = cfg.get("num_nf_cells")
        self.use_nf = self.num_nf_cells is not None
        self.n_latents = sum(gps)
        self.scaling_factor = 2 ** (self.n_blocks + self.num_scales - 1)
        self.logit_channels = self.num_mixtures + self.num_mixtures * 3 * self.img_channels

        # ---- pre-processing (model.py:97-130)
        mult = 1
        self.pre_cells: List[EncCell] = []
        for b in range(self.n_blocks):
            for c in range(self.n_cells_block):
                last = c == self.n_cells_block - 1
                ch = c0 * mult
                pfx = f"preprocessing_block.block_{b}.cell_{c}"
                if not last:
                    self.pre_cells.append(EncCell(pfx, ch, ch, False))
                else:
                    self.pre_cells.append(EncCell(pfx, ch, ch * 2, True))
                    mult *= 2
        # ---- encoder tower (model.py:132-189); executed s = S-1 .. 0, g = 0 .. G-1 (models.py:176-192)
        self.enc_scales: List[dict] = []
        res = self.image_resolution // (2 ** self.n_blocks)
        for s in range(self.num_scales - 1, -1, -1):
            ch = c0 * mult
            groups = []
            for g in range(self.groups_per_scale[s]):
                cells = [EncCell(f"encoder_tower.scale_{s}.group_{g}.cell_{c}", ch, ch, False)
                         for c in range(self.num_cells_per_group)]
                groups.append(cells)
            down = None
            if s > 0:
                down = EncCell(f"encoder_tower.scale_{s}.downsampling", ch, ch * 2, True)
                mult *= 2
            self.enc_scales.append({"s": s, "channels": ch, "res": res, "groups": groups, "down": down})
            if s > 0:
                res //= 2
        self.top_channels = c0 * mult
        self.top_res = res
        # ---- decoder tower (model.py:233-270)
        self.levels: List[Level] = []
        self.up_cells: Dict[int, DecCell] = {}
        dmult = mult
        for s in range(self.num_scales):
            ch = c0 * dmult
            for g in range(self.groups_per_scale[s]):
                lvl = Level(s, g, ch, res)
                if not (s == 0 and g == 0):
                    lvl.cells = [DecCell(f"decoder_tower.scale_{s}.group_{g}.cell_{c}", ch, ch, False, 6)
                                 for c in range(self.num_cells_per_group)]
                self.levels.append(lvl)
            if s < self.num_scales - 1:
                self.up_cells[s] = DecCell(f"decoder_tower.scale_{s}.upsampling", ch, ch // 2, True, 6)
                dmult //= 2
                res *= 2
        # ---- post-processing (model.py:272-298)
        self.post_cells: List[DecCell] = []
        for b in range(self.n_blocks):
            for c in range(self.n_cells_block):
                ch = c0 * dmult
                pfx = f"postprocessing_block.block_{b}.cell_{c}"
                if c != 0:
                    self.post_cells.append(DecCell(pfx, ch, ch, False, 3))
                else:
                    self.post_cells.append(DecCell(pfx, ch, ch // 2, True, 3))
                    dmult //= 2
        self.out_channels = c0 * dmult

    # ------------------------------------------------------------------ helpers
    def level_index(self, s: int, g: int) -> int:
        return sum(self.groups_per_scale[:s]) + g

    def all_enc_cells(self) -> List[EncCell]:
        out = list(self.pre_cells)
        for sc in self.enc_scales:
            for grp in sc["groups"]:
                out += grp
            if sc["down"] is not None:
                out.append(sc["down"])
        return out

    def all_dec_cells(self) -> List[DecCell]:
        out = []
        for s in range(self.num_scales):
            for lvl in self.levels:
                if lvl.s == s:
                    out += lvl.cells
            if s in self.up_cells:
                out.append(self.up_cells[s])
        out += self.post_cells
        return out

    # ------------------------------------------------------------------ state_dict layout
    @staticmethod
    def _wn_conv(sd, prefix, cout, cin, k, bias=True):
        sd[f"{prefix}.bias"] = (cout,) if bias else None
        sd[f"{prefix}.parametrizations.weight.original0"] = (cout, 1, 1, 1)
        sd[f"{prefix}.parametrizations.weight.original1"] = (cout, cin, k, k)

    @staticmethod
    def _bn(sd, prefix, c):
        sd[f"{prefix}.weight"] = (c,)
        sd[f"{prefix}.bias"] = (c,)
        sd[f"{prefix}.running_mean"] = (c,)
        sd[f"{prefix}.running_var"] = (c,)
        sd[f"{prefix}.num_batches_tracked"] = ()

    @staticmethod
    def _se(sd, prefix, c):
        h = max(c // 16, 4)
        sd[f"{prefix}.linear_1.weight"] = (h, c)
        sd[f"{prefix}.linear_1.bias"] = (h,)
        sd[f"{prefix}.linear_2.weight"] = (c, h)
        sd[f"{prefix}.linear_2.bias"] = (c,)

    def state_dict_shapes(self) -> "OrderedDict[str, tuple]":
        """key -> shape of the reference `AutoEncoder.state_dict()` (order is not significant)."""
        sd: "OrderedDict[str, Optional[tuple]]" = OrderedDict()
        c0 = self.base_channels
        r0 = self.image_resolution // self.scaling_factor
        sd["const_prior"] = (1, int(self.scaling_factor * c0), r0, r0)
        self._wn_conv(sd, "preprocessing_block.init_conv", c0, self.img_channels, 3)
        for cell in self.all_enc_cells():
            p = cell.prefix
            if cell.down:
                self._wn_conv(sd, f"{p}.skip_connection.conv", cell.cout, cell.cin, 1)
            self._bn(sd, f"{p}.residual.0", cell.cin)
            self._wn_conv(sd, f"{p}.residual.2", cell.cout, cell.cin, 3)
            self._bn(sd, f"{p}.residual.3", cell.cout)
            self._wn_conv(sd, f"{p}.residual.5", cell.cout, cell.cout, 3)
            self._se(sd, f"{p}.residual.6", cell.cout)
        for sc in self.enc_scales:
            for g in range(len(sc["groups"])):
                if not (sc["s"] == 0 and g == 0):
                    self._wn_conv(sd, f"encoder_combiners.combiner_{sc['s']}:{g}.conv", sc["channels"], sc["channels"], 1)
        self._wn_conv(sd, "encoder_0.1", self.top_channels, self.top_channels, 1)
        for lvl in self.levels:
            self._wn_conv(sd, f"enc_sampler.sampler_{lvl.s}:{lvl.g}", 2 * self.z, lvl.channels, 3)
            if not (lvl.s == 0 and lvl.g == 0):
                self._wn_conv(sd, f"dec_sampler.sampler_{lvl.s}:{lvl.g}.1", 2 * self.z, lvl.channels, 1)
            self._wn_conv(sd, f"decoder_combiners.combiner_{lvl.s}:{lvl.g}.conv", lvl.channels, lvl.channels + self.z, 1)
        for cell in self.all_dec_cells():
            p, o = cell.prefix, cell.off
            if cell.up:
                self._wn_conv(sd, f"{p}.skip_connection.conv", cell.cout, cell.cin, 1)
            self._bn(sd, f"{p}.residual.{0 + o}", cell.cin)
            sd[f"{p}.residual.{1 + o}.weight"] = (cell.hidden, cell.cin, 1, 1)
            self._bn(sd, f"{p}.residual.{2 + o}", cell.hidden)
            sd[f"{p}.residual.{4 + o}.weight"] = (cell.hidden, 1, 5, 5)
            self._bn(sd, f"{p}.residual.{5 + o}", cell.hidden)
            sd[f"{p}.residual.{7 + o}.weight"] = (cell.cout, cell.hidden, 1, 1)
            self._bn(sd, f"{p}.residual.{8 + o}", cell.cout)
            self._se(sd, f"{p}.residual.{9 + o}", cell.cout)
        self._wn_conv(sd, "to_logits.1", self.logit_channels, self.out_channels, 3)
        if self.use_nf:
            # NVAE/model.py:216-221: one nn.Sequential of `num_nf_cells` NFBlocks per latent level; NFBlock = cell1 (mirror False) +
            # cell2 (mirror True); NFCell.layers = [MaskedConv2d 3x3 z -> 6z, ELU, MaskedConv2d dw5x5, ELU, MaskedConv2d 1x1 6z -> z]
            # (architecture.py:221-253); every MaskedConv2d registers its `mask` as a buffer (:9-28)
            z, hz = self.z, 6 * self.z
            for lvl in self.levels:
                for n in range(self.num_nf_cells):
                    for cell in ("cell1", "cell2"):
                        q = f"nf_cells.nf_{lvl.s}:{lvl.g}.{n}.{cell}.layers"
                        for idx, shape in ((0, (hz, z, 3, 3)), (2, (hz, 1, 5, 5)), (4, (z, hz, 1, 1))):
                            sd[f"{q}.{idx}.weight"] = shape
                            sd[f"{q}.{idx}.bias"] = (shape[0],)
                            sd[f"{q}.{idx}.mask"] = shape
        return OrderedDict((k, v) for k, v in sd.items() if v is not None)

    @staticmethod
    def nf_mask(shape, mirror: bool, zero_diag: bool):
        """MaskedConv2d mask, architecture.py:17-28: taps in row-major order, the first (h*w)//2 (+1 with zero_diag) stay, mirrored = flipped"""
        import torch
        co, ci, h, w = shape
        m = torch.ones(co, ci, h * w)
        half = (h * w) // 2 + int(zero_diag)
        m[:, :, half:] = 0
        if mirror:
            m = torch.flip(m, dims=(2,))
        return m.view(co, ci, h, w)

    def nf_cells_of(self, s: int, g: int):
        """[(layers-prefix, mirror)] of the NF cells applied to z of level (s, g), in execution order"""
        if not self.use_nf:
            return []
        return [(f"nf_cells.nf_{s}:{g}.{n}.{cell}.layers", cell == "cell2") for n in range(self.num_nf_cells) for cell in ("cell1", "cell2")]

    def noise_shapes(self, batch: int) -> List[tuple]:
        """RNG draw order of one `__call__`: input noise, then one eps per latent level."""
        shapes = [(batch, self.img_channels, self.image_resolution, self.image_resolution)]
        for lvl in self.levels:
            shapes.append((batch, self.z, lvl.res, lvl.res))
        return shapes

    def describe(self) -> str:
        return (f"NVAE cfg={self.cfg} resolution={self.resolution} latents={self.n_latents} "
                f"groups_per_scale={self.groups_per_scale}")
