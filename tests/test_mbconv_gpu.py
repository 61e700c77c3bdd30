"""GPU: the fused decoder-cell kernel (expand 1x1 -> SiLU -> depthwise 5x5 -> SiLU -> project 1x1, hidden tensor on chip)
against the torch restatement of the three ops (tests/emu_ops.py) and against the three separate CUDA kernels."""
import math

import pytest
import torch

from gen_adversarial_b200 import ops
from gen_adversarial_b200._lib import ACT_NONE, ACT_SILU
from tests import emu_ops

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _cell(c, hidden, seed):
    g = torch.Generator().manual_seed(seed)
    we = torch.randn(hidden, c, generator=g) / math.sqrt(c)
    wp = torch.randn(c, hidden, generator=g) / math.sqrt(hidden)
    e = ops.ConvLayer(1, 1, 1, 0, c, hidden, post_act=ACT_SILU, name="expand")
    e.w_tc = we.to(torch.bfloat16).contiguous()
    e.bias = torch.randn(hidden, generator=g) * 0.3
    p = ops.ConvLayer(1, 1, 1, 0, hidden, c, post_act=ACT_NONE, name="project")
    p.w_tc = wp.to(torch.bfloat16).contiguous()
    p.bias = torch.randn(c, generator=g) * 0.3
    dw_w = torch.randn(25, hidden, generator=g) / 5.0
    dw_b = torch.randn(hidden, generator=g) * 0.3
    return e, dw_w, dw_b, p


def _dev(L):
    import copy
    D = copy.copy(L)
    D.w_tc, D.bias = L.w_tc.to(DEV), L.bias.to(DEV)
    return D


@pytest.mark.parametrize("n,w,c,hidden", [(3, 8, 256, 1536), (2, 8, 256, 64), (2, 16, 128, 768), (1, 16, 128, 128),
                                          (2, 32, 64, 384), (1, 32, 64, 192), (5, 32, 64, 64),
                                          # more tiles than SMs: the persistent CTAs walk 2-3 tiles each (cross-tile pipelining of the rings)
                                          (75, 32, 64, 128), (40, 32, 64, 64), (301, 8, 256, 128), (297, 8, 256, 64), (150, 16, 128, 192),
                                          (149, 16, 128, 64)])
def test_mbconv_fused_matches_three_kernels(n, w, c, hidden):
    e, dw_w, dw_b, p = _cell(c, hidden, seed=n + w + hidden)
    x = torch.randn(n, w, w, c, generator=torch.Generator().manual_seed(1)).to(torch.bfloat16)
    ref = emu_ops.mbconv_fused(x, e, ops.dw_weights_chunked(dw_w), dw_b, p).float()
    ed, pd, xd, wd, bd = _dev(e), _dev(p), x.to(DEV), dw_w.to(DEV), dw_b.to(DEV)
    assert ops.mbconv_fused_supported(xd, ed, pd)
    got = ops.mbconv_fused(xd, ed, ops.dw_weights_chunked(wd), bd, pd)
    torch.cuda.synchronize()
    scale = ref.abs().max().item()
    err = (got.float().cpu() - ref).abs().max().item()
    # the same cell through the three separate kernels (what the fused kernel replaces)
    h1, _ = ops.conv2d_tc(xd, ed)
    h2 = ops.dwconv5x5(h1, wd, bd, ACT_SILU, False, torch.bfloat16)
    r3, _ = ops.conv2d_tc(h2, pd)
    err3 = (got.float() - r3.float()).abs().max().item()
    print(f"mbconv n={n} w={w} c={c} hidden={hidden}: vs torch {err:.3e}, vs 3 kernels {err3:.3e} (scale {scale:.2f})")
    assert err <= 2e-2 * scale, (err, scale)
    assert err3 <= 2e-2 * scale, (err3, scale)


@pytest.mark.parametrize("n,w,c,hidden", [(3, 8, 256, 128), (301, 8, 256, 64), (2, 16, 128, 128), (150, 16, 128, 64), (3, 32, 64, 128), (40, 32, 64, 64)])
def test_mbconv_fused_channel_sums(n, w, c, hidden):
    """SE squeeze from the fused cell's epilogue: the sums of r AS STORED (bf16), per image and 128-pixel slice, in the layout `channel_sum`
    writes; r itself is unchanged by asking for them."""
    e, dw_w, dw_b, p = _cell(c, hidden, seed=n + w + hidden)
    x = torch.randn(n, w, w, c, generator=torch.Generator().manual_seed(2)).to(torch.bfloat16)
    ed, pd, xd, wd, bd = _dev(e), _dev(p), x.to(DEV), dw_w.to(DEV), dw_b.to(DEV)
    wc = ops.dw_weights_chunked(wd)
    r0 = ops.mbconv_fused(xd, ed, wc, bd, pd)
    r, sums = ops.mbconv_fused(xd, ed, wc, bd, pd, want_sums=True)
    ref = ops.channel_sum(r)
    torch.cuda.synchronize()
    assert torch.equal(r0, r)
    assert sums.shape == ref.shape
    # same addends, different association: fp32 rounding only
    scale = max(1.0, ref.abs().max().item())
    assert (sums - ref).abs().max().item() <= 2e-5 * scale * (min(w * w, 128) ** 0.5)


@pytest.mark.parametrize("n,w,c,hidden", [(3, 8, 256, 128), (301, 8, 256, 64), (2, 16, 128, 192), (150, 16, 128, 64), (3, 32, 64, 128), (40, 32, 64, 64)])
def test_mbconv_fused_tape(n, w, c, hidden):
    """Taping variant (attack path): r and the channel sums are those of the plain fused cell, and the two tapes are SiLU' of the expand /
    depthwise pre-activations -- compared with the three taping kernels it replaces (bf16 tapes of O(1) values: a few bf16 ulps; the fused cell
    keeps the hidden tile in fp16 where the three kernels round it to bf16)."""
    e, dw_w, dw_b, p = _cell(c, hidden, seed=n + w + hidden)
    x = torch.randn(n, w, w, c, generator=torch.Generator().manual_seed(3)).to(torch.bfloat16)
    ed, pd, xd, wd, bd = _dev(e), _dev(p), x.to(DEV), dw_w.to(DEV), dw_b.to(DEV)
    wc = ops.dw_weights_chunked(wd)
    r0, s0 = ops.mbconv_fused(xd, ed, wc, bd, pd, want_sums=True)
    r, s, de, dd = ops.mbconv_fused(xd, ed, wc, bd, pd, want_sums=True, want_tape=True)
    de3 = torch.empty(n, w, w, hidden, device=DEV, dtype=torch.bfloat16)
    h1, _ = ops.conv2d_tc(xd, ed, dact_out=de3)
    h2, dd3 = ops.dwconv5x5(h1, wd, bd, ACT_SILU, False, torch.bfloat16, want_dact=True)
    torch.cuda.synchronize()
    assert torch.equal(r0, r) and torch.equal(s0, s)
    assert de.shape == (n, w, w, hidden) and dd.shape == (n, w, w, hidden)
    assert (de.float() - de3.float()).abs().max().item() <= 1.6e-2          # same accumulator, same formula: at most one bf16 ulp of ~1
    err_dd = (dd.float() - dd3.float()).abs()
    assert err_dd.max().item() <= 6e-2 and err_dd.mean().item() <= 4e-3, (err_dd.max().item(), err_dd.mean().item())


@pytest.mark.parametrize("with_add", [True, False])
@pytest.mark.parametrize("n,w,c,hidden", [(3, 8, 256, 128), (301, 8, 256, 64), (2, 16, 128, 192), (150, 16, 128, 64), (3, 32, 64, 128), (40, 32, 64, 64)])
def test_mbconv_fused_backward_matches_three_kernels(n, w, c, hidden, with_add):
    """Input gradient of the cell in one kernel against the three kernels of the backward sweep it replaces (project dgrad x tape, transposed
    depthwise x tape, expand dgrad + skip gradient): same bf16 intermediates, fp32 accumulation, fp32 result."""
    e, dw_w, dw_b, p = _cell(c, hidden, seed=n + w + hidden)
    gen = torch.Generator().manual_seed(7)
    g = (torch.randn(n, w, w, c, generator=gen) * 1e-3).to(torch.bfloat16).to(DEV)
    dact_dw = (torch.rand(n, w, w, hidden, generator=gen) * 1.2 - 0.1).to(torch.bfloat16).to(DEV)
    dact_e = (torch.rand(n, w, w, hidden, generator=gen) * 1.2 - 0.1).to(torch.bfloat16).to(DEV)
    add = (torch.randn(n, w, w, c, generator=gen) * 1e-3).to(DEV) if with_add else None
    p_d = ops.ConvLayer(1, 1, 1, 0, c, hidden, post_act=ACT_NONE, name="project_dgrad")
    p_d.w_tc = p.w_tc.t().contiguous().to(DEV)           # [hidden][C]
    e_d = ops.ConvLayer(1, 1, 1, 0, hidden, c, post_act=ACT_NONE, name="expand_dgrad")
    e_d.w_tc = e.w_tc.t().contiguous().to(DEV)           # [C][hidden]
    dw_wT = dw_w.flip(0).contiguous().to(DEV)
    g_v2, _ = ops.conv2d_tc(g, p_d, mul=dact_dw)
    g_v1 = ops.dwconv5x5(g_v2, dw_wT, None, ACT_NONE, False, torch.bfloat16, mul=dact_e)
    _, ref = ops.conv2d_tc(g_v1, e_d, want_bf16=False, want_f32=True, add=add)
    got = ops.mbconv_fused_bwd(g, p_d, ops.dw_weights_chunked(dw_wT), dact_dw, dact_e, e_d, add=add)
    torch.cuda.synchronize()
    scale = ref.abs().max().item()
    err = (got - ref).abs().max().item()
    rel = ((got - ref).norm() / ref.norm()).item()
    print(f"mbconv bwd n={n} w={w} c={c} hidden={hidden} add={with_add}: max err {err:.3e} (scale {scale:.3e}), rel-L2 {rel:.3e}")
    assert torch.isfinite(got).all()
    assert rel <= 1e-2 and err <= 3e-2 * scale
