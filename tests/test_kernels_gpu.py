"""GPU: every kernel of libga_b200 against a torch restatement of the same op (tests/emu_ops.py, CPU fp32),
through the C-ABI.  fp32 kernels: tight tolerances; bf16 tensor-core kernels: bf16-rounding tolerances."""
import math

import pytest
import torch

from gen_adversarial_b200 import ops
from gen_adversarial_b200._lib import PRE_NONE, PRE_ELU, PRE_SILU, PRE_AFFINE_SILU, ACT_NONE, ACT_SILU, ACT_ELU, ACT_RELU
from oracle import nvae_ref
from tests import emu_ops

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _layer(cin, cout, k, stride=1, pad=0, pre_op=PRE_NONE, post_act=ACT_NONE, bias=True, up=1, tc=False, cin2=0, seed=0):
    g = torch.Generator().manual_seed(seed)
    w = torch.randn(cout, cin, k, k, generator=g) / math.sqrt(cin * k * k)
    L = ops.ConvLayer(k, k, stride, pad, cin, cout, pre_op=pre_op, post_act=post_act, up=up, name=f"t{cin}x{cout}k{k}")
    L.w_simt = w.permute(2, 3, 1, 0).reshape(k * k * cin, cout).contiguous()
    if tc:
        wk = w.permute(0, 2, 3, 1).reshape(cout, k * k * cin)
        if cin2:
            wk = torch.cat([wk, torch.randn(cout, cin2, generator=g) / math.sqrt(cin2)], dim=1)
            L.cin2 = cin2
        L.w_tc = wk.to(torch.bfloat16).contiguous()
    if bias:
        L.bias = torch.randn(cout, generator=g) * 0.1
    if pre_op == PRE_AFFINE_SILU:
        L.pre_scale = torch.rand(cin, generator=g) + 0.5
        L.pre_shift = torch.randn(cin, generator=g) * 0.2
    return L


def _to_dev(L):
    import copy
    D = copy.copy(L)
    for f in ("w_simt", "w_tc", "bias", "pre_scale", "pre_shift"):
        v = getattr(L, f)
        setattr(D, f, None if v is None else v.to(DEV))
    return D


SIMT_CASES = [
    # n, h, w, cin, cout, k, stride, pad, pre, post, add, up
    (2, 16, 16, 3, 32, 3, 1, 1, PRE_NONE, ACT_NONE, False, 1),
    (3, 16, 16, 32, 32, 3, 1, 1, PRE_AFFINE_SILU, ACT_SILU, False, 1),
    (2, 16, 16, 32, 64, 3, 2, 1, PRE_AFFINE_SILU, ACT_SILU, False, 1),
    (2, 16, 16, 32, 64, 1, 2, 0, PRE_SILU, ACT_NONE, False, 1),
    (2, 8, 8, 64, 40, 1, 1, 0, PRE_ELU, ACT_NONE, False, 1),
    (2, 8, 8, 24, 64, 1, 1, 0, PRE_NONE, ACT_NONE, True, 1),
    (5, 4, 4, 20, 13, 3, 1, 1, PRE_NONE, ACT_ELU, True, 1),
    (2, 8, 8, 64, 32, 3, 1, 1, PRE_NONE, ACT_NONE, False, 2),      # transposed (dgrad of stride 2)
    (130, 1, 1, 200, 100, 1, 1, 0, PRE_NONE, ACT_RELU, False, 1),  # linear
]


@pytest.mark.parametrize("case", SIMT_CASES)
def test_conv_simt_fp32(case):
    n, h, w, cin, cout, k, stride, pad, pre, post, use_add, up = case
    L = _layer(cin, cout, k, stride, pad, pre, post, up=up, seed=cin + cout)
    if up > 1:
        L.pad = k - 1 - pad
    x = torch.randn(n, h, w, cin)
    out_hw = (2 * h, 2 * w) if up > 1 else None
    ho, wo = out_hw if out_hw else ops.conv_out_hw(L, h, w)
    add = torch.randn(n, ho, wo, cout) if use_add else None
    ref = emu_ops.conv2d_simt(x, L, torch.float32, add=add, out_hw=out_hw)
    got = ops.conv2d_simt(x.to(DEV), _to_dev(L), torch.float32, add=add.to(DEV) if use_add else None, out_hw=out_hw)
    assert got.shape == ref.shape
    err = (got.cpu() - ref).abs().max().item()
    assert err <= 2e-5, err


def test_conv_simt_bf16_io():
    L = _layer(32, 64, 3, 2, 1, PRE_AFFINE_SILU, ACT_SILU)
    x = torch.randn(2, 16, 16, 32).to(torch.bfloat16)
    ref = emu_ops.conv2d_simt(x, L, torch.float32)
    got = ops.conv2d_simt(x.to(DEV), _to_dev(L), torch.bfloat16)
    assert (got.float().cpu() - ref).abs().max().item() <= 2e-2


TC_CASES = [
    # n, h, w, cin, cout, k, post, add('f32'|'bf16'|None), cin2
    (4, 8, 8, 64, 64, 1, ACT_NONE, None, 0),
    (2, 32, 32, 64, 64, 3, ACT_SILU, None, 0),
    (2, 16, 16, 128, 128, 3, ACT_NONE, None, 0),
    (4, 8, 8, 256, 256, 3, ACT_SILU, None, 0),
    (3, 8, 8, 256, 1536, 1, ACT_SILU, None, 0),
    (3, 8, 8, 1536, 256, 1, ACT_NONE, None, 0),
    (2, 16, 16, 128, 20, 3, ACT_NONE, None, 0),          # sampler: N = 20 (padded tile, scalar stores)
    (2, 16, 16, 128, 40, 1, ACT_NONE, None, 0),
    (2, 8, 8, 256, 256, 1, ACT_NONE, "f32", 0),           # encoder combiner: + stash
    (2, 8, 8, 256, 256, 1, ACT_NONE, "bf16", 24),         # decoder combiner: second K source (z, 24 ch)
    (2, 64, 64, 32, 32, 3, ACT_SILU, None, 0),            # Cin = 32 < 64: TMA zero fill in K
    (2, 64, 64, 96, 32, 1, ACT_NONE, None, 0),            # K = 96 (not a multiple of 64)
    (2, 64, 64, 32, 100, 3, ACT_NONE, None, 0),           # to_logits
    (2, 32, 32, 96, 64, 3, ACT_SILU, "bf16", 0),          # Cin = 96: three 32-channel K blocks per tap (64-byte swizzle path)
    (1, 128, 128, 32, 32, 3, ACT_NONE, None, 0),          # Cin = 32, W = 128 (one image row per tile)
    (3, 8, 8, 24, 256, 1, ACT_NONE, "f32", 0),            # level 0: z only + prior
    (5, 4, 4, 512, 512, 3, ACT_RELU, None, 0),            # VGG 4x4 (8 images per tile, ragged batch)
    (5, 2, 2, 512, 512, 3, ACT_RELU, None, 0),
    (130, 1, 1, 2048, 384, 1, ACT_RELU, None, 0),         # linear, ragged M
    (3, 1, 1, 25088, 100, 1, ACT_NONE, None, 0),          # long-K linear (392 K blocks)
]


@pytest.mark.parametrize("case", TC_CASES)
def test_conv_tc_bf16(case):
    n, h, w, cin, cout, k, post, addk, cin2 = case
    L = _layer(cin, cout, k, 1, 1 if k == 3 else 0, PRE_NONE, post, tc=True, cin2=cin2, seed=cin + cout + k)
    g = torch.Generator().manual_seed(1)
    x = torch.randn(n, h, w, cin, generator=g).to(torch.bfloat16)
    x2 = torch.randn(n, h, w, cin2, generator=g).to(torch.bfloat16) if cin2 else None
    add = None
    if addk == "f32":
        add = torch.randn(n, h, w, cout, generator=g)
    elif addk == "bf16":
        add = torch.randn(n, h, w, cout, generator=g).to(torch.bfloat16)
    LD = _to_dev(L)
    assert ops.conv2d_tc_supported(x.to(DEV), LD, x2.to(DEV) if cin2 else None)
    _, ref = emu_ops.conv2d_tc(x, L, want_bf16=False, want_f32=True, add=add, x2=x2)
    ob, of = ops.conv2d_tc(x.to(DEV), LD, want_bf16=True, want_f32=True, add=add.to(DEV) if add is not None else None,
                           x2=x2.to(DEV) if cin2 else None)
    torch.cuda.synchronize()
    scale = max(1.0, ref.abs().max().item())
    e32 = (of.cpu() - ref).abs().max().item()
    e16 = (ob.float().cpu() - ref).abs().max().item()
    assert e32 <= 2e-3 * scale, ("fp32 out", e32, scale)       # fp32 accumulation-order differences only
    assert e16 <= 1e-2 * scale, ("bf16 out", e16, scale)


HALO_CASES = [
    # persistent halo-reuse kernel (conv3x3_tc.cu): n, h, w, cin, cout, post, add, mul_mode (None = no mul), want (bf16, f32), dact
    (40, 32, 32, 64, 64, ACT_SILU, None, None, (True, False), True),      # encoder conv1, taping: 160 tiles > 148 CTAs, resident weights
    (150, 16, 16, 128, 128, ACT_NONE, None, None, (True, False), False),  # 150 tiles, streaming weights, two K blocks
    (37, 32, 32, 64, 128, ACT_RELU, None, None, (True, True), False),     # VGG-like: both outputs
    (20, 16, 16, 128, 256, ACT_RELU, None, None, (False, True), False),   # two N blocks
    (9, 16, 16, 256, 256, ACT_NONE, "f32", None, (False, True), False),   # four K blocks, + add
    (6, 64, 64, 64, 64, ACT_NONE, None, None, (True, False), False),      # W = 64: 4-row tiles
    (3, 32, 32, 64, 20, ACT_NONE, None, None, (False, True), False),      # sampler: 20 output channels, fp32 only
    (5, 32, 32, 128, 64, ACT_NONE, "f32", 0, (False, True), False),       # dgrad style: (conv + add) * mul, fp32 gradient out
    (5, 16, 16, 128, 128, ACT_NONE, None, 1, (True, False), False),       # VGG backward: times ReLU mask of mul
    (1, 32, 32, 64, 64, ACT_SILU, None, None, (True, False), False),      # fewer tiles than SMs
    (300, 8, 32, 64, 64, ACT_NONE, None, None, (True, False), False),     # H = 8: one tile per image, 300 tiles
    # images wider than 128 pixels (StyleGAN layers): 8 x 32 windows, the output slabs are row segments of the image
    (2, 16, 256, 64, 64, ACT_NONE, None, None, (True, False), False),     # 32 windows, resident weights
    (1, 8, 512, 64, 128, ACT_NONE, None, None, (False, True), False),     # phase conv of an up layer: fp32 conv output, streaming weights
    (3, 24, 128, 128, 64, ACT_SILU, None, None, (True, True), True),      # W = 128: full-width 2-row tiles, both outputs + tape
    (2, 24, 384, 128, 64, ACT_SILU, None, None, (True, True), True),      # W = 384: 12 windows per row, both outputs + tape
    (2, 32, 256, 64, 256, ACT_RELU, "f32", None, (True, False), False),   # two N blocks, + add
    (1, 64, 256, 128, 128, ACT_NONE, None, 0, (True, False), False),      # times mul
    (37, 8, 256, 64, 32, ACT_NONE, None, None, (True, False), False),     # 296 windows > 148 CTAs, 32 output channels
]


@pytest.mark.parametrize("case", HALO_CASES)
def test_conv3x3_halo_kernel(case):
    n, h, w, cin, cout, post, addk, mul_mode, (wb, wf), want_dact = case
    L = _layer(cin, cout, 3, 1, 1, PRE_NONE, post, tc=True, seed=cin + cout + n)
    g = torch.Generator().manual_seed(n)
    x = torch.randn(n, h, w, cin, generator=g).to(torch.bfloat16)
    add = torch.randn(n, h, w, cout, generator=g) if addk == "f32" else None
    mul = torch.randn(n, h, w, cout, generator=g).to(torch.bfloat16) if mul_mode is not None else None
    LD = _to_dev(L)
    dref = torch.empty(n, h, w, cout) if want_dact else None
    _, ref = emu_ops.conv2d_tc(x, L, want_bf16=False, want_f32=True, add=add, mul=mul, mul_mode=mul_mode or 0, dact_out=dref)
    dgot = torch.empty(n, h, w, cout, device=DEV, dtype=torch.bfloat16) if want_dact else None
    ob, of = ops.conv2d_tc(x.to(DEV), LD, want_bf16=wb, want_f32=wf, add=add.to(DEV) if add is not None else None,
                           mul=mul.to(DEV) if mul is not None else None, mul_mode=mul_mode or 0, dact_out=dgot)
    torch.cuda.synchronize()
    scale = max(1.0, ref.abs().max().item())
    if wf:
        e32 = (of.cpu() - ref).abs().max().item()
        assert e32 <= 2e-3 * scale, ("fp32 out", e32, scale)
    if wb:
        e16 = (ob.float().cpu() - ref).abs().max().item()
        assert e16 <= 1e-2 * scale, ("bf16 out", e16, scale)
    if want_dact:
        assert (dgot.float().cpu() - dref).abs().max().item() <= 2e-2


P1X1_CASES = [
    # 1x1 convolutions at M >= 16384 rows, where the persistent pipeline takes the plain epilogues (and, with GA_TC_P1X1=2, the add/mul ones):
    # n, h, w, cin, cout, post, add, mul_mode, want (bf16, f32), dact
    (64, 32, 32, 64, 64, ACT_NONE, "f32", None, (True, False), False),      # encoder combiner: + stash (lean add epilogue), resident weights
    (64, 32, 32, 64, 384, ACT_SILU, None, None, (True, False), True),       # decoder expand, taping: three N blocks, streaming weights
    (64, 32, 32, 384, 64, ACT_NONE, None, None, (True, False), False),      # decoder project: six K blocks, resident weights
    (64, 32, 32, 384, 64, ACT_NONE, "f32", None, (False, True), False),     # expand dgrad: fp32 stream gradient + add
    (64, 32, 32, 64, 384, ACT_NONE, None, 0, (True, False), False),         # project dgrad times SiLU' tape (bf16 mul)
    (100, 16, 16, 128, 40, ACT_NONE, None, None, (False, True), False),     # decoder sampler: 40 fp32 channels
    (257, 8, 8, 64, 128, ACT_RELU, None, None, (True, False), False),       # ragged: M = 16448 = 64 tiles of 256 + 64 rows
    (75, 16, 16, 128, 64, ACT_NONE, "f32", 1, (True, False), False),        # M = 19200 = 75 tiles of 256 with add and ReLU-mask mul
    (4, 64, 64, 1536, 256, ACT_NONE, None, None, (True, False), False),     # K = 1536 (24 blocks), two N blocks
]


@pytest.mark.parametrize("case", P1X1_CASES)
def test_conv1x1_persistent_kernel(case):
    n, h, w, cin, cout, post, addk, mul_mode, (wb, wf), want_dact = case
    L = _layer(cin, cout, 1, 1, 0, PRE_NONE, post, tc=True, seed=cin + cout + n)
    g = torch.Generator().manual_seed(n)
    x = torch.randn(n, h, w, cin, generator=g).to(torch.bfloat16)
    add = torch.randn(n, h, w, cout, generator=g) if addk == "f32" else None
    mul = torch.randn(n, h, w, cout, generator=g).to(torch.bfloat16) if mul_mode is not None else None
    LD = _to_dev(L)
    dref = torch.empty(n, h, w, cout) if want_dact else None
    _, ref = emu_ops.conv2d_tc(x, L, want_bf16=False, want_f32=True, add=add, mul=mul, mul_mode=mul_mode or 0, dact_out=dref)
    dgot = torch.empty(n, h, w, cout, device=DEV, dtype=torch.bfloat16) if want_dact else None
    ob, of = ops.conv2d_tc(x.to(DEV), LD, want_bf16=wb, want_f32=wf, add=add.to(DEV) if add is not None else None,
                           mul=mul.to(DEV) if mul is not None else None, mul_mode=mul_mode or 0, dact_out=dgot)
    torch.cuda.synchronize()
    scale = max(1.0, ref.abs().max().item())
    if wf:
        assert (of.cpu() - ref).abs().max().item() <= 2e-3 * scale
    if wb:
        assert (ob.float().cpu() - ref).abs().max().item() <= 1e-2 * scale
    if want_dact:
        assert (dgot.float().cpu() - dref).abs().max().item() <= 2e-2


@pytest.mark.parametrize("n,h,w,c", [(37, 32, 32, 64), (150, 16, 16, 128), (3, 64, 64, 64), (5, 16, 16, 256)])
def test_conv3x3_fused_channel_sums(n, h, w, c):
    """SE squeeze fused into the conv epilogue: the sums written next to the bf16 output equal ga_channel_sum of that output (same
    128-pixel slices; fp32 sums of the bf16-rounded values, only the order of the additions differs)"""
    L = _layer(c, c, 3, 1, 1, PRE_NONE, ACT_NONE, tc=True, seed=c + n)
    x = torch.randn(n, h, w, c, generator=torch.Generator().manual_seed(n)).to(torch.bfloat16).to(DEV)
    LD = _to_dev(L)
    assert ops.conv2d_tc_csum_supported(x, LD)
    sums = torch.full((n, ops.channel_sum_parts(n, h * w), c), float("nan"), device=DEV)
    r, _ = ops.conv2d_tc(x, LD, csum_out=sums)
    r_plain, _ = ops.conv2d_tc(x, LD)
    torch.cuda.synchronize()
    assert torch.equal(r, r_plain)
    ref = ops.channel_sum(r)
    assert ref.shape == sums.shape and not torch.isnan(sums).any()
    assert (sums - ref).abs().max().item() <= 2e-4 * max(1.0, ref.abs().max().item())
    # shapes the fused form does not cover say so instead of silently skipping the sums
    x8 = torch.randn(4, 8, 8, 256).to(torch.bfloat16).to(DEV)
    assert not ops.conv2d_tc_csum_supported(x8, _to_dev(_layer(256, 256, 3, 1, 1, PRE_NONE, ACT_NONE, tc=True)))


@pytest.mark.parametrize("n,h,w,cin,cout,k,stride,act", [(2, 32, 32, 64, 64, 3, 1, ACT_NONE), (3, 16, 16, 128, 256, 3, 2, ACT_NONE),
                                                        (2, 64, 64, 64, 128, 1, 2, ACT_NONE), (4, 8, 8, 512, 512, 3, 1, ACT_SILU),
                                                        (2, 128, 128, 64, 64, 3, 1, ACT_NONE), (1, 32, 32, 40, 72, 3, 1, ACT_RELU)])
def test_conv_tc_tf32_operands(n, h, w, cin, cout, k, stride, act):
    """kind::tf32 path of the tensor-core conv (fp32 activations / weights, 10-bit mantissas): against the fp32 torch convolution"""
    import torch.nn.functional as F
    from gen_adversarial_b200.fold import Folder
    g = torch.Generator().manual_seed(cin + cout + k)
    wt = torch.randn(cout, cin, k, k, generator=g, dtype=torch.float64) / math.sqrt(cin * k * k)
    b = torch.randn(cout, generator=g, dtype=torch.float64) * 0.2
    x = torch.randn(n, h, w, cin, generator=g)
    L = Folder({}, DEV, want_tc=True).conv(wt, b, stride=stride, pad=k // 2, post_act=act, simt=False, tf32=True)
    assert L.w_tf32 is not None and ops.conv2d_tc_supported(x.to(DEV), L, tf32=True)
    ob, of = ops.conv2d_tc(x.to(DEV), L, want_bf16=True, want_f32=True, tf32=True)
    y = F.conv2d(x.permute(0, 3, 1, 2).double(), wt, b, stride=stride, padding=k // 2)
    y = {ACT_NONE: y, ACT_SILU: F.silu(y), ACT_RELU: F.relu(y)}[act].permute(0, 2, 3, 1).float()
    scale = max(1.0, y.abs().max().item())
    e32 = (of.cpu() - y).abs().max().item()
    assert of.shape == y.shape and e32 <= 2e-3 * scale, (e32, scale)          # TF32 rounding: 2^-11 per operand
    assert (ob.float().cpu() - y).abs().max().item() <= 1e-2 * scale
    # operands pre-rounded to nearest TF32 (what the engines do: the MMA itself truncates): the product is then exact up to fp32 accumulation
    ops.f32_round_tf32(True)
    xr = ops.cast(x.to(DEV), torch.float32)
    ops.f32_round_tf32(False)
    assert (xr.cpu() - x).abs().max().item() <= 2 ** -11 * x.abs().max().item() and not torch.equal(xr.cpu(), x)
    assert torch.equal(ops.cast(x.to(DEV), torch.float32).cpu(), x)                 # switch off again: plain copy
    _, ofr = ops.conv2d_tc(xr, L, want_bf16=False, want_f32=True, tf32=True)
    wr = L.w_tf32.cpu().double().view(cout, k, k, cin).permute(0, 3, 1, 2)
    yr = F.conv2d(xr.cpu().permute(0, 3, 1, 2).double(), wr, b, stride=stride, padding=k // 2)
    yr = {ACT_NONE: yr, ACT_SILU: F.silu(yr), ACT_RELU: F.relu(yr)}[act].permute(0, 2, 3, 1).float()
    assert (ofr.cpu() - yr).abs().max().item() <= (2e-5 if act != ACT_SILU else 2e-3) * scale     # SiLU epilogue uses tanh.approx
    # and it is really more accurate than the bf16 operand path
    Lb = Folder({}, DEV, want_tc=True).conv(wt, b, stride=stride, pad=k // 2, post_act=act, simt=False)
    _, ofb = ops.conv2d_tc(x.to(DEV).bfloat16(), Lb, want_bf16=False, want_f32=True)
    assert e32 < 0.5 * (ofb.cpu() - y).abs().max().item()


def test_dwconv5x5():
    g = torch.Generator().manual_seed(0)
    for up in (False, True):
        x = torch.randn(2, 8, 8, 48, generator=g)
        w = torch.randn(25, 48, generator=g) * 0.2
        b = torch.randn(48, generator=g) * 0.1
        ref = emu_ops.dwconv5x5(x, w, b, ACT_SILU, up, torch.float32)
        got = ops.dwconv5x5(x.to(DEV), w.to(DEV), b.to(DEV), ACT_SILU, up, torch.float32)
        assert (got.cpu() - ref).abs().max().item() <= 1e-5
        gotb = ops.dwconv5x5(x.to(DEV).bfloat16(), w.to(DEV), b.to(DEV), ACT_SILU, up, torch.bfloat16)
        assert (gotb.float().cpu() - ref).abs().max().item() <= 3e-2


@pytest.mark.parametrize("n,h,w,c,up", [(3, 16, 16, 128, False), (2, 32, 32, 64, False), (2, 12, 16, 96, False), (5, 8, 8, 192, False),
                                        (2, 64, 64, 32, False), (2, 24, 40, 72, False), (3, 8, 8, 128, True), (2, 16, 16, 96, True),
                                        (1, 4, 4, 64, True), (300, 8, 8, 64, False)])
def test_dwconv5x5_bf16_tma_pipeline(n, h, w, c, up):
    """persistent TMA-pipelined bf16 kernel (dwconv_tma.cu): plain forward, taping forward (dact) and backward (flipped taps x mul);
    whole / partial tiles, partial channel blocks, nearest-x2 input, more tiles than resident CTAs"""
    g = torch.Generator().manual_seed(n * 1000 + c)
    x = torch.randn(n, h, w, c, generator=g).bfloat16()
    wt = torch.randn(25, c, generator=g) * 0.2
    b = torch.randn(c, generator=g) * 0.1
    s = 2 if up else 1
    m = torch.randn(n, h * s, w * s, c, generator=g).bfloat16()
    xd, wd, bd, md = x.to(DEV), wt.to(DEV), b.to(DEV), m.to(DEV)
    ref = emu_ops.dwconv5x5(x.float(), wt, b, ACT_SILU, up, torch.float32)
    got = ops.dwconv5x5(xd, wd, bd, ACT_SILU, up, torch.bfloat16)
    scale = max(1.0, ref.abs().max().item())
    assert got.shape == ref.shape and (got.float().cpu() - ref).abs().max().item() <= 1e-2 * scale
    ry, rd = emu_ops.dwconv5x5(x.float(), wt, b, ACT_SILU, up, torch.float32, want_dact=True)
    gy, gd = ops.dwconv5x5(xd, wd, bd, ACT_SILU, up, torch.bfloat16, want_dact=True)
    assert torch.equal(gy, got)                                              # the taping variant computes the same values
    assert (gd.float().cpu() - rd).abs().max().item() <= 1e-2
    rm = emu_ops.dwconv5x5(x.float(), wt, None, ACT_NONE, up, torch.float32, mul=m.float())
    gm = ops.dwconv5x5(xd, wd, None, ACT_NONE, up, torch.bfloat16, mul=md)
    assert (gm.float().cpu() - rm).abs().max().item() <= 1e-2 * max(1.0, rm.abs().max().item())
    assert torch.equal(gm, ops.dwconv5x5(xd, wd, None, ACT_NONE, up, torch.bfloat16, mul=md))     # deterministic


@pytest.mark.parametrize("c,hw", [(32, 64), (64, 32), (256, 8), (512, 4), (24, 8), (8, 16)])
def test_channel_sum_and_se_residual(c, hw):
    g = torch.Generator().manual_seed(c)
    n = 3
    r = torch.randn(n, hw, hw, c, generator=g)
    skip = torch.randn(n, hw, hw, c, generator=g)
    hid = max(c // 16, 4)
    se = (torch.randn(hid, c, generator=g) * 0.3, torch.randn(hid, generator=g) * 0.1,
          torch.randn(c, hid, generator=g) * 0.3, torch.randn(c, generator=g) * 0.1)
    aff = (torch.rand(c, generator=g) + 0.5, torch.randn(c, generator=g) * 0.1)
    sums_ref = emu_ops.channel_sum(r)
    sums = ops.channel_sum(r.to(DEV))
    assert (sums.sum(dim=1).cpu() - sums_ref[:, 0]).abs().max().item() <= 1e-3 * max(1.0, sums_ref.abs().max().item())
    assert torch.equal(sums, ops.channel_sum(r.to(DEV)))            # two-stage reduction: bit-reproducible
    ref = emu_ops.se_residual(r, sums_ref, se, 0.1, skip, torch.float32, want_out2=True, act_affine=aff, want_gate=True)
    got = ops.se_residual(r.to(DEV), sums, tuple(t.to(DEV) for t in se), 0.1, skip.to(DEV), torch.float32, want_out2=True,
                          act_affine=tuple(t.to(DEV) for t in aff), want_gate=True)
    assert (got[0].cpu() - ref[0]).abs().max().item() <= 1e-5
    assert (got[1].float().cpu() - ref[0]).abs().max().item() <= 3e-2
    assert (got[2].float().cpu() - ref[2].float()).abs().max().item() <= 3e-2
    assert (got[3].cpu() - ref[3]).abs().max().item() <= 1e-5
    # ELU copy without an affine (pre-activated input of the decoder sampler / logits head)
    from gen_adversarial_b200._lib import ACT_ELU as _ELU
    ref_e = emu_ops.se_residual(r, sums_ref, se, 0.1, skip, torch.float32, want_out2=True, act_plain=True, act_op=_ELU)
    got_e = ops.se_residual(r.to(DEV), sums, tuple(t.to(DEV) for t in se), 0.1, skip.to(DEV), torch.float32, want_out2=True,
                            act_plain=True, act_op=_ELU)
    assert (got_e[0].cpu() - ref_e[0]).abs().max().item() <= 1e-5
    assert got_e[2].dtype == torch.bfloat16 and (got_e[2].float().cpu() - ref_e[2].float()).abs().max().item() <= 3e-2


def test_latent_mix_explicit_noise():
    g = torch.Generator().manual_seed(0)
    n, h, z, zc = 3, 8, 20, 24
    q = torch.randn(n, h, h, z, generator=g) * 2
    p = torch.randn(n, h, h, 2 * z, generator=g) * 2
    eps = torch.randn(n, z, h, h, generator=g)
    alpha = torch.tensor([0.37])
    for pp in (None, p):
        ref = emu_ops.latent_mix(q, pp, eps, 0, 0, 0, alpha, 0.6, z, zc, torch.float32)
        got = ops.latent_mix(q.to(DEV), pp.to(DEV) if pp is not None else None, eps.to(DEV), 0, 0, 0, alpha.to(DEV), 0.6, z, zc,
                             torch.float32)
        assert (got.cpu() - ref).abs().max().item() <= 1e-5
        assert got[..., z:].abs().max().item() == 0.0


def test_latent_mix_philox_statistics_and_shard_independence():
    n, h, z, zc = 64, 16, 20, 24
    q = torch.zeros(n, h, h, z, device=DEV)
    alpha = torch.ones(1, device=DEV)
    full = ops.latent_mix(q, None, None, 1234, 3, 0, alpha, 1.0, z, zc, torch.float32)[..., :z]
    assert abs(full.mean().item()) < 0.02 and abs(full.std().item() - 1.0) < 0.02
    half = ops.latent_mix(q[32:], None, None, 1234, 3, 32, alpha, 1.0, z, zc, torch.float32)[..., :z]
    assert torch.equal(half, full[32:])       # keyed by the GLOBAL sample index


def test_latent_mix_philox_backward_regenerates_forward_noise():
    """Philox mode: the backward kernel re-draws eps from (seed, level, global sample, channel, pixel); it must be the forward's draw.
    eps is recovered from the forward output at alpha = 1 with the prior N(0,1) (z = eps * T) and fed explicitly to a second backward."""
    n, h, z, zc, seed, lvl = 3, 8, 20, 24, 99, 5
    g = torch.Generator().manual_seed(3)
    q = (torch.randn(n, h, h, z, generator=g)).to(DEV)
    p = (torch.randn(n, h, h, 2 * z, generator=g) * 0.5).to(DEV)
    gz = torch.randn(n, h, h, zc, generator=g).to(DEV)
    one, a = torch.ones(1, device=DEV), torch.full((1,), 0.4, device=DEV)
    eps_nhwc = ops.latent_mix(torch.zeros_like(q), None, None, seed, lvl, 7, one, 1.0, z, zc, torch.float32)[..., :z]
    eps = eps_nhwc.permute(0, 3, 1, 2).contiguous()
    z_phi = ops.latent_mix(q, p, None, seed, lvl, 7, a, 0.6, z, zc, torch.float32)
    z_exp = ops.latent_mix(q, p, eps, 0, lvl, 7, a, 0.6, z, zc, torch.float32)
    assert (z_phi - z_exp).abs().max().item() <= 1e-5
    gq1, gp1 = ops.latent_mix_bwd(gz, q, p, None, seed, lvl, 7, a, 0.6, z, zc)
    gq2, gp2 = ops.latent_mix_bwd(gz, q, p, eps, 0, lvl, 7, a, 0.6, z, zc)
    assert (gq1 - gq2).abs().max().item() <= 1e-6 and (gp1 - gp2).abs().max().item() <= 1e-5 * max(1.0, gp2.abs().max().item())
    # bf16 output stays within bf16 rounding of the fp32 result
    z_bf = ops.latent_mix(q, p, eps, 0, lvl, 7, a, 0.6, z, zc, torch.bfloat16)
    assert (z_bf.float() - z_exp).abs().max().item() <= 2e-2 * max(1.0, z_exp.abs().max().item())


def test_discmix_mean():
    g = torch.Generator().manual_seed(0)
    logits = torch.randn(3, 16, 16, 100, generator=g) * 1.5
    ref_p, ref_c = emu_ops.discmix_mean(logits, 10, torch.float32)
    got_p, got_c = ops.discmix_mean(logits.to(DEV), 10, torch.float32)
    assert (got_p.cpu() - ref_p).abs().max().item() <= 1e-5
    assert (got_c.cpu() - ref_c).abs().max().item() <= 1e-5


@pytest.mark.parametrize("hw,eps,blur", [(64, 2.0, False), (64, 0.0, True), (32, 1.0, True), (128, 4.0, True), (40, 1.0, True)])
def test_preprocess_matches_reference_semantics(hw, eps, blur):
    g = torch.Generator().manual_seed(hw)
    x = torch.rand(3, 3, hw, hw, generator=g)
    noise = torch.randn(3, 3, hw, hw, generator=g)
    ref = nvae_ref.preprocess(x, noise, eps, blur) if hw != 40 else None
    if ref is None:                                        # k = int(2**(sqrt(40)//2)-1) = 7
        ref = nvae_ref.preprocess(x, noise, eps, blur)
    got, pre = ops.preprocess(x.to(DEV), noise.to(DEV), eps, blur, torch.float32, normalize=True, save_pre=True)
    got = got.permute(0, 3, 1, 2).cpu()
    assert (got - (ref - 0.5) / 0.5).abs().max().item() <= 2e-5
    assert (pre.cpu().clamp(0, 1) - ref).abs().max().item() <= 1e-5


def test_preprocess_philox_noise_has_requested_l2_norm():
    x = torch.full((4, 3, 64, 64), 0.5, device=DEV)
    out, _ = ops.preprocess(x, None, 2.0, False, torch.float32, seed=99, sample0=0, normalize=False)
    d = (out.permute(0, 3, 1, 2) - x).flatten(1).norm(dim=1)
    assert (d - 2.0).abs().max().item() <= 1e-3
    out2, _ = ops.preprocess(x[2:], None, 2.0, False, torch.float32, seed=99, sample0=2, normalize=False)
    assert torch.equal(out2, out[2:])                      # Philox keyed by the global sample index; no atomics


def test_resampling_and_pool_and_cast():
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 8, 8, 16, generator=g)
    assert (ops.upsample_nearest2x(x.to(DEV)).cpu() - emu_ops.upsample_nearest2x(x)).abs().max().item() == 0
    assert (ops.upsample_bilinear2x(x.to(DEV)).cpu() - emu_ops.upsample_bilinear2x(x)).abs().max().item() <= 1e-5
    assert (ops.maxpool2x2(x.to(DEV)).cpu() - emu_ops.maxpool2x2(x)).abs().max().item() == 0
    sc, sh = torch.rand(16, generator=g) + 0.5, torch.randn(16, generator=g)
    got = ops.affine_act(x.to(DEV), sc.to(DEV), sh.to(DEV), ACT_SILU, torch.float32)
    assert (got.cpu() - emu_ops.affine_act(x, sc, sh, ACT_SILU, torch.float32)).abs().max().item() <= 1e-5
    xn = torch.randn(2, 3, 8, 8, generator=g)
    assert (ops.nchw_to_nhwc(xn.to(DEV), torch.float32, 2.0, -1.0).cpu() - emu_ops.nchw_to_nhwc(xn, torch.float32, 2.0, -1.0)).abs().max().item() <= 1e-6


def test_pgd_step_bit_exact():
    g = torch.Generator().manual_seed(0)
    x = torch.rand(5, 3, 64, 64, generator=g)
    xa = (x + (torch.rand(x.shape, generator=g) - 0.5) * 0.05).clamp(0, 1)
    gr = torch.randn(x.shape, generator=g)
    gr[0, 0, 0, :8] = 0.0
    ref = nvae_ref.pgd_linf_step(xa, gr, x, 2 / 255, 8 / 255)
    got = ops.pgd_linf_step_(xa.clone().to(DEV), gr.to(DEV), x.to(DEV), 2 / 255, 8 / 255)
    assert torch.equal(got.cpu(), ref)


def test_softmax_xent():
    g = torch.Generator().manual_seed(0)
    logits = torch.randn(37, 100, generator=g) * 3
    y = torch.randint(0, 100, (37,), generator=g)
    counter = torch.zeros(1, dtype=torch.int64, device=DEV)
    loss, dl, pred = ops.softmax_xent(logits.to(DEV), y.to(DEV), True, counter)
    lr = logits.clone().requires_grad_(True)
    ref = torch.nn.functional.cross_entropy(lr, y, reduction="none")
    ref.mean().backward()
    assert (loss.cpu() - ref.detach()).abs().max().item() <= 1e-5
    assert (dl.cpu() - lr.grad).abs().max().item() <= 1e-6
    assert torch.equal(pred.cpu().long(), logits.argmax(1))
    assert counter.item() == (logits.argmax(1) == y).sum().item()
