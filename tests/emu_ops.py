"""TEST INFRASTRUCTURE ONLY: a torch-CPU emulation of the C-ABI op wrappers in gen_adversarial_b200/ops.py.

It lets the `-m "not gpu"` suite exercise the HOST logic (weight folding, op ordering, tensor plumbing of
nvae_engine.py / vgg_engine.py) against the oracle without a GPU, by monkeypatching the `ops` module inside a
test.  It is never imported by the product; on a GPU the same host code drives the real kernels and is checked
against the same oracle by the `-m gpu` tests.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from gen_adversarial_b200 import ops as real_ops
from gen_adversarial_b200._lib import PRE_NONE, PRE_ELU, PRE_SILU, PRE_AFFINE_SILU, PRE_AFFINE, ACT_NONE, ACT_SILU, ACT_ELU, \
    ACT_RELU, ACT_LRELU_SQRT2, ACT_PRELU, MUL_VALUE, MUL_RELU_MASK, MUL_ELU_FROM_Y

ConvLayer = real_ops.ConvLayer
gaussian_taps = real_ops.gaussian_taps
conv_out_hw = real_ops.conv_out_hw
_launches = [0]


def _act(v, act):
    if act == ACT_SILU:
        return F.silu(v)
    if act == ACT_ELU:
        return F.elu(v)
    if act == ACT_RELU:
        return F.relu(v)
    if act == ACT_LRELU_SQRT2:
        return F.leaky_relu(v, 0.2) * (2 ** 0.5)
    return v


def _act_grad(v, act):
    if act == ACT_SILU:
        sg = torch.sigmoid(v)
        return sg * (1 + v * (1 - sg))
    if act == ACT_ELU:
        return torch.where(v > 0, torch.ones_like(v), torch.exp(v))
    if act == ACT_RELU:
        return (v > 0).to(v.dtype)
    return torch.ones_like(v)


def _mul_factor(m, mode):
    m = m.float()
    if mode == MUL_RELU_MASK:
        return (m > 0).float()
    if mode == MUL_ELU_FROM_Y:
        return torch.where(m > 0, torch.ones_like(m), m + 1)
    return m


def _pre(x, L):
    if L.pre_op == PRE_ELU:
        return F.elu(x)
    if L.pre_op == PRE_SILU:
        return F.silu(x)
    if L.pre_op == PRE_AFFINE_SILU:
        return F.silu(x * L.pre_scale + L.pre_shift)
    if L.pre_op == PRE_AFFINE:
        return x * L.pre_scale + L.pre_shift
    return x


def _epilogue(v_nchw, L, add, mul, mul_mode):
    """out = (act(v) + add) * f(mul), or act(v + add) with act_after_add; PReLU uses the per-channel slopes"""
    def act(t):
        if L.post_act == ACT_PRELU:
            return torch.where(t > 0, t, t * L.act_slope.float().view(1, -1, 1, 1))
        return _act(t, L.post_act)
    y = v_nchw if L.act_after_add else act(v_nchw)
    y = y.permute(0, 2, 3, 1)
    if add is not None:
        y = y + add.float()
    if L.act_after_add:
        y = act(y.permute(0, 3, 1, 2)).permute(0, 2, 3, 1)
    if mul is not None:
        y = y * _mul_factor(mul, mul_mode)
    return y.contiguous()


def _nchw(x):
    return x.float().permute(0, 3, 1, 2)


def _nhwc(x, dtype):
    return x.permute(0, 2, 3, 1).contiguous().to(dtype)


def conv2d_simt(x, L, out_dtype, add=None, out_hw=None, mul=None, mul_mode=0, want_dact=False, out=None):
    _launches[0] += 1
    xin = _pre(x.float(), L)
    w = L.w_simt.float().view(L.kh, L.kw, L.cin, L.cout).permute(3, 2, 0, 1)
    xn = xin.permute(0, 3, 1, 2)
    if L.up > 1:
        n, c, h, wd = xn.shape
        z = torch.zeros((n, c, (h - 1) * L.up + 1, (wd - 1) * L.up + 1), dtype=xn.dtype)
        z[:, :, ::L.up, ::L.up] = xn
        xn = z
        if out_hw is not None:     # output_padding of a transposed conv: pad bottom/right with zeros
            ho, wo = conv_out_hw(L, h, wd)
            xn = F.pad(xn, (0, out_hw[1] - wo, 0, out_hw[0] - ho))
    v = F.conv2d(xn, w, L.bias.float() if L.bias is not None else None, stride=L.stride, padding=L.pad)
    y = _epilogue(v, L, add, mul, mul_mode).to(out_dtype)
    if out is not None:
        out.copy_(y)
        y = out
    if want_dact:
        return y, _act_grad(v, L.post_act).permute(0, 2, 3, 1).contiguous().to(out_dtype)
    return y


def conv2d_tc_supported(x, L, x2=None, tf32=False):
    if tf32:
        if L.w_tf32 is None or x.dtype != torch.float32 or x2 is not None:
            return False
    elif L.w_tc is None or x.dtype != torch.bfloat16:
        return False
    if L.stride not in (1, 2) or L.up != 1 or (x2 is not None and L.stride != 1):
        return False
    n, h, w, c = x.shape
    if c % 8:
        return False
    h, w = conv_out_hw(L, h, w)
    if w >= 128:
        return w % 128 == 0
    if 128 % w:
        return False
    rows = 128 // w
    return True if h >= rows else (rows % h == 0)


def conv2d_tc_csum_supported(x, L):
    return False          # the emulation keeps the separate channel_sum op (the fused form is a kernel-level feature)


def conv2d_tc(x, L, want_bf16=True, want_f32=False, add=None, x2=None, mul=None, mul_mode=0, dact_out=None, out_bf16=None,
              out_f32=None, tf32=False, csum_out=None):
    _launches[0] += 1
    assert x.dtype == (torch.float32 if tf32 else torch.bfloat16)
    k1 = L.kh * L.kw * L.cin
    w = L.w_tf32.float() if tf32 else L.w_tc.float()
    w1 = w[:, :k1].view(L.cout, L.kh, L.kw, L.cin).permute(0, 3, 1, 2)
    y = F.conv2d(_nchw(x), w1, None, stride=L.stride, padding=L.pad)
    if x2 is not None:
        assert x2.dtype == torch.bfloat16 and w.shape[1] == k1 + x2.shape[3]
        y = y + F.conv2d(_nchw(x2), w[:, k1:].view(L.cout, -1, 1, 1))
    else:
        assert w.shape[1] == k1
    if L.bias is not None:
        y = y + L.bias.float().view(1, -1, 1, 1)
    if dact_out is not None:
        dact_out.copy_(_act_grad(y, L.post_act).permute(0, 2, 3, 1))
    y = _epilogue(y, L, add, mul, mul_mode)
    ob = y.to(torch.bfloat16) if want_bf16 else None
    of = y if want_f32 else None
    if ob is not None and out_bf16 is not None:
        out_bf16.copy_(ob); ob = out_bf16
    if of is not None and out_f32 is not None:
        out_f32.copy_(of); of = out_f32
    return ob, of


def dwconv5x5(x, weight, bias, act, up, out_dtype, mul=None, want_dact=False):
    _launches[0] += 1
    xn = _nchw(x)
    if up:
        xn = F.interpolate(xn, scale_factor=2, mode="nearest")
    c = xn.shape[1]
    w = weight.float().view(5, 5, c).permute(2, 0, 1).unsqueeze(1)
    v = F.conv2d(xn, w, bias.float() if bias is not None else None, padding=2, groups=c)
    y = _act(v, act)
    if mul is not None:
        y = y * _nchw(mul)
    if want_dact:
        return _nhwc(y, out_dtype), _nhwc(_act_grad(v, act), out_dtype)
    return _nhwc(y, out_dtype)


def channel_sum(r):
    _launches[0] += 1
    return r.float().sum(dim=(1, 2)).unsqueeze(1)      # [n, parts=1, c]


def mbconv_fused_supported(x, e, p):
    n, h, w, c = x.shape
    return x.dtype == torch.bfloat16 and e.w_tc is not None and p.w_tc is not None and h == w and (w, c) in ((8, 256), (16, 128), (32, 64))


def dw_weights_chunked(dw_w):
    return real_ops.dw_weights_chunked(dw_w)


def mbconv_fused(x, e, dw_w_chunked, dw_b, p, want_sums=False, want_tape=False):
    if want_tape:          # the three kernels the fused taping cell replaces, with their tapes
        dw_w = dw_w_chunked.permute(1, 0, 2).reshape(25, -1)
        dact_e = torch.empty(x.shape[:3] + (e.cout,), dtype=torch.bfloat16)
        h1, _ = conv2d_tc(x, e, dact_out=dact_e)
        h2, dact_dw = dwconv5x5(h1, dw_w, dw_b, ACT_SILU, False, torch.bfloat16, want_dact=True)
        r, _ = conv2d_tc(h2, p)
        _launches[0] -= 2
        s_ = channel_sum(r) if want_sums else None
        if want_sums:
            _launches[0] -= 1
        return r, s_, dact_e, dact_dw
    if want_sums:
        r = mbconv_fused(x, e, dw_w_chunked, dw_b, p)
        s = channel_sum(r)
        _launches[0] -= 1
        return r, s
    _launches[0] += 1
    dw_w = dw_w_chunked.permute(1, 0, 2).reshape(25, -1)
    h1, _ = conv2d_tc(x, e)
    _launches[0] -= 1
    h2 = dwconv5x5(h1, dw_w, dw_b, ACT_SILU, False, torch.bfloat16)
    _launches[0] -= 1
    r, _ = conv2d_tc(h2, p)
    _launches[0] -= 1
    return r


def mbconv_fused_bwd(g, p_d, dw_wT_chunked, dact_dw, dact_e, e_d, add=None):
    """the three kernels the fused backward cell replaces"""
    dw_wT = dw_wT_chunked.permute(1, 0, 2).reshape(25, -1)
    g_v2, _ = conv2d_tc(g, p_d, mul=dact_dw)
    g_v1 = dwconv5x5(g_v2, dw_wT, None, ACT_NONE, False, torch.bfloat16, mul=dact_e)
    _, o = conv2d_tc(g_v1, e_d, want_bf16=False, want_f32=True, add=add)
    _launches[0] -= 2
    return o


def se_residual(r, sums, se, res_scale, skip, out_dtype=torch.float32, want_out2=False, out2_dtype=torch.bfloat16,
                act_affine=None, act_dtype=torch.bfloat16, want_gate=False, act_op=ACT_SILU, act_plain=False):
    _launches[0] += 1
    w1, b1, w2, b2 = se
    mean = sums.sum(dim=1) / (r.shape[1] * r.shape[2])
    gate = torch.sigmoid(F.linear(F.relu(F.linear(mean, w1, b1)), w2, b2))
    out = skip.float() + res_scale * gate[:, None, None, :] * r.float()
    out2 = out.to(out2_dtype) if want_out2 else None
    act = None
    if act_affine is not None or act_plain:
        act = out * act_affine[0] + act_affine[1] if act_affine is not None else out
        act = (F.silu(act) if act_op == ACT_SILU else (F.elu(act) if act_op == ACT_ELU else act)).to(act_dtype)
    return out.to(out_dtype), out2, act, (gate if want_gate else None)


# ------------------------------------------------------------------------------------------------ encoder-side ops
def add_layernorm(x, y, gamma, beta, out_dtype, eps=1e-5, out2_dtype=None):
    _launches[0] += 1
    t = x.float() + (y.float() if y is not None else 0.0)
    o = F.layer_norm(t, (t.shape[-1],), gamma, beta, eps)
    return (o.to(out_dtype), o.to(out2_dtype)) if out2_dtype is not None else o.to(out_dtype)


def attention(q, q_off, k, k_off, v, v_off, heads, dh, out_dtype):
    _launches[0] += 2
    b, nq = q.shape[0], q.shape[1] * q.shape[2]
    s = k.shape[1] * k.shape[2]
    qf = q.float().reshape(b, nq, -1)[..., q_off:q_off + heads * dh].reshape(b, nq, heads, dh).permute(0, 2, 1, 3)
    kf = k.float().reshape(b, s, -1)[..., k_off:k_off + heads * dh].reshape(b, s, heads, dh).permute(0, 2, 1, 3)
    vf = v.float().reshape(b, s, -1)[..., v_off:v_off + heads * dh].reshape(b, s, heads, dh).permute(0, 2, 1, 3)
    p = torch.softmax(qf @ kf.transpose(-1, -2) / (dh ** 0.5), dim=-1)
    o = (p @ vf).permute(0, 2, 1, 3).reshape(b, nq, 1, heads * dh)
    return o.contiguous().to(out_dtype)


def codes_assemble(heads, heads_lb, use_w0, latent_avg, b, l, d):
    _launches[0] += 1
    h = heads.float().reshape(l, b, d).permute(1, 0, 2) if heads_lb else heads.float().reshape(b, l, d)
    out = h.clone()
    if use_w0:
        out[:, 1:] += h[:, :1]
    if latent_avg is not None:
        out = out + latent_avg.float().reshape(1, l, d)
    return out.contiguous()


def resize_bilinear(x, full_h, out_w, crop_y0, crop_h, out_dtype=None):
    _launches[0] += 1
    y = F.interpolate(_nchw(x), size=(full_h, out_w), mode="bilinear", align_corners=False)
    return _nhwc(y[:, :, crop_y0:crop_y0 + crop_h], out_dtype or x.dtype)


def image_pool_out(img, k1, k2=1, mask_rows=0, denorm=(0.5, 0.5), cls_dtype=None, want_purified=True):
    _launches[0] += 1
    y = F.avg_pool2d(_nchw(img)[:, :3], k1)
    if k2 == 2:
        if mask_rows:
            y = y.clone()
            y[:, :, :mask_rows] = -1.0
            y[:, :, -mask_rows:] = -1.0
        y = F.avg_pool2d(y, 2)
    pur = (y * denorm[0] + denorm[1]).contiguous() if want_purified else None
    cls = _nhwc(y, cls_dtype) if cls_dtype is not None else None
    return pur, cls


def philox_codes(seed, sample0, std, l, b, d, device):
    raise RuntimeError("the emulation only supports explicit noise")


def latent_mix(q, p, eps_nchw, seed, level, sample0, alpha_dev, temperature, zdim, zc, out_dtype):
    _launches[0] += 1
    assert eps_nchw is not None, "the emulation only supports explicit noise"
    a = float(alpha_dev[0])
    sc = lambda t: 5.0 * torch.tanh(t / 5.0)
    mu_q = q.float()[..., :zdim]
    e = eps_nchw.float().permute(0, 2, 3, 1)
    if p is None:
        z = (1 - a) * sc(mu_q) + a * (e * temperature)
    else:
        mu_p, ls_p = p.float()[..., :zdim], p.float()[..., zdim:]
        z = (1 - a) * sc(mu_p + mu_q) + a * (sc(mu_p) + e * (temperature * torch.exp(sc(ls_p))))
    out = torch.zeros(q.shape[:3] + (zc,), dtype=torch.float32)
    out[..., :zdim] = z
    return out.to(out_dtype)


def discmix_mean(logits, n_mix, cls_dtype=None):
    _launches[0] += 1
    from oracle.nvae_ref import disc_mix_logistic_mean
    rec = disc_mix_logistic_mean(_nchw(logits), n_mix)
    purified = rec * 0.5 + 0.5
    cls = _nhwc((purified - 0.5) / 0.5, cls_dtype) if cls_dtype is not None else None
    return purified.contiguous(), cls


def upsample_nearest2x(x, out_dtype=None):
    _launches[0] += 1
    return _nhwc(F.interpolate(_nchw(x), scale_factor=2, mode="nearest"), out_dtype or x.dtype)


def upsample_bilinear2x(x, out_dtype=None):
    _launches[0] += 1
    return _nhwc(F.interpolate(_nchw(x), scale_factor=2, mode="bilinear", align_corners=True), out_dtype or x.dtype)


def maxpool2x2(x, out_dtype=None):
    _launches[0] += 1
    return _nhwc(F.max_pool2d(_nchw(x), 2), out_dtype or x.dtype)


def subsample2x(x, out_dtype=None):
    _launches[0] += 1
    return x[:, ::2, ::2, :].contiguous().to(out_dtype or x.dtype)


def maxpool3x3s2(x, out_dtype=None):
    _launches[0] += 1
    return _nhwc(F.max_pool2d(_nchw(x), 3, 2, 1), out_dtype or x.dtype)


def global_avgpool(x, out_dtype=None):
    _launches[0] += 1
    return x.float().mean(dim=(1, 2), keepdim=True).to(out_dtype or x.dtype)


def affine_act(x, scale, shift, act, out_dtype):
    _launches[0] += 1
    v = x.float()
    if scale is not None:
        v = v * scale + shift
    return _act(v, act).to(out_dtype)


def cast(x, out_dtype):
    return affine_act(x, None, None, ACT_NONE, out_dtype)


def nchw_to_nhwc(x_nchw, out_dtype, scale=1.0, shift=0.0):
    _launches[0] += 1
    return _nhwc(x_nchw.float() * scale + shift, out_dtype)


def preprocess(x_nchw, noise_nchw, eps, blur, out_dtype, seed=0, sample0=0, normalize=True, save_pre=False, taps_cache=None):
    _launches[0] += 1
    from oracle import nvae_ref
    x = x_nchw.float()
    if blur:
        x = nvae_ref.gaussian_blur(x)
    if eps != 0.0:
        assert noise_nchw is not None, "the emulation only supports explicit noise"
        nrm = noise_nchw.reshape(noise_nchw.shape[0], -1).norm(dim=1).view(-1, 1, 1, 1)
        x = x + noise_nchw * (eps / nrm)
    pre = x.clone() if save_pre else None
    x = x.clamp(0, 1)
    if normalize:
        x = (x - 0.5) * 2.0
    return _nhwc(x, out_dtype), pre


# ------------------------------------------------------------------------------------------------ StyleGAN2 generator ops
def pixelnorm(x, out_dtype):
    _launches[0] += 1
    y = x * torch.rsqrt(torch.mean(x ** 2, dim=1, keepdim=True) + 1e-8)
    return y.view(x.shape[0], 1, 1, x.shape[1]).to(out_dtype)


def style_demod(s, wsq):
    _launches[0] += 1
    return torch.rsqrt((s ** 2) @ wsq.t() + 1e-8)


def channel_scale(x, s, out_dtype):
    _launches[0] += 1
    return (x.float() * s[:, None, None, :]).to(out_dtype)


def styled_bias_act(y, phases, demod, noise_hw, noise_w, bias, act, skip, out_dtype, scale_a=None, scale_b=None, want_out=True,
                    skip_up_kernel=None):
    _launches[0] += 1
    if skip is not None and skip_up_kernel is not None:
        skip = upfirdn2d(skip, skip_up_kernel, up=2, down=1, pad=(2, 1), out_dtype=torch.float32)
        _launches[0] -= 1
    v = y.float()
    if phases:
        n, hh, wh, c4 = v.shape
        ph = v.view(n, hh, wh, 2, 2, c4 // 4)                           # [n, h, w, py, px, c]
        v = ph.permute(0, 1, 3, 2, 4, 5).reshape(n, hh * 2, wh * 2, c4 // 4)
    if demod is not None:
        v = v * demod[:, None, None, :]
    if noise_hw is not None:
        v = v + noise_w * noise_hw.view(1, v.shape[1], v.shape[2], 1)
    if bias is not None:
        v = v + bias
    v = _act(v, act)
    if skip is not None:
        v = v + skip.float()
    out = None
    if want_out:
        out = (v * scale_a[:, None, None, :] if scale_a is not None else v).contiguous().to(out_dtype)
    if scale_b is not None:
        return out, (v * scale_b[:, None, None, :]).contiguous().to(out_dtype)
    return out


def torgb_fused(x, L, bias, skip=None, skip_up_kernel=None):
    _launches[0] += 1
    y = torch.einsum("nhwc,jc->nhwj", x.float(), L.w_tc.float())
    if bias is not None:
        y = y + bias
    if skip is not None:
        y = y + upfirdn2d(skip, skip_up_kernel, up=2, down=1, pad=(2, 1), out_dtype=torch.float32)
        _launches[0] -= 1
    return y.contiguous()


def upfirdn2d(x, kernel, up=1, down=1, pad=(0, 0), out_dtype=None):
    _launches[0] += 1
    from oracle.ref_import import upfirdn2d_cpu
    return _nhwc(upfirdn2d_cpu(_nchw(x), kernel.float(), up, down, pad), out_dtype or x.dtype)


def avgpool_to_nchw(x, k, out_c):
    _launches[0] += 1
    return F.avg_pool2d(_nchw(x)[:, :out_c], k).contiguous()


def latent_lerp(codes, styles, alphas_dev):
    _launches[0] += 1
    a = alphas_dev.view(1, -1, 1)
    return (1 - a) * codes + a * styles


# ------------------------------------------------------------------------------------------------ backward ops
def affine_act_bwd(g, x, scale, shift, act, out_dtype, add=None):
    _launches[0] += 1
    v = x.float()
    sc = 1.0
    if scale is not None:
        v = v * scale + shift
        sc = scale
    out = g.float() * _act_grad(v, act) * sc
    if add is not None:
        out = out + add.float()
    return out.to(out_dtype)


def add(a, b, out_dtype):
    _launches[0] += 1
    return (a.float() + b.float()).to(out_dtype)


def se_residual_bwd(g_out, r, sums, se, res_scale, out_dtype):
    _launches[0] += 2
    w1, b1, w2, b2 = se
    rr = r.float().detach().requires_grad_(True)
    with torch.enable_grad():
        gate = torch.sigmoid(F.linear(F.relu(F.linear(rr.mean(dim=(1, 2)), w1, b1)), w2, b2))
        out = res_scale * gate[:, None, None, :] * rr
        g_r, = torch.autograd.grad(out, [rr], g_out.float())
    return g_r.to(out_dtype)


def sumpool2x2(x, out_dtype, mul=None):
    _launches[0] += 1
    y = F.avg_pool2d(_nchw(x), 2) * 4.0
    y = y.permute(0, 2, 3, 1)
    if mul is not None:
        y = y * mul.float()
    return y.contiguous().to(out_dtype)


def f32_round_tf32(on):
    pass


def depth_to_space2(x):
    _launches[0] += 1
    n, h, w, c4 = x.shape
    c = c4 // 4
    return x.float().view(n, h, w, 2, 2, c).permute(0, 1, 3, 2, 4, 5).reshape(n, 2 * h, 2 * w, c).contiguous()


def upsample_bilinear2x_bwd(g_out, out_dtype):
    _launches[0] += 1
    n, h, w, c = g_out.shape
    x = torch.zeros((n, c, h // 2, w // 2), requires_grad=True)
    with torch.enable_grad():
        y = F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=True)
        g, = torch.autograd.grad(y, [x], _nchw(g_out))
    return _nhwc(g, out_dtype)


def maxpool2x2_bwd(x_in, g_out, relu, out_dtype):
    _launches[0] += 1
    x = _nchw(x_in).detach().requires_grad_(True)
    with torch.enable_grad():
        y = F.max_pool2d(F.relu(x) if relu else x, 2)
        g, = torch.autograd.grad(y, [x], _nchw(g_out))
    return _nhwc(g, out_dtype)


def latent_mix_bwd(g_z, q, p, eps_nchw, seed, level, sample0, alpha_dev, temperature, zdim, zc):
    _launches[0] += 1
    a = float(alpha_dev[0])
    sc = lambda t: 5.0 * torch.tanh(t / 5.0)
    qq = q.float()[..., :zdim].detach().requires_grad_(True)
    g = g_z.float()[..., :zdim]
    with torch.enable_grad():
        if p is None:
            z = (1 - a) * sc(qq)
            g_q, = torch.autograd.grad(z, [qq], g)
            g_p = None
        else:
            pp = p.float().detach().requires_grad_(True)
            e = eps_nchw.float().permute(0, 2, 3, 1)
            mu_p, ls_p = pp[..., :zdim], pp[..., zdim:]
            z = (1 - a) * sc(mu_p + qq) + a * (sc(mu_p) + e * (temperature * torch.exp(sc(ls_p))))
            g_q, g_p = torch.autograd.grad(z, [qq, pp], g)
    out = torch.zeros(q.shape[:3] + (zc,), dtype=torch.float32)
    out[..., :zdim] = g_q
    return out, g_p


def discmix_mean_bwd(logits, n_mix, g_purified_nchw, g_cls, pad_to=0):
    _launches[0] += 1
    from oracle.nvae_ref import disc_mix_logistic_mean
    l = _nchw(logits).detach().requires_grad_(True)
    with torch.enable_grad():
        v = disc_mix_logistic_mean(l, n_mix)
        gv = torch.zeros_like(v)
        if g_purified_nchw is not None:
            gv = gv + 0.5 * g_purified_nchw
        if g_cls is not None:
            gv = gv + _nchw(g_cls)
        g, = torch.autograd.grad(v, [l], gv)
    g = _nhwc(g, torch.float32)
    if pad_to > g.shape[3]:
        g = torch.cat([g, torch.zeros(g.shape[:3] + (pad_to - g.shape[3],))], dim=3).contiguous()
    return g


def preprocess_bwd(g_nhwc, pre_nchw, blur, normalize=True, taps_cache=None):
    _launches[0] += 1
    from oracle import nvae_ref
    n, h, w, c = g_nhwc.shape
    x = torch.zeros((n, c, h, w), requires_grad=True)
    mask = ((pre_nchw >= 0) & (pre_nchw <= 1)).float()
    g = _nchw(g_nhwc) * mask * (2.0 if normalize else 1.0)
    if not blur:
        return g.contiguous()
    with torch.enable_grad():
        y = nvae_ref.gaussian_blur(x)
        gx, = torch.autograd.grad(y, [x], g)
    return gx


def launch_count(reset=False):
    v = _launches[0]
    if reset:
        _launches[0] = 0
    return v
