"""GPU: backward (input-gradient) kernels against torch restatements, the autograd bridge against torch autograd
through the oracle, and the PGD-Linf loop (BASELINE config 5; parity unpinned -- compared with the oracle's PGD)."""
import pytest
import torch

from gen_adversarial_b200 import ops, synth
from gen_adversarial_b200._lib import ACT_NONE, ACT_SILU, ACT_ELU, ACT_RELU
from gen_adversarial_b200.attacks import PGDLinf
from gen_adversarial_b200.nvae_spec import NvaeSpec, NVAE_C32_CONFIG, NVAE_C32_RESOLUTION, tiny_config
from gen_adversarial_b200.defenses.ours.models import NVAEDefenseModel, CelebaIdentityClassifier
from oracle import nvae_ref
from tests import emu_ops

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_affine_act_bwd_and_add():
    g = torch.Generator().manual_seed(0)
    x, gr = torch.randn(2, 8, 8, 16, generator=g), torch.randn(2, 8, 8, 16, generator=g)
    sc, sh = torch.rand(16, generator=g) + 0.5, torch.randn(16, generator=g) * 0.2
    add = torch.randn(2, 8, 8, 16, generator=g)
    for act, s, h in ((ACT_SILU, sc, sh), (ACT_ELU, None, None), (ACT_SILU, None, None)):
        ref = emu_ops.affine_act_bwd(gr, x, s, h, act, torch.float32, add=add)
        got = ops.affine_act_bwd(gr.to(DEV), x.to(DEV), s.to(DEV) if s is not None else None, h.to(DEV) if h is not None else None,
                                 act, torch.float32, add=add.to(DEV))
        assert (got.cpu() - ref).abs().max().item() <= 1e-5
    assert (ops.add(x.to(DEV), gr.to(DEV).bfloat16(), torch.float32).cpu() - (x + gr.bfloat16().float())).abs().max().item() <= 1e-6


@pytest.mark.parametrize("c,hw", [(64, 32), (256, 8), (24, 8), (128, 16), (32, 64), (512, 4)])
def test_se_residual_bwd(c, hw):
    g = torch.Generator().manual_seed(c)
    r, go = torch.randn(2, hw, hw, c, generator=g), torch.randn(2, hw, hw, c, generator=g)
    hid = max(c // 16, 4)
    se = (torch.randn(hid, c, generator=g) * 0.3, torch.randn(hid, generator=g) * 0.1,
          torch.randn(c, hid, generator=g) * 0.3, torch.randn(c, generator=g) * 0.1)
    ref = emu_ops.se_residual_bwd(go, r, None, se, 0.1, torch.float32)
    sums = ops.channel_sum(r.to(DEV))
    got = ops.se_residual_bwd(go.to(DEV), r.to(DEV), sums, tuple(t.to(DEV) for t in se), 0.1, torch.float32)
    assert (got.cpu() - ref).abs().max().item() <= 2e-5 * max(1.0, ref.abs().max().item())
    # the attack path's dtypes: fp32 stream gradient, bf16 branch output, bf16 result
    rb = r.bfloat16()
    refb = emu_ops.se_residual_bwd(go, rb.float(), None, se, 0.1, torch.float32)
    gotb = ops.se_residual_bwd(go.to(DEV), rb.to(DEV), ops.channel_sum(rb.to(DEV)), tuple(t.to(DEV) for t in se), 0.1, torch.bfloat16)
    assert gotb.dtype == torch.bfloat16 and (gotb.float().cpu() - refb).abs().max().item() <= 1e-2 * max(1.0, refb.abs().max().item())


def test_resampling_backward_kernels():
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 16, 16, 8, generator=g)
    m = torch.randn(2, 8, 8, 8, generator=g)
    assert (ops.sumpool2x2(x.to(DEV), torch.float32, mul=m.to(DEV)).cpu() - emu_ops.sumpool2x2(x, torch.float32, mul=m)).abs().max().item() <= 1e-5
    assert (ops.upsample_bilinear2x_bwd(x.to(DEV), torch.float32).cpu() - emu_ops.upsample_bilinear2x_bwd(x, torch.float32)).abs().max().item() <= 1e-5
    xb, mb = torch.randn(3, 12, 20, 24, generator=g).bfloat16(), torch.randn(3, 6, 10, 24, generator=g).bfloat16()      # attack-path dtypes
    refb = emu_ops.sumpool2x2(xb.float(), torch.float32, mul=mb.float())
    gotb = ops.sumpool2x2(xb.to(DEV), torch.bfloat16, mul=mb.to(DEV))
    assert (gotb.float().cpu() - refb).abs().max().item() <= 2e-2 * max(1.0, refb.abs().max().item())
    x3 = torch.randn(2, 4, 4, 6, generator=g)                                                                          # C % 4 != 0: scalar kernel
    assert (ops.sumpool2x2(x3.to(DEV), torch.float32).cpu() - emu_ops.sumpool2x2(x3, torch.float32)).abs().max().item() <= 1e-5
    xin = torch.randn(2, 16, 16, 8, generator=g)
    go = torch.randn(2, 8, 8, 8, generator=g)
    for relu in (False, True):
        got = ops.maxpool2x2_bwd(xin.to(DEV), go.to(DEV), relu, torch.float32)
        assert (got.cpu() - emu_ops.maxpool2x2_bwd(xin, go, relu, torch.float32)).abs().max().item() <= 1e-6


def test_depth_to_space2_and_stride2_phase_dgrad():
    """depth-to-space kernel vs the torch restatement; stride-2 dgrad through the phase conv (tensor cores) vs autograd"""
    import torch.nn.functional as F
    from gen_adversarial_b200.nvae_engine import stride2_dgrad_phase_weights
    from gen_adversarial_b200.fold import Folder
    g = torch.Generator().manual_seed(0)
    x4 = torch.randn(3, 5, 7, 4 * 12, generator=g)
    assert torch.equal(ops.depth_to_space2(x4.to(DEV)).cpu(), emu_ops.depth_to_space2(x4))
    for k, pad in ((3, 1), (1, 0)):
        cin, cout, n, h = 32, 64, 4, 32
        w = torch.randn(cout, cin, k, k, generator=g, dtype=torch.float64) / (k * cin ** 0.5)
        x = torch.randn(n, cin, h, h, generator=g, dtype=torch.float64, requires_grad=True)
        y = F.conv2d(x, w, None, stride=2, padding=pad)
        go = torch.randn(y.shape, generator=g, dtype=torch.float64)
        ref, = torch.autograd.grad(y, [x], go)
        P = Folder({}, DEV, want_tc=True).conv(stride2_dgrad_phase_weights(w, pad), None, stride=1, pad=k // 2, simt=False)
        gb = go.permute(0, 2, 3, 1).contiguous().to(DEV, torch.bfloat16)
        assert ops.conv2d_tc_supported(gb, P)
        _, o4 = ops.conv2d_tc(gb, P, want_bf16=False, want_f32=True)
        got = ops.depth_to_space2(o4).permute(0, 3, 1, 2).cpu().double()
        assert (got - ref).abs().max().item() <= 2e-2 * ref.abs().max().item()


def test_dwconv_extras_and_conv_epilogue_extras():
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 8, 8, 32, generator=g)
    w, b = torch.randn(25, 32, generator=g) * 0.2, torch.randn(32, generator=g) * 0.1
    m = torch.randn(2, 8, 8, 32, generator=g)
    ry, rd = emu_ops.dwconv5x5(x, w, b, ACT_SILU, False, torch.float32, want_dact=True)
    gy, gd = ops.dwconv5x5(x.to(DEV), w.to(DEV), b.to(DEV), ACT_SILU, False, torch.float32, want_dact=True)
    assert (gy.cpu() - ry).abs().max().item() <= 1e-5 and (gd.cpu() - rd).abs().max().item() <= 1e-5
    rm = emu_ops.dwconv5x5(x, w, None, ACT_NONE, False, torch.float32, mul=m)
    gm = ops.dwconv5x5(x.to(DEV), w.to(DEV), None, ACT_NONE, False, torch.float32, mul=m.to(DEV))
    assert (gm.cpu() - rm).abs().max().item() <= 1e-5
    from tests.test_kernels_gpu import _layer, _to_dev
    L = _layer(64, 64, 3, 1, 1, post_act=ACT_SILU, tc=True)
    xs = torch.randn(2, 16, 16, 64, generator=g)
    mm = torch.randn(2, 16, 16, 64, generator=g)
    ry, rd = emu_ops.conv2d_simt(xs, L, torch.float32, mul=mm, mul_mode=2, want_dact=True)
    gy, gd = ops.conv2d_simt(xs.to(DEV), _to_dev(L), torch.float32, mul=mm.to(DEV), mul_mode=2, want_dact=True)
    assert (gy.cpu() - ry).abs().max().item() <= 2e-5 and (gd.cpu() - rd).abs().max().item() <= 2e-5
    # tensor-core kernel: mul epilogue + dact output
    xb = xs.bfloat16()
    dact_ref = torch.empty(2, 16, 16, 64)
    _, rf = emu_ops.conv2d_tc(xb, L, want_bf16=False, want_f32=True, mul=mm.bfloat16(), mul_mode=1, dact_out=dact_ref)
    dact = torch.empty(2, 16, 16, 64, device=DEV, dtype=torch.bfloat16)
    _, gf = ops.conv2d_tc(xb.to(DEV), _to_dev(L), want_bf16=False, want_f32=True, mul=mm.bfloat16().to(DEV), mul_mode=1, dact_out=dact)
    assert (gf.cpu() - rf).abs().max().item() <= 2e-3 * max(1.0, rf.abs().max().item())
    assert (dact.float().cpu() - dact_ref).abs().max().item() <= 2e-2


def test_latent_mix_and_discmix_backward():
    g = torch.Generator().manual_seed(0)
    n, h, z, zc = 2, 8, 20, 24
    q, p = torch.randn(n, h, h, z, generator=g) * 2, torch.randn(n, h, h, 2 * z, generator=g) * 2
    eps, gz = torch.randn(n, z, h, h, generator=g), torch.randn(n, h, h, zc, generator=g)
    alpha = torch.tensor([0.37])
    for pp in (None, p):
        rq, rp = emu_ops.latent_mix_bwd(gz, q, pp, eps, 0, 0, 0, alpha, 0.6, z, zc)
        gq, gp = ops.latent_mix_bwd(gz.to(DEV), q.to(DEV), pp.to(DEV) if pp is not None else None, eps.to(DEV), 0, 0, 0,
                                    alpha.to(DEV), 0.6, z, zc)
        assert (gq.cpu() - rq).abs().max().item() <= 1e-5
        if pp is not None:
            assert (gp.cpu() - rp).abs().max().item() <= 1e-5
    logits = torch.randn(2, 8, 8, 100, generator=g) * 1.5
    gp_, gc_ = torch.randn(2, 3, 8, 8, generator=g), torch.randn(2, 8, 8, 3, generator=g)
    ref = emu_ops.discmix_mean_bwd(logits, 10, gp_, gc_)
    got = ops.discmix_mean_bwd(logits.to(DEV), 10, gp_.to(DEV), gc_.to(DEV))
    assert (got.cpu() - ref).abs().max().item() <= 1e-5
    # ragged pixel count (not a multiple of the 128-pixel block) + channel padding for the dgrad conv + one gradient source only
    logits = torch.randn(3, 7, 9, 100, generator=g) * 1.5
    gc_ = torch.randn(3, 7, 9, 3, generator=g)
    ref = emu_ops.discmix_mean_bwd(logits, 10, None, gc_)
    got = ops.discmix_mean_bwd(logits.to(DEV), 10, None, gc_.to(DEV), pad_to=104)
    assert got.shape == (3, 7, 9, 104) and (got[..., :100].cpu() - ref).abs().max().item() <= 1e-5 and got[..., 100:].abs().max().item() == 0


@pytest.mark.parametrize("blur,eps", [(True, 1.0), (False, 2.0)])
def test_preprocess_backward(blur, eps):
    g = torch.Generator().manual_seed(0)
    x = torch.rand(2, 3, 32, 32, generator=g)
    noise = torch.randn(2, 3, 32, 32, generator=g)
    gr = torch.randn(2, 32, 32, 3, generator=g)
    xr = x.clone().requires_grad_(True)
    y = (nvae_ref.preprocess(xr, noise, eps, blur) - 0.5) / 0.5
    ref, = torch.autograd.grad(y, [xr], gr.permute(0, 3, 1, 2))
    _, pre = ops.preprocess(x.to(DEV), noise.to(DEV), eps, blur, torch.float32, save_pre=True)
    got = ops.preprocess_bwd(gr.to(DEV), pre, blur)
    assert (got.cpu() - ref).abs().max().item() <= 1e-5


@pytest.fixture(scope="module")
def c32_models():
    return synth.make_nvae_checkpoint(seed=0), synth.make_vgg11_checkpoint(100, seed=1)


def _oracle_grad(nv, vg, x, y, alphas, noises, eps, blur):
    spec = NvaeSpec(NVAE_C32_CONFIG, NVAE_C32_RESOLUTION)
    vgg = nvae_ref.build_vgg11(vg["state_dict"], 100)
    xr = x.clone().requires_grad_(True)
    logits, pur = nvae_ref.defense_call(nv["state_dict_temp=0.6"], spec, vgg, xr, alphas, noises, eps, blur)
    loss = torch.nn.functional.cross_entropy(logits, y)
    g, = torch.autograd.grad(loss, [xr])
    return g, logits.detach()


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_input_gradient_through_the_api_matches_oracle_autograd(c32_models, mode):
    """`torch.autograd.grad(loss, [x])` on NVAEDefenseModel + VGG11 (the attack path, untargeted.py:146) and the fused
    loss_input_grad primitive, against torch autograd through the oracle (C32, cosine alphas, blur + noise)."""
    nv, vg = c32_models
    spec = NvaeSpec(NVAE_C32_CONFIG, NVAE_C32_RESOLUTION)
    alphas = [0.7 * 0.5 * (1 - __import__("math").cos(__import__("math").pi * i / 24)) for i in range(1, 25)]
    x, _ = synth.synthetic_batch(2, seed=5)
    y = torch.tensor([3, 41])
    noises = synth.synthetic_noise(spec, 2, seed=6)
    g_ref, logits_ref = _oracle_grad(nv, vg, x, y, alphas, noises, 2.0, True)
    clf = CelebaIdentityClassifier(vg, DEV, mode=mode)
    dm = NVAEDefenseModel(clf, nv, [a / 0.7 for a in alphas], 0.7, 2.0, True, DEV, mode=mode).eval()
    dm.set_explicit_noise(noises)
    xd = x.to(DEV).requires_grad_(True)
    logits = dm(xd)
    loss = torch.nn.functional.cross_entropy(logits, y.to(DEV))
    g1, = torch.autograd.grad(loss, [xd], retain_graph=True)
    g1b, = torch.autograd.grad(2 * loss, [xd])                       # repeated backward through the same node
    _, g2, _ = dm.loss_input_grad(x.to(DEV), y.to(DEV))
    for name, g in (("autograd", g1), ("fused", g2)):
        g = g.cpu()
        rel_l2 = ((g - g_ref).norm() / g_ref.norm()).item()
        cos = torch.nn.functional.cosine_similarity(g.flatten(), g_ref.flatten(), dim=0).item()
        sign = (torch.sign(g) == torch.sign(g_ref)).float().mean().item()
        print(f"[{mode}] {name}: grad rel-L2 err {rel_l2:.3e}, cosine {cos:.6f}, sign agreement {sign:.4f}, |g_ref|max {g_ref.abs().max():.3e}")
        if mode == "fp32":
            assert rel_l2 <= 1e-2 and cos >= 0.9999
        else:
            # bf16: the purifier's gradient is accurate (cos 0.99998, test below); the random-init, BN-calibrated VGG11 is a
            # ReLU/max-pool network whose unit on/off pattern flips under bf16 rounding (the CPU emulation of the same bf16
            # arithmetic shows the same 0.94-0.96), so the end-to-end direction is only required to stay close; what the attack
            # needs -- the same robust-accuracy counts as the fp32 path -- is gated in test_pgd_50_steps_robust_counts below
            assert cos >= 0.93
    assert torch.allclose(g1b, 2 * g1, rtol=1e-3, atol=1e-9)
    assert ((g1 - g2).norm() / g2.norm()).item() <= 1e-4


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_pgd_linf_matches_oracle_pgd(c32_models, mode):
    """config 5 at test size: 3 PGD steps, eps 8/255, step 2/255, fixed explicit noise per step."""
    nv, vg = c32_models
    spec = NvaeSpec(NVAE_C32_CONFIG, NVAE_C32_RESOLUTION)
    import math
    alphas = [0.7 * 0.5 * (1 - math.cos(math.pi * i / 24)) for i in range(1, 25)]
    x, _ = synth.synthetic_batch(2, seed=8)
    vgg = nvae_ref.build_vgg11(vg["state_dict"], 100)
    steps = 3
    sched = [synth.synthetic_noise(spec, 2, seed=100 + i) for i in range(steps + 1)]
    with torch.no_grad():
        y = nvae_ref.defense_call(nv["state_dict_temp=0.6"], spec, vgg, x, alphas, sched[0], 2.0, False)[0].argmax(1)

    def logits_fn(xa, i):
        return nvae_ref.defense_call(nv["state_dict_temp=0.6"], spec, vgg, xa, alphas, sched[i], 2.0, False)[0]

    adv_ref = nvae_ref.pgd_linf_attack(logits_fn, x, y, 8 / 255, 2 / 255, steps)
    clf = CelebaIdentityClassifier(vg, DEV, mode=mode)
    dm = NVAEDefenseModel(clf, nv, [a / 0.7 for a in alphas], 0.7, 2.0, False, DEV, mode=mode).eval()
    succ, linf, adv = PGDLinf(8 / 255, 2 / 255, steps)(x.to(DEV), y.to(DEV), dm, noise_schedule=sched)
    same = ((adv.cpu() - adv_ref).abs() <= 1e-6).float().mean().item()
    print(f"[{mode}] PGD: {100 * same:.2f}% of adversarial pixels identical to the oracle's; linf {linf.tolist()}")
    assert linf.max().item() <= 8 / 255 + 1e-6
    assert adv.min().item() >= 0 and adv.max().item() <= 1
    assert same >= (0.97 if mode == "fp32" else 0.60)


def test_purifier_gradient_bf16_is_accurate(c32_models):
    """bf16 tensor-core backward through the NVAE purifier alone (gradient of a random linear functional of the
    purified image) against torch autograd through the oracle."""
    import math
    from gen_adversarial_b200 import autograd as ga
    from gen_adversarial_b200.nvae_engine import NvaeEngine
    nv, _ = c32_models
    spec = NvaeSpec(NVAE_C32_CONFIG, NVAE_C32_RESOLUTION)
    sd = nv["state_dict_temp=0.6"]
    alphas = [0.7 * 0.5 * (1 - math.cos(math.pi * i / 24)) for i in range(1, 25)]
    x, _ = synth.synthetic_batch(2, seed=5)
    noises = synth.synthetic_noise(spec, 2, seed=6)
    wgt = torch.randn(2, 3, 64, 64, generator=torch.Generator().manual_seed(0))
    xo = x.clone().requires_grad_(True)
    _, pur = nvae_ref.defense_call(sd, spec, None, xo, alphas, noises, 2.0, True)
    g_ref, = torch.autograd.grad((pur * wgt).sum(), [xo])
    eng = NvaeEngine(sd, spec, DEV, "bf16")
    tape = ga.Tape()
    xin, pre = ops.preprocess(x.to(DEV), noises[0].to(DEV), 2.0, True, eng.adt, save_pre=True)
    eng.purify(xin, torch.tensor(alphas, device=DEV), [n.to(DEV) for n in noises[1:]], tape=tape)
    gx = ops.preprocess_bwd(eng.backward(tape.nvae, wgt.to(DEV), None), pre, True).cpu()
    cos = torch.nn.functional.cosine_similarity(gx.flatten(), g_ref.flatten(), dim=0).item()
    rel = ((gx - g_ref).norm() / g_ref.norm()).item()
    print(f"[bf16] purifier-only gradient: rel-L2 {rel:.3e}, cosine {cos:.6f}")
    # two equally valid bf16 evaluation orders of the same network differ by about the bf16 noise itself (measured: the persistent 3x3 kernel
    # sums K in a different order than the per-tap kernel -- 1e-7 relative per conv -- and the purified images of the two differ by 7e-3,
    # the gradients by 4e-2 relative); the gate is the direction (cosine) plus a noise-level bound on the relative error
    assert cos >= 0.999 and rel <= 5e-2


def test_cuda_graph_replay_matches_eager_and_draws_fresh_noise():
    """CUDA-graph mode of the public API: same result as the eager call for the same seed salt state; every replay draws fresh
    noise (device-side salt); alphas stay run-time inputs; the graphed PGD loop respects the eps ball and flips predictions like
    the eager loop."""
    from gen_adversarial_b200 import graphs
    from gen_adversarial_b200.attacks import PGDLinf
    from gen_adversarial_b200.defenses.ours.models import NVAEDefenseModel, CelebaIdentityClassifier
    cfg, res = tiny_config(initial_channels=16, groups=2, scales=2, latent=4), (3, 32, 32)
    spec = NvaeSpec(cfg, res)
    clf = CelebaIdentityClassifier({"state_dict": synth.make_vgg11_state_dict(10, seed=3, device=DEV)}, DEV, mode="fp32", n_classes=10, image_size=32)
    n = spec.n_latents
    dm = NVAEDefenseModel(clf, synth.make_nvae_checkpoint(cfg, res, seed=3), [0.5] * n, 1.0, 1.0, True, DEV, mode="fp32")
    dm.noise_seed = 5
    x = synth.synthetic_batch(4, res, seed=2)[0].to(DEV)
    graphs.enable_seed_salt(DEV).zero_()
    eager = dm(x).clone()                                     # salt = 0: seed 5
    dm.enable_cuda_graph(True)
    graphs.enable_seed_salt(DEV).zero_()
    g1 = dm(x).clone()                                        # capture (warm-up + capture bump the salt), then first replay
    g2 = dm(x).clone()
    assert (g1 - g2).abs().max().item() > 1e-6               # fresh noise per replay
    salt = graphs.enable_seed_salt(DEV)
    s_before = int(salt.item())
    g3 = dm(x).clone()
    dm.enable_cuda_graph(False)
    v = (s_before + 0x9E3779B97F4A7C15) & ((1 << 64) - 1)     # the salt value the third replay ran with (uint64 add)
    # the salt applies to captured launches only: an eager call reproduces the replay when it is given seed + salt by value ...
    dm.noise_seed = (5 + v) & ((1 << 64) - 1)
    e3 = dm(x).clone()
    assert (g3 - e3).abs().max().item() <= 1e-5 * max(1.0, e3.abs().max().item())
    # ... and a fixed seed reproduces the first eager result no matter what the graphs did to the salt in between
    dm.noise_seed = 5
    assert (dm(x) - eager).abs().max().item() <= 1e-5 * max(1.0, eager.abs().max().item())
    # a taped eager forward and its backward see the same eps even if a replay bumps the salt in between
    xg = x.clone().requires_grad_(True)
    loss = dm(xg).square().sum()
    ga_, = torch.autograd.grad(loss, [xg], retain_graph=True)
    dm.enable_cuda_graph(True)
    dm(x)                                                      # graph replay between forward and the second backward
    dm.enable_cuda_graph(False)
    gb_, = torch.autograd.grad(loss, [xg])
    assert torch.equal(ga_, gb_)
    # alphas are read from device memory at replay time
    dm.enable_cuda_graph(True)
    a = dm(x).clone()
    dm.interpolation_alphas = [0.0] * n
    b = dm(x).clone()
    assert (a - b).abs().max().item() > 1e-4
    # graphed PGD
    y = dm(x).argmax(1)
    succ, linf, x_adv = PGDLinf(8 / 255, 2 / 255, 5)(x, y, dm)
    assert linf.max().item() <= 8 / 255 + 1e-6 and x_adv.min().item() >= 0 and x_adv.max().item() <= 1
    assert (x_adv - x).abs().max().item() > 1e-3
    graphs.enable_seed_salt(DEV).zero_()


def _tiny_attack_setup(batch, steps, seed):
    """tiny NVAE (2 scales x 2 groups, 32x32) + VGG11(10 classes): small enough that the oracle's 50-step PGD takes seconds on CPU"""
    cfg, res = tiny_config(initial_channels=16, groups=2, scales=2, latent=4), (3, 32, 32)
    spec = NvaeSpec(cfg, res)
    nv = synth.make_nvae_checkpoint(cfg, res, seed=3)
    vg = {"state_dict": synth.make_vgg11_state_dict(10, seed=3)}
    n = spec.n_latents
    alphas = [0.7 * 0.5 * (1 - __import__("math").cos(__import__("math").pi * i / n)) for i in range(1, n + 1)]
    x, _ = synth.synthetic_batch(batch, res, 10, seed=seed)
    sched = [synth.synthetic_noise(spec, batch, seed=1000 + 7 * i) for i in range(steps + 1)]
    return cfg, res, spec, nv, vg, alphas, x, sched


def test_pgd_50_steps_robust_counts_match_oracle():
    """BASELINE configs[4] at full attack length (eps 8/255, step 2/255, 50 steps, batch 16, explicit noise per step).
    PGD is a sign iteration: one flipped sign of a near-zero gradient component moves that pixel by 4/255 and the two trajectories
    then drift apart chaotically (measured: fp32 free-running 74% identical pixels after 50 steps although every success flag agrees),
    so pixel identity is checked TEACHER-FORCED -- at every one of the 50 steps the CUDA path takes the oracle's iterate and must produce
    the oracle's next iterate on >= 99% of the pixels (fp32) -- and the attack OUTCOME is checked free-running: identical per-image
    success flags (= identical robust-accuracy count) for the fp32 path; the bf16 path (what bench.py times) is gated on the count."""
    steps, batch = 50, 16
    cfg, res, spec, nv, vg, alphas, x, sched = _tiny_attack_setup(batch, steps, seed=21)
    vgg = nvae_ref.build_vgg11(vg["state_dict"], 10)
    sd = nv["state_dict_temp=0.6"]
    with torch.no_grad():
        y = nvae_ref.defense_call(sd, spec, vgg, x, alphas, sched[0], 1.0, True)[0].argmax(1)

    def logits_fn(xa, i):
        return nvae_ref.defense_call(sd, spec, vgg, xa, alphas, sched[i], 1.0, True)[0]

    traj = [x.clone()]                                         # the oracle's iterates x_adv^0 .. x_adv^50
    for i in range(steps):
        xa = traj[-1].clone().requires_grad_(True)
        g, = torch.autograd.grad(torch.nn.functional.cross_entropy(logits_fn(xa, i), y), [xa])
        traj.append(nvae_ref.pgd_linf_step(traj[-1], g, x, 2 / 255, 8 / 255).detach())
    adv_ref = traj[-1]
    with torch.no_grad():
        succ_ref = logits_fn(adv_ref, steps).argmax(1) != y
    print(f"oracle: {int(succ_ref.sum())}/{batch} attacks succeed (robust count {batch - int(succ_ref.sum())})")
    results = {}
    for mode in ("fp32", "bf16"):
        clf = CelebaIdentityClassifier({"state_dict": {k: v.to(DEV) for k, v in vg["state_dict"].items()}}, DEV, mode=mode, n_classes=10, image_size=32)
        dm = NVAEDefenseModel(clf, nv, [a / 0.7 for a in alphas], 0.7, 1.0, True, DEV, mode=mode).eval()
        # free-running attack: outcome
        succ, linf, adv = PGDLinf(8 / 255, 2 / 255, steps)(x.to(DEV), y.to(DEV), dm, noise_schedule=sched)
        same_free = ((adv.cpu() - adv_ref).abs() <= 1e-6).float().mean().item()
        flags = int((succ.cpu() == succ_ref).sum())
        assert linf.max().item() <= 8 / 255 + 1e-6 and adv.min().item() >= 0 and adv.max().item() <= 1
        # teacher-forced: one CUDA step from each oracle iterate
        worst, mean = 1.0, 0.0
        xd, yd = x.to(DEV), y.to(DEV)
        for i in range(steps):
            dm.set_explicit_noise(sched[i])
            xa = traj[i].to(DEV).contiguous().clone()
            _, grad, _ = dm.loss_input_grad(xa, yd)
            ops.pgd_linf_step_(xa, grad, xd, 2 / 255, 8 / 255)
            same = ((xa.cpu() - traj[i + 1]).abs() <= 1e-6).float().mean().item()
            worst, mean = min(worst, same), mean + same / steps
        dm.set_explicit_noise(None)
        results[mode] = (worst, mean, flags, int(succ.sum()))
        print(f"[{mode}] 50-step PGD: teacher-forced identical pixels per step: worst {100 * worst:.2f}% mean {100 * mean:.2f}%; free-running: "
              f"{100 * same_free:.2f}% identical after 50 steps, success flags equal on {flags}/{batch} images, "
              f"robust count {batch - int(succ.sum())} (oracle {batch - int(succ_ref.sum())})")
    assert results["fp32"][2] == batch, "fp32 path: per-image success flags must equal the oracle's"
    assert results["fp32"][0] >= 0.99, "fp32 path: every single PGD step must reproduce the oracle's step on >= 99% of the pixels"
    # bf16: the robust COUNT may differ by at most one image of 16 from the fp32 / oracle count
    assert abs(results["bf16"][3] - int(succ_ref.sum())) <= 1


def test_clean_counts_identical_on_64_images(c32_models):
    """north star: clean accuracy counts identical for the fp32 path -- C32 + VGG11, 64 images, explicit noise, every arg-max equal to
    the oracle's; the bf16 path is reported (count of differing arg-maxes)."""
    nv, vg = c32_models
    spec = NvaeSpec(NVAE_C32_CONFIG, NVAE_C32_RESOLUTION)
    import math
    alphas = [0.7 * 0.5 * (1 - math.cos(math.pi * i / 24)) for i in range(1, 25)]
    x, _ = synth.synthetic_batch(64, seed=31)
    noises = synth.synthetic_noise(spec, 64, seed=32)
    vgg = nvae_ref.build_vgg11(vg["state_dict"], 100)
    with torch.no_grad():
        ref = nvae_ref.defense_call(nv["state_dict_temp=0.6"], spec, vgg, x, alphas, noises, 2.0, True)[0]
    y = ref.argmax(1)
    assert y.unique().numel() > 4                                          # not a constant predictor
    for mode in ("fp32", "bf16"):
        clf = CelebaIdentityClassifier(vg, DEV, mode=mode)
        dm = NVAEDefenseModel(clf, nv, [a / 0.7 for a in alphas], 0.7, 2.0, True, DEV, mode=mode).eval()
        dm.set_explicit_noise(noises)
        with torch.no_grad():
            got = dm(x.to(DEV)).cpu()
        diff = int((got.argmax(1) != y).sum())
        rel = ((got - ref).abs().max() / ref.abs().max()).item()
        print(f"[{mode}] 64 images: {diff} arg-max differences, logits rel err {rel:.2e}")
        if mode == "fp32":
            assert diff == 0 and rel <= 1e-3
        else:
            # bf16 path: not a north-star gate (counts must be identical for the fp32 path only); measured 4 of 64 arg-maxes differ on
            # this random-init, near-tied 100-class head (logits rel err 2.3e-2) -- gated so that a regression shows
            assert diff <= 6


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_loss_input_grad_on_part_batches_changes_nothing(mode):
    """`set_streams(2)`: the fused attack primitive runs two half batches on two streams (taping forward + dgrad sweep each) with the
    loss gradient weighted by the part's share of the batch: loss, gradient and predictions are bit-identical to the unsplit call, and
    the graphed PGD loop built on it stays inside the eps ball."""
    from gen_adversarial_b200.attacks import PGDLinf
    from gen_adversarial_b200.defenses.ours.models import NVAEDefenseModel, CelebaIdentityClassifier
    cfg, res = tiny_config(initial_channels=16, groups=2, scales=2, latent=4), (3, 32, 32)
    spec = NvaeSpec(cfg, res)
    clf = CelebaIdentityClassifier({"state_dict": synth.make_vgg11_state_dict(10, seed=3, device=DEV)}, DEV, mode=mode, n_classes=10, image_size=32)
    dm = NVAEDefenseModel(clf, synth.make_nvae_checkpoint(cfg, res, seed=3), [0.5] * spec.n_latents, 1.0, 1.0, True, DEV, mode=mode)
    dm.noise_seed = 11
    x, y = synth.synthetic_batch(8, res, 10, seed=2)
    x, y = x.to(DEV), y.to(DEV)
    cnt = torch.zeros(1, dtype=torch.int64, device=DEV)
    l1, g1, p1 = dm.loss_input_grad(x, y, counter=cnt)
    c1 = int(cnt.item())
    dm.set_streams(2)
    outs = [dm.loss_input_grad(x, y, counter=cnt) for _ in range(3)]      # first: parts back to back; then forked
    torch.cuda.synchronize()
    assert int(cnt.item()) == 4 * c1
    for l, g, p in outs:
        assert torch.equal(l1, l) and torch.equal(p1, p) and torch.equal(g1, g)
    dm.noise_seed = None
    dm.enable_cuda_graph(True)
    succ, linf, x_adv = PGDLinf(8 / 255, 2 / 255, 5)(x, y, dm)
    dm.enable_cuda_graph(False)
    assert linf.max().item() <= 8 / 255 + 1e-6 and x_adv.min() >= 0 and x_adv.max() <= 1
    assert (x_adv - x).abs().max().item() > 1e-3
