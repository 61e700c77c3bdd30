"""CPU: host-side logic (folds, op order, plumbing) of the engines, driven through a torch emulation of the
kernels (tests/emu_ops.py) and compared with the oracle.  The kernels themselves are checked on the GPU."""
import pytest
import torch

from oracle import nvae_ref
from gen_adversarial_b200 import synth, nvae_engine, vgg_engine
from gen_adversarial_b200.nvae_spec import NvaeSpec, tiny_config
from tests import emu_ops


@pytest.fixture()
def emu(monkeypatch):
    monkeypatch.setattr(nvae_engine, "ops", emu_ops)
    monkeypatch.setattr(vgg_engine, "ops", emu_ops)
    return emu_ops


@pytest.mark.parametrize("mode,tol", [("fp32", 2e-5), ("bf16", 2e-2)])
def test_nvae_engine_host_logic_matches_oracle(emu, golden_tiny, mode, tol):
    g = golden_tiny
    spec = NvaeSpec(g["cfg"], g["resolution"])
    eng = nvae_engine.NvaeEngine(g["state_dict"], spec, "cpu", mode, _host_logic_test=True)
    for case in g["cases"]:
        alphas = torch.tensor([a * case["attenuation"] for a in case["alphas"]], dtype=torch.float32)
        xin, _ = emu.preprocess(g["x"], g["noises"][0], case["eps"], case["blur"], eng.adt)
        taps = {}
        eng.taps = taps
        pur, cls = eng.purify(xin, alphas, g["noises"][1:], cls_dtype=eng.adt)
        err = (pur - case["purified"]).abs().max().item()
        assert err <= tol, (mode, case["name"], err)
        assert cls.shape == (3, 32, 32, 3)


def test_nvae_engine_host_logic_3scale(emu):
    """a 3-scale / 3-group architecture with 16 base channels (exercises the tensor-core eligibility rules)."""
    cfg, res = tiny_config(initial_channels=16, groups=3, scales=3, latent=6), (3, 64, 64)
    spec = NvaeSpec(cfg, res)
    sd = synth.make_nvae_state_dict(cfg, res, seed=9)
    x, _ = synth.synthetic_batch(2, res, seed=1)
    noises = synth.synthetic_noise(spec, 2, seed=2)
    alphas = [0.7 * (i + 1) / spec.n_latents for i in range(spec.n_latents)]
    with torch.no_grad():
        _, ref = nvae_ref.defense_call(sd, spec, None, x, alphas, noises, 2.0, True)
    for mode, tol in (("fp32", 2e-5), ("bf16", 2e-2)):
        eng = nvae_engine.NvaeEngine(sd, spec, "cpu", mode, _host_logic_test=True)
        xin, _ = emu.preprocess(x, noises[0], 2.0, True, eng.adt)
        pur, _ = eng.purify(xin, torch.tensor(alphas), noises[1:])
        assert (pur - ref).abs().max().item() <= tol, mode


@pytest.mark.parametrize("nf", [1, 2])
def test_nvae_engine_normalizing_flow_cells(emu, nf):
    """A12: checkpoints built with num_nf_cells (NVAE/model.py:58-59,216-221; applied at src/defenses/ours/models.py:209-210,253-254):
    the engine folds the flow cells into the decoder-combiner biases (exact: the last masked 1x1 conv of every NFCell has an all-zero
    mask); compared with the oracle, which evaluates the masked convs in full."""
    cfg, res = tiny_config(initial_channels=8, groups=2, scales=2, latent=4), (3, 32, 32)
    cfg["num_nf_cells"] = nf
    spec = NvaeSpec(cfg, res)
    assert spec.use_nf
    sd = synth.make_nvae_state_dict(cfg, res, seed=11)
    assert any(k.endswith("cell2.layers.0.mask") for k in sd)
    x, _ = synth.synthetic_batch(2, res, seed=1)
    noises = synth.synthetic_noise(spec, 2, seed=2)
    alphas = [0.7 * (i + 1) / spec.n_latents for i in range(spec.n_latents)]
    with torch.no_grad():
        _, ref = nvae_ref.defense_call(sd, spec, None, x, alphas, noises, 1.0, False)
        cfg0 = dict(cfg); cfg0["num_nf_cells"] = None
        sd0 = {k: v for k, v in sd.items() if not k.startswith("nf_cells.")}
        _, ref0 = nvae_ref.defense_call(sd0, NvaeSpec(cfg0, res), None, x, alphas, noises, 1.0, False)
    assert (ref - ref0).abs().max().item() > 1e-3                  # the flow cells do change the result
    eng = nvae_engine.NvaeEngine(sd, spec, "cpu", "fp32", _host_logic_test=True)
    xin, _ = emu.preprocess(x, noises[0], 1.0, False, eng.adt)
    pur, _ = eng.purify(xin, torch.tensor(alphas), noises[1:])
    assert (pur - ref).abs().max().item() <= 2e-5


def test_vgg_engine_host_logic_matches_torchvision(emu):
    """pool-fold + BN-fold of the VGG11 head against torchvision's vgg11_bn (small head to keep the test light)."""
    torch.manual_seed(0)
    sd = synth.make_vgg11_state_dict(n_classes=10, seed=3, calibrate=True)
    model = nvae_ref.build_vgg11(sd, n_classes=10)
    x = torch.rand(3, 3, 64, 64)
    with torch.no_grad():
        ref = nvae_ref.classify(model, x)
    eng = vgg_engine.Vgg11Engine(sd, "cpu", "fp32", in_hw=64, _host_logic_test=True)
    xin = emu.nchw_to_nhwc(x, torch.float32, 2.0, -1.0)
    out = eng.forward(xin)
    rel = ((out - ref).abs().max() / ref.abs().max()).item()
    assert rel <= 1e-4, rel


@pytest.mark.parametrize("mode,tol", [("fp32", 2e-4), ("bf16", None)])
def test_backward_host_logic_matches_oracle_autograd(emu, golden_tiny, mode, tol):
    """dgrad-only reverse sweep of the engines (tape + backward) against torch autograd through the oracle."""
    from gen_adversarial_b200 import autograd as ga_autograd
    g = golden_tiny
    spec = NvaeSpec(g["cfg"], g["resolution"])
    sd = g["state_dict"]
    case = g["cases"][0]                                           # cosine alphas, noise eps 2
    alphas = [a * case["attenuation"] for a in case["alphas"]]
    x = g["x"][:2].clone()
    noises = [n[:2] for n in g["noises"]]
    wgt = torch.randn(2, 3, 32, 32, generator=torch.Generator().manual_seed(0))
    # oracle gradient
    xo = x.clone().requires_grad_(True)
    _, pur = nvae_ref.defense_call(sd, spec, None, xo, alphas, noises, 1.0, True)
    g_ref, = torch.autograd.grad((pur * wgt).sum(), [xo])
    # engine gradient (purified-image gradient only; the classifier leg is covered below)
    eng = nvae_engine.NvaeEngine(sd, spec, "cpu", mode, _host_logic_test=True)
    tape = ga_autograd.Tape()
    xin, pre = emu.preprocess(x, noises[0], 1.0, True, eng.adt, save_pre=True)
    pur2, _ = eng.purify(xin, torch.tensor(alphas), noises[1:], tape=tape)
    g_x = eng.backward(tape.nvae, wgt, None)
    gx = emu.preprocess_bwd(g_x, pre, True)
    rel = ((gx - g_ref).abs().max() / g_ref.abs().max()).item()
    cos = torch.nn.functional.cosine_similarity(gx.flatten(), g_ref.flatten(), dim=0).item()
    if tol is not None:
        assert rel <= tol, (mode, rel)
    else:   # bf16 gradients: direction matters (PGD uses sign(grad)); max-abs error is dominated by bf16 rounding
        assert cos >= 0.99, (mode, cos, rel)
    # the tape is not consumed: a second backward with another output gradient works (DeepFool / FAB pattern)
    gx2 = emu.preprocess_bwd(eng.backward(tape.nvae, 2 * wgt, None), pre, True)
    assert torch.allclose(gx2, 2 * gx, rtol=1e-3, atol=1e-6 if mode == "fp32" else 1e-3)


def test_vgg_backward_host_logic_matches_torchvision_autograd(emu):
    sd = synth.make_vgg11_state_dict(n_classes=10, seed=3, calibrate=True)
    model = nvae_ref.build_vgg11(sd, n_classes=10)
    x = torch.rand(2, 3, 64, 64, generator=torch.Generator().manual_seed(0))
    y = torch.tensor([3, 7])
    xr = x.clone().requires_grad_(True)
    loss = torch.nn.functional.cross_entropy(nvae_ref.classify(model, xr), y)
    g_ref, = torch.autograd.grad(loss, [xr])
    eng = vgg_engine.Vgg11Engine(sd, "cpu", "fp32", in_hw=64, _host_logic_test=True)
    tape = []
    logits = eng.forward(emu.nchw_to_nhwc(x, torch.float32, 2.0, -1.0), tape=tape)
    lg = logits.clone().requires_grad_(True)
    g_logits, = torch.autograd.grad(torch.nn.functional.cross_entropy(lg, y), [lg])
    g = eng.backward(tape, g_logits).permute(0, 3, 1, 2) * 2.0
    # a ReLU / max-pool unit whose pre-activation is within rounding error of 0 (or of its neighbour) can take the other
    # branch in two correct fp32 implementations, so the max-abs error is not a stable metric: use the relative L2 error
    rel_l2 = ((g - g_ref).norm() / g_ref.norm()).item()
    assert rel_l2 <= 5e-3, rel_l2


@pytest.mark.parametrize("k,pad", [(3, 1), (1, 0)])
def test_stride2_dgrad_phase_weights_match_autograd(k, pad):
    """input gradient of a stride-2 conv = one stride-1 conv over grad_out emitting the 4 output phases + depth-to-space
    (nvae_engine.stride2_dgrad_phase_weights) -- against torch autograd of F.conv2d."""
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(k)
    cin, cout, n, h = 8, 16, 2, 12
    w = torch.randn(cout, cin, k, k, generator=g, dtype=torch.float64)
    x = torch.randn(n, cin, h, h, generator=g, dtype=torch.float64, requires_grad=True)
    y = F.conv2d(x, w, None, stride=2, padding=pad)
    go = torch.randn(y.shape, generator=g, dtype=torch.float64)
    ref, = torch.autograd.grad(y, [x], go)
    wd = nvae_engine.stride2_dgrad_phase_weights(w, pad)
    o4 = F.conv2d(go, wd, None, stride=1, padding=k // 2)                                  # [n, 4*cin, h/2, h/2]
    got = emu_ops.depth_to_space2(o4.permute(0, 2, 3, 1).contiguous()).permute(0, 3, 1, 2)
    assert got.shape == ref.shape
    assert (got.double() - ref).abs().max().item() <= 1e-5 * ref.abs().max().item()


def test_alpha_schedules_match_reference_formulas():
    """alpha_learning/common_utils.py:15-22 (the reference module itself does not import here: botorch / data deps)"""
    import math
    from gen_adversarial_b200.alpha_schedules import get_cosine_alphas, get_linear_alphas
    for n in (16, 18, 24):
        assert get_linear_alphas(n) == [i / n for i in range(1, n + 1)]
        assert get_cosine_alphas(n) == [0.5 * (1 - math.cos(math.pi * (i / n))) for i in range(1, n + 1)]
        assert get_linear_alphas(n)[-1] == 1.0 and abs(get_cosine_alphas(n)[-1] - 1.0) < 1e-15


def test_ablation_models_refuse_cpu_tensors():
    from gen_adversarial_b200.defenses.ablations.models import GaussianBlurDefenseModel, GaussianNoiseDefenseModel
    x = torch.rand(2, 3, 64, 64)
    for m in (GaussianBlurDefenseModel(torch.nn.Identity()), GaussianNoiseDefenseModel(torch.nn.Identity(), eps=0.5)):
        with pytest.raises(RuntimeError):
            m(x)
