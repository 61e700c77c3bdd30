"""GPU: ragged / edge shapes of the purification call (the reference tests none of these; its product batch is 1 x EoT-32).
Per-sample results must not depend on the batch a sample sits in: odd batches leave half-empty two-image tiles at the 8 x 8 scale
(fused cell, tensor-core conv), batch 1 is the EoT unit, and an empty batch must come back empty."""
import pytest
import torch

from gen_adversarial_b200 import synth
from gen_adversarial_b200.defenses.ours.models import NVAEDefenseModel, CelebaIdentityClassifier
from gen_adversarial_b200.defenses.wrappers import EoTWrapper

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def dm():
    nv = synth.make_nvae_checkpoint(seed=0)
    vg = synth.make_vgg11_checkpoint(100, seed=1)
    clf = CelebaIdentityClassifier(vg, DEV, mode="bf16")
    m = NVAEDefenseModel(clf, nv, [0.5] * 24, 0.7, 1.0, True, DEV, mode="bf16")
    m.noise_seed = 11                       # Philox streams keyed by the global sample index
    return m


@pytest.mark.parametrize("b", [1, 3, 5])
def test_results_do_not_depend_on_the_batch_a_sample_sits_in(dm, b):
    x = synth.synthetic_batch(8, seed=5)[0].to(DEV)
    with torch.no_grad():
        full_logits, full = dm(x, preds_only=False)
        part_logits, part = dm(x[:b].contiguous(), preds_only=False)
    assert part.shape == (b, 3, 64, 64) and part_logits.shape == (b, 100)
    assert (part - full[:b]).abs().max().item() <= 1e-6
    assert (part_logits - full_logits[:b]).abs().max().item() <= 1e-3 * max(1.0, full_logits.abs().max().item())
    dm.sample_offset = 8 - b                # the same samples as the TAIL of the batch (different tile partners at 8 x 8)
    try:
        with torch.no_grad():
            _, tail = dm(x[8 - b:].contiguous(), preds_only=False)
    finally:
        dm.sample_offset = 0
    assert (tail - full[8 - b:]).abs().max().item() <= 1e-6


def test_empty_batch(dm):
    x = torch.empty(0, 3, 64, 64, device=DEV)
    with torch.no_grad():
        logits, pur = dm(x, preds_only=False)
    assert logits.shape == (0, 100) and pur.shape == (0, 3, 64, 64)


def test_eot_wrapper_reference_semantics(dm):
    """wrappers.py:15-24: (1,3,h,w) -> repeat eot_steps -> mean over replicas, keepdim; batched extension averages per image"""
    x = synth.synthetic_batch(3, seed=6)[0].to(DEV)
    alphas = dm.interpolation_alphas
    dm.interpolation_alphas = [0.0] * 24    # no resampling + eps below: deterministic, so the replicas agree
    eps, blur = dm.eps, dm.blur_input
    dm.eps, dm.blur_input = 0.0, False
    try:
        eot = EoTWrapper(dm, 4).eval()
        with torch.no_grad():
            single = dm(x[:1])
            avg = eot(x[:1])
            assert avg.shape == (1, 100) and (avg - single).abs().max().item() <= 1e-3 * max(1.0, single.abs().max().item())
            batched = eot(x)
            assert batched.shape == (3, 100)
            assert (batched[:1] - single).abs().max().item() <= 1e-3 * max(1.0, single.abs().max().item())
    finally:
        dm.interpolation_alphas, dm.eps, dm.blur_input = alphas, eps, blur


@pytest.mark.parametrize("kind,res", [("trans", 128), ("e4e", 256)])
def test_stylegan_defenses_batch_composition_independence(kind, res):
    """configs 3 / 4 (bf16 product mode, Philox noise): a sample's purified image does not depend on the batch it sits in -- batch 1 and an
    odd batch against the same samples inside a batch of 4 (generator batch slicing, tensor-core tiles spanning images at low resolution)"""
    from tests.test_stylegan_paths_gpu import _defense
    m = _defense(kind, "bf16")
    m.interpolation_alphas = [0.3] * len(m.interpolation_alphas)
    m.eps, m.blur_input, m.noise_seed = 1.0, kind == "trans", 23
    x = torch.rand(4, 3, res, res, generator=torch.Generator().manual_seed(9)).to(DEV)
    with torch.no_grad():
        _, full = m(x, preds_only=False)
        for b in (1, 3):
            _, part = m(x[:b].contiguous(), preds_only=False)
            assert part.shape == (b, 3, res, res)
            assert (part - full[:b]).abs().max().item() <= 1e-5, (kind, b)
        empty_logits, empty = m(x[:0], preds_only=False)
        assert empty.shape == (0, 3, res, res) and empty_logits.shape[0] == 0
