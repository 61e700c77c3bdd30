"""Which half of the attack path changes when the halo kernel is on?  (GPU box)"""
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gen_adversarial_b200 import ops, synth, _lib, autograd as ga
from gen_adversarial_b200.nvae_engine import NvaeEngine
from gen_adversarial_b200.nvae_spec import NvaeSpec, NVAE_C32_CONFIG, NVAE_C32_RESOLUTION

DEV = "cuda:0"
L_ = _lib.lib()
nv = synth.make_nvae_checkpoint(seed=0)
spec = NvaeSpec(NVAE_C32_CONFIG, NVAE_C32_RESOLUTION)
sd = nv["state_dict_temp=0.6"]
alphas = [0.7 * 0.5 * (1 - math.cos(math.pi * i / 24)) for i in range(1, 25)]
x, _ = synth.synthetic_batch(2, seed=5)
noises = synth.synthetic_noise(spec, 2, seed=6)
wgt = torch.randn(2, 3, 64, 64, generator=torch.Generator().manual_seed(9))
eng = NvaeEngine(sd, spec, DEV, "bf16")
a_dev = torch.tensor(alphas, device=DEV)


def run(h_fwd, h_bwd):
    tape = ga.Tape()
    L_.ga_tc_halo_enable(h_fwd)
    xin, pre = ops.preprocess(x.to(DEV), noises[0].to(DEV), 2.0, True, eng.adt, save_pre=True)
    pur, _ = eng.purify(xin, a_dev, [n.to(DEV) for n in noises[1:]], tape=tape)
    L_.ga_tc_halo_enable(h_bwd)
    gx = ops.preprocess_bwd(eng.backward(tape.nvae, wgt.to(DEV), None), pre, True)
    torch.cuda.synchronize()
    return pur.cpu(), gx.cpu(), tape


p00, g00, t00 = run(0, 0)
for hf, hb in ((0, 0), (1, 0), (0, 1), (1, 1)):
    p, g, t = run(hf, hb)
    print(f"halo fwd={hf} bwd={hb}: purified max diff vs (0,0) {(p - p00).abs().max():.3e}; grad rel-L2 diff vs (0,0) {((g - g00).norm() / g00.norm()):.3e}")
    if (hf, hb) == (1, 0):
        # compare tapes record by record
        for i, (ra, rb) in enumerate(zip(t00.nvae, t.nvae)):
            for j, (ta, tb) in enumerate(zip(ra, rb)):
                if torch.is_tensor(ta) and torch.is_tensor(tb) and ta.shape == tb.shape and ta.is_floating_point():
                    d = (ta.float() - tb.float()).abs().max().item()
                    s = ta.float().abs().max().item()
                    if d > 0.05 * max(s, 1e-3):
                        print(f"   tape rec {i} ({ra[0]}) field {j} shape {tuple(ta.shape)} dtype {ta.dtype}: max diff {d:.3e} (scale {s:.3e})")
L_.ga_tc_halo_enable(1)
