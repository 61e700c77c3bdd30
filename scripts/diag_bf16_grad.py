"""diagnostic: where does the bf16 input-gradient error come from? (purifier-only vs classifier-only)"""
import math, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gen_adversarial_b200 import synth, ops, autograd as ga
from gen_adversarial_b200.nvae_engine import NvaeEngine
from gen_adversarial_b200.vgg_engine import Vgg11Engine
from gen_adversarial_b200.nvae_spec import *
from oracle import nvae_ref
DEV = "cuda:0"
spec = NvaeSpec(NVAE_C32_CONFIG, NVAE_C32_RESOLUTION)
nv = synth.make_nvae_state_dict(seed=0)
alphas = [0.7 * 0.5 * (1 - math.cos(math.pi * i / 24)) for i in range(1, 25)]
x, _ = synth.synthetic_batch(2, seed=5)
noises = synth.synthetic_noise(spec, 2, seed=6)
wgt = torch.randn(2, 3, 64, 64, generator=torch.Generator().manual_seed(0))
xo = x.clone().requires_grad_(True)
_, pur = nvae_ref.defense_call(nv, spec, None, xo, alphas, noises, 2.0, True)
g_ref, = torch.autograd.grad((pur * wgt).sum(), [xo])

def met(name, g, r):
    g = g.cpu().float(); r = r.float()
    print(f"{name}: rel-L2 {((g - r).norm() / r.norm()).item():.3e} cos {torch.nn.functional.cosine_similarity(g.flatten(), r.flatten(), dim=0).item():.6f} "
          f"sign {(torch.sign(g) == torch.sign(r)).float().mean().item():.4f}")

for mode in ("fp32", "bf16"):
    eng = NvaeEngine(nv, spec, DEV, mode)
    tape = ga.Tape()
    xin, pre = ops.preprocess(x.to(DEV), noises[0].to(DEV), 2.0, True, eng.adt, save_pre=True)
    a_dev = torch.tensor(alphas, device=DEV)
    p2, _ = eng.purify(xin, a_dev, [n.to(DEV) for n in noises[1:]], tape=tape)
    gx = ops.preprocess_bwd(eng.backward(tape.nvae, wgt.to(DEV), None), pre, True)
    met(f"purifier-only grad [{mode}]", gx, g_ref)
# classifier only
vg = synth.make_vgg11_state_dict(100, seed=1)
model = nvae_ref.build_vgg11(vg, 100)
xi = pur.detach().clone()
y = torch.tensor([3, 41])
xr = xi.clone().requires_grad_(True)
loss = torch.nn.functional.cross_entropy(nvae_ref.classify(model, xr), y)
gc_ref, = torch.autograd.grad(loss, [xr])
for mode in ("fp32", "bf16"):
    eng = Vgg11Engine(vg, DEV, mode)
    tape = []
    xin = ops.nchw_to_nhwc(xi.to(DEV), eng.adt, 2.0, -1.0)
    logits = eng.forward(xin, tape=tape)
    _, dl, _ = ops.softmax_xent(logits, y.to(DEV))
    g = eng.backward(tape, dl).permute(0, 3, 1, 2) * 2.0
    met(f"classifier-only grad [{mode}]", g, gc_ref)
