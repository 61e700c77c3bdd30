"""GPU diagnostic: where does the bf16 error of the E4E path (config 3) come from -- encoder or generator?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gen_adversarial_b200 import ops, synth
from gen_adversarial_b200.defenses.ours import models as M

DEV = "cuda:0"
g = torch.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "e4e_gender_b2.pt"), weights_only=True)
x, noises = synth.synthetic_stylegan_inputs(g["batch"], 256, 18, seed=g["x_seed"])
ck, clf_ck = synth.make_e4e_checkpoint(1024), synth.make_resnet50_checkpoint()
dms = {}
for mode in ("fp32", "bf16"):
    clf = M.CelebaGenderClassifier(clf_ck, DEV, mode=mode)
    dm = M.E4EStyleGanDefenseModel(clf, ck, g["alphas"], g["attenuation"], g["eps"], g["blur"], DEV, mode=mode)
    dm.set_explicit_noise(noises)
    dms[mode] = dm
codes = {}
for mode, dm in dms.items():
    xin = dm._preprocessed(x.to(DEV))
    codes[mode] = dm._mix(dm._encode(xin))
print("codes: max-abs diff bf16 vs fp32", (codes["bf16"] - codes["fp32"]).abs().max().item(), "range", codes["fp32"].abs().max().item())
ref = g["purified"]
for enc in ("fp32", "bf16"):
    for dec in ("fp32", "bf16"):
        pur, _ = dms[dec]._decode(codes[enc], None)
        print(f"encoder {enc} + generator {dec}: purified max-abs err {(pur.cpu() - ref).abs().max().item():.3e}")
