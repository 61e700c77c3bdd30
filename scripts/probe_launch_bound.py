"""GPU probe: is the NVAE purify step launch-bound?  host enqueue time vs device time vs CUDA-graph replay."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gen_adversarial_b200 import ops, synth
from gen_adversarial_b200.defenses.ours.models import NVAEDefenseModel, CelebaIdentityClassifier
import bench

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
cfg = bench.LEARNED_BLUR_IDS
clf = CelebaIdentityClassifier({"state_dict": synth.make_vgg11_state_dict(100, seed=1, device=str(dev))}, dev, mode="bf16")
dm = NVAEDefenseModel(clf, synth.make_nvae_checkpoint(seed=0), cfg["interpolation_alphas"], cfg["alpha_attenuation"], cfg["initial_noise_eps"],
                      cfg["gaussian_blur_input"], dev, mode="bf16").eval()
dm.noise_seed = 123
x = synth.synthetic_batch(B, seed=42)[0].to(dev)
for _ in range(3):
    with torch.no_grad():
        out = dm(x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter(); e0.record()
for _ in range(3):
    with torch.no_grad():
        out = dm(x)
e1.record(); t_host = (time.perf_counter() - t0) / 3
torch.cuda.synchronize()
print(f"B={B}: host enqueue {1e3 * t_host:.1f} ms/step, device {e0.elapsed_time(e1) / 3:.1f} ms/step")
# CUDA graph capture of the whole call
try:
    dm._alphas_device()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        with torch.no_grad():
            dm(x)
    torch.cuda.current_stream().wait_stream(s)
    with torch.cuda.graph(g):
        with torch.no_grad():
            out_g = dm(x)
    torch.cuda.synchronize()
    for _ in range(2):
        g.replay()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(5):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    print(f"B={B}: CUDA-graph replay {e0.elapsed_time(e1) / 5:.1f} ms/step; max |graph - eager| logits {(out_g - out).abs().max().item():.3e}")
except Exception as ex:
    print("graph capture failed:", repr(ex)[:500])
