"""Experiment: one batch of 512 on one stream against two half batches on two streams (do the HBM-bound kernels of one half hide under the
tensor-/FMA-bound kernels of the other?).  python scripts/two_stream.py [steps]"""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from gen_adversarial_b200 import synth

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
dev = torch.device("cuda:0")
B = 512
dms = [bench.make_ours("purify", "bf16", dev) for _ in range(2)]
x, _ = synth.synthetic_batch(B, bench.RESOLUTION["purify"], bench.N_CLASSES["purify"], seed=42)
x = x.to(dev)
halves = [x[: B // 2].contiguous(), x[B // 2:].contiguous()]
streams = [torch.cuda.Stream(), torch.cuda.Stream()]


def one():
    with torch.no_grad():
        return dms[0](x)


def two():
    outs = []
    cur = torch.cuda.current_stream()
    for s in streams:
        s.wait_stream(cur)
    for i, s in enumerate(streams):
        with torch.cuda.stream(s), torch.no_grad():
            outs.append(dms[i](halves[i]))
    for s in streams:
        cur.wait_stream(s)
    return outs


def timeit(fn, name):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    print(f"{name}: {ms:.2f} ms / step of {B} images = {B / ms * 1e3:.0f} img/s", flush=True)


a = one()
b = torch.cat(two())
torch.cuda.synchronize()
print("max |one - two| logits:", (a - b).abs().max().item())
timeit(one, "one stream, batch 512")
timeit(two, "two streams, 2 x 256")
timeit(one, "one stream, batch 512")
timeit(two, "two streams, 2 x 256")
