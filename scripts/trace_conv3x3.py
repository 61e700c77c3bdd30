"""clock64 timeline of CTA 0 of the persistent 3x3 halo kernel (debug; GPU box).  python scripts/trace_conv3x3.py [c] [hw] [batch]"""
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gen_adversarial_b200 import ops, _lib
from gen_adversarial_b200._lib import ACT_NONE, ACT_SILU

DEV = "cuda:0"
c = int(sys.argv[1]) if len(sys.argv) > 1 else 64
hw = int(sys.argv[2]) if len(sys.argv) > 2 else 32
n = int(sys.argv[3]) if len(sys.argv) > 3 else 512
g = torch.Generator().manual_seed(0)
L = ops.ConvLayer(3, 3, 1, 1, c, c, post_act=ACT_SILU, name="t")
L.w_tc = (torch.randn(c, 9 * c, generator=g) / math.sqrt(9 * c)).to(torch.bfloat16).to(DEV)
L.bias = (torch.randn(c, generator=g) * 0.1).to(DEV)
x = torch.randn(n, hw, hw, c, generator=g).to(torch.bfloat16).to(DEV)
out = torch.empty_like(x)
for _ in range(3):
    ops.conv2d_tc(x, L, out_bf16=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    ops.conv2d_tc(x, L, out_bf16=out)
e1.record()
torch.cuda.synchronize()
us = e0.elapsed_time(e1) / 10 * 1e3
fl = 2.0 * n * hw * hw * c * 9 * c
print(f"conv3x3 c{c} hw{hw} batch {n}: {us:.1f} us per launch, {fl / us / 1e6:.0f} TFLOP/s")
buf = torch.zeros(3 * 16 * 16, dtype=torch.int64, device=DEV)
_lib.lib().ga_debug_c3_trace(buf.data_ptr())
ops.conv2d_tc(x, L, out_bf16=out)
torch.cuda.synchronize()
_lib.lib().ga_debug_c3_trace(None)
T = buf.cpu().view(3, 16, 16)
t0 = int(T[T > 0].min())
ns = 3 * (c // 64)
for it in range(6):
    if int(T[1, it, 0]) == 0:
        break
    print(f"tile {it}: producer issue " + " ".join(str(int(T[0, it, s]) - t0) for s in range(ns)))
    print(f"         mma: tmem_empty ok {int(T[1, it, 15]) - t0}; stage full at " + " ".join(str(int(T[1, it, s]) - t0) for s in range(ns)))
    print(f"         epilogue: wait start {int(T[2, it, 0]) - t0}, tmem_full ok {int(T[2, it, 1]) - t0}, store-wait0 ok {int(T[2, it, 6]) - t0}, sub0 done {int(T[2, it, 2]) - t0}, store-wait1 ok {int(T[2, it, 7]) - t0}, sub1 done {int(T[2, it, 3]) - t0}")
