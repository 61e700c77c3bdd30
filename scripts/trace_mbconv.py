"""Per-role clock64 timeline of the fused decoder-cell kernel at 32x32 (debug; GPU box).  python scripts/trace_mbconv.py [batch]"""
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gen_adversarial_b200 import ops, _lib
from gen_adversarial_b200._lib import ACT_NONE, ACT_SILU

DEV = "cuda:0"
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
c, hidden, w = 64, 384, 32
g = torch.Generator().manual_seed(0)
e = ops.ConvLayer(1, 1, 1, 0, c, hidden, post_act=ACT_SILU, name="expand")
e.w_tc = (torch.randn(hidden, c, generator=g) / math.sqrt(c)).to(torch.bfloat16).to(DEV)
e.bias = (torch.randn(hidden, generator=g) * 0.3).to(DEV)
p = ops.ConvLayer(1, 1, 1, 0, hidden, c, post_act=ACT_NONE, name="project")
p.w_tc = (torch.randn(c, hidden, generator=g) / math.sqrt(hidden)).to(torch.bfloat16).to(DEV)
p.bias = (torch.randn(c, generator=g) * 0.3).to(DEV)
dw = ops.dw_weights_chunked((torch.randn(25, hidden, generator=g) / 5.0).to(DEV))
db = (torch.randn(hidden, generator=g) * 0.3).to(DEV)
x = torch.randn(n, w, w, c, generator=g).to(torch.bfloat16).to(DEV)
for _ in range(3):
    ops.mbconv_fused(x, e, dw, db, p)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    ops.mbconv_fused(x, e, dw, db, p)
e1.record()
torch.cuda.synchronize()
print(f"plain kernel: {e0.elapsed_time(e1) / 5 * 1e3:.1f} us per launch (batch {n})")
buf = torch.zeros(8 * 13 * 24 * 8, dtype=torch.int64, device=DEV)
_lib.lib().ga_debug_mbconv_trace(buf.data_ptr())
ops.mbconv_fused(x, e, dw, db, p)
torch.cuda.synchronize()
e0.record()
ops.mbconv_fused(x, e, dw, db, p)
e1.record()
torch.cuda.synchronize()
print(f"traced kernel: {e0.elapsed_time(e1) * 1e3:.1f} us")
_lib.lib().ga_debug_mbconv_trace(None)
t = buf.cpu().view(8, 13, 24, 8)
nch = hidden // 64
for cta in (0, 5):
    T = t[cta]
    t0 = int(T[:, 0, 7].min())
    print(f"\n=== CTA slot {cta}: start 0, end {int(T[:, 1, 7].max()) - t0} clk; x_full at {int(T[0, 0, 6]) - t0}")
    print("control  k: expand-issued | project: wait-a2 start, a2 full, issued")
    for k in range(nch):
        print(f"  k={k}: {int(T[0, k, 0]) - t0:7d} | {int(T[0, k, 1]) - t0:7d} {int(T[0, k, 2]) - t0:7d} {int(T[0, k, 3]) - t0:7d}")
    for wname, wid in (("act warp 1", 1), ("act warp 4", 4)):
        print(f"{wname} k: loop top, exp_full ok, h_empty ok, done")
        for k in range(nch):
            print(f"  k={k}: " + " ".join(f"{int(T[wid, k, ev]) - t0:7d}" for ev in range(4)))
        print(f"  epilogue: wait proj {int(T[wid, 0, 4]) - t0}, proj ok {int(T[wid, 0, 5]) - t0}, exit {int(T[wid, 1, 7]) - t0}")
    for wid in (5, 8, 12):
        print(f"dw warp {wid} k: top, taps ok, h_full ok, pass0 accumulated, a2_empty ok, done")
        for k in range(nch):
            print(f"  k={k}: " + " ".join(f"{int(T[wid, k, ev]) - t0:7d}" for ev in range(6)))
