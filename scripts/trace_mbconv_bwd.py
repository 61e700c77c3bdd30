"""Per-role clock64 timeline of the fused BACKWARD decoder cell at 32x32 (debug; GPU box).  python scripts/trace_mbconv_bwd.py [batch]"""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gen_adversarial_b200 import ops, _lib
from gen_adversarial_b200._lib import ACT_NONE

DEV = "cuda:0"
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
c, hidden, w = 64, 384, 32
g = torch.Generator(device=DEV).manual_seed(0)
p_d = ops.ConvLayer(1, 1, 1, 0, c, hidden, post_act=ACT_NONE, name="project_dgrad")
p_d.w_tc = (torch.randn(hidden, c, device=DEV, generator=g) / math.sqrt(hidden)).bfloat16().contiguous()
e_d = ops.ConvLayer(1, 1, 1, 0, hidden, c, post_act=ACT_NONE, name="expand_dgrad")
e_d.w_tc = (torch.randn(c, hidden, device=DEV, generator=g) / math.sqrt(c)).bfloat16().contiguous()
dwT = ops.dw_weights_chunked(torch.randn(25, hidden, device=DEV, generator=g) / 5.0)
gr = (torch.randn(n, w, w, c, device=DEV, generator=g) * 1e-3).bfloat16()
dd = torch.rand(n, w, w, hidden, device=DEV, generator=g).bfloat16()
de = torch.rand(n, w, w, hidden, device=DEV, generator=g).bfloat16()
add = torch.randn(n, w, w, c, device=DEV, generator=g) * 1e-3
for _ in range(3):
    ops.mbconv_fused_bwd(gr, p_d, dwT, dd, de, e_d, add=add)
torch.cuda.synchronize()
buf = torch.zeros(8 * 13 * 24 * 8, dtype=torch.int64, device=DEV)
_lib.lib().ga_debug_mbconv_trace(buf.data_ptr())
ops.mbconv_fused_bwd(gr, p_d, dwT, dd, de, e_d, add=add)
torch.cuda.synchronize()
_lib.lib().ga_debug_mbconv_trace(None)
T = buf.cpu().view(8, 13, 24, 8)[0]
t0 = int(T[:, 0, 7].min())
nch = hidden // 64
print(f"=== CTA 0 (14 tiles): end {int(T[:, 1, 7].max()) - t0} clk; x_full at {int(T[0, 0, 6]) - t0}")
print("control  g: expand-issued | project: wait-a2 start, a2 full, issued")
for k in range(2 * nch):
    print(f"  g={k}: {int(T[0, k, 0]) - t0:7d} | {int(T[0, k, 1]) - t0:7d} {int(T[0, k, 2]) - t0:7d} {int(T[0, k, 3]) - t0:7d}")
for wname, wid in (("act warp 1", 1), ("act warp 4", 4)):
    print(f"{wname} g: loop top, exp_full ok, h_empty ok, done")
    for k in range(2 * nch):
        print(f"  g={k}: " + " ".join(f"{int(T[wid, k, ev]) - t0:7d}" for ev in range(4)))
    print(f"  epilogue tile 0: wait proj {int(T[wid, 0, 4]) - t0}, proj ok {int(T[wid, 0, 5]) - t0}; tile 1: {int(T[wid, 1, 4]) - t0}, {int(T[wid, 1, 5]) - t0}")
for wid in (5, 8, 12):
    print(f"dw warp {wid} g: top, taps ok, h_full ok, pass0 accumulated, a2_empty ok, done")
    for k in range(2 * nch):
        print(f"  g={k}: " + " ".join(f"{int(T[wid, k, ev]) - t0:7d}" for ev in range(6)))
