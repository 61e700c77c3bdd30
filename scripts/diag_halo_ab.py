"""A/B: persistent halo 3x3 kernel against the per-tap kernel on identical inputs (GPU box)."""
import itertools
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gen_adversarial_b200 import ops, _lib
from gen_adversarial_b200._lib import ACT_NONE, ACT_SILU, ACT_RELU

DEV = "cuda:0"
L_ = _lib.lib()
worst = 0.0
for (n, h, w, cin, cout) in [(2, 32, 32, 64, 64), (2, 16, 16, 128, 128), (2, 16, 16, 128, 256), (2, 32, 32, 64, 128), (2, 32, 32, 128, 64),
                             (2, 16, 16, 256, 128), (2, 32, 32, 64, 20), (2, 16, 16, 128, 20), (3, 32, 32, 64, 64), (2, 16, 16, 64, 256),
                             (2, 32, 32, 128, 256), (5, 16, 16, 256, 512)]:
    g = torch.Generator().manual_seed(n * 1000 + cin + cout)
    Ly = ops.ConvLayer(3, 3, 1, 1, cin, cout, name="t")
    Ly.w_tc = (torch.randn(cout, 9 * cin, generator=g) / math.sqrt(9 * cin)).to(torch.bfloat16).to(DEV)
    x = torch.randn(n, h, w, cin, generator=g).to(torch.bfloat16).to(DEV)
    add = torch.randn(n, h, w, cout, generator=g).to(DEV)
    mul = torch.randn(n, h, w, cout, generator=g).to(torch.bfloat16).to(DEV)
    for bias, act, use_add, use_mul, wb, wf in itertools.product((False, True), (ACT_NONE, ACT_SILU), (False, True), (False, True), (False, True), (False, True)):
        if not (wb or wf):
            continue
        Ly.bias = (torch.randn(cout, generator=torch.Generator().manual_seed(1)) * 0.1).to(DEV) if bias else None
        Ly.post_act = act
        outs = []
        for halo in (0, 1):
            L_.ga_tc_halo_enable(halo)
            ob, of = ops.conv2d_tc(x, Ly, want_bf16=wb, want_f32=wf, add=add if use_add else None, mul=mul if use_mul else None)
            torch.cuda.synchronize()
            outs.append((ob, of))
        for k in (0, 1):
            if outs[0][k] is None:
                continue
            a, b = outs[0][k].float(), outs[1][k].float()
            d = (a - b).abs().max().item()
            tol = (2e-2 if k == 0 else 1e-4) * max(1.0, a.abs().max().item())
            worst = max(worst, d)
            if d > tol:
                bad = (a - b).abs() > tol
                idx = bad.nonzero()
                print(f"MISMATCH shape {(n, h, w, cin, cout)} bias {bias} act {act} add {use_add} mul {use_mul} out {'bf16' if k == 0 else 'f32'}: max diff {d:.3e}, "
                      f"{int(bad.sum())} elements; first at {idx[0].tolist()} rows(y) {sorted(set(idx[:, 1].tolist()))[:12]} cols(x) {sorted(set(idx[:, 2].tolist()))[:12]} "
                      f"ch {sorted(set(idx[:, 3].tolist()))[:8]}")
L_.ga_tc_halo_enable(1)
print("worst diff", worst)
