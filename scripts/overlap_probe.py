"""Do the fused decoder cell and the SE / residual kernel of ANOTHER stream run side by side on the same SMs?
stream A: reps x mbconv_fused, stream B: reps x se_residual; wall time of both against each alone.  python scripts/overlap_probe.py"""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gen_adversarial_b200 import ops
from gen_adversarial_b200._lib import ACT_NONE, ACT_SILU

DEV = "cuda:0"
g = torch.Generator(device=DEV).manual_seed(0)
reps = 20
for (n, w, c) in [(256, 32, 64), (256, 16, 128), (256, 8, 256)]:
    hidden = 6 * c
    e = ops.ConvLayer(1, 1, 1, 0, c, hidden, post_act=ACT_SILU, name="expand")
    e.w_tc = (torch.randn(hidden, c, device=DEV, generator=g) / math.sqrt(c)).bfloat16().contiguous()
    e.bias = torch.randn(hidden, device=DEV, generator=g) * 0.3
    p = ops.ConvLayer(1, 1, 1, 0, hidden, c, post_act=ACT_NONE, name="project")
    p.w_tc = (torch.randn(c, hidden, device=DEV, generator=g) / math.sqrt(hidden)).bfloat16().contiguous()
    p.bias = torch.randn(c, device=DEV, generator=g) * 0.3
    dw = ops.dw_weights_chunked(torch.randn(25, hidden, device=DEV, generator=g) / 5.0)
    db = torch.randn(hidden, device=DEV, generator=g) * 0.3
    x = torch.randn(n, w, w, c, device=DEV, generator=g).bfloat16()
    r = torch.randn(n, w, w, c, device=DEV, generator=g).bfloat16()
    skip = torch.randn(n, w, w, c, device=DEV, generator=g)
    hid = max(c // 16, 4)
    se = (torch.randn(hid, c, device=DEV, generator=g) * 0.1, torch.zeros(hid, device=DEV), torch.randn(c, hid, device=DEV, generator=g) * 0.1,
          torch.zeros(c, device=DEV))
    sums = ops.channel_sum(r)
    sa, sb = torch.cuda.Stream(), torch.cuda.Stream()

    def run(a, b):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        cur = torch.cuda.current_stream()
        e0.record()
        sa.wait_stream(cur); sb.wait_stream(cur)
        if a:
            with torch.cuda.stream(sa):
                for _ in range(reps):
                    ops.mbconv_fused(x, e, dw, db, p)
        if b:
            with torch.cuda.stream(sb):
                for _ in range(b):
                    ops.se_residual(r, sums, se, 0.1, skip, torch.float32, want_out2=True)
        cur.wait_stream(sa); cur.wait_stream(sb)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) * 1e3

    run(True, reps)
    ta = min(run(True, 0) for _ in range(3))
    tb1 = min(run(False, reps) for _ in range(3))
    nb = max(1, int(reps * ta / tb1))           # as many SE launches as fit under the cells' time
    tb = min(run(False, nb) for _ in range(3))
    tab = min(run(True, nb) for _ in range(3))
    print(f"hw={w} c={c} n={n}: {reps} cells alone {ta:.0f} us, {nb} se_residual alone {tb:.0f} us, both streams {tab:.0f} us "
          f"(serial {ta + tb:.0f}; overlap hides {100 * (ta + tb - tab) / min(ta, tb):.0f}% of the shorter)")
