"""GPU: out-of-bounds guard bands.  The kernels added this round that write through computed addresses (persistent TMA depthwise with
partial tiles, depth-to-space, vectorised sum-pool / latent-mix backward, shared-memory DiscMix backward) are called through the C ABI
with their outputs placed INSIDE larger buffers filled with a sentinel; the bands before and after every output must stay untouched.
(compute-sanitizer is not available on the GPU pool.)"""
import pytest
import torch

from gen_adversarial_b200 import _lib, ops
from gen_adversarial_b200._lib import ACT_NONE, ACT_SILU
from gen_adversarial_b200.ops import gt, ptr, stream

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
BAND = 4096            # elements on each side


class Guarded:
    def __init__(self, shape, dtype):
        n = 1
        for d in shape:
            n *= d
        self.sentinel = 12345.0 if dtype == torch.float32 else 12352.0          # exactly representable in bf16
        self.big = torch.full((n + 2 * BAND,), self.sentinel, device=DEV, dtype=dtype)
        self.view = self.big[BAND:BAND + n].view(shape)
        assert self.view.is_contiguous() and self.view.data_ptr() % 16 == 0

    def check(self):
        torch.cuda.synchronize()
        assert bool((self.big[:BAND] == self.sentinel).all()) and bool((self.big[-BAND:] == self.sentinel).all()), "guard band overwritten"
        assert not bool((self.view == self.sentinel).all())                        # and the kernel did write its output
        return self.view


@pytest.mark.parametrize("n,h,w,c,up", [(2, 12, 16, 96, False), (2, 24, 40, 72, False), (3, 8, 8, 64, False), (2, 8, 8, 128, True),
                                        (1, 20, 12, 40, False)])
def test_dwconv_tma_stays_inside_its_outputs(n, h, w, c, up):
    g = torch.Generator(device=DEV).manual_seed(1)
    s = 2 if up else 1
    x = torch.randn(n, h, w, c, device=DEV, generator=g).bfloat16()
    wt = torch.randn(25, c, device=DEV, generator=g) * 0.2
    b = torch.randn(c, device=DEV, generator=g) * 0.1
    m = torch.randn(n, h * s, w * s, c, device=DEV, generator=g).bfloat16()
    L = _lib.lib()
    out, dact = Guarded((n, h * s, w * s, c), torch.bfloat16), Guarded((n, h * s, w * s, c), torch.bfloat16)
    _lib.check(L.ga_dwconv5x5_ex(gt(x), None, ptr(wt), ptr(b), ACT_SILU, int(up), gt(out.view), gt(dact.view), stream()), "dwconv taping")
    ref, refd = ops.dwconv5x5(x, wt, b, ACT_SILU, up, torch.bfloat16, want_dact=True)
    assert torch.equal(out.check(), ref) and torch.equal(dact.check(), refd)
    out2 = Guarded((n, h * s, w * s, c), torch.bfloat16)
    _lib.check(L.ga_dwconv5x5_ex(gt(x), gt(m), ptr(wt), None, ACT_NONE, int(up), gt(out2.view), None, stream()), "dwconv backward")
    assert torch.equal(out2.check(), ops.dwconv5x5(x, wt, None, ACT_NONE, up, torch.bfloat16, mul=m))
    out3 = Guarded((n, h * s, w * s, c), torch.bfloat16)
    _lib.check(L.ga_dwconv5x5_fwd(gt(x), ptr(wt), ptr(b), ACT_SILU, int(up), gt(out3.view), stream()), "dwconv plain")
    assert torch.equal(out3.check(), ref)


def test_elementwise_backward_kernels_stay_inside_their_outputs():
    g = torch.Generator(device=DEV).manual_seed(2)
    L = _lib.lib()
    # depth-to-space
    x4 = torch.randn(3, 5, 7, 4 * 12, device=DEV, generator=g)
    o = Guarded((3, 10, 14, 12), torch.float32)
    _lib.check(L.ga_depth_to_space2(gt(x4), gt(o.view), stream()), "d2s")
    assert torch.equal(o.check(), ops.depth_to_space2(x4))
    # sum-pool (vector path, bf16 in / out, with mul)
    xb = torch.randn(3, 12, 20, 24, device=DEV, generator=g).bfloat16()
    mb = torch.randn(3, 6, 10, 24, device=DEV, generator=g).bfloat16()
    o = Guarded((3, 6, 10, 24), torch.bfloat16)
    _lib.check(L.ga_sumpool2x2(gt(xb), gt(mb), gt(o.view), stream()), "sumpool")
    assert torch.equal(o.check(), ops.sumpool2x2(xb, torch.bfloat16, mul=mb))
    # DiscMix mean backward: ragged pixel count, channel padding
    logits = torch.randn(3, 7, 9, 100, device=DEV, generator=g) * 1.5
    gc = torch.randn(3, 7, 9, 3, device=DEV, generator=g)
    o = Guarded((3, 7, 9, 104), torch.float32)
    _lib.check(L.ga_discmix_mean_bwd(gt(logits), 10, None, gt(gc), gt(o.view), stream()), "discmix bwd")
    assert torch.equal(o.check(), ops.discmix_mean_bwd(logits, 10, None, gc, pad_to=104))
    # SE backward (vector path)
    r = torch.randn(2, 16, 16, 128, device=DEV, generator=g).bfloat16()
    go = torch.randn(2, 16, 16, 128, device=DEV, generator=g)
    se = (torch.randn(8, 128, device=DEV, generator=g) * 0.3, torch.randn(8, device=DEV, generator=g) * 0.1,
          torch.randn(128, 8, device=DEV, generator=g) * 0.3, torch.randn(128, device=DEV, generator=g) * 0.1)
    sums = ops.channel_sum(r)
    dots = torch.empty_like(sums)
    o = Guarded((2, 16, 16, 128), torch.bfloat16)
    _lib.check(L.ga_se_residual_bwd(gt(go), gt(r), ptr(sums), ptr(dots), ptr(se[0]), ptr(se[1]), ptr(se[2]), ptr(se[3]), 8, 0.1, gt(o.view),
                                    stream()), "se bwd")
    assert torch.equal(o.check(), ops.se_residual_bwd(go, r, sums, se, 0.1, torch.bfloat16))
