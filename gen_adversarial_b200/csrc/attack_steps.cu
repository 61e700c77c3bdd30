// Fused update steps of the reference's L2 attacks, batched over images: one CTA per image does the whole update --
// per-image norms included -- so a batch of attacks costs one launch per iteration instead of ~15 elementwise / reduction launches and a
// host sync per image.
//   APGD (/root/reference/src/attacks/untargeted.py:176-193):
//       z      = x_adv + step * grad / ||grad||
//       new    = clamp(x + (z - x) * min(1, bound / ||z - x||), 0, 1)
//       u      = x_adv + (new - x_adv) * a + (x_adv - x_adv_old) * (1 - a)
//       x_adv' = clamp(x + (u - x) * min(1, bound / ||u - x||), 0, 1)          (x_adv_old' = x_adv)
//   FGSM (untargeted.py:736-745):   x_adv = clamp(x + l2 * sign(grad) / ||sign(grad)||, 0, 1)      (grad of +CE)
// Norms are per image (the reference runs one image at a time: its whole-tensor norms ARE per-image norms).  Reductions: fixed-order
// block tree, no atomics -> bit-reproducible.
#include "ga_common.cuh"

namespace ga {

constexpr int ATK_THREADS = 512;

__device__ __forceinline__ float block_sum(float v, float* s_red) {
  v = warp_sum(v);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();                       // s_red may still be read from the previous reduction
  if (lane == 0) s_red[warp] = v;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int i = 0; i < ATK_THREADS / 32; ++i) t += s_red[i];
  return t;
}

__global__ void __launch_bounds__(ATK_THREADS) apgd_l2_step_kernel(float* __restrict__ x_adv, float* __restrict__ x_adv_old,
                                                                   const float* __restrict__ grad, const float* __restrict__ x_nat,
                                                                   const float* __restrict__ step_size, float a, float bound, int chw) {
  __shared__ float s_red[ATK_THREADS / 32];
  const int64_t base = (int64_t)blockIdx.x * chw;
  float* xa = x_adv + base;
  float* xo = x_adv_old + base;
  const float* g = grad + base;
  const float* x = x_nat + base;
  const float step = step_size[blockIdx.x];
  float acc = 0.f;
  for (int i = threadIdx.x; i < chw; i += ATK_THREADS) { const float v = g[i]; acc = fmaf(v, v, acc); }
  const float gscale = step / sqrtf(block_sum(acc, s_red));
  acc = 0.f;
  for (int i = threadIdx.x; i < chw; i += ATK_THREADS) { const float d = fmaf(gscale, g[i], xa[i]) - x[i]; acc = fmaf(d, d, acc); }
  const float n1 = sqrtf(block_sum(acc, s_red));
  const float s1 = fminf(bound, n1) / n1;                                 // normalize(d) * min(bound, ||d||)
  acc = 0.f;
  for (int i = threadIdx.x; i < chw; i += ATK_THREADS) {
    const float xv = x[i], av = xa[i];
    const float nw = fminf(fmaxf(fmaf(fmaf(gscale, g[i], av) - xv, s1, xv), 0.f), 1.f);
    const float u = av + (nw - av) * a + (av - xo[i]) * (1.f - a);
    const float d = u - xv;
    acc = fmaf(d, d, acc);
  }
  const float n2 = sqrtf(block_sum(acc, s_red));
  const float s2 = fminf(bound, n2) / n2;
  for (int i = threadIdx.x; i < chw; i += ATK_THREADS) {
    const float xv = x[i], av = xa[i];
    const float nw = fminf(fmaxf(fmaf(fmaf(gscale, g[i], av) - xv, s1, xv), 0.f), 1.f);
    const float u = av + (nw - av) * a + (av - xo[i]) * (1.f - a);
    xo[i] = av;
    xa[i] = fminf(fmaxf(fmaf(u - xv, s2, xv), 0.f), 1.f);
  }
}

__global__ void __launch_bounds__(ATK_THREADS) fgsm_l2_step_kernel(const float* __restrict__ x_nat, const float* __restrict__ grad,
                                                                   float l2, float* __restrict__ out, int chw) {
  __shared__ float s_red[ATK_THREADS / 32];
  const int64_t base = (int64_t)blockIdx.x * chw;
  float acc = 0.f;
  for (int i = threadIdx.x; i < chw; i += ATK_THREADS) acc += (grad[base + i] != 0.f) ? 1.f : 0.f;       // ||sign(g)||^2 = #nonzero
  const float scale = l2 / sqrtf(block_sum(acc, s_red));
  for (int i = threadIdx.x; i < chw; i += ATK_THREADS) {
    const float gv = grad[base + i];
    const float sg = (gv > 0.f) ? 1.f : ((gv < 0.f) ? -1.f : 0.f);
    out[base + i] = fminf(fmaxf(fmaf(scale, sg, x_nat[base + i]), 0.f), 1.f);
  }
}

// x_adv = clamp(x + bound * noise / ||noise||, 0, 1): the APGD starting point (untargeted.py:129-131)
__global__ void __launch_bounds__(ATK_THREADS) l2_ball_start_kernel(const float* __restrict__ x_nat, const float* __restrict__ noise,
                                                                    float bound, float* __restrict__ out, int chw) {
  __shared__ float s_red[ATK_THREADS / 32];
  const int64_t base = (int64_t)blockIdx.x * chw;
  float acc = 0.f;
  for (int i = threadIdx.x; i < chw; i += ATK_THREADS) { const float v = noise[base + i]; acc = fmaf(v, v, acc); }
  const float scale = bound / sqrtf(block_sum(acc, s_red));
  for (int i = threadIdx.x; i < chw; i += ATK_THREADS) out[base + i] = fminf(fmaxf(fmaf(scale, noise[base + i], x_nat[base + i]), 0.f), 1.f);
}

}  // namespace ga

using namespace ga;

extern "C" int ga_apgd_l2_step(float* x_adv, float* x_adv_old, const float* grad, const float* x_nat, const float* step_size, float a,
                               float bound, int n, int chw, void* stream) {
  GA_CHECK(x_adv && x_adv_old && grad && x_nat && step_size && n >= 0 && chw > 0, "ga_apgd_l2_step: bad arguments");
  if (n == 0) return 0;
  apgd_l2_step_kernel<<<n, ATK_THREADS, 0, (cudaStream_t)stream>>>(x_adv, x_adv_old, grad, x_nat, step_size, a, bound, chw);
  GA_LAUNCH_OK();
  return 0;
}

extern "C" int ga_fgsm_l2_step(const float* x_nat, const float* grad, float l2, float* out, int n, int chw, void* stream) {
  GA_CHECK(x_nat && grad && out && n >= 0 && chw > 0, "ga_fgsm_l2_step: bad arguments");
  if (n == 0) return 0;
  fgsm_l2_step_kernel<<<n, ATK_THREADS, 0, (cudaStream_t)stream>>>(x_nat, grad, l2, out, chw);
  GA_LAUNCH_OK();
  return 0;
}

extern "C" int ga_l2_ball_start(const float* x_nat, const float* noise, float bound, float* out, int n, int chw, void* stream) {
  GA_CHECK(x_nat && noise && out && n >= 0 && chw > 0, "ga_l2_ball_start: bad arguments");
  if (n == 0) return 0;
  l2_ball_start_kernel<<<n, ATK_THREADS, 0, (cudaStream_t)stream>>>(x_nat, noise, bound, out, chw);
  GA_LAUNCH_OK();
  return 0;
}
