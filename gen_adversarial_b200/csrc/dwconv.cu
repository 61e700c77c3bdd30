// Depthwise 5x5 convolution (+ folded-BN bias + SiLU) of the NVAE decoder cell
// (/root/reference/src/mlvgms_autoencoders/NVAE/modules/architecture.py:168-170; hidden = 6C channels).
//
// HBM/FMA balanced op (25 MAC per output element): a CTA stages an (8+4) x (TW+4) pixel halo tile of CH channels
// in shared memory once (16-byte vector loads, zero fill = the conv's zero padding), then every thread produces an
// 8-row x 2-column strip for TWO adjacent channels with a sliding window: 12 x 6 shared loads (bf16x2 words) feed
// 8 x 2 x 25 x 2 FMAs (11 FMA per LDS.32 and per bf16->fp32 unpack: the half-rate ALU pipe was the limiter of the
// 1-column version, ncu: ALU 60% / FMA 28%), weights live in registers.  Lanes run along channels (64 bf16
// channels per CTA = one warp wide), so shared reads are conflict-free and global accesses are 128-byte segments.
// `up` reads the input through the nearest x2 up-sampling of the up cells (architecture.py:162) without
// materialising it.
#include "ga_common.cuh"

namespace ga {

constexpr int DW_TH = 8;

template <typename T> struct DwTraits;
template <> struct DwTraits<__nv_bfloat16> { static constexpr int CH = 64; static constexpr int VEC = 8; };
template <> struct DwTraits<float> { static constexpr int CH = 32; static constexpr int VEC = 4; };

template <typename T> __device__ __forceinline__ float2 lds2(const T* p);
template <> __device__ __forceinline__ float2 lds2<float>(const float* p) { return *reinterpret_cast<const float2*>(p); }
template <> __device__ __forceinline__ float2 lds2<__nv_bfloat16>(const __nv_bfloat16* p) {
  return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(p));
}
template <typename T> __device__ __forceinline__ void stg2(T* p, float a, float b);
template <> __device__ __forceinline__ void stg2<float>(float* p, float a, float b) { *reinterpret_cast<float2*>(p) = make_float2(a, b); }
template <> __device__ __forceinline__ void stg2<__nv_bfloat16>(__nv_bfloat16* p, float a, float b) {
  *reinterpret_cast<__nv_bfloat162*>(p) = __floats2bfloat162_rn(a, b);
}

template <typename T> __device__ __forceinline__ uint4 mul_vec(uint4 a, uint4 b);
template <> __device__ __forceinline__ uint4 mul_vec<float>(uint4 a, uint4 b) {
  float4 x = *reinterpret_cast<float4*>(&a), y = *reinterpret_cast<float4*>(&b);
  float4 r = make_float4(x.x * y.x, x.y * y.y, x.z * y.z, x.w * y.w);
  return *reinterpret_cast<uint4*>(&r);
}
template <> __device__ __forceinline__ uint4 mul_vec<__nv_bfloat16>(uint4 a, uint4 b) {
  uint4 r;
  const uint32_t* pa = &a.x; const uint32_t* pb = &b.x; uint32_t* pr = &r.x;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 fa = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(pa + i));
    float2 fb = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(pb + i));
    pr[i] = pack_bf16x2(fa.x * fb.x, fa.y * fb.y);
  }
  return r;
}

// ACT (compile time): GA_ACT_NONE or GA_ACT_SILU.  EXTRAS: the backward / taping variant (mul, dact pointers live).
template <typename TIn, typename TOut, int TW, int ACT, bool EXTRAS>
__global__ void __launch_bounds__(TW * DwTraits<TIn>::CH / 4) dwconv5x5_tiled_kernel(
    const TIn* __restrict__ in, const TOut* __restrict__ mul, const float* __restrict__ w, const float* __restrict__ bias,
    int up, int H, int W, int C, int tiles_x, TOut* __restrict__ out, TOut* __restrict__ dact) {
  constexpr int act = ACT;
  constexpr int CH = DwTraits<TIn>::CH, VEC = DwTraits<TIn>::VEC;
  constexpr int NT = TW * CH / 4;      // (TW/2 column pairs) x (CH/2 channel pairs)
  constexpr int SH = DW_TH + 4, SW = TW + 4;
  __shared__ __align__(16) TIn s_in[SH][SW][CH];

  const int tid = threadIdx.x;
  const int n = blockIdx.z;
  const int c_blk = blockIdx.y * CH;
  const int oy0 = (blockIdx.x / tiles_x) * DW_TH, ox0 = (blockIdx.x % tiles_x) * TW;
  const int Hi = up ? H >> 1 : H, Wi = up ? W >> 1 : W;

  // ---- stage the halo tile (zero outside the image / beyond C).  All of a thread's global loads are issued before the
  // first shared store (fixed trip count, fully unrolled): with a rolled loop each thread had ONE load in flight and the
  // CTA spent ~8 us waiting on DRAM latency 7 times in a row.
  constexpr int VPP = CH / VEC;                      // vectors per pixel
  constexpr int TOTAL = SH * SW * VPP;
  constexpr int ITER = (TOTAL + NT - 1) / NT;
  uint4 vals[ITER];
#pragma unroll
  for (int it = 0; it < ITER; ++it) {
    const int i = tid + it * NT;
    const int v = i % VPP;
    const int px = (i / VPP) % SW;
    const int py = i / (VPP * SW);
    const int iy = oy0 + py - 2, ix = ox0 + px - 2;
    const int c = c_blk + v * VEC;
    vals[it] = make_uint4(0u, 0u, 0u, 0u);
    if (i < TOTAL && iy >= 0 && iy < H && ix >= 0 && ix < W && c < C) {
      const int sy = up ? iy >> 1 : iy, sx = up ? ix >> 1 : ix;
      vals[it] = __ldg(reinterpret_cast<const uint4*>(in + (((int64_t)n * Hi + sy) * Wi + sx) * C + c));
    }
  }
#pragma unroll
  for (int it = 0; it < ITER; ++it) {
    const int i = tid + it * NT;
    if (i < TOTAL) {
      const int v = i % VPP;
      const int px = (i / VPP) % SW;
      const int py = i / (VPP * SW);
      *reinterpret_cast<uint4*>(&s_in[py][px][v * VEC]) = vals[it];
    }
  }
  // ---- per-thread weights (2 channels x 25 taps) and bias
  const int cp = tid % (CH / 2);
  const int txp = tid / (CH / 2);               // column pair
  const int c0 = c_blk + 2 * cp;
  const bool c_ok = c0 < C;
  float2 wr[25];
#pragma unroll
  for (int t = 0; t < 25; ++t) wr[t] = c_ok ? __ldg(reinterpret_cast<const float2*>(w + t * C + c0)) : make_float2(0.f, 0.f);
  float2 b2 = (c_ok && bias != nullptr) ? __ldg(reinterpret_cast<const float2*>(bias + c0)) : make_float2(0.f, 0.f);
  float2 acc[DW_TH][2];
#pragma unroll
  for (int r = 0; r < DW_TH; ++r) { acc[r][0] = b2; acc[r][1] = b2; }
  __syncthreads();

#pragma unroll
  for (int ir = 0; ir < SH; ++ir) {
    float2 v[6];
#pragma unroll
    for (int dx = 0; dx < 6; ++dx) v[dx] = lds2<TIn>(&s_in[ir][2 * txp + dx][2 * cp]);
#pragma unroll
    for (int r = 0; r < DW_TH; ++r) {
      const int dy = ir - r;
      if (dy < 0 || dy > 4) continue;
#pragma unroll
      for (int dx = 0; dx < 5; ++dx) {
#pragma unroll
        for (int j = 0; j < 2; ++j) acc[r][j] = ffma2(v[dx + j], wr[dy * 5 + dx], acc[r][j]);   // 2 channels per FFMA2
      }
    }
  }
  if (!c_ok) return;
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const int ox = ox0 + 2 * txp + j;
    if (ox >= W) continue;
#pragma unroll
    for (int r = 0; r < DW_TH; ++r) {
      const int oy = oy0 + r;
      if (oy >= H) break;
      const int64_t o = (((int64_t)n * H + oy) * W + ox) * C + c0;
      const float ax = acc[r][j].x, ay = acc[r][j].y;
      if (EXTRAS) {
        if (dact != nullptr) stg2<TOut>(dact + o, act_grad(ax, act), act_grad(ay, act));
        if (mul != nullptr) {                      // backward: result times the saved derivative of the producer's activation
          const float2 m = lds2<TOut>(mul + o);
          stg2<TOut>(out + o, apply_act(ax, act) * m.x, apply_act(ay, act) * m.y);
          continue;
        }
      }
      if (sizeof(TOut) == 2)   // bf16 output: fast-math activation (error far below bf16 rounding)
        stg2<TOut>(out + o, apply_act_fast(ax, act), apply_act_fast(ay, act));
      else
        stg2<TOut>(out + o, apply_act(ax, act), apply_act(ay, act));
    }
  }
}

template <typename TIn, typename TOut, int ACT, bool EXTRAS>
static int launch_dw2(const ga_tensor* in, const void* mul, const float* weight, const float* bias, int up, const ga_tensor* out,
                      void* dact, cudaStream_t s) {
  constexpr int CH = DwTraits<TIn>::CH;
  const int H = out->h, W = out->w, C = out->c;
  const int cblocks = cdiv(C, CH);
  if (W >= 16) {
    const int tiles_x = cdiv(W, 16);
    dim3 grid(tiles_x * cdiv(H, DW_TH), cblocks, out->n);
    dwconv5x5_tiled_kernel<TIn, TOut, 16, ACT, EXTRAS><<<grid, 16 * CH / 4, 0, s>>>(
        (const TIn*)in->data, (const TOut*)mul, weight, bias, up, H, W, C, tiles_x, (TOut*)out->data, (TOut*)dact);
  } else {
    const int tiles_x = cdiv(W, 8);
    dim3 grid(tiles_x * cdiv(H, DW_TH), cblocks, out->n);
    dwconv5x5_tiled_kernel<TIn, TOut, 8, ACT, EXTRAS><<<grid, 8 * CH / 4, 0, s>>>(
        (const TIn*)in->data, (const TOut*)mul, weight, bias, up, H, W, C, tiles_x, (TOut*)out->data, (TOut*)dact);
  }
  GA_LAUNCH_OK();
  return 0;
}

template <typename TIn, typename TOut>
static int launch_dw(const ga_tensor* in, const void* mul, const float* weight, const float* bias, int act, int up,
                     const ga_tensor* out, void* dact, cudaStream_t s) {
  const bool extras = mul != nullptr || dact != nullptr;
  if (act == GA_ACT_SILU)
    return extras ? launch_dw2<TIn, TOut, GA_ACT_SILU, true>(in, mul, weight, bias, up, out, dact, s)
                  : launch_dw2<TIn, TOut, GA_ACT_SILU, false>(in, mul, weight, bias, up, out, dact, s);
  return extras ? launch_dw2<TIn, TOut, GA_ACT_NONE, true>(in, mul, weight, bias, up, out, dact, s)
                : launch_dw2<TIn, TOut, GA_ACT_NONE, false>(in, mul, weight, bias, up, out, dact, s);
}

// persistent TMA-pipelined bf16 kernel (dwconv_tma.cu)
bool dwconv_tma_supported(const ga_tensor* in, const ga_tensor* mul, const ga_tensor* out, const ga_tensor* dact);
int dwconv_tma_launch(const ga_tensor* in, const ga_tensor* mul, const float* weight, const float* bias, int act, int up,
                      const ga_tensor* out, const ga_tensor* dact, cudaStream_t s);

}  // namespace ga

using namespace ga;

static int dw_dispatch(const ga_tensor* in, const ga_tensor* mul, const float* weight, const float* bias, int act, int up,
                       const ga_tensor* out, const ga_tensor* dact, void* stream, const char* who) {
  GA_CHECK(in && weight && out, "%s: null argument", who);
  GA_CHECK(in->c == out->c && in->n == out->n, "%s: channels / batch must match", who);
  GA_CHECK(in->c % (in->dtype == GA_BF16 ? 8 : 4) == 0, "%s: channels must be a multiple of 8 (bf16) / 4 (fp32)", who);
  GA_CHECK(up ? (out->h == 2 * in->h && out->w == 2 * in->w) : (out->h == in->h && out->w == in->w), "%s: shape mismatch", who);
  GA_CHECK(out->n <= 65535, "%s: batch too large for grid.z", who);
  GA_CHECK(act == GA_ACT_NONE || act == GA_ACT_SILU, "%s: activation must be none or SiLU", who);
  GA_CHECK(in->dtype == out->dtype, "%s: input and output dtypes must match", who);
  GA_CHECK(!mul || (same_shape(mul, out) && mul->dtype == out->dtype), "%s: mul must match the output's shape and dtype", who);
  GA_CHECK(!dact || (same_shape(dact, out) && dact->dtype == out->dtype), "%s: dact must match the output's shape and dtype", who);
  if (numel(out) == 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  const void* m = mul ? mul->data : nullptr;
  void* da = dact ? dact->data : nullptr;
  if (in->dtype == GA_F32) return launch_dw<float, float>(in, m, weight, bias, act, up, out, da, s);
  if (dwconv_tma_supported(in, mul, out, dact)) return dwconv_tma_launch(in, mul, weight, bias, act, up, out, dact, s);
  return launch_dw<__nv_bfloat16, __nv_bfloat16>(in, m, weight, bias, act, up, out, da, s);
}

extern "C" int ga_dwconv5x5_fwd(const ga_tensor* in, const float* weight, const float* bias, int act, int up,
                                const ga_tensor* out, void* stream) {
  return dw_dispatch(in, nullptr, weight, bias, act, up, out, nullptr, stream, "ga_dwconv5x5_fwd");
}

extern "C" int ga_dwconv5x5_ex(const ga_tensor* in, const ga_tensor* mul, const float* weight, const float* bias, int act, int up,
                               const ga_tensor* out, const ga_tensor* dact, void* stream) {
  return dw_dispatch(in, mul, weight, bias, act, up, out, dact, stream, "ga_dwconv5x5_ex");
}
