// Depthwise 5x5 convolution (+ folded-BN bias + SiLU), bf16 NHWC, as a persistent TMA-pipelined kernel for sm_100a
// (/root/reference/src/mlvgms_autoencoders/NVAE/modules/architecture.py:168-170; backward = the same stencil with flipped taps).
//
// Why a second kernel: ncu of the one-tile-per-CTA kernel in dwconv.cu on the attack path (profiles/r01_ncu_dwconv_pgd_v8.md):
// 64% of the warp samples wait on global loads (tile staging, then the `mul` operand in the epilogue), 1400 of the 1800 instructions
// per warp are index arithmetic of the staging loop, FMA pipe 14%.  Here
//   * the (8+4) x (TW+4) x 64-channel halo tile -- and the 8 x TW tile of the backward's `mul` operand -- are fetched by ONE TMA box
//     each (hardware zero fill outside the image = the conv's zero padding, no index arithmetic, no staging registers);
//   * a CTA is persistent over the spatial tiles of ONE 64-channel block (taps stay in registers) and double-buffers its tiles:
//     the TMA of tile i+1 is in flight while tile i is computed and stored;
//   * `up` (nearest x2 of the up cells, architecture.py:162) fetches the low-resolution halo and indexes it with >> 1.
// Per tile: 8 x TW x 64 outputs x 25 FMA on the fp32 pipe (FFMA2 issues 2 FMAs per slot but runs at half rate on B200 --
// scripts/ubench_pipes.cu: 0.5 inst/clk/SMSP -- so the floor is 128 FMA/clk/SM = 1600 cycles per 8x16x64 tile).
#include <cuda.h>
#include <stdlib.h>
#include <type_traits>
#include "ga_common.cuh"
#include "tc_ptx.cuh"

namespace ga {

constexpr int DT_TH = 8;          // output rows per tile
constexpr int DT_CH = 64;         // channels per tile = one warp of channel pairs

struct DtParams {
  const float* w;                 // [25][C] taps
  const float* bias;              // [C] or null
  __nv_bfloat16* out;             // [N][H][W][C]
  __nv_bfloat16* dact;            // taping forward: act'(pre-activation), or null
  int has_mul;                    // backward: out = act(conv) * mul (mul tile arrives through tmMul)
  int H, W, C;                    // OUTPUT height / width, channels
  int cblocks;                    // ceil(C / 64)
  int tiles_x, tiles_y;
  int64_t tiles;                  // spatial tiles = N * tiles_y * tiles_x
};

__device__ __forceinline__ uint32_t dt_lds_b32(uint32_t a) { uint32_t v; asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ float2 dt_bf2_to_f2(uint32_t v) { return make_float2(__uint_as_float(v << 16), __uint_as_float(v & 0xffff0000u)); }

template <int TW, bool UP> struct DtGeom {
  static constexpr int IN_H = UP ? (DT_TH + 4) / 2 : DT_TH + 4;          // up: hi-res rows oy0-2 .. oy0+9 -> low-res oy0/2-1 .. oy0/2+4
  static constexpr int IN_W = UP ? TW / 2 + 2 : TW + 4;
  static constexpr int IN_BYTES = IN_H * IN_W * DT_CH * 2;
  static constexpr int MUL_BYTES = DT_TH * TW * DT_CH * 2;
};

template <int TW, int ACT, bool EXTRAS, bool UP>
__global__ void __launch_bounds__(TW * 16, TW == 16 ? 2 : 4) dwconv5x5_tma_kernel(const __grid_constant__ CUtensorMap tmIn,
                                                                                   const __grid_constant__ CUtensorMap tmMul,
                                                                                   const DtParams p) {
  using G = DtGeom<TW, UP>;
  constexpr int act = ACT;
  extern __shared__ uint8_t dt_smem_raw[];
  uint8_t* base = dt_smem_raw + ((1024u - (smem_u32(dt_smem_raw) & 1023u)) & 1023u);
  uint64_t* full = reinterpret_cast<uint64_t*>(base);                      // [2]
  uint8_t* s_in = base + 1024;                                             // [2][IN_H][IN_W][64] bf16
  uint8_t* s_mul = s_in + 2 * G::IN_BYTES;                                 // [2][8][TW][64] bf16 (EXTRAS with mul only)

  const int tid = threadIdx.x;
  const int cb = blockIdx.x % p.cblocks;                                   // this CTA's 64-channel block
  const int s0 = blockIdx.x / p.cblocks;
  const int sstep = gridDim.x / p.cblocks;
  const int tiles = (int)p.tiles;
  const int per_img = p.tiles_x * p.tiles_y;
  const bool use_mul = EXTRAS && p.has_mul != 0;

  if (tid == 0) {
    tma_prefetch_desc(&tmIn);
    if (use_mul) tma_prefetch_desc(&tmMul);
    mbar_init(&full[0], 1);
    mbar_init(&full[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  auto issue = [&](int s, int b) {
    const int n = s / per_img;
    const int rem = s - n * per_img;
    const int oy0 = (rem / p.tiles_x) * DT_TH, ox0 = (rem % p.tiles_x) * TW;
    mbar_expect_tx(&full[b], G::IN_BYTES + (use_mul ? G::MUL_BYTES : 0));
    tma_load_4d(&tmIn, &full[b], s_in + b * G::IN_BYTES, cb * DT_CH, UP ? ox0 / 2 - 1 : ox0 - 2, UP ? oy0 / 2 - 1 : oy0 - 2, n);
    if (use_mul) tma_load_4d(&tmMul, &full[b], s_mul + b * G::MUL_BYTES, cb * DT_CH, ox0, oy0, n);
  };
  if (tid == 0 && s0 < tiles) issue(s0, 0);

  // ---- per-thread taps (2 channels x 25) and bias: loaded once, the CTA stays on one channel block
  const int cp = tid & 31;                       // channel pair = lane: shared reads of a warp are 128 contiguous bytes
  const int txp = tid >> 5;                      // column pair = warp
  const int c0 = cb * DT_CH + 2 * cp;
  const bool c_ok = c0 < p.C;
  float2 wr[25];
#pragma unroll
  for (int t = 0; t < 25; ++t) wr[t] = c_ok ? __ldg(reinterpret_cast<const float2*>(p.w + (size_t)t * p.C + c0)) : make_float2(0.f, 0.f);
  const float2 b2 = (c_ok && p.bias != nullptr) ? __ldg(reinterpret_cast<const float2*>(p.bias + c0)) : make_float2(0.f, 0.f);

  int i = 0;
  for (int s = s0; s < tiles; s += sstep, ++i) {
    const int b = i & 1;
    // buffer b^1 was read in iteration i-1; every thread has passed that iteration's trailing barrier
    if (tid == 0 && s + sstep < tiles) issue(s + sstep, b ^ 1);
    const int n = s / per_img;
    const int rem = s - n * per_img;
    const int oy0 = (rem / p.tiles_x) * DT_TH, ox0 = (rem % p.tiles_x) * TW;
    mbar_wait(&full[b], (uint32_t)((i >> 1) & 1));

    const uint32_t in_a = smem_u32(s_in + b * G::IN_BYTES) + (uint32_t)cp * 4u;
    float2 acc[DT_TH][2];
#pragma unroll
    for (int r = 0; r < DT_TH; ++r) { acc[r][0] = b2; acc[r][1] = b2; }
#pragma unroll
    for (int ir = 0; ir < DT_TH + 4; ++ir) {
      float2 v[6];
      if (UP) {
        float2 u[3];
#pragma unroll
        for (int d = 0; d < 3; ++d) u[d] = dt_bf2_to_f2(dt_lds_b32(in_a + (uint32_t)(((ir >> 1) * G::IN_W + txp + d) * (DT_CH * 2))));
#pragma unroll
        for (int dx = 0; dx < 6; ++dx) v[dx] = u[dx >> 1];
      } else {
#pragma unroll
        for (int dx = 0; dx < 6; ++dx) v[dx] = dt_bf2_to_f2(dt_lds_b32(in_a + (uint32_t)((ir * G::IN_W + 2 * txp + dx) * (DT_CH * 2))));
      }
#pragma unroll
      for (int r = 0; r < DT_TH; ++r) {
        const int dy = ir - r;
        if (dy < 0 || dy > 4) continue;
#pragma unroll
        for (int dx = 0; dx < 5; ++dx) {
#pragma unroll
          for (int j = 0; j < 2; ++j) acc[r][j] = ffma2(v[dx + j], wr[dy * 5 + dx], acc[r][j]);
        }
      }
    }
    if (c_ok) {
      const uint32_t mul_a = smem_u32(s_mul + b * G::MUL_BYTES) + (uint32_t)((2 * txp) * (DT_CH * 2) + cp * 4);
      // one 64-bit address per tile; rows / the second column advance it by constant strides (the generic per-element index
      // arithmetic was 600 of the 1500 instructions per tile)
      const int64_t base_off = ((((int64_t)n * p.H + oy0) * p.W + ox0 + 2 * txp) * p.C + c0) * 2;
      const int64_t rstride = (int64_t)p.W * p.C * 2;
      const bool want_d = EXTRAS && p.dact != nullptr;
      auto store_tile = [&](auto full_tag) {
        constexpr bool FULL = decltype(full_tag)::value;
        uint8_t* orow = reinterpret_cast<uint8_t*>(p.out) + base_off;
        uint8_t* drow = reinterpret_cast<uint8_t*>(p.dact) + base_off;
        const bool col1 = FULL || ox0 + 2 * txp + 1 < p.W;
        const bool col0 = FULL || ox0 + 2 * txp < p.W;
#pragma unroll
        for (int r = 0; r < DT_TH; ++r) {
          if (!FULL && oy0 + r >= p.H) break;
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            if (!(j == 0 ? col0 : col1)) continue;
            const float ax = acc[r][j].x, ay = acc[r][j].y;
            float yx, yy;
            if (act == GA_ACT_SILU) {
              // one MUFU.TANH serves SiLU and its derivative: with h = v/2, t = tanh(h): silu = h + h t, silu' = 1/2 + (t + h (1 - t^2)) / 2
              const float hx = 0.5f * ax, hy = 0.5f * ay;
              const float tx = tanh_approx(hx), ty = tanh_approx(hy);
              yx = fmaf(hx, tx, hx); yy = fmaf(hy, ty, hy);
              if (want_d) {
                const float dx_ = fmaf(0.5f, fmaf(hx, fmaf(-tx, tx, 1.0f), tx), 0.5f);
                const float dy_ = fmaf(0.5f, fmaf(hy, fmaf(-ty, ty, 1.0f), ty), 0.5f);
                *reinterpret_cast<uint32_t*>(drow + j * (p.C * 2)) = pack_bf16x2(dx_, dy_);
              }
            } else {
              yx = ax; yy = ay;
              if (want_d) *reinterpret_cast<uint32_t*>(drow + j * (p.C * 2)) = 0x3f803f80u;      // bf16 (1, 1)
            }
            if (use_mul) {
              const float2 m = dt_bf2_to_f2(dt_lds_b32(mul_a + (uint32_t)((r * TW + j) * (DT_CH * 2))));
              yx *= m.x; yy *= m.y;
            }
            *reinterpret_cast<uint32_t*>(orow + j * (p.C * 2)) = pack_bf16x2(yx, yy);
          }
          orow += rstride; drow += rstride;
        }
      };
      if (oy0 + DT_TH <= p.H && ox0 + TW <= p.W) store_tile(std::true_type{});
      else store_tile(std::false_type{});
    }
    __syncthreads();                             // all reads of buffer b done before its next TMA (issued at the top of iteration i+1)
  }
}

// ----------------------------------------------------------------------------- host side
typedef CUresult (*PFN_dtEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                      const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                      CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_dtEncodeTiled dt_encode_fn() {
  static PFN_dtEncodeTiled fn = nullptr;
  if (fn == nullptr) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_dtEncodeTiled>(ptr);
  }
  return fn;
}

static int dt_encode(CUtensorMap* tm, const ga_tensor* t, int bw, int bh) {
  PFN_dtEncodeTiled enc = dt_encode_fn();
  GA_CHECK(enc != nullptr, "cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t dims[4] = {(cuuint64_t)t->c, (cuuint64_t)t->w, (cuuint64_t)t->h, (cuuint64_t)t->n};
  cuuint64_t strides[3] = {(cuuint64_t)t->c * 2, (cuuint64_t)t->w * t->c * 2, (cuuint64_t)t->h * t->w * t->c * 2};
  cuuint32_t box[4] = {(cuuint32_t)DT_CH, (cuuint32_t)bw, (cuuint32_t)bh, 1u};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, t->data, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  GA_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(depthwise n=%d h=%d w=%d c=%d box=%d,%d) failed: %d", t->n, t->h, t->w, t->c, bw, bh,
           (int)r);
  return 0;
}

static int dt_sm_count() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)
      sms = 148;
  }
  return sms;
}

template <int TW, int ACT, bool EXTRAS, bool UP>
static int dt_launch(const ga_tensor* in, const ga_tensor* mul, const DtParams& p0, cudaStream_t s) {
  using G = DtGeom<TW, UP>;
  DtParams p = p0;
  CUtensorMap tmIn, tmMul;
  if (dt_encode(&tmIn, in, G::IN_W, G::IN_H)) return 1;
  if (mul) { if (dt_encode(&tmMul, mul, TW, DT_TH)) return 1; }
  else tmMul = tmIn;
  const int smem = 1024 /*align*/ + 1024 /*barriers*/ + 2 * G::IN_BYTES + (mul ? 2 * G::MUL_BYTES : 0);
  static int configured = 0;
  if (configured < smem) {
    GA_CUDA(cudaFuncSetAttribute(dwconv5x5_tma_kernel<TW, ACT, EXTRAS, UP>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = smem;
  }
  const int occ = TW == 16 ? 2 : 4;
  int64_t per_cb = ((int64_t)dt_sm_count() * occ) / p.cblocks;             // CTAs per channel block
  if (per_cb < 1) per_cb = 1;
  if (per_cb > p.tiles) per_cb = p.tiles;
  const unsigned grid = (unsigned)(per_cb * p.cblocks);
  dwconv5x5_tma_kernel<TW, ACT, EXTRAS, UP><<<grid, TW * 16, smem, s>>>(tmIn, tmMul, p);
  GA_LAUNCH_OK();
  return 0;
}

template <int TW, int ACT, bool EXTRAS>
static int dt_launch_up(const ga_tensor* in, const ga_tensor* mul, const DtParams& p, int up, cudaStream_t s) {
  return up ? dt_launch<TW, ACT, EXTRAS, true>(in, mul, p, s) : dt_launch<TW, ACT, EXTRAS, false>(in, mul, p, s);
}

template <int TW>
static int dt_launch_act(const ga_tensor* in, const ga_tensor* mul, const DtParams& p, int act, int up, bool extras, cudaStream_t s) {
  if (act == GA_ACT_SILU)
    return extras ? dt_launch_up<TW, GA_ACT_SILU, true>(in, mul, p, up, s) : dt_launch_up<TW, GA_ACT_SILU, false>(in, mul, p, up, s);
  return extras ? dt_launch_up<TW, GA_ACT_NONE, true>(in, mul, p, up, s) : dt_launch_up<TW, GA_ACT_NONE, false>(in, mul, p, up, s);
}

// bf16 NHWC tensors whose channel pitch and base are 16-byte aligned (what TMA needs); everything else stays on dwconv5x5_tiled_kernel
bool dwconv_tma_supported(const ga_tensor* in, const ga_tensor* mul, const ga_tensor* out, const ga_tensor* dact) {
  static int enabled = -1;
  if (enabled < 0) { const char* e = getenv("GA_DW_TMA"); enabled = e ? atoi(e) : 1; }
  if (!enabled) return false;
  if (in->dtype != GA_BF16 || out->dtype != GA_BF16 || in->c % 8 != 0) return false;
  if ((((uintptr_t)in->data) & 15) != 0 || (mul && (((uintptr_t)mul->data) & 15) != 0)) return false;
  if ((((uintptr_t)out->data) & 3) != 0 || (dact && (((uintptr_t)dact->data) & 3) != 0)) return false;
  return true;
}

int dwconv_tma_launch(const ga_tensor* in, const ga_tensor* mul, const float* weight, const float* bias, int act, int up,
                      const ga_tensor* out, const ga_tensor* dact, cudaStream_t s) {
  DtParams p;
  p.w = weight; p.bias = bias;
  p.out = (__nv_bfloat16*)out->data;
  p.dact = dact ? (__nv_bfloat16*)dact->data : nullptr;
  p.has_mul = mul ? 1 : 0;
  p.H = out->h; p.W = out->w; p.C = out->c;
  p.cblocks = cdiv(out->c, DT_CH);
  const bool extras = mul != nullptr || dact != nullptr;
  const int tw = out->w >= 16 ? 16 : 8;
  p.tiles_x = cdiv(out->w, tw); p.tiles_y = cdiv(out->h, DT_TH);
  p.tiles = (int64_t)out->n * p.tiles_x * p.tiles_y;
  if (tw == 16) return dt_launch_act<16>(in, mul, p, act, up, extras, s);
  return dt_launch_act<8>(in, mul, p, act, up, extras, s);
}

}  // namespace ga
