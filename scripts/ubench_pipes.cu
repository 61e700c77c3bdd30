// Micro-benchmark of the SIMT pipes the depthwise / activation stages depend on (B200, sm_100a):
// warp-instructions per cycle per SM sub-partition for FFMA (3 registers), FFMA2 (packed pair), FMUL, MUFU.TANH, LDS.32 and the
// bf16x2 -> fp32 unpack pair, at 1/2/4/8 warps per sub-partition.  Build + run:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ubench scripts/ubench_pipes.cu && /tmp/ubench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int ITERS = 4096;
constexpr int NACC = 16;

__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  unsigned long long ra, rb = *reinterpret_cast<unsigned long long*>(&b), rc = *reinterpret_cast<unsigned long long*>(&c);
  unsigned long long rd;
  ra = *reinterpret_cast<unsigned long long*>(&a);
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
  return *reinterpret_cast<float2*>(&rd);
}

template <int MODE>
__global__ void k(float* out, const float* in, long long* cycles) {
  __shared__ float sm[4096];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = in[i & 255];
  __syncthreads();
  float a[NACC];
  float2 a2[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) { a[i] = in[(threadIdx.x + i) & 255]; a2[i] = make_float2(a[i], a[i] * 0.5f); }
  const float w0 = in[threadIdx.x & 127], w1 = in[(threadIdx.x + 7) & 127];
  const float2 w2 = make_float2(w0, w1), v2 = make_float2(w1, w0);
  uint32_t saddr = (threadIdx.x & 31) * 4;
  uint32_t h[NACC], g[NACC];
  const uint32_t hw = 0x3c003800u ^ (threadIdx.x & 1), hv = 0x38003c00u ^ (threadIdx.x & 2);      // fp16 (1.0, 0.5) / (0.5, 1.0)
#pragma unroll
  for (int i = 0; i < NACC; ++i) { h[i] = 0x3c003c00u + i; g[i] = threadIdx.x + i; }
  long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) {
      if (MODE == 0) a[i] = fmaf(a[i], w0, w1);                                  // FFMA, 3 distinct registers
      if (MODE == 1) a2[i] = ffma2(w2, v2, a2[i]);                               // FFMA2
      if (MODE == 2) a[i] = a[i] * w0;                                           // FMUL
      if (MODE == 3) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(a[i]));       // MUFU.TANH
      if (MODE == 4) { a[i] += sm[(saddr >> 2) + i * 32 + (it & 1) * 512]; }     // LDS.32 + FADD
      if (MODE == 5) {                                                          // unpack pair: SHL + LOP
        uint32_t v = __float_as_uint(a[i]);
        a[i] = __uint_as_float(v << 16) + __uint_as_float(v & 0xffff0000u);
      }
      if (MODE == 6) a[i] = fmaf(a[i], 1.0009765625f, w1);                       // FFMA immediate form
      if (MODE == 7) {                                                          // FFMA2 with per-iteration distinct operands (window)
        a2[i] = ffma2(a2[(i + 1) & (NACC - 1)], w2, a2[i]);
      }
      if (MODE == 8) asm volatile("fma.rn.f16x2 %0, %1, %2, %0;" : "+r"(h[i]) : "r"(hw), "r"(hv));                   // HFMA2 (fp16 pair)
      if (MODE == 9) asm volatile("fma.rn.bf16x2 %0, %1, %2, %0;" : "+r"(h[i]) : "r"(hw), "r"(hv));                  // HFMA2.BF16
      if (MODE == 10) asm volatile("fma.rn.f16x2 %0, %1, %2, %0;" : "+r"(h[i]) : "r"(h[(i + 1) & (NACC - 1)]), "r"(hv));   // HFMA2, rotating operands
      if (MODE == 11) {                                                         // FFMA2 + one ALU op each: does the ALU op issue in FFMA2's 2nd cycle?
        a2[i] = ffma2(w2, v2, a2[i]);
        h[i] = (h[i] << 3) ^ hw;
      }
      if (MODE == 13) { a2[i] = ffma2(w2, v2, a2[i]); a[i] = fmaf(a[i], w0, w1); }                  // FFMA2 + FFMA: does the scalar FMA ride a second pipe?
      if (MODE == 14) {                                                                             // HFMA2 + FFMA
        asm volatile("fma.rn.f16x2 %0, %1, %2, %0;" : "+r"(h[i]) : "r"(hw), "r"(hv));
        a[i] = fmaf(a[i], w0, w1);
      }
      if (MODE == 15) {                                                                             // HFMA2 + LDS.32 (the depthwise loop's pair)
        asm volatile("fma.rn.f16x2 %0, %1, %2, %0;" : "+r"(h[i]) : "r"(hw), "r"(__float_as_uint(sm[(saddr >> 2) + i * 32 + (it & 1) * 512])));
      }
      if (MODE == 16) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));      // MUFU.EX2
      if (MODE == 17) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(a[i]));      // MUFU.RCP
      if (MODE == 12) {                                                         // HFMA2 + one ALU op each
        asm volatile("fma.rn.f16x2 %0, %1, %2, %0;" : "+r"(h[i]) : "r"(hw), "r"(hv));
        g[i] = (g[i] << 3) ^ hw;
      }
    }
  }
  long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += a[i] + a2[i].x + a2[i].y + __uint_as_float(h[i] ^ g[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, float* out, float* in, long long* cyc) {
  printf("%-28s", name);
  for (int warps_per_smsp : {1, 2, 4, 8}) {
    const int threads = warps_per_smsp * 4 * 32;
    k<MODE><<<148, threads>>>(out, in, cyc);
    cudaDeviceSynchronize();
    k<MODE><<<148, threads>>>(out, in, cyc);
    cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double c = 0;
    for (int i = 0; i < 148; ++i) c += h[i];
    c /= 148;
    const double inst = (double)ITERS * NACC * warps_per_smsp;      // warp-instructions per sub-partition
    printf("  w/smsp=%d: %.3f inst/clk", warps_per_smsp, inst / c);
  }
  printf("\n");
}

int main() {
  float *out, *in;
  long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4);
  cudaMalloc(&in, 4096 * 4);
  cudaMalloc(&cyc, 148 * 8);
  float h[4096];
  for (int i = 0; i < 4096; ++i) h[i] = 0.001f * (i % 97) + 0.5f;
  cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
  printf("warp-instructions per clock per SM sub-partition (16 independent chains per thread)\n");
  run<0>("FFMA (3 regs)", out, in, cyc);
  run<6>("FFMA (imm)", out, in, cyc);
  run<1>("FFMA2", out, in, cyc);
  run<7>("FFMA2 (rotating operands)", out, in, cyc);
  run<2>("FMUL", out, in, cyc);
  run<3>("MUFU.TANH", out, in, cyc);
  run<4>("LDS.32 + FADD", out, in, cyc);
  run<5>("SHL + LOP3 + FADD", out, in, cyc);
  run<8>("HFMA2 (f16x2)", out, in, cyc);
  run<9>("HFMA2.BF16 (bf16x2)", out, in, cyc);
  run<10>("HFMA2 (rotating operands)", out, in, cyc);
  run<11>("FFMA2 + SHL/LOP3 pair", out, in, cyc);
  run<12>("HFMA2 + SHL/LOP3 pair", out, in, cyc);
  run<16>("MUFU.EX2", out, in, cyc);
  run<17>("MUFU.RCP", out, in, cyc);
  run<13>("FFMA2 + FFMA (pairs/clk)", out, in, cyc);
  run<14>("HFMA2 + FFMA (pairs/clk)", out, in, cyc);
  run<15>("HFMA2 fed by LDS.32", out, in, cyc);
  cudaError_t e = cudaGetLastError();
  printf("status: %s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
