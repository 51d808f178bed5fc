// Measured issue rates of the pipes the RMP2 pair loop is bound by, on the GPU it runs on:
// scalar FFMA, packed FFMA2 (fma.rn.f32x2, new on sm_100), MUFU (ex2/rcp/rsqrt), and FFMA2 with MUFU
// interleaved (do the two pipes overlap?).  Prints one JSON object; bench.py / DESIGN.md quote it as the
// measured FP32 and MUFU peaks next to the analytic 148 x 128 x 2 x f_SM figure.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/pipe_peaks tools/pipe_peaks.cu && tools/pipe_peaks
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

#define CHECK(x)                                                                      \
  do {                                                                                \
    cudaError_t e_ = (x);                                                             \
    if (e_ != cudaSuccess) {                                                          \
      fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_));      \
      exit(1);                                                                        \
    }                                                                                 \
  } while (0)

constexpr int kIters = 4096;
constexpr int kChains = 8;    // independent dependency chains per thread (latency 4 -> needs >= 4 with 1 warp)

__device__ __forceinline__ float mufu_ex2(float x) {
  float y;
  asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// mode 0: FFMA, 1: FFMA2, 2: MUFU.EX2, 3: FFMA2 + MUFU (1 MUFU per 4 FFMA2), 4: FFMA + MUFU (1 per 8)
template <int kMode>
__global__ void __launch_bounds__(256) pipe_kernel(float* out, float a, float b) {
  float2 acc[kChains];
#pragma unroll
  for (int c = 0; c < kChains; ++c) acc[c] = make_float2(threadIdx.x * 1e-3f + c, threadIdx.x * 2e-3f - c);
  const float2 a2 = make_float2(a, a), b2 = make_float2(b, b);
  float m = threadIdx.x * 1e-4f;
  float sc[2 * kChains];
  unsigned ic[2 * kChains];
  const unsigned ia = __float_as_uint(a) | 1u, ib = __float_as_uint(b) | 0xffff0000u;
#pragma unroll
  for (int c = 0; c < 2 * kChains; ++c) {
    sc[c] = threadIdx.x * 3e-3f + c;
    ic[c] = threadIdx.x * 77u + c;
  }
#pragma unroll 1
  for (int it = 0; it < kIters; ++it) {
    if (kMode == 0) {
#pragma unroll
      for (int c = 0; c < kChains; ++c) {
        acc[c].x = fmaf(acc[c].x, a, b);
        acc[c].y = fmaf(acc[c].y, a, b);
      }
    } else if (kMode == 1) {
#pragma unroll
      for (int c = 0; c < kChains; ++c) {
        acc[c] = __ffma2_rn(acc[c], a2, b2);
        acc[c] = __ffma2_rn(acc[c], a2, b2);
      }
    } else if (kMode == 2) {
#pragma unroll
      for (int c = 0; c < kChains; ++c) {
        acc[c].x = mufu_ex2(acc[c].x);
        acc[c].y = mufu_ex2(acc[c].y);
      }
    } else if (kMode == 3) {
#pragma unroll
      for (int c = 0; c < kChains; ++c) {
        acc[c] = __ffma2_rn(acc[c], a2, b2);
        acc[c] = __ffma2_rn(acc[c], a2, b2);
      }
      m = mufu_ex2(m);
      m = mufu_ex2(m);
      m = mufu_ex2(m);
      m = mufu_ex2(m);
    } else if (kMode == 4) {
#pragma unroll
      for (int c = 0; c < kChains; ++c) {
        acc[c].x = fmaf(acc[c].x, a, b);
        acc[c].y = fmaf(acc[c].y, a, b);
      }
      m = mufu_ex2(m);
      m = mufu_ex2(m);
    } else if (kMode >= 7) {
      // 16 FFMA2 + 8 (mode 7) or 16 (mode 8) FMNMX, or 16 integer LOP3 (mode 9): does ALU-pipe work
      // take FP32 lane time away from the packed instructions?
#pragma unroll
      for (int c = 0; c < kChains; ++c) {
        acc[c] = __ffma2_rn(acc[c], a2, b2);
        if (kMode == 9) {
          ic[c] = (ic[c] ^ ia) & ib;
          ic[c + kChains] = (ic[c + kChains] ^ ib) | ia;
        } else {
          sc[c] = fmaxf(sc[c], sc[c + kChains] * 0.f + a);
          if (kMode == 8) sc[c + kChains] = fminf(sc[c + kChains], b);
        }
        acc[c] = __ffma2_rn(acc[c], a2, b2);
      }
    } else {
      // 16 FFMA2 + (kMode == 5 ? 8 : 16) scalar FFMA on independent chains: does the second FP32
      // sub-pipe run scalar work under the packed instructions?
#pragma unroll
      for (int c = 0; c < kChains; ++c) {
        acc[c] = __ffma2_rn(acc[c], a2, b2);
        sc[c] = fmaf(sc[c], a, b);
        acc[c] = __ffma2_rn(acc[c], a2, b2);
        if (kMode == 6) sc[c + kChains] = fmaf(sc[c + kChains], a, b);
      }
    }
  }
  float s = m;
#pragma unroll
  for (int c = 0; c < 2 * kChains; ++c) s += sc[c] + __uint_as_float(ic[c]);
#pragma unroll
  for (int c = 0; c < kChains; ++c) s += acc[c].x + acc[c].y;
  if (s == 12345.678f) out[0] = s;    // never true: keeps the chains alive
}

// MUFU rate per operation: 0 ex2, 1 rcp, 2 rsqrt, 3 lg2, 4 sin, 5 sqrt, 6 tanh
template <int kOp>
__device__ __forceinline__ float mufu_op(float x) {
  float y;
  if (kOp == 0) asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  if (kOp == 1) asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  if (kOp == 2) asm volatile("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  if (kOp == 3) asm volatile("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  if (kOp == 4) asm volatile("sin.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  if (kOp == 5) asm volatile("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  if (kOp == 6) asm volatile("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int kOp>
__global__ void __launch_bounds__(256) mufu_kernel(float* out) {
  float acc[2 * kChains];
#pragma unroll
  for (int c = 0; c < 2 * kChains; ++c) acc[c] = 1.0f + threadIdx.x * 1e-3f + c;
#pragma unroll 1
  for (int it = 0; it < kIters; ++it) {
#pragma unroll
    for (int c = 0; c < 2 * kChains; ++c) acc[c] = mufu_op<kOp>(acc[c]);
  }
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < 2 * kChains; ++c) s += acc[c];
  if (s == 12345.678f) out[0] = s;
}

template <int kOp>
static double mufu_per_clk_per_sm(int blocks, float* out, int sms, int clock_khz) {
  cudaEvent_t e0, e1;
  CHECK(cudaEventCreate(&e0));
  CHECK(cudaEventCreate(&e1));
  mufu_kernel<kOp><<<blocks, 256>>>(out);
  CHECK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < 5; ++r) {
    CHECK(cudaEventRecord(e0));
    mufu_kernel<kOp><<<blocks, 256>>>(out);
    CHECK(cudaEventRecord(e1));
    CHECK(cudaEventSynchronize(e1));
    float ms;
    CHECK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  return (double)blocks * 256 * kIters * 2 * kChains / (best * 1e-3) / sms / (clock_khz * 1e3);
}

template <int kMode>
static double time_ms(int blocks, float* out) {
  cudaEvent_t e0, e1;
  CHECK(cudaEventCreate(&e0));
  CHECK(cudaEventCreate(&e1));
  pipe_kernel<kMode><<<blocks, 256>>>(out, 0.999f, 1e-3f);
  CHECK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < 5; ++r) {
    CHECK(cudaEventRecord(e0));
    pipe_kernel<kMode><<<blocks, 256>>>(out, 0.999f, 1e-3f);
    CHECK(cudaEventRecord(e1));
    CHECK(cudaEventSynchronize(e1));
    float ms;
    CHECK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  return best;
}

int main() {
  cudaDeviceProp prop;
  CHECK(cudaGetDeviceProperties(&prop, 0));
  int clock_khz = 0;
  CHECK(cudaDeviceGetAttribute(&clock_khz, cudaDevAttrClockRate, 0));
  float* out;
  CHECK(cudaMalloc(&out, 4));
  const int sms = prop.multiProcessorCount;
  const int blocks = sms * 8 * 4;                       // 8 resident blocks of 8 warps, 4 waves
  const double threads = (double)blocks * 256;
  const double ms_ffma = time_ms<0>(blocks, out);
  const double ms_ffma2 = time_ms<1>(blocks, out);
  const double ms_mufu = time_ms<2>(blocks, out);
  const double ms_mix2 = time_ms<3>(blocks, out);
  const double ms_mix1 = time_ms<4>(blocks, out);
  const double ms_p16s8 = time_ms<5>(blocks, out);
  const double ms_p16s16 = time_ms<6>(blocks, out);
  const double ms_a8 = time_ms<7>(blocks, out);
  const double ms_a16 = time_ms<8>(blocks, out);
  const double ms_i16 = time_ms<9>(blocks, out);
  const double n_ffma = threads * kIters * kChains * 2;           // scalar FFMA instructions (thread level)
  const double n_ffma2 = threads * kIters * kChains * 2;          // packed instructions, 2 FMAs each
  const double n_mufu = threads * kIters * kChains * 2;
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_mhz_max\": %.0f,\n", prop.name, sms, clock_khz / 1e3);
  printf(" \"ffma_tflops\": %.2f, \"ffma2_tflops\": %.2f, \"mufu_tops\": %.3f,\n",
         n_ffma * 2 / ms_ffma / 1e9, n_ffma2 * 4 / ms_ffma2 / 1e9, n_mufu / ms_mufu / 1e9);
  printf(" \"ffma_per_clk_per_sm\": %.1f, \"ffma2_fma_per_clk_per_sm\": %.1f, \"mufu_per_clk_per_sm\": %.2f,\n",
         n_ffma / (ms_ffma * 1e-3) / sms / (clock_khz * 1e3), n_ffma2 * 2 / (ms_ffma2 * 1e-3) / sms / (clock_khz * 1e3),
         n_mufu / (ms_mufu * 1e-3) / sms / (clock_khz * 1e3));
  // mixed kernels: time relative to the sum / max of the parts (1.0 = perfect overlap with the slower pipe)
  const double mufu_part2 = ms_mufu * 4.0 / (kChains * 2);
  const double mufu_part1 = ms_mufu * 2.0 / (kChains * 2);
  printf(" \"ffma2_plus_mufu_ms\": %.4f, \"ffma2_alone_ms\": %.4f, \"mufu_part_ms\": %.4f,\n", ms_mix2, ms_ffma2,
         mufu_part2);
  printf(" \"ffma_plus_mufu_ms\": %.4f, \"ffma_alone_ms\": %.4f, \"mufu_part1_ms\": %.4f,\n", ms_mix1, ms_ffma,
         mufu_part1);
  // packed + scalar mixes: FMAs per clock per SM (16 FFMA2 = 32 FMAs, plus 8 or 16 scalar per iteration)
  printf(" \"ffma2x16_plus_ffma8_ms\": %.4f, \"fma_per_clk_per_sm_p16s8\": %.1f,\n", ms_p16s8,
         threads * kIters * 40.0 / (ms_p16s8 * 1e-3) / sms / (clock_khz * 1e3));
  printf(" \"ffma2x16_plus_ffma16_ms\": %.4f, \"fma_per_clk_per_sm_p16s16\": %.1f}\n", ms_p16s16,
         threads * kIters * 48.0 / (ms_p16s16 * 1e-3) / sms / (clock_khz * 1e3));
  printf("{\"ffma2x16_ms\": %.4f, \"plus_fmnmx8_ms\": %.4f, \"plus_fmnmx16_ms\": %.4f, \"plus_lop16_ms\": %.4f}\n", ms_ffma2,
         ms_a8, ms_a16, ms_i16);
  printf("{\"mufu_per_clk_per_sm\": {\"ex2\": %.2f, \"rcp\": %.2f, \"rsqrt\": %.2f, \"lg2\": %.2f, \"sin\": %.2f, \"sqrt\": %.2f, \"tanh\": %.2f}}\n",
         mufu_per_clk_per_sm<0>(blocks, out, sms, clock_khz), mufu_per_clk_per_sm<1>(blocks, out, sms, clock_khz),
         mufu_per_clk_per_sm<2>(blocks, out, sms, clock_khz), mufu_per_clk_per_sm<3>(blocks, out, sms, clock_khz),
         mufu_per_clk_per_sm<4>(blocks, out, sms, clock_khz), mufu_per_clk_per_sm<5>(blocks, out, sms, clock_khz),
         mufu_per_clk_per_sm<6>(blocks, out, sms, clock_khz));
  return 0;
}
