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
    } else {
#pragma unroll
      for (int c = 0; c < kChains; ++c) {
        acc[c].x = fmaf(acc[c].x, a, b);
        acc[c].y = fmaf(acc[c].y, a, b);
      }
      m = mufu_ex2(m);
      m = mufu_ex2(m);
    }
  }
  float s = m;
#pragma unroll
  for (int c = 0; c < kChains; ++c) s += acc[c].x + acc[c].y;
  if (s == 12345.678f) out[0] = s;    // never true: keeps the chains alive
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
  printf(" \"ffma_plus_mufu_ms\": %.4f, \"ffma_alone_ms\": %.4f, \"mufu_part1_ms\": %.4f}\n", ms_mix1, ms_ffma,
         mufu_part1);
  return 0;
}
