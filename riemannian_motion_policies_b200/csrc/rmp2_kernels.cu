// sm_100a kernels of the RMP2 control step.
//
// One control step = up to four launches on the caller's stream (intermediates stay in L2/HBM, which
// this FP32-bound path uses at a few per cent of its bandwidth).  The two table-driven kernels (frames,
// step) have their bodies in rmp2_tree_kernels.cuh: here they are instantiated generically (tables as a
// kernel parameter); rmp2_jit.cu rebuilds them per tree with NVRTC (tables as compile-time constants).
//
//   rmp2_frames_kernel<N>    thread per environment: kinematic chain -> origin position, velocity and
//                            Jdot*qd of every frame that carries a sphere-obstacle leaf  (40 B records)
//   rmp2_spheres_kernel      thread per (environment, obstacle leaf): the O(frames x spheres) pair loop,
//                            two spheres per step in packed f32x2 arithmetic (FFMA2).  The sphere rows
//                            of the E environments of a block are staged into shared memory by 1-D
//                            TMA bulk copies (cp.async.bulk, mbarrier completion) at an odd pitch: HBM
//                            is read once in whole rows and the LDS.128 reads are bank-conflict free
//                            with immediate offsets.  This kernel carries >= 50 % of the step.
//                            Output: per (env, leaf) the 3x3 metric sum S and the force sum g, written
//                            over the input record.
//   rmp2_step_kernel<N>      thread per environment: chain again (cheap), target leaves, pullback of every
//                            frame's (S, g), configuration-space leaves -> combined (M, f); small batches
//                            keep the resolve fused in this kernel.
//   rmp2_resolve_kernel<N>   thread per environment: qdd = pinv(M) f (direct QR solve where provably clear
//                            of the pinv cutoff, truncated SVD by one-sided Jacobi otherwise);
//                            optional explicit-Euler sub-steps for closed-loop rollouts.
//
//   rmp2_fk_kernel<N>        FK value / velocity / Jacobian / Jdot*qd of one frame (Python API).
//   rmp2_leaf_kernel         one leaf policy at given task-space points (Python API).
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <mutex>
#include <unordered_map>

#include "rmp2_tree_kernels.cuh"
#include "rmp2_launch.h"

// ----------------------------------------------------------------------------- TMA / mbarrier PTX
RMP2_DEV uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

RMP2_DEV void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

RMP2_DEV void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}

RMP2_DEV void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  const uint32_t addr = smem_u32(bar);
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
  } while (!done);
}

// read-only 16-byte shared-memory load the compiler may schedule freely (see rmp2_spheres_kernel)
RMP2_DEV float4 lds128(uint32_t addr) {
  float4 v;
  asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}

// ------------------------------------------------------------------------------- frames kernel
// (body: rmp2_tree_kernels.cuh)
template <int N>
__global__ void __launch_bounds__(RMP2_BLOCK_THREADS)
    rmp2_frames_kernel(const __grid_constant__ StepTables T, const __grid_constant__ StepArgs A) {
  frames_body<N>(T, A);
}

// ------------------------------------------------------------------------------ spheres kernel
// Thread t of a block <-> (obstacle-leaf slot t / E, local environment t % E); the block owns a tile of
// E consecutive environments, so record loads/stores are contiguous across lanes.  (The early-out variant uses the
// order (environment t / L, slot t % L) when E is not a multiple of 8: see `slot_fastest` in the kernel.)
//
// Staging (kTma): the spheres of one environment are one contiguous row of 16*O bytes in HBM; lane e < E
// copies row e of the tile with one 1-D bulk copy (cp.async.bulk -- the TMA unit, completion on an
// mbarrier by transaction bytes) into shared memory at a pitch of 16*O + 16 bytes.  The odd pitch is
// what a swizzle would buy: the LDS.128 of 8 consecutive lanes (environments) lands in 8 distinct
// 16-byte bank groups, and a thread's sphere o sits at row + 16*o -- an immediate offset, so the pair
// loop has no per-thread address arithmetic (ALU instructions cost FP32 lane time on this machine).
// One tile per block on purpose: a persistent, double-buffered variant of this kernel measured 10 %
// SLOWER (1.09 vs 0.99 ms, config 4) -- its resident blocks run in lockstep, all warps of an SM want the
// MUFU pipe, then the FMA pipe, at the same time; short blocks that start whenever another one retires
// keep the warps of an SM spread over the phases of the loop.
#ifndef RMP2_SPHERES_MIN_BLOCKS
#define RMP2_SPHERES_MIN_BLOCKS 5     // resident blocks per SM the register allocation aims at (<= 102 registers)
#endif
#ifndef RMP2_SQRT_NEWTON
// 1: refine the pair distance |r| = dc2 * rsqrt(dc2) with one Newton step (4 lane operations per pair, +8 % kernel
// time).  Measured on a B200 (profiles/r2_parity_study.json): no effect on parity -- 7 instead of 8 of 4096 config-5
// environments beyond the 1e-5 bar, all of them explained by the float32 conditioning of the step itself -- so off.
#define RMP2_SQRT_NEWTON 0
#endif
#ifndef RMP2_SPHERES_SKIP_MIN_BLOCKS
#define RMP2_SPHERES_SKIP_MIN_BLOCKS 7  // the early-out variant is latency bound (shared-memory scoreboard): more warps help it
                                        // (measured 5 / 6 / 7 / 8 blocks: 1.070 / 1.068 / 1.028 / 1.065 ms per step)
#endif
#ifndef RMP2_SKIP_MIN_SPHERES
#define RMP2_SKIP_MIN_SPHERES 32      // rows shorter than this run every pair even with RMP2_OPT_EARLY_OUT set
#endif
#ifndef RMP2_SKIP_UNROLL
#define RMP2_SKIP_UNROLL 1            // packed steps per trip of the early-out pair loop (2: measured slower, 1.236 vs 1.170 ms/step)
#endif
#ifndef RMP2_SKIP_SORT
#define RMP2_SKIP_SORT 1              // re-deal the owners of a block by work before the early-out pair loop
#endif
#ifndef RMP2_SLOT_FASTEST
#define RMP2_SLOT_FASTEST (-1)        // early-out variant, thread <-> (slot, environment) with the slot fastest: -1 as the host decides, 0 / 1 forced
#endif
#ifndef RMP2_TMA_SPREAD
#define RMP2_TMA_SPREAD 2             // issue the row copies of a tile from all warps: bit 0 all-pairs variant, bit 1 early-out variant
#endif
#ifndef RMP2_SPHERES_STEPS_PER_TRIP
#define RMP2_SPHERES_STEPS_PER_TRIP 4 // packed (two-sphere) steps per loop trip
#endif

// Shared-memory pitch of one staged sphere row.  16 O + 16: an odd number of 16-byte units spreads the LDS.128 of 8
// consecutive lanes over all bank groups, and the unit behind the row holds the far-away sphere.  The sorted early-out
// (O <= 64) uses 67 units whatever O is: its sentinel spheres sit at the fixed slots 64 and 65.
__host__ __device__ inline uint32_t rmp2_spheres_pitch(int O, bool skip_variant) {
  return (skip_variant && RMP2_SKIP_SORT && O <= 64) ? 67u * 16u : (uint32_t)O * 16u + 16u;
}

RMP2_DEV void bulk_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// Early-out variant (kSkip, library default): pairs beyond the leaf's metric radius contribute exactly zero in
// the reference as well (rmp2.py:194), so only the others go through the pair arithmetic.  Three phases:
//   1. every thread tests its owner's (environment, obstacle leaf) spheres with a packed squared-distance test
//      (conservative by 1e-5) into two bit masks, even and odd spheres;
//   2. the owners of the block are SORTED by the number of packed steps they need (counting sort through shared
//      memory) and re-dealt to the threads in that order: the lanes of a warp then carry similar amounts of work
//      and the warp's trip count is close to its lanes' mean instead of their maximum over a random sample
//      (config 4, 20 % active pairs: 9.0 instead of 12.0 packed steps per warp, 7.6 being the mean);
//   3. the full pair only for the set bits, an even with an odd sphere per packed step (the shorter list is padded
//      with a sphere 1e15 m away, which adds exactly zero).
// Lane x of every accumulator still takes the even spheres of its owner in increasing order and lane y the odd
// ones -- which thread does the work changes, the sums do not: bit-identical to the all-pairs variant
// (tests/test_gpu_step.py::test_early_out_is_exact).
struct SkipOwner {                  // what travels with an owner when it is re-dealt (phase 2)
  float p[3], v[3];                  // (9 words: an odd stride keeps owners[t] conflict free across a warp)
  uint32_t mask_even, mask_odd;
  int32_t thread;                   // the owner's home thread (-> slot, environment: decode() in the kernel)
};

template <bool kTma, bool kSkip>
__global__ void __launch_bounds__(RMP2_SPHERES_BLOCK, (kSkip ? RMP2_SPHERES_SKIP_MIN_BLOCKS : RMP2_SPHERES_MIN_BLOCKS) * 128 / RMP2_SPHERES_BLOCK)
    rmp2_spheres_kernel(const __grid_constant__ SphereTables ST, const __grid_constant__ StepArgs A) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ SkipOwner owners[kSkip ? RMP2_SPHERES_BLOCK : 1];   // re-deal of the early-out variant (phase 2)
  __shared__ int hist[kSkip ? 33 : 1];
  const int L = ST.n_slots, E = ST.envs_per_block;
  const int t = threadIdx.x;
  if (kSkip) {                                      // zeroed ahead of the barrier(s) every thread passes before phase 2
    for (int i = t; i < 33; i += blockDim.x) hist[i] = 0;
    if (!kTma) __syncthreads();
  }
  // thread <-> (slot, local environment).  Environment-fastest (t = slot * E + e): consecutive lanes are consecutive
  // environments of one slot.  Slot-fastest (t = e * L + slot; early-out variant when the host asks for it,
  // SphereTables::slot_fastest): the L threads of an environment sit in adjacent lanes and read the SAME sphere in the
  // same instruction, so a warp's LDS.128 of the reach test touches 5-6 rows instead of 32 -- that kernel runs the
  // shared-memory data pipe at 91 % of its peak (ncu), a third of it bank conflicts of the reach test when E is not a
  // multiple of 8.  Measured (pair kernel, ms): L = 6 / E = 21 (Panda, merged control points) 0.4745 -> 0.4404;
  // L = 8 / E = 16 0.6183 -> 0.6232 (stays environment-fastest); the all-pairs variant does not care (0.849 either way).
  // (Divisions as multiply-shift: t / E = (t * magic) >> 16.)
  const bool slot_fastest = kSkip && (RMP2_SLOT_FASTEST >= 0 ? RMP2_SLOT_FASTEST != 0 : ST.slot_fastest != 0);
  auto decode = [&](int thr, int& s, int& e) {
    if (slot_fastest) {
      e = (thr * ST.div_magic_slots) >> 16;
      s = thr - e * L;
    } else {
      s = (thr * ST.div_magic) >> 16;
      e = thr - s * E;
    }
  };
  int slot, e_local;
  decode(t, slot, e_local);
  const long long env0 = (long long)blockIdx.x * E;
  long long env = env0 + e_local;
  const int O = A.n_spheres;
  bool active = (slot < L) && (e_local < E) && (env < A.B);
  const bool sorted = kSkip && RMP2_SKIP_SORT && O <= 64;   // one mask word per parity: owners can be re-dealt

  uint32_t row = 0;                                 // shared-memory address of this thread's sphere row
  uint32_t tile = 0;                                // ... of the tile's first row
  uint64_t* bar = nullptr;
  // row pitch in shared memory (rmp2_spheres_pitch, shared with the host): 16 O + 16, or 67 x 16 bytes for the sorted
  // early-out, whose exhausted lists read the two sentinel slots behind sphere 63 (see masked_pairs_rev)
  const uint32_t pitch = rmp2_spheres_pitch(O, kSkip);
  const long long rows = (A.B - env0 < E) ? (A.B - env0) : E;
  if (kTma) {
    const uint32_t row_bytes = (uint32_t)O * 16u;
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
    bar = reinterpret_cast<uint64_t*>(base + (size_t)E * pitch);
    if (t == 0) {
      mbar_init(bar, 1);
      mbar_expect_tx(bar, (uint32_t)rows * row_bytes);
    }
    __syncthreads();                              // barrier initialised before anyone copies or polls
    if (kSkip ? (RMP2_TMA_SPREAD & 2) : (RMP2_TMA_SPREAD & 1)) {
      // A bulk copy is a uniform-datapath instruction, issued once per elected lane: row e goes to lane e / W of warp
      // e % W (W warps per block), so the copies of a tile leave from all warps at once instead of one after the
      // other from the lanes of warp 0.  Measured: early-out variant 0.5028 -> 0.4962 ms (kept), all-pairs variant
      // 0.8465 -> 0.8635 ms (not used there).
      const int W = blockDim.x >> 5, e = (t & 31) * W + (t >> 5);
      if (e < rows)
        bulk_load_1d(smem_u32(base) + (uint32_t)e * pitch,
                     reinterpret_cast<const unsigned char*>(A.spheres) + (size_t)(env0 + e) * row_bytes, row_bytes, bar);
    } else {
      for (int e = t; e < rows; e += blockDim.x)    // lane e copies row e (one warp's worth for E <= 32)
        bulk_load_1d(smem_u32(base) + (uint32_t)e * pitch,
                     reinterpret_cast<const unsigned char*>(A.spheres) + (size_t)(env0 + e) * row_bytes, row_bytes, bar);
    }
    tile = smem_u32(base);
    row = tile + (uint32_t)e_local * pitch;
    // the 16 pad bytes behind every row hold a sphere that contributes exactly zero (beyond every metric radius;
    // 1e15 m away keeps all terms finite): index O of a row, what an exhausted list of the early-out loop yields
    if (kSkip && slot == 0 && e_local < E) {
      if (sorted) {     // slots 64 and 65: where index 62 + parity - 2 pos lands for pos = -1 (masked_pairs_rev)
        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %2, %2};" ::"r"(row + 64u * 16u), "f"(1e15f), "f"(0.f) : "memory");
        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %2, %2};" ::"r"(row + 65u * 16u), "f"(1e15f), "f"(0.f) : "memory");
      } else {
        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %2, %2};" ::"r"(row + (uint32_t)O * 16u), "f"(1e15f), "f"(0.f) : "memory");
      }
    }
  }
  if (!sorted && !active) return;                   // (the sorted variant keeps every thread for its barriers)

  // this thread's frame record and leaf parameters (loads in flight while the rows land)
  float* rec = A.rec + rmp2_rec_base(env, L) + slot * RMP2_REC_TILE;      // tiled record scratch, see rmp2_tables.h
  const int fstride = L * RMP2_REC_TILE;
  float px = 0.f, py = 0.f, pz = 0.f;
  float v[3] = {0.f, 0.f, 0.f};
  if (active) {
    px = rec[0 * fstride], py = rec[1 * fstride], pz = rec[2 * fstride];
    v[0] = rec[3 * fstride], v[1] = rec[4 * fstride], v[2] = rec[5 * fstride];
  }
  float p[SP_COUNT];
#pragma unroll
  for (int i = 0; i < SP_COUNT; ++i) p[i] = ST.p[active ? slot : 0][i];
  if (kTma) {
    mbar_wait(bar, 0);
    // the sphere loads below are plain (non-volatile) asm so that the scheduler may hoist them over
    // arithmetic; making their address depend on this statement keeps them after the wait
    asm volatile("" : "+r"(row) : : "memory");
  }
  // Two spheres per step in packed f32x2 arithmetic: lane x of every accumulator takes the even
  // spheres of the environment, lane y the odd ones (fixed assignment -> the early-out variant adds
  // the same terms to the same accumulators in the same order and stays bit-identical).
  float2 S[6], g[3];
  float vk[3] = {0.f, 0.f, 0.f}, vvk = 0.f;          // k v and |k v|^2 of this thread's owner (obstacle_pair2)
  auto scale_velocity = [&]() {
    const float k = p[SP_K_VEL];
#pragma unroll
    for (int i = 0; i < 3; ++i) vk[i] = k * v[i];
    vvk = fmaf(vk[0], vk[0], fmaf(vk[1], vk[1], vk[2] * vk[2]));
  };
#pragma unroll
  for (int i = 0; i < 6; ++i) S[i] = make_float2(0.f, 0.f);
#pragma unroll
  for (int i = 0; i < 3; ++i) g[i] = make_float2(0.f, 0.f);
  auto two_spheres = [&](const float4 s0, const float4 s1) {
    // pos_on_link = frame origin; pos_on_obstacle = closest surface point of the sphere
    const float2 rx = make_float2(px - s0.x, px - s1.x);
    const float2 ry = make_float2(py - s0.y, py - s1.y);
    const float2 rz = make_float2(pz - s0.z, pz - s1.z);
    // |r|^2 + 1e-24: a centre on the frame origin stays finite without a separate clamp
    const float2 dc2 = __ffma2_rn(rx, rx, __ffma2_rn(ry, ry, __ffma2_rn(rz, rz, bc2(1e-24f))));
    const float2 inv_dc = make_float2(fast_rsqrt(dc2.x), fast_rsqrt(dc2.y));
#if RMP2_SQRT_NEWTON
    // |r| to ~0.5 ulp instead of MUFU.RSQ's 2 ulp: one Newton step on dc = dc2 * rsqrt(dc2) (the residual
    // dc2 - dc^2 is exact in an FMA).  Experiment knob, see RMP2_SQRT_NEWTON above.
    const float2 dc0 = __fmul2_rn(dc2, inv_dc);
    const float2 res = __ffma2_rn(neg2(dc0), dc0, dc2);
    const float2 dc = __ffma2_rn(res, __fmul2_rn(inv_dc, bc2(0.5f)), dc0);
    const float2 sd = make_float2(dc.x - s0.w, dc.y - s1.w);                                   // signed surface distance
#else
    const float2 sd = make_float2(fmaf(dc2.x, inv_dc.x, -s0.w), fmaf(dc2.y, inv_dc.y, -s1.w));   // signed surface distance
#endif
    const float2 sgn = make_float2(copysignf(inv_dc.x, sd.x), copysignf(inv_dc.y, sd.y));       // inside: normal flips
    const float2 d = make_float2(fabsf(sd.x) + 1e-12f, fabsf(sd.y) + 1e-12f);                   // > 0 (FADD, not FMNMX)
    const float2 inv_d = make_float2(fast_rcp(d.x), fast_rcp(d.y));
    obstacle_pair2(p, __fmul2_rn(rx, sgn), __fmul2_rn(ry, sgn), __fmul2_rn(rz, sgn), d, inv_d, vk, vvk, S, g);
  };
  const float4* gs = reinterpret_cast<const float4*>(A.spheres) + (size_t)env * O;
  auto load_sphere = [&](int o) -> float4 { return kTma ? lds128(row + (uint32_t)o * 16u) : __ldg(gs + o); };
  if (!kSkip) {
    scale_velocity();
    // a sphere that contributes exactly zero (beyond every metric radius; d ~ 1e15 keeps all terms finite)
    const float4 far_away = make_float4(px + 1e15f, py, pz, 0.f);
    constexpr int kStep = 2 * RMP2_SPHERES_STEPS_PER_TRIP;      // spheres per loop trip
    int o = 0;
    for (; o + kStep <= O; o += kStep) {
#pragma unroll
      for (int k = 0; k < kStep; k += 2) two_spheres(load_sphere(o + k), load_sphere(o + k + 1));
    }
    for (; o + 1 < O; o += 2) two_spheres(load_sphere(o), load_sphere(o + 1));
    if (o < O) two_spheres(load_sphere(o), far_away);
  } else {
    // phase 1: which spheres are within reach (conservative), 64 at a time
    auto reach_masks = [&](int o0, uint32_t& mask_even, uint32_t& mask_odd) {
      const float reach = p[SP_REACH];
      const float4 nowhere = make_float4(px + 1e15f, py, pz, 0.f);
      mask_even = 0u, mask_odd = 0u;
      const int cnt = min(64, O - o0);
      for (int o = 0; o < cnt; o += 2) {
        const float4 s0 = load_sphere(o0 + o);
        const float4 s1 = (o + 1 < cnt) ? load_sphere(o0 + o + 1) : nowhere;
        const float2 rx = make_float2(px - s0.x, px - s1.x);
        const float2 ry = make_float2(py - s0.y, py - s1.y);
        const float2 rz = make_float2(pz - s0.z, pz - s1.z);
        const float2 dc2 = __ffma2_rn(rx, rx, __ffma2_rn(ry, ry, __fmul2_rn(rz, rz)));
        const float2 lim = make_float2(s0.w + reach, s1.w + reach);
        const float2 lim2 = __fmul2_rn(lim, lim);
        mask_even |= (uint32_t)(dc2.x <= lim2.x) << (o >> 1);
        mask_odd |= (uint32_t)(dc2.y <= lim2.y) << (o >> 1);
      }
    };
    // phase 1 for O <= 64, fully unrolled: per sphere one LDS.128, two FFMA2 on the register pairs the load delivers
    // -- (x, y) -> (px - x, py - y) and (z, r) -> (pz - z, reach + r) --, two FMUL2 for the four squares, three FADD for
    // w = |r|^2 - (reach + radius)^2, and one funnel shift that pushes the SIGN of w into the mask: 9 instructions
    // where the generic loop needs 15.  (w = +0 exactly on the boundary counts as outside: the reach carries a
    // 1e-5 relative margin, and a pair wrongly kept would contribute exactly zero anyway.)
    auto reach_masks_64 = [&](uint32_t& mask_even, uint32_t& mask_odd, bool reversed) {
      const float2 pxy = make_float2(px, py), pzr = make_float2(pz, p[SP_REACH]);
      const float2 mm = make_float2(-1.f, -1.f), mp = make_float2(-1.f, 1.f);
      uint32_t acc_even = 0u, acc_odd = 0u;                          // bit 31 - k <-> sphere pair k, reversed at the end
      auto margin = [&](const float4 s) -> uint32_t {
        const float2 a = __ffma2_rn(make_float2(s.x, s.y), mm, pxy);
        const float2 b = __ffma2_rn(make_float2(s.z, s.w), mp, pzr);
        const float2 a2 = __fmul2_rn(a, a), b2 = __fmul2_rn(b, b);
        return __float_as_uint((a2.x + a2.y) + (b2.x - b2.y));       // < 0 inside the reach
      };
#pragma unroll
      for (int c = 0; c < 8; ++c) {                                  // 8 spheres per chunk, one bounds check per chunk
        if (8 * c + 8 <= O) {
#pragma unroll
          for (int k = 4 * c; k < 4 * c + 4; ++k) {
            acc_even = __funnelshift_l(margin(load_sphere(2 * k)), acc_even, 1);
            acc_odd = __funnelshift_l(margin(load_sphere(2 * k + 1)), acc_odd, 1);
          }
        } else if (8 * c < O) {
#pragma unroll
          for (int k = 4 * c; k < 4 * c + 4; ++k) {
            acc_even = __funnelshift_l((2 * k < O) ? margin(load_sphere(2 * k)) : 0u, acc_even, 1);
            acc_odd = __funnelshift_l((2 * k + 1 < O) ? margin(load_sphere(2 * k + 1)) : 0u, acc_odd, 1);
          }
        } else {
          acc_even <<= 4;
          acc_odd <<= 4;
        }
      }
      mask_even = reversed ? acc_even : __brev(acc_even);
      mask_odd = reversed ? acc_odd : __brev(acc_odd);
    };
    // phase 3: the pairs of the set bits, an even with an odd sphere per packed step; an exhausted list yields the
    // far-away sphere (staged rows: the pad slot, index O -- one select on the index instead of four on the data)
    auto masked_pairs = [&](int o0, uint32_t mask_even, uint32_t mask_odd) {
      const float4 far_away = make_float4(px + 1e15f, py, pz, 0.f);
      auto next = [&](uint32_t& mask, int parity) -> float4 {
        const bool have = mask != 0u;
        const int k = __ffs((int)mask) - 1;
        mask &= mask - 1u;                                        // 0 stays 0
        if (kTma && o0 == 0) return load_sphere(have ? 2 * k + parity : O);
        const float4 s = load_sphere(o0 + (have ? 2 * k + parity : 0));
        return have ? s : far_away;
      };
      while (mask_even | mask_odd) {
        const float4 a0 = next(mask_even, 0), a1 = next(mask_odd, 1);
#if RMP2_SKIP_UNROLL >= 2
        const float4 b0 = next(mask_even, 0), b1 = next(mask_odd, 1);
        two_spheres(a0, a1);
        two_spheres(b0, b1);
#else
        two_spheres(a0, a1);
#endif
      }
    };
    // The same on staged rows with the masks in REVERSED bit order (bit 31 - k <-> sphere pair k, the order the reach
    // test's funnel shifts produce): the leading one (one FLO) is the next pair in increasing sphere order, the
    // sphere's address is base - 32 pos, and an exhausted list (pos = -1) lands on the sentinel slots 64 / 65 of the
    // row by itself -- 5 instructions per list and step (FLO, IMAD, LDS, SHF, LOP3) where the forward order needs 9
    // (BREV, FLO, ISETP, 2 k (+1), SEL, LEA, LDS, mask & (mask - 1)).
    auto masked_pairs_rev = [&](uint32_t rev_even, uint32_t rev_odd) {
      const uint32_t base_even = row + 62u * 16u, base_odd = row + 63u * 16u;
      while (rev_even | rev_odd) {
        uint32_t pe, po;                                   // bfind = FLO: position of the leading one, 0xffffffff for 0
        asm("bfind.u32 %0, %1;" : "=r"(pe) : "r"(rev_even));
        asm("bfind.u32 %0, %1;" : "=r"(po) : "r"(rev_odd));
        rev_even &= ~(1u << (pe & 31u));
        rev_odd &= ~(1u << (po & 31u));
        two_spheres(lds128(base_even - 32u * pe), lds128(base_odd - 32u * po));
      }
    };
    if (!sorted) {
      scale_velocity();
      for (int o0 = 0; o0 < O; o0 += 64) {
        uint32_t me, mo;
        reach_masks(o0, me, mo);
        masked_pairs(o0, me, mo);
      }
    } else {
      uint32_t me = 0u, mo = 0u;
      if (active) reach_masks_64(me, mo, kTma);
      const int steps = max(__popc(me), __popc(mo));            // 0 .. 32
      const int rank = atomicAdd(&hist[steps], 1);
      __syncthreads();
      // heaviest owners first: position = (owners with more steps) + rank.  Every warp forms the suffix sums of the
      // histogram on its own lanes (no third barrier, no serial loop): lane l holds sum_{s >= l, s < 32} hist[s]
      const int lane = t & 31;
      int suffix = hist[lane];
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int up = __shfl_down_sync(0xffffffffu, suffix, d);
        if (lane + d < 32) suffix += up;
      }
      const int above31 = hist[32];
      const int more = __shfl_sync(0xffffffffu, suffix, min(steps + 1, 31));   // sum_{s > steps, s < 32} (steps <= 30)
      const int before = (steps >= 32) ? 0 : above31 + ((steps >= 31) ? 0 : more);
      {
        SkipOwner& w = owners[before + rank];
        w.p[0] = px, w.p[1] = py, w.p[2] = pz;
#pragma unroll
        for (int i = 0; i < 3; ++i) w.v[i] = v[i];
        w.mask_even = me, w.mask_odd = mo;
        w.thread = active ? t : -1;
      }
      __syncthreads();
      const SkipOwner& r = owners[t];
      px = r.p[0], py = r.p[1], pz = r.p[2];
#pragma unroll
      for (int i = 0; i < 3; ++i) v[i] = r.v[i];
      me = r.mask_even, mo = r.mask_odd;
      const int home = r.thread;
      active = home >= 0;
      if (!active) return;                                        // no barrier below this line
      decode(home, slot, e_local);
      env = env0 + e_local;
      rec = A.rec + rmp2_rec_base(env, L) + slot * RMP2_REC_TILE;
      gs = reinterpret_cast<const float4*>(A.spheres) + (size_t)env * O;
      if (kTma) row = tile + (uint32_t)e_local * pitch;
#pragma unroll
      for (int i = 0; i < SP_COUNT; ++i) p[i] = ST.p[slot][i];
      scale_velocity();
      if (kTma) masked_pairs_rev(me, mo);
      else masked_pairs(0, me, mo);
    }
  }
  // The pair loop left the n.a part of the curvature term out of g and accumulated k^2 g (obstacle_pair2): the step
  // kernel finishes with g = g_sum * weight / k^2 - S a (a = Jdot qd of the frame origin is at hand there).
  // SP_WEIGHT: this leaf stands for a group of obstacle leaves that share its control point (rmp2_tree_create); the
  // group's S is the weight times its own (1 for an ordinary leaf: the product is exact)
  const float wgt = ST.p[slot][SP_WEIGHT];
#pragma unroll
  for (int i = 0; i < 6; ++i) rec[i * fstride] = (S[i].x + S[i].y) * wgt;       // fields 0..5: S, 6..8: raw g sums (in place)
#pragma unroll
  for (int i = 0; i < 3; ++i) rec[(6 + i) * fstride] = g[i].x + g[i].y;
}

// --------------------------------------------------------------------------------- step kernel
// (body: rmp2_tree_kernels.cuh)
template <int N, bool kSplit>
__global__ void __launch_bounds__(RMP2_BLOCK_THREADS, kSplit ? RMP2_SPLIT_MIN_BLOCKS(N) : RMP2_STEP_MIN_BLOCKS(N))
    rmp2_step_kernel(const __grid_constant__ StepTables T, const __grid_constant__ StepArgs A) {
  step_body<N, kSplit>(T, A);
}

// -------------------------------------------------------------------------------- resolve kernels
// Stage 1 (split mode): qdd = pinv(M) f by the direct solves of rmp2_step.cuh, environments that do not qualify
// are handed to the fallback kernel (see "resolve hand-off" in rmp2_tree_kernels.cuh).
// kQr: trees without an isotropic metric leaf, i.e. possibly rank deficient (pivoted QR + rank-revealing solve).
template <int N, bool kQr>
__global__ void __launch_bounds__(RMP2_BLOCK_THREADS, RMP2_RESOLVE_MIN_BLOCKS(N))
    rmp2_resolve_kernel(const __grid_constant__ ResolveArgs R, const __grid_constant__ StepArgs A) {
  const long long env = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = env < A.B;                 // full-warp votes inside: no early return
  const long long e = active ? env : A.B - 1;
  const int n = R.n;
  const bool rollout = A.n_sim_steps > 0;
  float M[N][N], f[N];
  const float* in = A.mf + e;
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int j = 0; j < N; ++j) M[i][j] = __ldg(in + (size_t)(i * N + j) * A.B);
#pragma unroll
  for (int i = 0; i < N; ++i) f[i] = __ldg(in + (size_t)(N * N + i) * A.B);
  float q[N], qd[N];
#pragma unroll
  for (int j = 0; j < N; ++j) {
    q[j] = (rollout && j < n) ? A.q_rw[e * n + j] : 0.f;
    qd[j] = (rollout && j < n) ? A.qd_rw[e * n + j] : 0.f;
  }
  resolve_or_defer<N, kQr, true>(A, M, f, n, R.rcond, e, active, rollout, q, qd);
}

// Stage 2: truncated SVD by one-sided Jacobi for the environments on the work list, starting from the
// factorised problem [R | Q^T f | perm] stage 1 left in the (M, f) scratch.  Launched with a fixed grid after
// every step; blocks beyond the list exit at once, and the last block to finish clears the list for the next
// step (so a captured CUDA graph replays correctly and no memset is needed).
template <int N, bool kQr>
__global__ void __launch_bounds__(RMP2_BLOCK_THREADS, RMP2_FALLBACK_MIN_BLOCKS(N))
    rmp2_resolve_fallback_kernel(const __grid_constant__ ResolveArgs R, const __grid_constant__ StepArgs A) {
  __shared__ int s_count;
  if (threadIdx.x == 0) s_count = *reinterpret_cast<volatile int*>(A.fb);
  __syncthreads();
  const int count = s_count;
  const int n = R.n;
  const bool rollout = A.n_sim_steps > 0;
  for (int base = blockIdx.x * blockDim.x; base < count; base += gridDim.x * blockDim.x) {
    const int i = base + threadIdx.x;
    const bool valid = i < count;                // lanes beyond the list idle through the sweeps
    const long long e = A.fb[2 + (valid ? i : count - 1)];
    float G[N][N], y[N], xs[N], qdd[N];
    int perm[N];
    float row[RMP2_HANDOFF_ROW(N)];                // [R upper triangle | Q^T f | perm], see defer_to_fallback
    if (!A.split) {
      const float4* in4 = reinterpret_cast<const float4*>(A.mf + (size_t)e * RMP2_HANDOFF_ROW(N));
#pragma unroll
      for (int i = 0; i < RMP2_HANDOFF_ROW(N) / 4; ++i) {
        const float4 v = in4[i];
        row[4 * i] = v.x, row[4 * i + 1] = v.y, row[4 * i + 2] = v.z, row[4 * i + 3] = v.w;
      }
    } else {
      const float* in = A.mf + e;
#pragma unroll
      for (int i = 0; i < RMP2_HANDOFF_FIELDS(N); ++i) row[i] = in[(size_t)i * A.B];
    }
    int k = 0;
#pragma unroll
    for (int r = 0; r < N; ++r)
#pragma unroll
      for (int c = 0; c < N; ++c) G[r][c] = (c >= r) ? row[k++] : 0.f;
#pragma unroll
    for (int r = 0; r < N; ++r) y[r] = row[k++];
#pragma unroll
    for (int r = 0; r < N; ++r) perm[r] = __float_as_int(row[k++]);
#pragma unroll
    for (int r = 0; r < N; ++r) xs[r] = 0.f;
    resolve_jacobi<N, kQr>(G, y, perm, !valid, xs, R.rcond, qdd);
    float q[N], qd[N];
#pragma unroll
    for (int j = 0; j < N; ++j) {
      q[j] = (rollout && j < n) ? A.q_rw[e * n + j] : 0.f;
      qd[j] = (rollout && j < n) ? A.qd_rw[e * n + j] : 0.f;
    }
    finish_step<N>(A, n, e, valid, rollout, q, qd, qdd);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const int ticket = atomicAdd(A.fb + 1, 1);
    if (ticket == (int)gridDim.x - 1) {          // every block has read the length: clear for the next step
      A.fb[0] = 0;
      A.fb[1] = 0;
      __threadfence();
    }
  }
}

// x = pinv(M) f for row-major M [B][n][n], f [B][n] -- the resolve on its own (rmp2_pinv_solve; reference:
// rmp.py:153-154).  mode 0: as in the step (direct where provable, Jacobi otherwise); mode 1: Jacobi for every
// environment (cross-check of the two solvers).
template <int N, bool kQr, bool kDirect>
__global__ void __launch_bounds__(RMP2_BLOCK_THREADS)
    rmp2_pinv_kernel(int n, float rcond, long long B, const float* __restrict__ Min, const float* __restrict__ fin,
                     float* __restrict__ x) {
  const long long env = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = env < B;
  const long long e = active ? env : B - 1;
  float M[N][N], f[N], out[N];
#pragma unroll
  for (int i = 0; i < N; ++i) {
#pragma unroll
    for (int j = 0; j < N; ++j) M[i][j] = (i < n && j < n) ? Min[(e * n + i) * n + j] : 0.f;
    f[i] = (i < n) ? fin[e * n + i] : 0.f;
  }
  resolve_pinv<N, kQr, kDirect>(M, f, n, rcond, out);
  if (active) {
#pragma unroll
    for (int j = 0; j < N; ++j)
      if (j < n) x[e * n + j] = out[j];
  }
}

// --------------------------------------------------------------------------------- feed kernel
// On-GPU stand-in for Simulation.calculate_distances (reference: simulation.py:462-484, p.getClosestPoints between
// a link's collision geometry and an obstacle) with primitive geometry on both sides: every listed frame carries a
// capsule fixed in the frame (FeedLinks; degenerate = the frame origin), obstacles are spheres (c, r) and capsules
// (a, b, r, 0).  For every (environment, listed frame, obstacle) one closest-point pair in the reference's
// distance_data layout -- pair row (pos_on_link, pos_on_obstacle, 0, 0) and aux row (distance, normal from obstacle
// to link).  Marker leaves carry the listing index of their frame in LeafTab::pair_set.
RMP2_DEV float clamp01(float x) { return fminf(fmaxf(x, 0.f), 1.f); }

// closest points of the segments P1 + s d1 and P2 + t d2, s, t in [0, 1] (either may be a point)
RMP2_DEV void closest_on_segments(const float (&P1)[3], const float (&d1)[3], const float (&P2)[3], const float (&d2)[3],
                                  float& s, float& t) {
  const float r[3] = {P1[0] - P2[0], P1[1] - P2[1], P1[2] - P2[2]};
  const float a = fmaf(d1[0], d1[0], fmaf(d1[1], d1[1], d1[2] * d1[2]));
  const float e = fmaf(d2[0], d2[0], fmaf(d2[1], d2[1], d2[2] * d2[2]));
  const float f = fmaf(d2[0], r[0], fmaf(d2[1], r[1], d2[2] * r[2]));
  const float c = fmaf(d1[0], r[0], fmaf(d1[1], r[1], d1[2] * r[2]));
  const float eps = 1e-12f;
  if (a <= eps) {                                  // first segment is a point
    s = 0.f;
    t = (e <= eps) ? 0.f : clamp01(f / e);
    return;
  }
  if (e <= eps) {                                  // second segment is a point
    t = 0.f;
    s = clamp01(-c / a);
    return;
  }
  const float b = fmaf(d1[0], d2[0], fmaf(d1[1], d2[1], d1[2] * d2[2]));
  const float denom = fmaf(a, e, -b * b);          // >= 0, 0 for parallel segments
  s = (denom > eps * a * e) ? clamp01(fmaf(b, f, -c * e) / denom) : 0.f;
  t = fmaf(b, s, f) / e;
  if (t < 0.f) {
    t = 0.f;
    s = clamp01(-c / a);
  } else if (t > 1.f) {
    t = 1.f;
    s = clamp01((b - c) / a);
  }
}

template <int N>
__global__ void __launch_bounds__(RMP2_BLOCK_THREADS)
    rmp2_feed_kernel(const __grid_constant__ StepTables T, const __grid_constant__ FeedLinks LK,
                     const __grid_constant__ FeedArgs A) {
  extern __shared__ float slots[];
  const long long env = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (env >= A.B) return;
  const int n = T.n;
  float q[N], qd[N];
#pragma unroll
  for (int j = 0; j < N; ++j) {
    q[j] = (j < n) ? A.q[env * n + j] : 0.f;
    qd[j] = 0.f;
  }
  Chain ch;
  chain_reset(ch);
  const int K = A.n_spheres + A.n_capsules;
  for (int fi = 0; fi < T.n_frames; ++fi) {
    visit_frame<N, false>(T, fi, q, qd, ch, nullptr, slots);
    const FrameTab& F = T.frames[fi];
    for (int li = F.leaf_begin; li < F.leaf_end; ++li) {
      const int listing = T.leaves[li].pair_set;
      float* prow = A.pairs + ((size_t)env * A.n_listed * K + (size_t)listing * K) * RMP2_PAIR_FLOATS;
      float* arow = A.aux ? A.aux + ((size_t)env * A.n_listed * K + (size_t)listing * K) * 4 : nullptr;
      // the frame's control capsule in world coordinates
      const float* lc = LK.c[listing];
      float la[3], lb[3], P1[3], d1[3];
      matvec3(ch.R, lc, la);
      matvec3(ch.R, lc + 3, lb);
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        P1[i] = ch.p[i] + la[i];
        d1[i] = lb[i] - la[i];
      }
      const float link_r = lc[6];
      for (int o = 0; o < K; ++o) {
        float P2[3], d2[3] = {0.f, 0.f, 0.f}, rad;
        if (o < A.n_spheres) {
          const float4 sp = __ldg(reinterpret_cast<const float4*>(A.spheres) + (size_t)env * A.n_spheres + o);
          P2[0] = sp.x, P2[1] = sp.y, P2[2] = sp.z, rad = sp.w;
        } else {
          const float4* cp = reinterpret_cast<const float4*>(A.capsules) + ((size_t)env * A.n_capsules + (o - A.n_spheres)) * 2;
          const float4 c0 = __ldg(cp), c1 = __ldg(cp + 1);          // (ax ay az bx) (by bz r 0)
          P2[0] = c0.x, P2[1] = c0.y, P2[2] = c0.z;
          d2[0] = c0.w - c0.x, d2[1] = c1.x - c0.y, d2[2] = c1.y - c0.z;
          rad = c1.z;
        }
        float s, t;
        closest_on_segments(P1, d1, P2, d2, s, t);
        const float c1x = fmaf(s, d1[0], P1[0]), c1y = fmaf(s, d1[1], P1[1]), c1z = fmaf(s, d1[2], P1[2]);
        const float c2x = fmaf(t, d2[0], P2[0]), c2y = fmaf(t, d2[1], P2[1]), c2z = fmaf(t, d2[2], P2[2]);
        const float rx = c1x - c2x, ry = c1y - c2y, rz = c1z - c2z;
        const float dc = sqrtf(fmaxf(fmaf(rx, rx, fmaf(ry, ry, rz * rz)), 1e-24f));
        const float inv = 1.f / dc;
        const float nx = rx * inv, ny = ry * inv, nz = rz * inv;
        float* pr = prow + (size_t)o * RMP2_PAIR_FLOATS;
        pr[0] = fmaf(-link_r, nx, c1x), pr[1] = fmaf(-link_r, ny, c1y), pr[2] = fmaf(-link_r, nz, c1z);
        pr[3] = fmaf(rad, nx, c2x), pr[4] = fmaf(rad, ny, c2y), pr[5] = fmaf(rad, nz, c2z);
        pr[6] = 0.f, pr[7] = 0.f;
        if (arow) {
          float* ar = arow + (size_t)o * 4;
          ar[0] = dc - rad - link_r, ar[1] = nx, ar[2] = ny, ar[3] = nz;
        }
      }
    }
  }
}

// --------------------------------------------------------------------------------------- FK kernel
// Frames of `T` are the path base -> requested frame (serial).  Outputs follow the reference's
// layout: x = row-major vec of the 4x4 transform (kinematics.py:262), J [16][n].
template <int N>
__global__ void __launch_bounds__(128)
    rmp2_fk_kernel(const __grid_constant__ StepTables T, long long B, const float* __restrict__ qin,
                   const float* __restrict__ qdin, float* __restrict__ x, float* __restrict__ xd,
                   float* __restrict__ J, float* __restrict__ c) {
  const long long env = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (env >= B) return;
  const int n = T.n;
  float q[N], qd[N];
#pragma unroll
  for (int j = 0; j < N; ++j) {
    q[j] = (j < n) ? qin[env * n + j] : 0.f;
    qd[j] = (j < n && qdin) ? qdin[env * n + j] : 0.f;
  }
  float zj[N][3], pj[N][3];
#pragma unroll
  for (int j = 0; j < N; ++j) zj[j][0] = zj[j][1] = zj[j][2] = pj[j][0] = pj[j][1] = pj[j][2] = 0.f;
  Chain ch;
  chain_reset(ch);
  uint32_t anc = 0;
  for (int fi = 0; fi < T.n_frames; ++fi) {
    const FrameTab& F = T.frames[fi];
    float qi = 0.f, qdi = 0.f;
#pragma unroll
    for (int j = 0; j < N; ++j)
      if (j == F.qidx) {
        qi = q[j];
        qdi = qd[j];
      }
    float z[3];
    chain_advance(ch, F, qi, qdi, z);
#pragma unroll
    for (int j = 0; j < N; ++j)
      if (j == F.qidx) {
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          zj[j][i] = z[i];
          pj[j][i] = ch.p[i];
        }
      }
    anc = F.anc_mask;
  }
  float* xo = x + env * 16;
#pragma unroll
  for (int r = 0; r < 3; ++r) {
#pragma unroll
    for (int k = 0; k < 3; ++k) xo[4 * r + k] = ch.R[3 * r + k];
    xo[4 * r + 3] = ch.p[r];
  }
  xo[12] = xo[13] = xo[14] = 0.f;
  xo[15] = 1.f;
  if (xd) {                                     // Rdot = [w]x R ; pdot = v
    float* o = xd + env * 16;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const float col[3] = {ch.R[k], ch.R[3 + k], ch.R[6 + k]};
      float d[3];
      cross3(ch.w, col, d);
#pragma unroll
      for (int r = 0; r < 3; ++r) o[4 * r + k] = d[r];
    }
#pragma unroll
    for (int r = 0; r < 3; ++r) o[4 * r + 3] = ch.v[r];
    o[12] = o[13] = o[14] = o[15] = 0.f;
  }
  if (c) {                                      // Rddot = [al]x R + [w]x [w]x R ; pddot = a
    float* o = c + env * 16;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const float col[3] = {ch.R[k], ch.R[3 + k], ch.R[6 + k]};
      float d1[3], d2[3], d3[3];
      cross3(ch.al, col, d1);
      cross3(ch.w, col, d2);
      cross3(ch.w, d2, d3);
#pragma unroll
      for (int r = 0; r < 3; ++r) o[4 * r + k] = d1[r] + d3[r];
    }
#pragma unroll
    for (int r = 0; r < 3; ++r) o[4 * r + 3] = ch.a[r];
    o[12] = o[13] = o[14] = o[15] = 0.f;
  }
  if (J) {
    float* o = J + env * 16 * n;
    for (int i = 0; i < 16 * n; ++i) o[i] = 0.f;
#pragma unroll
    for (int j = 0; j < N; ++j) {
      if (j < n && (anc & (1u << j))) {
        if (T.prismatic_mask & (1u << j)) {
#pragma unroll
          for (int r = 0; r < 3; ++r) o[(4 * r + 3) * n + j] = zj[j][r];
        } else {
          const float rr[3] = {ch.p[0] - pj[j][0], ch.p[1] - pj[j][1], ch.p[2] - pj[j][2]};
          float d[3];
          cross3(zj[j], rr, d);
#pragma unroll
          for (int r = 0; r < 3; ++r) o[(4 * r + 3) * n + j] = d[r];
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            const float col[3] = {ch.R[k], ch.R[3 + k], ch.R[6 + k]};
            float dr[3];
            cross3(zj[j], col, dr);
#pragma unroll
            for (int r = 0; r < 3; ++r) o[(4 * r + k) * n + j] = dr[r];
          }
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------- leaf kernel
// One thread per task-space point.  D = RMP2_MAX_JOINTS covers every task dimension in use.
__global__ void __launch_bounds__(128)
    rmp2_leaf_kernel(const __grid_constant__ LeafTab L, const __grid_constant__ LeafVec V, int m, long long K,
                     const float* __restrict__ xin, const float* __restrict__ xdin, const float* __restrict__ aux,
                     float* __restrict__ xdd_out, float* __restrict__ M_out) {
  constexpr int D = RMP2_MAX_JOINTS;
  const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  float x[D], xd[D], xdd[D];
#pragma unroll
  for (int i = 0; i < D; ++i) {
    x[i] = (i < m) ? xin[k * m + i] : 0.f;
    xd[i] = (i < m) ? xdin[k * m + i] : 0.f;
    xdd[i] = 0.f;
  }
  float* Mo = M_out + k * m * m;
  for (int i = 0; i < m * m; ++i) Mo[i] = 0.f;
  if (L.type == RMP2_LEAF_OBSTACLE_AVOIDANCE) {
    float a, mm;
    obstacle_scalar(L.p, x[0], xd[0], a, mm);
    xdd[0] = a;
    Mo[0] = mm;
  } else if (L.type == RMP2_LEAF_COLLISION_AVOIDANCE) {
    const float vec[3] = {aux[4 * k + 1], aux[4 * k + 2], aux[4 * k + 3]};
    const float xd3[3] = {xd[0], xd[1], xd[2]};
    float f3[3], w;
    collision_avoidance_v1(L.p, aux[4 * k], vec, xd3, f3, w);
    for (int i = 0; i < 3; ++i) {
      xdd[i] = f3[i];
      Mo[4 * i] = w;
    }
  } else if (L.type == RMP2_LEAF_TARGET_POLICY || L.type == RMP2_LEAF_TARGET_ATTRACTOR) {
    float zeta[D], iso, dir;
#pragma unroll
    for (int i = 0; i < D; ++i) zeta[i] = 0.f;
    if (L.type == RMP2_LEAF_TARGET_POLICY) {
      float goal[D];
#pragma unroll
      for (int i = 0; i < D; ++i) goal[i] = (i < m) ? V.v[i] : 0.f;
      target_policy<D>(L.p, x, xd, goal, m, xdd, zeta, iso, dir);
    } else {
      const float x3[3] = {x[0], x[1], x[2]}, xd3[3] = {xd[0], xd[1], xd[2]}, g3[3] = {V.v[0], V.v[1], V.v[2]};
      float a3[3], z3[3];
      target_attractor(L.p, x3, xd3, g3, a3, z3, iso, dir);
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        xdd[i] = a3[i];
        zeta[i] = z3[i];
      }
    }
    for (int i = 0; i < m; ++i)
      for (int j = 0; j < m; ++j) Mo[i * m + j] = fmaf(dir * zeta[i], zeta[j], (i == j) ? iso : 0.f);
  } else if (L.type == RMP2_LEAF_CONFIG_BIASING || L.type == RMP2_LEAF_JOINT_DAMPING ||
             L.type == RMP2_LEAF_CSPACE_BIASING) {
    float mm;
    if (L.type == RMP2_LEAF_CONFIG_BIASING)
      leaf_config_biasing<D>(L.p, V.v, m, x, xd, xdd, mm);
    else if (L.type == RMP2_LEAF_JOINT_DAMPING)
      leaf_joint_damping<D>(L.p, m, xd, xdd, mm);
    else
      leaf_cspace_biasing<D>(L.p, V.v, m, x, xd, xdd, mm);
    for (int i = 0; i < m; ++i) Mo[i * m + i] = mm;
  } else if (L.type == RMP2_LEAF_VELOCITY_CAP) {
    float diag[D], w;
    leaf_velocity_cap<D>(L.p, m, xd, xdd, diag, w);
    for (int i = 0; i < m; ++i)
      for (int j = 0; j < m; ++j) Mo[i * m + j] = (i == j) ? diag[i] : w;
  } else if (L.type == RMP2_LEAF_JOINT_LIMIT) {
    float zeta[D], w[D];
    leaf_joint_limit<D>(L.p, V.v, m, x, xd, xdd, zeta, w);
    const float beta = L.p[JL_BETA];
    for (int i = 0; i < m; ++i)
      for (int j = 0; j < m; ++j) Mo[i * m + j] = fmaf(beta * zeta[i], zeta[j], (i == j) ? (1.f - beta) : 0.f) * w[j];
  }
  for (int i = 0; i < m; ++i) xdd_out[k * m + i] = xdd[i];
}

// -------------------------------------------------------------------------------- host launchers
// Opt a kernel in to `smem` bytes of dynamic shared memory (needed above 48 KB).  The largest size granted per
// kernel is remembered, so the steady state makes no runtime call (the attribute only ever grows).
static cudaError_t allow_dynamic_smem(const void* fn, size_t smem) {
  static std::mutex mu;
  static std::unordered_map<const void*, size_t> granted;
  std::lock_guard<std::mutex> lock(mu);
  size_t& have = granted[fn];
  if (smem <= have) return cudaSuccess;
  if (smem > 227 * 1024) return cudaErrorInvalidConfiguration;
  cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e == cudaSuccess) have = smem;
  return e;
}

size_t rmp2_step_smem(const StepTables& T, int block) {
  return ((size_t)T.n_slots * RMP2_CHAIN_FLOATS + 6 * (size_t)rmp2_pick_width(T.n)) * block * sizeof(float);
}

int rmp2_pick_width(int n) {
  if (n <= 2) return 2;
  if (n <= 7) return 7;
  if (n <= 9) return 9;
  return 12;
}

#define RMP2_DISPATCH_N(n, CALL)            \
  switch (rmp2_pick_width(n)) {             \
    case 2: { constexpr int NN = 2; CALL; } break;   \
    case 7: { constexpr int NN = 7; CALL; } break;   \
    case 9: { constexpr int NN = 9; CALL; } break;   \
    default: { constexpr int NN = 12; CALL; } break; \
  }

cudaError_t rmp2_launch_frames(const StepTables& T, const StepArgs& A, int block, cudaStream_t stream) {
  const long long blocks = (A.B + block - 1) / block;
  if (blocks <= 0) return cudaSuccess;
  const size_t smem = (size_t)T.n_slots * RMP2_CHAIN_FLOATS * block * sizeof(float);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaSuccess;
    RMP2_DISPATCH_N(T.n, (e = allow_dynamic_smem((const void*)rmp2_frames_kernel<NN>, smem)));
    if (e != cudaSuccess) return e;
  }
  RMP2_DISPATCH_N(T.n, (rmp2_frames_kernel<NN><<<(unsigned)blocks, block, smem, stream>>>(T, A)));
  return cudaGetLastError();
}

size_t rmp2_spheres_smem(const SphereTables& ST, int n_spheres, bool use_tma, bool early_out) {
  if (!use_tma) return 0;
  const size_t pitch = rmp2_spheres_pitch(n_spheres, early_out && n_spheres >= RMP2_SKIP_MIN_SPHERES);
  return 128 + (size_t)ST.envs_per_block * pitch + sizeof(uint64_t);
}

cudaError_t rmp2_launch_spheres(const SphereTables& ST, const StepArgs& A, bool use_tma, cudaStream_t stream) {
  const long long blocks = (A.B + ST.envs_per_block - 1) / ST.envs_per_block;
  if (blocks <= 0) return cudaSuccess;
  if (blocks > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
  const int threads = ((ST.envs_per_block * ST.n_slots + 31) / 32) * 32;
  const size_t smem = rmp2_spheres_smem(ST, A.n_spheres, use_tma, A.early_out != 0);
  const unsigned nb = (unsigned)blocks;
  // the early-out variant pays a reach test per pair and a re-deal of the block's work: below ~32 spheres per
  // environment that costs more than the skipped pairs save (measured, config 3 with 16 spheres: 0.8x), so short
  // rows run every pair -- the results are identical either way
  const bool skip = A.early_out && A.n_spheres >= RMP2_SKIP_MIN_SPHERES;
  if (use_tma) {
    const void* fn = skip ? (const void*)rmp2_spheres_kernel<true, true> : (const void*)rmp2_spheres_kernel<true, false>;
    cudaError_t e = allow_dynamic_smem(fn, smem);
    if (e != cudaSuccess) return e;
    if (skip) rmp2_spheres_kernel<true, true><<<nb, threads, smem, stream>>>(ST, A);
    else rmp2_spheres_kernel<true, false><<<nb, threads, smem, stream>>>(ST, A);
  } else {
    if (skip) rmp2_spheres_kernel<false, true><<<nb, threads, 0, stream>>>(ST, A);
    else rmp2_spheres_kernel<false, false><<<nb, threads, 0, stream>>>(ST, A);
  }
  return cudaGetLastError();
}

cudaError_t rmp2_launch_step(const StepTables& T, const StepArgs& A, int block, cudaStream_t stream) {
  const long long blocks = (A.B + block - 1) / block;
  if (blocks <= 0) return cudaSuccess;
  const size_t smem = rmp2_step_smem(T, block);
  cudaError_t e = cudaSuccess;
  if (smem > 48 * 1024) {
    if (A.split) { RMP2_DISPATCH_N(T.n, (e = allow_dynamic_smem((const void*)rmp2_step_kernel<NN, true>, smem))); }
    else { RMP2_DISPATCH_N(T.n, (e = allow_dynamic_smem((const void*)rmp2_step_kernel<NN, false>, smem))); }
    if (e != cudaSuccess) return e;
  }
  if (A.split) {
    RMP2_DISPATCH_N(T.n, (rmp2_step_kernel<NN, true><<<(unsigned)blocks, block, smem, stream>>>(T, A)));
  } else {
    RMP2_DISPATCH_N(T.n, (rmp2_step_kernel<NN, false><<<(unsigned)blocks, block, smem, stream>>>(T, A)));
  }
  return cudaGetLastError();
}

cudaError_t rmp2_launch_resolve(const StepTables& T, const StepArgs& A, int block, cudaStream_t stream) {
  const long long blocks = (A.B + block - 1) / block;
  if (blocks <= 0) return cudaSuccess;
  ResolveArgs R;
  R.n = T.n;
  R.rcond = T.rcond;
  if (T.precondition) {
    RMP2_DISPATCH_N(T.n, (rmp2_resolve_kernel<NN, true><<<(unsigned)blocks, block, 0, stream>>>(R, A)));
  } else {
    RMP2_DISPATCH_N(T.n, (rmp2_resolve_kernel<NN, false><<<(unsigned)blocks, block, 0, stream>>>(R, A)));
  }
  return cudaGetLastError();
}

cudaError_t rmp2_launch_fallback(const StepTables& T, const StepArgs& A, int max_blocks, cudaStream_t stream) {
  const int block = RMP2_BLOCK_THREADS;
  long long blocks = (A.B + block - 1) / block;
  if (blocks <= 0) return cudaSuccess;
  if (blocks > max_blocks) blocks = max_blocks;
  ResolveArgs R;
  R.n = T.n;
  R.rcond = T.rcond;
  if (T.precondition) {
    RMP2_DISPATCH_N(T.n, (rmp2_resolve_fallback_kernel<NN, true><<<(unsigned)blocks, block, 0, stream>>>(R, A)));
  } else {
    RMP2_DISPATCH_N(T.n, (rmp2_resolve_fallback_kernel<NN, false><<<(unsigned)blocks, block, 0, stream>>>(R, A)));
  }
  return cudaGetLastError();
}

cudaError_t rmp2_launch_pinv(int n, float rcond, bool pivot, int mode, long long B, const float* M, const float* f,
                             float* x, cudaStream_t stream) {
  const long long blocks = (B + RMP2_BLOCK_THREADS - 1) / RMP2_BLOCK_THREADS;
  if (blocks <= 0) return cudaSuccess;
  const unsigned nb = (unsigned)blocks;
  if (mode == 1) {
    if (pivot) { RMP2_DISPATCH_N(n, (rmp2_pinv_kernel<NN, true, false><<<nb, RMP2_BLOCK_THREADS, 0, stream>>>(n, rcond, B, M, f, x))); }
    else { RMP2_DISPATCH_N(n, (rmp2_pinv_kernel<NN, false, false><<<nb, RMP2_BLOCK_THREADS, 0, stream>>>(n, rcond, B, M, f, x))); }
  } else {
    if (pivot) { RMP2_DISPATCH_N(n, (rmp2_pinv_kernel<NN, true, true><<<nb, RMP2_BLOCK_THREADS, 0, stream>>>(n, rcond, B, M, f, x))); }
    else { RMP2_DISPATCH_N(n, (rmp2_pinv_kernel<NN, false, true><<<nb, RMP2_BLOCK_THREADS, 0, stream>>>(n, rcond, B, M, f, x))); }
  }
  return cudaGetLastError();
}

cudaError_t rmp2_kernel_attributes(int n, int which, bool use_tma, int block, size_t smem, int* regs,
                                   int* blocks_per_sm) {
  cudaFuncAttributes attr;
  cudaError_t e = cudaSuccess;
  if (which == 1) {
    const void* fn = use_tma ? (const void*)rmp2_spheres_kernel<true, false> : (const void*)rmp2_spheres_kernel<false, false>;
    e = cudaFuncGetAttributes(&attr, fn);
    if (e != cudaSuccess) return e;
    if (smem > 0) {
      e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return e;
    }
    e = use_tma ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, rmp2_spheres_kernel<true, false>, block, smem)
                : cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, rmp2_spheres_kernel<false, false>, block, smem);
  } else if (which == 0) {
    RMP2_DISPATCH_N(n, (e = cudaFuncGetAttributes(&attr, rmp2_frames_kernel<NN>),
                        e = (e == cudaSuccess) ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(
                                                     blocks_per_sm, rmp2_frames_kernel<NN>, block, smem) : e));
  } else if (which == 2) {
    RMP2_DISPATCH_N(n, (e = cudaFuncGetAttributes(&attr, rmp2_step_kernel<NN, false>),
                        e = (e == cudaSuccess) ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(
                                                     blocks_per_sm, rmp2_step_kernel<NN, false>, block, smem) : e));
  } else if (which == 3) {
    RMP2_DISPATCH_N(n, (e = cudaFuncGetAttributes(&attr, rmp2_step_kernel<NN, true>),
                        e = (e == cudaSuccess) ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(
                                                     blocks_per_sm, rmp2_step_kernel<NN, true>, block, smem) : e));
  } else if (which == 4) {
    RMP2_DISPATCH_N(n, (e = cudaFuncGetAttributes(&attr, rmp2_resolve_kernel<NN, true>),
                        e = (e == cudaSuccess) ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(
                                                     blocks_per_sm, rmp2_resolve_kernel<NN, true>, block, 0) : e));
  } else {
    RMP2_DISPATCH_N(n, (e = cudaFuncGetAttributes(&attr, rmp2_resolve_fallback_kernel<NN, true>),
                        e = (e == cudaSuccess) ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(
                                                     blocks_per_sm, rmp2_resolve_fallback_kernel<NN, true>, block, 0) : e));
  }
  if (e != cudaSuccess) return e;
  *regs = attr.numRegs;
  return cudaSuccess;
}

cudaError_t rmp2_launch_feed(const StepTables& T, const FeedLinks& LK, const FeedArgs& A, cudaStream_t stream) {
  const int block = RMP2_BLOCK_THREADS;
  const long long blocks = (A.B + block - 1) / block;
  if (blocks <= 0) return cudaSuccess;
  const size_t smem = (size_t)T.n_slots * RMP2_CHAIN_FLOATS * block * sizeof(float);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaSuccess;
    RMP2_DISPATCH_N(T.n, (e = allow_dynamic_smem((const void*)rmp2_feed_kernel<NN>, smem)));
    if (e != cudaSuccess) return e;
  }
  RMP2_DISPATCH_N(T.n, (rmp2_feed_kernel<NN><<<(unsigned)blocks, block, smem, stream>>>(T, LK, A)));
  return cudaGetLastError();
}

cudaError_t rmp2_launch_fk(const StepTables& T, long long B, const float* q, const float* qd, float* x, float* xd,
                           float* J, float* c, cudaStream_t stream) {
  const long long blocks = (B + 127) / 128;
  if (blocks <= 0) return cudaSuccess;
  RMP2_DISPATCH_N(T.n, (rmp2_fk_kernel<NN><<<(unsigned)blocks, 128, 0, stream>>>(T, B, q, qd, x, xd, J, c)));
  return cudaGetLastError();
}

cudaError_t rmp2_launch_leaf(const LeafTab& L, const LeafVec& V, int m, long long K, const float* x, const float* xd,
                             const float* aux, float* xdd, float* M, cudaStream_t stream) {
  const long long blocks = (K + 127) / 128;
  if (blocks <= 0) return cudaSuccess;
  rmp2_leaf_kernel<<<(unsigned)blocks, 128, 0, stream>>>(L, V, m, K, x, xd, aux, xdd, M);
  return cudaGetLastError();
}
