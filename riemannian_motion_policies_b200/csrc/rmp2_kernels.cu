// sm_100a kernels of the RMP2 control step.
//
//   rmp2_step_kernel<N, kTma>   THE hot path: one thread per environment walks the kinematic tree
//                               once per sphere tile, evaluates every leaf, pulls back, resolves.
//                               Sphere obstacles of the 32 environments of a warp are staged into
//                               shared memory by one TMA 2-D tiled bulk copy per 8 spheres
//                               (cp.async.bulk.tensor, 128-byte swizzle, warp-private mbarrier),
//                               so HBM is read once, in full 128-byte lines, and the per-thread
//                               row reads from shared memory are bank-conflict free.
//   rmp2_fk_kernel<N>           FK value / velocity / Jacobian / Jdot*qd of one frame (Python API).
//   rmp2_leaf_kernel            one leaf policy at given task-space points (Python API).
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "rmp2_step.cuh"
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

RMP2_DEV void tma_load_2d(void* dst, const CUtensorMap* tmap, int32_t x, int32_t y, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(smem_u32(dst)), "l"(tmap), "r"(x), "r"(y), "r"(smem_u32(bar))
      : "memory");
}

// --------------------------------------------------------------------------------- per-env context
struct WarpTile {
  const char* base;   // this warp's staged spheres: [box][32 rows][128 B], 128B-swizzled
  uint32_t base32;    // the same, as a shared-window address
  uint64_t* bar;      // this warp's mbarrier
  uint32_t phase;     // parity of the next completion to wait for
  int boxes;          // boxes per tile (<= 4)
};

// One environment, one control step: q, qd -> qdd.
template <int N, bool kTma>
RMP2_DEV void evaluate_env(const StepTables& T, const StepArgs& A, const CUtensorMap* tmap, long long env,
                           long long warp_env0, const float (&q)[N], const float (&qd)[N], float* slots,
                           WarpTile& wt, float (&qdd)[N]) {
  const int n = T.n;
  const uint32_t lane = threadIdx.x & 31u;
  float Msym[N * (N + 1) / 2];
  float f[N];
#pragma unroll
  for (int i = 0; i < N * (N + 1) / 2; ++i) Msym[i] = 0.f;
#pragma unroll
  for (int i = 0; i < N; ++i) f[i] = 0.f;

  const int O = A.n_spheres;
  const int n_tiles = (T.uses_spheres && O > 0) ? (O + RMP2_TILE_SPHERES - 1) / RMP2_TILE_SPHERES : 1;

  for (int tile = 0; tile < n_tiles; ++tile) {
    const int tile_first = tile * RMP2_TILE_SPHERES;
    const int tile_count = T.uses_spheres ? min(RMP2_TILE_SPHERES, O - tile_first) : 0;
    bool tile_ready = !kTma;
    if (kTma && tile_count > 0) {
      __syncwarp();                              // every lane is done reading the previous tile
      if (lane == 0) {
        const int boxes = (tile_count + 7) >> 3;
        mbar_expect_tx(wt.bar, static_cast<uint32_t>(boxes) * 4096u);
        for (int b = 0; b < boxes; ++b)
          tma_load_2d(const_cast<char*>(wt.base) + b * 4096, tmap, (tile_first + 8 * b) * 4,
                      static_cast<int32_t>(warp_env0), wt.bar);
      }
    }

    float zj[N][3], pj[N][3];
#pragma unroll
    for (int j = 0; j < N; ++j) zj[j][0] = zj[j][1] = zj[j][2] = pj[j][0] = pj[j][1] = pj[j][2] = 0.f;
    Chain ch;
    chain_reset(ch);

    for (int fi = 0; fi < T.n_frames; ++fi) {
      const FrameTab& F = T.frames[fi];
      if (F.restore_slot == RMP2_SLOT_BASE) {
        chain_reset(ch);
      } else if (F.restore_slot >= 0) {
        const float* s = slots + (size_t)F.restore_slot * RMP2_CHAIN_FLOATS * blockDim.x + threadIdx.x;
        float* cf = reinterpret_cast<float*>(&ch);
#pragma unroll
        for (int i = 0; i < RMP2_CHAIN_FLOATS; ++i) cf[i] = s[i * blockDim.x];
      }
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
      if (F.save_slot >= 0) {
        float* s = slots + (size_t)F.save_slot * RMP2_CHAIN_FLOATS * blockDim.x + threadIdx.x;
        const float* cf = reinterpret_cast<const float*>(&ch);
#pragma unroll
        for (int i = 0; i < RMP2_CHAIN_FLOATS; ++i) s[i * blockDim.x] = cf[i];
      }
      if (F.leaf_begin >= F.leaf_end) continue;

      float S[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      float g[3] = {0.f, 0.f, 0.f};
      bool contrib = false;
      const float vv = fmaf(ch.v[0], ch.v[0], fmaf(ch.v[1], ch.v[1], ch.v[2] * ch.v[2]));

      for (int li = F.leaf_begin; li < F.leaf_end; ++li) {
        const LeafTab& L = T.leaves[li];
        if (L.space == RMP2_SPACE_FRAME_POSITION) {
          if (tile != 0) continue;
          float goal[3], xdd[3], zeta[3], iso, dir;
#pragma unroll
          for (int i = 0; i < 3; ++i)
            goal[i] = (L.goal_slot >= 0) ? __ldg(A.goals + (env * A.n_goal_slots + L.goal_slot) * 3 + i)
                                         : T.vecpool[L.vec_off + i];
          if (L.type == RMP2_LEAF_TARGET_POLICY)
            target_policy<3>(L.p, ch.p, ch.v, goal, 3, xdd, zeta, iso, dir);
          else
            target_attractor(L.p, ch.p, ch.v, goal, xdd, zeta, iso, dir);
          const float e[3] = {xdd[0] - ch.a[0], xdd[1] - ch.a[1], xdd[2] - ch.a[2]};
          const float ze = dir * fmaf(zeta[0], e[0], fmaf(zeta[1], e[1], zeta[2] * e[2]));
          const float dz[3] = {dir * zeta[0], dir * zeta[1], dir * zeta[2]};
          S[0] += fmaf(dz[0], zeta[0], iso);
          S[1] = fmaf(dz[0], zeta[1], S[1]);
          S[2] = fmaf(dz[0], zeta[2], S[2]);
          S[3] += fmaf(dz[1], zeta[1], iso);
          S[4] = fmaf(dz[1], zeta[2], S[4]);
          S[5] += fmaf(dz[2], zeta[2], iso);
#pragma unroll
          for (int i = 0; i < 3; ++i) g[i] += fmaf(iso, e[i], ze * zeta[i]);
          contrib = true;
        } else if (L.space == RMP2_SPACE_FRAME_DISTANCE_SPHERES) {
          if (tile_count <= 0) continue;
          if (kTma && !tile_ready) {
            mbar_wait(wt.bar, wt.phase);
            wt.phase ^= 1u;
            tile_ready = true;
          }
          auto one_sphere = [&](const float4 sp) {
            // pos_on_link = frame origin; pos_on_obstacle = closest surface point of the sphere
            const float rx = ch.p[0] - sp.x, ry = ch.p[1] - sp.y, rz = ch.p[2] - sp.z;
            const float dc2 = fmaxf(fmaf(rx, rx, fmaf(ry, ry, rz * rz)), 1e-24f);
            const float inv_dc = fast_rsqrt(dc2);
            const float sd = fmaf(dc2, inv_dc, -sp.w);             // signed surface distance
            const float sgn = (sd < 0.f) ? -inv_dc : inv_dc;
            const float d = fmaxf(fabsf(sd), 1e-12f);
            obstacle_pair(L.p, rx * sgn, ry * sgn, rz * sgn, d, fast_rcp(d), ch.v, ch.a, vv, S, g);
          };
          if (kTma) {
            // row `lane` of each staged box holds 8 spheres of this environment; the 16-byte chunk c
            // of a row sits at chunk (c ^ (lane & 7)) (128-byte TMA swizzle) -> conflict-free LDS.128
            const uint32_t row = wt.base32 + lane * 128u;
            const uint32_t x7 = (lane & 7u) << 4;
            const int nbox = tile_count >> 3;
            for (int hb = 0; hb < 2 * nbox; ++hb) {          // half a box (4 spheres) per trip: the
#pragma unroll                                               // unrolled body stays inside the L0 I-cache
              for (int c4 = 0; c4 < 4; ++c4) {
                float4 sp;
                const uint32_t c8 = (uint32_t)(hb & 1) * 4u + (uint32_t)c4;
                const uint32_t addr = row + (uint32_t)(hb >> 1) * 4096u + ((c8 << 4) ^ x7);
                asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                             : "=f"(sp.x), "=f"(sp.y), "=f"(sp.z), "=f"(sp.w)
                             : "r"(addr));
                one_sphere(sp);
              }
            }
          } else {
            const float4* gsph = reinterpret_cast<const float4*>(A.spheres + ((size_t)env * O + tile_first) * 4);
#pragma unroll 2
            for (int o = 0; o < tile_count; ++o) one_sphere(__ldg(gsph + o));
          }
          contrib = true;
        } else {  // RMP2_SPACE_FRAME_DISTANCE_PAIRS: explicit (pos_on_link, pos_on_obstacle) pairs
          if (tile != 0) continue;
          const int k0 = A.pair_off[L.pair_set], k1 = A.pair_off[L.pair_set + 1];
          const float* pp = A.pairs + ((size_t)env * A.pair_total + k0) * 6;
          for (int k = 0; k < k1 - k0; ++k) {
            const float rx = __ldg(pp + 6 * k + 0) - __ldg(pp + 6 * k + 3);
            const float ry = __ldg(pp + 6 * k + 1) - __ldg(pp + 6 * k + 4);
            const float rz = __ldg(pp + 6 * k + 2) - __ldg(pp + 6 * k + 5);
            const float d2 = fmaxf(fmaf(rx, rx, fmaf(ry, ry, rz * rz)), 1e-24f);
            const float inv_d = fast_rsqrt(d2);
            obstacle_pair(L.p, rx * inv_d, ry * inv_d, rz * inv_d, d2 * inv_d, inv_d, ch.v, ch.a, vv, S, g);
          }
          contrib = true;
        }
      }
      if (contrib) pullback<N>(zj, pj, ch.p, F.anc_mask, T.prismatic_mask, S, g, Msym, f);
    }
    if (kTma && tile_count > 0 && !tile_ready) {   // tree asked for spheres but no frame consumed them
      mbar_wait(wt.bar, wt.phase);
      wt.phase ^= 1u;
    }
  }

  // ---- configuration-space leaves and the resolve, on the full matrix --------------------------
  float M[N][N];
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int j = 0; j < N; ++j) M[i][j] = (j <= i) ? Msym[i * (i + 1) / 2 + j] : Msym[j * (j + 1) / 2 + i];

  for (int li = T.n_frame_leaves; li < T.n_leaves; ++li) {
    const LeafTab& L = T.leaves[li];
    const float* vec = T.vecpool + L.vec_off;
    float xdd[N];
    if (L.type == RMP2_LEAF_CONFIG_BIASING || L.type == RMP2_LEAF_JOINT_DAMPING ||
        L.type == RMP2_LEAF_CSPACE_BIASING) {
      float m;
      if (L.type == RMP2_LEAF_CONFIG_BIASING)
        leaf_config_biasing<N>(L.p, vec, n, q, qd, xdd, m);
      else if (L.type == RMP2_LEAF_JOINT_DAMPING)
        leaf_joint_damping<N>(L.p, n, qd, xdd, m);
      else
        leaf_cspace_biasing<N>(L.p, vec, n, q, qd, xdd, m);
#pragma unroll
      for (int i = 0; i < N; ++i)
        if (i < n) {
          M[i][i] += m;
          f[i] = fmaf(m, xdd[i], f[i]);
        }
    } else if (L.type == RMP2_LEAF_VELOCITY_CAP) {
      float diag[N], w;
      leaf_velocity_cap<N>(L.p, n, qd, xdd, diag, w);
      float sum = 0.f;
#pragma unroll
      for (int i = 0; i < N; ++i) sum += xdd[i];
#pragma unroll
      for (int i = 0; i < N; ++i)
        if (i < n) {
          f[i] += fmaf(w, sum - xdd[i], diag[i] * xdd[i]);
#pragma unroll
          for (int j = 0; j < N; ++j)
            if (j < n) M[i][j] += (i == j) ? diag[i] : w;
        }
    } else if (L.type == RMP2_LEAF_JOINT_LIMIT) {
      float zeta[N], w[N];
      leaf_joint_limit<N>(L.p, vec, n, q, qd, xdd, zeta, w);
      const float beta = L.p[JL_BETA];
      float t = 0.f;
#pragma unroll
      for (int j = 0; j < N; ++j) t = fmaf(zeta[j] * w[j], xdd[j], t);
#pragma unroll
      for (int i = 0; i < N; ++i) {
        f[i] += fmaf(beta * zeta[i], t, (1.f - beta) * w[i] * xdd[i]);
#pragma unroll
        for (int j = 0; j < N; ++j) M[i][j] += fmaf(beta * zeta[i], zeta[j], (i == j) ? (1.f - beta) : 0.f) * w[j];
      }
    } else {  // RMP2_LEAF_TARGET_POLICY on the identity task map
      float goal[N], zeta[N], iso, dir;
#pragma unroll
      for (int i = 0; i < N; ++i) goal[i] = (i < n) ? vec[i] : 0.f;
      target_policy<N>(L.p, q, qd, goal, n, xdd, zeta, iso, dir);
      float zx = 0.f;
#pragma unroll
      for (int j = 0; j < N; ++j) zx = fmaf(zeta[j], xdd[j], zx);
#pragma unroll
      for (int i = 0; i < N; ++i)
        if (i < n) {
          f[i] += fmaf(iso, xdd[i], dir * zeta[i] * zx);
#pragma unroll
          for (int j = 0; j < N; ++j) M[i][j] += fmaf(dir * zeta[i], zeta[j], (i == j) ? iso : 0.f);
        }
    }
  }
  resolve_pinv<N>(M, f, T.rcond, qdd);
}

// ------------------------------------------------------------------------------------- step kernel
template <int N, bool kTma>
__global__ void __launch_bounds__(RMP2_BLOCK_THREADS)
    rmp2_step_kernel(const __grid_constant__ StepTables T, const __grid_constant__ StepArgs A,
                     const __grid_constant__ CUtensorMap tmap) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const int warps = blockDim.x >> 5;
  const int warp = threadIdx.x >> 5;
  const uint32_t lane = threadIdx.x & 31u;
  const long long env_raw = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long warp_env0 = env_raw - lane;
  if (warp_env0 >= A.B) return;                  // whole warp out of range (uniform per warp)
  const bool active = env_raw < A.B;
  const long long env = active ? env_raw : A.B - 1;
  const int n = T.n;

  // shared memory carve-up: [sphere tiles | mbarriers | chain-state slots]
  WarpTile wt;
  const int boxes = kTma ? min(4, (A.n_spheres + 7) >> 3) : 0;
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  wt.base = reinterpret_cast<const char*>(base + (size_t)warp * boxes * 4096);
  wt.base32 = smem_u32(wt.base);
  uint64_t* bars = reinterpret_cast<uint64_t*>(base + (size_t)warps * boxes * 4096);
  wt.bar = bars + warp;
  wt.phase = 0;
  wt.boxes = boxes;
  float* slots = reinterpret_cast<float*>(bars + warps);
  if (kTma) {
    if (lane == 0) mbar_init(wt.bar, 1);
    __syncwarp();
  }

  float q[N], qd[N], qdd[N];
#pragma unroll
  for (int j = 0; j < N; ++j) {
    const bool in = j < n;
    const float* qs = (A.n_sim_steps > 0) ? A.q_rw : A.q;
    const float* qds = (A.n_sim_steps > 0) ? A.qd_rw : A.qd;
    q[j] = in ? qs[env * n + j] : 0.f;
    qd[j] = in ? qds[env * n + j] : 0.f;
    qdd[j] = 0.f;
  }

  // One call site (one inlined copy of the step): a plain step is a rollout of one control step
  // without integration.  Closed-loop rollout = explicit Euler at dt, control every `control_every`.
  const bool rollout = A.n_sim_steps > 0;
  const int n_steps = rollout ? A.n_sim_steps : 1;
  for (int step = 0; step < n_steps; ++step) {
    if (!rollout || step % A.control_every == 0)
      evaluate_env<N, kTma>(T, A, &tmap, env, warp_env0, q, qd, slots, wt, qdd);
    if (rollout) {
#pragma unroll
      for (int j = 0; j < N; ++j) {
        qd[j] = fmaf(qdd[j], A.dt, qd[j]);
        q[j] = fmaf(qd[j], A.dt, q[j]);
      }
    }
  }
  if (rollout && active) {
#pragma unroll
    for (int j = 0; j < N; ++j)
      if (j < n) {
        A.q_rw[env * n + j] = q[j];
        A.qd_rw[env * n + j] = qd[j];
      }
  }
  if (active) {
#pragma unroll
    for (int j = 0; j < N; ++j)
      if (j < n) A.qdd[env * n + j] = qdd[j];
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
                     const float* __restrict__ xin, const float* __restrict__ xdin, float* __restrict__ xdd_out,
                     float* __restrict__ M_out) {
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
template <int N>
static cudaError_t launch_step_n(const StepTables& T, const StepArgs& A, const CUtensorMap* tmap, bool use_tma,
                                 int block, size_t smem, cudaStream_t stream) {
  const long long blocks = (A.B + block - 1) / block;
  if (blocks <= 0) return cudaSuccess;
  CUtensorMap dummy;
  memset(&dummy, 0, sizeof(dummy));
  if (use_tma) {
    cudaError_t e = cudaFuncSetAttribute(rmp2_step_kernel<N, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    rmp2_step_kernel<N, true><<<(unsigned)blocks, block, smem, stream>>>(T, A, *tmap);
  } else {
    cudaError_t e = cudaFuncSetAttribute(rmp2_step_kernel<N, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    rmp2_step_kernel<N, false><<<(unsigned)blocks, block, smem, stream>>>(T, A, dummy);
  }
  return cudaGetLastError();
}

int rmp2_pick_width(int n) {
  if (n <= 2) return 2;
  if (n <= 7) return 7;
  if (n <= 9) return 9;
  return 12;
}

cudaError_t rmp2_launch_step(const StepTables& T, const StepArgs& A, const CUtensorMap* tmap, bool use_tma,
                             int block, size_t smem, cudaStream_t stream) {
  switch (rmp2_pick_width(T.n)) {
    case 2: return launch_step_n<2>(T, A, tmap, use_tma, block, smem, stream);
    case 7: return launch_step_n<7>(T, A, tmap, use_tma, block, smem, stream);
    case 9: return launch_step_n<9>(T, A, tmap, use_tma, block, smem, stream);
    default: return launch_step_n<12>(T, A, tmap, use_tma, block, smem, stream);
  }
}

template <int N>
static cudaError_t step_attr_n(bool use_tma, cudaFuncAttributes* attr, int block, size_t smem, int* blocks_per_sm) {
  const void* fn = use_tma ? (const void*)rmp2_step_kernel<N, true> : (const void*)rmp2_step_kernel<N, false>;
  cudaError_t e = cudaFuncGetAttributes(attr, fn);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  if (use_tma)
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, rmp2_step_kernel<N, true>, block, smem);
  return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, rmp2_step_kernel<N, false>, block, smem);
}

cudaError_t rmp2_step_attributes(int n, bool use_tma, int block, size_t smem, int* regs, int* static_smem,
                                 int* blocks_per_sm) {
  cudaFuncAttributes attr;
  cudaError_t e;
  switch (rmp2_pick_width(n)) {
    case 2: e = step_attr_n<2>(use_tma, &attr, block, smem, blocks_per_sm); break;
    case 7: e = step_attr_n<7>(use_tma, &attr, block, smem, blocks_per_sm); break;
    case 9: e = step_attr_n<9>(use_tma, &attr, block, smem, blocks_per_sm); break;
    default: e = step_attr_n<12>(use_tma, &attr, block, smem, blocks_per_sm); break;
  }
  if (e != cudaSuccess) return e;
  *regs = attr.numRegs;
  *static_smem = (int)attr.sharedSizeBytes;
  return cudaSuccess;
}

cudaError_t rmp2_launch_fk(const StepTables& T, long long B, const float* q, const float* qd, float* x, float* xd,
                           float* J, float* c, cudaStream_t stream) {
  const long long blocks = (B + 127) / 128;
  if (blocks <= 0) return cudaSuccess;
  switch (rmp2_pick_width(T.n)) {
    case 2: rmp2_fk_kernel<2><<<(unsigned)blocks, 128, 0, stream>>>(T, B, q, qd, x, xd, J, c); break;
    case 7: rmp2_fk_kernel<7><<<(unsigned)blocks, 128, 0, stream>>>(T, B, q, qd, x, xd, J, c); break;
    case 9: rmp2_fk_kernel<9><<<(unsigned)blocks, 128, 0, stream>>>(T, B, q, qd, x, xd, J, c); break;
    default: rmp2_fk_kernel<12><<<(unsigned)blocks, 128, 0, stream>>>(T, B, q, qd, x, xd, J, c); break;
  }
  return cudaGetLastError();
}

cudaError_t rmp2_launch_leaf(const LeafTab& L, const LeafVec& V, int m, long long K, const float* x, const float* xd,
                             float* xdd, float* M, cudaStream_t stream) {
  const long long blocks = (K + 127) / 128;
  if (blocks <= 0) return cudaSuccess;
  rmp2_leaf_kernel<<<(unsigned)blocks, 128, 0, stream>>>(L, V, m, K, x, xd, xdd, M);
  return cudaGetLastError();
}
