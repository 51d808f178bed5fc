// Device code of the two table-driven kernels of the control step -- the frames kernel and the step kernel --
// as function bodies over a `StepTables`.  Compiled twice:
//   * ahead of time (rmp2_kernels.cu): generic kernels, the tables arrive as a __grid_constant__ parameter and
//     every frame / leaf is interpreted at run time;
//   * at run time by NVRTC (rmp2_jit.cu, rmp2_tree_specialize): the tables of ONE tree are a compile-time
//     constant, RMP2_UNROLL_SPEC unrolls the frame and leaf loops, and every table-driven branch, joint
//     selection and constant transform folds away (config 4: 7100 -> 4600 executed instructions per
//     environment in the step kernel, 3000 -> 1800 in the frames kernel).
// Nothing in here may need a host header: NVRTC compiles it without an include path.
#pragma once
#include "rmp2_step.cuh"

#ifndef RMP2_BLOCK_THREADS
#define RMP2_BLOCK_THREADS 128
#endif

// Threads per block as a kernel body sees it: the launch's own value (kBlock = 0), or a compile-time constant when a
// specialised kernel is built for one block size -- the shared-memory offsets of the chain-state slots and joint columns
// are immediates then (config 4: 6090 -> 5710 instructions in the specialised step kernel).
#define RMP2_BLOCKDIM (kBlock ? kBlock : (int)blockDim.x)

#ifdef RMP2_JIT
#define RMP2_UNROLL_SPEC _Pragma("unroll")
#else
#define RMP2_UNROLL_SPEC
#endif

// ------------------------------------------------------------------------------ chain walking
// Visit frame `fi` of the depth-first execution list: restore / advance / save the chain state and,
// when kCols, record the world axis and origin of the joint column the frame drives.
template <int N, bool kCols, int kBlock = 0>
RMP2_DEV void visit_frame(const StepTables& T, int fi, const float (&q)[N], const float (&qd)[N], Chain& ch,
                          float* cols, float* slots) {
  const FrameTab& F = T.frames[fi];
  if (F.restore_slot == RMP2_SLOT_BASE) {
    chain_reset(ch);
  } else if (F.restore_slot >= 0) {
    const float* s = slots + (size_t)F.restore_slot * RMP2_CHAIN_FLOATS * RMP2_BLOCKDIM + threadIdx.x;
    float* cf = reinterpret_cast<float*>(&ch);
#pragma unroll
    for (int i = 0; i < RMP2_CHAIN_FLOATS; ++i) cf[i] = s[i * RMP2_BLOCKDIM];
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
  if (kCols && F.qidx >= 0) {                    // joint column -> shared memory (see pullback)
    float* c = cols + (size_t)F.qidx * 6 * RMP2_BLOCKDIM;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      c[i * RMP2_BLOCKDIM] = z[i];
      c[(3 + i) * RMP2_BLOCKDIM] = ch.p[i];
    }
  }
  if (F.save_slot >= 0) {
    float* s = slots + (size_t)F.save_slot * RMP2_CHAIN_FLOATS * RMP2_BLOCKDIM + threadIdx.x;
    const float* cf = reinterpret_cast<const float*>(&ch);
#pragma unroll
    for (int i = 0; i < RMP2_CHAIN_FLOATS; ++i) s[i * RMP2_BLOCKDIM] = cf[i];
  }
}

// ------------------------------------------------------------------------------- frames kernel
// rec[field][slot][env] = (p, v) of the frame origin for every sphere-obstacle leaf slot (fields 0..5; the pair kernel
// overwrites fields 0..8 with its sums).  a = Jdot qd of the origin is not stored: only the step kernel needs it (for
// the S a part of the curvature term), and it walks the chain itself.
template <int N, int kBlock = 0>
RMP2_DEV void frames_body(const StepTables& T, const StepArgs& A) {
  extern __shared__ float slots[];
  const long long env = (long long)blockIdx.x * RMP2_BLOCKDIM + threadIdx.x;
  if (env >= A.B) return;
  const int n = T.n;
  float q[N], qd[N];
#pragma unroll
  for (int j = 0; j < N; ++j) {
    q[j] = (j < n) ? A.q[env * n + j] : 0.f;
    qd[j] = (j < n) ? A.qd[env * n + j] : 0.f;
  }
  Chain ch;
  chain_reset(ch);
  // tiled record scratch (rmp2_rec_base): consecutive environments are consecutive floats -> every store below is
  // one full line per warp, at an immediate offset from one base pointer in the specialised kernel
  float* rec = A.rec + rmp2_rec_base(env, T.n_sphere_slots);
  const int fstride = T.n_sphere_slots * RMP2_REC_TILE;
  RMP2_UNROLL_SPEC
  for (int fi = 0; fi < T.n_frames; ++fi) {
    visit_frame<N, false, kBlock>(T, fi, q, qd, ch, nullptr, slots);
    const FrameTab& F = T.frames[fi];
    RMP2_UNROLL_SPEC
    for (int li = F.leaf_begin; li < F.leaf_end; ++li) {
      const LeafTab& L = T.leaves[li];
      if (L.space != RMP2_SPACE_FRAME_DISTANCE_SPHERES) continue;
      float* r = rec + L.sphere_slot * RMP2_REC_TILE;
      r[0 * fstride] = ch.p[0];
      r[1 * fstride] = ch.p[1];
      r[2 * fstride] = ch.p[2];
      r[3 * fstride] = ch.v[0];
      r[4 * fstride] = ch.v[1];
      r[5 * fstride] = ch.v[2];
    }
  }
}

// --------------------------------------------------------------------------------- step kernel
// Tail shared by the fused step kernel and the resolve kernel: optional explicit-Euler sub-steps with
// the command held (reference loop: control at 10 Hz, simulation at 100 Hz --
// experiments/franka_panda/05_obstacle_avoidance.py:92-97), then the stores.
template <int N>
RMP2_DEV void finish_step(const StepArgs& A, int n, long long e, bool active, bool rollout, float (&q)[N],
                          float (&qd)[N], const float (&qdd)[N]) {
  if (rollout) {
    for (int s = 0; s < A.n_sim_steps; ++s) {
#pragma unroll
      for (int j = 0; j < N; ++j) {
        qd[j] = fmaf(qdd[j], A.dt, qd[j]);
        q[j] = fmaf(qd[j], A.dt, q[j]);
      }
    }
    if (active) {
#pragma unroll
      for (int j = 0; j < N; ++j)
        if (j < n) {
          A.q_rw[e * n + j] = q[j];
          A.qd_rw[e * n + j] = qd[j];
        }
    }
  }
  if (active) {
#pragma unroll
    for (int j = 0; j < N; ++j)
      if (j < n) A.qdd[e * n + j] = qdd[j];
  }
}

// ------------------------------------------------------------------------------ resolve hand-off
// The resolve runs in two stages.  Stage 1 (here, inside the step kernel or rmp2_resolve_kernel): factorise and
// solve directly where the matrix is provably clear of the pinv cutoff (resolve_direct) -- every environment of
// the trees with an isotropic metric leaf, ~97 % of the rank-deficient config-4 tree.  Environments that do not
// qualify are appended to a work list (one atomic per warp) together with their factorised problem
// [R | Q^T f | perm], written over the environment's own column of the (M, f) scratch; stage 2
// (rmp2_resolve_fallback_kernel) runs the Jacobi sweeps for exactly those.  The ~30 KB of unrolled Jacobi code,
// its registers and its warp divergence stay out of the kernels every environment passes through.
//   A.fb: [0] = list length, [1] = block ticket of the fallback kernel, [2 ...] = environment indices
#define RMP2_HANDOFF_FIELDS(N) ((N) * ((N) + 1) / 2 + 2 * (N))
// Fused step kernel (A.split == 0): the mf scratch holds nothing else, and an environment's problem is ONE contiguous,
// 16-byte aligned row of RMP2_HANDOFF_ROW(N) floats at mf + e * ROW -- a handful of 128-bit stores here, and a handful
// of lines (one or two pages) per environment in the fallback kernel, whose lanes gather scattered environments: with
// the field-major layout its 42 loads per lane each touched another 4 MB plane of the scratch and took 40 % of that
// kernel's time (ncu: long_scoreboard at its head).  Split mode keeps the field-major layout [k][B] written over the
// environment's own column: there the scratch still holds other environments' unread (M, f).
#define RMP2_HANDOFF_ROW(N) ((RMP2_HANDOFF_FIELDS(N) + 3) / 4 * 4)

template <int N, bool kFieldMajor>     // kFieldMajor: called from the stand-alone resolve kernel (split mode, A.split == 1)
RMP2_DEV void defer_to_fallback(const StepArgs& A, long long e, bool need, const float (&G)[N][N],
                                const float (&y)[N], const int (&perm)[N]) {
  const unsigned ballot = __ballot_sync(0xffffffffu, need);
  if (ballot == 0u) return;
  const int lane = threadIdx.x & 31;
  const int leader = __ffs((int)ballot) - 1;
  int base = 0;
  if (lane == leader) base = atomicAdd(A.fb, __popc(ballot));
  base = __shfl_sync(0xffffffffu, base, leader);
  if (!need) return;
  A.fb[2 + base + __popc(ballot & ((1u << lane) - 1u))] = (int)e;
  if (!kFieldMajor) {
    float row[RMP2_HANDOFF_ROW(N)];
    int k = 0;
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
      for (int j = i; j < N; ++j) row[k++] = G[i][j];
#pragma unroll
    for (int i = 0; i < N; ++i) row[k++] = y[i];
#pragma unroll
    for (int i = 0; i < N; ++i) row[k++] = __int_as_float(perm[i]);
#pragma unroll
    for (; k < RMP2_HANDOFF_ROW(N); ++k) row[k] = 0.f;
    float4* o4 = reinterpret_cast<float4*>(A.mf + (size_t)e * RMP2_HANDOFF_ROW(N));
#pragma unroll
    for (int i = 0; i < RMP2_HANDOFF_ROW(N) / 4; ++i) o4[i] = make_float4(row[4 * i], row[4 * i + 1], row[4 * i + 2], row[4 * i + 3]);
    return;
  }
  float* o = A.mf + e;
  int k = 0;
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int j = i; j < N; ++j) o[(size_t)(k++) * A.B] = G[i][j];
#pragma unroll
  for (int i = 0; i < N; ++i) o[(size_t)(k++) * A.B] = y[i];
#pragma unroll
  for (int i = 0; i < N; ++i) o[(size_t)(k++) * A.B] = __int_as_float(perm[i]);
}

template <int N, bool kQr, bool kFieldMajor = false>
RMP2_DEV void resolve_or_defer(const StepArgs& A, float (&M)[N][N], float (&f)[N], int n, float rcond, long long e,
                               bool active, bool rollout, float (&q)[N], float (&qd)[N]) {
  int perm[N];
  float qdd[N];
  const bool solved = resolve_direct<N, kQr>(M, f, perm, n, rcond, qdd);
  defer_to_fallback<N, kFieldMajor>(A, e, active && !solved, M, f, perm);
  finish_step<N>(A, n, e, active && solved, rollout, q, qd, qdd);
}

#ifndef RMP2_RESOLVE_MIN_BLOCKS
#define RMP2_RESOLVE_MIN_BLOCKS(N) ((N) <= 7 ? 4 : ((N) <= 9 ? 3 : 2))
#endif
#ifndef RMP2_FALLBACK_MIN_BLOCKS
#define RMP2_FALLBACK_MIN_BLOCKS(N) ((N) <= 7 ? 6 : ((N) <= 9 ? 3 : 2))
#endif
#ifndef RMP2_SPLIT_MIN_BLOCKS
#define RMP2_SPLIT_MIN_BLOCKS(N) ((N) <= 7 ? 4 : ((N) <= 9 ? 3 : 2))
#endif
// resident blocks per SM the register allocation aims at (N <= 7: 128 registers -> 16 warps/SM)
#ifndef RMP2_STEP_MIN_BLOCKS
#define RMP2_STEP_MIN_BLOCKS(N) ((N) <= 7 ? 4 : ((N) <= 9 ? 3 : 2))
#endif
// The tree-specialised step kernels are ~110 KB of straight-line code and their top stall is instruction fetch
// (ncu: no_instruction 2.5 of 6.8 stalled warps per issue).  Warps of one block start together and stay close, so a
// fetched line serves several of them: 256-thread blocks (two warps per SM sub-partition and block, two blocks per
// SM -- the same 128-register budget as four blocks of 128) run the config-4 step kernel in 0.270 ms instead of
// 0.303; 512-thread blocks 0.283 (one block per SM: no latency diversity), block barriers between the frames
// (forced lock step) cost more than they save (0.283 at 256, 0.311 at 512).  Upper bound of the launch only: small
// batches still launch 32 / 64 / 128 threads.  N > 7 keeps 128 (its register budget needs 3 or 2 blocks per SM).
#ifndef RMP2_SPEC_STEP_THREADS
#define RMP2_SPEC_STEP_THREADS(N) ((N) <= 7 ? 256 : RMP2_BLOCK_THREADS)
#endif
#ifndef RMP2_SPEC_STEP_MIN_BLOCKS
#define RMP2_SPEC_STEP_MIN_BLOCKS(N) ((N) <= 7 ? 2 : RMP2_STEP_MIN_BLOCKS(N))
#endif
#ifndef RMP2_SPEC_SPLIT_MIN_BLOCKS
#define RMP2_SPEC_SPLIT_MIN_BLOCKS(N) ((N) <= 7 ? 2 : RMP2_SPLIT_MIN_BLOCKS(N))
#endif
// kSplit: stop after the combined (M, f) and hand them to rmp2_resolve_kernel through A.mf
// (field-major [N*N + N][B]); used for large batches, where two small kernels beat one big one.
template <int N, bool kSplit, int kBlock = 0>
RMP2_DEV void step_body(const StepTables& T, const StepArgs& A) {
  extern __shared__ float slots[];
  const long long env = (long long)blockIdx.x * RMP2_BLOCKDIM + threadIdx.x;
  // no early return: the resolve uses full-warp votes; out-of-range lanes redo the last environment
  const bool active = env < A.B;
  const long long e = active ? env : A.B - 1;
  const int n = T.n;
  const bool rollout = A.n_sim_steps > 0;

  float q[N], qd[N], qdd[N];
#pragma unroll
  for (int j = 0; j < N; ++j) {
    const bool in = j < n;
    q[j] = in ? (rollout ? A.q_rw : A.q)[e * n + j] : 0.f;
    qd[j] = in ? (rollout ? A.qd_rw : A.qd)[e * n + j] : 0.f;
  }

  float Msym[N * (N + 1) / 2];
  float f[N];
#pragma unroll
  for (int i = 0; i < N * (N + 1) / 2; ++i) Msym[i] = 0.f;
#pragma unroll
  for (int i = 0; i < N; ++i) f[i] = 0.f;

  {
    // shared memory: [chain-state slots | joint columns (6 N floats per thread)]
    float* cols = slots + (size_t)T.n_slots * RMP2_CHAIN_FLOATS * RMP2_BLOCKDIM + threadIdx.x;
    const int cstride = RMP2_BLOCKDIM;
    Chain ch;
    chain_reset(ch);
#ifdef RMP2_JIT
    // Specialised build: the (S, g) sums of the next frame's sphere leaves (written by rmp2_spheres_kernel)
    // are pulled into L1 while this frame is processed -- consumed cold they are this kernel's top stall.
    // (In the generic kernel the table walk this needs costs more than it saves.)
    auto prefetch_sums = [&](int fn) {
      if (A.n_spheres <= 0 || fn >= T.n_frames) return;
      RMP2_UNROLL_SPEC
      for (int li = T.frames[fn].leaf_begin; li < T.frames[fn].leaf_end; ++li)
        if (T.leaves[li].space == RMP2_SPACE_FRAME_DISTANCE_SPHERES) {
          const float* r = A.rec + rmp2_rec_base(e, T.n_sphere_slots) + T.leaves[li].sphere_slot * RMP2_REC_TILE;
          const int fs = T.n_sphere_slots * RMP2_REC_TILE;
#pragma unroll
          for (int i = 0; i < 9; ++i) asm volatile("prefetch.global.L1 [%0];" ::"l"(r + i * fs));
        }
    };
    prefetch_sums(0);
#endif
    RMP2_UNROLL_SPEC
    for (int fi = 0; fi < T.n_frames; ++fi) {
#ifdef RMP2_JIT
      prefetch_sums(fi + 1);
#endif
      visit_frame<N, true, kBlock>(T, fi, q, qd, ch, cols, slots);
      const FrameTab& F = T.frames[fi];
      if (F.leaf_begin >= F.leaf_end) continue;

      float S[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      float g[3] = {0.f, 0.f, 0.f};
      bool contrib = false;
      RMP2_UNROLL_SPEC
      for (int li = F.leaf_begin; li < F.leaf_end; ++li) {
        const LeafTab& L = T.leaves[li];
        if (L.space == RMP2_SPACE_FRAME_POSITION) {
          float goal[3], xdd[3], zeta[3], iso, dir;
#pragma unroll
          for (int i = 0; i < 3; ++i)
            goal[i] = (L.goal_slot >= 0) ? __ldg(A.goals + (e * A.n_goal_slots + L.goal_slot) * 3 + i)
                                         : T.vecpool[L.vec_off + i];
          if (L.type == RMP2_LEAF_TARGET_POLICY)
            target_policy<3>(L.p, ch.p, ch.v, goal, 3, xdd, zeta, iso, dir);
          else
            target_attractor(L.p, ch.p, ch.v, goal, xdd, zeta, iso, dir);
          const float er[3] = {xdd[0] - ch.a[0], xdd[1] - ch.a[1], xdd[2] - ch.a[2]};
          const float ze = dir * fmaf(zeta[0], er[0], fmaf(zeta[1], er[1], zeta[2] * er[2]));
          const float dz[3] = {dir * zeta[0], dir * zeta[1], dir * zeta[2]};
          S[0] += fmaf(dz[0], zeta[0], iso);
          S[1] = fmaf(dz[0], zeta[1], S[1]);
          S[2] = fmaf(dz[0], zeta[2], S[2]);
          S[3] += fmaf(dz[1], zeta[1], iso);
          S[4] = fmaf(dz[1], zeta[2], S[4]);
          S[5] += fmaf(dz[2], zeta[2], iso);
#pragma unroll
          for (int i = 0; i < 3; ++i) g[i] += fmaf(iso, er[i], ze * zeta[i]);
          contrib = true;
        } else if (L.space == RMP2_SPACE_FRAME_EULER) {
          // orientation leaf: x = Euler angles, J = E J_omega (J_omega[:, j] = z_j for revolute joints on the
          // path, 0 for prismatic ones), so the pullback runs on the angular columns with
          // S' = E^T A E and g' = E^T A (xdd - c), A = iso I + dir zeta zeta^T
          float th[3], thd[3], cc[3], E[9], goal[3], xdd[3], zeta[3], iso, dir;
          euler_map(ch.R, ch.w, ch.al, th, thd, cc, E);
#pragma unroll
          for (int i = 0; i < 3; ++i)
            goal[i] = (L.goal_slot >= 0) ? __ldg(A.goals + (e * A.n_goal_slots + L.goal_slot) * 3 + i)
                                         : T.vecpool[L.vec_off + i];
          if (L.type == RMP2_LEAF_TARGET_POLICY)
            target_policy<3>(L.p, th, thd, goal, 3, xdd, zeta, iso, dir);
          else
            target_attractor(L.p, th, thd, goal, xdd, zeta, iso, dir);
          const float er[3] = {xdd[0] - cc[0], xdd[1] - cc[1], xdd[2] - cc[2]};
          float Ez[3], Ee[3];                                       // E^T zeta, E^T er
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            Ez[i] = fmaf(E[i], zeta[0], fmaf(E[3 + i], zeta[1], E[6 + i] * zeta[2]));
            Ee[i] = fmaf(E[i], er[0], fmaf(E[3 + i], er[1], E[6 + i] * er[2]));
          }
          const float ze = dir * fmaf(zeta[0], er[0], fmaf(zeta[1], er[1], zeta[2] * er[2]));
          float Se[6], ge[3];
          int k = 0;
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            ge[i] = fmaf(iso, Ee[i], ze * Ez[i]);
#pragma unroll
            for (int j = i; j < 3; ++j)
              Se[k++] = fmaf(iso, fmaf(E[i], E[j], fmaf(E[3 + i], E[3 + j], E[6 + i] * E[6 + j])), dir * Ez[i] * Ez[j]);
          }
          const float origin[3] = {0.f, 0.f, 0.f};
          pullback<N>(cols, cstride, origin, F.anc_mask & ~T.prismatic_mask, 0xffffffffu, Se, ge, Msym, f);
        } else if (L.space == RMP2_SPACE_FRAME_DISTANCE_SPHERES) {
          if (A.n_spheres <= 0) continue;
          // sums over this leaf's spheres, produced by rmp2_spheres_kernel
          const float* r = A.rec + rmp2_rec_base(e, T.n_sphere_slots) + L.sphere_slot * RMP2_REC_TILE;
          const int fstride = T.n_sphere_slots * RMP2_REC_TILE;
          // fields 0..5: S = weight * sum_o m n n^T; 6..8: the raw force sums k^2 (g + S a)/weight (obstacle_pair2 leaves
          // the n.a part of the curvature term out of its loop: sum_o m (n.a) n = S a, subtracted here, where a = Jdot qd
          // of this frame's origin is at hand)
          float Ss[6], gs[3];
#pragma unroll
          for (int i = 0; i < 6; ++i) Ss[i] = __ldg(r + i * fstride);
#pragma unroll
          for (int i = 0; i < 3; ++i) gs[i] = __ldg(r + (6 + i) * fstride);
          const float sc = L.p[OA_G_SCALE];
          g[0] += fmaf(gs[0], sc, -fmaf(Ss[0], ch.a[0], fmaf(Ss[1], ch.a[1], Ss[2] * ch.a[2])));
          g[1] += fmaf(gs[1], sc, -fmaf(Ss[1], ch.a[0], fmaf(Ss[3], ch.a[1], Ss[4] * ch.a[2])));
          g[2] += fmaf(gs[2], sc, -fmaf(Ss[2], ch.a[0], fmaf(Ss[4], ch.a[1], Ss[5] * ch.a[2])));
#pragma unroll
          for (int i = 0; i < 6; ++i) S[i] += Ss[i];
          contrib = true;
        } else if (L.space == RMP2_SPACE_FRAME_DISTANCE_PAIRS) {  // explicit (pos_on_link, pos_on_obstacle)
          const int k0 = A.pair_off[L.pair_set], k1 = A.pair_off[L.pair_set + 1];
          const float* pp = A.pairs + ((size_t)e * A.pair_total + k0) * RMP2_PAIR_FLOATS;
          const float vv = fmaf(ch.v[0], ch.v[0], fmaf(ch.v[1], ch.v[1], ch.v[2] * ch.v[2]));
          for (int k = 0; k < k1 - k0; ++k) {
            const float* row = pp + RMP2_PAIR_FLOATS * k;
            const float rx = __ldg(row + 0) - __ldg(row + 3);
            const float ry = __ldg(row + 1) - __ldg(row + 4);
            const float rz = __ldg(row + 2) - __ldg(row + 5);
            const float d2 = fmaxf(fmaf(rx, rx, fmaf(ry, ry, rz * rz)), 1e-24f);
            const float inv_d = fast_rsqrt(d2);
            obstacle_pair(L.p, rx * inv_d, ry * inv_d, rz * inv_d, d2 * inv_d, inv_d, ch.v, ch.a, vv, S, g);
          }
          contrib = true;
        } else {  // RMP2_SPACE_FRAME_POINTS: points fixed in the frame (v1 CollisionAvoidance)
          // x = p + R rel;  xd = v + w x rho;  c = a + alpha x rho + w x (w x rho)   (rho = R rel);
          // the point's Jacobian is the frame-origin Jacobian evaluated at x, so every pair is
          // pulled back on its own (reference: taskmap.py:83-99 via autodiff).
          const int k0 = A.pair_off[L.pair_set], k1 = A.pair_off[L.pair_set + 1];
          const float* pp = A.pairs + ((size_t)e * A.pair_total + k0) * RMP2_PAIR_FLOATS;
          for (int k = 0; k < k1 - k0; ++k) {
            const float* row = pp + RMP2_PAIR_FLOATS * k;
            const float rel[3] = {__ldg(row + 0), __ldg(row + 1), __ldg(row + 2)};
            const float dist = __ldg(row + 3);
            const float vec[3] = {__ldg(row + 4), __ldg(row + 5), __ldg(row + 6)};
            float rho[3], wr[3], wwr[3], ar[3];
            matvec3(ch.R, rel, rho);
            cross3(ch.w, rho, wr);
            cross3(ch.w, wr, wwr);
            cross3(ch.al, rho, ar);
            const float xp[3] = {ch.p[0] + rho[0], ch.p[1] + rho[1], ch.p[2] + rho[2]};
            const float xd[3] = {ch.v[0] + wr[0], ch.v[1] + wr[1], ch.v[2] + wr[2]};
            const float cp[3] = {ch.a[0] + ar[0] + wwr[0], ch.a[1] + ar[1] + wwr[1], ch.a[2] + ar[2] + wwr[2]};
            float fl[3], w;
            collision_avoidance_v1(L.p, dist, vec, xd, fl, w);
            const float Sp[6] = {w, 0.f, 0.f, w, 0.f, w};
            const float gp[3] = {w * (fl[0] - cp[0]), w * (fl[1] - cp[1]), w * (fl[2] - cp[2])};
            pullback<N>(cols, cstride, xp, F.anc_mask, T.prismatic_mask, Sp, gp, Msym, f);
          }
        }
      }
      if (contrib) pullback<N>(cols, cstride, ch.p, F.anc_mask, T.prismatic_mask, S, g, Msym, f);
    }
  }

  // ---- configuration-space leaves and the resolve, on the full matrix --------------------------
  float M[N][N];
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int j = 0; j < N; ++j) M[i][j] = (j <= i) ? Msym[i * (i + 1) / 2 + j] : Msym[j * (j + 1) / 2 + i];

  RMP2_UNROLL_SPEC
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
      float tt = 0.f;
#pragma unroll
      for (int j = 0; j < N; ++j) tt = fmaf(zeta[j] * w[j], xdd[j], tt);
#pragma unroll
      for (int i = 0; i < N; ++i) {
        f[i] += fmaf(beta * zeta[i], tt, (1.f - beta) * w[i] * xdd[i]);
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
  if (kSplit) {
    if (active) {
      float* o = A.mf + e;
#pragma unroll
      for (int i = 0; i < N; ++i)
#pragma unroll
        for (int j = 0; j < N; ++j) o[(size_t)(i * N + j) * A.B] = M[i][j];
#pragma unroll
      for (int i = 0; i < N; ++i) o[(size_t)(N * N + i) * A.B] = f[i];
    }
    return;
  }
  if (T.precondition)                            // same arithmetic as the stand-alone resolve kernel
    resolve_or_defer<N, true>(A, M, f, n, T.rcond, e, active, rollout, q, qd);
  else
    resolve_or_defer<N, false>(A, M, f, n, T.rcond, e, active, rollout, q, qd);
}
