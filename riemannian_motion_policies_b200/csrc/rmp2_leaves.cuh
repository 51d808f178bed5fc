// Leaf policies as device functions (float32).  Each cites the reference lines it restates.
// The derived parameter layout p[] is produced on the host by derive_leaf_params()
// (rmp2_api.cu) from the raw constructor arguments documented in include/rmp2_b200.h.
#pragma once
#include "rmp2_tables.h"

#define RMP2_DEV __device__ __forceinline__

// ---- derived parameter slots ------------------------------------------------------------------
// TARGET_POLICY
#define TP_ALPHA 0
#define TP_BETA 1
#define TP_C 2
#define TP_INV_C 3
// CONFIG_BIASING
#define CB_GAMMA_P 0
#define CB_GAMMA_D 1
#define CB_W 2
// JOINT_LIMIT   (vecpool: lower[n], upper[n], 1/(upper-lower)[n])
#define JL_GAMMA_P 0
#define JL_GAMMA_D 1
#define JL_C3 2
#define JL_C2 3
#define JL_R 4
#define JL_INV_QDMAX 5
#define JL_BETA 6
#define JL_C 7
#define JL_INV_C 8
// TARGET_ATTRACTOR
#define TA_PGAIN 0
#define TA_DGAIN 1
#define TA_EPS 2
#define TA_EPS10 3
#define TA_INV_ALEN 4
#define TA_MIN_ALPHA 5
#define TA_SMAX 6
#define TA_SMIN 7
#define TA_BOOST 8
#define TA_INV_BLEN 9
// VELOCITY_CAP
#define VC_CUTOFF 0
#define VC_GAIN 1
#define VC_CLIP 2
#define VC_INV_REGION 3
#define VC_WEIGHT 4
// JOINT_DAMPING
#define JD_GAIN 0
#define JD_SCALAR 1
#define JD_INERTIA 2
// OBSTACLE_AVOIDANCE
#define OA_MARGIN 0
#define OA_R 1
#define OA_INV_R 2
#define OA_MSCALAR 3
#define OA_INV_ESTD 4
#define OA_EEPS 5
#define OA_RGAIN 6
#define OA_INV_RSTD 7
#define OA_INV_VLEN 8
#define OA_DGAIN 9
#define OA_INV_DSTD 10
#define OA_DEPS 11
#define OA_K_REP 12       // -log2(e) / repulsion_std_dev
#define OA_K_VEL 13       //  log2(e) / damping_velocity_gate_length_scale
#define OA_G_SCALE 14     // sphere path: what the step kernel multiplies the pair kernel's raw force sums by: weight / k^2
// CSPACE_BIASING
#define CS_METRIC 0
#define CS_PGAIN 1
#define CS_DGAIN 2
#define CS_THRESH 3

// COLLISION_AVOIDANCE (v1)
#define CA_ETA_REP 0
#define CA_INV_NU_REP 1
#define CA_ETA_DAMP 2
#define CA_INV_NU_DAMP 3
#define CA_R 4
#define CA_C3 5
#define CA_C2 6

// ---- single-instruction special functions (MUFU), used in the per-pair hot loop -----------------
// Relative error ~1e-7 each (2 ulp); the distance leaf tolerates that: see DESIGN.md "numerics".
RMP2_DEV float fast_rcp(float x) {                                               // MUFU.RCP, x normal
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
RMP2_DEV float fast_rsqrt(float x) {                                             // MUFU.RSQ, x normal
  float y;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
RMP2_DEV float fast_exp2(float x) {                                              // MUFU.EX2
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// exp(x) = 2^(x log2e) with the rounding error of the product folded back in (x up to ~ +-90)
RMP2_DEV float fast_exp(float x) {
  const float t = x * 1.4426950408889634f;
  const float e = fmaf(x, 1.4426950408889634f, -t) + x * 1.9259629911266175e-8f;   // low part
  return fast_exp2(t) * fmaf(e, 0.6931471805599453f, 1.f);
}

// ---- metrics of the form  A = iso * I + dir * zeta zeta^T  ----------------------------------------
// TargetPolicy  (reference: rmp.py:241-260, helper/rmp_helper.py:62-74)
template <int D>
RMP2_DEV void target_policy(const float* __restrict__ p, const float (&x)[D], const float (&xd)[D],
                            const float (&goal)[D], int dim, float (&xdd)[D], float (&zeta)[D],
                            float& iso, float& dir) {
  float v[D];
  float nv2 = 0.f;
#pragma unroll
  for (int i = 0; i < D; ++i) {
    v[i] = (i < dim) ? goal[i] - x[i] : 0.f;
    nv2 = fmaf(v[i], v[i], nv2);
  }
  const float nv = sqrtf(nv2);                                           // tf.norm(v)   rmp.py:243
  const float c = p[TP_C];
  const float h = nv + c * log1pf(expf(-2.f * c * nv));                  // c*log(...)   rmp.py:244
  const float inv_h = 1.f / h;
  float nx2 = 0.f;
#pragma unroll
  for (int i = 0; i < D; ++i) {
    xdd[i] = (i < dim) ? p[TP_ALPHA] * (inv_h * v[i]) - p[TP_BETA] * xd[i] : 0.f;   // rmp.py:245-246
    nx2 = fmaf(xdd[i], xdd[i], nx2);
  }
  const float beta = 1.f - expf(-0.5f * nv2);                            // sigma_H = 1  rmp.py:253
  const float nx = sqrtf(nx2);
  const float hh = nx + p[TP_INV_C] * log1pf(expf(-2.f * c * nx));       // soft_norm    rmp_helper.py:64
  const float inv_hh = 1.f / hh;
#pragma unroll
  for (int i = 0; i < D; ++i) zeta[i] = xdd[i] * inv_hh;
  const float w = expf(-nv * (1.f / 3.f));                               // sigma_w = 3  rmp.py:257
  iso = w * (1.f - beta);
  dir = w * beta;
}

// TargetAttractor  (reference: rmp2.py:52-83)
RMP2_DEV void target_attractor(const float* __restrict__ p, const float (&x)[3], const float (&xd)[3],
                               const float (&goal)[3], float (&xdd)[3], float (&zeta)[3], float& iso,
                               float& dir) {
  float d[3];
  float dn2 = 0.f;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    d[i] = goal[i] - x[i];
    dn2 = fmaf(d[i], d[i], dn2);
  }
  const float dn = sqrtf(dn2);
  const float inv_soft = 1.f / fmaxf(dn, p[TA_EPS10]);                   // rmp2.py:68-69
  const float inv_acc = 1.f / (dn + p[TA_EPS]);                          // rmp2.py:58
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    xdd[i] = p[TA_PGAIN] * d[i] * inv_acc - p[TA_DGAIN] * xd[i];
    zeta[i] = d[i] * inv_soft;
  }
  const float sd = dn * p[TA_INV_ALEN];
  const float a = (1.f - p[TA_MIN_ALPHA]) * expf(-0.5f * sd * sd) + p[TA_MIN_ALPHA];   // rmp2.py:74
  const float bd = dn * p[TA_INV_BLEN];
  const float ba = expf(-0.5f * bd * bd);                                               // rmp2.py:79
  const float boost = ba * p[TA_BOOST] + (1.f - ba);                                    // rmp2.py:80
  iso = boost * (a * p[TA_SMAX]);                                                       // rmp2.py:76,82
  dir = boost * ((1.f - a) * p[TA_SMIN]);
}

// ---- orientation task map: xyz Euler angles of a frame ------------------------------------------------
// theta = (theta_x, theta_y, theta_z) with R = R_z(theta_z) R_y(theta_y) R_x(theta_x)  (reference:
// kinematics.py:74-96: theta_y = -asin(r20), theta_z = atan2(r10, r00), theta_x = atan2(r21, r22)).  The
// reference differentiates this through the 4x4 matrix entries by autodiff (taskmap.py:57-67); on rotation
// matrices that is the Euler-rate map of helper/trigonometry_helper.py:18-38, omega = H(theta) thetadot:
//     thetadot = E omega,   E = H^-1 = 1/cb [[cg, sg, 0], [-sg cb, cg cb, 0], [cg sb, sg sb, cb]]
//     c = d/dt(E omega) at zero joint acceleration = Edot omega + E alpha      (w, al: Chain::w, Chain::al)
// R row-major 3x3.  E row-major.
RMP2_DEV void euler_map(const float* __restrict__ R, const float (&w)[3], const float (&al)[3], float (&th)[3],
                        float (&thd)[3], float (&cc)[3], float (&E)[9]) {
  const float r00 = R[0], r10 = R[3], r20 = R[6], r21 = R[7], r22 = R[8];
  th[0] = atan2f(r21, r22);
  th[1] = -asinf(fminf(fmaxf(r20, -1.f), 1.f));
  th[2] = atan2f(r10, r00);
  const float sb = -r20;
  const float cb = fmaxf(sqrtf(fmaxf(fmaf(-r20, r20, 1.f), 0.f)), 1e-6f);   // cos(theta_y) >= 0; gimbal-lock guard
  const float inv_cb = 1.f / cb;
  const float inv_h = rsqrtf(fmaxf(fmaf(r00, r00, r10 * r10), 1e-24f));
  const float cg = r00 * inv_h, sg = r10 * inv_h;
  const float u = fmaf(cg, w[0], sg * w[1]), v = fmaf(cg, w[1], -sg * w[0]);
  const float ad = u * inv_cb, bd = v, gd = fmaf(sb, ad, w[2]);
  thd[0] = ad;
  thd[1] = bd;
  thd[2] = gd;
  const float ud = fmaf(cg, al[0], fmaf(sg, al[1], gd * v));
  const float vd = fmaf(cg, al[1], fmaf(-sg, al[0], -gd * u));
  const float add = fmaf(ad * sb, bd, ud) * inv_cb;
  cc[0] = add;
  cc[1] = vd;
  cc[2] = fmaf(sb, add, fmaf(cb * bd, ad, al[2]));
  E[0] = cg * inv_cb, E[1] = sg * inv_cb, E[2] = 0.f;
  E[3] = -sg, E[4] = cg, E[5] = 0.f;
  E[6] = cg * sb * inv_cb, E[7] = sg * sb * inv_cb, E[8] = 1.f;
}

// ---- ObstacleAvoidance on one closest-point pair --------------------------------------------------
// n = unit vector obstacle -> link, d = distance (the task coordinate x of the reference),
// inv_d = 1/d.  v, a = velocity and Jdot*qd of the frame origin, vv = |v|^2.
// Distance map (reference: taskmap.py:120-138 via autodiff): xdot = n.v,
// c = n.a + (|v|^2 - xdot^2)/d.  Leaf (reference: rmp2.py:184-196).
// Accumulates S += M n n^T (xx, xy, xz, yy, yz, zz) and g += M (xdd - c) n, which is the
// per-pair pullback of rmp.py:165-167 onto the frame origin.
RMP2_DEV void obstacle_pair(const float* __restrict__ p, float nx, float ny, float nz, float d,
                            float inv_d, const float (&v)[3], const float (&a)[3], float vv,
                            float (&S)[6], float (&g)[3]) {
  const float xdot = fmaf(nx, v[0], fmaf(ny, v[1], nz * v[2]));
  const float curv = fmaf(-xdot, xdot, vv) * inv_d;
  const float c = fmaf(nx, a[0], fmaf(ny, a[1], fmaf(nz, a[2], curv)));
  const float x = fmaxf(d - p[OA_MARGIN], 0.f);                          // rmp2.py:185-186
  // one reciprocal for the two leaf denominators (MUFU is the co-limiting pipe of the pair loop):
  // 1/den1 = den2 / (den1 den2), 1/den2 = den1 / (den1 den2)
  const float den1 = fmaf(x, p[OA_INV_ESTD], p[OA_EEPS]);
  const float den2 = fmaf(x, p[OA_INV_DSTD], p[OA_DEPS]);
  const float r12 = fast_rcp(den1 * den2);
  const float inv1 = r12 * den2, inv2 = r12 * den1;
  const float base = p[OA_MSCALAR] * inv1;                               // rmp2.py:187
  const float gt = fmaf(x, p[OA_INV_R], -1.f);
  const float gate = gt * gt;                                            // rmp2.py:172
  const float rep = p[OA_RGAIN] * fast_exp2(x * p[OA_K_REP]);            // rmp2.py:189
  const float one_minus_sig = fast_rcp(1.f + fast_exp2(xdot * p[OA_K_VEL]));     // 1 - sigmoid  rmp2.py:190
  const float damp = -one_minus_sig * p[OA_DGAIN] * xdot * inv2;         // rmp2.py:191
  const float acc = rep + damp;
  const float m = (x > p[OA_R]) ? 0.f : one_minus_sig * base * gate;     // rmp2.py:194
  const float h = m * (acc - c);
  const float mx = m * nx, my = m * ny, mz = m * nz;
  S[0] = fmaf(mx, nx, S[0]);
  S[1] = fmaf(mx, ny, S[1]);
  S[2] = fmaf(mx, nz, S[2]);
  S[3] = fmaf(my, ny, S[3]);
  S[4] = fmaf(my, nz, S[4]);
  S[5] = fmaf(mz, nz, S[5]);
  g[0] = fmaf(h, nx, g[0]);
  g[1] = fmaf(h, ny, g[1]);
  g[2] = fmaf(h, nz, g[2]);
}

// ---- packed form of the pair, two spheres per thread step (rmp2_spheres_kernel) ---------------------
// Blackwell issues fma/mul/add.f32x2 (SASS FFMA2) on register pairs: half the issue slots of the scalar
// form for the same FP32-lane work, which lets the MUFU pipe run underneath it.  Measured on a B200
// (tools/pipe_peaks.cu): FFMA2 sustains 124-128 FMA/clk/SM with MUFU overlapped; scalar FFMA 97-107;
// and every ALU-pipe instruction (FMNMX, FSEL, FSETP, LOP3) takes the same lane time as a scalar FP32
// instruction away from FFMA2.  So the cost of a pair is its count of lane operations, FP32 and ALU
// alike, and the leaf is arranged to minimise that count (49 lane operations and 5 MUFU per pair;
// the first scalar version had 61 and 6):
//   xs = sat((d - margin)/r) in ONE FFMA.SAT replaces max(.,0), the x > r test and its select:
//        gate = (xs - 1)^2 is exactly 0 beyond the radius (rmp2.py:170-174, 194), r is folded into the
//        coefficients that multiply x;
//   one reciprocal q = 1 / (den1' den2' (1 + e_v)) gives  (1-sigmoid) m0 = q den2'  and
//        D (1-sigmoid)/den2 = q den1'   (den1' = den1 / metric_scalar, den2' = den2 / damping_gain,
//        e_v = exp(xdot / l_v)): the damping gain D rides on the reciprocal for free;
//   xdd - c' = G e_rep - (q den1') xdot - curv  as two FMAs, c' = c - n.a;
//   the n.a term of the curvature c (three FMAs per pair) never enters the loop:
//        sum_o m (n.a) n = (sum_o m n n^T) a = S a   -- the caller subtracts S a once per (environment, leaf);
//   the velocity enters pre-scaled, v' = k v with k = log2(e)/l_v, so xdot' = n.v' is the exponent of e_v as it
//        stands; everything downstream that is quadratic in the velocity carries k^2 (|v'|^2, the damping term through
//        den2' = den2 / (D k), the repulsion through G' = G k^2) and the caller divides g by k^2 once;
//   g += (xdd - c')(m n): the product with m is shared with the metric's m n.
// 41 lane operations per pair (round 1: 61, then 49, 47, 43).
// Sphere-row parameters, derived on the host in double (fill_sphere_row, rmp2_api.cu):
#define SP_XA 0           // 1 / r
#define SP_XB 1           // -margin / r
#define SP_GT_A 2         // 1            (0 when metric_scalar == 0: the leaf's metric vanishes)
#define SP_GT_B 3         // -1           (0 when metric_scalar == 0)
#define SP_D1A 4          // r / (metric_exploder_std_dev * metric_scalar)
#define SP_D1B 5          // metric_exploder_eps / metric_scalar
#define SP_D2A 6          // r / (damping_std_dev * damping_gain * k)      (0 when damping_gain == 0)
#define SP_D2B 7          // damping_robustness_eps / (damping_gain * k)   (2^60 when damping_gain == 0: no damping term)
#define SP_K_VEL 8        // k = log2(e) / damping_velocity_gate_length_scale
#define SP_K_REP 9        // -log2(e) r / repulsion_std_dev
#define SP_RGAIN 10       // repulsion_gain * k^2
#define SP_REACH 11       // (r + margin) * (1 + 1e-5): conservative bound of the early-out test
#define SP_INV_K2 12      // 1 / k^2
#define SP_COUNT 13        // parameters the pair loop keeps in registers
#define SP_WEIGHT 13       // number of obstacle leaves this row stands for (coincident control points merged), read once at the end

RMP2_DEV float2 bc2(float s) { return make_float2(s, s); }
RMP2_DEV float2 neg2(float2 a) { return make_float2(-a.x, -a.y); }
RMP2_DEV float fma_sat(float a, float b, float c) {                               // FFMA.SAT: clamp to [0, 1]
  float y;
  asm("fma.rn.sat.f32 %0, %1, %2, %3;" : "=f"(y) : "f"(a), "f"(b), "f"(c));
  return y;
}

// n = unit vectors obstacle -> link of the two pairs, d = distances, inv_d = 1/d; vk = k v (velocity of the frame
// origin, pre-scaled), vvk = |k v|^2.  S[i], g[i]: lane x accumulates the even spheres of the environment, lane y the
// odd ones.  g is accumulated as k^2 (g + S a): the caller finishes with g = g / k^2 - S a.
RMP2_DEV void obstacle_pair2(const float* __restrict__ p, float2 nx, float2 ny, float2 nz, float2 d, float2 inv_d,
                             const float (&vk)[3], float vvk, float2 (&S)[6], float2 (&g)[3]) {
  const float2 xdot = __ffma2_rn(nx, bc2(vk[0]), __ffma2_rn(ny, bc2(vk[1]), __fmul2_rn(nz, bc2(vk[2]))));   // k n.v
  const float2 curv = __fmul2_rn(__ffma2_rn(neg2(xdot), xdot, bc2(vvk)), inv_d);  // k^2 (|v|^2 - xdot^2)/d  taskmap.py:120-138
  // xs = clamp((d - margin) / r, 0, 1)                                           rmp2.py:185-186, 170-174
  const float2 xs = make_float2(fma_sat(d.x, p[SP_XA], p[SP_XB]), fma_sat(d.y, p[SP_XA], p[SP_XB]));
  const float2 den1 = __ffma2_rn(xs, bc2(p[SP_D1A]), bc2(p[SP_D1B]));           // (x/s_e + eps_e) / scalar  rmp2.py:187
  const float2 den2 = __ffma2_rn(xs, bc2(p[SP_D2A]), bc2(p[SP_D2B]));           // (x/s_d + eps_d) / (D k)   rmp2.py:191
  const float2 ev = make_float2(fast_exp2(xdot.x), fast_exp2(xdot.y));           // 1/(1 - sigmoid) = 1 + ev  rmp2.py:190
  const float2 den12 = __fmul2_rn(den1, den2);
  const float2 Q = __ffma2_rn(den12, ev, den12);
  const float2 q = make_float2(fast_rcp(Q.x), fast_rcp(Q.y));                    // 0 when ev overflowed: no metric, no force
  const float2 w1 = __fmul2_rn(q, den2);                                          // (1-sig) scalar / (x/s_e + eps_e)
  const float2 w2 = __fmul2_rn(q, den1);                                          // (1-sig) D k / (x/s_d + eps_d)
  const float2 gt = __ffma2_rn(xs, bc2(p[SP_GT_A]), bc2(p[SP_GT_B]));           // x/r - 1, 0 beyond the radius
  const float2 m = __fmul2_rn(w1, __fmul2_rn(gt, gt));                            // rmp2.py:172,194
  const float2 tr = __fmul2_rn(xs, bc2(p[SP_K_REP]));
  const float2 er = make_float2(fast_exp2(tr.x), fast_exp2(tr.y));               // rmp2.py:189
  const float2 ac = __ffma2_rn(bc2(p[SP_RGAIN]), er, neg2(__ffma2_rn(w2, xdot, curv)));   // k^2 (xdd - (c - n.a))
  const float2 mx = __fmul2_rn(m, nx), my = __fmul2_rn(m, ny), mz = __fmul2_rn(m, nz);
  S[0] = __ffma2_rn(mx, nx, S[0]);
  S[1] = __ffma2_rn(mx, ny, S[1]);
  S[2] = __ffma2_rn(mx, nz, S[2]);
  S[3] = __ffma2_rn(my, ny, S[3]);
  S[4] = __ffma2_rn(my, nz, S[4]);
  S[5] = __ffma2_rn(mz, nz, S[5]);
  g[0] = __ffma2_rn(ac, mx, g[0]);
  g[1] = __ffma2_rn(ac, my, g[1]);
  g[2] = __ffma2_rn(ac, mz, g[2]);
}

// scalar form used by rmp2_leaf_evaluate: x, xd -> xdd, M
RMP2_DEV void obstacle_scalar(const float* __restrict__ p, float xin, float xdot, float& xdd, float& M) {
  const float x = fmaxf(xin - p[OA_MARGIN], 0.f);
  const float base = p[OA_MSCALAR] / fmaf(x, p[OA_INV_ESTD], p[OA_EEPS]);
  const float gt = fmaf(x, p[OA_INV_R], -1.f);
  const float rep = p[OA_RGAIN] * fast_exp2(x * p[OA_K_REP]);
  const float one_minus_sig = 1.f / (1.f + fast_exp2(xdot * p[OA_K_VEL]));
  const float damp = -one_minus_sig * p[OA_DGAIN] * xdot / fmaf(x, p[OA_INV_DSTD], p[OA_DEPS]);
  xdd = rep + damp;
  M = (x > p[OA_R]) ? 0.f : one_minus_sig * base * (gt * gt);
}

// CollisionAvoidance, v1 (reference: rmp.py:283-315).  d, vec = distance and unit normal of the pair
// (external data), xd = velocity of the point.  The metric's directional stretching is multiplied by
// beta = 0 (rmp.py:312), so A = w * I exactly.
RMP2_DEV void collision_avoidance_v1(const float* __restrict__ p, float d, const float (&vec)[3],
                                     const float (&xd)[3], float (&f)[3], float& w) {
  const float alpha_rep = p[CA_ETA_REP] * expf(-d * p[CA_INV_NU_REP]);            // rmp.py:285
  const float alpha_damp = p[CA_ETA_DAMP] / (d * p[CA_INV_NU_DAMP] + 1e-6f);      // rmp.py:289-290
  const float vx = fmaf(vec[0], xd[0], fmaf(vec[1], xd[1], vec[2] * xd[2]));      // vec . xd
  const float scaling = fmaxf(0.f, -vx);                                          // rmp.py:291
  const float damp = alpha_damp * scaling * vx;                                   // P_obs xd = scaling vec (vec.xd)
#pragma unroll
  for (int i = 0; i < 3; ++i) f[i] = (alpha_rep - damp) * vec[i];                 // rmp.py:286,293-295
  const float spline = fmaf(p[CA_C3] * d, d * d, fmaf(p[CA_C2] * d, d, 1.f));     // rmp.py:301-305
  w = (d > p[CA_R]) ? 0.f : spline;                                               // rmp.py:306
}

// ---- configuration-space leaves: add A into the full matrix M and A*xdd into f ---------------------
// All take q, qd (first n entries valid), M row-major [N][N], f [N].
template <int N>
RMP2_DEV void leaf_config_biasing(const float* __restrict__ p, const float* __restrict__ q0, int n,
                                  const float (&q)[N], const float (&qd)[N], float (&xdd)[N], float& m) {
  // reference: rmp.py:330-347
#pragma unroll
  for (int i = 0; i < N; ++i)
    xdd[i] = (i < n) ? p[CB_GAMMA_P] * (q0[i] - q[i]) - p[CB_GAMMA_D] * qd[i] : 0.f;
  m = p[CB_W];
}

template <int N>
RMP2_DEV void leaf_joint_damping(const float* __restrict__ p, int n, const float (&qd)[N], float (&xdd)[N],
                                 float& m) {
  // reference: rmp2.py:127-137
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < N; ++i) s = fmaf(qd[i], qd[i], s);     // entries >= n are zero
  const float nq = sqrtf(s);
  const float gain = p[JD_GAIN] * nq;
#pragma unroll
  for (int i = 0; i < N; ++i) xdd[i] = -gain * qd[i];
  m = fmaf(p[JD_SCALAR], nq, p[JD_INERTIA]);
}

template <int N>
RMP2_DEV void leaf_cspace_biasing(const float* __restrict__ p, const float* __restrict__ goal, int n,
                                  const float (&q)[N], const float (&qd)[N], float (&xdd)[N], float& m) {
  // reference: rmp2.py:212-226
  float e[N];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < N; ++i) {
    e[i] = (i < n) ? q[i] - goal[i] : 0.f;
    s = fmaf(e[i], e[i], s);
  }
  const float ne = sqrtf(s);
  const bool near = ne < p[CS_THRESH];
  const float scale = near ? -p[CS_PGAIN] : -p[CS_THRESH] * p[CS_PGAIN] / ne;
#pragma unroll
  for (int i = 0; i < N; ++i) xdd[i] = fmaf(scale, e[i], -p[CS_DGAIN] * qd[i]);
  m = p[CS_METRIC];
}

// JointVelocityCap: A_ii = w / (1 - ratio_i^2), A_ij = w  (reference: rmp2.py:100-112)
template <int N>
RMP2_DEV void leaf_velocity_cap(const float* __restrict__ p, int n, const float (&qd)[N], float (&xdd)[N],
                                float (&diag)[N], float& w) {
  w = p[VC_WEIGHT];
#pragma unroll
  for (int i = 0; i < N; ++i) {
    const float av = fabsf(qd[i]);
    const float dv = av - p[VC_CUTOFF];
    const float sgn = (qd[i] > 0.f) ? 1.f : ((qd[i] < 0.f) ? -1.f : 0.f);
    const float acc = -fabsf(p[VC_GAIN] * dv) * sgn;
    xdd[i] = (i < n && !(av < p[VC_CUTOFF])) ? acc : 0.f;
    const float ratio = fminf(dv, p[VC_CLIP]) * p[VC_INV_REGION];
    diag[i] = (i < n) ? w / (1.f - ratio * ratio) : 0.f;
  }
}

// JointLimitAvoidance: A = (beta zeta zeta^T + (1-beta) I) diag(w)  (reference: rmp.py:357-382)
template <int N>
RMP2_DEV void leaf_joint_limit(const float* __restrict__ p, const float* __restrict__ vec, int n,
                               const float (&q)[N], const float (&qd)[N], float (&xdd)[N], float (&zeta)[N],
                               float (&w)[N]) {
  float s = 0.f;
  float vv[N];
#pragma unroll
  for (int i = 0; i < N; ++i) {
    if (i < n) {
      const float lo = vec[i], up = vec[n + i], inv_range = vec[2 * n + i];
      const float d = fminf((up - q[i]) * inv_range, (q[i] - lo) * inv_range);
      const float spline = fmaf(p[JL_C3] * d, d * d, fmaf(p[JL_C2] * d, d, 1.f));
      w[i] = (d > p[JL_R]) ? 0.f : spline;
      vv[i] = qd[i] * p[JL_INV_QDMAX];
      xdd[i] = -p[JL_GAMMA_P] * q[i] - p[JL_GAMMA_D] * qd[i];
    } else {
      w[i] = 0.f;
      vv[i] = 0.f;
      xdd[i] = 0.f;
    }
    s = fmaf(vv[i], vv[i], s);
  }
  const float nv = sqrtf(s);
  const float hh = nv + p[JL_INV_C] * log1pf(expf(-2.f * p[JL_C] * nv));
  const float inv_hh = 1.f / hh;
#pragma unroll
  for (int i = 0; i < N; ++i) zeta[i] = vv[i] * inv_hh;
}
