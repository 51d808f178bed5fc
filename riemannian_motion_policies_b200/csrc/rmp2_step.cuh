// Per-environment device code of the control step: kinematic chain recursion with analytic
// velocity / Jdot*qd, pullback to configuration space, and the truncated-SVD resolve.
// One thread owns one environment; every array below lives in registers (all indices are
// compile-time after unrolling).
#pragma once
#include "rmp2_leaves.cuh"

// --------------------------------------------------------------------------------------------------
// Kinematic chain state of the frame most recently visited.
//   R, p : world rotation and origin            (reference: kinematics.py:235-246, the chain product)
//   w, v : angular velocity, origin velocity    (xd = J qd,  kinematics.py:265)
//   al, a: angular / origin acceleration at zero joint acceleration, i.e. Jdot qd (kinematics.py:267)
// The reference obtains v, J and a by TensorFlow autodiff of the 4x4 product; here they are the
// closed-form rigid-body recursion (verified against autodiff in tests/test_oracle_kinematics.py).
// --------------------------------------------------------------------------------------------------
struct Chain {
  float R[9], p[3], w[3], v[3], al[3], a[3];
};

RMP2_DEV void chain_reset(Chain& c) {
#pragma unroll
  for (int i = 0; i < 9; ++i) c.R[i] = (i % 4 == 0) ? 1.f : 0.f;
#pragma unroll
  for (int i = 0; i < 3; ++i) c.p[i] = c.w[i] = c.v[i] = c.al[i] = c.a[i] = 0.f;
}

RMP2_DEV void cross3(const float* a, const float* b, float* o) {
  o[0] = fmaf(a[1], b[2], -a[2] * b[1]);
  o[1] = fmaf(a[2], b[0], -a[0] * b[2]);
  o[2] = fmaf(a[0], b[1], -a[1] * b[0]);
}

RMP2_DEV void matvec3(const float* R, const float* x, float* o) {
#pragma unroll
  for (int i = 0; i < 3; ++i) o[i] = fmaf(R[3 * i], x[0], fmaf(R[3 * i + 1], x[1], R[3 * i + 2] * x[2]));
}

RMP2_DEV void matmul3(const float* A, const float* B, float* C) {
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      C[3 * i + j] = fmaf(A[3 * i], B[j], fmaf(A[3 * i + 1], B[3 + j], A[3 * i + 2] * B[6 + j]));
}

// Advance the chain across one URDF joint: T_child = T_parent * T_const * T_var(q_i).
// z receives the joint axis in world coordinates (valid for movable joints).
RMP2_DEV void chain_advance(Chain& c, const FrameTab& F, float qi, float qdi, float (&z)[3]) {
  float rho[3], Rc[9];
  matvec3(c.R, F.t, rho);                       // offset parent origin -> joint origin, world
  if (F.const_rot_identity) {
#pragma unroll
    for (int i = 0; i < 9; ++i) Rc[i] = c.R[i];
  } else {
    matmul3(c.R, F.R, Rc);
  }
  z[0] = z[1] = z[2] = 0.f;
  if (F.type != RMP2_JOINT_FIXED) matvec3(Rc, F.axis, z);
  if (F.type == RMP2_JOINT_PRISMATIC) {         // T_var = translation q * axis  (kinematics.py:231-233)
#pragma unroll
    for (int i = 0; i < 3; ++i) rho[i] = fmaf(z[i], qi, rho[i]);
  }
  float wr[3], wwr[3], ar[3];
  cross3(c.w, rho, wr);
  cross3(c.w, wr, wwr);
  cross3(c.al, rho, ar);
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    c.p[i] += rho[i];
    c.v[i] += wr[i];
    c.a[i] += ar[i] + wwr[i];
  }
  if (F.type == RMP2_JOINT_PRISMATIC) {
    float wz[3];
    cross3(c.w, z, wz);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      c.v[i] = fmaf(z[i], qdi, c.v[i]);
      c.a[i] = fmaf(2.f * qdi, wz[i], c.a[i]);
    }
  }
  if (F.type == RMP2_JOINT_REVOLUTE) {          // Rodrigues, unit axis (kinematics.py:99-121)
    float s, co;
    sincosf(qi, &s, &co);
    if (F.axis_is_z) {                           // R_new = Rc * R_z(q): mixes columns 0 and 1
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const float a0 = Rc[3 * r], a1 = Rc[3 * r + 1];
        c.R[3 * r] = fmaf(co, a0, s * a1);
        c.R[3 * r + 1] = fmaf(co, a1, -s * a0);
        c.R[3 * r + 2] = Rc[3 * r + 2];
      }
    } else {
      const float t = 1.f - co;
      const float ux = F.axis[0], uy = F.axis[1], uz = F.axis[2];
      float Rv[9];
      Rv[0] = fmaf(t * ux, ux, co);
      Rv[1] = fmaf(t * ux, uy, -s * uz);
      Rv[2] = fmaf(t * ux, uz, s * uy);
      Rv[3] = fmaf(t * uy, ux, s * uz);
      Rv[4] = fmaf(t * uy, uy, co);
      Rv[5] = fmaf(t * uy, uz, -s * ux);
      Rv[6] = fmaf(t * uz, ux, -s * uy);
      Rv[7] = fmaf(t * uz, uy, s * ux);
      Rv[8] = fmaf(t * uz, uz, co);
      matmul3(Rc, Rv, c.R);
    }
    float wz[3];
    cross3(c.w, z, wz);                          // uses the parent's angular velocity
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      c.al[i] = fmaf(wz[i], qdi, c.al[i]);
      c.w[i] = fmaf(z[i], qdi, c.w[i]);
    }
  } else {
#pragma unroll
    for (int i = 0; i < 9; ++i) c.R[i] = Rc[i];
  }
}

// --------------------------------------------------------------------------------------------------
// Pullback of everything attached to one frame origin (reference: rmp.py:165-167, summed over the
// pairs/leaves of the frame as rmp.py:149-150 does afterwards):
//   M += J^T S J,   f += J^T g,   J[:, j] = z_j x (p_k - p_j)  (revolute) | z_j (prismatic)
// S symmetric 3x3 as (xx, xy, xz, yy, yz, zz).  Msym holds the lower triangle, row-major.
// --------------------------------------------------------------------------------------------------
// Joint columns live in shared memory, one float per (joint, component, thread):
//   cols[(j * 6 + i) * stride] = z_j[i] (i < 3) | p_j[i - 3] (i >= 3), `cols` already offset by threadIdx.x
// (conflict-free, ~30-cycle loads, and 6n registers less than keeping them per thread).
template <int N>
RMP2_DEV void pullback(const float* __restrict__ cols, int stride, const float (&pk)[3], uint32_t anc,
                       uint32_t prismatic, const float (&S)[6], const float (&g)[3],
                       float (&Msym)[N * (N + 1) / 2], float (&f)[N]) {
  // One pass over the joint columns: column i is built, multiplied by S (u_i = S J_i, kept for the rows
  // below) and contracted with every u_j, j <= i, at once -- only u stays live, not the columns.
  float u[N][3];
#pragma unroll
  for (int i = 0; i < N; ++i) {
    u[i][0] = u[i][1] = u[i][2] = 0.f;
    if (anc & (1u << i)) {                       // warp-uniform
      float col[3];
      const float z[3] = {cols[(i * 6 + 0) * stride], cols[(i * 6 + 1) * stride], cols[(i * 6 + 2) * stride]};
      if (prismatic & (1u << i)) {
        col[0] = z[0];
        col[1] = z[1];
        col[2] = z[2];
      } else {
        const float r[3] = {pk[0] - cols[(i * 6 + 3) * stride], pk[1] - cols[(i * 6 + 4) * stride],
                            pk[2] - cols[(i * 6 + 5) * stride]};
        cross3(z, r, col);
      }
      u[i][0] = fmaf(S[0], col[0], fmaf(S[1], col[1], S[2] * col[2]));
      u[i][1] = fmaf(S[1], col[0], fmaf(S[3], col[1], S[4] * col[2]));
      u[i][2] = fmaf(S[2], col[0], fmaf(S[4], col[1], S[5] * col[2]));
      f[i] = fmaf(col[0], g[0], fmaf(col[1], g[1], fmaf(col[2], g[2], f[i])));
#pragma unroll
      for (int j = 0; j <= i; ++j) {             // u_j = 0 for joints off the path
        const int idx = i * (i + 1) / 2 + j;
        Msym[idx] = fmaf(col[0], u[j][0], fmaf(col[1], u[j][1], fmaf(col[2], u[j][2], Msym[idx])));
      }
    }
  }
}

// --------------------------------------------------------------------------------------------------
// Resolve: x = pinv(M) f with tf.linalg.pinv's default cutoff (reference: rmp.py:153-154;
// TensorFlow 2.10 linalg_impl.pinv: singular values <= 10*max(rows,cols)*eps * sigma_max are dropped).
// One-sided Jacobi on the ROWS of the augmented matrix [M | f]:  rotations Q^T are applied from
// the left until the rows g_i of G = Q^T M are mutually orthogonal; then g_i = sigma_i v_i^T,
// y = Q^T f, and  pinv(M) f = sum_{sigma_i > cutoff} g_i^T y_i / sigma_i^2.  No U or V is stored.
// M may be non-symmetric (joint-limit leaf) or indefinite (velocity-cap leaf).
// --------------------------------------------------------------------------------------------------
#ifndef RMP2_JACOBI_MAX_SWEEPS
#define RMP2_JACOBI_MAX_SWEEPS 16
#endif
#define RMP2_JACOBI_TOL 2.4e-7f          // relative orthogonality |g| <= tol sqrt(a b)  (4 eps32)
#define RMP2_JACOBI_ANGLE 1e-8f          // see below
#define RMP2_JACOBI_DROP (1.f / 16.f)    // see below

// Round-robin (circle method) schedule: round r, slot k -> the pair (p, q).  Consecutive slots of a
// round touch disjoint rows, so the serial scalar part of one rotation overlaps the row updates of
// the previous one.  MM = N rounded up to even; pairs that involve the dummy row MM-1 >= N are skipped.
template <int N>
struct JacobiSchedule {
  static constexpr int MM = (N % 2 == 0) ? N : N + 1;
  static constexpr int rounds = MM - 1;
  static constexpr int slots = MM / 2;
  __host__ __device__ static constexpr int first(int r, int k) {
    return k == 0 ? (MM - 1) : (r + k) % (MM - 1);
  }
  __host__ __device__ static constexpr int second(int r, int k) {
    return k == 0 ? r : (r - k + (MM - 1)) % (MM - 1);
  }
};

// Householder QR with column pivoting, in place on the augmented matrix:  [G | y] <- [R | Q^T y],
// columns of G permuted so that |R_11| >= |R_22| >= ...;  perm[j] = original column now at position j.
// Purpose: preconditioner.  The rows of R are graded and R R^T is far closer to diagonal than M M^T,
// which cuts the Jacobi sweeps on rank-deficient trees from ~7 to ~4 (Drmac & Veselic, SIMAX 2008, use
// the same device).  pinv(M) f = P pinv(R) Q^T f, so the truncation rule is unchanged.
template <int N>
RMP2_DEV void qr_column_pivoting(float (&G)[N][N], float (&y)[N], int (&perm)[N]) {
  float cn[N];                                   // squared norms of the trailing part of each column
#pragma unroll
  for (int j = 0; j < N; ++j) {
    perm[j] = j;
    float a = 0.f;
#pragma unroll
    for (int i = 0; i < N; ++i) a = fmaf(G[i][j], G[i][j], a);
    cn[j] = a;
  }
#pragma unroll
  for (int k = 0; k < N - 1; ++k) {
    int piv = k;
    float best = cn[k];
#pragma unroll
    for (int j = k + 1; j < N; ++j)
      if (cn[j] > best) {
        best = cn[j];
        piv = j;
      }
#pragma unroll
    for (int j = k + 1; j < N; ++j) {
      const bool sw = (piv == j);
#pragma unroll
      for (int i = 0; i < N; ++i) {
        const float a = G[i][k], b = G[i][j];
        G[i][k] = sw ? b : a;
        G[i][j] = sw ? a : b;
      }
      const float ca = cn[k], cb = cn[j];
      cn[k] = sw ? cb : ca;
      cn[j] = sw ? ca : cb;
      const int pa = perm[k], pb = perm[j];
      perm[k] = sw ? pb : pa;
      perm[j] = sw ? pa : pb;
    }
    float nrm2 = 0.f;
#pragma unroll
    for (int i = k; i < N; ++i) nrm2 = fmaf(G[i][k], G[i][k], nrm2);
    const float nrm = sqrtf(nrm2);
    const float x0 = G[k][k];
    const float alpha = (x0 > 0.f) ? -nrm : nrm;             // R_kk
    const float v0 = x0 - alpha;                              // v = x - alpha e_k (no cancellation)
    const float denom = nrm2 - alpha * x0;                    // = v^T v / 2
    const float beta = (denom > 0.f) ? 1.f / denom : 0.f;     // H = I - beta v v^T
#pragma unroll
    for (int j = k + 1; j < N; ++j) {
      float d = v0 * G[k][j];
#pragma unroll
      for (int i = k + 1; i < N; ++i) d = fmaf(G[i][k], G[i][j], d);
      d *= beta;
      G[k][j] = fmaf(-d, v0, G[k][j]);
#pragma unroll
      for (int i = k + 1; i < N; ++i) G[i][j] = fmaf(-d, G[i][k], G[i][j]);
      cn[j] = fmaxf(cn[j] - G[k][j] * G[k][j], 0.f);
    }
    {
      float d = v0 * y[k];
#pragma unroll
      for (int i = k + 1; i < N; ++i) d = fmaf(G[i][k], y[i], d);
      d *= beta;
      y[k] = fmaf(-d, v0, y[k]);
#pragma unroll
      for (int i = k + 1; i < N; ++i) y[i] = fmaf(-d, G[i][k], y[i]);
    }
    G[k][k] = (denom > 0.f) ? alpha : x0;
#pragma unroll
    for (int i = k + 1; i < N; ++i) G[i][k] = 0.f;
  }
}

// Householder QR without pivoting, in place: [G | y] <- [R | Q^T y]; then, if R is provably far from
// the pinv cutoff, the plain solve x = R^-1 Q^T y.  "Provably": sigma_min(R) >= 1 / |R^-1|_F and
// sigma_max(R) <= |R|_F (Frobenius norms bound the spectral ones), so
//     1 / (|R^-1|_F |R|_F) > 4 rcond                                        (4: rounding of R itself)
// implies that no singular value of M is at or below tf.linalg.pinv's cutoff rcond * sigma_max
// (rmp.py:153): nothing is truncated and pinv(M) f = M^-1 f.  The test loses at most a factor N against
// the true sigma_min / sigma_max.  Trees with an isotropic metric leaf pass it for essentially every
// environment (configs 2, 3, 5: sigma_min/sigma_max ~ 0.2) and skip the Jacobi sweeps altogether.
// Rows/columns >= n (kernel width padding) are zero in M; they get a unit diagonal here so that R stays
// invertible (x_j = 0 there), which only makes the test more conservative, and the caller resets it.
// Returns true when x holds the solution.
template <int N>
RMP2_DEV bool qr_solve_if_well_conditioned(float (&G)[N][N], float (&y)[N], int n, float rcond, float (&x)[N]) {
#pragma unroll
  for (int j = 0; j < N; ++j)
    if (j >= n) G[j][j] = 1.f;
#pragma unroll
  for (int k = 0; k < N - 1; ++k) {
    float nrm2 = 0.f;
#pragma unroll
    for (int i = k; i < N; ++i) nrm2 = fmaf(G[i][k], G[i][k], nrm2);
    const float nrm = sqrtf(nrm2);
    const float x0 = G[k][k];
    const float alpha = (x0 > 0.f) ? -nrm : nrm;             // R_kk
    const float v0 = x0 - alpha;                              // v = x - alpha e_k (no cancellation)
    const float denom = nrm2 - alpha * x0;                    // = v^T v / 2
    const float beta = (denom > 0.f) ? 1.f / denom : 0.f;     // H = I - beta v v^T
#pragma unroll
    for (int j = k + 1; j < N; ++j) {
      float d = v0 * G[k][j];
#pragma unroll
      for (int i = k + 1; i < N; ++i) d = fmaf(G[i][k], G[i][j], d);
      d *= beta;
      G[k][j] = fmaf(-d, v0, G[k][j]);
#pragma unroll
      for (int i = k + 1; i < N; ++i) G[i][j] = fmaf(-d, G[i][k], G[i][j]);
    }
    {
      float d = v0 * y[k];
#pragma unroll
      for (int i = k + 1; i < N; ++i) d = fmaf(G[i][k], y[i], d);
      d *= beta;
      y[k] = fmaf(-d, v0, y[k]);
#pragma unroll
      for (int i = k + 1; i < N; ++i) y[i] = fmaf(-d, G[i][k], y[i]);
    }
    G[k][k] = (denom > 0.f) ? alpha : x0;
#pragma unroll
    for (int i = k + 1; i < N; ++i) G[i][k] = 0.f;
  }
  // W = R^-1 (upper triangular) row by row, its Frobenius norm, |R|_F, and x = W y on the way
  float r2 = 0.f, w2 = 0.f;
  bool finite = true;
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int j = i; j < N; ++j) r2 = fmaf(G[i][j], G[i][j], r2);
  float inv_diag[N];
#pragma unroll
  for (int j = 0; j < N; ++j) {
    finite = finite && (G[j][j] != 0.f);
    inv_diag[j] = 1.f / G[j][j];
  }
#pragma unroll
  for (int i = 0; i < N; ++i) {
    float W[N];
    W[i] = inv_diag[i];
    float acc = W[i] * y[i];
    w2 = fmaf(W[i], W[i], w2);
#pragma unroll
    for (int j = i + 1; j < N; ++j) {
      float sum = 0.f;
#pragma unroll
      for (int k = i; k < j; ++k) sum = fmaf(W[k], G[k][j], sum);
      W[j] = -sum * inv_diag[j];
      acc = fmaf(W[j], y[j], acc);
      w2 = fmaf(W[j], W[j], w2);
    }
    x[i] = acc;
  }
#pragma unroll
  for (int j = 0; j < N; ++j)
    if (j >= n) G[j][j] = 0.f;                   // back to the padded problem for the Jacobi fallback
  // 1 / (|W|_F |R|_F) > 4 rcond   <=>   16 rcond^2 r2 w2 < 1     (NaN / inf compare false)
  return finite && (16.f * rcond * rcond * r2 * w2 < 1.f);
}

// Rank-revealing direct solve for trees whose metric may be rank deficient (after qr_column_pivoting).
// With M P = Q [R11 R12; 0 R22] and r = the number of leading columns for which
//     sigma_min(R11) >= 1 / |R11^-1|_F > 2 rcond |R|_F >= 2 rcond sigma_max            (r-th singular value clear of the cutoff)
// the remaining singular values are bounded by |R22|_F (interlacing).  If additionally
//     |R22|_F <= rcond |R_11| / 2  <= rcond sigma_max / 2      (all of them clear below the cutoff)
//     |R22|_F |R11^-1|_F <= 1e-3                                (gap: second-order term <= 1e-6)
// then tf.linalg.pinv (rmp.py:153) keeps exactly r singular values and the truncated-SVD solution equals, up
// to (|R22| / sigma_r)^2, the minimum-norm solution of [R11 R12] z = (Q^T f)_1 (complete orthogonal
// decomposition), obtained here without a second factorisation:
//     b = R11^-1 y1,  C = R11^-1 R12,  z2 = (I + C^T C)^-1 C^T b,  z1 = b - C z2,  x = P z.
// (I + C^T C is SPD with eigenvalues >= 1: Cholesky without pivoting; column pivoting keeps |C_ij| = O(1).)
// r differs per lane, so every loop runs at full width N and r acts through selects: the instruction stream
// is uniform across the warp.  The factor 2 on either side of the cutoff is what float32 needs: the bounds are
// rigorous for the computed R, whose singular values differ from those of M by ~eps32 sigma_max, 70 times less than
// the cutoff itself.  On the config-4 tree (target + joint limits + obstacles: rank 5..7, a continuum of singular
// values down to zero) 98.5 % of the environments qualify; measured against the exact truncated-SVD solution of
// the same float32 matrix the deviation is <= 7.4e-7 (median 7e-15, 4096 envs).
// G, y, perm are left untouched: lanes that do not qualify continue with the Jacobi sweeps on them.
// Returns true when x holds the solution.
template <int N>
RMP2_DEV bool cod_solve_if_gap(const float (&G)[N][N], const float (&y)[N], const int (&perm)[N], float rcond,
                               float (&x)[N]) {
  float rn[N], cs[N], inv_diag[N];
  float nF2 = 0.f;
#pragma unroll
  for (int i = 0; i < N; ++i) {
    float a = 0.f;
#pragma unroll
    for (int j = i; j < N; ++j) a = fmaf(G[i][j], G[i][j], a);
    rn[i] = a;
    nF2 += a;
    cs[i] = 0.f;
    inv_diag[i] = 1.f / G[i][i];                  // inf for an exactly zero pivot: stops the rank count below
  }
  // W = R^-1 row by row (its leading r x r block is R11^-1 for every r), column sums of squares
  float W[N][N];
#pragma unroll
  for (int i = 0; i < N; ++i) {
    W[i][i] = inv_diag[i];
    cs[i] = fmaf(W[i][i], W[i][i], cs[i]);
#pragma unroll
    for (int j = i + 1; j < N; ++j) {
      float sum = 0.f;
#pragma unroll
      for (int k = i; k < j; ++k) sum = fmaf(W[i][k], G[k][j], sum);
      W[i][j] = -sum * inv_diag[j];
      cs[j] = fmaf(W[i][j], W[i][j], cs[j]);
    }
  }
  // r = leading columns with 1 / |R11^-1|_F > 2 rcond |R|_F   (NaN / inf compare false and end the count)
  const float thr = 4.f * rcond * rcond * nF2;
  int r = 0;
  bool alive = true;
  float F2 = 0.f, F2r = 0.f;
#pragma unroll
  for (int j = 0; j < N; ++j) {
    F2 += cs[j];
    alive = alive && (thr * F2 < 1.f);
    r += alive ? 1 : 0;
    F2r = alive ? F2 : F2r;
  }
  float tail2 = 0.f;
#pragma unroll
  for (int i = 0; i < N; ++i) tail2 += (i >= r) ? rn[i] : 0.f;
  const bool ok = (4.f * tail2 <= rcond * rcond * G[0][0] * G[0][0]) && (tail2 * F2r <= 1e-6f);
  // b = R11^-1 y1 and C = R11^-1 R12, masked by selects (entries of W beyond column r may be inf / NaN)
  float b[N];
#pragma unroll
  for (int i = 0; i < N; ++i) {
    float acc = 0.f;
#pragma unroll
    for (int j = i; j < N; ++j) {
      W[i][j] = (j < r) ? W[i][j] : 0.f;
      acc = fmaf(W[i][j], y[j], acc);
    }
    b[i] = acc;
  }
  float C[N][N];                                   // strictly upper part used: C[i][j], i < j
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int j = i + 1; j < N; ++j) {
      float sum = 0.f;
#pragma unroll
      for (int k = i; k < j; ++k) sum = fmaf(W[i][k], G[k][j], sum);
      C[i][j] = (j >= r) ? sum : 0.f;
    }
  // K = I + C^T C (lower triangle), t = C^T b; rows / columns < r are those of the identity
  float K[N][N], t[N];
#pragma unroll
  for (int l = 0; l < N; ++l) {
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < l; ++i) acc = fmaf(C[i][l], b[i], acc);
    t[l] = acc;
#pragma unroll
    for (int j = 0; j <= l; ++j) {
      float s = (j == l) ? 1.f : 0.f;
#pragma unroll
      for (int i = 0; i < j; ++i) s = fmaf(C[i][j], C[i][l], s);
      K[l][j] = s;
    }
  }
  // Cholesky K = L L^T in place (diagonal holds 1 / L_jj), L u = t, L^T z2 = u
#pragma unroll
  for (int j = 0; j < N; ++j) {
    float d = K[j][j];
#pragma unroll
    for (int k = 0; k < j; ++k) d = fmaf(-K[j][k], K[j][k], d);
    const float inv = rsqrtf(d);
    K[j][j] = inv;
#pragma unroll
    for (int i = j + 1; i < N; ++i) {
      float s = K[i][j];
#pragma unroll
      for (int k = 0; k < j; ++k) s = fmaf(-K[i][k], K[j][k], s);
      K[i][j] = s * inv;
    }
  }
#pragma unroll
  for (int i = 0; i < N; ++i) {
    float s = t[i];
#pragma unroll
    for (int k = 0; k < i; ++k) s = fmaf(-K[i][k], t[k], s);
    t[i] = s * K[i][i];
  }
#pragma unroll
  for (int i = N - 1; i >= 0; --i) {
    float s = t[i];
#pragma unroll
    for (int k = i + 1; k < N; ++k) s = fmaf(-K[k][i], t[k], s);
    t[i] = s * K[i][i];                           // t now holds z2 (zero for columns < r)
  }
  float z[N];
#pragma unroll
  for (int i = 0; i < N; ++i) {
    float s = b[i];
#pragma unroll
    for (int j = i + 1; j < N; ++j) s = fmaf(-C[i][j], t[j], s);
    z[i] = s + t[i];
  }
#pragma unroll
  for (int j = 0; j < N; ++j) {                    // x[perm[k]] = z[k]
    x[j] = 0.f;
#pragma unroll
    for (int k = 0; k < N; ++k) x[j] = (perm[k] == j) ? z[k] : x[j];
  }
  return ok;
}

// First half of the resolve: factorise and, where the matrix is provably clear of the pinv cutoff, solve.
//   kQr  (trees that may be rank deficient): Householder QR with column pivoting, then the rank-revealing
//        direct solve above;
//   !kQr (trees with an isotropic metric leaf): plain QR, then the full-rank direct solve.
// Leaves [R | Q^T y] (and perm) behind for resolve_jacobi.  Returns true when xs holds the solution.
template <int N, bool kQr>
RMP2_DEV bool resolve_direct(float (&G)[N][N], float (&y)[N], int (&perm)[N], int n, float rcond, float (&xs)[N]) {
  if (kQr) {
    qr_column_pivoting<N>(G, y, perm);
    return cod_solve_if_gap<N>(G, y, perm, rcond, xs);
  }
#pragma unroll
  for (int j = 0; j < N; ++j) perm[j] = j;
  return qr_solve_if_well_conditioned<N>(G, y, n, rcond, xs);
}

// Second half: truncated SVD by one-sided Jacobi on [R | Q^T y] for the lanes with solved == false (the others
// idle through it and keep xs).  A lane's result never depends on its neighbours: the warp votes only skip
// work nobody needs.
template <int N, bool kQr>
RMP2_DEV void resolve_jacobi(float (&G)[N][N], float (&y)[N], const int (&perm)[N], bool solved, const float (&xs)[N],
                             float rcond, float (&x)[N]) {
  using Sch = JacobiSchedule<N>;
  float nrm[N];
  for (int sweep = 0; sweep < RMP2_JACOBI_MAX_SWEEPS; ++sweep) {
    // squared row norms: exact at the start of every sweep, updated in closed form inside it
    float smax = 0.f;
#pragma unroll
    for (int i = 0; i < N; ++i) {
      float a0 = 0.f, a1 = 0.f;
#pragma unroll
      for (int j = 0; j + 1 < N; j += 2) {
        a0 = fmaf(G[i][j], G[i][j], a0);
        a1 = fmaf(G[i][j + 1], G[i][j + 1], a1);
      }
      if (N % 2) a0 = fmaf(G[i][N - 1], G[i][N - 1], a0);
      nrm[i] = a0 + a1;
      smax = fmaxf(smax, nrm[i]);
    }
    // A row whose norm is below drop2 ends below the pinv cutoff whatever happens next (the smaller
    // row of a pair only loses energy), so its own direction is irrelevant.  Such a row is still
    // rotated against a large row while the rotation angle is visible in y (> RMP2_JACOBI_ANGLE),
    // but no longer just to keep it orthogonal relative to its own, ever shrinking, norm -- which in
    // float32 never terminates for exactly rank-deficient M.
    const float drop2 = rcond * rcond * smax * RMP2_JACOBI_DROP;
    bool rotated = false;
#pragma unroll
    for (int r = 0; r < Sch::rounds; ++r) {
#pragma unroll
      for (int k = 0; k < Sch::slots; ++k) {
        const int p0 = Sch::first(r, k), q0 = Sch::second(r, k);
        if (p0 >= N || q0 >= N) continue;            // compile-time after unrolling
        const int p = p0 < q0 ? p0 : q0, q = p0 < q0 ? q0 : p0;
        float g0 = 0.f, g1 = 0.f;
#pragma unroll
        for (int j = 0; j + 1 < N; j += 2) {
          g0 = fmaf(G[p][j], G[q][j], g0);
          g1 = fmaf(G[p][j + 1], G[q][j + 1], g1);
        }
        if (N % 2) g0 = fmaf(G[p][N - 1], G[q][N - 1], g0);
        const float g = g0 + g1;
        const float a = nrm[p], b = nrm[q];
        const float g2 = g * g;
        const float mx = fmaxf(a, b), mn = fminf(a, b);
        const bool rot = !solved && (g2 > (RMP2_JACOBI_TOL * RMP2_JACOBI_TOL) * a * b) &&
                         (mn >= drop2 || g2 > (RMP2_JACOBI_ANGLE * RMP2_JACOBI_ANGLE) * mx * mx);
        if (!__any_sync(0xffffffffu, rot)) continue;  // warp-uniform: nobody needs this pair
        rotated |= rot;
        // tan of the rotation angle, smaller root of t^2 + 2 zeta t - 1 = 0 with zeta = (b-a)/(2g):
        //   t = 2 g sign(b-a) / (|b-a| + sqrt((b-a)^2 + 4 g^2)).
        // Lanes that do not rotate use g = 0, which yields t = 0, s = 0 and c ~ 1 by itself.  Any (c, s)
        // gives an orthogonal-up-to-scale row transform and the solution below is invariant to row
        // scaling, so the approximate MUFU results only affect the convergence rate.
        const float ge = rot ? g : 0.f;
        const float tau = b - a;
        const float hyp2 = fmaf(tau, tau, 4.f * ge * ge);
        const float hyp = hyp2 * fast_rsqrt(fmaxf(hyp2, 1e-37f));
        const float t = ((tau < 0.f) ? -2.f * ge : 2.f * ge) * fast_rcp(fmaxf(fabsf(tau) + hyp, 1e-37f));
        const float c = fast_rsqrt(fmaf(t, t, 1.f));
        const float s = c * t;
        const float tg = t * ge;                      // sign(t g) = sign(b - a): the larger row gains
        nrm[p] = a - tg;
        nrm[q] = b + tg;
#pragma unroll
        for (int j = 0; j < N; ++j) {
          const float gp = G[p][j], gq = G[q][j];
          G[p][j] = fmaf(c, gp, -s * gq);
          G[q][j] = fmaf(s, gp, c * gq);
        }
        const float yp = y[p], yq = y[q];
        y[p] = fmaf(c, yp, -s * yq);
        y[q] = fmaf(s, yp, c * yq);
      }
    }
    if (!__any_sync(0xffffffffu, rotated)) break;
  }
  float sig2[N];
  float smax = 0.f;
#pragma unroll
  for (int i = 0; i < N; ++i) {
    float a = 0.f;
#pragma unroll
    for (int j = 0; j < N; ++j) a = fmaf(G[i][j], G[i][j], a);
    sig2[i] = a;
    smax = fmaxf(smax, a);
  }
  const float cut2 = rcond * rcond * smax;
  float xp[N];                                   // solution (in pivoted column order with QRCP)
#pragma unroll
  for (int j = 0; j < N; ++j) xp[j] = 0.f;
#pragma unroll
  for (int i = 0; i < N; ++i) {
    const float coef = (sig2[i] > cut2) ? y[i] / sig2[i] : 0.f;
#pragma unroll
    for (int j = 0; j < N; ++j) xp[j] = fmaf(G[i][j], coef, xp[j]);
  }
  if (kQr) {
#pragma unroll
    for (int j = 0; j < N; ++j) {                // x[perm[k]] = xp[k]
      float v = 0.f;
#pragma unroll
      for (int k = 0; k < N; ++k) v = (perm[k] == j) ? xp[k] : v;
      x[j] = solved ? xs[j] : v;
    }
  } else {
#pragma unroll
    for (int j = 0; j < N; ++j) x[j] = solved ? xs[j] : xp[j];
  }
}

// Whole resolve inside one kernel: direct solve, Jacobi only in warps where some lane did not qualify.
// kDirect = false skips the direct solves (every lane takes the Jacobi path; used to cross-check the solvers).
template <int N, bool kQr, bool kDirect = true>
RMP2_DEV void resolve_pinv(float (&G)[N][N], float (&y)[N], int n, float rcond, float (&x)[N]) {
  int perm[N];
  float xs[N];
  bool solved = resolve_direct<N, kQr>(G, y, perm, n, rcond, xs);
  if (!kDirect) solved = false;
  if (__all_sync(0xffffffffu, solved)) {
#pragma unroll
    for (int j = 0; j < N; ++j) x[j] = xs[j];
    return;
  }
  resolve_jacobi<N, kQr>(G, y, perm, solved, xs, rcond, x);
}
