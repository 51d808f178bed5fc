// C ABI of rmp2_b200 (declared in include/rmp2_b200.h): handle lifetime, tree compilation
// (host only), argument checking, kernel launches, the host-buffer pipeline.
#include <cuda.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <functional>
#include <string>
#include <vector>

#include "rmp2_launch.h"
#include "rmp2_jit.h"
#include "rmp2_leaves.cuh"

namespace {

thread_local std::string g_last_error;
std::atomic<long long> g_launches{0};

int fail(int code, const std::string& msg) {
  g_last_error = msg;
  return code;
}

int cuda_fail(cudaError_t e, const char* what) {
  return fail(RMP2_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}

}  // namespace

struct rmp2_robot {
  int F = 0, n = 0;
  std::vector<float> T_const;   // [F][16]
  std::vector<float> axis;      // [F][3]
  std::vector<int8_t> jtype;
  std::vector<int32_t> parent, qidx;
};

struct HostStage {                // device staging for rmp2_step_host, one per pipeline slot
  float* buf = nullptr;
  size_t floats = 0;
  cudaStream_t stream = nullptr;
};

#define RMP2_N_CLOCKS 5

struct KernelClock {              // optional per-kernel timing (rmp2_tree_profile)
  std::vector<cudaEvent_t> pending;   // (start, stop) pairs not yet read
  double ms = 0.0;
  long long launches = 0;
};

struct rmp2_tree {
  StepTables tab;
  rmp2_robot robot;                       // copy of the kinematics the tables were compiled from (recompiled on demand)
  std::vector<rmp2_leaf_desc> leaves;     // as given, tree order
  std::vector<int> table_index;           // tree order -> index in tab.leaves (a merged leaf: its representative's)
  bool merge_coincident = true;           // RMP2_OPT_MERGE_COINCIDENT
  int n_merged = 0;                       // sphere-obstacle leaves represented by another leaf's record slot
  int n_pair_sets = 0;
  int n_goal_slots_used = 0;
  SphereTables sph;                       // parameters of the sphere-obstacle leaves, by record slot
  float* rec = nullptr;                   // sphere-record scratch of rmp2_step (field-major, see rmp2_tables.h)
  size_t rec_floats = 0;
  float* mf = nullptr;                    // (M, f) scratch between step and resolve kernel / problems handed to the fallback
  size_t mf_floats = 0;
  int32_t* fb = nullptr;                  // fallback work list (count, ticket, environment indices)
  size_t fb_ints = 0;
  long long reserved_envs = 0;            // rmp2_tree_reserve
  int reserved_spheres = 0;
  bool profiling = false;
  bool early_out = true;                  // RMP2_OPT_EARLY_OUT
  bool use_tma = true;                    // RMP2_OPT_TMA
  int split_resolve = -1;                 // RMP2_OPT_SPLIT_RESOLVE: -1 by batch size, 0 fused, 1 split
  int force_block = 0;                    // RMP2_OPT_BLOCK_THREADS: 0 by batch size, else 32 / 64 / 128
  long long chunk_envs = 0;               // RMP2_OPT_CHUNK_ENVS: environments per internal chunk (0 = 2^20)
  int spec_step_block = 0;                // tuning hook RMP2_SPEC_STEP_BLOCK: block size of the specialised step kernel for large batches (0: 256 for n <= 7)
  bool respecialize = false;              // rebuild the specialised kernels when a leaf changes
  KernelClock clock[RMP2_N_CLOCKS];       // frames, spheres, step, resolve, resolve fallback
  HostStage stage[3];
  SpecModule* spec = nullptr;             // tree-specialised frames / step kernels (rmp2_tree_specialize)
};

#define RMP2_STEP_CHUNK (1LL << 20)       // environments per internal chunk (bounds the scratch)

// Sphere-row parameters of an ObstacleAvoidance leaf for the packed pair loop (SP_* in rmp2_leaves.cuh).
// The kernel works with xs = clamp((x - margin)/r, 0, 1): r is folded into every coefficient that
// multiplies x, metric_scalar into the first denominator.  A leaf whose metric vanishes identically
// (metric_scalar == 0, or a radius <= 0: rmp2.py:187,194) gets a zero gate.
static void fill_sphere_row(const rmp2_leaf_desc& d, float* row) {
  const float* r = d.params;
  for (int i = 0; i < RMP2_LEAF_PARAMS; ++i) row[i] = 0.f;
  const double margin = r[0], rad = r[7], ms = r[8];
  const bool live = (ms != 0.0) && (rad > 0.0);
  const double fold = live ? ms : 1.0, R = live ? rad : 1.0;
  row[SP_XA] = (float)(1.0 / R);
  row[SP_XB] = (float)(-margin / R);
  row[SP_GT_A] = live ? 1.f : 0.f;
  row[SP_GT_B] = live ? -1.f : 0.f;
  row[SP_D1A] = (float)(R / ((double)r[9] * fold));
  row[SP_D1B] = (float)((double)r[10] / fold);
  // the damping gain D and the velocity scale k divide den2: q = 1/(den1' den2/(D k) (1 + e_v)) then carries them
  // (w2 = q den1' = D k (1-sig)/den2, w1 = q den2/(D k) unchanged).  D == 0: a constant den2' = 2^60 makes the
  // damping term vanish below float32 resolution.  The pair loop works with v' = k v (see obstacle_pair2).
  const double D = r[1], kv = 1.4426950408889634 / (double)r[4];
  const bool damped = std::fabs(D) > 1e-30;
  row[SP_D2A] = damped ? (float)(R / ((double)r[2] * D * kv)) : 0.f;
  row[SP_D2B] = damped ? (float)((double)r[3] / (D * kv)) : 1152921504606846976.f;
  row[SP_K_VEL] = (float)kv;
  row[SP_K_REP] = (float)(-1.4426950408889634 * R / (double)r[6]);
  row[SP_RGAIN] = (float)((double)r[5] * kv * kv);
  row[SP_INV_K2] = (float)(1.0 / (kv * kv));
  row[SP_REACH] = (float)((rad + margin) * 1.00001);
}

// ------------------------------------------------------------------------------ parameter derivation
// Raw constructor arguments (layout in include/rmp2_b200.h) -> kernel parameters (rmp2_leaves.cuh).
// Python evaluates these expressions in double before they meet a float32 tensor, so they are
// evaluated in double here and rounded once.
static int derive_leaf_params(const rmp2_leaf_desc& d, int n, LeafTab& L, float* vec, int* vec_len) {
  const float* r = d.params;
  for (int i = 0; i < RMP2_LEAF_PARAMS; ++i) L.p[i] = 0.f;
  *vec_len = 0;
  const int dim = (d.space == RMP2_SPACE_CONFIG) ? n : 3;
  switch (d.type) {
    case RMP2_LEAF_TARGET_POLICY:
      L.p[TP_ALPHA] = r[0];
      L.p[TP_BETA] = r[1];
      L.p[TP_C] = r[2];
      L.p[TP_INV_C] = (float)(1.0 / (double)r[2]);
      for (int i = 0; i < dim; ++i) vec[i] = d.vec[i];
      *vec_len = dim;
      break;
    case RMP2_LEAF_CONFIG_BIASING:
      L.p[CB_GAMMA_P] = r[0];
      L.p[CB_GAMMA_D] = r[1];
      L.p[CB_W] = r[2];
      for (int i = 0; i < n; ++i) vec[i] = d.vec[i];
      *vec_len = n;
      break;
    case RMP2_LEAF_JOINT_LIMIT: {
      const double rr = 0.15;                                   // rmp.py:364
      L.p[JL_GAMMA_P] = r[0];
      L.p[JL_GAMMA_D] = r[1];
      L.p[JL_C3] = (float)(2.0 / (rr * rr * rr));
      L.p[JL_C2] = (float)(-3.0 / (rr * rr));
      L.p[JL_R] = (float)rr;
      L.p[JL_INV_QDMAX] = (float)(1.0 / (20.0 * (2.0 * M_PI) / 60.0));   // rmp.py:374
      L.p[JL_BETA] = 0.9f;                                               // rmp.py:376
      L.p[JL_C] = 5.f;
      L.p[JL_INV_C] = 0.2f;
      for (int i = 0; i < n; ++i) {
        vec[i] = d.vec[i];
        vec[n + i] = d.vec[n + i];
        vec[2 * n + i] = 1.f / (d.vec[n + i] - d.vec[i]);
      }
      *vec_len = 3 * n;
      break;
    }
    case RMP2_LEAF_TARGET_ATTRACTOR:
      L.p[TA_PGAIN] = r[0];
      L.p[TA_DGAIN] = r[1];
      L.p[TA_EPS] = r[2];
      L.p[TA_EPS10] = (float)((double)r[2] / 10.0);
      L.p[TA_INV_ALEN] = (float)(1.0 / (double)r[3]);
      L.p[TA_MIN_ALPHA] = r[4];
      L.p[TA_SMAX] = r[5];
      L.p[TA_SMIN] = r[6];
      L.p[TA_BOOST] = r[7];
      L.p[TA_INV_BLEN] = (float)(1.0 / (double)r[8]);
      for (int i = 0; i < 3; ++i) vec[i] = d.vec[i];
      *vec_len = 3;
      break;
    case RMP2_LEAF_VELOCITY_CAP:
      L.p[VC_CUTOFF] = (float)((double)r[0] - (double)r[1]);   // rmp2.py:97
      L.p[VC_GAIN] = r[2];
      L.p[VC_CLIP] = (float)((double)r[1] - 1e-6);             // rmp2.py:104
      L.p[VC_INV_REGION] = (float)(1.0 / (double)r[1]);
      L.p[VC_WEIGHT] = r[3];
      break;
    case RMP2_LEAF_JOINT_DAMPING:
      L.p[JD_GAIN] = r[0];
      L.p[JD_SCALAR] = r[1];
      L.p[JD_INERTIA] = r[2];
      break;
    case RMP2_LEAF_OBSTACLE_AVOIDANCE:
      L.p[OA_MARGIN] = r[0];
      L.p[OA_DGAIN] = r[1];
      L.p[OA_INV_DSTD] = (float)(1.0 / (double)r[2]);
      L.p[OA_DEPS] = r[3];
      L.p[OA_INV_VLEN] = (float)(1.0 / (double)r[4]);
      L.p[OA_RGAIN] = r[5];
      L.p[OA_INV_RSTD] = (float)(1.0 / (double)r[6]);
      L.p[OA_R] = r[7];
      L.p[OA_INV_R] = (float)(1.0 / (double)r[7]);
      L.p[OA_MSCALAR] = r[8];
      L.p[OA_INV_ESTD] = (float)(1.0 / (double)r[9]);
      L.p[OA_EEPS] = r[10];
      L.p[OA_K_REP] = (float)(-1.4426950408889634 / (double)r[6]);
      L.p[OA_K_VEL] = (float)(1.4426950408889634 / (double)r[4]);
      break;
    case RMP2_LEAF_COLLISION_AVOIDANCE: {
      const double rr = (double)r[4];
      L.p[CA_ETA_REP] = r[0];
      L.p[CA_INV_NU_REP] = (float)(1.0 / (double)r[1]);
      L.p[CA_ETA_DAMP] = r[2];
      L.p[CA_INV_NU_DAMP] = (float)(1.0 / (double)r[3]);
      L.p[CA_R] = r[4];
      L.p[CA_C3] = (float)(2.0 / (rr * rr * rr));               // rmp.py:303-304
      L.p[CA_C2] = (float)(-3.0 / (rr * rr));
      break;
    }
    case RMP2_LEAF_CSPACE_BIASING:
      L.p[CS_METRIC] = (float)((double)r[0] + (double)r[4]);   // rmp2.py:224
      L.p[CS_PGAIN] = r[1];
      L.p[CS_DGAIN] = r[2];
      L.p[CS_THRESH] = r[3];
      for (int i = 0; i < n; ++i) vec[i] = d.vec[i];
      *vec_len = n;
      break;
    default:
      return fail(RMP2_ERR_INVALID, "unknown leaf type " + std::to_string(d.type));
  }
  return RMP2_OK;
}

static int check_leaf(const rmp2_leaf_desc& d, int F, int idx) {
  const std::string at = "leaf " + std::to_string(idx) + ": ";
  const bool config_leaf = d.type == RMP2_LEAF_CONFIG_BIASING || d.type == RMP2_LEAF_JOINT_LIMIT ||
                           d.type == RMP2_LEAF_VELOCITY_CAP || d.type == RMP2_LEAF_JOINT_DAMPING ||
                           d.type == RMP2_LEAF_CSPACE_BIASING;
  switch (d.space) {
    case RMP2_SPACE_CONFIG:
      if (!config_leaf && d.type != RMP2_LEAF_TARGET_POLICY)
        return fail(RMP2_ERR_UNSUPPORTED, at + "this leaf type is not implemented on the identity task map");
      if (d.goal_slot >= 0) return fail(RMP2_ERR_UNSUPPORTED, at + "per-environment goals need a frame position / orientation task map");
      break;
    case RMP2_SPACE_FRAME_POSITION:
    case RMP2_SPACE_FRAME_EULER:
      if (d.type != RMP2_LEAF_TARGET_POLICY && d.type != RMP2_LEAF_TARGET_ATTRACTOR)
        return fail(RMP2_ERR_UNSUPPORTED, at + "only TargetPolicy / TargetAttractor live on a frame position / orientation");
      break;
    case RMP2_SPACE_FRAME_DISTANCE_SPHERES:
    case RMP2_SPACE_FRAME_DISTANCE_PAIRS:
      if (d.type != RMP2_LEAF_OBSTACLE_AVOIDANCE)
        return fail(RMP2_ERR_UNSUPPORTED, at + "only ObstacleAvoidance lives on a distance task map");
      break;
    case RMP2_SPACE_FRAME_POINTS:
      if (d.type != RMP2_LEAF_COLLISION_AVOIDANCE)
        return fail(RMP2_ERR_UNSUPPORTED, at + "only CollisionAvoidance lives on frame-fixed points");
      break;
    default:
      return fail(RMP2_ERR_INVALID, at + "unknown space " + std::to_string(d.space));
  }
  if (d.space != RMP2_SPACE_CONFIG && (d.frame < 0 || d.frame >= F))
    return fail(RMP2_ERR_INVALID, at + "frame index out of range");
  if (d.goal_slot >= RMP2_MAX_GOAL_SLOTS) return fail(RMP2_ERR_INVALID, at + "goal_slot out of range");
  return RMP2_OK;
}

// Path base -> frame as a serial table (used by rmp2_fk and as a building block below).
static void fill_frame(const rmp2_robot& rb, int k, FrameTab& ft) {
  const float* T = &rb.T_const[(size_t)k * 16];
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) ft.R[3 * r + c] = T[4 * r + c];
    ft.t[r] = T[4 * r + 3];
    ft.axis[r] = rb.axis[(size_t)k * 3 + r];
  }
  ft.type = rb.jtype[k];
  ft.qidx = (rb.qidx[k] >= 0 && rb.qidx[k] < rb.n && ft.type != RMP2_JOINT_FIXED) ? rb.qidx[k] : -1;
  ft.restore_slot = -1;
  ft.save_slot = -1;
  ft.anc_mask = 0;
  ft.leaf_begin = ft.leaf_end = 0;
  ft.ref_index = k;
  bool ident = true;
  for (int i = 0; i < 9; ++i) ident = ident && (ft.R[i] == ((i % 4 == 0) ? 1.f : 0.f));
  ft.const_rot_identity = ident ? 1 : 0;
  ft.axis_is_z = (ft.axis[0] == 0.f && ft.axis[1] == 0.f && ft.axis[2] == 1.f) ? 1 : 0;
}

static uint32_t ancestor_mask(const rmp2_robot& rb, int k) {
  uint32_t m = 0;
  for (int i = k; i >= 0; i = rb.parent[i])
    if (rb.qidx[i] >= 0 && rb.qidx[i] < rb.n && rb.jtype[i] != RMP2_JOINT_FIXED) m |= 1u << rb.qidx[i];
  return m;
}

static uint32_t prismatic_mask(const rmp2_robot& rb) {
  uint32_t m = 0;
  for (int i = 0; i < rb.F; ++i)
    if (rb.jtype[i] == RMP2_JOINT_PRISMATIC && rb.qidx[i] >= 0 && rb.qidx[i] < rb.n) m |= 1u << rb.qidx[i];
  return m;
}

// Kernel tables of a tree from its robot and leaf descriptors (rmp2_tree_create; again whenever a leaf or the
// merge option changes).  The tree's tables are only replaced when the whole compilation succeeds.
static int compile_tables_into(rmp2_tree* tr);
static int compile_tables(rmp2_tree* tr) {
  const StepTables tab = tr->tab;
  const SphereTables sph = tr->sph;
  const std::vector<int> table_index = tr->table_index;
  const int n_pair_sets = tr->n_pair_sets, n_goal_slots_used = tr->n_goal_slots_used, n_merged = tr->n_merged;
  const int rc = compile_tables_into(tr);
  if (rc != RMP2_OK) {
    tr->tab = tab, tr->sph = sph, tr->table_index = table_index;
    tr->n_pair_sets = n_pair_sets, tr->n_goal_slots_used = n_goal_slots_used, tr->n_merged = n_merged;
  }
  return rc;
}

static int compile_tables_into(rmp2_tree* tr) {
  const rmp2_robot& rb = tr->robot;
  const rmp2_leaf_desc* leaves = tr->leaves.data();
  const int n_leaves = (int)tr->leaves.size();
  memset(&tr->tab, 0, sizeof(StepTables));
  StepTables& T = tr->tab;
  T.n = rb.n;
  T.rcond = (float)(10.0 * rb.n * 1.1920928955078125e-07);   // 10 * max(rows, cols) * eps(float32)
  T.prismatic_mask = prismatic_mask(rb);
  tr->table_index.assign(n_leaves, -1);

  // Sphere-obstacle leaves that share their control point.  The distance map differentiates through the frame origin
  // only (taskmap.py:124-128) and the sphere path takes the frame origin as pos_on_link, so two ObstacleAvoidance
  // leaves with equal parameters whose frame origins coincide for every q -- frame k has a zero constant translation
  // and a revolute or fixed joint, so its origin is its parent's (Panda: joint2 on joint1, joint6 on joint5) -- yield
  // the same (M, f): same x, xd, c, and Jacobians that differ by a zero column.  One leaf of such a group (the one on
  // the shallowest frame) runs the pair loop and the pullback; its sums carry the size of the group as a weight.
  std::vector<int> rep(n_leaves), weight(n_leaves, 1);     // representative of leaf i (i itself when not merged)
  for (int i = 0; i < n_leaves; ++i) rep[i] = i;
  tr->n_merged = 0;
  if (tr->merge_coincident) {
    auto origin_root = [&](int k) {
      while (rb.parent[k] >= 0 && rb.jtype[k] != RMP2_JOINT_PRISMATIC && rb.T_const[16 * k + 3] == 0.f &&
             rb.T_const[16 * k + 7] == 0.f && rb.T_const[16 * k + 11] == 0.f)
        k = rb.parent[k];
      return k;
    };
    auto same_policy = [&](const rmp2_leaf_desc& a, const rmp2_leaf_desc& b) {
      return a.type == b.type && memcmp(a.params, b.params, sizeof(a.params)) == 0;
    };
    for (int i = 0; i < n_leaves; ++i) {
      if (leaves[i].space != RMP2_SPACE_FRAME_DISTANCE_SPHERES) continue;
      for (int j = 0; j < i; ++j) {
        if (leaves[j].space != RMP2_SPACE_FRAME_DISTANCE_SPHERES || rep[j] != j) continue;
        if (origin_root(leaves[j].frame) != origin_root(leaves[i].frame) || !same_policy(leaves[i], leaves[j])) continue;
        rep[i] = j;
        break;
      }
    }
    // the member on the shallowest frame (smallest index: parents precede children) represents its group
    std::vector<int> group = rep;                            // first member in tree order
    for (int j = 0; j < n_leaves; ++j) {
      if (group[j] != j) continue;
      int best = j, members = 0;
      for (int i = j; i < n_leaves; ++i)
        if (group[i] == j) {
          ++members;
          if (leaves[i].frame < leaves[best].frame) best = i;
        }
      if (members == 1) continue;
      for (int i = j; i < n_leaves; ++i)
        if (group[i] == j) rep[i] = best, weight[i] = 0;
      weight[best] = members;
      tr->n_merged += members - 1;
    }
    // A control point that cannot move -- its origin frame hangs on the base through fixed joints only and is not
    // prismatic itself (Panda: joint1, and joint2 on top of it) -- has J = 0 identically: the leaf's pulled-back
    // (M, f) is exactly zero in the reference as well (rmp.py:165-167 with J = 0).  Such a group runs no pair loop.
    for (int i = 0; i < n_leaves; ++i) {
      if (leaves[i].space != RMP2_SPACE_FRAME_DISTANCE_SPHERES || rep[i] != i) continue;
      const int r = origin_root(leaves[i].frame);
      bool immobile = rb.jtype[r] != RMP2_JOINT_PRISMATIC;
      for (int a = rb.parent[r]; a >= 0 && immobile; a = rb.parent[a]) immobile = rb.jtype[a] == RMP2_JOINT_FIXED;
      if (!immobile) continue;
      tr->n_merged += 1;                              // (its merged companions are counted above)
      for (int j = 0; j < n_leaves; ++j)
        if (rep[j] == i) rep[j] = -2, weight[j] = 0;  // no representative: dropped from the tables (rep[i] itself included)
    }
  }

  // frames that carry a leaf, and everything between them and the base
  std::vector<char> needed(rb.F, 0);
  for (int i = 0; i < n_leaves; ++i)
    if (leaves[i].space != RMP2_SPACE_CONFIG && rep[i] == i)
      for (int k = leaves[i].frame; k >= 0; k = rb.parent[k]) needed[k] = 1;
  std::vector<std::vector<int>> children(rb.F + 1);          // index F = base
  for (int k = 0; k < rb.F; ++k)
    if (needed[k]) children[rb.parent[k] < 0 ? rb.F : rb.parent[k]].push_back(k);

  // leaf bookkeeping shared by frame-attached and configuration-space leaves
  int leaf_cursor = 0, pair_sets = 0, goal_slots = 0, vec_cursor = 0;
  std::vector<int> pair_set_of(n_leaves, -1);
  for (int i = 0; i < n_leaves; ++i) {
    if (leaves[i].space == RMP2_SPACE_FRAME_DISTANCE_PAIRS || leaves[i].space == RMP2_SPACE_FRAME_POINTS)
      pair_set_of[i] = pair_sets++;
    if (leaves[i].goal_slot >= 0) goal_slots = std::max(goal_slots, leaves[i].goal_slot + 1);
    if (leaves[i].space == RMP2_SPACE_FRAME_DISTANCE_SPHERES) T.uses_spheres = 1;
  }
  memset(&tr->sph, 0, sizeof(SphereTables));
  if (pair_sets > RMP2_MAX_PAIR_SETS) { return fail(RMP2_ERR_UNSUPPORTED, "too many explicit-pair leaves"); }
  tr->n_pair_sets = pair_sets;
  tr->n_goal_slots_used = goal_slots;
  auto add_leaf = [&](int i) -> int {
    LeafTab& L = T.leaves[leaf_cursor];
    L.type = leaves[i].type;
    L.space = leaves[i].space;
    L.goal_slot = leaves[i].goal_slot;
    L.pair_set = pair_set_of[i];
    L.sphere_slot = -1;
    float vec[3 * RMP2_MAX_JOINTS];
    int vlen = 0;
    int rc = derive_leaf_params(leaves[i], rb.n, L, vec, &vlen);
    if (rc != RMP2_OK) return rc;
    if (L.space == RMP2_SPACE_FRAME_DISTANCE_SPHERES) {
      L.sphere_slot = T.n_sphere_slots++;
      fill_sphere_row(leaves[i], tr->sph.p[L.sphere_slot]);
      tr->sph.p[L.sphere_slot][SP_WEIGHT] = (float)weight[i];
      L.p[OA_G_SCALE] = tr->sph.p[L.sphere_slot][SP_INV_K2] * (float)weight[i];
    }
    if (vec_cursor + vlen > RMP2_VECPOOL) return fail(RMP2_ERR_UNSUPPORTED, "vector-parameter pool exhausted");
    L.vec_off = vec_cursor;
    for (int j = 0; j < vlen; ++j) T.vecpool[vec_cursor + j] = vec[j];
    vec_cursor += vlen;
    tr->table_index[i] = leaf_cursor++;
    return RMP2_OK;
  };

  // depth-first execution order.  A frame with several needed children saves its chain state in
  // slot `depth` (stack discipline: its descendants only use deeper slots); the first child
  // continues from the live state, later children reload the slot.
  int n_slots = 0;
  std::function<int(int, int, int)> visit = [&](int k, int restore, int depth) -> int {
    if (T.n_frames >= RMP2_MAX_FRAMES) return fail(RMP2_ERR_UNSUPPORTED, "too many frames");
    FrameTab& ft = T.frames[T.n_frames++];
    fill_frame(rb, k, ft);
    ft.anc_mask = ancestor_mask(rb, k);
    ft.restore_slot = restore;
    ft.leaf_begin = leaf_cursor;
    for (int i = 0; i < n_leaves; ++i)
      if (leaves[i].space != RMP2_SPACE_CONFIG && leaves[i].frame == k && rep[i] == i) {
        int rc = add_leaf(i);
        if (rc != RMP2_OK) return rc;
      }
    ft.leaf_end = leaf_cursor;
    const std::vector<int>& ch = children[k];
    if (ch.size() > 1) {
      if (depth >= RMP2_MAX_SLOTS) return fail(RMP2_ERR_UNSUPPORTED, "kinematic tree branches too deeply");
      ft.save_slot = depth;
      n_slots = std::max(n_slots, depth + 1);
      for (size_t ci = 0; ci < ch.size(); ++ci) {
        int rc = visit(ch[ci], ci == 0 ? -1 : depth, depth + 1);
        if (rc != RMP2_OK) return rc;
      }
    } else if (ch.size() == 1) {
      return visit(ch[0], -1, depth);
    }
    return RMP2_OK;
  };
  for (int k : children[rb.F]) {
    int rc = visit(k, RMP2_SLOT_BASE, 0);
    if (rc != RMP2_OK) return rc;
  }
  T.n_slots = n_slots;
  T.n_frame_leaves = leaf_cursor;
  for (int i = 0; i < n_leaves; ++i)
    if (leaves[i].space == RMP2_SPACE_CONFIG) {
      int rc = add_leaf(i);
      if (rc != RMP2_OK) return rc;
    }
  T.n_leaves = leaf_cursor;
  for (int i = 0; i < n_leaves; ++i)
    if (rep[i] >= 0 && rep[i] != i) tr->table_index[i] = tr->table_index[rep[i]];
  // leaves whose metric is a positive multiple of the identity keep M well conditioned
  T.precondition = 1;
  for (int i = 0; i < n_leaves; ++i) {
    const int t = leaves[i].type;
    if (t == RMP2_LEAF_CONFIG_BIASING || t == RMP2_LEAF_JOINT_DAMPING || t == RMP2_LEAF_CSPACE_BIASING) T.precondition = 0;
  }
  tr->sph.n_slots = T.n_sphere_slots;
  if (T.n_sphere_slots > 0) {
    // E environments per block: E * L threads <= 128, E <= 32 (box rows), shared memory bounded
    int E = RMP2_SPHERES_BLOCK / T.n_sphere_slots;
    E = std::max(1, std::min(E, 32));
    tr->sph.envs_per_block = E;
    tr->sph.div_magic = 65536 / E + 1;
    tr->sph.div_magic_slots = 65536 / T.n_sphere_slots + 1;
    // the slot-fastest order pays when the environment-fastest one misaligns its quarter-warps with the slots
    // (measured: E = 21 gains 7 % on the early-out pair kernel, E = 16 loses 0.5 %)
    tr->sph.slot_fastest = (E % 8 != 0) ? 1 : 0;
  }
  return RMP2_OK;
}

extern "C" {

const char* rmp2_last_error(void) { return g_last_error.c_str(); }
const char* rmp2_version(void) { return "rmp2_b200 0.1 (sm_100a)"; }
int64_t rmp2_launch_count(void) { return g_launches.load(); }

int rmp2_robot_create(const float* T_const, const float* axis, const int8_t* jtype, const int32_t* parent,
                      const int32_t* qidx, int32_t F, int32_t n, rmp2_robot** out) {
  if (!T_const || !axis || !jtype || !parent || !qidx || !out) return fail(RMP2_ERR_INVALID, "null argument");
  if (F <= 0 || F > RMP2_MAX_FRAMES) return fail(RMP2_ERR_UNSUPPORTED, "frame count must be in 1.." + std::to_string(RMP2_MAX_FRAMES));
  if (n <= 0 || n > RMP2_MAX_JOINTS) return fail(RMP2_ERR_UNSUPPORTED, "joint count must be in 1.." + std::to_string(RMP2_MAX_JOINTS));
  for (int k = 0; k < F; ++k) {
    if (parent[k] >= k || parent[k] < -1) return fail(RMP2_ERR_INVALID, "parent[" + std::to_string(k) + "] must be -1 or an earlier frame");
    if (jtype[k] < 0 || jtype[k] > 2) return fail(RMP2_ERR_INVALID, "jtype[" + std::to_string(k) + "] unknown");
    if (jtype[k] == RMP2_JOINT_REVOLUTE) {
      const float* a = axis + 3 * k;
      const double len = sqrt((double)a[0] * a[0] + (double)a[1] * a[1] + (double)a[2] * a[2]);
      if (fabs(len - 1.0) > 1e-4)
        return fail(RMP2_ERR_UNSUPPORTED, "revolute axis of frame " + std::to_string(k) + " is not unit length");
    }
    if (qidx[k] >= n) return fail(RMP2_ERR_INVALID, "qidx[" + std::to_string(k) + "] out of range");
  }
  rmp2_robot* rb = new rmp2_robot;
  rb->F = F;
  rb->n = n;
  rb->T_const.assign(T_const, T_const + (size_t)F * 16);
  rb->axis.assign(axis, axis + (size_t)F * 3);
  rb->jtype.assign(jtype, jtype + F);
  rb->parent.assign(parent, parent + F);
  rb->qidx.assign(qidx, qidx + F);
  *out = rb;
  return RMP2_OK;
}

void rmp2_robot_destroy(rmp2_robot* robot) { delete robot; }

int rmp2_tree_create(const rmp2_robot* rb, const rmp2_leaf_desc* leaves, int32_t n_leaves, rmp2_tree** out) {
  if (!rb || !out || (n_leaves > 0 && !leaves)) return fail(RMP2_ERR_INVALID, "null argument");
  if (n_leaves < 0 || n_leaves > RMP2_MAX_LEAVES)
    return fail(RMP2_ERR_UNSUPPORTED, "at most " + std::to_string(RMP2_MAX_LEAVES) + " leaves per tree");
  for (int i = 0; i < n_leaves; ++i) {
    int rc = check_leaf(leaves[i], rb->F, i);
    if (rc != RMP2_OK) return rc;
  }
  rmp2_tree* tr = new rmp2_tree;
  tr->robot = *rb;
  tr->leaves.assign(leaves, leaves + n_leaves);
  // tuning hooks, read once per tree (never on the step path); rmp2_tree_set_option overrides them
  if (getenv("RMP2_DISABLE_TMA")) tr->use_tma = false;
  if (const char* v = getenv("RMP2_SPLIT_RESOLVE")) tr->split_resolve = (v[0] == '1') ? 1 : 0;
  if (const char* v = getenv("RMP2_FORCE_BLOCK")) {
    const int b = atoi(v);
    if (b == 32 || b == 64 || b == 128) tr->force_block = b;
  }
  if (const char* v = getenv("RMP2_CHUNK_ENVS")) tr->chunk_envs = std::max(0LL, atoll(v));
  if (const char* v = getenv("RMP2_SPEC_STEP_BLOCK")) {
    const int b = atoi(v);
    if (b == 128 || b == 256) tr->spec_step_block = b;
  }
  if (const char* v = getenv("RMP2_MERGE_COINCIDENT")) tr->merge_coincident = (v[0] != '0');
  int rc = compile_tables(tr);
  if (rc != RMP2_OK) { delete tr; return rc; }
  *out = tr;
  return RMP2_OK;
}

void rmp2_tree_destroy(rmp2_tree* tree) {
  if (!tree) return;
  for (auto& s : tree->stage) {
    if (s.buf) cudaFree(s.buf);
    if (s.stream) cudaStreamDestroy(s.stream);
  }
  if (tree->spec || tree->rec || tree->mf || tree->fb) cudaDeviceSynchronize();   // queued steps may still use them
  if (tree->rec) cudaFree(tree->rec);
  if (tree->mf) cudaFree(tree->mf);
  if (tree->fb) cudaFree(tree->fb);
  rmp2_jit_destroy(tree->spec);
  for (auto& c : tree->clock)
    for (auto ev : c.pending) cudaEventDestroy(ev);
  delete tree;
}

int rmp2_tree_specialize(rmp2_tree* tree, int32_t flags) {
  if (!tree) return fail(RMP2_ERR_INVALID, "null argument");
  const bool compile_only = (flags & RMP2_SPECIALIZE_COMPILE_ONLY) != 0;
  SpecModule* m = nullptr;
  std::string err;
  if (rmp2_jit_build(tree->tab, rmp2_pick_width(tree->tab.n), compile_only, &m, err) != 0)
    return fail(err.rfind("NVRTC not", 0) == 0 ? RMP2_ERR_UNSUPPORTED : RMP2_ERR_CUDA, "rmp2_tree_specialize: " + err);
  if (!compile_only) {
    if (tree->spec) cudaDeviceSynchronize();      // kernels of the old module may still be queued
    rmp2_jit_destroy(tree->spec);
    tree->spec = m;
    tree->respecialize = true;
  }
  return RMP2_OK;
}

int rmp2_tree_is_specialized(const rmp2_tree* tree, double* compile_seconds) {
  if (compile_seconds) *compile_seconds = (tree && tree->spec) ? rmp2_jit_seconds(tree->spec) : 0.0;
  return (tree && tree->spec) ? 1 : 0;
}

int rmp2_tree_update_leaf(rmp2_tree* tree, int32_t index, const rmp2_leaf_desc* leaf) {
  if (!tree || !leaf) return fail(RMP2_ERR_INVALID, "null argument");
  if (index < 0 || index >= (int)tree->leaves.size()) return fail(RMP2_ERR_INVALID, "leaf index out of range");
  const rmp2_leaf_desc& old = tree->leaves[index];
  if (old.type != leaf->type || old.space != leaf->space || old.frame != leaf->frame || old.goal_slot != leaf->goal_slot)
    return fail(RMP2_ERR_INVALID, "update_leaf may change params/vec only; rebuild the tree to change its structure");
  // the tables are recompiled as a whole (host work of microseconds): a changed obstacle leaf may leave or join a group
  // of merged leaves.  On failure the tree keeps its previous tables and descriptor.
  const rmp2_leaf_desc previous = old;
  tree->leaves[index] = *leaf;
  const int rc = compile_tables(tree);
  if (rc != RMP2_OK) {
    tree->leaves[index] = previous;
    return rc;
  }
  if (tree->spec) {
    // The tables are compile-time constants of the specialised kernels: rebuild them for the new values (NVRTC,
    // tens of milliseconds once the compiler library is loaded), so that the reference idiom
    // `target_rmp.goal = ...` keeps the fast path.  If the rebuild fails the generic kernels take over.
    SpecModule* m = nullptr;
    std::string err;
    const int jrc = rmp2_jit_build(tree->tab, rmp2_pick_width(tree->tab.n), false, &m, err);
    cudaDeviceSynchronize();              // steps queued with the old module finish before it is unloaded
    rmp2_jit_destroy(tree->spec);
    tree->spec = (jrc == 0) ? m : nullptr;
  }
  return RMP2_OK;
}

}  // extern "C"

// ------------------------------------------------------------------------------------ step launch
namespace {

int pick_block(const rmp2_tree* tree, long long B) {
  if (tree->force_block) return tree->force_block;
  return (B >= 148LL * 4 * 128) ? 128 : (B >= 148LL * 4 * 64 ? 64 : 32);
}

// The staged (TMA bulk-copy) sphere path needs one tile of E rows (+ barrier) in shared memory.
bool tma_eligible(const rmp2_tree* tree, const StepArgs& A) {
  const int O = A.n_spheres;
  if (!tree->use_tma || O <= 0 || A.spheres == nullptr) return false;
  return rmp2_spheres_smem(tree->sph, O, true, true) <= 96 * 1024;
}

struct ScopedClock {               // brackets one launch with events when profiling is on
  KernelClock* c = nullptr;
  cudaStream_t stream;
  cudaEvent_t stop = nullptr;
  ScopedClock(rmp2_tree* tree, int which, cudaStream_t s) : stream(s) {
    if (!tree->profiling) return;
    c = &tree->clock[which];
    if (c->pending.size() >= 2 * 4096) drain(*c);       // bounded: fold finished pairs into the totals
    cudaEvent_t start;
    cudaEventCreate(&start);
    cudaEventCreate(&stop);
    cudaEventRecord(start, stream);
    c->pending.push_back(start);
    c->pending.push_back(stop);
  }
  ~ScopedClock() {
    if (c) cudaEventRecord(stop, stream);
  }
  static cudaError_t drain(KernelClock& c) {
    for (size_t i = 0; i + 1 < c.pending.size(); i += 2) {
      cudaError_t e = cudaEventSynchronize(c.pending[i + 1]);
      if (e != cudaSuccess) return e;
      float t = 0.f;
      cudaEventElapsedTime(&t, c.pending[i], c.pending[i + 1]);
      c.ms += t;
      c.launches += 1;
      cudaEventDestroy(c.pending[i]);
      cudaEventDestroy(c.pending[i + 1]);
    }
    c.pending.clear();
    return cudaSuccess;
  }
};

int build_args(const rmp2_tree* tree, const rmp2_step_io* io, StepArgs& A) {
  if (!tree || !io) return fail(RMP2_ERR_INVALID, "null argument");
  if (io->B < 0) return fail(RMP2_ERR_INVALID, "B must be >= 0");
  if (io->B > 0 && (!io->q || !io->qd || !io->qdd)) return fail(RMP2_ERR_INVALID, "q, qd and qdd are required");
  memset(&A, 0, sizeof(A));
  A.B = io->B;
  A.q = io->q;
  A.qd = io->qd;
  A.qdd = io->qdd;
  A.goals = io->goals;
  A.n_goal_slots = io->n_goal_slots;
  A.spheres = io->spheres;
  A.n_spheres = io->spheres ? io->n_spheres : 0;
  A.pairs = io->pairs;
  if (tree->n_goal_slots_used > 0 && (!io->goals || io->n_goal_slots < tree->n_goal_slots_used))
    return fail(RMP2_ERR_INVALID, "tree has per-environment goals: io.goals with n_goal_slots >= " +
                                      std::to_string(tree->n_goal_slots_used) + " is required");
  if (tree->tab.uses_spheres && io->n_spheres > 0 && !io->spheres)
    return fail(RMP2_ERR_INVALID, "n_spheres > 0 but spheres is NULL");
  if (io->n_spheres < 0) return fail(RMP2_ERR_INVALID, "n_spheres must be >= 0");
  // sphere rows are read 16 bytes at a time (TMA bulk copies / LDG.128)
  if (A.spheres && ((uintptr_t)A.spheres % 16) != 0)
    return fail(RMP2_ERR_INVALID, "io.spheres must be 16-byte aligned (rows of float4)");
  if (io->n_pair_sets != tree->n_pair_sets)
    return fail(RMP2_ERR_INVALID, "io.n_pair_sets (" + std::to_string(io->n_pair_sets) + ") != explicit-pair leaves of the tree (" +
                                      std::to_string(tree->n_pair_sets) + ")");
  int total = 0;
  for (int i = 0; i < tree->n_pair_sets; ++i) {
    if (io->pair_counts[i] < 0) return fail(RMP2_ERR_INVALID, "negative pair count");
    A.pair_off[i] = total;
    total += io->pair_counts[i];
  }
  A.pair_off[tree->n_pair_sets] = total;
  A.pair_total = total;
  A.early_out = tree->early_out ? 1 : 0;
  if (total > 0 && !io->pairs) return fail(RMP2_ERR_INVALID, "pair counts > 0 but pairs is NULL");
  return RMP2_OK;
}

// One control step over A.B environments on `stream`; A.rec / A.mf / A.fb are the scratch of this chunk.
int launch_chunk(rmp2_tree* tree, const StepArgs& A, cudaStream_t stream) {
  if (A.B == 0) return RMP2_OK;
  const StepTables& T = tree->tab;
  const int block = pick_block(tree, A.B);
  cudaError_t e;
  std::string jit_err;
  if (T.n_sphere_slots > 0 && A.n_spheres > 0) {
    {
      ScopedClock clk(tree, 0, stream);
      if (tree->spec) {
        const size_t smem = (size_t)T.n_slots * RMP2_CHAIN_FLOATS * block * sizeof(float);
        e = rmp2_jit_launch(tree->spec, 0, A, (unsigned)((A.B + block - 1) / block), block, smem, stream, jit_err);
      } else {
        e = rmp2_launch_frames(T, A, block, stream);
      }
    }
    if (e != cudaSuccess) return cuda_fail(e, (std::string("rmp2_frames_kernel launch") + (jit_err.empty() ? "" : ": " + jit_err)).c_str());
    const bool use_tma = tma_eligible(tree, A);
    {
      ScopedClock clk(tree, 1, stream);
      e = rmp2_launch_spheres(tree->sph, A, use_tma, stream);
    }
    if (e != cudaSuccess) return cuda_fail(e, "rmp2_spheres_kernel launch");
    g_launches.fetch_add(2);
  }
  {
    ScopedClock clk(tree, 2, stream);
    if (tree->spec) {
      // large batches: 256-thread blocks (instruction fetch, see RMP2_SPEC_STEP_THREADS in rmp2_tree_kernels.cuh)
      int sb = block;
      if (block == 128 && A.B >= 148LL * 2 * 256 * 2 && rmp2_pick_width(T.n) <= 7)
        sb = tree->spec_step_block ? tree->spec_step_block : 256;
      // which: 1 fused / 2 split at the launch's block size; 3 = the fused instance compiled for exactly
      // RMP2_SPEC_STEP_THREADS threads (shared-memory offsets as immediates)
      const int which = A.split ? 2 : ((sb == rmp2_jit_big_block(tree->spec, rmp2_pick_width(T.n))) ? 3 : 1);
      e = rmp2_jit_launch(tree->spec, which, A, (unsigned)((A.B + sb - 1) / sb), sb, rmp2_step_smem(T, sb), stream, jit_err);
    } else
      e = rmp2_launch_step(T, A, block, stream);
  }
  if (e != cudaSuccess) return cuda_fail(e, (std::string("rmp2_step_kernel launch") + (jit_err.empty() ? "" : ": " + jit_err)).c_str());
  g_launches.fetch_add(1);
  if (A.split) {
    {
      ScopedClock clk(tree, 3, stream);
      e = rmp2_launch_resolve(T, A, block, stream);
    }
    if (e != cudaSuccess) return cuda_fail(e, "rmp2_resolve_kernel launch");
    g_launches.fetch_add(1);
  }
  {
    ScopedClock clk(tree, 4, stream);
    e = rmp2_launch_fallback(T, A, 148 * 4, stream);
  }
  if (e != cudaSuccess) return cuda_fail(e, "rmp2_resolve_fallback_kernel launch");
  g_launches.fetch_add(1);
  return RMP2_OK;
}

// The direct resolve runs inside the step kernel (default), or as its own kernel behind the (M, f) scratch
// (RMP2_OPT_SPLIT_RESOLVE = 1).  Round 1 split them for large batches because the Jacobi sweeps sat in that code; with
// the sweeps moved to the fallback kernel the fused form is faster or equal on every configuration (measured, 2^20
// environments, ms per step fused / split: config 2 0.109 / 0.153, config 3 0.646 / 0.675, config 4 1.364 / 1.366,
// config 5 1.300 / 1.334) and saves the write + read of (M, f).
bool split_resolve(const rmp2_tree* tree, long long /*B*/) { return tree->split_resolve == 1; }

long long chunk_envs(const rmp2_tree* tree) { return tree->chunk_envs > 0 ? tree->chunk_envs : RMP2_STEP_CHUNK; }

size_t mf_floats_for(const rmp2_tree* tree, long long B) {
  const int N = rmp2_pick_width(tree->tab.n);
  // split mode: the combined (M, f), field-major; fused mode: the fallback's rows (RMP2_HANDOFF_ROW(N) floats each)
  const int handoff_row = (N * (N + 1) / 2 + 2 * N + 3) / 4 * 4;
  return (size_t)B * std::max(N * N + N, handoff_row);
}

size_t fb_ints_for(long long B) { return (size_t)B + 4; }

size_t rec_floats_for(const rmp2_tree* tree, long long B, int n_spheres) {
  if (tree->tab.n_sphere_slots == 0 || n_spheres <= 0) return 0;
  const size_t tiles = (size_t)((B + RMP2_REC_TILE - 1) / RMP2_REC_TILE);        // tiled over environments: rmp2_rec_base
  return tiles * RMP2_REC_TILE * tree->tab.n_sphere_slots * RMP2_REC_FLOATS;
}

// Size the tree's scratch for chunks of `chunk` environments.  Stream-ordered (cudaMallocAsync / cudaFreeAsync
// on the caller's stream): no device-wide stall; inside a stream capture growing is refused -- reserve first.
int ensure_scratch(rmp2_tree* tree, long long chunk, int n_spheres, cudaStream_t stream) {
  const size_t need_rec = rec_floats_for(tree, chunk, n_spheres);
  const size_t need_mf = mf_floats_for(tree, chunk);
  const size_t need_fb = fb_ints_for(chunk);
  if (need_rec <= tree->rec_floats && need_mf <= tree->mf_floats && need_fb <= tree->fb_ints) return RMP2_OK;
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(stream, &cap) == cudaSuccess && cap != cudaStreamCaptureStatusNone)
    return fail(RMP2_ERR_INVALID, "the tree's scratch must grow but the stream is being captured: call rmp2_tree_reserve "
                                  "(or run one step) before capturing");
  cudaError_t e;
  auto grow = [&](void** ptr, size_t* have, size_t need, size_t elem) -> cudaError_t {
    if (need <= *have) return cudaSuccess;
    if (*ptr) {
      cudaError_t e2 = cudaFreeAsync(*ptr, stream);       // ordered after the steps already queued on this stream
      if (e2 != cudaSuccess) return e2;
      *ptr = nullptr;
      *have = 0;
    }
    cudaError_t e2 = cudaMallocAsync(ptr, need * elem, stream);
    if (e2 == cudaSuccess) *have = need;
    return e2;
  };
  if ((e = grow((void**)&tree->rec, &tree->rec_floats, need_rec, sizeof(float))) != cudaSuccess)
    return cuda_fail(e, "allocating the sphere-record scratch");
  if ((e = grow((void**)&tree->mf, &tree->mf_floats, need_mf, sizeof(float))) != cudaSuccess)
    return cuda_fail(e, "allocating the (M, f) scratch");
  const size_t had_fb = tree->fb_ints;
  if ((e = grow((void**)&tree->fb, &tree->fb_ints, need_fb, sizeof(int32_t))) != cudaSuccess)
    return cuda_fail(e, "allocating the fallback work list");
  if (tree->fb_ints != had_fb) {
    e = cudaMemsetAsync(tree->fb, 0, 2 * sizeof(int32_t), stream);      // list length and block ticket
    if (e != cudaSuccess) return cuda_fail(e, "clearing the fallback work list");
  }
  return RMP2_OK;
}

// Device-pointer step, chunked so that the scratch stays bounded.
int launch(rmp2_tree* tree, const StepArgs& A0, cudaStream_t stream) {
  const long long B = A0.B;
  if (B == 0) return RMP2_OK;
  const long long chunk = std::min<long long>(B, chunk_envs(tree));
  int rc = ensure_scratch(tree, chunk, A0.n_spheres, stream);
  if (rc != RMP2_OK) return rc;
  const bool split = split_resolve(tree, chunk);
  const int n = tree->tab.n;
  for (long long e0 = 0; e0 < B; e0 += chunk) {
    StepArgs A = A0;
    A.B = std::min(chunk, B - e0);
    A.q = A0.q + e0 * n;
    A.qd = A0.qd + e0 * n;
    A.qdd = A0.qdd + e0 * n;
    if (A0.q_rw) A.q_rw = A0.q_rw + e0 * n;
    if (A0.qd_rw) A.qd_rw = A0.qd_rw + e0 * n;
    if (A0.goals) A.goals = A0.goals + e0 * A0.n_goal_slots * 3;
    if (A0.spheres) A.spheres = A0.spheres + e0 * (long long)A0.n_spheres * 4;
    if (A0.pairs) A.pairs = A0.pairs + e0 * (long long)A0.pair_total * RMP2_PAIR_FLOATS;
    A.rec = tree->rec;
    A.mf = tree->mf;
    A.fb = tree->fb;
    A.split = split ? 1 : 0;
    rc = launch_chunk(tree, A, stream);
    if (rc != RMP2_OK) return rc;
  }
  return RMP2_OK;
}

}  // namespace

extern "C" {

int rmp2_tree_reserve(rmp2_tree* tree, int64_t B, int32_t n_spheres, void* stream) {
  if (!tree) return fail(RMP2_ERR_INVALID, "null argument");
  if (B < 0 || n_spheres < 0) return fail(RMP2_ERR_INVALID, "B and n_spheres must be >= 0");
  if (B == 0) return RMP2_OK;
  tree->reserved_envs = std::max<long long>(tree->reserved_envs, B);
  tree->reserved_spheres = std::max(tree->reserved_spheres, n_spheres);
  return ensure_scratch(tree, std::min<long long>(B, chunk_envs(tree)), n_spheres, (cudaStream_t)stream);
}

int rmp2_step(const rmp2_tree* tree, const rmp2_step_io* io, void* stream) {
  StepArgs A;
  int rc = build_args(tree, io, A);
  if (rc != RMP2_OK) return rc;
  // the handle owns scratch: steps of one tree are serialised by the caller (one stream at a time)
  return launch(const_cast<rmp2_tree*>(tree), A, (cudaStream_t)stream);
}

int rmp2_rollout(const rmp2_tree* tree, const rmp2_step_io* io, float* q_inout, float* qd_inout, float dt,
                 int32_t n_steps, int32_t control_every, void* stream) {
  if (!tree || !io) return fail(RMP2_ERR_INVALID, "null argument");
  if (!q_inout || !qd_inout) return fail(RMP2_ERR_INVALID, "q_inout and qd_inout are required");
  if (n_steps <= 0 || control_every <= 0) return fail(RMP2_ERR_INVALID, "n_steps and control_every must be positive");
  rmp2_step_io tmp = *io;
  tmp.q = q_inout;
  tmp.qd = qd_inout;
  StepArgs A;
  int rc = build_args(tree, &tmp, A);
  if (rc != RMP2_OK) return rc;
  A.q_rw = q_inout;
  A.qd_rw = qd_inout;
  A.dt = dt;
  // control step + `control_every` held-command Euler sub-steps, repeated; all asynchronous
  for (int done = 0; done < n_steps; done += control_every) {
    A.n_sim_steps = std::min(control_every, n_steps - done);
    rc = launch(const_cast<rmp2_tree*>(tree), A, (cudaStream_t)stream);
    if (rc != RMP2_OK) return rc;
  }
  return RMP2_OK;
}

int rmp2_step_host(rmp2_tree* tree, const rmp2_step_io* io) {
  StepArgs A0;
  int rc = build_args(tree, io, A0);
  if (rc != RMP2_OK) return rc;
  const long long B = io->B;
  if (B == 0) return RMP2_OK;
  const int n = tree->tab.n;
  const int G = io->goals ? io->n_goal_slots : 0;
  const int O = A0.n_spheres;
  const int K = A0.pair_total;
  // floats per environment, each section padded to 4 floats so every section stays 16-byte aligned
  auto pad4 = [](size_t v) { return (v + 3) & ~size_t(3); };
  const long long chunk = std::min<long long>(B, 65536);
  const size_t off_q = 0;
  const size_t off_qd = off_q + pad4((size_t)chunk * n);
  const size_t off_qdd = off_qd + pad4((size_t)chunk * n);
  const size_t off_goal = off_qdd + pad4((size_t)chunk * n);
  const size_t off_sph = off_goal + pad4((size_t)chunk * G * 3);
  const size_t off_pair = off_sph + pad4((size_t)chunk * O * 4);
  const size_t off_rec = off_pair + pad4((size_t)chunk * K * RMP2_PAIR_FLOATS);
  const size_t off_mf = off_rec + pad4(rec_floats_for(tree, chunk, O));
  const size_t off_fb = off_mf + pad4(mf_floats_for(tree, chunk));
  const bool split = split_resolve(tree, chunk);
  const size_t total = off_fb + pad4(fb_ints_for(chunk));
  // on any error below: wait for the copies already queued on the caller's host buffers before returning
  auto drain = [&]() {
    for (auto& s : tree->stage)
      if (s.stream) cudaStreamSynchronize(s.stream);
  };
  for (auto& s : tree->stage) {
    if (!s.stream) {
      cudaError_t e = cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking);
      if (e != cudaSuccess) return cuda_fail(e, "cudaStreamCreate");
    }
    if (s.floats < total) {
      if (s.buf) cudaFree(s.buf);
      s.buf = nullptr;
      s.floats = 0;
      cudaError_t e = cudaMalloc(&s.buf, total * sizeof(float));
      if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc staging");
      s.floats = total;
    }
    // (re)laid-out staging: the work list of this slot starts empty
    cudaError_t e = cudaMemsetAsync(s.buf + off_fb, 0, 2 * sizeof(int32_t), s.stream);
    if (e != cudaSuccess) return cuda_fail(e, "clearing the fallback work list");
  }
  int slot = 0;
  for (long long e0 = 0; e0 < B; e0 += chunk, slot = (slot + 1) % 3) {
    const long long cb = std::min(chunk, B - e0);
    HostStage& s = tree->stage[slot];
    cudaError_t e;
#define RMP2_H2D(dst, src, count)                                                                         \
  if ((count) > 0) {                                                                                       \
    e = cudaMemcpyAsync(s.buf + (dst), (src), (size_t)(count) * sizeof(float), cudaMemcpyHostToDevice, s.stream); \
    if (e != cudaSuccess) {                                                                                 \
      drain();                                                                                              \
      return cuda_fail(e, "H2D copy");                                                                      \
    }                                                                                                       \
  }
    RMP2_H2D(off_q, io->q + e0 * n, cb * n);
    RMP2_H2D(off_qd, io->qd + e0 * n, cb * n);
    RMP2_H2D(off_goal, io->goals ? io->goals + e0 * G * 3 : nullptr, cb * G * 3);
    RMP2_H2D(off_sph, io->spheres ? io->spheres + e0 * O * 4 : nullptr, cb * O * 4);
    RMP2_H2D(off_pair, io->pairs ? io->pairs + e0 * K * RMP2_PAIR_FLOATS : nullptr, cb * K * RMP2_PAIR_FLOATS);
#undef RMP2_H2D
    StepArgs A = A0;
    A.B = cb;
    A.q = s.buf + off_q;
    A.qd = s.buf + off_qd;
    A.qdd = s.buf + off_qdd;
    A.goals = G ? s.buf + off_goal : nullptr;
    A.spheres = O ? s.buf + off_sph : nullptr;
    A.pairs = K ? s.buf + off_pair : nullptr;
    A.rec = s.buf + off_rec;
    A.mf = s.buf + off_mf;
    A.fb = reinterpret_cast<int32_t*>(s.buf + off_fb);
    A.split = split ? 1 : 0;
    rc = launch_chunk(tree, A, s.stream);
    if (rc != RMP2_OK) {
      drain();
      return rc;
    }
    e = cudaMemcpyAsync(io->qdd + e0 * n, s.buf + off_qdd, (size_t)cb * n * sizeof(float), cudaMemcpyDeviceToHost, s.stream);
    if (e != cudaSuccess) {
      drain();
      return cuda_fail(e, "D2H copy");
    }
  }
  for (auto& s : tree->stage) {
    cudaError_t e = cudaStreamSynchronize(s.stream);
    if (e != cudaSuccess) return cuda_fail(e, "rmp2_step_host");
  }
  return RMP2_OK;
}

int rmp2_pinv_solve(int32_t n, int64_t B, const float* M, const float* f, float* x, int32_t pivot, int32_t mode,
                    void* stream) {
  if (!M || !f || !x) return fail(RMP2_ERR_INVALID, "M, f and x are required");
  if (n <= 0 || n > RMP2_MAX_JOINTS) return fail(RMP2_ERR_INVALID, "n out of range");
  if (mode != 0 && mode != 1) return fail(RMP2_ERR_INVALID, "mode must be 0 (as in the step) or 1 (Jacobi only)");
  if (B <= 0) return B == 0 ? RMP2_OK : fail(RMP2_ERR_INVALID, "B must be >= 0");
  const float rcond = (float)(10.0 * n * 1.1920928955078125e-07);
  cudaError_t e = rmp2_launch_pinv(n, rcond, pivot != 0, mode, B, M, f, x, (cudaStream_t)stream);
  if (e != cudaSuccess) return cuda_fail(e, "rmp2_pinv_solve launch");
  g_launches.fetch_add(1);
  return RMP2_OK;
}

int rmp2_fk(const rmp2_robot* rb, int32_t frame, int64_t B, const float* q, const float* qd, float* x, float* xd,
            float* J, float* c, void* stream) {
  if (!rb || !q || !x) return fail(RMP2_ERR_INVALID, "robot, q and x are required");
  if (frame < 0 || frame >= rb->F) return fail(RMP2_ERR_INVALID, "frame index out of range");
  if ((xd || c) && !qd) return fail(RMP2_ERR_INVALID, "xd / c need qd");
  if (B <= 0) return B == 0 ? RMP2_OK : fail(RMP2_ERR_INVALID, "B must be >= 0");
  StepTables T;
  memset(&T, 0, sizeof(T));
  T.n = rb->n;
  T.prismatic_mask = prismatic_mask(*rb);
  std::vector<int> path;
  for (int k = frame; k >= 0; k = rb->parent[k]) path.push_back(k);
  std::reverse(path.begin(), path.end());
  for (int k : path) {
    FrameTab& ft = T.frames[T.n_frames++];
    fill_frame(*rb, k, ft);
    ft.anc_mask = ancestor_mask(*rb, k);
  }
  cudaError_t e = rmp2_launch_fk(T, B, q, qd, x, xd, J, c, (cudaStream_t)stream);
  if (e != cudaSuccess) return cuda_fail(e, "rmp2_fk launch");
  g_launches.fetch_add(1);
  return RMP2_OK;
}

int rmp2_obstacle_feed(const rmp2_robot* rb, const int32_t* frames, const float* link_capsules, int32_t n_frames,
                       int64_t B, const float* q, const float* spheres, int32_t n_spheres, const float* capsules,
                       int32_t n_capsules, float* pairs, float* aux, void* stream) {
  if (!rb || !frames || !q || !pairs) return fail(RMP2_ERR_INVALID, "robot, frames, q and pairs are required");
  if (n_frames <= 0 || n_frames > RMP2_MAX_LEAVES) return fail(RMP2_ERR_INVALID, "n_frames out of range");
  if (n_spheres < 0 || n_capsules < 0 || (n_spheres > 0 && !spheres) || (n_capsules > 0 && !capsules))
    return fail(RMP2_ERR_INVALID, "obstacle arrays do not match their counts");
  if ((spheres && ((uintptr_t)spheres % 16) != 0) || (capsules && ((uintptr_t)capsules % 16) != 0))
    return fail(RMP2_ERR_INVALID, "spheres and capsules must be 16-byte aligned (rows of float4)");
  if (B <= 0) return B == 0 ? RMP2_OK : fail(RMP2_ERR_INVALID, "B must be >= 0");
  // marker leaves make the tree compiler produce the pruned, depth-first frame table
  std::vector<rmp2_leaf_desc> marks(n_frames);
  for (int i = 0; i < n_frames; ++i) {
    memset(&marks[i], 0, sizeof(rmp2_leaf_desc));
    marks[i].type = RMP2_LEAF_OBSTACLE_AVOIDANCE;
    marks[i].space = RMP2_SPACE_FRAME_DISTANCE_SPHERES;
    marks[i].frame = frames[i];
    marks[i].goal_slot = -1;
    for (int k = 0; k < 11; ++k) marks[i].params[k] = 1.f;
  }
  rmp2_tree* tmp = nullptr;
  int rc = rmp2_tree_create(rb, marks.data(), n_frames, &tmp);
  if (rc == RMP2_OK) rc = rmp2_tree_set_option(tmp, RMP2_OPT_MERGE_COINCIDENT, 0);   // every listed frame keeps its own rows
  if (rc != RMP2_OK) {
    rmp2_tree_destroy(tmp);
    return rc;
  }
  for (int i = 0; i < n_frames; ++i) tmp->tab.leaves[tmp->table_index[i]].pair_set = i;
  FeedArgs A;
  A.B = B;
  A.q = q;
  A.spheres = spheres;
  A.capsules = capsules;
  A.pairs = pairs;
  A.aux = aux;
  A.n_spheres = n_spheres;
  A.n_capsules = n_capsules;
  A.n_listed = n_frames;
  FeedLinks LK;
  memset(&LK, 0, sizeof(LK));
  if (link_capsules)
    for (int i = 0; i < n_frames; ++i)
      for (int k = 0; k < 7; ++k) LK.c[i][k] = link_capsules[(size_t)i * 8 + k];
  cudaError_t e = rmp2_launch_feed(tmp->tab, LK, A, (cudaStream_t)stream);
  rmp2_tree_destroy(tmp);
  if (e != cudaSuccess) return cuda_fail(e, "rmp2_obstacle_feed launch");
  g_launches.fetch_add(1);
  return RMP2_OK;
}

int rmp2_leaf_evaluate(const rmp2_leaf_desc* leaf, int32_t m, int64_t K, const float* x, const float* xd,
                       const float* aux, float* xdd, float* M, void* stream) {
  if (!leaf || !x || !xd || !xdd || !M) return fail(RMP2_ERR_INVALID, "null argument");
  if (leaf->type == RMP2_LEAF_COLLISION_AVOIDANCE && (m != 3 || !aux))
    return fail(RMP2_ERR_INVALID, "CollisionAvoidance is three-dimensional and needs aux = (distance, normal)");
  if (m <= 0 || m > RMP2_MAX_JOINTS) return fail(RMP2_ERR_INVALID, "task dimension out of range");
  if (K <= 0) return K == 0 ? RMP2_OK : fail(RMP2_ERR_INVALID, "K must be >= 0");
  if (leaf->type == RMP2_LEAF_OBSTACLE_AVOIDANCE && m != 1) return fail(RMP2_ERR_INVALID, "ObstacleAvoidance is one-dimensional");
  if (leaf->type == RMP2_LEAF_TARGET_ATTRACTOR && m != 3) return fail(RMP2_ERR_INVALID, "TargetAttractor is three-dimensional here");
  LeafTab L;
  LeafVec V;
  memset(&L, 0, sizeof(L));
  memset(&V, 0, sizeof(V));
  L.type = leaf->type;
  L.space = leaf->space;
  rmp2_leaf_desc d = *leaf;
  d.space = RMP2_SPACE_CONFIG;                    // vec length follows m
  int vlen = 0;
  int rc = derive_leaf_params(d, m, L, V.v, &vlen);
  if (rc != RMP2_OK) return rc;
  cudaError_t e = rmp2_launch_leaf(L, V, m, K, x, xd, aux, xdd, M, (cudaStream_t)stream);
  if (e != cudaSuccess) return cuda_fail(e, "rmp2_leaf_evaluate launch");
  g_launches.fetch_add(1);
  return RMP2_OK;
}

int rmp2_tree_kernel_info(const rmp2_tree* tree, int32_t which, int32_t n_spheres, int32_t* regs, int32_t* smem_bytes,
                          int32_t* blocks_per_sm, int32_t* block_threads) {
  if (!tree) return fail(RMP2_ERR_INVALID, "null argument");
  if (which < 0 || which > 5)
    return fail(RMP2_ERR_INVALID, "which must be 0 (frames), 1 (spheres), 2 (step, fused), 3 (step, split), 4 (resolve) or "
                                  "5 (resolve fallback)");
  int block = RMP2_BLOCK_THREADS;
  size_t smem = (which >= 4) ? 0 : (which >= 2 ? rmp2_step_smem(tree->tab, block)
                                                 : (size_t)tree->tab.n_slots * RMP2_CHAIN_FLOATS * block * sizeof(float));
  bool use_tma = false;
  if (which == 1) {
    if (tree->tab.n_sphere_slots == 0) return fail(RMP2_ERR_INVALID, "tree has no sphere-obstacle leaves");
    use_tma = n_spheres > 0 && rmp2_spheres_smem(tree->sph, n_spheres, true, true) <= 96 * 1024;
    block = ((tree->sph.envs_per_block * tree->sph.n_slots + 31) / 32) * 32;
    smem = rmp2_spheres_smem(tree->sph, n_spheres, use_tma, tree->early_out);
  }
  int r = 0, bps = 0;
  cudaError_t e = rmp2_kernel_attributes(tree->tab.n, which, use_tma, block, smem, &r, &bps);
  if (e != cudaSuccess) return cuda_fail(e, "rmp2_tree_kernel_info");
  if (tree->spec) {                         // specialised frames / step kernels: their own register count
    const int w = (which == 0) ? 0 : (which == 2 ? 1 : (which == 3 ? 2 : -1));
    const int r2 = (w >= 0) ? rmp2_jit_registers(tree->spec, w) : 0;
    if (r2 > 0) r = r2;
  }
  if (regs) *regs = r;
  if (smem_bytes) *smem_bytes = (int)smem;
  if (blocks_per_sm) *blocks_per_sm = bps;
  if (block_threads) *block_threads = block;
  return RMP2_OK;
}

int rmp2_tree_set_option(rmp2_tree* tree, int32_t option, int32_t value) {
  if (!tree) return fail(RMP2_ERR_INVALID, "null argument");
  switch (option) {
    case RMP2_OPT_EARLY_OUT:
      tree->early_out = value != 0;
      return RMP2_OK;
    case RMP2_OPT_TMA:
      tree->use_tma = value != 0;
      return RMP2_OK;
    case RMP2_OPT_SPLIT_RESOLVE:
      if (value < -1 || value > 1) return fail(RMP2_ERR_INVALID, "RMP2_OPT_SPLIT_RESOLVE takes -1 (by batch size), 0 or 1");
      tree->split_resolve = value;
      return RMP2_OK;
    case RMP2_OPT_BLOCK_THREADS:
      if (value != 0 && value != 32 && value != 64 && value != 128)
        return fail(RMP2_ERR_INVALID, "RMP2_OPT_BLOCK_THREADS takes 0 (by batch size), 32, 64 or 128");
      tree->force_block = value;
      return RMP2_OK;
    case RMP2_OPT_CHUNK_ENVS:
      if (value < 0) return fail(RMP2_ERR_INVALID, "RMP2_OPT_CHUNK_ENVS must be >= 0");
      tree->chunk_envs = value;
      return RMP2_OK;
    case RMP2_OPT_MERGE_COINCIDENT: {
      const bool on = value != 0;
      if (on == tree->merge_coincident) return RMP2_OK;
      tree->merge_coincident = on;
      const int rc = compile_tables(tree);
      if (rc != RMP2_OK) {
        tree->merge_coincident = !on;
        return rc;
      }
      if (tree->spec) {                       // the tables are compile-time constants of the specialised kernels
        SpecModule* m = nullptr;
        std::string err;
        const int jrc = rmp2_jit_build(tree->tab, rmp2_pick_width(tree->tab.n), false, &m, err);
        cudaDeviceSynchronize();              // steps queued with the old module finish before it is unloaded
        rmp2_jit_destroy(tree->spec);
        tree->spec = (jrc == 0) ? m : nullptr;
      }
      return RMP2_OK;
    }
    default:
      return fail(RMP2_ERR_INVALID, "unknown option " + std::to_string(option));
  }
}

int rmp2_tree_obstacle_slots(const rmp2_tree* tree, int32_t* n_leaves, int32_t* n_slots) {
  if (!tree) return fail(RMP2_ERR_INVALID, "null argument");
  if (n_leaves) *n_leaves = tree->tab.n_sphere_slots + tree->n_merged;
  if (n_slots) *n_slots = tree->tab.n_sphere_slots;
  return RMP2_OK;
}

int rmp2_tree_profile(rmp2_tree* tree, int32_t enable) {
  if (!tree) return fail(RMP2_ERR_INVALID, "null argument");
  tree->profiling = enable != 0;
  return RMP2_OK;
}

int rmp2_tree_profile_read(rmp2_tree* tree, double* ms, int64_t* launches) {
  if (!tree || !ms || !launches) return fail(RMP2_ERR_INVALID, "null argument");
  for (int k = 0; k < RMP2_N_CLOCKS; ++k) {
    KernelClock& c = tree->clock[k];
    cudaError_t e = ScopedClock::drain(c);
    if (e != cudaSuccess) return cuda_fail(e, "rmp2_tree_profile_read");
    ms[k] = c.ms;
    launches[k] = c.launches;
    c.ms = 0.0;
    c.launches = 0;
  }
  return RMP2_OK;
}

}  // extern "C"
