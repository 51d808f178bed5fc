// Kernel-side tables of one compiled RMP tree.  Built on the host by rmp2_tree_create and
// passed BY VALUE as a __grid_constant__ kernel parameter, so every field is read through the
// constant bank (uniform across the warp) and no global state is shared between trees.
#pragma once
#include <stdint.h>
#include "../../include/rmp2_b200.h"

#define RMP2_MAX_SLOTS 4          // saved chain states for branching kinematic trees
#define RMP2_VECPOOL 160          // floats: goals / q0 / limits of all leaves
#define RMP2_SLOT_BASE (-2)       // restore_slot value meaning "start from the base link"
#define RMP2_CHAIN_FLOATS 24      // R(9) p(3) w(3) v(3) alpha(3) a(3)
#define RMP2_PAIR_FLOATS 8        // floats per explicit pair row (rmp2_step_io.pairs)
#define RMP2_REC_FLOATS 9         // fields of one frame record (p, v, a); the (S, g) sums overwrite them in place
#define RMP2_REC_TILE 128         // environments per tile of the record scratch (see rmp2_rec_base)

// The record scratch is tiled over the environment axis: rec[env / TILE][field * L + slot][env % TILE] (L = record
// slots per environment).  Consecutive environments stay consecutive floats (full 128-byte lines per warp access), and
// the distance between two fields of one environment is (L * TILE) floats -- a compile-time constant in the
// tree-specialised kernels, so every record access there is base + immediate instead of a 64-bit add per field (the
// field-major layout [field][slot][B] cost ~2 address instructions per access in the step kernel).
// -> offset of (field 0, slot 0) of environment `env`; field f / slot s add (f * L + s) * RMP2_REC_TILE.
#if defined(__CUDACC__)
__host__ __device__
#endif
inline size_t rmp2_rec_base(long long env, int L) {
  return (size_t)(env / RMP2_REC_TILE) * ((size_t)RMP2_REC_FLOATS * L * RMP2_REC_TILE) + (size_t)(env % RMP2_REC_TILE);
}

struct FrameTab {
  float R[9];            // constant rotation  (reference: kinematics.py:202, R_x R_y R_z order)
  float t[3];            // constant translation
  float axis[3];         // joint axis, in the frame after the constant transform
  int32_t type;          // RMP2_JOINT_*
  int32_t qidx;          // column of q, -1 = none
  int32_t restore_slot;  // -1: parent is the previous frame in execution order
                         // RMP2_SLOT_BASE: parent is the base; >= 0: reload saved chain state
  int32_t save_slot;     // >= 0: store the chain state after this frame (it has several children)
  uint32_t anc_mask;     // bit j: joint column j lies on the path base -> this frame
  int32_t leaf_begin;    // leaves attached to this frame: [leaf_begin, leaf_end)
  int32_t leaf_end;
  int32_t ref_index;     // index in the reference's frame order (for diagnostics)
  int32_t const_rot_identity;  // R == I: skip the constant-rotation product
  int32_t axis_is_z;           // axis == (0,0,1): the joint rotation only mixes two columns
};

struct LeafTab {
  int32_t type;          // RMP2_LEAF_*
  int32_t space;         // RMP2_SPACE_*
  int32_t goal_slot;     // >= 0: per-environment goal
  int32_t vec_off;       // offset of this leaf's vector parameters in vecpool
  int32_t pair_set;      // FRAME_DISTANCE_PAIRS: index into the per-call pair offsets
  int32_t sphere_slot;   // FRAME_DISTANCE_SPHERES: record slot of this leaf in StepArgs::rec
  float p[RMP2_LEAF_PARAMS];   // derived parameters, see leaf_params.h
};

struct StepTables {
  int32_t n;                    // controllable joints
  int32_t n_frames;             // frames executed (unused ones pruned), depth-first order
  int32_t n_frame_leaves;       // leaves [0, n_frame_leaves) hang on frames
  int32_t n_leaves;             // leaves [n_frame_leaves, n_leaves) are configuration-space
  int32_t n_slots;
  int32_t uses_spheres;         // some leaf reads io.spheres
  int32_t n_sphere_slots;       // number of FRAME_DISTANCE_SPHERES leaves (record slots per environment)
  int32_t precondition;         // no leaf adds a positive multiple of I: M may be rank deficient -> pivoted-QR
                                // preconditioning of the Jacobi resolve pays off
  uint32_t prismatic_mask;      // bit j: joint column j is prismatic
  float rcond;                  // 10 * n * eps32 (tf.linalg.pinv default, rmp.py:153)
  FrameTab frames[RMP2_MAX_FRAMES];
  LeafTab leaves[RMP2_MAX_LEAVES];
  float vecpool[RMP2_VECPOOL];
};

// Parameters of the sphere-obstacle leaves, one row per record slot (rmp2_spheres_kernel).
struct SphereTables {
  int32_t n_slots;              // L
  int32_t envs_per_block;       // E: environments per thread block (E * L <= 128)
  int32_t div_magic;            // 65536 / E + 1: t / E == (t * div_magic) >> 16 for every thread index t < 512
  int32_t div_magic_slots;      // 65536 / L + 1: t / L likewise
  int32_t slot_fastest;         // early-out variant: threads ordered (environment, slot) instead of (slot, environment)
  float p[RMP2_MAX_LEAVES][RMP2_LEAF_PARAMS];
};

struct FeedArgs {
  long long B;
  const float* q;
  const float* spheres;   // [B][n_spheres][4]
  const float* capsules;  // [B][n_capsules][8]
  float* pairs;           // [B][n_listed * (n_spheres + n_capsules)][8]
  float* aux;             // [B][n_listed * (n_spheres + n_capsules)][4] or NULL
  int32_t n_spheres, n_capsules, n_listed;
};

// Control geometry of the listed frames for the obstacle feed: one capsule per frame in FRAME coordinates
// (ax ay az bx by bz radius 0); a zero-length, zero-radius capsule at the origin = the frame origin as control point.
struct FeedLinks {
  float c[RMP2_MAX_LEAVES][8];
};

struct ResolveArgs {
  int32_t n;
  float rcond;
};

// per-call arguments (device pointers)
struct StepArgs {
  long long B;
  const float* q;
  const float* qd;
  float* qdd;
  const float* goals;
  const float* spheres;
  const float* pairs;
  float* rec;            // [ceil(B / TILE)][9 * n_sphere_slots][TILE] scratch (rmp2_rec_base): frame records in, (S, g) sums out
  float* mf;             // [N*N + N][B] scratch (N = kernel width): combined M and f between the step and the resolve
                         // kernel (split mode); in either mode the factorised problems handed to the fallback kernel
  int32_t* fb;           // fallback work list: [0] length, [1] block ticket, [2 ...] environment indices
  int32_t split;         // 1: the step kernel stops at (M, f) and rmp2_resolve_kernel follows; 0: resolve fused
  int32_t n_goal_slots;
  int32_t n_spheres;
  int32_t pair_total;
  int32_t early_out;     // spheres kernel: skip pairs beyond the metric radius (exact, see rmp2.py:194)
  int32_t pair_off[RMP2_MAX_PAIR_SETS + 1];
  // rollout (rmp2_rollout): when n_sim_steps > 0 the step kernel integrates in place afterwards
  float* q_rw;
  float* qd_rw;
  float dt;
  int32_t n_sim_steps;
  int32_t control_every;
};
