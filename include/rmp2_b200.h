/* rmp2_b200 -- C ABI of the B200-native RMP2 control-step engine.
 *
 * One control step of an RMP tree (URDF forward kinematics -> task maps -> leaf policies ->
 * pullback sums J^T M J and J^T M (a - Jdot qdot) -> pseudo-inverse resolve), batched over B
 * independent robot environments.  Plain C types, device pointers and sizes only: no torch
 * types cross this boundary.  The reference (TomGoesGitHub/Riemannian-Motion-Policies) has no
 * FFI of its own -- its boundary is the Python class API -- so every entry point below names the
 * reference Python interface it stands in for.  INTEGRATION.md shows the ctypes binding.
 *
 * Conventions
 *   - all arrays are float32, row-major, densely packed; "device" pointers are CUDA device
 *     memory of the current device, borrowed for the duration of the call (never retained);
 *   - every call that launches work is asynchronous on the caller's CUDA stream (`stream` is a
 *     cudaStream_t passed as void*; NULL = the legacy default stream);
 *   - functions return 0 (RMP2_OK) or an error code; the message is read with
 *     rmp2_last_error() (thread-local).  Nothing throws across the ABI;
 *   - there is no CPU fallback anywhere: without a CUDA device the compute calls fail with
 *     RMP2_ERR_CUDA.
 */
#ifndef RMP2_B200_H_
#define RMP2_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RMP2_OK 0
#define RMP2_ERR_INVALID 1      /* bad argument (message says which)                       */
#define RMP2_ERR_CUDA 2         /* CUDA runtime error, no device, launch failure            */
#define RMP2_ERR_UNSUPPORTED 3  /* a tree / robot shape the kernels do not implement       */

#define RMP2_MAX_FRAMES 24      /* URDF joints (= frames) per robot                         */
#define RMP2_MAX_JOINTS 12      /* controllable joints n                                    */
#define RMP2_MAX_LEAVES 40      /* leaf policies per tree                                   */
#define RMP2_LEAF_PARAMS 16
#define RMP2_MAX_GOAL_SLOTS 4   /* per-environment 3-D goals                                 */
#define RMP2_MAX_PAIR_SETS 24   /* leaves fed by explicit closest-point pairs                */

/* joint types, reference: kinematics.py:205-209 */
enum { RMP2_JOINT_FIXED = 0, RMP2_JOINT_REVOLUTE = 1, RMP2_JOINT_PRISMATIC = 2 };

/* leaf policies; params[] layout is given next to each */
enum {
  /* rmp.py:226-261  TargetPolicy: params = {alpha, beta, c}; vec = goal[dim]               */
  RMP2_LEAF_TARGET_POLICY = 1,
  /* rmp.py:318-347  ConfigurationSpaceBiasing: params = {gamma_p, gamma_d, w}; vec = q0[n] */
  RMP2_LEAF_CONFIG_BIASING = 2,
  /* rmp.py:349-382  JointLimitAvoidance: params = {gamma_p, gamma_d};
   *                 vec = lower[n] followed by upper[n]                                      */
  RMP2_LEAF_JOINT_LIMIT = 3,
  /* rmp2.py:31-83   TargetAttractor: params = {accel_p_gain, accel_d_gain, accel_norm_eps,
   *                 metric_alpha_length_scale, min_metric_alpha, max_metric_scalar,
   *                 min_metric_scalar, proximity_metric_boost_scalar,
   *                 proximity_metric_boost_length_scale}; vec = goal[3]                      */
  RMP2_LEAF_TARGET_ATTRACTOR = 4,
  /* rmp2.py:86-112  JointVelocityCap: params = {max_velocity, velocity_damping_region,
   *                 damping_gain, metric_weight}                                             */
  RMP2_LEAF_VELOCITY_CAP = 5,
  /* rmp2.py:115-137 JointDamping: params = {accel_d_gain, metric_scalar, inertia}          */
  RMP2_LEAF_JOINT_DAMPING = 6,
  /* rmp2.py:140-196 ObstacleAvoidance: params = {margin, damping_gain, damping_std_dev,
   *                 damping_robustness_eps, damping_velocity_gate_length_scale,
   *                 repulsion_gain, repulsion_std_dev, metric_modulation_radius,
   *                 metric_scalar, metric_exploder_std_dev, metric_exploder_eps}             */
  RMP2_LEAF_OBSTACLE_AVOIDANCE = 7,
  /* rmp2.py:198-226 CSpaceBiasing: params = {metric_scalar, position_gain, damping_gain,
   *                 robust_position_term_thresh, inertia}; vec = goal[n]                     */
  RMP2_LEAF_CSPACE_BIASING = 8,
  /* rmp.py:264-315  CollisionAvoidance (v1): params = {eta_rep, nu_rep, eta_damp, nu_damp, r, c};
   *                 distance d and unit normal vec of every pair come with the pair rows             */
  RMP2_LEAF_COLLISION_AVOIDANCE = 9
};

/* task map a leaf lives on (the closed set of chains the reference's experiments build) */
enum {
  /* taskmap.py:13-20   IdentityTaskmap                                                      */
  RMP2_SPACE_CONFIG = 0,
  /* taskmap.py:22-54   chain [TaskmapByForwardKinematic(frame), TaskmapFrom4x4ToPosition]   */
  RMP2_SPACE_FRAME_POSITION = 1,
  /* taskmap.py:115-138 chain [FK(frame), TaskmapJointFrame4x4ToDistance]; the K pairs are the
   * O spheres of rmp2_step_io.spheres: pos_on_link = frame origin, pos_on_obstacle = closest
   * point of the sphere surface to it                                                         */
  RMP2_SPACE_FRAME_DISTANCE_SPHERES = 2,
  /* same chain, the K pairs given explicitly (the Datamanager feed, data_management.py:22-37);
   * pair row = (pos_on_link xyz, pos_on_obstacle xyz, 0, 0)                                    */
  RMP2_SPACE_FRAME_DISTANCE_PAIRS = 3,
  /* taskmap.py:79-99   chain [FK(frame), TaskmapRelative4x4(relative_pos), TaskmapFrom4x4ToPosition]:
   * K points fixed in the frame; pair row = (relative_pos xyz, distance, normal_vec xyz, 0)     */
  RMP2_SPACE_FRAME_POINTS = 4,
  /* taskmap.py:57-67   chain [FK(frame), TaskmapFrom4x4ToEuler]: xyz Euler angles of the frame's rotation
   * (kinematics.py:74-96), differentiated analytically with the Euler-rate map omega = H(theta) thetadot
   * (helper/trigonometry_helper.py:18-38): J = H^-1 J_omega, c = d/dt(H^-1) omega + H^-1 alpha.
   * Leaves: TargetPolicy / TargetAttractor with a 3-vector goal of Euler angles.                        */
  RMP2_SPACE_FRAME_EULER = 5
};

typedef struct rmp2_robot rmp2_robot; /* constant kinematic tables of one URDF              */
typedef struct rmp2_tree rmp2_tree;   /* one RmpCore: robot + ordered leaf list              */

/* One leaf policy.  Stands in for one entry of RmpCore.rmps (rmp.py:115,127-128). */
typedef struct rmp2_leaf_desc {
  int32_t type;      /* RMP2_LEAF_*                                                          */
  int32_t space;     /* RMP2_SPACE_*                                                         */
  int32_t frame;     /* frame index (order of rmp2_robot_create) for FRAME_* spaces, else -1 */
  int32_t goal_slot; /* >= 0: goal read per environment from rmp2_step_io.goals; -1: vec     */
  float params[RMP2_LEAF_PARAMS];
  float vec[2 * RMP2_MAX_JOINTS];
} rmp2_leaf_desc;

/* Inputs and output of one batched control step.  Stands in for the arguments of
 * RmpCore.evaluate(q, qd) (rmp.py:133) plus the state the reference leaves capture by
 * reference: target_rmp.goal and the Datamanager variables (data_management.py:8-17). */
typedef struct rmp2_step_io {
  int64_t B;              /* number of environments                                          */
  const float* q;         /* [B][n]                                                          */
  const float* qd;        /* [B][n]                                                          */
  float* qdd;             /* [B][n]  out                                                     */
  const float* goals;     /* [B][n_goal_slots][3] or NULL                                    */
  int32_t n_goal_slots;
  int32_t n_spheres;      /* O, spheres per environment                                      */
  const float* spheres;   /* [B][O][4] = (cx, cy, cz, radius) or NULL; 16-byte aligned        */
  const float* pairs;     /* [B][K_total][8] pair rows (layout per space, see above) or NULL   */
  int32_t n_pair_sets;    /* number of FRAME_DISTANCE_PAIRS / FRAME_POINTS leaves, tree order  */
  int32_t pair_counts[RMP2_MAX_PAIR_SETS]; /* K of each such leaf; K_total = their sum        */
} rmp2_step_io;

/* The entry points below are host functions; device code that only needs the constants and structs above
 * (the library's own NVRTC build of its kernels) skips them. */
#ifndef __CUDACC_RTC__

/* ---- lifetime -------------------------------------------------------------------------- */

/* Build the constant tables of a robot.  Stands in for UrdfForwardKinematic._build
 * (kinematics.py:163-210); URDF parsing stays on the host side of the ABI.
 *   T_const [F][16]  row-major 4x4 joint-origin transforms (kinematics.py:200-203)
 *   axis    [F][3]   joint axes (kinematics.py:204); revolute axes must be unit length
 *   jtype   [F]      RMP2_JOINT_*
 *   parent  [F]      index of the parent frame, -1 for children of the base link; parent < own
 *   qidx    [F]      column of q driving this joint, -1 if none (kinematics.py:197)           */
int rmp2_robot_create(const float* T_const, const float* axis, const int8_t* jtype,
                      const int32_t* parent, const int32_t* qidx, int32_t F, int32_t n,
                      rmp2_robot** out);
void rmp2_robot_destroy(rmp2_robot* robot);

/* Compile an ordered leaf list into kernel tables.  Stands in for RmpCore.add_rmp
 * (rmp.py:127-128); leaf order = dict insertion order.  The robot may be destroyed afterwards. */
int rmp2_tree_create(const rmp2_robot* robot, const rmp2_leaf_desc* leaves, int32_t n_leaves,
                     rmp2_tree** out);
void rmp2_tree_destroy(rmp2_tree* tree);
/* Overwrite params/vec of leaf `index` (same type/space/frame), e.g. `target_rmp.goal = ...`
 * (experiments/franka_panda/06_cluttered_environment.py:142). */
int rmp2_tree_update_leaf(rmp2_tree* tree, int32_t index, const rmp2_leaf_desc* leaf);

/* ---- the hot path ------------------------------------------------------------------------ */

/* qdd = pinv(sum_l J_l^T M_l J_l) * sum_l J_l^T M_l (xdd_l - Jdot_l qd) for B environments.
 * Stands in for RmpCore.evaluate (rmp.py:133-155).  All io pointers are device memory.
 * Launches up to four kernels on `stream` (frames -> spheres -> step -> resolve fallback; five with
 * RMP2_OPT_SPLIT_RESOLVE).  The
 * tree handle owns scratch buffers (per-(environment, obstacle leaf) records of 40 B, the combined metric, the
 * work list of the fallback resolve; at most 2^20 environments at a time).  They are sized by rmp2_tree_reserve,
 * or grown on demand with stream-ordered allocation on `stream` (no device-wide stall; refused while `stream`
 * is being captured); therefore the steps of ONE tree must be issued from one thread on one stream at a time. */
int rmp2_step(const rmp2_tree* tree, const rmp2_step_io* io, void* stream);

/* Size the tree's scratch for steps of up to B environments with up to n_spheres spheres each, so that no
 * later rmp2_step / rmp2_rollout allocates (required before capturing steps into a CUDA graph). */
int rmp2_tree_reserve(rmp2_tree* tree, int64_t B, int32_t n_spheres, void* stream);

/* Same step with HOST buffers: every io pointer is host memory (pinned for full overlap);
 * inputs are copied to device staging owned by the tree, the step runs, qdd is copied back,
 * chunked and pipelined on internal streams.  Returns after qdd is complete. */
int rmp2_step_host(rmp2_tree* tree, const rmp2_step_io* io);

/* Closed-loop rollout: `n_steps` simulation steps of explicit Euler
 * (qd += qdd*dt; q += qd*dt), re-evaluating the tree every `control_every` steps (command held in between)
 * (the 100 Hz / 10 Hz loop of experiments/franka_panda/05_obstacle_avoidance.py:92-97 with
 * simulation.step replaced by an integrator).  q and qd are updated in place; io->qdd receives
 * the last command. */
int rmp2_rollout(const rmp2_tree* tree, const rmp2_step_io* io, float* q_inout, float* qd_inout,
                 float dt, int32_t n_steps, int32_t control_every, void* stream);

/* ---- stage-level entry points (same kernels' device functions, for the Python API) ------- */

/* x = vec(T_frame(q)) [B][16], xd = J qd [B][16], J [B][16][n], c = Jdot qd [B][16].
 * Stands in for UrdfForwardKinematic.forward / .differentiate (kinematics.py:212-270).
 * qd, xd, J, c may be NULL (forward only). */
int rmp2_fk(const rmp2_robot* robot, int32_t frame, int64_t B, const float* q, const float* qd,
            float* x, float* xd, float* J, float* c, void* stream);

/* xdd [K][m], M [K][m][m] of one leaf at task-space points x, xd [K][m].
 * Stands in for RiemannianMotionPolicy.evaluate (rmp.py:202-206, rmp2.py:25-29).
 * aux: only for RMP2_LEAF_COLLISION_AVOIDANCE, [K][4] = (distance, normal_vec xyz) -- the data the
 * reference leaf holds as self.d / self.vec (rmp.py:269-270); NULL otherwise. */
int rmp2_leaf_evaluate(const rmp2_leaf_desc* leaf, int32_t m, int64_t K, const float* x,
                       const float* xd, const float* aux, float* xdd, float* M, void* stream);

/* On-GPU obstacle feed with primitive geometry on both sides.  Stands in for Simulation.calculate_distances
 * (simulation.py:462-484: p.getClosestPoints between a link's collision shape and an obstacle): for every
 * (environment, listed frame, obstacle) it emits one closest-point pair in the reference's distance_data layout.
 *   frames        [n_frames]            frame indices (order of rmp2_robot_create)
 *   link_capsules [n_frames][8] HOST    the link geometry riding on each listed frame, one capsule in FRAME coordinates
 *                                       (ax, ay, az, bx, by, bz, radius, 0); NULL = the frame origin as control point
 *   spheres       [B][n_spheres][4]     (cx, cy, cz, radius); 16-byte aligned
 *   capsules      [B][n_capsules][8]    (ax, ay, az, bx, by, bz, radius, 0) -- the experiments' cylinders; 16-byte aligned
 *   pairs    [B][n_frames*K][8] out     K = n_spheres + n_capsules; (pos_on_link, pos_on_obstacle, 0, 0), both on the
 *                                       surfaces, i.e. the rows rmp2_step_io.pairs expects for FRAME_DISTANCE_PAIRS leaves
 *   aux      [B][n_frames*K][4] out     (surface distance, normal_vec from obstacle to link), may be NULL */
int rmp2_obstacle_feed(const rmp2_robot* robot, const int32_t* frames, const float* link_capsules, int32_t n_frames,
                       int64_t B, const float* q, const float* spheres, int32_t n_spheres, const float* capsules,
                       int32_t n_capsules, float* pairs, float* aux, void* stream);

/* x = pinv(M) f for B independent n x n systems, M [B][n][n], f [B][n], x [B][n] (device, row-major), with
 * tf.linalg.pinv's default cutoff 10 n eps32 sigma_max.  Stands in for the last two lines of RmpCore.evaluate
 * (rmp.py:153-154) on its own.  pivot != 0: the variant for possibly rank-deficient M (pivoted QR +
 * rank-revealing direct solve); 0: the variant for well-conditioned M.  mode 0: as in the step (direct solve
 * where the spectrum is provably clear of the cutoff, one-sided Jacobi SVD otherwise); mode 1: Jacobi SVD for
 * every system (cross-check of the two solvers). */
int rmp2_pinv_solve(int32_t n, int64_t B, const float* M, const float* f, float* x, int32_t pivot,
                    int32_t mode, void* stream);

/* ---- introspection ----------------------------------------------------------------------- */
const char* rmp2_last_error(void);
const char* rmp2_version(void);
/* kernels launched by this library in this process so far (bench.py's gpu_launches) */
int64_t rmp2_launch_count(void);
/* registers/thread, dynamic shared memory, max resident blocks/SM and block size of one of the
 * kernels a step launches: which = 0 frames (chain -> frame records), 1 spheres (the obstacle
 * pair loop; n_spheres selects the staging layout), 2 step with the direct resolve fused (small batches),
 * 3 step without it and 4 the stand-alone direct-resolve kernel (large batches), 5 the fallback resolve
 * (Jacobi SVD for the environments the direct solve handed over). */
int rmp2_tree_kernel_info(const rmp2_tree* tree, int32_t which, int32_t n_spheres, int32_t* regs,
                          int32_t* smem_bytes, int32_t* blocks_per_sm, int32_t* block_threads);
/* Options of a tree.  RMP2_OPT_EARLY_OUT (default 1): the obstacle kernel evaluates only the
 * (frame, sphere) pairs within the leaf's metric_modulation_radius; the others contribute exactly
 * zero in the reference as well (rmp2.py:194), so results are unchanged.  Set to 0 to force every
 * pair through the full arithmetic (used for the roofline measurement). */
#define RMP2_OPT_EARLY_OUT 0
/* RMP2_OPT_TMA (default 1): stage sphere rows through shared memory with TMA bulk copies (0: LDG.128).
 * RMP2_OPT_SPLIT_RESOLVE (default -1 = 0; 0 / 1): run the direct resolve inside the step kernel (measured faster
 * or equal on every configuration) or as its own kernel behind the (M, f) scratch.  RMP2_OPT_BLOCK_THREADS (default 0 = by batch size; 32 / 64 / 128).
 * RMP2_OPT_CHUNK_ENVS (default 0 = 2^20): environments per internal chunk of a step (bounds the scratch). */
#define RMP2_OPT_TMA 1
#define RMP2_OPT_SPLIT_RESOLVE 2
#define RMP2_OPT_BLOCK_THREADS 3
#define RMP2_OPT_CHUNK_ENVS 4
/* RMP2_OPT_MERGE_COINCIDENT (default 1): ObstacleAvoidance leaves on the sphere path whose parameters are equal and
 * whose frame origins coincide for every q (a frame with zero constant translation and a revolute or fixed joint sits
 * on its parent's origin; Panda: joint2 on joint1, joint6 on joint5) have the same x, xd, c and the same Jacobian up to a
 * zero column (the reference differentiates the distance through the frame origin only, taskmap.py:124-128), hence the
 * same pulled-back (M, f).  One leaf of the group runs the pair loop and the pullback, its sums are multiplied by the
 * size of the group.  Results equal the unmerged tree's up to rounding (M1 + M1 = 2 M1 is exact; the sum over leaves
 * is formed in another order).  A group whose control point cannot move at all (its origin frame hangs on the base
 * through fixed joints only and is not prismatic; Panda: joint1, joint2) has J = 0 and a pulled-back (M, f) of exactly
 * zero in the reference too: it runs no pair loop.  0: every leaf on its own (used for the roofline measurement).  Changing the option
 * recompiles the tree's tables (and its specialised kernels, when loaded). */
#define RMP2_OPT_MERGE_COINCIDENT 5
int rmp2_tree_set_option(rmp2_tree* tree, int32_t option, int32_t value);
/* Sphere-path ObstacleAvoidance leaves of the tree (*n_leaves) and the pair loops that run for them per environment
 * (*n_slots = n_leaves minus the leaves merged under RMP2_OPT_MERGE_COINCIDENT).  Either pointer may be NULL. */
int rmp2_tree_obstacle_slots(const rmp2_tree* tree, int32_t* n_leaves, int32_t* n_slots);

/* Tree-specialised kernels.  The frames and step kernels interpret the tree's tables at run time; for a
 * large batch that shares one tree that interpretation is pure overhead.  rmp2_tree_specialize rebuilds
 * the two kernels for THIS tree with NVRTC (tables as a compile-time constant, frame and leaf loops
 * unrolled; a few seconds, once) and uses them for every later rmp2_step / rmp2_rollout of the tree.
 * Results agree with the generic kernels to rounding (same source, different constant folding).
 * rmp2_tree_update_leaf rebuilds the specialisation for the new values (the tables are compile-time constants
 * of the kernels); if that fails the generic kernels take over.
 * flags: RMP2_SPECIALIZE_COMPILE_ONLY = run NVRTC but do not load (needs no GPU; build checks).
 * Fails with RMP2_ERR_UNSUPPORTED when NVRTC (libnvrtc.so.12) cannot be loaded. */
#define RMP2_SPECIALIZE_COMPILE_ONLY 1
int rmp2_tree_specialize(rmp2_tree* tree, int32_t flags);
/* 1 when specialised kernels are loaded for the tree; *compile_seconds (may be NULL) = NVRTC time. */
int rmp2_tree_is_specialized(const rmp2_tree* tree, double* compile_seconds);

/* Per-kernel device timing with CUDA events on the launching stream (bench.py's roofline leg).
 * rmp2_tree_profile_read waits for the recorded launches, returns the accumulated milliseconds and
 * launch counts of {frames, spheres, step, resolve, resolve fallback} since the last read, and resets them. */
#define RMP2_PROFILE_KERNELS 5
int rmp2_tree_profile(rmp2_tree* tree, int32_t enable);
int rmp2_tree_profile_read(rmp2_tree* tree, double* ms /*[5]*/, int64_t* launches /*[5]*/);

#endif /* __CUDACC_RTC__ */

#ifdef __cplusplus
}
#endif
#endif /* RMP2_B200_H_ */
