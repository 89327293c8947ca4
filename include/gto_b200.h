/*
 * gto_b200.h -- C-ABI of libgto_b200.so: batched grasp-trajectory optimisation on NVIDIA B200 (sm_100a).
 *
 * This is the drop-in boundary for the hot path of IRVLUTD/GraspTrajOpt (reference commit 4703ba2).  In the
 * reference the path sits behind the Python object protocol of optas.CasADiSolver (optas/solver.py:323-421),
 * driven only from gto.GTOPlanner (gto/gto_planner.py:142,160-176,222-239) and gto.IKSolver
 * (gto/ik_solver.py:75-90); beneath it is casadi.nlpsol("solver","ipopt",...) (optas/solver.py:384,397).
 * The reference has no FFI of its own -- a maintainer binds these entry points with ctypes (see
 * INTEGRATION.md; grasptrajopt_b200/capi.py is that binding).
 *
 * Conventions
 *   - plain C, opaque handle, caller-owned host buffers, library-owned device buffers;
 *   - every entry point returns 0 (GTO_OK) or a negative code; no C++ exception crosses the ABI;
 *     gto_last_error() gives the message of the last failure on that context;
 *   - one context per (process, device); a context is not thread-safe; calls are synchronous on an internal stream;
 *   - arrays are C-order; joint vectors `q` have `ndof` entries in URDF actuated-joint order
 *     (optas/models.py:349-354); `nopt` optimised joints are a subset (opt_qidx), the rest are parameters.
 */
#ifndef GTO_B200_H
#define GTO_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GTO_ABI_VERSION 5

/* error codes */
#define GTO_OK 0
#define GTO_ERR_INVALID -1    /* bad argument / inconsistent sizes */
#define GTO_ERR_CUDA -2       /* CUDA runtime or driver error */
#define GTO_ERR_NO_DEVICE -3  /* no CUDA device / wrong architecture */
#define GTO_ERR_STATE -4      /* call order (robot or field not set, nothing uploaded, ...) */
#define GTO_ERR_NOMEM -5

/* per-problem status */
#define GTO_STATUS_CONVERGED 0
#define GTO_STATUS_MAX_ITER 1
#define GTO_STATUS_NAN 2
#define GTO_STATUS_STALLED 3  /* damping hit lambda_max without an acceptable step */
#define GTO_STATUS_SLOW 4     /* |dq| fell below tol_step only under heavy damping (> lambda_conv): the iterate rests on a
                                 gradient jump (cell face) of the piecewise-trilinear field, where no descent step longer than
                                 tol_step is accepted; the trajectory is returned but not counted as converged */

/* joint types in the robot table */
#define GTO_JOINT_REVOLUTE 1
#define GTO_JOINT_PRISMATIC 2

#define GTO_MAX_OPT 16  /* optimised joints */
#define GTO_MAX_MOV 32  /* movable joints kept in the table */
#define GTO_MAX_LINKS 32

/* flags for gto_batch_in.flags */
#define GTO_FLAG_NO_JROWS 1u      /* do not materialise the Jacobian rows in HBM (assembly still fused) */
#define GTO_FLAG_OBS_LINEAR 2u    /* obstacle term w_obs * sum c(W) instead of w_obs * sum c(W)^2: the reference IK solver's term
                                     (gto/ik_solver.py:69).  Value w*c, gradient w*dc/dq, no Gauss-Newton curvature; the row block
                                     still holds sqrt(w)*dc/dq | sqrt(w)*c */
#define GTO_FLAG_NO_CULL 32u      /* every link is treated as touching a non-zero node of the field (A-B check of the culling) */

typedef struct gto_ctx gto_ctx;

/*
 * Flattened kinematic tree (replaces the symbolic chain FK the reference re-traces for every (link, knot):
 * optas/models.py:826-868, gto/gto_models.py:83-101).  Only movable joints are listed, parents first; runs of
 * fixed joints are folded into `mov_origin`.
 *   T_j = T_parent(j) * mov_origin[j] * motion_j(q[mov_qidx[j]])
 *   visual frame of collision link l = T_{link_mov[l]} * link_tf[l]          (link_mov < 0: root)
 *   gripper link frame             = T_{grip_mov} * grip_tf               (plain link frame, gto_planner.py:79-82)
 * All 3x4 matrices are row-major [R|t].
 */
typedef struct gto_robot_desc {
  int32_t ndof;            /* length of every q vector */
  int32_t nopt;            /* optimised joints (<= GTO_MAX_OPT) */
  const int32_t* opt_qidx; /* [nopt] index into q */
  const double* lo;        /* [nopt] position limits (optas/builder.py:472-510) */
  const double* hi;        /* [nopt] */
  int32_t nmov;
  const int32_t* mov_parent;  /* [nmov] index of parent movable joint or -1 */
  const int32_t* mov_type;    /* [nmov] GTO_JOINT_* */
  const double* mov_origin;   /* [nmov][12] */
  const double* mov_axis;     /* [nmov][3] unit axis in the joint frame */
  const int32_t* mov_qidx;    /* [nmov] index into q */
  const int32_t* mov_opt;     /* [nmov] index among optimised joints or -1 */
  int32_t nlinks;
  const int32_t* link_mov;       /* [nlinks] */
  const double* link_tf;         /* [nlinks][12] */
  const int32_t* link_pt_start;  /* [nlinks] */
  const int32_t* link_pt_count;  /* [nlinks] */
  const uint32_t* link_optmask;  /* [nlinks] bit k set: optimised joint k moves this link */
  int32_t npoints;
  const float* points;    /* [npoints][3] surface points in their link's visual frame (gto_models.py:62-80) */
  int32_t grip_mov;       /* frame of link_gripper */
  const double* grip_tf;  /* [12] */
  int32_t grip_pt_start;  /* gripper point set = points[grip_pt_start .. +grip_pt_count) (gto_planner.py:37) */
  int32_t grip_pt_count;
  uint32_t grip_optmask;
} gto_robot_desc;

/* Solver options; gto_default_options() fills the values used by the oracle (oracle/gto_oracle.py SolverOptions). */
typedef struct gto_options {
  int32_t max_iter;    /* 100 = reference max_iter (gto/gto_planner.py:141) */
  double tol_step;     /* |dq|_inf of an accepted step           (1e-6) */
  double tol_grad;     /* |projected gradient|_inf               (1e-6) */
  double lambda0;      /* initial Levenberg-Marquardt damping    (1e-3) */
  double lambda_min;   /* 1e-9 */
  double lambda_max;   /* 1e9  */
  double eta;          /* step acceptance ratio                  (1e-4) */
  double noise_rel;    /* cost reductions below noise_rel*point-cost count as fp32 noise (1e-6) */
  double bound_eps;    /* 1e-12 */
  int32_t check_every; /* host polls the device convergence counter every this many iterations (4) */
  double ftol;         /* GTO_STATUS_SLOW: accepted step with cost reduction <= ftol*f ...          (1e-6) */
  double lambda_slow;  /* ... while the damping that produced it was >= lambda_slow  (1e30 = the test is off: iterates resting
                          on a kink of the trilinear field are iterated until |dq| <= tol_step like any other) */
  int32_t slow_window; /* GTO_STATUS_SLOW as well when the cost fell by <= slow_ftol*f over the last slow_window
                          iterations (default 0 = disabled: it also fires during slow but healthy linear convergence; at most 16) */
  double slow_ftol;    /* (1e-3) */
  int32_t as_rounds;   /* active-set rounds per step (1): free variables that the step pushes beyond a joint limit are moved
                          exactly onto it and the other variables are re-solved; 0 = plain clipping of the step */
  double lambda_reject; /* a rejected step raises the damping to at least this value (1e-4): from lambda_min ~ 1e-9 the
                           doubling rule alone needs ~7 rejections before the damping changes the step at all */
  double lambda_conv;   /* |dq| <= tol_step counts as GTO_STATUS_CONVERGED only when the damping that produced the step was
                           <= lambda_conv (1e-2), i.e. the step was the Gauss-Newton step to within 1 %; otherwise GTO_STATUS_SLOW */
  int32_t bundle;       /* pieces of the gradient bundle (3; 0 = plain Levenberg-Marquardt, at most GTO_BUNDLE_MAX).  The trilinear
                           field makes the objective piecewise smooth: its gradient jumps at cell faces, and a minimiser usually
                           lies ON such a kink, where every one-sided quadratic model predicts a descent that the other side
                           takes back.  The solver therefore keeps the (cost, gradient) of the last `bundle` points it evaluated
                           but does not stand on (rejected trial points, iterates it left) as cutting planes
                           piece_k(s) = (f_k - f)/2 + g_k.(s - (y_k - x)) and minimises  max_k piece_k(s) + s'(H + lambda D)s/2 :
                           one block-tridiagonal factorisation with bundle+1 right-hand sides and a (bundle+1)-variable dual QP.
                           The step tends to zero under light damping at a kink minimiser, so tol_step certifies it. */
  double bundle_radius; /* a piece whose point is further than this from the standing point (|.|_inf, 3e-3 rad) is ignored and is
                           the first to be replaced: far from the iterate the planes say nothing about the kinks around it, and
                           using them makes the path depend on 1e-9 perturbations of the cost (measured: with no radius 11 of
                           244 C2 problems end in another local minimum under 1e-9 relative noise, with 3e-3 none do) */
} gto_options;
#define GTO_BUNDLE_MAX 4

/*
 * One batch of B independent (seed x grasp) problems == B reference plan() calls (gto/gto_planner.py:42-182).
 * Residual blocks per problem (f = sum r^2):
 *   goal      sqrt(w_goal) * ( W_g(Q_{T-1}, x_k) - goal_tf[b][0] x_k )                     3*Pg rows
 *   stand-off sqrt(w_goal) * ( W_g(Q_{T+standoff_offset}, x_k) - goal_tf[b][1] x_k )        3*Pg rows  (use_standoff)
 *   obstacle  sqrt(w_obs)  * c( W_link(Q_t, x_i) + base_position[b] ),  c = trilinear field  T*P rows  (collision_avoidance)
 *             knots t <  T+standoff_offset read field slot field_all[b], later knots field_obs[b] (gto_planner.py:117-131)
 *   velocity  sqrt(w_vel)  * (Q_{t+1}-Q_t)/dt                                                analytic, never materialised
 * Constraints (gto_planner.py:59-72,138): optimised rows of knots 0 and 1 equal qc; lo <= Q_t <= hi.
 */
typedef struct gto_batch_in {
  int32_t B;
  int32_t T;
  double dt;
  const double* qc;             /* [B][ndof] current configuration */
  const double* q_seed;         /* [B][T][ndof] initial trajectory; parameter-joint entries are kept as given */
  const double* goal_tf;        /* [B][2][12]: RT.G and RT.S.G as row-major 3x4 (gto_planner.py:93-100) */
  const double* base_position;  /* [B][3] */
  const int32_t* field_all;     /* [B] field slot for knots <  T+standoff_offset, -1 = zero field (plan(), Q4) */
  const int32_t* field_obs;     /* [B] field slot for knots >= T+standoff_offset, -1 = zero field */
  int32_t standoff_offset;      /* -10 */
  int32_t use_standoff;
  int32_t collision_avoidance;
  double w_goal, w_obs, w_vel;  /* 1, 10, 0.01 */
  uint32_t flags;
} gto_batch_in;

typedef struct gto_batch_out {
  double* Q;        /* [B][T][ndof] */
  double* dQ;       /* [B][T-1][ndof]  optimised rows = diff(Q)/dt, parameter rows = 0 (optas/solver.py:126-159) */
  double* cost;     /* [B] objective f at Q (goal + w_obs*obstacle + w_vel*velocity, Q10) */
  int32_t* iters;   /* [B] */
  int32_t* status;  /* [B] GTO_STATUS_* */
} gto_batch_out;

/* Output of one linearisation (parity tests): dense Jacobian rows and the per-knot Gauss-Newton blocks.
 * Row layout per problem: [rows][nopt+1] = [J (nopt) | r]; rows = obstacle [t][point] (T*P, if collision_avoidance),
 * then goal [axis][k] (3*Pg), then stand-off [axis][k] (3*Pg, if use_standoff).  Any pointer may be NULL. */
typedef struct gto_eval_out {
  float* rows;   /* [B][nrows][nopt+1] */
  float* H;      /* [B][T][nopt][nopt]  sum_rows j j^T per knot (error-compensated TF32 tensor-core contraction) */
  double* g;     /* [B][T][nopt]        sum_rows j r   per knot (float32 products, float64 accumulation) */
  double* cost;  /* [B][T]              sum_rows r^2   per knot (float64 accumulation) */
} gto_eval_out;

/* Timings measured with CUDA events on the library's stream. */
typedef struct gto_profile {
  double solve_ms;           /* device time of the last gto_solve_* (first kernel .. last kernel) */
  double linearize_ms;       /* summed duration of the linearise launches inside it */
  double step_ms;            /* summed duration of the LM-step launches */
  int32_t linearize_launches;
  int32_t step_launches;
  int32_t iterations;        /* outer iterations executed (max over problems) */
  int64_t knot_items;        /* (problem,knot) work items processed by the linearise launches (exact, from the device) */
  int64_t jrow_bytes;        /* bytes of Jacobian rows written to HBM (exact) */
  int64_t problem_iterations;            /* sum over linearise launches of the problems still active */
  int32_t linearize_launches_with_work;  /* launches that had at least one active problem */
  double h2d_ms, d2h_ms;     /* last upload / download */
  int64_t h2d_bytes, d2h_bytes;
  int64_t links_tested;      /* (problem, knot, link) triples whose node box was tested against the field's summed-volume table */
  int64_t links_active;      /* ... of which touched a non-zero node and went through the point kernel (the rest: zero rows) */
  int64_t kernel_launches;   /* kernels launched by the last gto_solve_* */
} gto_profile;

int gto_abi_version(void);
int gto_create(gto_ctx** ctx, int device);
void gto_destroy(gto_ctx* ctx);
const char* gto_last_error(gto_ctx* ctx);
void gto_default_options(gto_options* opts);

/* Upload the robot table and point sets (replaces GTORobotModel.setup_fk_functions + surface_pc_map). */
int gto_set_robot(gto_ctx* ctx, const gto_robot_desc* robot);

/* Upload one voxel cost field into `slot` (0 <= slot < 4096): cost[dims[0]][dims[1]][dims[2]] float32, nodes at
 * origin + k*pitch (gto/gto_models.py:155-187).  Replaces passing sdf_cost_all / sdf_cost_obstacle as NLP parameters
 * on every call (gto/gto_planner.py:53-54,228-235). */
int gto_set_field(gto_ctx* ctx, int slot, const float* cost, const int32_t dims[3], const double origin[3], double pitch);

/* Solve a batch: H2D of the per-problem inputs, LM iterations to convergence on the device, D2H of the results. */
int gto_solve_batch(gto_ctx* ctx, const gto_batch_in* in, const gto_options* opts, gto_batch_out* out);

/* The same in three steps, so a caller can keep inputs resident in HBM and time the solve alone. */
int gto_upload_batch(gto_ctx* ctx, const gto_batch_in* in);
int gto_solve_resident(gto_ctx* ctx, const gto_options* opts);
int gto_download_batch(gto_ctx* ctx, gto_batch_out* out);

/* Device pointer to the packed float32 result of the last solve, [B][nopt*T + 2] = (optimised rows of Q
 * knot-major, cost, status) -- the payload of the multi-GPU all-gather. */
int gto_result_device_ptr(gto_ctx* ctx, void** ptr, int64_t* nfloats_per_problem);

/* One linearisation at in->q_seed (no projection, no iteration). */
int gto_eval_batch(gto_ctx* ctx, const gto_batch_in* in, gto_eval_out* out);

int gto_get_profile(gto_ctx* ctx, gto_profile* prof);

/* Run-time tuning knobs of a context (defaults in parentheses; the environment variable of the same upper-case name with the
 * prefix GTO_ is read once in gto_create):
 *   "jrows_budget_mb" (24576)  size of the Jacobian-row buffer; larger batches are solved in chunks that fit
 *   "pdl" (1)                  programmatic dependent launch of the solver kernels
 *   "launch_events" (0)        CUDA events between the launches instead of in-kernel time stamps (profile cross-check)
 *   "step_fk" (0)              the step kernel also writes the item records once at most this many problems are active
 *   "cull_nslot" (4), "cons_warps" (0 = automatic), "slot_floats" (0 = automatic)   shared-memory ring of k_linearize_cull
 *   "step_dbg" (0)             print clock64() phase times of CTA 0 of the step / FK kernels to stderr
 *   "fused" (0)                1: k_solve_fused, one persistent CTA per problem runs the whole solver loop (identical results)
 * Unknown keys return GTO_ERR_INVALID. */
int gto_configure(gto_ctx* ctx, const char* key, double value);

/* Value-only pass: sum over knots and points of the nearest-node cost along given plans -- the reference's seed
 * ranking GTORobotModel.compute_plan_cost (gto/gto_models.py:204-215).  plans [n][T][ndof]; cost[n], dist[n]. */
int gto_plan_cost(gto_ctx* ctx, int32_t n, int32_t T, const double* plans, int32_t field_slot, const double base_position[3],
                  double* cost, double* dist);

/*
 * Scene side: the reference's DepthPointCloud (mesh_to_sdf/depth_point_cloud.py:9-91,127-142).
 * gto_cloud_set uploads the world-frame point cloud of a depth image (what the reference puts into a scikit-learn KD-tree, :20-25);
 * gto_cloud_query returns, for N query points, the distance to the nearest cloud point, negative where the query is hidden behind
 * the visible surface (is_outside, :127-142: projection with the intrinsics K into the depth image through cam_inv, the inverse
 * camera pose, row-major 4x4) -- mode 0, get_sdf :57-62 -- or the CHOMP-style cost of that distance -- mode 1, get_sdf_cost :65-91 --
 * or the visibility test alone, 1.0 = visible / outside, 0.0 = hidden -- mode 2, is_outside :127-142 (needs no cloud).
 * kernel_ms (may be NULL) receives the device time of the query kernel.
 * gto_cloud_backproject turns a depth image into the world-frame cloud (DepthPointCloud.__init__ / backproject_camera, :15-19,32-52):
 * pixels with 0 < depth < threshold and target_mask == 0 (target_mask may be NULL); points[H*W][3] receives every pixel's point,
 * valid[H*W] which of them pass the test (compact in row-major pixel order to get the reference's `points`).
 */
int gto_cloud_backproject(gto_ctx* ctx, const float* depth, const uint8_t* target_mask, int32_t H, int32_t W, const double Kinv[9],
                          const double cam_pose[16], double threshold, double* points, uint8_t* valid);
int gto_cloud_set(gto_ctx* ctx, const double* points, int64_t M);
int gto_cloud_query(gto_ctx* ctx, const double* query, int64_t N, const float* depth, int32_t H, int32_t W, const double K[9],
                    const double cam_inv[16], int32_t mode, double epsilon, double w_inside, float* out, double* kernel_ms);

/*
 * Mobile-base placement: the reference's BasePlanner (gto/base_planner.py:35-168).  One problem ==
 * one BasePlanner.plan_goalset(qc, RTs) call: find the planar base motion y = (x, y, theta) and one arm configuration per goal
 * minimising  w_effort |y|^2 + sum_i sum_k | F_gripper(q_i) x_k - T_b(y) RT_i G x_k |^2  subject to -pi <= theta <= pi and the
 * joint limits (:44-89), every arm seeded with qc and y = 0 (:101-118).  B problems (the random grasp subsets of the reference's
 * rejection loop, examples/pybullet_gto_planning_mobile.py:187-201) are solved to convergence in one kernel launch; `collision`
 * is the occupancy-grid count of the robot's surface points at qc seen from the new base (:150-165; grid layout and indexing of
 * GTORobotModel.setup_occupancy_grid / points_to_offsets_occupancy_numpy, gto/gto_models.py:219-271).
 */
typedef struct gto_base_in {
  int32_t B;               /* problems */
  int32_t n_goals;         /* goals per problem, 1..32 (builder T = goal_size, base_planner.py:38) */
  const double* qc;        /* [ndof] */
  const double* goal_tf;   /* [B][n_goals][12]: RT_i . G as row-major 3x4, current base frame */
  double w_effort;         /* base_effort_weight (0.01) */
  const float* occupancy;  /* [occ_dims[0]][occ_dims[1]] cell values, or NULL (collision is then 0) */
  int32_t occ_dims[2];
  double occ_origin[2];
  double occ_resolution;
} gto_base_in;

typedef struct gto_base_out {
  double* Q;          /* [B][n_goals][ndof] arm configuration per goal (parameter joints = qc) */
  double* y;          /* [B][3] (x, y, theta): old base in new base (base_planner.py:49-52) */
  double* cost;       /* [B] objective, may be NULL */
  double* collision;  /* [B] may be NULL */
  int32_t* iters;     /* [B] may be NULL */
  int32_t* status;    /* [B] GTO_STATUS_*, may be NULL */
} gto_base_out;

/* opts: max_iter, tol_step, tol_grad, lambda0/min/max, eta, bound_eps are used (NULL = defaults). */
int gto_base_place(gto_ctx* ctx, const gto_base_in* in, const gto_options* opts, gto_base_out* out, double* kernel_ms);

#ifdef __cplusplus
}
#endif
#endif /* GTO_B200_H */
