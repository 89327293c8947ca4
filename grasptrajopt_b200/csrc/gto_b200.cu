// gto_b200.cu -- CUDA kernels (sm_100a) + C-ABI of libgto_b200.so.  See include/gto_b200.h.
//
// Hot path of IRVLUTD/GraspTrajOpt re-designed for B200 (reference citations are file:line in the reference tree):
//   k_linearize   ONE fused kernel per Gauss-Newton iteration: chain FK of every link (optas/models.py:826-868,
//                 gto/gto_models.py:83-101) -> world points (gto/gto_planner.py:111-128) -> trilinear SDF value + gradient
//                 from a shared-memory brick staged by TMA (replaces the nearest-node gather gto/gto_models.py:174-187 and
//                 gto/sdf_callback.py) -> obstacle / goal / stand-off residuals (gto_planner.py:86-131) -> chained analytic
//                 Jacobian rows (geometric Jacobian, optas/models.py:1203-1268) -> coalesced fp32 row stores to HBM ->
//                 per-knot Gauss-Newton blocks: J^T J on the tensor cores (mma.sync m16n8k8 TF32; an 8x8..16x16 output cannot
//                 fill a tcgen05 M=64 tile), J^T r and the cost in fp32 with warp-shuffle reduction.
//   k_step        per problem: step acceptance (Levenberg-Marquardt, Nielsen damping), projected bound handling
//                 (optas/builder.py:472-510), block-tridiagonal factor/solve in fp64 (the velocity term gto_planner.py:134-135
//                 couples neighbouring knots), next trial point.  Replaces IPOPT (optas/solver.py:384-400).
//   k_init / k_finalize   constraint elimination (gto_planner.py:59-72) and result unpacking (optas/solver.py:126-159).
// There is no CPU fallback: every entry point fails with GTO_ERR_NO_DEVICE without a CUDA device.

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <math.h>
#include <math_constants.h>
#include <string>
#include <vector>
#include <thread>
#include <functional>
#include <algorithm>

#include "../../include/gto_b200.h"

#define MAX_FIELDS 4096
#define MAX_CHUNKS 1024
#define CULL_NAXC 7      // per-axis TMA box sizes 8,12,...,32 (one tensor map per combination and field)
#define CULL_MAX_CONS 8  // consumer warps of k_linearize_cull

// ------------------------------------------------------------------------------------------------------------------
// device-side tables
// ------------------------------------------------------------------------------------------------------------------
struct RobotDev {
  int ndof, nopt, nmov, nlinks, npoints;
  int opt_qidx[GTO_MAX_OPT];
  int opt_mov[GTO_MAX_OPT];
  int mov_parent[GTO_MAX_MOV], mov_type[GTO_MAX_MOV], mov_qidx[GTO_MAX_MOV], mov_opt[GTO_MAX_MOV];
  float mov_origin[GTO_MAX_MOV][12];
  float mov_axis[GTO_MAX_MOV][4];
  double mov_origin_d[GTO_MAX_MOV][12];
  double mov_axis_d[GTO_MAX_MOV][4];
  double link_tf_d[GTO_MAX_LINKS][12];
  double grip_tf_d[12];
  int link_mov[GTO_MAX_LINKS];
  float link_tf[GTO_MAX_LINKS][12];
  int link_pt_start[GTO_MAX_LINKS], link_pt_count[GTO_MAX_LINKS];
  unsigned link_optmask[GTO_MAX_LINKS];
  float link_center[GTO_MAX_LINKS][4];  // AABB of the link's points in its visual frame
  float link_half[GTO_MAX_LINKS][4];
  int link_chunk0[GTO_MAX_LINKS + 1];   // chunk range per link
  int grip_mov;
  float grip_tf[12];
  int grip_pt_start, grip_pt_count;
  unsigned grip_optmask;
  int nchunks;
  double lo[GTO_MAX_OPT], hi[GTO_MAX_OPT];
};

struct FieldDev {
  const float* data;  // [nx][ny][nzp]
  int nx, ny, nz, nzp;
  float ox, oy, oz, inv_pitch;
  double org_d[3], inv_pitch_d;  // float64 copies for the voxel coordinate of a point (k_item_fk)
  const CUtensorMap* maps2;  // [7*7*7] TMA tile maps with per-axis box sizes 8,12,...,32; NULL if unavailable
  const unsigned* svt;       // [(nx+1)][(ny+1)][(nz+1)] summed-volume table of the non-zero nodes (culling test); NULL if unavailable
};

struct LinParams {
  const RobotDev* robot;
  const int* chunk_start;  // [nchunks] first point of each 32-point chunk
  const int* chunk_count;  // [nchunks]
  const float* pts3;       // [3][npad] surface points x | y | z (one bulk copy into shared memory per CTA)
  int npad;                // npoints rounded up to a multiple of 4
  const float* px;
  const float* py;
  const float* pz;
  const double* q;         // [B][T][ndof] configuration at which to linearise
  const double* goal_tf;   // [B][2][12]
  const float* base;       // [B][4]
  const int* field_ids;    // [B][2]
  const FieldDev* fields;
  const int* active;       // problem ids (NULL: identity)
  const int* nactive;      // device counter (NULL: use nproblems)
  int nproblems;
  int b0;                  // first problem of the chunk (rows buffer is indexed by b - b0)
  const int* bufsel;       // [B] accepted buffer per problem; output goes to the other one (NULL: buffer 0)
  float* H;                // [2][Bcap][T][nopt*nopt]
  double* g;               // [2][Bcap][T][nopt]   J^T r (float64 accumulation)
  double* costp;           // [2][Bcap][T]         sum r^2
  long long buf_stride_H, buf_stride_g, buf_stride_c;     // accepted / trial buffer
  float* rows;             // [Bchunk][nrows][RS] or NULL
  long long rows_per_problem;
  int T, t_lo, knot_standoff, use_standoff, collision;
  float sw_obs, sw_goal;
  unsigned flags;
};

// ------------------------------------------------------------------------------------------------------------------
// PTX helpers: mbarrier + TMA (cp.async.bulk.tensor) + TF32 MMA
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  uint32_t spins = 0;
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (!ok && ++spins > (1u << 24)) __trap();  // a lost transaction must abort the launch, never hang the device
  }
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
      ::"r"(smem_u32(dst)), "l"((unsigned long long)map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ uint32_t f2tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ void mma_tf32_16x8x8(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                                uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// In-kernel launch timing: ts[0] = max over CTAs of ~start, ts[1] = max over warps of end (ns, %globaltimer); both zeroed
// before the solve.  Replaces CUDA events between the launches, which cost ~3 us of stream serialisation each.
__device__ __forceinline__ unsigned long long gto_globaltimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void stamp_begin(unsigned long long* ts) {
  if (ts && threadIdx.x == 0) atomicMax(ts, ~gto_globaltimer());
}
__device__ __forceinline__ void stamp_end(unsigned long long* ts) {
  if (ts && (threadIdx.x & 31) == 0) atomicMax(ts + 1, gto_globaltimer());
}

// Programmatic dependent launch (PDL): a kernel launched with the attribute may start while its predecessor in the stream is
// still running; pdl_wait() blocks until the predecessor has completed and its writes are visible, pdl_trigger() lets the
// successor start.  Everything a kernel does before pdl_wait() (shared-memory set-up, static tables) overlaps the tail of
// the previous kernel; the ~300 dependent launches of a solve otherwise pay ~5 us each.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// 3x4 product C = A * B (both [R|t] row-major, implicit last row 0 0 0 1)
// a0*b0 + a1*b1 + a2*b2 with a fixed association (explicit fused multiply-adds): every code shape that forms a frame
// entry -- unrolled 3x4 product, one entry per lane -- gives the same bits
template <typename T>
__device__ __forceinline__ T dot3(T a0, T b0, T a1, T b1, T a2, T b2) {
  return fma(a2, b2, fma(a1, b1, a0 * b0));
}
template <typename TA, typename TB, typename TC>
__device__ __forceinline__ void mul34(const TA* A, const TB* B, TC* C) {
#pragma unroll
  for (int r = 0; r < 3; ++r) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      TC s = dot3<TC>((TC)A[r * 4 + 0], (TC)B[c], (TC)A[r * 4 + 1], (TC)B[4 + c], (TC)A[r * 4 + 2], (TC)B[8 + c]);
      if (c == 3) s += A[r * 4 + 3];
      C[r * 4 + c] = s;
    }
  }
}

// Warp-level: add this chunk's rows (staged in shared memory as [32][RS]) into the tensor-core accumulators.
// A[m][k] = J[point k][col m], B[k][n] = J[point k][col n]  =>  D = J^T J.  NP = 8: one 16x8 tile (rows 8..15 unused);
// NP = 16: two n-tiles.  Error-compensated TF32 ("3xTF32"): every operand is split into hi = tf32(x) and lo = tf32(x - hi) and
// D += hi*hi + hi*lo + lo*hi, which brings the contraction from the 10-bit TF32 mantissa (5e-4 relative per product) to
// float32 accuracy -- the LM path (accept / reject decisions near kinks of the field) is sensitive to 1e-6 changes of J^T J.
__device__ __forceinline__ void tf32_split(float x, uint32_t& hi, uint32_t& lo) {
  hi = f2tf32(x);
  lo = f2tf32(x - __uint_as_float(hi));
}
template <int NP>
__device__ __forceinline__ void mma_rows(const float* st, int RS, int nrows32, float (&acc0)[4], float (&acc1)[4], int lane) {
  const int gq = lane >> 2, tq = lane & 3;
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    if (ks * 8 >= nrows32) break;
    const float* r0 = st + (ks * 8 + tq) * RS;
    const float* r1 = r0 + 4 * RS;
    uint32_t a0, a2, l0, l2;
    tf32_split(r0[gq], a0, l0);
    tf32_split(r1[gq], a2, l2);
    if (NP == 8) {
      mma_tf32_16x8x8(acc0, a0, 0u, a2, 0u, l0, l2);
      mma_tf32_16x8x8(acc0, l0, 0u, l2, 0u, a0, a2);
      mma_tf32_16x8x8(acc0, a0, 0u, a2, 0u, a0, a2);
    } else {
      uint32_t a1, a3, l1, l3;
      tf32_split(r0[gq + 8], a1, l1);
      tf32_split(r1[gq + 8], a3, l3);
      mma_tf32_16x8x8(acc0, a0, a1, a2, a3, l0, l2);
      mma_tf32_16x8x8(acc0, l0, l1, l2, l3, a0, a2);
      mma_tf32_16x8x8(acc0, a0, a1, a2, a3, a0, a2);
      mma_tf32_16x8x8(acc1, a0, a1, a2, a3, l1, l3);
      mma_tf32_16x8x8(acc1, l0, l1, l2, l3, a1, a3);
      mma_tf32_16x8x8(acc1, a0, a1, a2, a3, a1, a3);
    }
  }
}

// Warp-level: copy `nfl` floats of staged rows to global memory with coalesced stores (128-bit when aligned).
__device__ __forceinline__ void store_rows(float* __restrict__ dst, const float* st, int nfl, int lane) {
  if ((((uintptr_t)dst & 15) == 0) && ((nfl & 3) == 0)) {
    const float4* s4 = reinterpret_cast<const float4*>(st);
    float4* d4 = reinterpret_cast<float4*>(dst);
    for (int i = lane; i < (nfl >> 2); i += 32) __stcs(d4 + i, s4[i]);
  } else {
    for (int i = lane; i < nfl; i += 32) __stcs(dst + i, st[i]);
  }
}


#include "lin_cull.cuh"

// ------------------------------------------------------------------------------------------------------------------
// k_init: eliminate the equality constraints (gto/gto_planner.py:59-72): optimised rows of knots 0,1 = qc; clip the seed
// to the position limits (:138); parameter-joint entries are kept from the seed (optas/solver.py:126-159).
// ------------------------------------------------------------------------------------------------------------------
struct StateParams {
  const RobotDev* robot;
  int B, T;
  double dt, w_vel;
  const double* qc;      // [B][ndof]
  const double* q_seed;  // [B][T][ndof]
  double* Qc;            // [B][T][nopt] accepted point
  double* Qt;            // [B][T][nopt] trial point
  double* q_trial;       // [B][T][ndof] what k_linearize reads (float64: FK runs in double)
  int* bufsel;
  double *F, *Fp, *lam, *nu, *pred, *stepn;
  int *iters, *status;
  double lambda0;
  int project;
  // finalize
  double* outQ;
  double* outdQ;
  double* outcost;
  float* result;  // [B][nopt*T+2]
};

__global__ void k_init(const StateParams p) {
  const RobotDev& R = *p.robot;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)p.B * p.T) return;
  const int b = (int)(i / p.T), t = (int)(i % p.T);
  const double* s = p.q_seed + i * R.ndof;
  double* qt = p.q_trial + i * R.ndof;
  for (int j = 0; j < R.ndof; ++j) qt[j] = s[j];
  if (!p.project) return;
  for (int k = 0; k < R.nopt; ++k) {
    const int j = R.opt_qidx[k];
    double v = fmin(fmax(s[j], R.lo[k]), R.hi[k]);
    if (t < 2) v = p.qc[(long long)b * R.ndof + j];
    p.Qc[i * R.nopt + k] = v;
    p.Qt[i * R.nopt + k] = v;
    qt[j] = v;
  }
  if (t == 0) {
    p.bufsel[b] = 0;
    p.F[b] = 0.0; p.Fp[b] = 0.0; p.lam[b] = p.lambda0; p.nu[b] = 2.0; p.pred[b] = 0.0; p.stepn[b] = 0.0;
    p.iters[b] = 0; p.status[b] = -1;
  }
}

__global__ void k_finalize(const StateParams p) {
  const RobotDev& R = *p.robot;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)p.B * p.T) return;
  const int b = (int)(i / p.T), t = (int)(i % p.T);
  const int n = R.nopt;
  const double* s = p.q_seed + i * R.ndof;
  double* Q = p.outQ + i * R.ndof;
  for (int j = 0; j < R.ndof; ++j) Q[j] = s[j];
  for (int k = 0; k < n; ++k) {
    const double v = p.Qc[i * n + k];
    Q[R.opt_qidx[k]] = v;
    p.result[(long long)b * (n * p.T + 2) + t * n + k] = (float)v;
  }
  if (t < p.T - 1) {
    double* dQ = p.outdQ + ((long long)b * (p.T - 1) + t) * R.ndof;
    for (int j = 0; j < R.ndof; ++j) dQ[j] = 0.0;
    for (int k = 0; k < n; ++k) dQ[R.opt_qidx[k]] = (p.Qc[(i + 1) * n + k] - p.Qc[i * n + k]) / p.dt;
  }
  if (t == 0) {
    p.outcost[b] = p.F[b];
    p.result[(long long)b * (n * p.T + 2) + n * p.T] = (float)p.F[b];
    p.result[(long long)b * (n * p.T + 2) + n * p.T + 1] = (float)p.status[b];
  }
}

// ------------------------------------------------------------------------------------------------------------------
// parameters of k_step_cr (step_cr.cuh); mirrors oracle/gto_oracle.py solve_lm / lm_step (float64)
// ------------------------------------------------------------------------------------------------------------------
struct StepParams {
  const RobotDev* robot;
  int T;
  double dt, w_vel;
  int max_iter;
  double tol_step, tol_grad, lambda_min, lambda_max, eta, noise_rel, bound_eps, ftol, lambda_slow, slow_ftol;
  int slow_window;
  int as_rounds;
  double lambda_reject, lambda_conv;
  int bundle;      // pieces of the gradient bundle (0 .. GTO_BUNDLE_MAX)
  double bundle_radius;  // pieces further than this (|.|_inf, rad) from the standing point are not used
  double* gB;      // [B][GTO_BUNDLE_MAX][T-2][n] half gradients at the bundle points
  double* dyB;     // [B][GTO_BUNDLE_MAX][T-2][n] bundle point - standing point
  double* FB;      // [B][GTO_BUNDLE_MAX] cost at the bundle points
  int* nbund;      // [B] pieces held
  double* Fhist;  // [B][16] accepted cost per iteration (ring)
  double* Qc;
  double* Qt;
  double* q_trial;
  const float* H;
  const double* g;
  double* costp;
  long long buf_stride_H, buf_stride_g, buf_stride_c;
  int* bufsel;
  double *F, *Fp, *lam, *nu, *pred, *stepn;
  int *iters, *status;
  const int* active_in;
  const int* nactive_in;
  int* active_out;
  int* nactive_out;
  int iter;        // number of LM steps already taken by the problems in the active list
  long long* dbg;  // k_step_cr: clock64() at phase boundaries of CTA 0 in the launch of iteration dbg_iter (diagnostics, NULL: off)
  int dbg_iter, dbg_cta;
  unsigned long long* ts;  // k_step_cr: launch time stamps (NULL: off)
  int fk_robot_smem;  // k_step_cr: the robot table fits into the shared storage that is free during the FK phase
  int do_fk;       // k_step_cr: also write the item records (FK, brick placement, culling test) of the new trial point
  CullParams fk;   // what item_fk_body needs (robot, fields, records, Gauss-Newton buffers)
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__device__ __forceinline__ double shfl_d(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }

#include "step_cr.cuh"
#include "solve_fused.cuh"
#include "cloud_sdf.cuh"
#include "base_place.cuh"

// ------------------------------------------------------------------------------------------------------------------
// k_plan_cost: nearest-node cost of whole plans (seed ranking, gto/gto_models.py:204-215; clip-then-truncate indexing of
// points_to_offsets_numpy :190-201).  One block per (plan, knot).
// ------------------------------------------------------------------------------------------------------------------
__global__ void k_plan_cost(const RobotDev* robot, const float* px, const float* py, const float* pz, const float* q, int T,
                            FieldDev f, float bx, float by, float bz, double* cost) {
  __shared__ float Tm[GTO_MAX_MOV][12];
  __shared__ float frames[GTO_MAX_LINKS][12];
  __shared__ double part[32];
  const RobotDev& R = *robot;
  const int n = blockIdx.x / T, t = blockIdx.x % T;
  const float* qq = q + ((long long)n * T + t) * R.ndof;
  if (threadIdx.x == 0) {
    for (int j = 0; j < R.nmov; ++j) {
      const float qj = qq[R.mov_qidx[j]];
      const float ax = R.mov_axis[j][0], ay = R.mov_axis[j][1], az = R.mov_axis[j][2];
      float M[12], A[12];
      if (R.mov_type[j] == GTO_JOINT_REVOLUTE) {
        float s, c;
        sincosf(qj, &s, &c);
        const float v = 1.f - c;
        M[0] = 1.f - v * (ay * ay + az * az); M[1] = -s * az + v * ax * ay; M[2] = s * ay + v * ax * az; M[3] = 0.f;
        M[4] = s * az + v * ax * ay; M[5] = 1.f - v * (ax * ax + az * az); M[6] = -s * ax + v * ay * az; M[7] = 0.f;
        M[8] = -s * ay + v * ax * az; M[9] = s * ax + v * ay * az; M[10] = 1.f - v * (ax * ax + ay * ay); M[11] = 0.f;
      } else {
        M[0] = 1.f; M[1] = 0.f; M[2] = 0.f; M[3] = qj * ax;
        M[4] = 0.f; M[5] = 1.f; M[6] = 0.f; M[7] = qj * ay;
        M[8] = 0.f; M[9] = 0.f; M[10] = 1.f; M[11] = qj * az;
      }
      mul34(R.mov_origin[j], M, A);
      if (R.mov_parent[j] < 0) {
        for (int e = 0; e < 12; ++e) Tm[j][e] = A[e];
      } else {
        float C[12];
        mul34(Tm[R.mov_parent[j]], A, C);
        for (int e = 0; e < 12; ++e) Tm[j][e] = C[e];
      }
    }
    for (int l = 0; l < R.nlinks; ++l) {
      if (R.link_mov[l] < 0) {
        for (int e = 0; e < 12; ++e) frames[l][e] = R.link_tf[l][e];
      } else {
        float C[12];
        mul34(Tm[R.link_mov[l]], R.link_tf[l], C);
        for (int e = 0; e < 12; ++e) frames[l][e] = C[e];
      }
    }
  }
  __syncthreads();
  double s = 0.0;
  for (int l = 0; l < R.nlinks; ++l) {
    const float* F = frames[l];
    for (int i = threadIdx.x; i < R.link_pt_count[l]; i += blockDim.x) {
      const int pi = R.link_pt_start[l] + i;
      const float x = px[pi], y = py[pi], z = pz[pi];
      const float wx = F[0] * x + F[1] * y + F[2] * z + F[3] + bx;
      const float wy = F[4] * x + F[5] * y + F[6] * z + F[7] + by;
      const float wz = F[8] * x + F[9] * y + F[10] * z + F[11] + bz;
      const int ix = (int)fminf(fmaxf((wx - f.ox) * f.inv_pitch, 0.f), (float)(f.nx - 1));
      const int iy = (int)fminf(fmaxf((wy - f.oy) * f.inv_pitch, 0.f), (float)(f.ny - 1));
      const int iz = (int)fminf(fmaxf((wz - f.oz) * f.inv_pitch, 0.f), (float)(f.nz - 1));
      s += (double)f.data[((long long)ix * f.ny + iy) * f.nzp + iz];
    }
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += part[w];
    atomicAdd(cost + n, tot);
  }
}

// ==================================================================================================================
// host side: context + C-ABI
// ==================================================================================================================
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t cap = 0;
  cudaError_t ensure(size_t n) {
    if (n <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    cudaError_t e = cudaMalloc((void**)&p, n * sizeof(T));
    if (e == cudaSuccess) cap = n;
    return e;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
};

struct FieldHost {
  float* data = nullptr;
  CUtensorMap* maps2 = nullptr;
  unsigned* svt = nullptr;
  size_t svt_n = 0;
  unsigned nonzero = 0;
  int nx = 0, ny = 0, nz = 0, nzp = 0;
  double origin[3] = {0, 0, 0};
  double pitch = 0;
  bool set = false;
};

struct gto_ctx {
  int device = 0;
  int sm_count = 0;
  int max_smem_optin = 0;
  int smem_per_sm = 0;
  cudaStream_t stream = nullptr;
  std::string err;
  PFN_encodeTiled encode = nullptr;
  // robot
  bool has_robot = false;
  RobotDev robot_h;
  RobotDev* robot_d = nullptr;
  DevBuf<float> pts3;  // [3][npad]: x | y | z of the surface points
  int npad = 0;
  DevBuf<int> chunk_start, chunk_count;
  int pipe_cons = 8;
  double max_link_diag = 0.0;  // largest |half extent|_2 over links
  // fields
  std::vector<FieldHost> fields;
  FieldDev* fields_d = nullptr;
  double min_pitch = 0.0;
  // batch (resident)
  bool has_batch = false;
  bool solved = false;
  int B = 0, T = 0;
  double dt = 0, w_goal = 1, w_obs = 10, w_vel = 0.01;
  int standoff_offset = -10, use_standoff = 1, collision = 1;
  unsigned flags = 0;
  DevBuf<double> qc, q_seed, Qc, Qt, F, Fp, lam, nu, pred, stepn, Fhist, outQ, outdQ, outcost, gB, dyB, FB;
  DevBuf<int> nbund;
  DevBuf<double> q_trial, goal_tf;
  DevBuf<float> base, H, rows, result;
  DevBuf<double> g, costp;
  DevBuf<int> field_ids, bufsel, iters, status, active, nactive, work_ctr;
  DevBuf<unsigned long long> stats;
  DevBuf<long long> dbg;
  DevBuf<float4> cloud;      // depth point cloud (gto_cloud_set), padded to a multiple of CLOUD_TILE
  long long cloud_n = 0;
  DevBuf<double> cloud_q;
  DevBuf<float> cloud_depth, cloud_out, cloud_tiles;
  DevBuf<unsigned char> cloud_mask;
  DevBuf<unsigned long long> tstamps;
  double grip_mom[16] = {0};  // sum_k [x_k;1][x_k;1]^T over the gripper point set (base placement)
  DevBuf<double> base_d;      // base placement: inputs and outputs, one allocation
  DevBuf<float> base_occ;
  DevBuf<int> base_i;
  DevBuf<CullCtx> recs, rec_dummy;
  bool use_pdl = true;       // programmatic dependent launch of the solver kernels (gto_configure "pdl")
  // run-time tuning (gto_configure; defaults from the environment, read once in gto_create)
  double tune_jrows_budget_mb = 24576.0;
  int tune_step_fk = 0, tune_launch_events = 0, tune_step_dbg = 0, tune_cons = 0, tune_nslot = 4, tune_slot_floats = 0;
  int tune_fused = 0;  // 1: one persistent CTA per problem runs the whole solver loop (k_solve_fused, no launches, no chunking); default 0:
                       // one launch per phase and iteration -- measured faster on C2, where a problem's 28 knots spread over 28 CTAs
  long long cull_smem_set = -1, step_smem_set = -1, fused_smem_set = -1;
  int fused_occ = 0;
  DevBuf<int> queue;
  DevBuf<unsigned long long> phase_ns;  // dynamic shared memory the kernels were last configured for
  int cull_occ = 0;
  int* h_counter = nullptr;  // pinned, 16 ints
  cudaEvent_t ev_poll[2] = {nullptr, nullptr};  // convergence polls in flight (two parities)
  int tune_step_dbg_cta = 0;  // CTA (position in the active list) whose phase clocks step_dbg records
  int tune_blocking_sync = 0;  // 1: the host thread sleeps in the convergence polls instead of spinning (many contexts per core)
  long long rows_per_problem = 0;
  int Bchunk = 0;
  // profiling
  gto_profile prof;
  std::vector<cudaEvent_t> ev;
};

struct EventPair {  // two timing events that are destroyed on every return path
  cudaEvent_t a = nullptr, b = nullptr;
  cudaError_t create() {
    cudaError_t e = cudaEventCreate(&a);
    return e != cudaSuccess ? e : cudaEventCreate(&b);
  }
  ~EventPair() {
    if (a) cudaEventDestroy(a);
    if (b) cudaEventDestroy(b);
  }
};

static int fail(gto_ctx* c, int code, const std::string& msg) {
  if (c) c->err = msg;
  return code;
}
#define CK(call)                                                                                         \
  do {                                                                                                   \
    cudaError_t e_ = (call);                                                                             \
    if (e_ != cudaSuccess) return fail(ctx, GTO_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
  } while (0)

extern "C" int gto_abi_version(void) { return GTO_ABI_VERSION; }

extern "C" void gto_default_options(gto_options* o) {
  if (!o) return;
  o->max_iter = 100;
  o->tol_step = 1e-6;
  o->tol_grad = 1e-6;
  o->lambda0 = 1e-3;
  o->lambda_min = 1e-9;
  o->lambda_max = 1e9;
  o->eta = 1e-4;
  o->noise_rel = 1e-6;
  o->bound_eps = 1e-12;
  o->check_every = 4;
  o->ftol = 1e-6;
  o->lambda_slow = 1e30;
  o->slow_window = 0;
  o->slow_ftol = 1e-3;
  o->as_rounds = 1;
  o->lambda_reject = 1e-4;
  o->lambda_conv = 1e-2;
  o->bundle = 3;
  o->bundle_radius = 3e-3;
}

extern "C" int gto_configure(gto_ctx* ctx, const char* key, double value) {
  if (!ctx) return GTO_ERR_INVALID;
  if (!key) return fail(ctx, GTO_ERR_INVALID, "null or out-of-range argument");
  const std::string k(key);
  if (k == "jrows_budget_mb") ctx->tune_jrows_budget_mb = value > 0 ? value : 24576.0;
  else if (k == "pdl") ctx->use_pdl = value != 0;
  else if (k == "launch_events") ctx->tune_launch_events = value != 0;
  else if (k == "step_fk") ctx->tune_step_fk = (int)value;
  else if (k == "cull_nslot") { ctx->tune_nslot = (int)value; ctx->cull_smem_set = ctx->fused_smem_set = -1; }
  else if (k == "cons_warps") { ctx->tune_cons = (int)value; ctx->cull_smem_set = ctx->fused_smem_set = -1; }
  else if (k == "slot_floats") { ctx->tune_slot_floats = (int)value; ctx->cull_smem_set = ctx->fused_smem_set = -1; }
  else if (k == "step_dbg") ctx->tune_step_dbg = (int)value;  // iteration whose step launch records its phase clocks (0: off)
  else if (k == "fused") ctx->tune_fused = value != 0;
  else if (k == "step_dbg_cta") ctx->tune_step_dbg_cta = (int)value;
  else if (k == "blocking_sync") {  // the poll events are re-created with / without cudaEventBlockingSync at the next solve
    ctx->tune_blocking_sync = value != 0;
    for (int i = 0; i < 2; ++i)
      if (ctx->ev_poll[i]) { cudaEventDestroy(ctx->ev_poll[i]); ctx->ev_poll[i] = nullptr; }
  }
  else return fail(ctx, GTO_ERR_INVALID, "gto_configure: unknown key '" + k + "'");
  return GTO_OK;
}

extern "C" const char* gto_last_error(gto_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

extern "C" int gto_create(gto_ctx** out, int device) {
  if (!out) return GTO_ERR_INVALID;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) return GTO_ERR_NO_DEVICE;  // no CPU fallback, by design
  if (device < 0 || device >= ndev) return GTO_ERR_INVALID;
  gto_ctx* ctx = new gto_ctx();
  ctx->device = device;
  memset(&ctx->prof, 0, sizeof(ctx->prof));
  if (cudaSetDevice(device) != cudaSuccess) { delete ctx; return GTO_ERR_CUDA; }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete ctx; return GTO_ERR_CUDA; }
  if (prop.major < 10) {  // kernels are built for sm_100a only
    delete ctx;
    return GTO_ERR_NO_DEVICE;
  }
  ctx->sm_count = prop.multiProcessorCount;
  ctx->max_smem_optin = (int)prop.sharedMemPerBlockOptin;
  ctx->smem_per_sm = (int)prop.sharedMemPerMultiprocessor;
  if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return GTO_ERR_CUDA; }
  cudaDriverEntryPointQueryResult qres;
  void* fn = nullptr;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
    ctx->encode = (PFN_encodeTiled)fn;
  {  // tuning defaults from the environment (read once; gto_configure changes them afterwards)
    static const char* keys[] = {"jrows_budget_mb", "pdl", "launch_events", "step_fk", "cull_nslot", "cons_warps", "slot_floats", "step_dbg", "fused", "blocking_sync"};
    for (const char* k : keys) {
      std::string e = std::string("GTO_") + k;
      for (auto& ch : e) ch = (char)toupper((unsigned char)ch);
      if (const char* v = getenv(e.c_str())) gto_configure(ctx, k, atof(v));
    }
  }
  ctx->fields.resize(MAX_FIELDS);
  if (cudaMalloc((void**)&ctx->robot_d, sizeof(RobotDev)) != cudaSuccess ||
      cudaMalloc((void**)&ctx->fields_d, sizeof(FieldDev) * MAX_FIELDS) != cudaSuccess ||
      cudaMallocHost((void**)&ctx->h_counter, sizeof(int) * 16) != cudaSuccess) {
    gto_destroy(ctx);
    return GTO_ERR_NOMEM;
  }
  cudaMemset(ctx->fields_d, 0, sizeof(FieldDev) * MAX_FIELDS);
  *out = ctx;
  return GTO_OK;
}

extern "C" void gto_destroy(gto_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  for (auto& f : ctx->fields) {
    if (f.data) cudaFree(f.data);
    if (f.maps2) cudaFree(f.maps2);
    if (f.svt) cudaFree(f.svt);
  }
  for (auto e : ctx->ev) cudaEventDestroy(e);
  for (int i = 0; i < 2; ++i)
    if (ctx->ev_poll[i]) cudaEventDestroy(ctx->ev_poll[i]);
  if (ctx->robot_d) cudaFree(ctx->robot_d);
  if (ctx->fields_d) cudaFree(ctx->fields_d);
  if (ctx->h_counter) cudaFreeHost(ctx->h_counter);
  ctx->pts3.release(); ctx->chunk_start.release(); ctx->chunk_count.release();
  ctx->qc.release(); ctx->q_seed.release(); ctx->Qc.release(); ctx->Qt.release(); ctx->F.release(); ctx->Fp.release();
  ctx->lam.release(); ctx->nu.release(); ctx->pred.release(); ctx->stepn.release(); ctx->Fhist.release(); ctx->gB.release(); ctx->dyB.release(); ctx->FB.release(); ctx->nbund.release();
  ctx->outQ.release(); ctx->outdQ.release(); ctx->outcost.release();
  ctx->q_trial.release(); ctx->goal_tf.release(); ctx->base.release(); ctx->H.release(); ctx->g.release(); ctx->costp.release();
  ctx->rows.release(); ctx->result.release(); ctx->field_ids.release(); ctx->bufsel.release(); ctx->iters.release();
  ctx->status.release(); ctx->active.release(); ctx->nactive.release(); ctx->work_ctr.release(); ctx->stats.release(); ctx->dbg.release(); ctx->tstamps.release();
  ctx->cloud.release(); ctx->cloud_q.release(); ctx->cloud_depth.release(); ctx->cloud_out.release(); ctx->cloud_tiles.release(); ctx->cloud_mask.release(); ctx->recs.release(); ctx->rec_dummy.release(); ctx->queue.release(); ctx->phase_ns.release();
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

extern "C" int gto_set_robot(gto_ctx* ctx, const gto_robot_desc* r) {
  if (!ctx) return GTO_ERR_INVALID;
  if (!r) return fail(ctx, GTO_ERR_INVALID, "null or out-of-range argument");
  CK(cudaSetDevice(ctx->device));
  if (r->nopt < 1 || r->nopt > GTO_MAX_OPT || r->nmov < 1 || r->nmov > GTO_MAX_MOV || r->nlinks < 1 || r->nlinks > GTO_MAX_LINKS || r->nlinks > CULL_REC_LINKS ||
      r->ndof < r->nopt || r->npoints < 1)
    return fail(ctx, GTO_ERR_INVALID, "robot table sizes out of range");
  RobotDev& h = ctx->robot_h;
  memset(&h, 0, sizeof(h));
  h.ndof = r->ndof; h.nopt = r->nopt; h.nmov = r->nmov; h.nlinks = r->nlinks; h.npoints = r->npoints;
  for (int k = 0; k < r->nopt; ++k) {
    if (r->opt_qidx[k] < 0 || r->opt_qidx[k] >= r->ndof) return fail(ctx, GTO_ERR_INVALID, "opt_qidx out of range");
    h.opt_qidx[k] = r->opt_qidx[k];
    h.opt_mov[k] = -1;
    h.lo[k] = r->lo[k];
    h.hi[k] = r->hi[k];
    if (!(h.lo[k] <= h.hi[k])) return fail(ctx, GTO_ERR_INVALID, "joint limits must satisfy lo <= hi");
  }
  for (int j = 0; j < r->nmov; ++j) {
    if (r->mov_parent[j] >= j) return fail(ctx, GTO_ERR_INVALID, "movable joints must be listed parents first");
    if (r->mov_type[j] != GTO_JOINT_REVOLUTE && r->mov_type[j] != GTO_JOINT_PRISMATIC) return fail(ctx, GTO_ERR_INVALID, "unsupported joint type");
    if (r->mov_qidx[j] < 0 || r->mov_qidx[j] >= r->ndof) return fail(ctx, GTO_ERR_INVALID, "mov_qidx out of range");
    h.mov_parent[j] = r->mov_parent[j]; h.mov_type[j] = r->mov_type[j]; h.mov_qidx[j] = r->mov_qidx[j]; h.mov_opt[j] = r->mov_opt[j];
    for (int e = 0; e < 12; ++e) { h.mov_origin[j][e] = (float)r->mov_origin[j * 12 + e]; h.mov_origin_d[j][e] = r->mov_origin[j * 12 + e]; }
    for (int e = 0; e < 3; ++e) { h.mov_axis[j][e] = (float)r->mov_axis[j * 3 + e]; h.mov_axis_d[j][e] = r->mov_axis[j * 3 + e]; }
    if (r->mov_opt[j] >= 0) {
      if (r->mov_opt[j] >= r->nopt) return fail(ctx, GTO_ERR_INVALID, "mov_opt out of range");
      h.opt_mov[r->mov_opt[j]] = j;
    }
  }
  std::vector<float> hx(r->npoints), hy(r->npoints), hz(r->npoints);
  for (int i = 0; i < r->npoints; ++i) { hx[i] = r->points[3 * i]; hy[i] = r->points[3 * i + 1]; hz[i] = r->points[3 * i + 2]; }
  std::vector<int> cs, cc;
  ctx->max_link_diag = 0.0;
  for (int l = 0; l < r->nlinks; ++l) {
    const int s = r->link_pt_start[l], n = r->link_pt_count[l];
    if (s < 0 || n < 0 || s + n > r->npoints || r->link_mov[l] >= r->nmov) return fail(ctx, GTO_ERR_INVALID, "link table out of range");
    h.link_mov[l] = r->link_mov[l]; h.link_pt_start[l] = s; h.link_pt_count[l] = n; h.link_optmask[l] = r->link_optmask[l];
    for (int e = 0; e < 12; ++e) { h.link_tf[l][e] = (float)r->link_tf[l * 12 + e]; h.link_tf_d[l][e] = r->link_tf[l * 12 + e]; }
    float mn[3] = {1e30f, 1e30f, 1e30f}, mx[3] = {-1e30f, -1e30f, -1e30f};
    for (int i = s; i < s + n; ++i) {
      const float v[3] = {hx[i], hy[i], hz[i]};
      for (int a = 0; a < 3; ++a) { mn[a] = std::min(mn[a], v[a]); mx[a] = std::max(mx[a], v[a]); }
    }
    double diag = 0;
    for (int a = 0; a < 3; ++a) {
      h.link_center[l][a] = n ? 0.5f * (mn[a] + mx[a]) : 0.f;
      h.link_half[l][a] = n ? 0.5f * (mx[a] - mn[a]) : 0.f;
      diag += (double)h.link_half[l][a] * h.link_half[l][a];
    }
    ctx->max_link_diag = std::max(ctx->max_link_diag, sqrt(diag));
    h.link_chunk0[l] = (int)cs.size();
    for (int o = 0; o < n; o += 32) { cs.push_back(s + o); cc.push_back(std::min(32, n - o)); }
  }
  h.link_chunk0[r->nlinks] = (int)cs.size();
  h.nchunks = (int)cs.size();
  if (h.nchunks > MAX_CHUNKS * 64) return fail(ctx, GTO_ERR_INVALID, "too many surface points");
  h.grip_mov = r->grip_mov; h.grip_pt_start = r->grip_pt_start; h.grip_pt_count = r->grip_pt_count; h.grip_optmask = r->grip_optmask;
  if (r->grip_pt_start < 0 || r->grip_pt_count < 1 || r->grip_pt_start + r->grip_pt_count > r->npoints || r->grip_mov >= r->nmov)
    return fail(ctx, GTO_ERR_INVALID, "gripper point set out of range");
  for (int e = 0; e < 12; ++e) { h.grip_tf[e] = (float)r->grip_tf[e]; h.grip_tf_d[e] = r->grip_tf[e]; }
  for (int e = 0; e < 16; ++e) ctx->grip_mom[e] = 0.0;
  for (int i = r->grip_pt_start; i < r->grip_pt_start + r->grip_pt_count; ++i) {
    const double v[4] = {(double)hx[i], (double)hy[i], (double)hz[i], 1.0};
    for (int a = 0; a < 4; ++a)
      for (int c = 0; c < 4; ++c) ctx->grip_mom[4 * a + c] += v[a] * v[c];
  }
  {  // consumer warps of k_linearize_cull: chunks are dealt round-robin over the whole item
    int bestc = 8;
    double beff = 0;
    for (int w = CULL_MAX_CONS; w >= 5; --w) {  // prefer more warps on ties (latency hiding)
      const double eff = (double)h.nchunks / ((double)((h.nchunks + w - 1) / w) * w);
      if (eff > beff + 1e-9) { beff = eff; bestc = w; }
    }
    ctx->pipe_cons = bestc;
  }
  ctx->npad = (r->npoints + 3) & ~3;
  CK(ctx->pts3.ensure((size_t)3 * ctx->npad));
  CK(cudaMemset(ctx->pts3.p, 0, sizeof(float) * 3 * ctx->npad));
  CK(ctx->chunk_start.ensure(cs.size())); CK(ctx->chunk_count.ensure(cs.size()));
  CK(cudaMemcpy(ctx->pts3.p, hx.data(), sizeof(float) * r->npoints, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(ctx->pts3.p + ctx->npad, hy.data(), sizeof(float) * r->npoints, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(ctx->pts3.p + 2 * ctx->npad, hz.data(), sizeof(float) * r->npoints, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(ctx->chunk_start.p, cs.data(), sizeof(int) * cs.size(), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(ctx->chunk_count.p, cc.data(), sizeof(int) * cc.size(), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(ctx->robot_d, &h, sizeof(RobotDev), cudaMemcpyHostToDevice));
  ctx->has_robot = true;
  ctx->has_batch = false;
  return GTO_OK;
}

extern "C" int gto_set_field(gto_ctx* ctx, int slot, const float* cost, const int32_t dims[3], const double origin[3], double pitch) {
  if (!ctx) return GTO_ERR_INVALID;
  if (!cost || !dims || !origin) return fail(ctx, GTO_ERR_INVALID, "null or out-of-range argument");
  if (slot < 0 || slot >= MAX_FIELDS) return fail(ctx, GTO_ERR_INVALID, "field slot out of range");
  if (dims[0] < 2 || dims[1] < 2 || dims[2] < 2 || !(pitch > 0)) return fail(ctx, GTO_ERR_INVALID, "field needs >= 2 nodes per axis and pitch > 0");
  CK(cudaSetDevice(ctx->device));
  FieldHost& f = ctx->fields[slot];
  f.set = false;  // until every upload below has succeeded, the slot counts as unset (a failed re-upload must not leave stale maps in use)
  const int nzp = (dims[2] + 3) & ~3;  // TMA: global strides must be multiples of 16 bytes
  const size_t n = (size_t)dims[0] * dims[1] * nzp;
  if (f.data && (size_t)f.nx * f.ny * f.nzp != n) { cudaFree(f.data); f.data = nullptr; }
  if (!f.data) CK(cudaMalloc((void**)&f.data, n * sizeof(float)));
  CK(cudaMemset(f.data, 0, n * sizeof(float)));
  CK(cudaMemcpy2D(f.data, (size_t)nzp * sizeof(float), cost, (size_t)dims[2] * sizeof(float), (size_t)dims[2] * sizeof(float),
                  (size_t)dims[0] * dims[1], cudaMemcpyHostToDevice));
  f.nx = dims[0]; f.ny = dims[1]; f.nz = dims[2]; f.nzp = nzp; f.pitch = pitch;
  for (int a = 0; a < 3; ++a) f.origin[a] = origin[a];
  FieldDev d;
  d.data = f.data; d.nx = f.nx; d.ny = f.ny; d.nz = f.nz; d.nzp = nzp;
  d.ox = (float)origin[0]; d.oy = (float)origin[1]; d.oz = (float)origin[2]; d.inv_pitch = (float)(1.0 / pitch);
  d.org_d[0] = origin[0]; d.org_d[1] = origin[1]; d.org_d[2] = origin[2]; d.inv_pitch_d = 1.0 / pitch;
  d.maps2 = nullptr;
  if (ctx->encode) {  // one tile map per combination of per-axis box sizes
    std::vector<CUtensorMap> m2((size_t)CULL_NAXC * CULL_NAXC * CULL_NAXC);
    bool ok = true;
    for (int cx = 0; cx < CULL_NAXC && ok; ++cx)
      for (int cy = 0; cy < CULL_NAXC && ok; ++cy)
        for (int cz = 0; cz < CULL_NAXC && ok; ++cz) {
          const cuuint64_t gdim[3] = {(cuuint64_t)f.nz, (cuuint64_t)f.ny, (cuuint64_t)f.nx};
          const cuuint64_t gstr[2] = {(cuuint64_t)nzp * sizeof(float), (cuuint64_t)f.ny * nzp * sizeof(float)};
          const cuuint32_t box[3] = {(cuuint32_t)(8 + 4 * cz), (cuuint32_t)(8 + 4 * cy), (cuuint32_t)(8 + 4 * cx)};
          const cuuint32_t estr[3] = {1, 1, 1};
          CUresult rc = ctx->encode(&m2[((size_t)cx * CULL_NAXC + cy) * CULL_NAXC + cz], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)f.data, gdim,
                                    gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                    CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
          ok = (rc == CUDA_SUCCESS);
        }
    if (ok) {
      if (!f.maps2) CK(cudaMalloc((void**)&f.maps2, m2.size() * sizeof(CUtensorMap)));
      CK(cudaMemcpy(f.maps2, m2.data(), m2.size() * sizeof(CUtensorMap), cudaMemcpyHostToDevice));
      d.maps2 = f.maps2;
    }
  }
  {  // summed-volume table of the non-zero nodes: S[i][j][k] = #{cost != 0 in [0,i) x [0,j) x [0,k)} (culling test of k_linearize_cull)
    const size_t ex = (size_t)f.nx + 1, ey = (size_t)f.ny + 1, ez = (size_t)f.nz + 1, ns = ex * ey * ez;
    std::vector<unsigned> S(ns, 0u);
    // two passes on the host cores (a 256^3 field has 17 M entries): (1) the 2-D prefix sums of every x-slab, slabs in parallel;
    // (2) the running sum over the slabs, (y, z) rows in parallel.  Integer arithmetic: the same table as a single sweep.
    const unsigned nth = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    auto parallel_for = [&](size_t n, const std::function<void(size_t, size_t)>& body) {
      std::vector<std::thread> th;
      const size_t chunk = (n + nth - 1) / nth;
      for (unsigned w = 0; w < nth; ++w) {
        const size_t lo = std::min(n, w * chunk), hi = std::min(n, lo + chunk);
        if (lo < hi) th.emplace_back(body, lo, hi);
      }
      for (auto& t : th) t.join();
    };
    unsigned* Sp = S.data();
    parallel_for(ex - 1, [&](size_t lo, size_t hi) {
      for (size_t i = lo + 1; i <= hi; ++i)
        for (size_t j = 1; j < ey; ++j) {
          const float* src = cost + ((i - 1) * f.ny + (j - 1)) * f.nz;
          unsigned* row = Sp + (i * ey + j) * ez;
          const unsigned* up = Sp + (i * ey + (j - 1)) * ez;  // this slab, row j-1
          unsigned run = 0;                                   // non-zero nodes of this z-line so far
          for (size_t k = 1; k < ez; ++k) {
            run += (src[k - 1] != 0.0f) ? 1u : 0u;
            row[k] = run + up[k];
          }
        }
    });
    parallel_for(ey - 1, [&](size_t lo, size_t hi) {
      for (size_t i = 2; i < ex; ++i)
        for (size_t j = lo + 1; j <= hi; ++j) {
          unsigned* row = Sp + (i * ey + j) * ez;
          const unsigned* back = Sp + ((i - 1) * ey + j) * ez;  // previous slab, same row
          for (size_t k = 1; k < ez; ++k) row[k] += back[k];
        }
    });
    if (f.svt && f.svt_n != ns) { cudaFree(f.svt); f.svt = nullptr; }
    if (!f.svt) CK(cudaMalloc((void**)&f.svt, ns * sizeof(unsigned)));
    f.svt_n = ns;
    CK(cudaMemcpy(f.svt, S.data(), ns * sizeof(unsigned), cudaMemcpyHostToDevice));
    d.svt = f.svt;
    f.nonzero = S[ns - 1];
  }
  CK(cudaMemcpy(ctx->fields_d + slot, &d, sizeof(d), cudaMemcpyHostToDevice));
  f.set = true;
  ctx->min_pitch = 0.0;
  for (auto& ff : ctx->fields)
    if (ff.set) ctx->min_pitch = (ctx->min_pitch == 0.0) ? ff.pitch : std::min(ctx->min_pitch, ff.pitch);
  return GTO_OK;
}

static long long rows_per_problem(const gto_ctx* c) {
  const RobotDev& R = c->robot_h;
  return (c->collision ? (long long)c->T * R.npoints : 0) + 3LL * R.grip_pt_count * (1 + (c->use_standoff ? 1 : 0));
}

static int validate_batch(gto_ctx* ctx, const gto_batch_in* in) {
  if (!ctx->has_robot) return fail(ctx, GTO_ERR_STATE, "gto_set_robot has not been called");
  if (!in || in->B < 1 || in->T < 3 || !(in->dt > 0) || !in->qc || !in->q_seed || !in->goal_tf)
    return fail(ctx, GTO_ERR_INVALID, "batch needs B >= 1, T >= 3, dt > 0 and qc/q_seed/goal_tf");
  const int ks = in->T + in->standoff_offset;
  if (in->use_standoff && (ks < 0 || ks >= in->T)) return fail(ctx, GTO_ERR_INVALID, "stand-off knot outside the trajectory");
  if (in->collision_avoidance) {
    if (!in->field_all || !in->field_obs) return fail(ctx, GTO_ERR_INVALID, "collision_avoidance needs field_all/field_obs");
    for (int b = 0; b < in->B; ++b)
      for (int w = 0; w < 2; ++w) {
        const int s = w ? in->field_obs[b] : in->field_all[b];
        if (s >= MAX_FIELDS || (s >= 0 && !ctx->fields[s].set)) return fail(ctx, GTO_ERR_INVALID, "batch references a field slot that was never set");
      }
  }
  return GTO_OK;
}

extern "C" int gto_upload_batch(gto_ctx* ctx, const gto_batch_in* in) {
  if (!ctx) return GTO_ERR_INVALID;
  int rc = validate_batch(ctx, in);
  if (rc) return rc;
  CK(cudaSetDevice(ctx->device));
  const RobotDev& R = ctx->robot_h;
  const int B = in->B, T = in->T, n = R.nopt, nd = R.ndof, m = T - 2;
  ctx->B = B; ctx->T = T; ctx->dt = in->dt;
  ctx->w_goal = in->w_goal; ctx->w_obs = in->w_obs; ctx->w_vel = in->w_vel;
  ctx->standoff_offset = in->standoff_offset; ctx->use_standoff = in->use_standoff; ctx->collision = in->collision_avoidance;
  ctx->flags = in->flags;
  ctx->has_batch = false;
  ctx->solved = false;
  EventPair evp;
  CK(evp.create());
  cudaEvent_t e0 = evp.a, e1 = evp.b;
  CK(ctx->qc.ensure((size_t)B * nd)); CK(ctx->q_seed.ensure((size_t)B * T * nd));
  CK(ctx->goal_tf.ensure((size_t)B * 24)); CK(ctx->base.ensure((size_t)B * 4)); CK(ctx->field_ids.ensure((size_t)B * 2));
  CK(ctx->Qc.ensure((size_t)B * T * n)); CK(ctx->Qt.ensure((size_t)B * T * n)); CK(ctx->q_trial.ensure((size_t)B * T * nd));
  CK(ctx->F.ensure(B)); CK(ctx->Fp.ensure(B)); CK(ctx->lam.ensure(B)); CK(ctx->nu.ensure(B)); CK(ctx->pred.ensure(B)); CK(ctx->stepn.ensure(B));
  CK(ctx->bufsel.ensure(B)); CK(ctx->iters.ensure(B)); CK(ctx->status.ensure(B)); CK(ctx->active.ensure((size_t)2 * B));
  CK(ctx->H.ensure((size_t)2 * B * T * n * n)); CK(ctx->g.ensure((size_t)2 * B * T * n));
  CK(ctx->costp.ensure((size_t)2 * B * T));
  (void)m;
  CK(ctx->outQ.ensure((size_t)B * T * nd)); CK(ctx->outdQ.ensure((size_t)B * (T - 1) * nd)); CK(ctx->outcost.ensure(B));
  CK(ctx->result.ensure((size_t)B * (n * T + 2)));
  // host-side repacking of the small per-problem inputs (float32 copies for the point kernel)
  std::vector<float> bs((size_t)B * 4, 0.f);
  std::vector<int> fid((size_t)B * 2, -1);
  for (int b = 0; b < B; ++b) {
    if (in->base_position)
      for (int a = 0; a < 3; ++a) bs[4 * b + a] = (float)in->base_position[3 * b + a];
    if (in->collision_avoidance) { fid[2 * b] = in->field_all[b]; fid[2 * b + 1] = in->field_obs[b]; }
  }
  CK(cudaEventRecord(e0, ctx->stream));
  CK(cudaMemcpyAsync(ctx->qc.p, in->qc, sizeof(double) * B * nd, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(ctx->q_seed.p, in->q_seed, sizeof(double) * B * T * nd, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(ctx->goal_tf.p, in->goal_tf, sizeof(double) * B * 24, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(ctx->base.p, bs.data(), sizeof(float) * bs.size(), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(ctx->field_ids.p, fid.data(), sizeof(int) * fid.size(), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaEventRecord(e1, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  ctx->prof.h2d_ms = ms;
  ctx->prof.h2d_bytes = (long long)sizeof(double) * B * (nd * (1 + T) + 24) + sizeof(float) * bs.size() + sizeof(int) * fid.size();
  ctx->rows_per_problem = rows_per_problem(ctx);
  ctx->has_batch = true;
  return GTO_OK;
}

// launches one linearisation on ctx->stream
template <typename P>
static cudaError_t launch_pdl(void (*kern)(const P), unsigned grid, unsigned block, size_t smem, cudaStream_t stream, bool pdl, const P& params) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid, 1, 1);
  cfg.blockDim = dim3(block, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, params);
}

static void fill_lin_params(gto_ctx* ctx, LinParams& p, const double* q, const int* active, const int* nactive, int nproblems, int b0,
                            const int* bufsel, float* rows, int t_lo, unsigned flags) {
  const RobotDev& R = ctx->robot_h;
  memset(&p, 0, sizeof(p));
  p.robot = ctx->robot_d; p.chunk_start = ctx->chunk_start.p; p.chunk_count = ctx->chunk_count.p;
  p.pts3 = ctx->pts3.p; p.npad = ctx->npad;
  p.px = ctx->pts3.p; p.py = ctx->pts3.p + ctx->npad; p.pz = ctx->pts3.p + 2 * ctx->npad;
  p.q = q; p.goal_tf = ctx->goal_tf.p; p.base = ctx->base.p; p.field_ids = ctx->field_ids.p;
  p.fields = ctx->fields_d;
  p.active = active; p.nactive = nactive; p.nproblems = nproblems; p.b0 = b0; p.bufsel = bufsel;
  p.H = ctx->H.p; p.g = ctx->g.p; p.costp = ctx->costp.p;
  p.buf_stride_H = (long long)ctx->B * ctx->T * R.nopt * R.nopt;
  p.buf_stride_g = (long long)ctx->B * ctx->T * R.nopt;
  p.buf_stride_c = (long long)ctx->B * ctx->T;
  p.rows = rows; p.rows_per_problem = ctx->rows_per_problem;
  p.T = ctx->T; p.t_lo = t_lo; p.knot_standoff = ctx->T + ctx->standoff_offset; p.use_standoff = ctx->use_standoff;
  p.collision = ctx->collision;
  p.sw_obs = (float)sqrt(ctx->w_obs); p.sw_goal = (float)sqrt(ctx->w_goal);
  p.flags = flags;
}

// the linearise kernel stages SDF bricks by TMA: every field needs its tile maps (encoded at gto_set_field)
static bool fields_have_tile_maps(const gto_ctx* ctx) {
  for (auto& ff : ctx->fields)
    if (ff.set && !ff.maps2) return false;
  return true;
}
static int brick_slot_floats(const gto_ctx* ctx) {
  if (ctx->tune_slot_floats > 0) return std::max(512, ctx->tune_slot_floats & ~127);
  const int n3 = ctx->min_pitch > 0 ? (int)ceil(2.0 * ctx->max_link_diag / ctx->min_pitch) + 5 : 8;
  return n3 <= 24 ? 4096 : (n3 <= 32 ? 8192 : 12288);
}

// brick ring slots: as many as requested, but not so many that a second CTA no longer fits on the SM
static int pick_nslot(const gto_ctx* ctx, int nopt, int nc, int slot_floats, size_t extra) {
  const int want = std::min(CULL_NSLOT_MAX, std::max(2, ctx->tune_nslot));
  for (int ns = want; ns >= 2; --ns)
    if (2 * (cull_smem_bytes(nopt, nc, ns, slot_floats, ctx->npad) + extra + 1024) <= (size_t)ctx->smem_per_sm) return ns;
  return want;
}

typedef void (*cull_kernel_t)(const CullParams);
static cull_kernel_t pick_cull_kernel(int nopt) {
  if (nopt == 7) return k_linearize_cull<8, 7>;
  if (nopt == 8) return k_linearize_cull<8, 8>;
  if (nopt == 10) return k_linearize_cull<16, 10>;
  return nopt < 8 ? k_linearize_cull<8, 0> : k_linearize_cull<16, 0>;
}

// One linearisation on ctx->stream = k_item_fk (per-item records) + k_linearize_cull (points, rows, Gauss-Newton blocks).
// `have_recs`: the records of this launch were already written by the step kernel that produced the trial point.
static int launch_linearize(gto_ctx* ctx, const double* q, const int* active, const int* nactive, int nproblems, int b0, const int* bufsel,
                            float* rows, int t_lo, unsigned flags, int* work_counter, bool have_recs = false, unsigned long long* ts = nullptr) {
  cudaStream_t stream = ctx->stream;
  const RobotDev& R = ctx->robot_h;
  if (!fields_have_tile_maps(ctx)) return fail(ctx, GTO_ERR_STATE, "a cost field has no TMA tile maps (cuTensorMapEncodeTiled unavailable)");
  CullParams cp;
  memset(&cp, 0, sizeof(cp));
  fill_lin_params(ctx, cp.lin, q, active, nactive, nproblems, b0, bufsel, rows, t_lo, flags);
  const int slot_floats = brick_slot_floats(ctx);
  cp.slot_floats = slot_floats;
  const int nc = ctx->tune_cons > 0 ? std::min(CULL_MAX_CONS, ctx->tune_cons) : ctx->pipe_cons;
  cp.ncons = nc;
  const int nslot = pick_nslot(ctx, R.nopt, nc, slot_floats, 0);
  cp.nslot = nslot;
  cp.count_early = have_recs ? 0 : 1;
  cp.work_counter = work_counter;
  cp.stats = ctx->stats.p;
  const size_t sm = cull_smem_bytes(R.nopt, nc, nslot, slot_floats, ctx->npad);
  const int threads = (nc + 2) * 32;  // consumers + producer + zero-row warp
  cull_kernel_t kern = pick_cull_kernel(R.nopt);
  if (ctx->cull_smem_set != (long long)sm) {
    int occ = 0;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, sm);
    if (e != cudaSuccess) return fail(ctx, GTO_ERR_CUDA, std::string("linearise launch setup: ") + cudaGetErrorString(e));
    if (occ < 1) return fail(ctx, GTO_ERR_INVALID, "linearise kernel does not fit on an SM (brick slots too large)");
    ctx->cull_smem_set = (long long)sm;
    ctx->cull_occ = occ;
  }
  const long long max_items = (long long)nproblems * (ctx->T - t_lo);
  const int grid = (int)std::max(1LL, std::min<long long>((long long)ctx->sm_count * ctx->cull_occ, max_items));
  cudaError_t e = ctx->recs.ensure((size_t)std::max<long long>(max_items, (long long)ctx->Bchunk * ctx->T));
  if (e == cudaSuccess) e = ctx->rec_dummy.ensure(1);
  if (e != cudaSuccess) return fail(ctx, GTO_ERR_NOMEM, "item records");
  cp.recs = ctx->recs.p;
  cp.rec_dummy = ctx->rec_dummy.p;
  cp.dbg = ctx->tune_step_dbg ? ctx->dbg.p : nullptr;
  cp.ts_fk = ts;
  cp.ts_lin = ts ? ts + 2 : nullptr;
  if (!have_recs) {
    const size_t fk_smem = ((sizeof(RobotDev) + 15) & ~(size_t)15) + (size_t)8 * 2 * R.nmov * 12 * sizeof(double);
    e = launch_pdl<CullParams>(k_item_fk, (unsigned)((max_items + 7) / 8), 128, fk_smem, stream, ctx->use_pdl, cp);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) return fail(ctx, GTO_ERR_CUDA, std::string("k_item_fk launch: ") + cudaGetErrorString(e));
    ctx->prof.kernel_launches += 1;
  }
  e = launch_pdl<CullParams>(kern, (unsigned)grid, (unsigned)threads, sm, stream, ctx->use_pdl, cp);
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) return fail(ctx, GTO_ERR_CUDA, std::string("k_linearize_cull launch: ") + cudaGetErrorString(e));
  ctx->prof.kernel_launches += 1;
  return GTO_OK;
}

static cudaEvent_t get_event(gto_ctx* ctx, size_t i) {
  while (ctx->ev.size() <= i) {
    cudaEvent_t e;
    cudaEventCreate(&e);
    ctx->ev.push_back(e);
  }
  return ctx->ev[i];
}

typedef void (*step_kernel_t)(const StepParams);
static step_kernel_t pick_step_kernel(int n) {
  if (n == 7) return k_step_cr<7, true>;
  if (n == 8) return k_step_cr<8, true>;
  if (n == 10) return k_step_cr<10, true>;
  return n < 8 ? k_step_cr<8, false> : k_step_cr<16, false>;
}

extern "C" int gto_solve_resident(gto_ctx* ctx, const gto_options* user_opts) {
  if (!ctx) return GTO_ERR_INVALID;
  if (!ctx->has_batch) return fail(ctx, GTO_ERR_STATE, "no batch uploaded");
  CK(cudaSetDevice(ctx->device));
  gto_options o;
  gto_default_options(&o);
  if (user_opts) o = *user_opts;
  if (o.max_iter < 1 || o.check_every < 1) return fail(ctx, GTO_ERR_INVALID, "max_iter and check_every must be >= 1");
  const RobotDev& R = ctx->robot_h;
  const int B = ctx->B, T = ctx->T, n = R.nopt;
  const bool want_rows = !(ctx->flags & GTO_FLAG_NO_JROWS);
  // Jacobian-row buffer: problems are processed in chunks that fit the budget
  const double budget = ctx->tune_jrows_budget_mb * 1048576.0;
  const double per_problem = (double)ctx->rows_per_problem * (n + 1) * sizeof(float);
  int Bchunk = B;
  if (want_rows) {
    Bchunk = (int)std::max(1.0, std::min((double)B, floor(budget / per_problem)));
    CK(ctx->rows.ensure((size_t)Bchunk * ctx->rows_per_problem * (n + 1)));
  }
  ctx->Bchunk = Bchunk;
  CK(ctx->stats.ensure(4));
  CK(cudaMemsetAsync(ctx->stats.p, 0, sizeof(unsigned long long) * 4, ctx->stream));

  StateParams sp;
  memset(&sp, 0, sizeof(sp));
  sp.robot = ctx->robot_d; sp.B = B; sp.T = T; sp.dt = ctx->dt; sp.w_vel = ctx->w_vel;
  sp.qc = ctx->qc.p; sp.q_seed = ctx->q_seed.p; sp.Qc = ctx->Qc.p; sp.Qt = ctx->Qt.p; sp.q_trial = ctx->q_trial.p;
  sp.bufsel = ctx->bufsel.p; sp.F = ctx->F.p; sp.Fp = ctx->Fp.p; sp.lam = ctx->lam.p; sp.nu = ctx->nu.p; sp.pred = ctx->pred.p;
  sp.stepn = ctx->stepn.p; sp.iters = ctx->iters.p; sp.status = ctx->status.p; sp.lambda0 = o.lambda0; sp.project = 1;
  sp.outQ = ctx->outQ.p; sp.outdQ = ctx->outdQ.p; sp.outcost = ctx->outcost.p; sp.result = ctx->result.p;

  StepParams st;
  memset(&st, 0, sizeof(st));
  st.robot = ctx->robot_d; st.T = T; st.dt = ctx->dt; st.w_vel = ctx->w_vel; st.max_iter = o.max_iter;
  st.tol_step = o.tol_step; st.tol_grad = o.tol_grad; st.lambda_min = o.lambda_min; st.lambda_max = o.lambda_max; st.eta = o.eta;
  st.noise_rel = o.noise_rel; st.bound_eps = o.bound_eps; st.ftol = o.ftol; st.lambda_slow = o.lambda_slow;
  st.slow_ftol = o.slow_ftol; st.slow_window = std::min(16, std::max(0, (int)o.slow_window));
  st.as_rounds = std::min(4, std::max(0, (int)o.as_rounds)); st.lambda_reject = o.lambda_reject; st.lambda_conv = o.lambda_conv;
  CK(ctx->Fhist.ensure((size_t)B * 16));
  st.Fhist = ctx->Fhist.p;
  st.bundle = std::min(GTO_BUNDLE_MAX, std::max(0, (int)o.bundle));
  st.bundle_radius = o.bundle_radius;
  CK(ctx->gB.ensure((size_t)B * GTO_BUNDLE_MAX * (T - 2) * n));
  CK(ctx->dyB.ensure((size_t)B * GTO_BUNDLE_MAX * (T - 2) * n));
  CK(ctx->FB.ensure((size_t)B * GTO_BUNDLE_MAX));
  CK(ctx->nbund.ensure((size_t)B));
  CK(cudaMemsetAsync(ctx->nbund.p, 0, sizeof(int) * B, ctx->stream));
  st.gB = ctx->gB.p; st.dyB = ctx->dyB.p; st.FB = ctx->FB.p; st.nbund = ctx->nbund.p;
  st.Qc = ctx->Qc.p; st.Qt = ctx->Qt.p; st.q_trial = ctx->q_trial.p; st.H = ctx->H.p; st.g = ctx->g.p; st.costp = ctx->costp.p;
  st.buf_stride_H = (long long)B * T * n * n; st.buf_stride_g = (long long)B * T * n; st.buf_stride_c = (long long)B * T;
  st.bufsel = ctx->bufsel.p; st.F = ctx->F.p; st.Fp = ctx->Fp.p; st.lam = ctx->lam.p; st.nu = ctx->nu.p; st.pred = ctx->pred.p;
  st.stepn = ctx->stepn.p; st.iters = ctx->iters.p; st.status = ctx->status.p;
  // LM step: block cyclic reduction, one CTA per problem (k_step_cr)
  step_kernel_t step_kern = pick_step_kernel(n);
  const size_t cr_smem = step_cr_smem_bytes(T, n, st.bundle);
  if (cr_smem > (size_t)ctx->max_smem_optin)
    return fail(ctx, GTO_ERR_INVALID, "(T-2) * nopt^2 too large: the block-tridiagonal factor must fit in shared memory");
  if (ctx->step_smem_set != (long long)cr_smem) {
    CK(cudaFuncSetAttribute(step_kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cr_smem));
    ctx->step_smem_set = (long long)cr_smem;
  }
  // optionally the step kernel also writes the item records of its new trial point (tune "step_fk" = problem count below
  // which it does; measured slower than a k_item_fk launch on C2, off by default)
  const int step_fk_limit = ctx->tune_step_fk;
  const bool step_fk_ok = step_fk_limit > 0 && (size_t)(STEP_CR_THREADS / 16) * 2 * R.nmov * 12 <= (size_t)3 * (T - 2) * n * n;
  CK(ctx->recs.ensure((size_t)Bchunk * T));

  gto_profile& pf = ctx->prof;
  pf.solve_ms = pf.linearize_ms = pf.step_ms = 0;
  pf.linearize_launches = pf.step_launches = pf.iterations = 0;
  pf.knot_items = 0; pf.jrow_bytes = 0; pf.problem_iterations = 0; pf.linearize_launches_with_work = 0;
  pf.links_tested = pf.links_active = 0;
  pf.kernel_launches = 2;  // k_init + k_finalize; the linearise / step launches are added where they happen
  size_t nev = 0;
  std::vector<int> ev_kind;
  cudaEvent_t ev_begin = get_event(ctx, nev++);
  CK(cudaEventRecord(ev_begin, ctx->stream));
  {
    const long long tot = (long long)B * T;
    k_init<<<(unsigned)((tot + 127) / 128), 128, 0, ctx->stream>>>(sp);
    CK(cudaGetLastError());
  }
  if (ctx->tune_fused) {
    // ---- default: one persistent CTA per problem runs the whole solver loop (solve_fused.cuh) ----
    if (!fields_have_tile_maps(ctx)) return fail(ctx, GTO_ERR_STATE, "a cost field has no TMA tile maps (cuTensorMapEncodeTiled unavailable)");
    FusedParams fp;
    memset(&fp, 0, sizeof(fp));
    const int slot_floats = brick_slot_floats(ctx);
    const int nc = ctx->tune_cons > 0 ? std::min(CULL_MAX_CONS, ctx->tune_cons) : ctx->pipe_cons;
    const int nslot = pick_nslot(ctx, n, nc, slot_floats, fused_robot_bytes());
    const int threads = (nc + 2) * 32;
    const size_t phase_smem = std::max(std::max(cull_smem_bytes(n, nc, nslot, slot_floats, ctx->npad), (cr_smem + 127) & ~(size_t)127),
                                       (fused_fk_smem_bytes(R.nmov, threads) + 127) & ~(size_t)127);
    const size_t smem = fused_robot_bytes() + phase_smem;
    if (smem > (size_t)ctx->max_smem_optin) return fail(ctx, GTO_ERR_INVALID, "fused solver kernel does not fit in shared memory");
    void (*kern)(const FusedParams) = nullptr;
    if (n == 7) kern = k_solve_fused<8, 7, 7, true>;
    else if (n == 8) kern = k_solve_fused<8, 8, 8, true>;
    else if (n == 10) kern = k_solve_fused<16, 10, 10, true>;
    else if (n < 8) kern = k_solve_fused<8, 0, 8, false>;
    else kern = k_solve_fused<16, 0, 16, false>;
    if (ctx->fused_smem_set != (long long)smem) {
      int occ = 0;
      CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, smem));
      if (occ < 1) return fail(ctx, GTO_ERR_INVALID, "fused solver kernel does not fit on an SM");
      ctx->fused_smem_set = (long long)smem;
      ctx->fused_occ = occ;
    }
    const int grid = std::max(1, std::min(B, ctx->sm_count * ctx->fused_occ));
    // the row buffer and the item records are per CTA slot, not per problem: no chunking, whatever the batch size
    if (want_rows) CK(ctx->rows.ensure((size_t)grid * ctx->rows_per_problem * (n + 1)));
    CK(ctx->recs.ensure((size_t)grid * T));
    CK(ctx->rec_dummy.ensure(1));
    CK(ctx->queue.ensure(1));
    CK(ctx->phase_ns.ensure(4));
    CK(cudaMemsetAsync(ctx->queue.p, 0, sizeof(int), ctx->stream));
    CK(cudaMemsetAsync(ctx->phase_ns.p, 0, 4 * sizeof(unsigned long long), ctx->stream));
    fill_lin_params(ctx, fp.cull.lin, ctx->q_trial.p, nullptr, nullptr, B, 0, ctx->bufsel.p, want_rows ? ctx->rows.p : nullptr, 2, ctx->flags);
    fp.cull.slot_floats = slot_floats; fp.cull.ncons = nc; fp.cull.nslot = nslot; fp.cull.count_early = 0;
    fp.cull.stats = ctx->stats.p; fp.cull.recs = ctx->recs.p; fp.cull.rec_dummy = ctx->rec_dummy.p;
    fp.step = st;
    fp.B = B; fp.queue = ctx->queue.p; fp.phase_ns = ctx->phase_ns.p;
    ctx->Bchunk = grid;
    kern<<<grid, threads, smem, ctx->stream>>>(fp);
    CK(cudaGetLastError());
    pf.kernel_launches += 1;
    {
      const long long tot = (long long)B * T;
      k_finalize<<<(unsigned)((tot + 127) / 128), 128, 0, ctx->stream>>>(sp);
      CK(cudaGetLastError());
    }
    cudaEvent_t ev_end = get_event(ctx, nev++);
    CK(cudaEventRecord(ev_end, ctx->stream));
    std::vector<int> h_it(B);
    unsigned long long hs[4] = {0, 0, 0, 0}, hp[4] = {0, 0, 0, 0};
    CK(cudaMemcpyAsync(h_it.data(), ctx->iters.p, sizeof(int) * B, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(hs, ctx->stats.p, sizeof(hs), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(hp, ctx->phase_ns.p, sizeof(hp), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, ev_begin, ev_end));
    pf.solve_ms = ms;
    // exact work accounting: a problem that stopped after `it` steps was linearised it + 1 times (all T knots the first
    // time, T - 2 afterwards; a problem that ran into max_iter is linearised once more to judge its last trial point)
    const long long rows_it = ctx->rows_per_problem - 2LL * (ctx->collision ? R.npoints : 0);
    for (int b = 0; b < B; ++b) {
      const long long nl = (long long)std::min(h_it[b], o.max_iter) + 1;
      pf.problem_iterations += nl;
      pf.knot_items += T + (nl - 1) * (T - 2);
      if (want_rows) pf.jrow_bytes += (ctx->rows_per_problem + (nl - 1) * rows_it) * (n + 1) * 4;
      pf.iterations = std::max(pf.iterations, h_it[b]);
    }
    pf.linearize_launches = pf.step_launches = 0;
    pf.linearize_launches_with_work = 0;
    pf.links_tested = (long long)hs[1];
    pf.links_active = (long long)hs[2];
    // phase times: ns summed over the CTAs / number of CTAs = average time a CTA spent in each phase
    pf.linearize_ms = (double)(hp[0] + hp[1]) * 1e-6 / grid;
    pf.step_ms = (double)hp[2] * 1e-6 / grid;
    ctx->solved = true;
    return GTO_OK;
  }
  std::vector<int> ident(B);
  for (int b = 0; b < B; ++b) ident[b] = b;
  for (int i = 0; i < 2; ++i)
    if (!ctx->ev_poll[i]) CK(cudaEventCreateWithFlags(&ctx->ev_poll[i], cudaEventDisableTiming | (ctx->tune_blocking_sync ? cudaEventBlockingSync : 0)));
  const size_t cstride = (size_t)o.max_iter + 3;  // one counter per iteration
  CK(ctx->nactive.ensure(cstride));
  CK(ctx->work_ctr.ensure(cstride));
  if (ctx->tune_step_dbg) {
    CK(ctx->dbg.ensure(64));
    CK(cudaMemsetAsync(ctx->dbg.p, 0, 64 * sizeof(long long), ctx->stream));
    st.dbg = ctx->dbg.p;
    st.dbg_iter = ctx->tune_step_dbg;
    st.dbg_cta = ctx->tune_step_dbg_cta;
  }
  std::vector<int> h_nact(cstride);
  // per-launch durations for the profile (linearize_ms / step_ms): in-kernel %globaltimer stamps by default; CUDA events
  // between the launches on request (they cost ~3 us of stream serialisation each)
  const bool launch_events = ctx->tune_launch_events != 0;
  const bool launch_stamps = !launch_events;
  const size_t ts_n = cstride * 6;
  if (launch_stamps) {
    CK(ctx->tstamps.ensure(ts_n));
    CK(cudaMemsetAsync(ctx->tstamps.p, 0, ts_n * sizeof(unsigned long long), ctx->stream));
  }

  for (int b0 = 0; b0 < B; b0 += Bchunk) {
    const int nb = std::min(Bchunk, B - b0);
    CK(cudaMemsetAsync(ctx->nactive.p, 0, sizeof(int) * cstride, ctx->stream));
    CK(cudaMemsetAsync(ctx->work_ctr.p, 0, sizeof(int) * cstride, ctx->stream));
    CK(cudaMemcpyAsync(ctx->active.p + b0, ident.data() + b0, sizeof(int) * nb, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->nactive.p, &nb, sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    int* act0 = ctx->active.p + b0;
    int* act1 = ctx->active.p + B + b0;
    int known = nb;  // upper bound of the problems still active (last polled count; it only decreases): sizes the grids
    bool recs_ready = false, done = false;
    size_t nblk = 0;
    for (int it = 0; it <= o.max_iter && !done; ++it) {
      int* ain = (it & 1) ? act1 : act0;
      int* aout = (it & 1) ? act0 : act1;
      cudaEvent_t a = nullptr, bE = nullptr, c = nullptr;
      if (launch_events) {
        a = get_event(ctx, nev++); bE = get_event(ctx, nev++); c = get_event(ctx, nev++);
        CK(cudaEventRecord(a, ctx->stream));
      }
      int rc = launch_linearize(ctx, ctx->q_trial.p, ain, ctx->nactive.p + it, known, b0, ctx->bufsel.p, want_rows ? ctx->rows.p : nullptr,
                                it == 0 ? 0 : 2, ctx->flags, ctx->work_ctr.p + it, recs_ready && it > 0,
                                launch_stamps ? ctx->tstamps.p + (size_t)it * 6 : nullptr);
      if (rc) return rc;
      if (launch_events) CK(cudaEventRecord(bE, ctx->stream));
      st.active_in = ain; st.nactive_in = ctx->nactive.p + it; st.active_out = aout; st.nactive_out = ctx->nactive.p + it + 1;
      st.iter = it;
      st.ts = launch_stamps ? ctx->tstamps.p + (size_t)it * 6 + 4 : nullptr;
      const bool step_fk = step_fk_ok && known <= step_fk_limit;
      st.do_fk = step_fk ? 1 : 0;
      st.fk_robot_smem = sizeof(RobotDev) <= (size_t)T * n * 8 + (size_t)(T - 2) * n * 8 + (size_t)2 * (T - 2) * n * n * 4 ? 1 : 0;
      recs_ready = step_fk;
      if (step_fk) {
        memset(&st.fk, 0, sizeof(st.fk));
        fill_lin_params(ctx, st.fk.lin, ctx->q_trial.p, nullptr, nullptr, nb, b0, ctx->bufsel.p, want_rows ? ctx->rows.p : nullptr, 2, ctx->flags);
        st.fk.slot_floats = brick_slot_floats(ctx);
        st.fk.recs = ctx->recs.p;
        CK(ctx->rec_dummy.ensure(1));
        st.fk.rec_dummy = ctx->rec_dummy.p;
        st.fk.stats = ctx->stats.p;
      }
      CK(launch_pdl<StepParams>(step_kern, (unsigned)known, STEP_CR_THREADS, cr_smem, ctx->stream, ctx->use_pdl, st));
      pf.kernel_launches += 1;
      CK(cudaGetLastError());
      if (launch_events) {
        CK(cudaEventRecord(c, ctx->stream));
        ev_kind.push_back(0);
      }
      pf.linearize_launches++;
      pf.step_launches++;
      pf.iterations = std::max(pf.iterations, it);
      if (((it + 1) % o.check_every) == 0 || it == o.max_iter) {
        // convergence poll, one block behind: the counter of this block is requested now and looked at after the NEXT block
        // has been enqueued, so the GPU never waits for the host (cost: up to check_every empty iterations at the end)
        const int par = (int)(nblk & 1);
        CK(cudaMemcpyAsync(ctx->h_counter + par, ctx->nactive.p + it + 1, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaEventRecord(ctx->ev_poll[par], ctx->stream));
        if (it == o.max_iter) {
          CK(cudaStreamSynchronize(ctx->stream));
          done = true;
        } else if (nblk > 0) {
          CK(cudaEventSynchronize(ctx->ev_poll[par ^ 1]));
          if (ctx->h_counter[par ^ 1] == 0) done = true;
          else known = std::min(known, ctx->h_counter[par ^ 1]);
        }
        ++nblk;
      }
    }
    if (launch_stamps) {  // per-launch durations of this chunk; the stamp array is reused by the next chunk
      std::vector<unsigned long long> hts(ts_n);
      CK(cudaMemcpyAsync(hts.data(), ctx->tstamps.p, ts_n * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
      CK(cudaMemsetAsync(ctx->tstamps.p, 0, ts_n * sizeof(unsigned long long), ctx->stream));
      CK(cudaStreamSynchronize(ctx->stream));
      for (size_t i = 0; i + 1 < ts_n; i += 2) {
        if (!hts[i + 1] || !hts[i]) continue;
        const unsigned long long t0 = ~hts[i], t1 = hts[i + 1];
        if (t1 <= t0) continue;
        const double msd = (double)(t1 - t0) * 1e-6;
        if ((i / 2) % 3 == 2) pf.step_ms += msd;
        else pf.linearize_ms += msd;
      }
    }
    // exact work accounting for the roofline: problems active in every linearise launch of this chunk
    CK(cudaMemcpyAsync(h_nact.data(), ctx->nactive.p, sizeof(int) * cstride, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    for (int it = 0; it <= o.max_iter; ++it) {
      const long long na = h_nact[it];
      if (na <= 0) break;
      pf.problem_iterations += na;
      pf.linearize_launches_with_work++;
      pf.knot_items += na * (it == 0 ? T : T - 2);
      if (want_rows)
        pf.jrow_bytes += na * (it == 0 ? ctx->rows_per_problem : ctx->rows_per_problem - 2LL * (ctx->collision ? R.npoints : 0)) * (n + 1) * 4;
    }
  }
  {
    const long long tot = (long long)B * T;
    k_finalize<<<(unsigned)((tot + 127) / 128), 128, 0, ctx->stream>>>(sp);
    CK(cudaGetLastError());
  }
  cudaEvent_t ev_end = get_event(ctx, nev++);
  CK(cudaEventRecord(ev_end, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  {
    unsigned long long hs[4] = {0, 0, 0, 0};
    CK(cudaMemcpy(hs, ctx->stats.p, sizeof(hs), cudaMemcpyDeviceToHost));
    pf.links_tested = (long long)hs[1];
    pf.links_active = (long long)hs[2];
  }
  if (st.dbg) {
    long long hd[64];
    CK(cudaMemcpy(hd, ctx->dbg.p, sizeof(hd), cudaMemcpyDeviceToHost));
    fprintf(stderr, "[gto] k_step_cr phase clocks (cycles between marks, launch of iteration %d, CTA %d):", ctx->tune_step_dbg, ctx->tune_step_dbg_cta);
    for (int i = 1; i < 32 && hd[i]; ++i) fprintf(stderr, " %lld", hd[i] - hd[i - 1]);
    fprintf(stderr, "\n[gto] k_step_cr launch of iteration %d: %lld CTAs reached the solve, %lld re-solved for the active set, %lld used bundle planes", ctx->tune_step_dbg,
            hd[40], hd[41], hd[42]);
    fprintf(stderr, "\n[gto] k_item_fk phase clocks (cycles since pdl_wait):");
    for (int i = 33; i < 40 && hd[i]; ++i) fprintf(stderr, " %lld", hd[i] - hd[32]);
    fprintf(stderr, "\n");
  }
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, ev_begin, ev_end));
  pf.solve_ms = ms;
  for (size_t i = 0; i < ev_kind.size(); ++i) {
    float m1 = 0, m2 = 0;
    cudaEventElapsedTime(&m1, ctx->ev[1 + 3 * i], ctx->ev[2 + 3 * i]);
    cudaEventElapsedTime(&m2, ctx->ev[2 + 3 * i], ctx->ev[3 + 3 * i]);
    pf.linearize_ms += m1;
    pf.step_ms += m2;
  }
  ctx->solved = true;
  return GTO_OK;
}

extern "C" int gto_download_batch(gto_ctx* ctx, gto_batch_out* out) {
  if (!ctx) return GTO_ERR_INVALID;
  if (!out) return fail(ctx, GTO_ERR_INVALID, "null or out-of-range argument");
  if (!ctx->solved) return fail(ctx, GTO_ERR_STATE, "no solved batch to download");
  CK(cudaSetDevice(ctx->device));
  const int B = ctx->B, T = ctx->T, nd = ctx->robot_h.ndof;
  EventPair evp;
  CK(evp.create());
  cudaEvent_t e0 = evp.a, e1 = evp.b;
  CK(cudaEventRecord(e0, ctx->stream));
  long long bytes = 0;
  if (out->Q) { CK(cudaMemcpyAsync(out->Q, ctx->outQ.p, sizeof(double) * B * T * nd, cudaMemcpyDeviceToHost, ctx->stream)); bytes += sizeof(double) * B * T * nd; }
  if (out->dQ) { CK(cudaMemcpyAsync(out->dQ, ctx->outdQ.p, sizeof(double) * B * (T - 1) * nd, cudaMemcpyDeviceToHost, ctx->stream)); bytes += sizeof(double) * B * (T - 1) * nd; }
  if (out->cost) { CK(cudaMemcpyAsync(out->cost, ctx->outcost.p, sizeof(double) * B, cudaMemcpyDeviceToHost, ctx->stream)); bytes += sizeof(double) * B; }
  if (out->iters) { CK(cudaMemcpyAsync(out->iters, ctx->iters.p, sizeof(int) * B, cudaMemcpyDeviceToHost, ctx->stream)); bytes += sizeof(int) * B; }
  if (out->status) { CK(cudaMemcpyAsync(out->status, ctx->status.p, sizeof(int) * B, cudaMemcpyDeviceToHost, ctx->stream)); bytes += sizeof(int) * B; }
  CK(cudaEventRecord(e1, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  ctx->prof.d2h_ms = ms;
  ctx->prof.d2h_bytes = bytes;
  return GTO_OK;
}

extern "C" int gto_solve_batch(gto_ctx* ctx, const gto_batch_in* in, const gto_options* opts, gto_batch_out* out) {
  int rc = gto_upload_batch(ctx, in);
  if (rc) return rc;
  rc = gto_solve_resident(ctx, opts);
  if (rc) return rc;
  return gto_download_batch(ctx, out);
}

extern "C" int gto_result_device_ptr(gto_ctx* ctx, void** ptr, int64_t* nfloats) {
  if (!ctx) return GTO_ERR_INVALID;
  if (!ptr) return fail(ctx, GTO_ERR_INVALID, "null or out-of-range argument");
  if (!ctx->solved) return fail(ctx, GTO_ERR_STATE, "no solved batch");
  *ptr = (void*)ctx->result.p;
  if (nfloats) *nfloats = (int64_t)ctx->robot_h.nopt * ctx->T + 2;
  return GTO_OK;
}

extern "C" int gto_eval_batch(gto_ctx* ctx, const gto_batch_in* in, gto_eval_out* out) {
  if (!ctx) return GTO_ERR_INVALID;
  if (!out) return fail(ctx, GTO_ERR_INVALID, "null or out-of-range argument");
  int rc = gto_upload_batch(ctx, in);
  if (rc) return rc;
  const RobotDev& R = ctx->robot_h;
  const int B = ctx->B, T = ctx->T, n = R.nopt;
  const size_t nrows = (size_t)B * ctx->rows_per_problem * (n + 1);
  if (out->rows) CK(ctx->rows.ensure(nrows));
  ctx->Bchunk = B;
  StateParams sp;
  memset(&sp, 0, sizeof(sp));
  sp.robot = ctx->robot_d; sp.B = B; sp.T = T; sp.q_seed = ctx->q_seed.p; sp.q_trial = ctx->q_trial.p; sp.project = 0;
  const long long tot = (long long)B * T;
  k_init<<<(unsigned)((tot + 127) / 128), 128, 0, ctx->stream>>>(sp);
  CK(cudaGetLastError());
  CK(ctx->work_ctr.ensure(4));
  CK(ctx->stats.ensure(4));
  CK(cudaMemsetAsync(ctx->work_ctr.p, 0, sizeof(int) * 4, ctx->stream));
  CK(cudaMemsetAsync(ctx->stats.p, 0, sizeof(unsigned long long) * 4, ctx->stream));
  rc = launch_linearize(ctx, ctx->q_trial.p, nullptr, nullptr, B, 0, nullptr, out->rows ? ctx->rows.p : nullptr, 0, ctx->flags, ctx->work_ctr.p);
  if (rc) return rc;
  CK(cudaStreamSynchronize(ctx->stream));
  if (out->rows) CK(cudaMemcpy(out->rows, ctx->rows.p, nrows * sizeof(float), cudaMemcpyDeviceToHost));
  if (out->H) CK(cudaMemcpy(out->H, ctx->H.p, sizeof(float) * B * T * n * n, cudaMemcpyDeviceToHost));
  if (out->g) CK(cudaMemcpy(out->g, ctx->g.p, sizeof(double) * B * T * n, cudaMemcpyDeviceToHost));
  if (out->cost) CK(cudaMemcpy(out->cost, ctx->costp.p, sizeof(double) * B * T, cudaMemcpyDeviceToHost));
  return GTO_OK;
}

extern "C" int gto_get_profile(gto_ctx* ctx, gto_profile* prof) {
  if (!ctx) return GTO_ERR_INVALID;
  if (!prof) return fail(ctx, GTO_ERR_INVALID, "null or out-of-range argument");
  *prof = ctx->prof;
  return GTO_OK;
}

extern "C" int gto_plan_cost(gto_ctx* ctx, int32_t nplans, int32_t T, const double* plans, int32_t slot, const double base_position[3],
                             double* cost, double* dist) {
  if (!ctx) return GTO_ERR_INVALID;
  if (!plans || !cost || nplans < 1 || T < 1) return fail(ctx, GTO_ERR_INVALID, "null or out-of-range argument");
  if (!ctx->has_robot) return fail(ctx, GTO_ERR_STATE, "gto_set_robot has not been called");
  if (slot < 0 || slot >= MAX_FIELDS || !ctx->fields[slot].set) return fail(ctx, GTO_ERR_INVALID, "field slot not set");
  CK(cudaSetDevice(ctx->device));
  const int nd = ctx->robot_h.ndof;
  const size_t nq = (size_t)nplans * T * nd;
  std::vector<float> qf(nq);
  for (size_t i = 0; i < nq; ++i) qf[i] = (float)plans[i];
  float* dq = nullptr;
  double* dc = nullptr;
  CK(cudaMalloc((void**)&dq, nq * sizeof(float)));
  if (cudaMalloc((void**)&dc, nplans * sizeof(double)) != cudaSuccess) { cudaFree(dq); return fail(ctx, GTO_ERR_NOMEM, "plan cost buffer"); }
  cudaMemcpyAsync(dq, qf.data(), nq * sizeof(float), cudaMemcpyHostToDevice, ctx->stream);
  cudaMemsetAsync(dc, 0, nplans * sizeof(double), ctx->stream);
  FieldDev f;
  cudaMemcpyAsync(&f, ctx->fields_d + slot, sizeof(f), cudaMemcpyDeviceToHost, ctx->stream);
  cudaStreamSynchronize(ctx->stream);
  const float bx = base_position ? (float)base_position[0] : 0.f, by = base_position ? (float)base_position[1] : 0.f,
              bz = base_position ? (float)base_position[2] : 0.f;
  k_plan_cost<<<nplans * T, 128, 0, ctx->stream>>>(ctx->robot_d, ctx->pts3.p, ctx->pts3.p + ctx->npad, ctx->pts3.p + 2 * ctx->npad, dq, T, f, bx, by, bz, dc);
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaMemcpyAsync(cost, dc, nplans * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  cudaFree(dq);
  cudaFree(dc);
  if (e != cudaSuccess) return fail(ctx, GTO_ERR_CUDA, std::string("k_plan_cost: ") + cudaGetErrorString(e));
  if (dist)
    for (int i = 0; i < nplans; ++i) {
      double s = 0;
      for (int j = 0; j < nd; ++j) {
        const double d = plans[((size_t)i * T) * nd + j] - plans[((size_t)i * T + T - 1) * nd + j];
        s += d * d;
      }
      dist[i] = sqrt(s);
    }
  return GTO_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// Scene side: depth point cloud -> signed distance / cost at query points (DepthPointCloud)
// ------------------------------------------------------------------------------------------------------------------
static inline unsigned morton_spread10(unsigned v) {  // 10 bits -> every third bit
  v &= 1023u;
  v = (v | (v << 16)) & 0x030000FFu;
  v = (v | (v << 8)) & 0x0300F00Fu;
  v = (v | (v << 4)) & 0x030C30C3u;
  v = (v | (v << 2)) & 0x09249249u;
  return v;
}

extern "C" int gto_cloud_set(gto_ctx* ctx, const double* points, int64_t M) {
  if (!ctx) return GTO_ERR_INVALID;
  if (!points || M < 1) return fail(ctx, GTO_ERR_INVALID, "gto_cloud_set: needs at least one point");
  CK(cudaSetDevice(ctx->device));
  const size_t Mpad = ((size_t)M + CLOUD_TILE - 1) / CLOUD_TILE * CLOUD_TILE;
  // order the cloud along a Morton curve (10 bits per axis over its bounding box): consecutive points are neighbours in
  // space, so the 256-point tiles of the pruned query kernel have small bounding boxes
  double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
  for (int64_t i = 0; i < M; ++i)
    for (int a = 0; a < 3; ++a) { lo[a] = std::min(lo[a], points[3 * i + a]); hi[a] = std::max(hi[a], points[3 * i + a]); }
  std::vector<std::pair<unsigned, unsigned>> key((size_t)M);
  for (int64_t i = 0; i < M; ++i) {
    unsigned code = 0;
    for (int a = 0; a < 3; ++a) {
      const double ext = hi[a] - lo[a];
      const unsigned q = ext > 0 ? (unsigned)std::min(1023.0, (points[3 * i + a] - lo[a]) / ext * 1024.0) : 0u;
      code |= morton_spread10(q) << a;
    }
    key[(size_t)i] = std::make_pair(code, (unsigned)i);
  }
  std::sort(key.begin(), key.end());
  std::vector<float4> h(Mpad);
  for (size_t i = 0; i < (size_t)M; ++i) {
    const size_t j = key[i].second;
    h[i] = make_float4((float)points[3 * j], (float)points[3 * j + 1], (float)points[3 * j + 2], 0.f);
  }
  for (size_t i = (size_t)M; i < Mpad; ++i) h[i] = make_float4(1.0e18f, 1.0e18f, 1.0e18f, 0.f);  // never the nearest
  const size_t ntiles = ((size_t)M + CLOUD_PTILE - 1) / CLOUD_PTILE;
  std::vector<float> tb(ntiles * 6);
  for (size_t t = 0; t < ntiles; ++t) {
    float l3[3] = {3.0e38f, 3.0e38f, 3.0e38f}, h3[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
    for (size_t i = t * CLOUD_PTILE; i < std::min((size_t)M, (t + 1) * CLOUD_PTILE); ++i) {
      const float v[3] = {h[i].x, h[i].y, h[i].z};
      for (int a = 0; a < 3; ++a) { l3[a] = std::min(l3[a], v[a]); h3[a] = std::max(h3[a], v[a]); }
    }
    for (int a = 0; a < 3; ++a) { tb[6 * t + a] = l3[a]; tb[6 * t + 3 + a] = h3[a]; }
  }
  CK(ctx->cloud.ensure(Mpad));
  CK(ctx->cloud_tiles.ensure(tb.size()));
  CK(cudaMemcpy(ctx->cloud.p, h.data(), Mpad * sizeof(float4), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(ctx->cloud_tiles.p, tb.data(), tb.size() * sizeof(float), cudaMemcpyHostToDevice));
  ctx->cloud_n = M;
  return GTO_OK;
}

extern "C" int gto_cloud_query(gto_ctx* ctx, const double* query, int64_t N, const float* depth, int32_t H, int32_t W, const double K[9],
                               const double cam_inv[16], int32_t mode, double epsilon, double w_inside, float* out, double* kernel_ms) {
  if (!ctx) return GTO_ERR_INVALID;
  if (!query || !depth || !K || !cam_inv || !out || N < 1 || H < 1 || W < 1 || mode < 0 || mode > 2)
    return fail(ctx, GTO_ERR_INVALID, "gto_cloud_query: null argument, empty query / image or unknown mode");
  if (mode != 2 && ctx->cloud_n < 1) return fail(ctx, GTO_ERR_STATE, "gto_cloud_set has not been called");
  CK(cudaSetDevice(ctx->device));
  CK(ctx->cloud_q.ensure((size_t)N * 3));
  CK(ctx->cloud_out.ensure((size_t)N));
  CK(ctx->cloud_depth.ensure((size_t)H * W));
  CK(cudaMemcpyAsync(ctx->cloud_q.p, query, sizeof(double) * 3 * N, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(ctx->cloud_depth.p, depth, sizeof(float) * H * W, cudaMemcpyHostToDevice, ctx->stream));
  CloudParams p;
  memset(&p, 0, sizeof(p));
  p.pts = ctx->cloud.p;
  p.Mpad = (int)(((size_t)ctx->cloud_n + CLOUD_TILE - 1) / CLOUD_TILE * CLOUD_TILE);
  p.query = ctx->cloud_q.p; p.N = N; p.depth = ctx->cloud_depth.p; p.H = H; p.W = W;
  for (int i = 0; i < 9; ++i) p.K[i] = K[i];
  for (int i = 0; i < 12; ++i) p.RT[i] = cam_inv[i];
  p.mode = mode;
  p.eps = (float)epsilon; p.half_eps = (float)(epsilon / 2); p.two_eps = (float)(2 * epsilon); p.w_inside = (float)w_inside;
  p.out = ctx->cloud_out.p;
  EventPair ev;
  CK(ev.create());
  CK(cudaEventRecord(ev.a, ctx->stream));
  p.tiles = ctx->cloud_tiles.p;
  p.ntiles = (int)(((size_t)ctx->cloud_n + CLOUD_PTILE - 1) / CLOUD_PTILE);
  if (mode == 2) {  // visibility test only (is_outside)
    k_cloud_outside<<<(unsigned)((N + 255) / 256), 256, 0, ctx->stream>>>(p);
  } else if (getenv("GTO_CLOUD_BRUTE")) {  // A/B reference: every query against every point
    const long long per = (long long)CLOUD_THREADS * CLOUD_QPT;
    k_cloud_query<<<(unsigned)((N + per - 1) / per), CLOUD_THREADS, 0, ctx->stream>>>(p);
  } else {
    const long long nwarps = (N + CLOUD_WQ - 1) / CLOUD_WQ;
    k_cloud_query_pruned<<<(unsigned)((nwarps + 3) / 4), 128, 0, ctx->stream>>>(p);
  }
  CK(cudaGetLastError());
  CK(cudaEventRecord(ev.b, ctx->stream));
  CK(cudaMemcpyAsync(out, ctx->cloud_out.p, sizeof(float) * N, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  float ms = 0;
  cudaEventElapsedTime(&ms, ev.a, ev.b);
  if (kernel_ms) *kernel_ms = ms;
  return GTO_OK;
}

extern "C" int gto_cloud_backproject(gto_ctx* ctx, const float* depth, const uint8_t* target_mask, int32_t H, int32_t W, const double Kinv[9],
                                     const double cam_pose[16], double threshold, double* points, uint8_t* valid) {
  if (!ctx) return GTO_ERR_INVALID;
  if (!depth || !Kinv || !cam_pose || !points || !valid || H < 1 || W < 1) return fail(ctx, GTO_ERR_INVALID, "gto_cloud_backproject: null argument or empty image");
  CK(cudaSetDevice(ctx->device));
  const size_t n = (size_t)H * W;
  CK(ctx->cloud_depth.ensure(n));
  CK(ctx->cloud_q.ensure(n * 3));
  CK(ctx->cloud_mask.ensure(2 * n));
  CK(cudaMemcpyAsync(ctx->cloud_depth.p, depth, sizeof(float) * n, cudaMemcpyHostToDevice, ctx->stream));
  if (target_mask) CK(cudaMemcpyAsync(ctx->cloud_mask.p, target_mask, n, cudaMemcpyHostToDevice, ctx->stream));
  BackprojParams p;
  memset(&p, 0, sizeof(p));
  p.depth = ctx->cloud_depth.p; p.mask = target_mask ? ctx->cloud_mask.p : nullptr; p.H = H; p.W = W;
  for (int i = 0; i < 9; ++i) p.Kinv[i] = Kinv[i];
  for (int i = 0; i < 12; ++i) p.pose[i] = cam_pose[i];
  p.threshold = (float)threshold;
  p.points = ctx->cloud_q.p; p.valid = ctx->cloud_mask.p + n;
  k_cloud_backproject<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(p);
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(points, ctx->cloud_q.p, sizeof(double) * 3 * n, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(valid, ctx->cloud_mask.p + n, n, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return GTO_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// Mobile-base placement (BasePlanner): B problems x n goals in one launch
// ------------------------------------------------------------------------------------------------------------------
extern "C" int gto_base_place(gto_ctx* ctx, const gto_base_in* in, const gto_options* user_opts, gto_base_out* out, double* kernel_ms) {
  if (!ctx) return GTO_ERR_INVALID;
  if (!in || !out || !in->qc || !in->goal_tf || !out->Q || !out->y) return fail(ctx, GTO_ERR_INVALID, "null or out-of-range argument");
  if (!ctx->has_robot) return fail(ctx, GTO_ERR_STATE, "gto_set_robot has not been called");
  if (in->B < 1 || in->n_goals < 1 || in->n_goals > 32) return fail(ctx, GTO_ERR_INVALID, "base placement: need B >= 1 and 1 <= n_goals <= 32");
  if (in->occupancy && (in->occ_dims[0] < 1 || in->occ_dims[1] < 1 || !(in->occ_resolution > 0)))
    return fail(ctx, GTO_ERR_INVALID, "base placement: bad occupancy grid");
  CK(cudaSetDevice(ctx->device));
  gto_options o;
  if (user_opts) o = *user_opts; else gto_default_options(&o);
  const RobotDev& h = ctx->robot_h;
  const int B = in->B, n = in->n_goals, nd = h.ndof, nopt = h.nopt, P = h.npoints;
  // one float64 allocation: qc | goals | wp | Qx | y | cost | collision
  const size_t o_qc = 0, o_goal = o_qc + (size_t)nd, o_wp = o_goal + (size_t)B * n * 12, o_Qx = o_wp + (size_t)P * 3,
               o_y = o_Qx + (size_t)B * n * nopt, o_cost = o_y + (size_t)B * 3, o_coll = o_cost + (size_t)B, total = o_coll + (size_t)B;
  CK(ctx->base_d.ensure(total));
  CK(ctx->base_i.ensure((size_t)2 * B));
  double* d = ctx->base_d.p;
  CK(cudaMemcpyAsync(d + o_qc, in->qc, sizeof(double) * nd, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(d + o_goal, in->goal_tf, sizeof(double) * (size_t)B * n * 12, cudaMemcpyHostToDevice, ctx->stream));
  BaseParams p;
  memset(&p, 0, sizeof(p));
  p.robot = ctx->robot_d;
  p.B = B; p.n = n;
  p.qc = d + o_qc; p.goal = d + o_goal; p.w_effort = in->w_effort;
  {  // movable joints from the root to the gripper link
    int tmp[GTO_MAX_MOV], c = 0;
    for (int j = h.grip_mov; j >= 0; j = h.mov_parent[j]) tmp[c++] = j;
    p.nchain = c;
    for (int i = 0; i < c; ++i) p.chain[i] = tmp[c - 1 - i];
  }
  for (int e = 0; e < 16; ++e) p.mom[e] = ctx->grip_mom[e];
  p.max_iter = o.max_iter; p.tol_step = o.tol_step; p.tol_grad = o.tol_grad; p.lambda0 = o.lambda0; p.lambda_min = o.lambda_min;
  p.lambda_max = o.lambda_max; p.eta = o.eta; p.bound_eps = o.bound_eps;
  if (in->occupancy) {
    const size_t cells = (size_t)in->occ_dims[0] * in->occ_dims[1];
    CK(ctx->base_occ.ensure(cells));
    CK(cudaMemcpyAsync(ctx->base_occ.p, in->occupancy, sizeof(float) * cells, cudaMemcpyHostToDevice, ctx->stream));
    p.occ = ctx->base_occ.p; p.onx = in->occ_dims[0]; p.ony = in->occ_dims[1];
    p.oox = in->occ_origin[0]; p.ooy = in->occ_origin[1]; p.ores = in->occ_resolution;
  }
  p.wp = d + o_wp; p.npoints = P;
  p.Qx = d + o_Qx; p.y = d + o_y; p.cost = d + o_cost; p.collision = d + o_coll;
  p.iters = ctx->base_i.p; p.status = ctx->base_i.p + B;
  EventPair ev;
  CK(ev.create());
  CK(cudaEventRecord(ev.a, ctx->stream));
  if (in->occupancy) k_base_points<<<1, 256, 0, ctx->stream>>>(ctx->robot_d, ctx->pts3.p, ctx->pts3.p + ctx->npad, ctx->pts3.p + 2 * ctx->npad, p.qc, d + o_wp);
  const bool v1 = getenv("GTO_BASE_V1") != nullptr || nopt > 12;  // local-memory kernel: A/B reference, and robots with > 12 optimised joints
  if (v1) {
    if (nopt <= 8) k_base_place<8><<<B, 32, 0, ctx->stream>>>(p);
    else k_base_place<16><<<B, 32, 0, ctx->stream>>>(p);
  } else {
    const int gpw = 32 / n;
    const unsigned grid = (unsigned)((B + gpw - 1) / gpw);
#define GTO_BASE_LAUNCH(NP_, NOPT_)                                                                              \
  do {                                                                                                         \
    const size_t smem = sizeof(double) * 32 * base_sm_doubles_per_lane<NP_>();                                 \
    CK(cudaFuncSetAttribute(k_base_place_sm<NP_, NOPT_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    k_base_place_sm<NP_, NOPT_><<<grid, 32, smem, ctx->stream>>>(p);                                            \
  } while (0)
    switch (nopt) {  // the shipped robots get fully unrolled instances (Panda / Fetch 7, Fetch-8, Fetch-10)
      case 7: GTO_BASE_LAUNCH(8, 7); break;
      case 8: GTO_BASE_LAUNCH(8, 8); break;
      case 10: GTO_BASE_LAUNCH(12, 10); break;
      default:
        if (nopt <= 8) GTO_BASE_LAUNCH(8, 0);
        else GTO_BASE_LAUNCH(12, 0);
    }
#undef GTO_BASE_LAUNCH
  }
  CK(cudaGetLastError());
  CK(cudaEventRecord(ev.b, ctx->stream));
  // results: optimised rows come back packed, the parameter joints are re-inflated from qc on the host (optas/solver.py:126-159)
  std::vector<double> qx((size_t)B * n * nopt);
  std::vector<int> is((size_t)2 * B);
  CK(cudaMemcpyAsync(qx.data(), d + o_Qx, sizeof(double) * qx.size(), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(out->y, d + o_y, sizeof(double) * (size_t)B * 3, cudaMemcpyDeviceToHost, ctx->stream));
  if (out->cost) CK(cudaMemcpyAsync(out->cost, d + o_cost, sizeof(double) * B, cudaMemcpyDeviceToHost, ctx->stream));
  if (out->collision) CK(cudaMemcpyAsync(out->collision, d + o_coll, sizeof(double) * B, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(is.data(), ctx->base_i.p, sizeof(int) * is.size(), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  for (size_t bi = 0; bi < (size_t)B * n; ++bi) {
    double* q = out->Q + bi * nd;
    for (int j = 0; j < nd; ++j) q[j] = in->qc[j];
    for (int k = 0; k < nopt; ++k) q[h.opt_qidx[k]] = qx[bi * nopt + k];
  }
  for (int b = 0; b < B; ++b) {
    if (out->iters) out->iters[b] = is[b];
    if (out->status) out->status[b] = is[B + b];
  }
  float ms = 0;
  cudaEventElapsedTime(&ms, ev.a, ev.b);
  if (kernel_ms) *kernel_ms = ms;
  return GTO_OK;
}
