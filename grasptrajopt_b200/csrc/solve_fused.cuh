// solve_fused.cuh -- k_solve_fused: the whole solver loop of a problem inside ONE persistent CTA.
// Included by gto_b200.cu after lin_cull.cuh and step_cr.cuh (uses item_fk_body, cull_body, step_body).
//
// The launch-per-iteration structure (k_item_fk -> k_linearize_cull -> k_step_cr, ~100 x 3 dependent launches per solve,
// every problem waiting at every launch boundary for the slowest one) is latency bound: a C2 solve is 101 iterations of
// ~74 us of which the HBM traffic needs ~5 us.  Problems are independent (one reference plan() call each,
// gto/gto_planner.py:145-182), so here a CTA takes a problem from a global queue and runs
//     FK records of its knots -> linearisation (TMA bricks, Jacobian rows to HBM, J^T J / J^T r) -> LM step
// in a loop until the problem has converged, then takes the next problem (continuous batching: a batch of any size keeps
// every SM busy, and a slow problem delays nobody).  Measured on B200 (C2): 27 ms per 256-problem solve against 10.5 ms of
// the launch path -- the linearisation of ONE problem's 28 knots inside one CTA is paced by the per-item latency chain (record
// copy -> TMA bricks -> per-item reduction, ~6 us each), which the launch path hides by spreading the items over all SMs.
// Selected with gto_configure("fused", 1); kept as the structure to build on (records in shared memory, bricks prefetched
// for all knots) and because it needs no Jacobian-row chunking for large batches.  No launches, no host polling, no active lists; the three phases reuse
// the bodies of the per-iteration kernels bit for bit, so both paths give identical results.
//   * the Jacobian-row buffer is indexed by CTA slot (gridDim.x slots), not by problem: the rows of one linearisation are
//     materialised once (coalesced 128-bit stores / bulk zero stores) and overwritten by the next one of that slot;
//   * shared memory: a persistent copy of the robot table + one region shared by the three phases (they never overlap);
//   * the per-problem Gauss-Newton blocks (both the accepted and the trial buffer) stay in global memory (L2 resident).
#pragma once

struct FusedParams {
  CullParams cull;     // linearisation (recs = [gridDim.x][T] item records, rows = [gridDim.x][rows_per_problem][nopt+1])
  StepParams step;
  int B;               // problems of the batch
  int* queue;          // next problem to start (zero before the launch)
  unsigned long long* phase_ns;  // [4] summed over CTAs: ns spent in FK / linearise / step, and CTA-iterations (profile; NULL: off)
};

__host__ __device__ inline size_t fused_fk_smem_bytes(int nmov, int nthreads) { return (size_t)(nthreads / 16) * 2 * nmov * 12 * sizeof(double); }
__host__ __device__ inline size_t fused_robot_bytes() { return (sizeof(RobotDev) + 127) & ~(size_t)127; }

// NPC / NOPT_CT: tile width and joint count of the linearisation (cull_body); NPS / EXACT: block order of the step (step_body)
template <int NPC, int NOPT_CT, int NPS, bool EXACT>
__global__ void __launch_bounds__((CULL_MAX_CONS + 2) * 32, 2) k_solve_fused(const __grid_constant__ FusedParams P) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ int s_next;
  const int tid = threadIdx.x;
  RobotDev& R = *reinterpret_cast<RobotDev*>(smem_raw);
  unsigned char* phase = smem_raw + fused_robot_bytes();
  {
    const unsigned long long* src = reinterpret_cast<const unsigned long long*>(P.cull.lin.robot);
    unsigned long long* dst = reinterpret_cast<unsigned long long*>(smem_raw);
    for (int i = tid; i < (int)(sizeof(RobotDev) / 8); i += blockDim.x) dst[i] = __ldg(src + i);
  }
  const LinParams& lp = P.cull.lin;
  const int T = lp.T, slot = blockIdx.x;
  CullCtx* recs = P.cull.recs + (size_t)slot * T;
  const int hl = tid & 15, grp = tid >> 4, ngrp = blockDim.x >> 4, hshift = tid & 16;
  unsigned long long t_fk = 0, t_lin = 0, t_step = 0, n_it = 0;
  const bool prof = P.phase_ns != nullptr && tid == 0;

  for (;;) {
    __syncthreads();  // (also: the robot table is in place; the previous problem's last phase is over)
    if (tid == 0) s_next = atomicAdd(P.queue, 1);
    __syncthreads();
    const int b = s_next;
    if (b >= P.B) break;
    for (int it = 0;; ++it) {
      const int t_lo = it == 0 ? 0 : 2, nitems = T - t_lo;
      unsigned long long c0 = prof ? gto_globaltimer() : 0ull;
      // ---- phase 1: item records (float64 chain FK, brick placement, culling test), 16 lanes per knot ----
      {
        const int obuf = 1 - lp.bufsel[b];
        double* A = reinterpret_cast<double*>(phase) + (size_t)grp * 2 * R.nmov * 12;
        double* Tm = A + (size_t)R.nmov * 12;
        for (int i0 = 0; i0 < nitems; i0 += ngrp) {
          if (((i0 + grp) & ~1) < nitems) {  // the two halves of a warp run in lock step
            const bool valid = (i0 + grp) < nitems;
            const int item = valid ? i0 + grp : nitems - 1;
            const int t = t_lo + item;
            item_fk_body(P.cull, recs, R, item, valid, b, t, obuf, lp.q + ((long long)b * T + t) * R.ndof, A, Tm, hl, hshift);
          }
        }
        asm volatile("fence.proxy.async;" ::: "memory");  // the records are read back by bulk copies (async proxy)
      }
      __syncthreads();
      unsigned long long c1 = prof ? gto_globaltimer() : 0ull;
      // ---- phase 2: linearisation of the problem's knots ----
      cull_body<NPC, NOPT_CT, true>(P.cull, phase, R, recs, nitems, slot);
      __syncthreads();
      if (tid == 0) {  // the barriers' storage is reused by the step phase
        CullShared& S = *reinterpret_cast<CullShared*>(phase);
        for (int s = 0; s < P.cull.nslot; ++s) {
          asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(smem_u32(&S.slot_full[s])) : "memory");
          asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(smem_u32(&S.slot_empty[s])) : "memory");
        }
        for (int c = 0; c < CULL_NCTX; ++c) {
          asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(smem_u32(&S.ctx_full[c])) : "memory");
          asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(smem_u32(&S.ctx_empty[c])) : "memory");
        }
        for (int z = 0; z < CULL_ZQ; ++z) {
          asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(smem_u32(&S.zq_full[z])) : "memory");
          asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(smem_u32(&S.zq_empty[z])) : "memory");
        }
        asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(smem_u32(&S.pts_full)) : "memory");
      }
      __syncthreads();
      unsigned long long c2 = prof ? gto_globaltimer() : 0ull;
      // ---- phase 3: LM step ----
      const bool active = step_body<NPS, EXACT, true>(P.step, phase, R, b, it);
      __syncthreads();
      if (prof) {
        const unsigned long long c3 = gto_globaltimer();
        t_fk += c1 - c0; t_lin += c2 - c1; t_step += c3 - c2; n_it += 1;
      }
      if (!active) break;
    }
  }
  if (prof) {
    atomicAdd(P.phase_ns + 0, t_fk);
    atomicAdd(P.phase_ns + 1, t_lin);
    atomicAdd(P.phase_ns + 2, t_step);
    atomicAdd(P.phase_ns + 3, n_it);
  }
}
