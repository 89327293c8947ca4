// step_cr.cuh -- k_step_cr: Levenberg-Marquardt bookkeeping + damped projected Gauss-Newton step, one CTA per problem.
// Included by gto_b200.cu after k_step (shares StepParams, warp_sum / warp_max / shfl_d).
//
// Same mathematics as k_step (mirrors oracle/gto_oracle.py solve_lm / lm_step, float64), but the SPD block-tridiagonal
// system  (H_t + a2 c_t I)(1 + lambda) on the diagonal, -a2 I off it  (the velocity term gto/gto_planner.py:134-135
// couples neighbouring knots) is solved by *block cyclic reduction* instead of the sequential block Thomas sweep:
// at stride s every block i = s (mod 2s) is eliminated at once -- one warp per block: Gauss-Jordan inverse of the 7x7..16x16
// diagonal block in registers (pivot rows travel by warp shuffle), W_L = D^-1 L, W_U = D^-1 U, w = D^-1 b -- and its two
// neighbours i -+ s absorb the Schur complement.  log2(T) levels instead of T-2 dependent block steps: the launch is
// latency bound (a few thousand flops per problem), so the depth of the dependency chain is what sets its duration.
// Elimination in any symmetric order is stable for an SPD matrix; a non-positive pivot means the damped matrix is not
// positive definite and the factorisation is retried with more damping, exactly as in k_step.
//
// Gradient bundle (gto_options.bundle, oracle lm_step / solve_lm): the trilinear field makes the objective piecewise smooth,
// and a minimiser usually lies on a gradient jump (cell face) where a one-sided quadratic model keeps predicting a descent that
// the other side takes back.  The (cost, gradient) of up to `bundle` points that were evaluated but are not stood on --
// rejected trial points, iterates that were left -- are kept per problem as cutting planes piece_k(s) = e_k + g_k.s, and the
// step minimises  max_k piece_k(s) + s'(H + lambda D)s/2 : the SAME factorisation is applied to bundle+1 right-hand sides
// (-g_0 .. -g_K), a (K+1)-variable dual QP gives the convex weights, the step is the weighted sum.  Both gradients are already
// there: the linearise kernel writes J^T r of the accepted AND of the trial point (double buffer).
#pragma once

#define STEP_CR_THREADS 256
#define STEP_RED2 (16 * (GTO_BUNDLE_MAX + GTO_BUNDLE_MAX * GTO_BUNDLE_MAX))  // scratch of cta_sum_n: [values][<= 16 warps]

__host__ __device__ inline size_t step_cr_smem_bytes(int T, int n, int bundle) {
  const size_t m = (size_t)(T - 2), nn = (size_t)n * n, nr = (size_t)bundle + 1;
  const size_t d = (size_t)2 * T * n + (5 + 2 * nr) * m * n + 3 * m * nn + 64 + STEP_RED2 + 8;
  return ((d * sizeof(double) + 2 * (m * nn) * sizeof(float) + (m + 4) * sizeof(unsigned)) + 15) & ~(size_t)15;
}

// deterministic block reductions (fixed tree): every thread of the CTA must call them
__device__ __forceinline__ double cta_sum(double v, double* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  double s = 0.0;
  for (int w = 0; w < nw; ++w) s += red[w];
  return s;
}
__device__ __forceinline__ void cta_sum3(double& a, double& b, double& c, double* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  a = warp_sum(a); b = warp_sum(b); c = warp_sum(c);
  __syncthreads();
  if (lane == 0) { red[warp] = a; red[16 + warp] = b; red[32 + warp] = c; }
  __syncthreads();
  a = 0.0; b = 0.0; c = 0.0;
  for (int w = 0; w < nw; ++w) { a += red[w]; b += red[16 + w]; c += red[32 + w]; }
}
__device__ __forceinline__ double cta_max(double v, double* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  double s = red[0];
  for (int w = 1; w < nw; ++w) s = fmax(s, red[w]);
  return s;
}

// Up to NV reductions at once (fixed tree): value i takes part when bit i of `mask` is set (uniform over the CTA); bit i of
// `maxmask` selects max instead of sum.  red2: [NV][16]
template <int NV>
__device__ __forceinline__ void cta_reduce_n(double (&v)[NV], unsigned mask, unsigned maxmask, double* red2) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
  for (int i = 0; i < NV; ++i)
    if ((mask >> i) & 1u) v[i] = ((maxmask >> i) & 1u) ? warp_max(v[i]) : warp_sum(v[i]);
  __syncthreads();
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i)
      if ((mask >> i) & 1u) red2[i * 16 + warp] = v[i];
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < NV; ++i)
    if ((mask >> i) & 1u) {
      double s = red2[i * 16];
      if ((maxmask >> i) & 1u)
        for (int w = 1; w < nw; ++w) s = fmax(s, red2[i * 16 + w]);
      else
        for (int w = 1; w < nw; ++w) s += red2[i * 16 + w];
      v[i] = s;
    }
}

// Dual of the bundle model (oracle bundle_dual / _bundle_dual, same loop): maximise b.theta + theta' M theta / 2 over
// theta_1..K >= 0, sum <= 1 by pairwise exchange; theta_0 = 1 - sum.  Row / column 0 of M and b_0 are zero, entries beyond K
// are zero.  Every array index is a compile-time constant (selects instead of M[ib][jb]) so that all of it stays in registers.
__device__ __forceinline__ void bundle_dual(int K, const double (&bq)[GTO_BUNDLE_MAX + 1], const double (&M)[GTO_BUNDLE_MAX + 1][GTO_BUNDLE_MAX + 1],
                                            double (&theta)[GTO_BUNDLE_MAX + 1]) {
  constexpr int KM = GTO_BUNDLE_MAX;
#pragma unroll
  for (int k = 0; k <= KM; ++k) theta[k] = (k == 0) ? 1.0 : 0.0;
  // the dual gradients of the pieces with weight all vanish at an interior optimum: the stopping tolerance is relative to the
  // largest |b_k|, not to the gradients themselves (which would never pass it and always run into the iteration cap)
  double scale = 0.0;
#pragma unroll
  for (int k = 1; k <= KM; ++k) scale = fmax(scale, fabs(bq[k]));
  for (int iter = 0; iter < 12; ++iter) {
    double G[KM + 1];
#pragma unroll
    for (int k = 0; k <= KM; ++k) {
      double g = bq[k];
#pragma unroll
      for (int j = 1; j <= KM; ++j) g += M[k][j] * theta[j];
      G[k] = g;
    }
    int ib = 0, jb = -1;
    double Gi = G[0], Gj = 0.0, thj = 0.0;
#pragma unroll
    for (int k = 1; k <= KM; ++k)
      if (k <= K && G[k] > Gi) { ib = k; Gi = G[k]; }
#pragma unroll
    for (int k = 0; k <= KM; ++k)
      if (theta[k] > 0.0 && (jb < 0 || G[k] < Gj)) { jb = k; Gj = G[k]; thj = theta[k]; }
    if (jb < 0 || ib == jb || Gi - Gj <= 1e-13 * scale) break;
    double Mii = 0.0, Mij = 0.0, Mjj = 0.0;
#pragma unroll
    for (int k = 0; k <= KM; ++k)
#pragma unroll
      for (int j = 0; j <= KM; ++j) {
        const double mv = M[k][j];
        if (k == ib && j == ib) Mii = mv;
        if (k == ib && j == jb) Mij = mv;
        if (k == jb && j == jb) Mjj = mv;
      }
    const double curv = -(Mii - 2.0 * Mij + Mjj);
    double delta = curv > 0.0 ? (Gi - Gj) / curv : thj;
    if (delta > thj) delta = thj;
#pragma unroll
    for (int k = 0; k <= KM; ++k) {
      if (k == ib) theta[k] += delta;
      if (k == jb) {
        theta[k] -= delta;
        if (theta[k] < 1e-15) theta[k] = 0.0;
      }
    }
  }
}

// In-register Gauss-Jordan inverse of an SPD block: lane r < NP owns row r.  Returns false (uniformly) on a
// non-positive pivot.
template <int NP>
__device__ __forceinline__ bool gj_inverse(double (&row)[NP], int r) {
  bool ok = true;
#pragma unroll
  for (int k = 0; k < NP; ++k) {
    const double piv = shfl_d(row[k], k);
    if (!(piv > 0.0)) ok = false;  // uniform: every lane sees the same pivot
    float ipf;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(ipf) : "f"((float)piv));  // seed; two Newton steps in float64 follow
    double ip = (double)ipf;
    ip = ip * (2.0 - piv * ip);
    ip = ip * (2.0 - piv * ip);
    const bool isk = (r == k);
    const double mult = -row[k] * ip;
    const double coef = isk ? ip - 1.0 : mult;
#pragma unroll
    for (int c = 0; c < NP; ++c) {
      if (c == k) continue;
      row[c] = fma(coef, shfl_d(row[c], k), row[c]);
    }
    row[k] = isk ? ip : mult;
  }
  return ok;
}

// step_body: one LM step of problem b (iteration counter `it`): judge the trial point of the previous call, update the
// damping, test convergence, solve for the next trial point.  Returns true (uniformly over the CTA) when the problem stays
// active.  Two callers: k_step_cr (one launch per iteration, a CTA per active problem; FUSED = false) and k_solve_fused
// (solve_fused.cuh: the CTA that owns the problem loops over the iterations; FUSED = true).
// step_smem: step_cr_smem_bytes(T, n) bytes, 16-byte aligned; every thread of the CTA must call it.
template <int NP, bool EXACT, bool FUSED>
__device__ __forceinline__ bool step_body(const StepParams& p, unsigned char* step_smem, const RobotDev& R, const int b, const int it) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, NT = blockDim.x, NW = NT >> 5;
  int dbg_i = 0;
#define STEP_MARK() do { if (p.dbg && p.iter == p.dbg_iter && (int)blockIdx.x == p.dbg_cta && tid == 0 && dbg_i < 63) p.dbg[dbg_i] = clock64(); ++dbg_i; } while (0)
  // Programmatic dependent launch: the per-problem solver state read below (active list, damping, accepted / trial point) was
  // written by the PREVIOUS step kernel, which completed before the linearise kernel ahead of us even started -- only the
  // Gauss-Newton blocks and costs need pdl_wait(), so the dependent round trips for the state overlap the linearise kernel.
  STEP_MARK();
  const int n = EXACT ? NP : R.nopt, T = p.T, m = T - 2, nn = n * n;
  const double a2 = p.w_vel / (p.dt * p.dt);
  const int KB = p.bundle, mn = m * n;
  double* X = reinterpret_cast<double*>(step_smem);  // [T][n] accepted point
  double* gt = X + (size_t)T * n;                    // [m][n] gradient
  double* dd = gt + (size_t)m * n;                   // [m][n] clipped step
  double* xs = dd + (size_t)m * n;                   // [KB+1][m][n] solutions of the linear system (0: LM step -> combined step)
  double* bb = xs + (size_t)(KB + 1) * mn;           // [KB+1][m][n] right-hand sides -> D^-1 b of eliminated blocks
  double* Dm = bb + (size_t)(KB + 1) * mn;           // [m][n*n] diagonal blocks -> their inverses
  double* Lm = Dm + (size_t)m * nn;                  // [m][n*n] coupling to block i - s -> D^-1 L
  double* Um = Lm + (size_t)m * nn;                  // [m][n*n] coupling to block i + s -> D^-1 U
  double* red = Um + (size_t)m * nn;                 // [64] reduction scratch
  double* red2 = red + 64;                           // [STEP_RED2] scratch of cta_sum_n
  double* thS = red2 + STEP_RED2;                    // [8] convex weights of the bundle pieces
  double* X2 = thS + 8;                              // [T][n] the other of (accepted, trial) point while the decision is open
  double* dfix = X2 + (size_t)T * n;                 // [m][n] prescribed step of the variables held at a joint limit
  double* gS = dfix + (size_t)m * n;                 // [2][m][n] J^T r of both buffers (knots 2..T-1)
  float* HS = reinterpret_cast<float*>(gS + (size_t)2 * m * n);  // [2][m][n*n] Gauss-Newton blocks of both buffers
  unsigned* fm = reinterpret_cast<unsigned*>(HS + (size_t)2 * m * nn);  // [m] bit k: variable k of knot i+2 is held at a bound
  int* sflag = reinterpret_cast<int*>(fm + m);           // [0] factorisation failed, [1] slot in the next active list

  double* Xc = p.Qc + (long long)b * T * n;
  double* Xt = p.Qt + (long long)b * T * n;
  // ---- everything that only depends on b is requested in one batch: the launch is latency bound, and every dependent
  //      round trip to L2 / HBM costs the better part of a microsecond ----
  int cur = p.bufsel[b];
  const int cur0 = cur;
  double* gBb = p.gB + (long long)b * GTO_BUNDLE_MAX * mn;    // [KB][m][n] half gradients at the bundle points y_k
  double* dyBb = p.dyB + (long long)b * GTO_BUNDLE_MAX * mn;  // [KB][m][n] y_k - x
  double* FBb = p.FB + (long long)b * GTO_BUNDLE_MAX;
  int nb = (KB > 0 && it > 0) ? p.nbund[b] : 0;
  double eB[GTO_BUNDLE_MAX + 1];  // value of piece k at the standing point (half-cost units, <= 0); uniform over the CTA
#pragma unroll
  for (int k = 0; k <= GTO_BUNDLE_MAX; ++k) eB[k] = 0.0;
  double lam = p.lam[b], nu = p.nu[b];
  const double Fcur = p.F[b], Fpcur = p.Fp[b], pred = p.pred[b], step = p.stepn[b];
  const int tri = 1 - cur;
  bool accepted = false;
  double s0 = 0.0, s1 = 0.0, sv = 0.0;
  for (int i = tid; i < T * n; i += NT) {
    const double xt = Xt[i];
    X[i] = xt;         // trial point
    X2[i] = Xc[i];     // accepted point
    if (i >= n) {
      const double d = xt - Xt[i - n];
      sv += d * d;
    }
  }
  // Bundle sums that do not depend on the decision about the trial point (the bundle was written by the previous step kernel):
  // [k] g_k.dy_k, [KM+k] g_k.(trial - accepted), [2KM+k] |dy_k|_inf, [3KM+k] |dy_k - (trial - accepted)|_inf -- ahead of pdl_wait().
  constexpr int KM = GTO_BUNDLE_MAX;
  double pre[4 * KM], FBv[KM];
#pragma unroll
  for (int k = 0; k < KM; ++k) {
    pre[k] = 0.0; pre[KM + k] = 0.0; pre[2 * KM + k] = 0.0; pre[3 * KM + k] = 0.0;
    FBv[k] = (k < nb) ? FBb[k] : 0.0;
  }
  if (nb > 0) {
    for (int idx = tid; idx < mn; idx += NT) {
      const double dprev = Xt[2 * n + idx] - Xc[2 * n + idx];
#pragma unroll
      for (int k = 0; k < KM; ++k)
        if (k < nb) {
          const double gk = gBb[(size_t)k * mn + idx], dy = dyBb[(size_t)k * mn + idx];
          pre[k] = fma(gk, dy, pre[k]);
          pre[KM + k] = fma(gk, dprev, pre[KM + k]);
          pre[2 * KM + k] = fmax(pre[2 * KM + k], fabs(dy));
          pre[3 * KM + k] = fmax(pre[3 * KM + k], fabs(dy - dprev));
        }
    }
    const unsigned used = (1u << nb) - 1u;
    cta_reduce_n<4 * KM>(pre, used | (used << KM) | (used << (2 * KM)) | (used << (3 * KM)), (used << (2 * KM)) | (used << (3 * KM)), red2);
  }
  if (!FUSED) {
    pdl_wait();     // from here on: results of the linearise kernel before us
    pdl_trigger();  // the successor may be scheduled (it blocks in its own pdl_wait until we are done)
    stamp_begin(p.ts);
  }
  for (int buf = 0; buf < 2; ++buf) {  // Gauss-Newton blocks of both buffers (which one is "accepted" is decided below)
    const float* Hg = p.H + buf * p.buf_stride_H + (long long)b * T * nn + 2 * nn;
    const double* gg = p.g + buf * p.buf_stride_g + (long long)b * T * n + 2 * n;
    for (int i = tid; i < m * nn; i += NT)
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(HS + (size_t)buf * m * nn + i)), "l"(Hg + i) : "memory");
    for (int i = tid; i < m * n; i += NT)
      asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(gS + (size_t)buf * m * n + i)), "l"(gg + i) : "memory");
  }
  asm volatile("cp.async.commit_group;" ::: "memory");

  // ---------------- evaluate the trial point produced by the previous call ----------------
  int slot = -1;    // bundle slot that receives the new cutting plane (-1: none)
  double Ft0 = 0.0;  // cost of the trial point
  {
    for (int t = tid; t < T; t += NT) {
      s0 += p.costp[(long long)b * T + t];
      s1 += p.costp[p.buf_stride_c + (long long)b * T + t];
    }
    STEP_MARK();  // 1: loads issued
    cta_sum3(s0, s1, sv, red);
    STEP_MARK();  // 2: trial cost known
    double* ct = p.costp + tri * p.buf_stride_c + (long long)b * T;
    const double Fp_t = tri ? s1 : s0;
    const double Ft = Fp_t + a2 * sv;
    Ft0 = Ft;
    int done = -1;  // -1: keep running, otherwise final status
    if (!isfinite(Ft)) {
      done = GTO_STATUS_NAN;
      if (it == 0 && tid == 0) p.F[b] = CUDART_INF;  // no accepted point yet: the reported cost must not be the initial 0
    } else if (it == 0) {  // initial point: accept unconditionally
      // knots 0 and 1 never move and are linearised only once: keep their cost in both buffers
      if (tid < 2) {
        const double c01 = ct[tid];
        for (int buf = 0; buf < 2; ++buf) p.costp[buf * p.buf_stride_c + (long long)b * T + tid] = c01;
      }
      cur = tri;
      accepted = true;
      if (tid == 0) { p.bufsel[b] = cur; p.F[b] = Ft; p.Fp[b] = Fp_t; }
    } else {
      const double ared = 0.5 * (Fcur - Ft);
      const double noise = p.noise_rel * fmax(Fpcur, Fp_t);
      if (pred > 0.0 && ared + noise >= p.eta * pred) {
        const double rho = ared / pred;
        cur = tri;
        accepted = true;
        if (tid == 0) { p.bufsel[b] = cur; p.F[b] = Ft; p.Fp[b] = Fp_t; }
        const double lam_used = lam;
        // gain ratio clamped to [0, 1]: a step accepted only thanks to the noise allowance can have rho << 0, and Nielsen's
        // cubic would then multiply the damping by hundreds in one step (float32 cost noise near convergence)
        const double w = 2.0 * fmin(fmax(rho, 0.0), 1.0) - 1.0;
        lam = fmax(p.lambda_min, lam * fmax(1.0 / 3.0, 1.0 - w * w * w));
        nu = 2.0;
        // a small step only certifies a stationary point when it was (nearly) the undamped Gauss-Newton step; under heavy
        // damping it means the iterate rests on a gradient jump of the trilinear field (GTO_STATUS_SLOW, not converged)
        if (step <= p.tol_step) done = (lam_used <= p.lambda_conv) ? GTO_STATUS_CONVERGED : GTO_STATUS_SLOW;
        else if (lam_used >= p.lambda_slow && ared <= p.ftol * Fcur) done = GTO_STATUS_SLOW;
      } else {
        if (pred <= 0.0 && step <= p.tol_step) {
          done = (lam <= p.lambda_conv) ? GTO_STATUS_CONVERGED : GTO_STATUS_SLOW;
        } else {
          lam = fmin(p.lambda_max, fmax(lam * nu, p.lambda_reject));
          nu *= 2.0;
          if (lam >= p.lambda_max) done = GTO_STATUS_STALLED;
        }
      }
      if (done < 0 && p.slow_window > 0) {  // windowed progress test on the accepted cost
        const double Fnow = accepted ? Ft : Fcur;
        double* hist = p.Fhist + (long long)b * 16;
        if (it >= p.slow_window) {
          const double Fold = hist[(it - p.slow_window) & 15];
          if (Fold - Fnow <= p.slow_ftol * Fnow) done = GTO_STATUS_SLOW;
        }
        __syncthreads();
        if (tid == 0) hist[it & 15] = Fnow;
      }
    }
    if (it == 0 && done < 0 && p.slow_window > 0 && tid == 0) p.Fhist[(long long)b * 16] = Ft;
    if (done < 0 && it >= p.max_iter) done = GTO_STATUS_MAX_ITER;
    // ---- bundle update, part 1 (oracle solve_lm): kept pieces are re-based to the new standing point; the point we do not stand on
    //      after this decision becomes a cutting plane (its gradient is formed in the gradient loop below); a full bundle replaces
    //      its least active piece ----
    if (done < 0 && it > 0 && KB > 0) {
      const double Fnew = accepted ? Ft : Fcur;
      slot = nb;
      double worst = 0.0;
#pragma unroll
      for (int k = 0; k < KM; ++k)
        if (k < nb) {  // a piece further than bundle_radius from the standing point is not used (and is the first to be replaced)
          const double s1 = accepted ? pre[k] - pre[KM + k] : pre[k], rk = accepted ? pre[3 * KM + k] : pre[2 * KM + k];
          const double e = rk > p.bundle_radius ? -1e300 : -fabs(0.5 * (FBv[k] - Fnew) - s1);
          eB[k + 1] = e;
          if (nb >= KB && (k == 0 || e < worst)) { worst = e; slot = k; }
        }
      if (accepted) {
        for (int idx = tid; idx < mn; idx += NT) {
          const double dprev = X[2 * n + idx] - X2[2 * n + idx];
#pragma unroll
          for (int k = 0; k < KM; ++k)
            if (k < nb && k != slot) dyBb[(size_t)k * mn + idx] -= dprev;
        }
      }
    }
    STEP_MARK();  // 3: bundle re-based
    // the accepted point (the trial point becomes the accepted one)
    if (accepted && it > 0)
      for (int i = tid; i < T * n; i += NT) Xc[i] = X[i];  // same thread -> same entries as in the staging loop above
    if (done >= 0) {
      if (tid == 0) { p.status[b] = done; p.iters[b] = it; p.lam[b] = lam; p.nu[b] = nu; }
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      if (!FUSED) stamp_end(p.ts);
      return false;
    }
  }
  const double* Xst = accepted ? X : X2;  // the standing point (accepted trial, or the old accepted point after a rejection)
  const double* Xot = accepted ? X2 : X;  // the other one of the two: the iterate that was left / the rejected trial
  for (int i = tid; i < m; i += NT) fm[i] = 0u;
  for (int i = tid; i < m * n; i += NT) dfix[i] = 0.0;
  __syncthreads();
  STEP_MARK();  // 4: accept / reject done

  // ---------------- gradient with the analytic velocity terms, active set, projected-gradient test ----------------
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  const float* Hc = HS + (size_t)cur * m * nn;  // knots 2..T-1 of the accepted buffer
  const double* gc = gS + (size_t)cur * m * n;
  {
    // bundle update, part 2: the new cutting plane (gradient of the point we do not stand on: the other Gauss-Newton buffer + the
    // analytic velocity term there) is formed in the same pass as the gradient of the standing point
    const double* go = gS + (size_t)(1 - cur) * mn;
    double pr[2] = {0.0, 0.0};  // [0] |projected gradient|_inf, [1] g_new . dy_new
    for (int idx = tid; idx < m * n; idx += NT) {
      const int i = idx / n, k = idx - i * n, t = i + 2;
      const double x = Xst[t * n + k];
      double gv = x - Xst[(t - 1) * n + k];
      if (t < T - 1) gv -= Xst[(t + 1) * n + k] - x;
      const double gtv = gc[idx] + a2 * gv;
      const bool fixed = (x <= R.lo[k] + p.bound_eps && gtv > 0.0) || (x >= R.hi[k] - p.bound_eps && gtv < 0.0);
      gt[idx] = gtv;
      if (fixed) atomicOr(fm + i, 1u << k);
      else pr[0] = fmax(pr[0], fabs(gtv));
      if (slot >= 0) {
        const double xo = Xot[t * n + k];
        double go_v = xo - Xot[(t - 1) * n + k];
        if (t < T - 1) go_v -= Xot[(t + 1) * n + k] - xo;
        const double gtot = go[idx] + a2 * go_v, dy = xo - x;
        gBb[(size_t)slot * mn + idx] = gtot;
        dyBb[(size_t)slot * mn + idx] = dy;
        pr[1] = fma(gtot, dy, pr[1]);
      }
    }
    cta_reduce_n<2>(pr, slot >= 0 ? 3u : 1u, 1u, red2);  // (the barriers inside also make the new piece visible to the whole CTA)
    if (2.0 * pr[0] <= p.tol_grad) {
      if (tid == 0) { p.status[b] = GTO_STATUS_CONVERGED; p.iters[b] = it; p.lam[b] = lam; p.nu[b] = nu; }
      if (!FUSED) stamp_end(p.ts);
      return false;
    }
    if (slot >= 0) {
      const double Fst = accepted ? Ft0 : Fcur, Fot = accepted ? Fcur : Ft0;
      const double e = step > p.bundle_radius ? -1e300 : -fabs(0.5 * (Fot - Fst) - pr[1]);
#pragma unroll
      for (int k = 0; k < KM; ++k)
        if (k == slot) eB[k + 1] = e;
      if (nb < KB) ++nb;
      if (tid == 0) { FBb[slot] = Fot; p.nbund[b] = nb; }
    }
  }
  __syncthreads();
  STEP_MARK();  // 5: gradient, active set

  // ---------------- damped projected Gauss-Newton step: block cyclic reduction in float64 ----------------
  constexpr int OUT_A = (2 * NP * NP + (GTO_BUNDLE_MAX + 1) * NP + 31) / 32;  // results a lane holds in phase A / B before they are written back
  constexpr int OUT_B = (3 * NP * NP + (GTO_BUNDLE_MAX + 1) * NP + 31) / 32;
  // no piece within bundle_radius of the standing point (the usual case while the steps are long): plain Levenberg-Marquardt step.
  // Pieces that are switched off (e = -1e300) never get weight in the dual, so skipping their solves does not change the result.
  {
    bool any = false;
#pragma unroll
    for (int k = 0; k < GTO_BUNDLE_MAX; ++k)
      if (k < nb && eB[k + 1] > -1e299) any = true;
    if (!any) nb = 0;
  }
  const int nrhs = nb + 1;  // right-hand side 0: -g (the Levenberg-Marquardt step); k: -g_k of bundle piece k
  // Active-set rounds (gto_options.as_rounds, oracle lm_step): a free variable that the step pushes beyond a joint limit is
  // moved exactly onto the limit (prescribed step dfix) and the other variables are re-solved with that step on the
  // right-hand side.  Clipping alone distorts the coupled step (the model then often predicts an increase).
  bool ok = false;
  int as_round = 0;
  for (int attempt = 0; attempt < 8 && !ok; ++attempt) {
    // build the masked, damped system
    for (int idx = tid; idx < m * nn; idx += NT) {
      const int i = idx / nn, rc = idx - i * nn, r = rc / n, c = rc - r * n;
      const unsigned mi = fm[i];
      const bool fr = (mi >> r) & 1u, fc = (mi >> c) & 1u;
      double v = (double)Hc[idx];
      if (r == c) {
        v += a2 * ((i + 2 < T - 1) ? 2.0 : 1.0);
        v += lam * v;
      }
      if (fr || fc) v = (r == c) ? 1.0 : 0.0;
      Dm[idx] = v;
      double l = 0.0, u = 0.0;
      if (r == c && !fr) {
        if (i > 0 && !((fm[i - 1] >> r) & 1u)) l = -a2;
        if (i < m - 1 && !((fm[i + 1] >> r) & 1u)) u = -a2;
      }
      Lm[idx] = l;
      Um[idx] = u;
    }
    for (int idx = tid; idx < m * n; idx += NT) {
      const int i = idx / n, k = idx - i * n;
      const unsigned mi = fm[i];
      if ((mi >> k) & 1u) {
        for (int q = 0; q < nrhs; ++q) bb[(size_t)q * mn + idx] = dfix[idx];
      } else {  // free row: -g - (coupling to the prescribed steps of the held variables)
        double cpl = 0.0;
        if (as_round > 0) {
          const float* Hr = Hc + (size_t)i * nn + k * n;
          for (int c = 0; c < n; ++c)
            if (c != k && ((mi >> c) & 1u)) cpl -= (double)Hr[c] * dfix[i * n + c];
          if (i > 0 && ((fm[i - 1] >> k) & 1u)) cpl += a2 * dfix[idx - n];
          if (i < m - 1 && ((fm[i + 1] >> k) & 1u)) cpl += a2 * dfix[idx + n];
        }
        bb[idx] = -gt[idx] + cpl;
        for (int q = 1; q < nrhs; ++q) bb[(size_t)q * mn + idx] = -gBb[(size_t)(q - 1) * mn + idx] + cpl;
      }
    }
    if (tid == 0) sflag[0] = 0;
    __syncthreads();
    STEP_MARK();  // 6: system built

    // forward: strides 1, 2, 4, ...; the last pass (s >= m) eliminates block 0, which has no neighbour left
    int s = 1;
    for (;; s <<= 1) {
      const bool last = (s >= m);
      // ---- phase A: one warp per eliminated block i = s (mod 2s) ----
      const int nel = last ? 1 : (m - s + 2 * s - 1) / (2 * s);
      for (int e = warp; e < nel; e += NW) {
        const int i = last ? 0 : s + 2 * s * e;
        double* Di = Dm + (size_t)i * nn;
        double* Li = Lm + (size_t)i * nn;
        double* Ui = Um + (size_t)i * nn;
        double* bi = bb + (size_t)i * n;
        const bool hasL = !last;                 // i - s >= 0 always holds for an eliminated block
        const bool hasU = !last && (i + s < m);
        double row[NP];
#pragma unroll
        for (int c = 0; c < NP; ++c) row[c] = (lane < n && c < n) ? Di[lane * n + c] : ((lane == c) ? 1.0 : 0.0);
        const bool good = gj_inverse<NP>(row, lane);
        if (!good && lane == 0) sflag[0] = 1;
        __syncwarp();
        if (lane < n) {
#pragma unroll
          for (int c = 0; c < NP; ++c)
            if (c < n) Di[lane * n + c] = row[c];
        }
        __syncwarp();
        // W_L = D^-1 L, W_U = D^-1 U, w = D^-1 b : outputs dealt over the 32 lanes, written back in place afterwards
        double outv[OUT_A];
        const int nout = 2 * nn + nrhs * n;
#pragma unroll
        for (int o = 0; o < OUT_A; ++o) {
          const int id = lane + 32 * o;
          double acc = 0.0;
          if (id < nout) {
            if (id < 2 * nn) {
              const bool isU = id >= nn;
              const int rc = isU ? id - nn : id, r = rc / n, c = rc - r * n;
              const double* M = isU ? Ui : Li;
              if (isU ? hasU : hasL) {
                if (s == 1) acc = Di[r * n + c] * M[c * n + c];  // first level: the couplings are still diagonal (-a2 I, masked)
                else
                  for (int k = 0; k < n; ++k) acc = fma(Di[r * n + k], M[k * n + c], acc);
              }
            } else {
              const int qr = id - 2 * nn, q = qr / n, r = qr - q * n;
              const double* bq = bi + (size_t)q * mn;
              for (int k = 0; k < n; ++k) acc = fma(Di[r * n + k], bq[k], acc);
            }
          }
          outv[o] = acc;
        }
        __syncwarp();
#pragma unroll
        for (int o = 0; o < OUT_A; ++o) {
          const int id = lane + 32 * o;
          if (id < nn) Li[id] = outv[o];
          else if (id < 2 * nn) Ui[id - nn] = outv[o];
          else if (id < nout) {
            const int qr = id - 2 * nn, q = qr / n, r = qr - q * n;
            bi[(size_t)q * mn + r] = outv[o];
          }
        }
      }
      __syncthreads();
      STEP_MARK();  // phase A of this level
      if (last) break;
      // ---- phase B: one warp per kept block j = 0 (mod 2s): absorb the Schur complements of j - s and j + s ----
      const int nkeep = (m + 2 * s - 1) / (2 * s);
      for (int e = warp; e < nkeep; e += NW) {
        const int j = 2 * s * e;
        const int i1 = j - s, i2 = j + s;
        const bool has1 = (i1 >= 0), has2 = (i2 < m);
        if (!has1 && !has2) continue;
        double* Dj = Dm + (size_t)j * nn;
        double* Lj = Lm + (size_t)j * nn;
        double* Uj = Um + (size_t)j * nn;
        double* bj = bb + (size_t)j * n;
        const double* WL1 = Lm + (size_t)(has1 ? i1 : 0) * nn;
        const double* WU1 = Um + (size_t)(has1 ? i1 : 0) * nn;
        const double* w1 = bb + (size_t)(has1 ? i1 : 0) * n;
        const double* WL2 = Lm + (size_t)(has2 ? i2 : 0) * nn;
        const double* WU2 = Um + (size_t)(has2 ? i2 : 0) * nn;
        const double* w2 = bb + (size_t)(has2 ? i2 : 0) * n;
        const bool has1L = has1 && (i1 - s >= 0);  // j - 2s exists
        const bool has2U = has2 && (i2 + s < m);   // j + 2s exists
        double outv[OUT_B];
        const int nout = 3 * nn + nrhs * n;
#pragma unroll
        for (int o = 0; o < OUT_B; ++o) {
          const int id = lane + 32 * o;
          double acc = 0.0;
          if (id < nout) {
            if (s == 1) {  // first level: L_j, U_j are diagonal, the products are row scalings
              if (id < nn) {
                const int r = id / n, c = id - r * n;
                acc = Dj[id];
                if (has2) acc = fma(-Uj[r * n + r], WL2[r * n + c], acc);
                if (has1) acc = fma(-Lj[r * n + r], WU1[r * n + c], acc);
              } else if (id < 2 * nn) {
                const int rc = id - nn, r = rc / n, c = rc - r * n;
                if (has1L) acc = -Lj[r * n + r] * WL1[r * n + c];
              } else if (id < 3 * nn) {
                const int rc = id - 2 * nn, r = rc / n, c = rc - r * n;
                if (has2U) acc = -Uj[r * n + r] * WU2[r * n + c];
              } else {
                const int qr = id - 3 * nn, q = qr / n, r = qr - q * n;
                const size_t qo = (size_t)q * mn;
                acc = bj[qo + r];
                if (has2) acc = fma(-Uj[r * n + r], w2[qo + r], acc);
                if (has1) acc = fma(-Lj[r * n + r], w1[qo + r], acc);
              }
            } else if (id < nn) {  // D_j - U_j W_L(j+s) - L_j W_U(j-s)
              const int r = id / n, c = id - r * n;
              acc = Dj[id];
              if (has2)
                for (int k = 0; k < n; ++k) acc = fma(-Uj[r * n + k], WL2[k * n + c], acc);
              if (has1)
                for (int k = 0; k < n; ++k) acc = fma(-Lj[r * n + k], WU1[k * n + c], acc);
            } else if (id < 2 * nn) {  // new L_j = -L_j W_L(j-s)
              const int rc = id - nn, r = rc / n, c = rc - r * n;
              if (has1L)
                for (int k = 0; k < n; ++k) acc = fma(-Lj[r * n + k], WL1[k * n + c], acc);
            } else if (id < 3 * nn) {  // new U_j = -U_j W_U(j+s)
              const int rc = id - 2 * nn, r = rc / n, c = rc - r * n;
              if (has2U)
                for (int k = 0; k < n; ++k) acc = fma(-Uj[r * n + k], WU2[k * n + c], acc);
            } else {  // b_j - U_j w(j+s) - L_j w(j-s)
              const int qr = id - 3 * nn, q = qr / n, r = qr - q * n;
              const size_t qo = (size_t)q * mn;
              acc = bj[qo + r];
              if (has2)
                for (int k = 0; k < n; ++k) acc = fma(-Uj[r * n + k], w2[qo + k], acc);
              if (has1)
                for (int k = 0; k < n; ++k) acc = fma(-Lj[r * n + k], w1[qo + k], acc);
            }
          }
          outv[o] = acc;
        }
        __syncwarp();
#pragma unroll
        for (int o = 0; o < OUT_B; ++o) {
          const int id = lane + 32 * o;
          if (id < nn) Dj[id] = outv[o];
          else if (id < 2 * nn) Lj[id - nn] = outv[o];
          else if (id < 3 * nn) Uj[id - 2 * nn] = outv[o];
          else if (id < nout) {
            const int qr = id - 3 * nn, q = qr / n, r = qr - q * n;
            bj[(size_t)q * mn + r] = outv[o];
          }
        }
      }
      __syncthreads();
      STEP_MARK();  // phase B of this level
    }
    ok = (sflag[0] == 0);
    __syncthreads();
    if (!ok) {
      lam = fmin(p.lambda_max, lam * 10.0);  // not positive definite: add damping and refactor
      continue;
    }
    // back substitution: x_0 = w_0, then the eliminated blocks level by level, coarsest first
    for (int idx = tid; idx < nrhs * n; idx += NT) {
      const int q = idx / n, r = idx - q * n;
      xs[(size_t)q * mn + r] = bb[(size_t)q * mn + r];
    }
    __syncthreads();
    for (s >>= 1; s >= 1; s >>= 1) {
      const int nel = (m - s + 2 * s - 1) / (2 * s);
      for (int idx = tid; idx < nrhs * nel * n; idx += NT) {
        const int q = idx / (nel * n), er = idx - q * nel * n;
        const int e = er / n, r = er - e * n;
        const int i = s + 2 * s * e;
        const size_t qo = (size_t)q * mn;
        const double* WL = Lm + (size_t)i * nn + r * n;
        const double* WU = Um + (size_t)i * nn + r * n;
        double acc = bb[qo + i * n + r];
        const double* xl = xs + qo + (size_t)(i - s) * n;
        for (int k = 0; k < n; ++k) acc = fma(-WL[k], xl[k], acc);
        if (i + s < m) {
          const double* xu = xs + qo + (size_t)(i + s) * n;
          for (int k = 0; k < n; ++k) acc = fma(-WU[k], xu[k], acc);
        }
        xs[qo + i * n + r] = acc;
      }
      __syncthreads();
    }
    STEP_MARK();  // bundle weights, combined step, active-set round
  if (p.dbg && p.iter == p.dbg_iter && tid == 0) {  // counters of this launch: CTAs, CTAs that re-solved for the active set, CTAs with planes in use
    atomicAdd(reinterpret_cast<unsigned long long*>(p.dbg + 40), 1ull);
    if (as_round > 0) atomicAdd(reinterpret_cast<unsigned long long*>(p.dbg + 41), 1ull);
    if (nb > 0) atomicAdd(reinterpret_cast<unsigned long long*>(p.dbg + 42), 1ull);
  }
    if (nb > 0) {  // convex weights of the pieces (dual QP of the bundle model), combined step into xs[0]
      constexpr int KM = GTO_BUNDLE_MAX, NV = KM + KM * KM;
      double acc[NV];  // [k]: (g_k - g).d_0, [KM + k KM + j]: (g_k - g).(d_j - d_0)
#pragma unroll
      for (int v = 0; v < NV; ++v) acc[v] = 0.0;
      for (int idx = tid; idx < mn; idx += NT) {
        const double g0 = gt[idx], x0 = xs[idx];
#pragma unroll
        for (int k = 0; k < KM; ++k)
          if (k < nb) {
            const double dg = gBb[(size_t)k * mn + idx] - g0;
            acc[k] = fma(dg, x0, acc[k]);
#pragma unroll
            for (int j = 0; j < KM; ++j)
              if (j < nb) acc[KM + k * KM + j] = fma(dg, xs[(size_t)(j + 1) * mn + idx] - x0, acc[KM + k * KM + j]);
          }
      }
      unsigned used = 0u;
#pragma unroll
      for (int k = 0; k < KM; ++k)
        if (k < nb) {
          used |= 1u << k;
#pragma unroll
          for (int j = 0; j < KM; ++j)
            if (j < nb) used |= 1u << (KM + k * KM + j);
        }
      cta_reduce_n<NV>(acc, used, 0u, red2);
      if (tid == 0) {
        double bq[KM + 1], M[KM + 1][KM + 1], th[KM + 1];
#pragma unroll
        for (int k = 0; k <= KM; ++k) {
          bq[k] = (k >= 1 && k <= nb) ? eB[k] + acc[k >= 1 ? k - 1 : 0] : 0.0;
#pragma unroll
          for (int j = 0; j <= KM; ++j)
            M[k][j] = (k >= 1 && j >= 1 && k <= nb && j <= nb) ? 0.5 * (acc[KM + (k >= 1 ? k - 1 : 0) * KM + (j >= 1 ? j - 1 : 0)] + acc[KM + (j >= 1 ? j - 1 : 0) * KM + (k >= 1 ? k - 1 : 0)]) : 0.0;
        }
        bundle_dual(nb, bq, M, th);
#pragma unroll
        for (int k = 0; k <= KM; ++k) thS[k] = th[k];
      }
      __syncthreads();
      for (int idx = tid; idx < mn; idx += NT) {
        const double x0 = xs[idx];
        double sx = x0;
        for (int k = 1; k <= nb; ++k) {
          const double th = thS[k];
          if (th != 0.0) sx = fma(th, xs[(size_t)k * mn + idx] - x0, sx);
        }
        xs[idx] = sx;
      }
      __syncthreads();
    }
    if (as_round < p.as_rounds) {  // free variables pushed beyond a limit: hold them on it and solve again
      int viol = 0;
      for (int idx = tid; idx < m * n; idx += NT) {
        const int i = idx / n, r = idx - i * n;
        if ((fm[i] >> r) & 1u) continue;
        const double xc = Xst[(i + 2) * n + r], xn = xc + xs[idx];
        if (xn < R.lo[r]) { dfix[idx] = R.lo[r] - xc; viol = 1; }
        else if (xn > R.hi[r]) { dfix[idx] = R.hi[r] - xc; viol = 1; }
        if (xn < R.lo[r] || xn > R.hi[r]) atomicOr(fm + i, 1u << r);
      }
      if (__syncthreads_or(viol)) {
        ++as_round;
        ok = false;
        attempt = -1;  // the positive-definiteness retries start over for the new system
      }
    }
  }
  if (!ok) {
    if (tid == 0) { p.status[b] = GTO_STATUS_NAN; p.iters[b] = it; }
    if (!FUSED) stamp_end(p.ts);
    return false;
  }

  STEP_MARK();  // back substitution done
  // ---------------- trial point = clip(X + x); predicted reduction with the undamped, unmasked model ----------------
  double stepmax = 0.0, gdot = 0.0;
  for (int idx = tid; idx < m * n; idx += NT) {
    const int i = idx / n, r = idx - i * n, t = i + 2;
    const double xc = Xst[t * n + r];
    const double xn = fmin(fmax(xc + xs[idx], R.lo[r]), R.hi[r]);
    const double d = xn - xc;
    Xt[t * n + r] = xn;
    p.q_trial[((long long)b * T + t) * R.ndof + R.opt_qidx[r]] = xn;
    dd[idx] = d;
    stepmax = fmax(stepmax, fabs(d));
    gdot += gt[idx] * d;
  }
  stepmax = cta_max(stepmax, red);
  gdot = cta_sum(gdot, red);  // (the barriers inside also publish dd)
  if (nb > 0) {  // linear part of the bundle model: max over the pieces
    constexpr int KM = GTO_BUNDLE_MAX;
    double lk[KM];
#pragma unroll
    for (int k = 0; k < KM; ++k) lk[k] = 0.0;
    for (int idx = tid; idx < mn; idx += NT) {
      const double dr = dd[idx];
#pragma unroll
      for (int k = 0; k < KM; ++k)
        if (k < nb) lk[k] = fma(gBb[(size_t)k * mn + idx], dr, lk[k]);
    }
    cta_reduce_n<KM>(lk, (1u << nb) - 1u, 0u, red2);
#pragma unroll
    for (int k = 0; k < KM; ++k)
      if (k < nb) gdot = fmax(gdot, eB[k + 1] + lk[k]);
  }
  double quad = 0.0;
  for (int idx = tid; idx < m * n; idx += NT) {
    const int i = idx / n, r = idx - i * n;
    const double dg = a2 * ((i + 2 < T - 1) ? 2.0 : 1.0);
    const double dr = dd[idx];
    double hd = dg * dr;
    const float* Hr = Hc + (size_t)i * nn + r * n;
    for (int c = 0; c < n; ++c) hd += (double)Hr[c] * dd[i * n + c];
    quad += dr * hd;
    if (i < m - 1) quad -= 2.0 * a2 * dr * dd[idx + n];
  }
  quad = cta_sum(quad, red);
  STEP_MARK();  // trial point, predicted reduction
  if (tid == 0) {
    p.pred[b] = -(gdot + 0.5 * quad);
    p.stepn[b] = stepmax;
    p.lam[b] = lam;
    p.nu[b] = nu;
    p.iters[b] = it + 1;
    if (!FUSED) {
      const int slot = atomicAdd(p.nactive_out, 1);
      p.active_out[slot] = b;
      sflag[1] = slot;
    }
  }
  if (FUSED) return true;
  if (!p.do_fk) { stamp_end(p.ts); return true; }
  // ---------------- item records of the trial point (what k_item_fk would compute in a launch of its own) ----------------
  __syncthreads();  // q_trial of this problem and the slot are visible to the whole CTA
  {
    const int slot = sflag[1];
    // the robot table moves into shared memory (the storage of X2 / HS / gS is free by now) when it fits: the serial FK chain
    // then never waits on L2
    const RobotDev* Rf = &R;
    if (p.fk_robot_smem) {
      const unsigned long long* src = reinterpret_cast<const unsigned long long*>(p.robot);
      unsigned long long* dst = reinterpret_cast<unsigned long long*>(X2);
      for (int i = tid; i < (int)(sizeof(RobotDev) / 8); i += NT) dst[i] = __ldg(src + i);
      __syncthreads();
      Rf = reinterpret_cast<const RobotDev*>(X2);
    }
    const int hl = tid & 15, grp = tid >> 4, ngrp = NT >> 4, hshift = tid & 16;
    double* A = Dm + (size_t)grp * 2 * R.nmov * 12;  // the factorisation is done: its storage is free
    double* Tm = A + (size_t)R.nmov * 12;
    for (int i0 = 0; i0 < m; i0 += ngrp) {
      const bool valid = (i0 + grp) < m;
      const int i = valid ? i0 + grp : m - 1;
      const int t = i + 2;
      item_fk_body(p.fk, p.fk.recs, *Rf, slot * m + i, valid, b, t, 1 - cur, p.q_trial + ((long long)b * T + t) * R.ndof, A, Tm, hl, hshift);
      __syncwarp();
    }
  }
  stamp_end(p.ts);
  return true;
}

template <int NP, bool EXACT>
__global__ void __launch_bounds__(STEP_CR_THREADS, 2) k_step_cr(const StepParams p) {
  extern __shared__ __align__(16) unsigned char step_smem[];
  if ((int)blockIdx.x >= *p.nactive_in) return;  // written by the step kernel two launches back: may be read ahead of pdl_wait
  step_body<NP, EXACT, false>(p, step_smem, *p.robot, p.active_in[blockIdx.x], p.iter);
}

