// base_place.cuh -- mobile-base placement (reference gto/base_planner.py:35-168).
//
// One problem = one reference BasePlanner.plan_goalset(qc, RTs) call: unknowns y = (x, y, theta) of the new base and one
// arm configuration q_i per goal; cost  w_e |y|^2 + sum_i sum_k | F(q_i) x_k - T_b(y) A_i x_k |^2  with A_i = RT_i.G,
// T_b(y) = [Rz(theta) | (x, y, 0)], bounds -pi <= theta <= pi and the joint limits.  The reference builds one such NLP, hands it
// to IPOPT and repeats it in a rejection loop over random grasp subsets until the occupancy-grid collision count of the robot at
// the new base is zero (examples/pybullet_gto_planning_mobile.py:187-201).  Here B such problems (B grasp subsets) are solved at
// once, to convergence, in ONE launch:
//
//   * one warp per problem, one lane per goal (n <= 32); everything is float64 and lives in the lane's registers / local memory;
//   * the sum over the gripper points never touches the points: with D = [F_R - M_R | F_t - M_t] (3x4) the cost of a goal is
//     tr(D Mom D^T), Mom = sum_k [x_k;1][x_k;1]^T (4x4, computed once on the host), so a goal costs O(nopt^2) flops per
//     iteration instead of O(Pg nopt^2), and the Gauss-Newton blocks  <E_a Mom, E_b>  (E_a = dD/d(parameter a)) are exactly the
//     J^T J of the per-point rows;
//   * the Gauss-Newton system is an arrow: per-goal blocks H_i (nopt x nopt), couplings C_i (nopt x 3), one 3x3 base block.
//     Each lane factors its own damped, bound-masked H_i (Cholesky) and solves for [C_i | -g_i]; the 3x3 Schur complement is
//     summed with warp shuffles and solved redundantly by every lane;
//   * the LM bookkeeping (damping policy of oracle/gto_oracle.py solve_lm, projected on the bounds) is warp-uniform, so the
//     whole iteration loop runs inside the kernel -- no host round trips;
//   * afterwards the warp counts the occupied cells under the robot's surface points seen from the new base
//     (gto/base_planner.py:150-165, gto/gto_models.py:261-271).
//
// This is latency-bound small dense algebra (a few kflop per goal and iteration), not HBM- or tensor-bound; the measure that
// matters is problems per second (tools/bench_rows_f.py).
#pragma once

// NP = padded number of optimised joints (8 or 16); a goal sees NP + 3 parameters: its arm joints + (x, y, theta)

struct BaseParams {
  const RobotDev* robot;
  int B, n;
  const double* qc;    // [ndof]
  const double* goal;  // [B][n][12]  A_i = RT_i . G
  double w_effort;
  int nchain;
  int chain[GTO_MAX_MOV];  // movable joints from the root to the gripper link
  double mom[16];          // moment matrix of the gripper point set
  int max_iter;
  double tol_step, tol_grad, lambda0, lambda_min, lambda_max, eta, bound_eps;
  const float* occ;  // [onx][ony] or nullptr
  int onx, ony;
  double oox, ooy, ores;
  const double* wp;  // [npoints][3] robot surface points at qc, current base frame
  int npoints;
  double* Qx;  // [B][n][nopt]
  double* y;   // [B][3]
  double* cost;
  double* collision;
  int* iters;
  int* status;
};

// world points of every collision link at qc (float64 FK): one block
__global__ void k_base_points(const RobotDev* robot, const float* px, const float* py, const float* pz, const double* qc, double* wp) {
  __shared__ double Tm[GTO_MAX_MOV][12];
  __shared__ double frames[GTO_MAX_LINKS][12];
  const RobotDev& R = *robot;
  if (threadIdx.x == 0) {
    for (int j = 0; j < R.nmov; ++j) {
      const double qj = qc[R.mov_qidx[j]];
      const double ax = R.mov_axis_d[j][0], ay = R.mov_axis_d[j][1], az = R.mov_axis_d[j][2];
      double M[12], A[12];
      if (R.mov_type[j] == GTO_JOINT_REVOLUTE) {
        double s, c;
        sincos(qj, &s, &c);
        const double v = 1.0 - c;
        M[0] = 1.0 - v * (ay * ay + az * az); M[1] = -s * az + v * ax * ay; M[2] = s * ay + v * ax * az; M[3] = 0.0;
        M[4] = s * az + v * ax * ay; M[5] = 1.0 - v * (ax * ax + az * az); M[6] = -s * ax + v * ay * az; M[7] = 0.0;
        M[8] = -s * ay + v * ax * az; M[9] = s * ax + v * ay * az; M[10] = 1.0 - v * (ax * ax + ay * ay); M[11] = 0.0;
      } else {
        M[0] = 1.0; M[1] = 0.0; M[2] = 0.0; M[3] = qj * ax;
        M[4] = 0.0; M[5] = 1.0; M[6] = 0.0; M[7] = qj * ay;
        M[8] = 0.0; M[9] = 0.0; M[10] = 1.0; M[11] = qj * az;
      }
      mul34(R.mov_origin_d[j], M, A);
      if (R.mov_parent[j] < 0) {
        for (int e = 0; e < 12; ++e) Tm[j][e] = A[e];
      } else {
        double C[12];
        mul34(Tm[R.mov_parent[j]], A, C);
        for (int e = 0; e < 12; ++e) Tm[j][e] = C[e];
      }
    }
    for (int l = 0; l < R.nlinks; ++l) {
      if (R.link_mov[l] < 0) {
        for (int e = 0; e < 12; ++e) frames[l][e] = R.link_tf_d[l][e];
      } else {
        double C[12];
        mul34(Tm[R.link_mov[l]], R.link_tf_d[l], C);
        for (int e = 0; e < 12; ++e) frames[l][e] = C[e];
      }
    }
  }
  __syncthreads();
  for (int l = 0; l < R.nlinks; ++l) {
    const double* F = frames[l];
    for (int i = threadIdx.x; i < R.link_pt_count[l]; i += blockDim.x) {
      const int pi = R.link_pt_start[l] + i;
      const double x = px[pi], y = py[pi], z = pz[pi];
      wp[3 * pi + 0] = F[0] * x + F[1] * y + F[2] * z + F[3];
      wp[3 * pi + 1] = F[4] * x + F[5] * y + F[6] * z + F[7];
      wp[3 * pi + 2] = F[8] * x + F[9] * y + F[10] * z + F[11];
    }
  }
}

// Linearisation of one goal at (q, y): cost, half gradient g[a] = <E_a, D Mom>, Gram matrix G[a][b] = <E_a Mom, E_b> over the
// parameters a = arm joints 0..nopt-1, then x, y, theta at nopt..nopt+2.
template <int NP>
__device__ __noinline__ void base_goal_linearize(const BaseParams& P, const double* __restrict__ qx, const double* __restrict__ yv,
                                                 const double* __restrict__ A, double& cost, double* __restrict__ g,
                                                 double (*__restrict__ G)[NP + 3]) {
  const RobotDev& R = *P.robot;
  const int nopt = R.nopt, nv = nopt + 3;
  double E[NP + 3][12];
  double om[NP][3], mm[NP][3];
  bool on_chain[NP];
  for (int k = 0; k < nopt; ++k) on_chain[k] = false;
  // chain FK root -> gripper link
  double T[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
  for (int c = 0; c < P.nchain; ++c) {
    const int j = P.chain[c];
    double U[12];
    mul34(T, R.mov_origin_d[j], U);
    const double ax = R.mov_axis_d[j][0], ay = R.mov_axis_d[j][1], az = R.mov_axis_d[j][2];
    const double zx = U[0] * ax + U[1] * ay + U[2] * az, zy = U[4] * ax + U[5] * ay + U[6] * az, zz = U[8] * ax + U[9] * ay + U[10] * az;
    const int k = R.mov_opt[j];
    const double qj = k >= 0 ? qx[k] : P.qc[R.mov_qidx[j]];
    if (R.mov_type[j] == GTO_JOINT_REVOLUTE) {
      if (k >= 0) {
        on_chain[k] = true;
        om[k][0] = zx; om[k][1] = zy; om[k][2] = zz;
        mm[k][0] = U[7] * zz - U[11] * zy;  // o x z
        mm[k][1] = U[11] * zx - U[3] * zz;
        mm[k][2] = U[3] * zy - U[7] * zx;
      }
      double s, cs;
      sincos(qj, &s, &cs);
      const double v = 1.0 - cs;
      double M[12];
      M[0] = 1.0 - v * (ay * ay + az * az); M[1] = -s * az + v * ax * ay; M[2] = s * ay + v * ax * az; M[3] = 0.0;
      M[4] = s * az + v * ax * ay; M[5] = 1.0 - v * (ax * ax + az * az); M[6] = -s * ax + v * ay * az; M[7] = 0.0;
      M[8] = -s * ay + v * ax * az; M[9] = s * ax + v * ay * az; M[10] = 1.0 - v * (ax * ax + ay * ay); M[11] = 0.0;
      mul34(U, M, T);
    } else {
      if (k >= 0) {
        on_chain[k] = true;
        om[k][0] = 0.0; om[k][1] = 0.0; om[k][2] = 0.0;
        mm[k][0] = zx; mm[k][1] = zy; mm[k][2] = zz;
      }
      for (int e = 0; e < 12; ++e) T[e] = U[e];
      T[3] += qj * zx; T[7] += qj * zy; T[11] += qj * zz;
    }
  }
  double F[12];
  mul34(T, R.grip_tf_d, F);
  // goal frame in the new base: M = T_b(y) A
  double s, c;
  sincos(yv[2], &s, &c);
  double D[12], RA[12];  // RA = dRz/dtheta . A
  for (int col = 0; col < 4; ++col) {
    const double a0 = A[col], a1 = A[4 + col], a2 = A[8 + col];
    const double m0 = c * a0 - s * a1 + (col == 3 ? yv[0] : 0.0);
    const double m1 = s * a0 + c * a1 + (col == 3 ? yv[1] : 0.0);
    D[col] = F[col] - m0;
    D[4 + col] = F[4 + col] - m1;
    D[8 + col] = F[8 + col] - a2;
    RA[col] = -s * a0 - c * a1;
    RA[4 + col] = c * a0 - s * a1;
    RA[8 + col] = 0.0;
  }
  // parameter directions E_a = dD/da
  for (int k = 0; k < nopt; ++k) {
    if (!on_chain[k]) {
      for (int e = 0; e < 12; ++e) E[k][e] = 0.0;
      continue;
    }
    const double wx = om[k][0], wy = om[k][1], wz = om[k][2];
    for (int col = 0; col < 4; ++col) {
      const double f0 = F[col], f1 = F[4 + col], f2 = F[8 + col];
      E[k][col] = wy * f2 - wz * f1 + (col == 3 ? mm[k][0] : 0.0);
      E[k][4 + col] = wz * f0 - wx * f2 + (col == 3 ? mm[k][1] : 0.0);
      E[k][8 + col] = wx * f1 - wy * f0 + (col == 3 ? mm[k][2] : 0.0);
    }
  }
  for (int e = 0; e < 12; ++e) { E[nopt][e] = 0.0; E[nopt + 1][e] = 0.0; E[nopt + 2][e] = -RA[e]; }
  E[nopt][3] = -1.0;
  E[nopt + 1][7] = -1.0;
  // D Mom, cost, gradient, Gram matrix
  double DM[12];
  double cs = 0.0;
  for (int r = 0; r < 3; ++r)
    for (int col = 0; col < 4; ++col) {
      double v = 0.0;
      for (int m = 0; m < 4; ++m) v += D[4 * r + m] * P.mom[4 * m + col];
      DM[4 * r + col] = v;
      cs += v * D[4 * r + col];
    }
  cost = cs;
  for (int a = 0; a < nv; ++a) {
    double v = 0.0;
    for (int e = 0; e < 12; ++e) v += E[a][e] * DM[e];
    g[a] = v;
  }
  for (int a = 0; a < nv; ++a) {
    double EM[12];
    for (int r = 0; r < 3; ++r)
      for (int col = 0; col < 4; ++col) {
        double v = 0.0;
        for (int m = 0; m < 4; ++m) v += E[a][4 * r + m] * P.mom[4 * m + col];
        EM[4 * r + col] = v;
      }
    for (int b = a; b < nv; ++b) {
      double v = 0.0;
      for (int e = 0; e < 12; ++e) v += EM[e] * E[b][e];
      G[a][b] = v;
      G[b][a] = v;
    }
  }
}

// solve the SPD system M X = R (order n <= NP, 4 right-hand sides) by Cholesky, in place in R
template <int NP>
__device__ __forceinline__ void base_chol_solve4(double (*M)[NP], double (*Rh)[4], int n) {
  for (int j = 0; j < n; ++j) {
    double d = M[j][j];
    for (int k = 0; k < j; ++k) d -= M[j][k] * M[j][k];
    d = sqrt(d);
    M[j][j] = d;
    const double inv = 1.0 / d;
    for (int i = j + 1; i < n; ++i) {
      double v = M[i][j];
      for (int k = 0; k < j; ++k) v -= M[i][k] * M[j][k];
      M[i][j] = v * inv;
    }
  }
  for (int r = 0; r < 4; ++r) {
    for (int i = 0; i < n; ++i) {
      double v = Rh[i][r];
      for (int k = 0; k < i; ++k) v -= M[i][k] * Rh[k][r];
      Rh[i][r] = v / M[i][i];
    }
    for (int i = n - 1; i >= 0; --i) {
      double v = Rh[i][r];
      for (int k = i + 1; k < n; ++k) v -= M[k][i] * Rh[k][r];
      Rh[i][r] = v / M[i][i];
    }
  }
}

template <int NP>
__global__ void __launch_bounds__(32) k_base_place(const __grid_constant__ BaseParams P) {
  const RobotDev& R = *P.robot;
  const int b = blockIdx.x, lane = threadIdx.x;
  const int nopt = R.nopt, nv = nopt + 3;
  const bool act = lane < P.n;
  const double BIG = 1e30, PI = 3.14159265358979323846;
  const double ylo[3] = {-BIG, -BIG, -PI}, yhi[3] = {BIG, BIG, PI};
  double A[12];
  for (int e = 0; e < 12; ++e) A[e] = act ? P.goal[((long long)b * P.n + lane) * 12 + e] : 0.0;
  double qx[NP], qn[NP], yv[3] = {0.0, 0.0, 0.0}, yn[3];
  for (int k = 0; k < nopt; ++k) qx[k] = P.qc[R.opt_qidx[k]];
  double Ga[NP + 3][NP + 3], ga[NP + 3], Gb[NP + 3][NP + 3], gb[NP + 3];
  double(*G)[NP + 3] = Ga, (*Gt)[NP + 3] = Gb;  // current / trial linearisation, swapped on acceptance
  double *g = ga, *gtr = gb;
  for (int a = 0; a < nv; ++a) {
    g[a] = 0.0;
    gtr[a] = 0.0;
    for (int c = 0; c < nv; ++c) { G[a][c] = 0.0; Gt[a][c] = 0.0; }
  }
  double ci = 0.0;
  if (act) base_goal_linearize<NP>(P, qx, yv, A, ci, g, G);
  double F = warp_sum(act ? ci : 0.0) + P.w_effort * (yv[0] * yv[0] + yv[1] * yv[1] + yv[2] * yv[2]);
  double lam = P.lambda0, nu = 2.0;
  int status = GTO_STATUS_MAX_ITER, it = 0;
  while (it < P.max_iter) {
    // ---- base block and gradient (warp sums), active sets ----
    double S[3][3], gy[3];
    for (int a = 0; a < 3; ++a) {
      gy[a] = warp_sum(g[nopt + a]) + P.w_effort * yv[a];
      for (int c = a; c < 3; ++c) {
        S[a][c] = warp_sum(G[nopt + a][nopt + c]) + (a == c ? P.w_effort : 0.0);
        S[c][a] = S[a][c];
      }
    }
    bool fy[3], fq[NP];
    double pgmax = 0.0;
    for (int a = 0; a < 3; ++a) {
      fy[a] = (yv[a] <= ylo[a] + P.bound_eps && gy[a] > 0.0) || (yv[a] >= yhi[a] - P.bound_eps && gy[a] < 0.0);
      if (!fy[a]) pgmax = fmax(pgmax, fabs(gy[a]));
    }
    double pgq = 0.0;
    for (int k = 0; k < nopt; ++k) {
      fq[k] = (qx[k] <= R.lo[k] + P.bound_eps && g[k] > 0.0) || (qx[k] >= R.hi[k] - P.bound_eps && g[k] < 0.0) ||
              !((R.grip_optmask >> k) & 1u) || !act;
      if (!fq[k]) pgq = fmax(pgq, fabs(g[k]));
    }
    pgmax = 2.0 * fmax(pgmax, warp_max(pgq));
    if (pgmax <= P.tol_grad) { status = GTO_STATUS_CONVERGED; break; }
    // ---- per-goal solve H_i Z = [C_i | -g_i], Schur complement on the base block ----
    double Hd[NP][NP], Z[NP][4], Cm[NP][3];
    for (int k = 0; k < nopt; ++k) {
      for (int l = 0; l < nopt; ++l) Hd[k][l] = (fq[k] || fq[l]) ? 0.0 : G[k][l];
      Hd[k][k] = fq[k] ? 1.0 : G[k][k] + lam * G[k][k];
      for (int a = 0; a < 3; ++a) {
        Cm[k][a] = (fq[k] || fy[a]) ? 0.0 : G[k][nopt + a];
        Z[k][a] = Cm[k][a];
      }
      Z[k][3] = fq[k] ? 0.0 : -g[k];
    }
    base_chol_solve4<NP>(Hd, Z, nopt);
    double M[3][4];
    for (int a = 0; a < 3; ++a) {
      double v = 0.0;
      for (int k = 0; k < nopt; ++k) v += Cm[k][a] * Z[k][3];
      M[a][3] = (fy[a] ? 0.0 : -gy[a]) - warp_sum(v);
      for (int c = 0; c < 3; ++c) {
        double u = 0.0;
        for (int k = 0; k < nopt; ++k) u += Cm[k][a] * Z[k][c];
        const double base = (fy[a] || fy[c]) ? (a == c ? 1.0 : 0.0) : S[a][c] + (a == c ? lam * S[a][a] : 0.0);
        M[a][c] = base - warp_sum(u);
      }
    }
    double dy[3];  // 3x3 Gaussian elimination (SPD: no pivoting)
    for (int p = 0; p < 3; ++p) {
      const double inv = 1.0 / M[p][p];
      for (int r = p + 1; r < 3; ++r) {
        const double f = M[r][p] * inv;
        for (int c = p; c < 4; ++c) M[r][c] -= f * M[p][c];
      }
    }
    for (int p = 2; p >= 0; --p) {
      double v = M[p][3];
      for (int c = p + 1; c < 3; ++c) v -= M[p][c] * dy[c];
      dy[p] = v / M[p][p];
    }
    // ---- projected trial point ----
    double dq[NP], stepm = 0.0;
    for (int k = 0; k < nopt; ++k) {
      const double d = Z[k][3] - (Z[k][0] * dy[0] + Z[k][1] * dy[1] + Z[k][2] * dy[2]);
      qn[k] = fmin(fmax(qx[k] + d, R.lo[k]), R.hi[k]);
      dq[k] = act ? qn[k] - qx[k] : 0.0;
      stepm = fmax(stepm, fabs(dq[k]));
    }
    for (int a = 0; a < 3; ++a) {
      yn[a] = fmin(fmax(yv[a] + dy[a], ylo[a]), yhi[a]);
      dy[a] = yn[a] - yv[a];
    }
    stepm = fmax(warp_max(stepm), fmax(fabs(dy[0]), fmax(fabs(dy[1]), fabs(dy[2]))));
    // ---- predicted reduction with the undamped model ----
    double pq = 0.0, cd[3] = {0.0, 0.0, 0.0};
    for (int k = 0; k < nopt; ++k) {
      double ad = G[k][nopt] * dy[0] + G[k][nopt + 1] * dy[1] + G[k][nopt + 2] * dy[2];
      for (int l = 0; l < nopt; ++l) ad += G[k][l] * dq[l];
      pq += g[k] * dq[k] + 0.5 * dq[k] * ad;
      for (int a = 0; a < 3; ++a) cd[a] += G[k][nopt + a] * dq[k];
    }
    pq = warp_sum(pq);
    double py = 0.0;
    for (int a = 0; a < 3; ++a) {
      const double ady = S[a][0] * dy[0] + S[a][1] * dy[1] + S[a][2] * dy[2] + warp_sum(cd[a]);
      py += gy[a] * dy[a] + 0.5 * dy[a] * ady;
    }
    const double pred = -(pq + py);
    ++it;
    // ---- trial linearisation, acceptance ----
    double ct = 0.0;
    if (act) base_goal_linearize<NP>(P, qn, yn, A, ct, gtr, Gt);
    const double Ft = warp_sum(act ? ct : 0.0) + P.w_effort * (yn[0] * yn[0] + yn[1] * yn[1] + yn[2] * yn[2]);
    if (!(Ft == Ft) || fabs(Ft) > 1e300) { status = GTO_STATUS_NAN; break; }
    const double ared = 0.5 * (F - Ft);
    if (pred > 0.0 && ared >= P.eta * pred) {
      const double rho = ared / pred;
      const double t = 2.0 * fmin(rho, 1.0) - 1.0;
      for (int k = 0; k < nopt; ++k) qx[k] = qn[k];
      for (int a = 0; a < 3; ++a) yv[a] = yn[a];
      {
        double(*tG)[NP + 3] = G; G = Gt; Gt = tG;
        double* tg = g; g = gtr; gtr = tg;
      }
      F = Ft;
      lam = fmax(P.lambda_min, lam * fmax(1.0 / 3.0, 1.0 - t * t * t));
      nu = 2.0;
      if (stepm <= P.tol_step) { status = GTO_STATUS_CONVERGED; break; }
    } else {
      if (pred <= 0.0 && stepm <= P.tol_step) { status = GTO_STATUS_CONVERGED; break; }
      lam = fmin(P.lambda_max, lam * nu);
      nu *= 2.0;
      if (lam >= P.lambda_max) { status = GTO_STATUS_STALLED; break; }
    }
  }
  // ---- results ----
  if (act)
    for (int k = 0; k < nopt; ++k) P.Qx[((long long)b * P.n + lane) * nopt + k] = qx[k];
  // occupancy count of the robot at qc seen from the new base: p' = Rz(theta)^T (p - (x, y, 0))
  double coll = 0.0;
  if (P.occ != nullptr) {
    double s, c;
    sincos(yv[2], &s, &c);
    for (int i = lane; i < P.npoints; i += 32) {
      const double px = P.wp[3 * i] - yv[0], py = P.wp[3 * i + 1] - yv[1];
      const double ux = c * px + s * py, uy = -s * px + c * py;
      const double fx = floor((ux - P.oox) / P.ores), fyy = floor((uy - P.ooy) / P.ores);
      const int ix = (int)fmin(fmax(fx, 0.0), (double)(P.onx - 1)), iy = (int)fmin(fmax(fyy, 0.0), (double)(P.ony - 1));
      coll += (double)P.occ[(long long)ix * P.ony + iy];
    }
    coll = warp_sum(coll);
  }
  if (lane == 0) {
    P.y[3 * b] = yv[0]; P.y[3 * b + 1] = yv[1]; P.y[3 * b + 2] = yv[2];
    P.cost[b] = F;
    P.collision[b] = coll;
    P.iters[b] = it;
    P.status[b] = status;
  }
}

// ==================================================================================================================
// k_base_place_sm: the default kernel.  Same arithmetic as k_base_place above (kept as the A/B reference, GTO_BASE_V1=1),
// but (1) the per-goal arrays (current and trial Gram matrices, gradients, and one scratch area shared by the linearisation's
// E / twist tables and the step's H_i / Z / C_i) live in shared memory, element-major with a 32-lane stride, so a lane's
// accesses are conflict-free and nothing spills to local memory / DRAM; (2) a warp carries floor(32 / n) problems, one per
// group of n consecutive lanes: the warp sums become group sums (n shuffles, fixed order), and a group that has finished
// idles (its state is frozen) until the slowest group of the warp is done.
// ==================================================================================================================
struct LaneArr {  // element e of this lane's array
  double* p;
  __device__ __forceinline__ double& operator[](int e) const { return p[e * 32]; }
};

template <int NP>
__host__ __device__ constexpr int base_sm_doubles_per_lane() {
  constexpr int NV = NP + 3;
  constexpr int s1 = 12 * NV, s2 = NP * NP + 7 * NP;
  return NV * (NV + 1) / 2 + NV + (s1 > s2 ? s1 : s2);
}

// packed upper triangle of a symmetric NV x NV matrix
template <int NV>
__device__ __forceinline__ int sym_ix(int a, int b) {
  const int lo = a < b ? a : b, hi = a < b ? b : a;
  return lo * NV - (lo * (lo - 1)) / 2 + (hi - lo);
}

__device__ __forceinline__ double group_sum(double v, int gbase, int n) {
  double s = 0.0;
  for (int j = 0; j < n; ++j) s += __shfl_sync(0xffffffffu, v, (gbase + j) & 31);
  return s;
}
__device__ __forceinline__ double group_max(double v, int gbase, int n) {
  double s = 0.0;
  for (int j = 0; j < n; ++j) s = fmax(s, __shfl_sync(0xffffffffu, v, (gbase + j) & 31));
  return s;
}

// linearisation of one goal (see base_goal_linearize); G (packed), g [NV] and the scratch area are the lane's shared-memory arrays.
// JAC = false: cost only (no shared memory is touched) -- the trial point of an iteration, which is linearised only if accepted.
template <int NP, int NOPT, bool JAC>
__device__ __noinline__ void base_goal_linearize_sm(const BaseParams& P, const double* __restrict__ qx, const double* __restrict__ yv,
                                                    const double* __restrict__ A, double& cost, LaneArr g, LaneArr G, LaneArr scr) {
  constexpr int NV = NP + 3;
  const RobotDev& R = *P.robot;
  const int nopt = NOPT > 0 ? NOPT : R.nopt, nv = nopt + 3;
  const LaneArr E = scr;  // [NV][12]; until the gripper frame is known, E_k[0..5] holds the joint's twist (omega, m)
  unsigned on_chain = 0u;
  double T[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
  for (int c = 0; c < P.nchain; ++c) {
    const int j = P.chain[c];
    double U[12];
    mul34(T, R.mov_origin_d[j], U);
    const double ax = R.mov_axis_d[j][0], ay = R.mov_axis_d[j][1], az = R.mov_axis_d[j][2];
    const double zx = U[0] * ax + U[1] * ay + U[2] * az, zy = U[4] * ax + U[5] * ay + U[6] * az, zz = U[8] * ax + U[9] * ay + U[10] * az;
    const int k = R.mov_opt[j];
    const double qj = k >= 0 ? qx[k] : P.qc[R.mov_qidx[j]];
    if (R.mov_type[j] == GTO_JOINT_REVOLUTE) {
      if (JAC && k >= 0) {
        on_chain |= 1u << k;
        E[12 * k] = zx; E[12 * k + 1] = zy; E[12 * k + 2] = zz;
        E[12 * k + 3] = U[7] * zz - U[11] * zy;  // o x z
        E[12 * k + 4] = U[11] * zx - U[3] * zz;
        E[12 * k + 5] = U[3] * zy - U[7] * zx;
      }
      double s, cs;
      sincos(qj, &s, &cs);
      const double v = 1.0 - cs;
      double M[12];
      M[0] = 1.0 - v * (ay * ay + az * az); M[1] = -s * az + v * ax * ay; M[2] = s * ay + v * ax * az; M[3] = 0.0;
      M[4] = s * az + v * ax * ay; M[5] = 1.0 - v * (ax * ax + az * az); M[6] = -s * ax + v * ay * az; M[7] = 0.0;
      M[8] = -s * ay + v * ax * az; M[9] = s * ax + v * ay * az; M[10] = 1.0 - v * (ax * ax + ay * ay); M[11] = 0.0;
      mul34(U, M, T);
    } else {
      if (JAC && k >= 0) {
        on_chain |= 1u << k;
        E[12 * k] = 0.0; E[12 * k + 1] = 0.0; E[12 * k + 2] = 0.0;
        E[12 * k + 3] = zx; E[12 * k + 4] = zy; E[12 * k + 5] = zz;
      }
#pragma unroll
      for (int e = 0; e < 12; ++e) T[e] = U[e];
      T[3] += qj * zx; T[7] += qj * zy; T[11] += qj * zz;
    }
  }
  double F[12];
  mul34(T, R.grip_tf_d, F);
  double s, c;
  sincos(yv[2], &s, &c);
  double D[12], RA[12];
#pragma unroll
  for (int col = 0; col < 4; ++col) {
    const double a0 = A[col], a1 = A[4 + col], a2 = A[8 + col];
    const double m0 = c * a0 - s * a1 + (col == 3 ? yv[0] : 0.0);
    const double m1 = s * a0 + c * a1 + (col == 3 ? yv[1] : 0.0);
    D[col] = F[col] - m0;
    D[4 + col] = F[4 + col] - m1;
    D[8 + col] = F[8 + col] - a2;
    RA[col] = -s * a0 - c * a1;
    RA[4 + col] = c * a0 - s * a1;
    RA[8 + col] = 0.0;
  }
  double DM[12];
  double cs = 0.0;
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int col = 0; col < 4; ++col) {
      double v = 0.0;
#pragma unroll
      for (int m = 0; m < 4; ++m) v += D[4 * r + m] * P.mom[4 * m + col];
      DM[4 * r + col] = v;
      cs += v * D[4 * r + col];
    }
  cost = cs;
  if constexpr (JAC) {
#pragma unroll
  for (int k = 0; k < nopt; ++k) {
    if (!((on_chain >> k) & 1u)) {
#pragma unroll
      for (int e = 0; e < 12; ++e) E[12 * k + e] = 0.0;
      continue;
    }
    const double wx = E[12 * k], wy = E[12 * k + 1], wz = E[12 * k + 2];
    const double m0 = E[12 * k + 3], m1 = E[12 * k + 4], m2 = E[12 * k + 5];
#pragma unroll
    for (int col = 0; col < 4; ++col) {
      const double f0 = F[col], f1 = F[4 + col], f2 = F[8 + col];
      E[12 * k + col] = wy * f2 - wz * f1 + (col == 3 ? m0 : 0.0);
      E[12 * k + 4 + col] = wz * f0 - wx * f2 + (col == 3 ? m1 : 0.0);
      E[12 * k + 8 + col] = wx * f1 - wy * f0 + (col == 3 ? m2 : 0.0);
    }
  }
#pragma unroll
  for (int e = 0; e < 12; ++e) { E[12 * nopt + e] = 0.0; E[12 * (nopt + 1) + e] = 0.0; E[12 * (nopt + 2) + e] = -RA[e]; }
  E[12 * nopt + 3] = -1.0;
  E[12 * (nopt + 1) + 7] = -1.0;
#pragma unroll
  for (int a = 0; a < nv; ++a) {
    double Ea[12], EM[12];
#pragma unroll
    for (int e = 0; e < 12; ++e) Ea[e] = E[12 * a + e];
    double v = 0.0;
#pragma unroll
    for (int e = 0; e < 12; ++e) v += Ea[e] * DM[e];
    g[a] = v;
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int col = 0; col < 4; ++col) {
        double u = 0.0;
#pragma unroll
        for (int m = 0; m < 4; ++m) u += Ea[4 * r + m] * P.mom[4 * m + col];
        EM[4 * r + col] = u;
      }
#pragma unroll
    for (int b = a; b < nv; ++b) {
      double u = 0.0;
#pragma unroll
      for (int e = 0; e < 12; ++e) u += EM[e] * E[12 * b + e];
      G[sym_ix<NV>(a, b)] = u;
    }
  }
  }
}

// NOPT > 0: the number of optimised joints is a compile-time constant (loops fully unrolled); NOPT = 0: read from the robot table
template <int NP, int NOPT>
__global__ void __launch_bounds__(32) k_base_place_sm(const __grid_constant__ BaseParams P) {
  constexpr int NV = NP + 3;
  extern __shared__ double base_sm[];
  const RobotDev& R = *P.robot;
  const int lane = threadIdx.x, n = P.n, gpw = 32 / n;
  const int grp = lane / n, gbase = grp * n, gl = lane - gbase;
  const int b = blockIdx.x * gpw + grp;
  const int nopt = NOPT > 0 ? NOPT : R.nopt;
  const bool act = grp < gpw && b < P.B;
  const double BIG = 1e30, PI = 3.14159265358979323846;
  const double ylo[3] = {-BIG, -BIG, -PI}, yhi[3] = {BIG, BIG, PI};
  constexpr int NG = NV * (NV + 1) / 2;  // packed symmetric Gram matrix
  const LaneArr G = {base_sm + lane}, g = {base_sm + NG * 32 + lane};
  const LaneArr scr = {base_sm + (NG + NV) * 32 + lane};
  const LaneArr Hd = scr, Z = {scr.p + NP * NP * 32}, Cm = {scr.p + (NP * NP + 4 * NP) * 32};
  double A[12];
#pragma unroll
  for (int e = 0; e < 12; ++e) A[e] = act ? P.goal[((long long)b * n + gl) * 12 + e] : 0.0;
  double qx[NP], qn[NP], dq[NP], yv[3] = {0.0, 0.0, 0.0}, yn[3];
#pragma unroll
  for (int k = 0; k < NP; ++k) qx[k] = k < nopt ? P.qc[R.opt_qidx[k]] : 0.0;
  for (int a = 0; a < NV; ++a) g[a] = 0.0;
  for (int e = 0; e < NG; ++e) G[e] = 0.0;
  double ci = 0.0;
  if (act) base_goal_linearize_sm<NP, NOPT, true>(P, qx, yv, A, ci, g, G, scr);
  double F = group_sum(act ? ci : 0.0, gbase, n);
  double lam = P.lambda0, nu = 2.0;
  int status = GTO_STATUS_MAX_ITER, it = 0;
  bool done = !act;
  while (!__all_sync(0xffffffffu, done || it >= P.max_iter)) {
    bool live = !done && it < P.max_iter;
    // ---- base block and gradient (group sums), active sets ----
    double S[3][3], gy[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      gy[a] = group_sum(g[nopt + a], gbase, n) + P.w_effort * yv[a];
#pragma unroll
      for (int c = a; c < 3; ++c) {
        S[a][c] = group_sum(G[sym_ix<NV>(nopt + a, nopt + c)], gbase, n) + (a == c ? P.w_effort : 0.0);
        S[c][a] = S[a][c];
      }
    }
    bool fy[3];
    unsigned fq = 0u;
    double pgmax = 0.0;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      fy[a] = (yv[a] <= ylo[a] + P.bound_eps && gy[a] > 0.0) || (yv[a] >= yhi[a] - P.bound_eps && gy[a] < 0.0);
      if (!fy[a]) pgmax = fmax(pgmax, fabs(gy[a]));
    }
    double pgq = 0.0;
#pragma unroll
    for (int k = 0; k < NP; ++k) {
      if (k < nopt) {
        const double gk = g[k];
        const bool f = (qx[k] <= R.lo[k] + P.bound_eps && gk > 0.0) || (qx[k] >= R.hi[k] - P.bound_eps && gk < 0.0) ||
                       !((R.grip_optmask >> k) & 1u) || !act;
        if (f) fq |= 1u << k;
        else pgq = fmax(pgq, fabs(gk));
      }
    }
    pgmax = 2.0 * fmax(pgmax, group_max(pgq, gbase, n));
    if (live && pgmax <= P.tol_grad) { status = GTO_STATUS_CONVERGED; done = true; live = false; }
    // ---- per-goal solve H_i Z = [C_i | -g_i] (Cholesky in shared memory), Schur complement on the base block ----
#pragma unroll
    for (int k = 0; k < nopt; ++k) {
      const bool fk = (fq >> k) & 1u;
#pragma unroll
      for (int l = 0; l < nopt; ++l) Hd[k * NP + l] = (fk || ((fq >> l) & 1u)) ? 0.0 : G[sym_ix<NV>(k, l)];
      const double gkk = G[sym_ix<NV>(k, k)];
      Hd[k * NP + k] = fk ? 1.0 : gkk + lam * gkk;
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        const double cv = (fk || fy[a]) ? 0.0 : G[sym_ix<NV>(k, nopt + a)];
        Cm[3 * k + a] = cv;
        Z[4 * k + a] = cv;
      }
      Z[4 * k + 3] = fk ? 0.0 : -g[k];
    }
#pragma unroll
    for (int j = 0; j < nopt; ++j) {
      double d = Hd[j * NP + j];
#pragma unroll
      for (int k = 0; k < j; ++k) { const double v = Hd[j * NP + k]; d -= v * v; }
      d = sqrt(d);
      Hd[j * NP + j] = d;
      const double inv = 1.0 / d;
#pragma unroll
      for (int i = j + 1; i < nopt; ++i) {
        double v = Hd[i * NP + j];
#pragma unroll
        for (int k = 0; k < j; ++k) v -= Hd[i * NP + k] * Hd[j * NP + k];
        Hd[i * NP + j] = v * inv;
      }
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
#pragma unroll
      for (int i = 0; i < nopt; ++i) {
        double v = Z[4 * i + r];
#pragma unroll
        for (int k = 0; k < i; ++k) v -= Hd[i * NP + k] * Z[4 * k + r];
        Z[4 * i + r] = v / Hd[i * NP + i];
      }
#pragma unroll
      for (int i = nopt - 1; i >= 0; --i) {
        double v = Z[4 * i + r];
#pragma unroll
        for (int k = i + 1; k < nopt; ++k) v -= Hd[k * NP + i] * Z[4 * k + r];
        Z[4 * i + r] = v / Hd[i * NP + i];
      }
    }
    double M[3][4];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      double v = 0.0;
#pragma unroll
      for (int k = 0; k < nopt; ++k) v += Cm[3 * k + a] * Z[4 * k + 3];
      M[a][3] = (fy[a] ? 0.0 : -gy[a]) - group_sum(v, gbase, n);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        double u = 0.0;
#pragma unroll
        for (int k = 0; k < nopt; ++k) u += Cm[3 * k + a] * Z[4 * k + c];
        const double base = (fy[a] || fy[c]) ? (a == c ? 1.0 : 0.0) : S[a][c] + (a == c ? lam * S[a][a] : 0.0);
        M[a][c] = base - group_sum(u, gbase, n);
      }
    }
    double dy[3];
#pragma unroll
    for (int p = 0; p < 3; ++p) {
      const double inv = 1.0 / M[p][p];
#pragma unroll
      for (int r = p + 1; r < 3; ++r) {
        const double f = M[r][p] * inv;
#pragma unroll
        for (int c = p; c < 4; ++c) M[r][c] -= f * M[p][c];
      }
    }
#pragma unroll
    for (int p = 2; p >= 0; --p) {
      double v = M[p][3];
#pragma unroll
      for (int c = p + 1; c < 3; ++c) v -= M[p][c] * dy[c];
      dy[p] = v / M[p][p];
    }
    // ---- projected trial point ----
    double stepm = 0.0;
#pragma unroll
    for (int k = 0; k < NP; ++k) {
      if (k < nopt) {
        const double d = Z[4 * k + 3] - (Z[4 * k] * dy[0] + Z[4 * k + 1] * dy[1] + Z[4 * k + 2] * dy[2]);
        qn[k] = fmin(fmax(qx[k] + d, R.lo[k]), R.hi[k]);
        dq[k] = act ? qn[k] - qx[k] : 0.0;
        stepm = fmax(stepm, fabs(dq[k]));
      } else {
        qn[k] = 0.0;
        dq[k] = 0.0;
      }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      yn[a] = fmin(fmax(yv[a] + dy[a], ylo[a]), yhi[a]);
      dy[a] = yn[a] - yv[a];
    }
    stepm = fmax(group_max(stepm, gbase, n), fmax(fabs(dy[0]), fmax(fabs(dy[1]), fabs(dy[2]))));
    // ---- predicted reduction with the undamped model ----
    double pq = 0.0, cd[3] = {0.0, 0.0, 0.0};
#pragma unroll
    for (int k = 0; k < NP; ++k) {
      if (k < nopt) {
        const double c0 = G[sym_ix<NV>(k, nopt)], c1 = G[sym_ix<NV>(k, nopt + 1)], c2 = G[sym_ix<NV>(k, nopt + 2)];
        double ad = c0 * dy[0] + c1 * dy[1] + c2 * dy[2];
#pragma unroll
        for (int l = 0; l < NP; ++l)
          if (l < nopt) ad += G[sym_ix<NV>(k, l)] * dq[l];
        pq += g[k] * dq[k] + 0.5 * dq[k] * ad;
        cd[0] += c0 * dq[k]; cd[1] += c1 * dq[k]; cd[2] += c2 * dq[k];
      }
    }
    pq = group_sum(pq, gbase, n);
    double py = 0.0;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const double ady = S[a][0] * dy[0] + S[a][1] * dy[1] + S[a][2] * dy[2] + group_sum(cd[a], gbase, n);
      py += gy[a] * dy[a] + 0.5 * dy[a] * ady;
    }
    const double pred = -(pq + py);
    if (live) ++it;
    // ---- cost of the trial point, acceptance; an accepted point is linearised in place ----
    double ct = 0.0;
    if (act) base_goal_linearize_sm<NP, NOPT, false>(P, qn, yn, A, ct, g, G, scr);
    bool relin = false;
    const double Ft = group_sum(act ? ct : 0.0, gbase, n) + P.w_effort * (yn[0] * yn[0] + yn[1] * yn[1] + yn[2] * yn[2]);
    if (live && (!(Ft == Ft) || fabs(Ft) > 1e300)) { status = GTO_STATUS_NAN; done = true; live = false; }
    if (live) {
      const double ared = 0.5 * (F - Ft);
      if (pred > 0.0 && ared >= P.eta * pred) {
        const double rho = ared / pred;
        const double t = 2.0 * fmin(rho, 1.0) - 1.0;
#pragma unroll
        for (int k = 0; k < NP; ++k) qx[k] = qn[k];
#pragma unroll
        for (int a = 0; a < 3; ++a) yv[a] = yn[a];
        relin = true;
        F = Ft;
        lam = fmax(P.lambda_min, lam * fmax(1.0 / 3.0, 1.0 - t * t * t));
        nu = 2.0;
        if (stepm <= P.tol_step) { status = GTO_STATUS_CONVERGED; done = true; }
      } else {
        if (pred <= 0.0 && stepm <= P.tol_step) { status = GTO_STATUS_CONVERGED; done = true; }
        else {
          lam = fmin(P.lambda_max, lam * nu);
          nu *= 2.0;
          if (lam >= P.lambda_max) { status = GTO_STATUS_STALLED; done = true; }
        }
      }
    }
    if (relin && !done && act) base_goal_linearize_sm<NP, NOPT, true>(P, qx, yv, A, ct, g, G, scr);
  }
  // ---- results ----
  if (act)
#pragma unroll
    for (int k = 0; k < nopt; ++k) P.Qx[((long long)b * n + gl) * nopt + k] = qx[k];
  double coll = 0.0;
  if (P.occ != nullptr && act) {
    double s, c;
    sincos(yv[2], &s, &c);
    for (int i = gl; i < P.npoints; i += n) {
      const double px = P.wp[3 * i] - yv[0], py = P.wp[3 * i + 1] - yv[1];
      const double ux = c * px + s * py, uy = -s * px + c * py;
      const double fx = floor((ux - P.oox) / P.ores), fyy = floor((uy - P.ooy) / P.ores);
      const int ix = (int)fmin(fmax(fx, 0.0), (double)(P.onx - 1)), iy = (int)fmin(fmax(fyy, 0.0), (double)(P.ony - 1));
      coll += (double)P.occ[(long long)ix * P.ony + iy];
    }
  }
  coll = group_sum(coll, gbase, n);
  if (act && gl == 0) {
    P.y[3 * b] = yv[0]; P.y[3 * b + 1] = yv[1]; P.y[3 * b + 2] = yv[2];
    P.cost[b] = F;
    P.collision[b] = coll;
    P.iters[b] = it;
    P.status[b] = status;
  }
}
