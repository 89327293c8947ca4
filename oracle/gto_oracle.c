/*
 * gto_oracle.c -- C restatement of oracle/gto_oracle.py (float64, pthreads over problems).
 *
 * TEST INFRASTRUCTURE ONLY: this is the CPU baseline timed by bench.py (`cpu_baseline`, `--impl reference`) and a
 * second checker for tests/.  Nothing under grasptrajopt_b200/ links or loads it.
 *
 * It follows the NumPy oracle function by function (which in turn cites the reference, IRVLUTD/GraspTrajOpt @ 4703ba2):
 *   fk_movable / joint twists      optas/models.py:826-868, 1203-1268; optas/spatialmath.py:91-100
 *   link frames, world points      gto/gto_models.py:83-121; gto/gto_planner.py:111-128
 *   trilinear field lookup         SURVEY.md Appendix A (replaces gto/gto_models.py:174-187, gto/sdf_callback.py)
 *   residual blocks                gto/gto_planner.py:86-135
 *   constraints                    gto/gto_planner.py:59-72,138 (eliminated / projected)
 *   solve                          projected bundle Levenberg-Marquardt of solve_lm() / lm_step() (stands in for IPOPT,
 *                                  optas/solver.py:384-400): solve_masked, bundle_dual, solve_one
 * Parity pinning: checked against the NumPy oracle in tests/test_oracle_c.py (the NumPy oracle is the one pinned
 * against reference-generated golden vectors).  The reference's own CPU path (CasADi + IPOPT) cannot be built here.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <stdatomic.h>
#include <unistd.h>
#include "../include/gto_b200.h"

typedef struct {
  const float* cost; /* [nx][ny][nz] */
  int32_t nx, ny, nz;
  double ox, oy, oz, pitch;
} oracle_field;

static void mul34(const double* A, const double* B, double* C) {
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 4; ++c) {
      double s = A[r * 4 + 0] * B[c] + A[r * 4 + 1] * B[4 + c] + A[r * 4 + 2] * B[8 + c];
      if (c == 3) s += A[r * 4 + 3];
      C[r * 4 + c] = s;
    }
}

static void fk_movable(const gto_robot_desc* R, const double* q, double* Tm /* [nmov][12] */) {
  for (int j = 0; j < R->nmov; ++j) {
    const double qj = q[R->mov_qidx[j]];
    const double* ax = R->mov_axis + 3 * j;
    double M[12], A[12];
    if (R->mov_type[j] == GTO_JOINT_REVOLUTE) {
      const double s = sin(qj), c = cos(qj), v = 1.0 - c, x = ax[0], y = ax[1], z = ax[2];
      M[0] = 1.0 - v * (y * y + z * z); M[1] = -s * z + v * x * y; M[2] = s * y + v * x * z; M[3] = 0;
      M[4] = s * z + v * x * y; M[5] = 1.0 - v * (x * x + z * z); M[6] = -s * x + v * y * z; M[7] = 0;
      M[8] = -s * y + v * x * z; M[9] = s * x + v * y * z; M[10] = 1.0 - v * (x * x + y * y); M[11] = 0;
    } else {
      M[0] = 1; M[1] = 0; M[2] = 0; M[3] = qj * ax[0];
      M[4] = 0; M[5] = 1; M[6] = 0; M[7] = qj * ax[1];
      M[8] = 0; M[9] = 0; M[10] = 1; M[11] = qj * ax[2];
    }
    mul34(R->mov_origin + 12 * j, M, A);
    if (R->mov_parent[j] < 0) memcpy(Tm + 12 * j, A, sizeof(A));
    else mul34(Tm + 12 * R->mov_parent[j], A, Tm + 12 * j);
  }
}

static void trilinear(const oracle_field* f, double wx, double wy, double wz, double* val, double* g) {
  const double u[3] = {(wx - f->ox) / f->pitch, (wy - f->oy) / f->pitch, (wz - f->oz) / f->pitch};
  const int N[3] = {f->nx, f->ny, f->nz};
  int i[3];
  double fr[3];
  int inb[3];
  for (int a = 0; a < 3; ++a) {
    int k = (int)floor(u[a]);
    if (k < 0) k = 0;
    if (k > N[a] - 2) k = N[a] - 2;
    i[a] = k;
    double t = u[a] - k;
    inb[a] = (t >= 0.0 && t <= 1.0);
    fr[a] = t < 0 ? 0 : (t > 1 ? 1 : t);
  }
  const long sy = f->nz, sx = (long)f->ny * f->nz;
  const float* p = f->cost + i[0] * sx + i[1] * sy + i[2];
  const double c000 = p[0], c001 = p[1], c010 = p[sy], c011 = p[sy + 1], c100 = p[sx], c101 = p[sx + 1], c110 = p[sx + sy],
               c111 = p[sx + sy + 1];
  const double d00 = c001 - c000, d01 = c011 - c010, d10 = c101 - c100, d11 = c111 - c110;
  const double z00 = c000 + fr[2] * d00, z01 = c010 + fr[2] * d01, z10 = c100 + fr[2] * d10, z11 = c110 + fr[2] * d11;
  const double y0 = z00 + fr[1] * (z01 - z00), y1 = z10 + fr[1] * (z11 - z10);
  *val = y0 + fr[0] * (y1 - y0);
  const double dy0 = z01 - z00, dy1 = z11 - z10;
  const double dz0 = d00 + fr[1] * (d01 - d00), dz1 = d10 + fr[1] * (d11 - d10);
  g[0] = inb[0] ? (y1 - y0) / f->pitch : 0.0;
  g[1] = inb[1] ? (dy0 + fr[0] * (dy1 - dy0)) / f->pitch : 0.0;
  g[2] = inb[2] ? (dz0 + fr[0] * (dz1 - dz0)) / f->pitch : 0.0;
}

typedef struct {
  const gto_robot_desc* R;
  const oracle_field* fields;
  const gto_batch_in* in;
  int b;
  double* Tm;  /* scratch [nmov][12] */
  double* om;  /* [nopt][3] */
  double* mm;  /* [nopt][3] */
} prob_ctx;

/* H [T][n][n], g [T][n], cost [T] at trajectory Q [T][ndof]; t_lo: first knot to (re)compute */
static void linearize(const prob_ctx* P, const double* Q, int t_lo, double* H, double* g, double* cost) {
  const gto_robot_desc* R = P->R;
  const gto_batch_in* in = P->in;
  const int T = in->T, n = R->nopt, nd = R->ndof, b = P->b;
  const double sw = sqrt(in->w_obs), sg = sqrt(in->w_goal);
  const int ks = T + in->standoff_offset;
  const double* base = in->base_position ? in->base_position + 3 * b : NULL;
  for (int t = t_lo; t < T; ++t) {
    double* Ht = H + (size_t)t * n * n;
    double* gt = g + (size_t)t * n;
    memset(Ht, 0, sizeof(double) * n * n);
    memset(gt, 0, sizeof(double) * n);
    double ct = 0.0;
    fk_movable(R, Q + (size_t)t * nd, P->Tm);
    for (int k = 0; k < n; ++k) P->om[3 * k] = P->om[3 * k + 1] = P->om[3 * k + 2] = P->mm[3 * k] = P->mm[3 * k + 1] = P->mm[3 * k + 2] = 0.0;
    for (int j = 0; j < R->nmov; ++j) {
      const int k = R->mov_opt[j];
      if (k < 0) continue;
      const double* Tj = P->Tm + 12 * j;
      const double* ax = R->mov_axis + 3 * j;
      const double z[3] = {Tj[0] * ax[0] + Tj[1] * ax[1] + Tj[2] * ax[2], Tj[4] * ax[0] + Tj[5] * ax[1] + Tj[6] * ax[2],
                           Tj[8] * ax[0] + Tj[9] * ax[1] + Tj[10] * ax[2]};
      if (R->mov_type[j] == GTO_JOINT_REVOLUTE) {
        const double o[3] = {Tj[3], Tj[7], Tj[11]};
        P->om[3 * k] = z[0]; P->om[3 * k + 1] = z[1]; P->om[3 * k + 2] = z[2];
        P->mm[3 * k] = o[1] * z[2] - o[2] * z[1]; P->mm[3 * k + 1] = o[2] * z[0] - o[0] * z[2]; P->mm[3 * k + 2] = o[0] * z[1] - o[1] * z[0];
      } else {
        P->mm[3 * k] = z[0]; P->mm[3 * k + 1] = z[1]; P->mm[3 * k + 2] = z[2];
      }
    }
    double J[GTO_MAX_OPT];
    if (in->collision_avoidance) {
      const int fid = (t < ks) ? in->field_all[b] : in->field_obs[b];
      if (fid >= 0) {
        const oracle_field* f = P->fields + fid;
        for (int l = 0; l < R->nlinks; ++l) {
          double F[12];
          if (R->link_mov[l] < 0) memcpy(F, R->link_tf + 12 * l, sizeof(F));
          else mul34(P->Tm + 12 * R->link_mov[l], R->link_tf + 12 * l, F);
          const unsigned mask = R->link_optmask[l];
          for (int i = R->link_pt_start[l]; i < R->link_pt_start[l] + R->link_pt_count[l]; ++i) {
            const double x = R->points[3 * i], y = R->points[3 * i + 1], zc = R->points[3 * i + 2];
            const double w[3] = {F[0] * x + F[1] * y + F[2] * zc + F[3], F[4] * x + F[5] * y + F[6] * zc + F[7], F[8] * x + F[9] * y + F[10] * zc + F[11]};
            double val, gr[3];
            trilinear(f, w[0] + (base ? base[0] : 0), w[1] + (base ? base[1] : 0), w[2] + (base ? base[2] : 0), &val, gr);
            const double r = sw * val;
            if (in->flags & GTO_FLAG_OBS_LINEAR) ct += sw * r;  /* unsquared term w*c (gto/ik_solver.py:69) */
            else ct += r * r;
            if (gr[0] == 0.0 && gr[1] == 0.0 && gr[2] == 0.0) continue;
            gr[0] *= sw; gr[1] *= sw; gr[2] *= sw;
            const double nx = w[1] * gr[2] - w[2] * gr[1], ny = w[2] * gr[0] - w[0] * gr[2], nz = w[0] * gr[1] - w[1] * gr[0];
            for (int k = 0; k < n; ++k)
              J[k] = ((mask >> k) & 1u) ? P->om[3 * k] * nx + P->om[3 * k + 1] * ny + P->om[3 * k + 2] * nz + P->mm[3 * k] * gr[0] + P->mm[3 * k + 1] * gr[1] + P->mm[3 * k + 2] * gr[2] : 0.0;
            if (in->flags & GTO_FLAG_OBS_LINEAR) {  /* half gradient (w/2) dc/dq, no Gauss-Newton curvature */
              for (int a = 0; a < n; ++a) gt[a] += 0.5 * sw * J[a];
              continue;
            }
            for (int a = 0; a < n; ++a) {
              if (J[a] == 0.0) continue;
              gt[a] += J[a] * r;
              for (int c = 0; c < n; ++c) Ht[a * n + c] += J[a] * J[c];
            }
          }
        }
      }
    }
    for (int which = 0; which < 2; ++which) {
      if (which == 0 && t != T - 1) continue;
      if (which == 1 && !(in->use_standoff && t == ks)) continue;
      double Fg[12];
      if (R->grip_mov < 0) memcpy(Fg, R->grip_tf, sizeof(Fg));
      else mul34(P->Tm + 12 * R->grip_mov, R->grip_tf, Fg);
      const double* M = in->goal_tf + (size_t)b * 24 + 12 * which;
      for (int i = R->grip_pt_start; i < R->grip_pt_start + R->grip_pt_count; ++i) {
        const double x = R->points[3 * i], y = R->points[3 * i + 1], zc = R->points[3 * i + 2];
        double w[3], rr[3];
        for (int a = 0; a < 3; ++a) {
          w[a] = Fg[a * 4] * x + Fg[a * 4 + 1] * y + Fg[a * 4 + 2] * zc + Fg[a * 4 + 3];
          rr[a] = sg * ((Fg[a * 4] - M[a * 4]) * x + (Fg[a * 4 + 1] - M[a * 4 + 1]) * y + (Fg[a * 4 + 2] - M[a * 4 + 2]) * zc + (Fg[a * 4 + 3] - M[a * 4 + 3]));
        }
        for (int a3 = 0; a3 < 3; ++a3) {
          for (int k = 0; k < n; ++k) {
            double v = 0.0;
            if ((R->grip_optmask >> k) & 1u) {
              const double* o = P->om + 3 * k;
              const double* m = P->mm + 3 * k;
              if (a3 == 0) v = o[1] * w[2] - o[2] * w[1] + m[0];
              else if (a3 == 1) v = o[2] * w[0] - o[0] * w[2] + m[1];
              else v = o[0] * w[1] - o[1] * w[0] + m[2];
            }
            J[k] = sg * v;
          }
          const double r = rr[a3];
          ct += r * r;
          for (int a = 0; a < n; ++a) {
            gt[a] += J[a] * r;
            for (int c = 0; c < n; ++c) Ht[a * n + c] += J[a] * J[c];
          }
        }
      }
    }
    cost[t] = ct;
  }
}

/* in-place Gauss-Jordan inverse of an SPD n x n block; returns 0 if a pivot is not positive */
static int gj_inverse(double* S, int n) {
  for (int k = 0; k < n; ++k) {
    const double piv = S[k * n + k];
    if (!(piv > 0.0)) return 0;
    const double ip = 1.0 / piv;
    for (int r = 0; r < n; ++r) {
      if (r == k) continue;
      const double f = S[r * n + k] * ip;
      for (int c = 0; c < n; ++c)
        if (c != k) S[r * n + c] -= f * S[k * n + c];
      S[r * n + k] = -f;
    }
    for (int c = 0; c < n; ++c)
      if (c != k) S[k * n + c] *= ip;
    S[k * n + k] = ip;
  }
  return 1;
}

static double velocity_cost(const gto_batch_in* in, const double* X, int n) {
  double s = 0.0;
  for (int i = 0; i < (in->T - 1) * n; ++i) {
    const double d = X[i + n] - X[i];
    s += d * d;
  }
  return in->w_vel / (in->dt * in->dt) * s;
}

typedef struct {
  const gto_robot_desc* R;
  const oracle_field* fields;
  const gto_batch_in* in;
  const gto_options* opt;
  gto_batch_out* out;
  atomic_int next;
  atomic_int failed;
} job_t;

/* Block-tridiagonal solve (block Thomas) of the masked, damped Gauss-Newton system with `nr` right-hand sides:
 * diagonal blocks (H_t + a2 c_t I)(1 + lam on the diagonal), couplings -a2 I, rows / columns of held variables replaced by
 * identity.  rhs, x: [nr][m*n]; vv: scratch [nr][m*n]; Sinv: scratch [m][n*n].  Returns 0 when a pivot block is not
 * positive definite. */
static int solve_masked(int m, int n, int T, const double* H /* [T][n][n] */, double a2, double lam, const unsigned char* fx,
                        double* Sinv, int nr, double* const* rhs, double* const* x, double* vv) {
  const int nn = n * n;
  double u[16];
  for (int i = 0; i < m; ++i) {
    const int t = i + 2;
    const double cnt = (t < T - 1) ? 2.0 : 1.0;
    double* S = Sinv + (size_t)i * nn;
    for (int r = 0; r < n; ++r)
      for (int c = 0; c < n; ++c) {
        const int fr = fx[i * n + r], fc = fx[i * n + c];
        double v = H[(size_t)t * nn + r * n + c];
        if (r == c) { v += a2 * cnt; v += lam * v; }
        if (fr || fc) v = (r == c) ? 1.0 : 0.0;
        if (i > 0) {
          const double cr = (fr || fx[(i - 1) * n + r]) ? 0.0 : a2, cc = (fc || fx[(i - 1) * n + c]) ? 0.0 : a2;
          v -= cr * cc * Sinv[(size_t)(i - 1) * nn + r * n + c];
        }
        S[r * n + c] = v;
      }
    if (!gj_inverse(S, n)) return 0;
    for (int q = 0; q < nr; ++q) {
      double* v = vv + (size_t)q * m * n;
      for (int r = 0; r < n; ++r) {
        double uu = rhs[q][i * n + r];
        if (i > 0) uu += ((fx[i * n + r] || fx[(i - 1) * n + r]) ? 0.0 : a2) * v[(i - 1) * n + r];
        u[r] = uu;
      }
      for (int r = 0; r < n; ++r) {
        double s = 0.0;
        for (int c = 0; c < n; ++c) s += S[r * n + c] * u[c];
        v[i * n + r] = s;
      }
    }
  }
  for (int q = 0; q < nr; ++q) {
    const double* v = vv + (size_t)q * m * n;
    double* xq = x[q];
    for (int i = m - 1; i >= 0; --i)
      for (int r = 0; r < n; ++r) {
        double s = v[i * n + r];
        if (i < m - 1)
          for (int c = 0; c < n; ++c)
            s += Sinv[(size_t)i * nn + r * n + c] * ((fx[i * n + c] || fx[(i + 1) * n + c]) ? 0.0 : a2) * xq[(i + 1) * n + c];
        xq[i * n + r] = s;
      }
  }
  return 1;
}

/* half gradient of f over the free knots 2..T-1: J^T r of the point rows + the analytic velocity term */
static void total_gradient(int T, int n, double a2, const double* X, const double* g, double* gt) {
  for (int i = 0; i < T - 2; ++i)
    for (int k = 0; k < n; ++k) {
      const int t = i + 2;
      double gv = X[t * n + k] - X[(t - 1) * n + k];
      if (t < T - 1) gv -= X[(t + 1) * n + k] - X[t * n + k];
      gt[i * n + k] = g[t * n + k] + a2 * gv;
    }
}

/* Dual of the bundle model: maximise  b.theta + theta' M theta / 2  over theta_1..K >= 0, sum <= 1 (theta_0 = 1 - sum is the
 * weight of the model at the standing point; row / column 0 of M and b_0 are zero).  Pairwise exchange (SMO): move weight
 * from the active piece with the smallest dual gradient to the piece with the largest.  Same loop in gto_oracle.py and
 * step_cr.cuh. */
static void bundle_dual(int K, const double* bq, double M[][GTO_BUNDLE_MAX + 1], double* theta) {
  theta[0] = 1.0;
  for (int k = 1; k <= K; ++k) theta[k] = 0.0;
  /* the dual gradients of the pieces with weight all vanish at an interior optimum: the stopping tolerance is relative to the
   * largest |b_k|, not to the gradients themselves (which would never pass it and always run into the iteration cap) */
  double scale = 0.0;
  for (int k = 1; k <= K; ++k) scale = fmax(scale, fabs(bq[k]));
  for (int iter = 0; iter < 12; ++iter) {
    double G[GTO_BUNDLE_MAX + 1];
    for (int k = 0; k <= K; ++k) {
      G[k] = bq[k];
      for (int j = 1; j <= K; ++j) G[k] += M[k][j] * theta[j];
    }
    int ib = 0, jb = -1;
    for (int k = 1; k <= K; ++k)
      if (G[k] > G[ib]) ib = k;
    for (int k = 0; k <= K; ++k)
      if (theta[k] > 0.0 && (jb < 0 || G[k] < G[jb])) jb = k;
    if (jb < 0 || ib == jb || G[ib] - G[jb] <= 1e-13 * scale) break;
    const double curv = -(M[ib][ib] - 2.0 * M[ib][jb] + M[jb][jb]);
    double delta = curv > 0.0 ? (G[ib] - G[jb]) / curv : theta[jb];
    if (delta > theta[jb]) delta = theta[jb];
    theta[ib] += delta;
    theta[jb] -= delta;
    if (theta[jb] < 1e-15) theta[jb] = 0.0;
  }
}

static void solve_one(job_t* J, int b) {
  const gto_robot_desc* R = J->R;
  const oracle_field* fields = J->fields;
  const gto_batch_in* in = J->in;
  const gto_options* opt = J->opt;
  gto_batch_out* out = J->out;
  const int T = in->T, n = R->nopt, nd = R->ndof, m = T - 2, nn = n * n, mn = m * n;
  const double a2 = in->w_vel / (in->dt * in->dt);
  const int KB = opt->bundle < 0 ? 0 : (opt->bundle > GTO_BUNDLE_MAX ? GTO_BUNDLE_MAX : opt->bundle);
  prob_ctx P;
  P.R = R; P.fields = fields; P.in = in; P.b = b;
  const size_t wsz = (size_t)R->nmov * 12 + 6 * n + 2 * ((size_t)T * nd + (size_t)T * nn + (size_t)T * n + T) + 2 * (size_t)T * n +
                     (size_t)m * nn + (size_t)mn * (5 + 3 * (GTO_BUNDLE_MAX + 1) + 2 * GTO_BUNDLE_MAX);
  double* W = (double*)calloc(wsz, sizeof(double));
  unsigned char* fx = (unsigned char*)calloc((size_t)mn + 1, 1);
  if (!W || !fx) {
    atomic_store(&J->failed, 1);
    free(W);
    free(fx);
    return;
  }
  double* w = W;
  P.Tm = w; w += (size_t)R->nmov * 12;
  P.om = w; w += 3 * n;
  P.mm = w; w += 3 * n;
  double* Q = w; w += (size_t)T * nd;
  double* Qt = w; w += (size_t)T * nd;
  double* H = w; w += (size_t)T * nn;
  double* Ht = w; w += (size_t)T * nn;
  double* g = w; w += (size_t)T * n;
  double* gtr = w; w += (size_t)T * n;
  double* cp = w; w += T;
  double* cpt = w; w += T;
  double* X = w; w += (size_t)T * n;
  double* Xt = w; w += (size_t)T * n;
  double* Sinv = w; w += (size_t)m * nn;
  double* gt = w; w += mn;    /* half gradient at the standing point */
  double* gtt = w; w += mn;   /* half gradient at the trial point */
  double* dd = w; w += mn;    /* clipped step */
  double* dfix = w; w += mn;  /* prescribed step of the variables held at a bound */
  double* sst = w; w += mn;   /* combined step before clipping */
  double* rhsb = w; w += (size_t)mn * (GTO_BUNDLE_MAX + 1);
  double* solb = w; w += (size_t)mn * (GTO_BUNDLE_MAX + 1);
  double* vvb = w; w += (size_t)mn * (GTO_BUNDLE_MAX + 1);
  double* gB = w; w += (size_t)mn * GTO_BUNDLE_MAX;   /* bundle: half gradients at the other points y_k */
  double* dyB = w; w += (size_t)mn * GTO_BUNDLE_MAX;  /* y_k - x */
  double FB[GTO_BUNDLE_MAX], eB[GTO_BUNDLE_MAX + 1];
  int nb = 0;
  double* rhs[GTO_BUNDLE_MAX + 1];
  double* sol[GTO_BUNDLE_MAX + 1];
  for (int q = 0; q <= GTO_BUNDLE_MAX; ++q) { rhs[q] = rhsb + (size_t)q * mn; sol[q] = solb + (size_t)q * mn; }
  /* initial trajectory: seed projected on the constraints */
  memcpy(Q, in->q_seed + (size_t)b * T * nd, sizeof(double) * T * nd);
  for (int t = 0; t < T; ++t)
    for (int k = 0; k < n; ++k) {
      const int j = R->opt_qidx[k];
      double v = Q[(size_t)t * nd + j];
      if (v < R->lo[k]) v = R->lo[k];
      if (v > R->hi[k]) v = R->hi[k];
      if (t < 2) v = in->qc[(size_t)b * nd + j];
      Q[(size_t)t * nd + j] = v;
      X[(size_t)t * n + k] = v;
    }
  linearize(&P, Q, 0, H, g, cp);
  double Fp = 0;
  for (int t = 0; t < T; ++t) Fp += cp[t];
  double F = Fp + velocity_cost(in, X, n);
  double lam = opt->lambda0, nu = 2.0;
  int status = GTO_STATUS_MAX_ITER, it = 0;
  double fhist[16];
  const int win = opt->slow_window > 15 ? 15 : opt->slow_window;
  fhist[0] = F;
  eB[0] = 0.0;
  while (it < opt->max_iter) {
    /* ---- gradient, active set, projected-gradient test ---- */
    double pgmax = 0.0;
    total_gradient(T, n, a2, X, g, gt);
    for (int i = 0; i < mn; ++i) {
      const int k = i % n;
      const double x = X[2 * n + i];
      const int fixed = (x <= R->lo[k] + opt->bound_eps && gt[i] > 0.0) || (x >= R->hi[k] - opt->bound_eps && gt[i] < 0.0);
      fx[i] = (unsigned char)fixed;
      if (!fixed && fabs(gt[i]) > pgmax) pgmax = fabs(gt[i]);
    }
    if (2.0 * pgmax <= opt->tol_grad) { status = GTO_STATUS_CONVERGED; break; }
    /* ---- bundle step: d_0 = Levenberg-Marquardt step, d_k = the same system solved for the gradient of bundle piece k;
     *      piece_k(s) = e_k + g_k.s in half-cost units relative to f/2 (e_k <= 0: its value at the standing point).
     * Active-set rounds (gto_options.as_rounds): a free variable that the step pushes beyond a joint limit is moved exactly onto
     * the limit (prescribed step dfix) and the remaining variables are re-solved with that step on the right-hand side. ---- */
    for (int i = 0; i < mn; ++i) dfix[i] = 0.0;
    int as_round = 0, ok;
    double theta[GTO_BUNDLE_MAX + 1];
  resolve:
    ok = 0;
    for (int attempt = 0; attempt < 8 && !ok; ++attempt) {
      for (int q = 0; q <= nb; ++q) {
        const double* gq = q == 0 ? gt : gB + (size_t)(q - 1) * mn;
        for (int i = 0; i < m; ++i)
          for (int r = 0; r < n; ++r) {
            const int t = i + 2;
            double uu;
            if (fx[i * n + r]) {
              uu = dfix[i * n + r];
            } else { /* free row: -g - (coupling to the prescribed steps of held variables) */
              uu = -gq[i * n + r];
              for (int c = 0; c < n; ++c)
                if (c != r && fx[i * n + c]) uu -= H[(size_t)t * nn + r * n + c] * dfix[i * n + c];
              if (i > 0 && fx[(i - 1) * n + r]) uu += a2 * dfix[(i - 1) * n + r];
              if (i < m - 1 && fx[(i + 1) * n + r]) uu += a2 * dfix[(i + 1) * n + r];
            }
            rhs[q][i * n + r] = uu;
          }
      }
      ok = solve_masked(m, n, T, H, a2, lam, fx, Sinv, nb + 1, rhs, sol, vvb);
      if (!ok) lam = fmin(opt->lambda_max, lam * 10.0);
    }
    if (!ok) { status = GTO_STATUS_NAN; break; }
    theta[0] = 1.0;
    for (int k = 1; k <= nb; ++k) theta[k] = 0.0;
    if (nb > 0) {
      double bq[GTO_BUNDLE_MAX + 1], M[GTO_BUNDLE_MAX + 1][GTO_BUNDLE_MAX + 1];
      for (int k = 0; k <= nb; ++k) { bq[k] = 0.0; for (int j = 0; j <= nb; ++j) M[k][j] = 0.0; }
      for (int k = 1; k <= nb; ++k) {
        const double* gk = gB + (size_t)(k - 1) * mn;
        double s = eB[k];
        for (int i = 0; i < mn; ++i) s += (gk[i] - gt[i]) * sol[0][i];
        bq[k] = s;
        for (int j = 1; j <= nb; ++j) {
          double mm_ = 0.0;
          for (int i = 0; i < mn; ++i) mm_ += (gk[i] - gt[i]) * (sol[j][i] - sol[0][i]);
          M[k][j] = mm_;
        }
      }
      for (int k = 1; k <= nb; ++k)
        for (int j = k + 1; j <= nb; ++j) { const double s = 0.5 * (M[k][j] + M[j][k]); M[k][j] = M[j][k] = s; }
      bundle_dual(nb, bq, M, theta);
    }
    for (int i = 0; i < mn; ++i) {
      double s = sol[0][i];
      for (int k = 1; k <= nb; ++k)
        if (theta[k] != 0.0) s += theta[k] * (sol[k][i] - sol[0][i]);
      sst[i] = s;
    }
    if (as_round < opt->as_rounds) {
      int nviol = 0;
      for (int i = 0; i < mn; ++i) {
        if (fx[i]) continue;
        const int r = i % n;
        const double xc = X[2 * n + i], xn = xc + sst[i];
        if (xn < R->lo[r]) { fx[i] = 1; dfix[i] = R->lo[r] - xc; ++nviol; }
        else if (xn > R->hi[r]) { fx[i] = 1; dfix[i] = R->hi[r] - xc; ++nviol; }
      }
      if (nviol) { ++as_round; goto resolve; }
    }
    /* ---- trial point = clip(X + s); predicted reduction with the undamped, unmasked bundle model ---- */
    double stepmax = 0.0;
    memcpy(Xt, X, sizeof(double) * T * n);
    for (int i = 0; i < mn; ++i) {
      const int r = i % n;
      const double xc = X[2 * n + i];
      double xn = xc + sst[i];
      if (xn < R->lo[r]) xn = R->lo[r];
      if (xn > R->hi[r]) xn = R->hi[r];
      Xt[2 * n + i] = xn;
      dd[i] = xn - xc;
      if (fabs(dd[i]) > stepmax) stepmax = fabs(dd[i]);
    }
    double lin = 0.0, quad = 0.0;
    for (int i = 0; i < mn; ++i) lin += gt[i] * dd[i];
    for (int k = 1; k <= nb; ++k) {
      const double* gk = gB + (size_t)(k - 1) * mn;
      double s = eB[k];
      for (int i = 0; i < mn; ++i) s += gk[i] * dd[i];
      if (s > lin) lin = s;
    }
    for (int i = 0; i < m; ++i) {
      const int t = i + 2;
      const double dg = a2 * ((t < T - 1) ? 2.0 : 1.0);
      for (int r = 0; r < n; ++r) {
        double hd = dg * dd[i * n + r];
        for (int c = 0; c < n; ++c) hd += H[(size_t)t * nn + r * n + c] * dd[i * n + c];
        quad += dd[i * n + r] * hd;
        if (i < m - 1) quad -= 2.0 * a2 * dd[i * n + r] * dd[(i + 1) * n + r];
      }
    }
    const double pred = -(lin + 0.5 * quad);
    it += 1;
    /* ---- evaluate the trial ---- */
    memcpy(Qt, Q, sizeof(double) * T * nd);
    for (int t = 2; t < T; ++t)
      for (int k = 0; k < n; ++k) Qt[(size_t)t * nd + R->opt_qidx[k]] = Xt[t * n + k];
    cpt[0] = cp[0]; cpt[1] = cp[1];
    linearize(&P, Qt, 2, Ht, gtr, cpt);
    double Fp_t = 0;
    for (int t = 0; t < T; ++t) Fp_t += cpt[t];
    const double Ft = Fp_t + velocity_cost(in, Xt, n);
    if (!isfinite(Ft)) { status = GTO_STATUS_NAN; break; }
    const double ared = 0.5 * (F - Ft), noise = opt->noise_rel * fmax(Fp, Fp_t);
    const int acc = pred > 0.0 && ared + noise >= opt->eta * pred;
    /* ---- bundle update: the point we do not stand on after this decision becomes a cutting plane.  Kept pieces are re-based to
     *      the new standing point; when the bundle is full the least active piece (most negative value e_k there) is replaced ---- */
    if (KB > 0) {
      const double Fnew = acc ? Ft : F;
      int slot = nb;
      double worst = 0.0;
      for (int k = 0; k < nb; ++k) {
        double* dk = dyB + (size_t)k * mn;
        const double* gk = gB + (size_t)k * mn;
        double s1 = 0.0;
        if (acc)
          for (int i = 0; i < mn; ++i) dk[i] -= dd[i];
        double rk = 0.0;
        for (int i = 0; i < mn; ++i) {
          s1 += gk[i] * dk[i];
          if (fabs(dk[i]) > rk) rk = fabs(dk[i]);
        }
        /* a piece further than bundle_radius from the standing point is not used (and is the first to be replaced) */
        eB[k + 1] = rk > opt->bundle_radius ? -1e300 : -fabs(0.5 * (FB[k] - Fnew) - s1);
        if (nb >= KB && (k == 0 || eB[k + 1] < worst)) { worst = eB[k + 1]; slot = k; }
      }
      double* gs = gB + (size_t)slot * mn;
      double* ds = dyB + (size_t)slot * mn;
      double s1 = 0.0;
      if (acc) {
        for (int i = 0; i < mn; ++i) { gs[i] = gt[i]; ds[i] = -dd[i]; s1 += gs[i] * ds[i]; }
        FB[slot] = F;
      } else {
        total_gradient(T, n, a2, Xt, gtr, gtt);
        for (int i = 0; i < mn; ++i) { gs[i] = gtt[i]; ds[i] = dd[i]; s1 += gs[i] * ds[i]; }
        FB[slot] = Ft;
      }
      eB[slot + 1] = stepmax > opt->bundle_radius ? -1e300 : -fabs(0.5 * (FB[slot] - Fnew) - s1);
      if (nb < KB) ++nb;
    }
    if (acc) {
      const double rho = ared / pred, lam_used = lam, F_before = F;
      double* tmp;
      tmp = Q; Q = Qt; Qt = tmp;
      tmp = X; X = Xt; Xt = tmp;
      tmp = H; H = Ht; Ht = tmp;
      tmp = g; g = gtr; gtr = tmp;
      tmp = cp; cp = cpt; cpt = tmp;
      F = Ft; Fp = Fp_t;
      /* gain ratio clamped to [0, 1]: a step accepted only thanks to the noise allowance can have rho << 0, and Nielsen's
       * cubic would then multiply the damping by hundreds in one step */
      const double ww = 2.0 * fmin(fmax(rho, 0.0), 1.0) - 1.0;
      lam = fmax(opt->lambda_min, lam * fmax(1.0 / 3.0, 1.0 - ww * ww * ww));
      nu = 2.0;
      /* a small step certifies a stationary point only when it was (nearly) the undamped step of the bundle model; under heavy
       * damping the iterate rests on gradient jumps of the trilinear field that the bundle does not resolve (GTO_STATUS_SLOW) */
      if (stepmax <= opt->tol_step) { status = (lam_used <= opt->lambda_conv) ? GTO_STATUS_CONVERGED : GTO_STATUS_SLOW; break; }
      if (lam_used >= opt->lambda_slow && ared <= opt->ftol * F_before) { status = GTO_STATUS_SLOW; break; }
    } else {
      if (pred <= 0.0 && stepmax <= opt->tol_step) { status = (lam <= opt->lambda_conv) ? GTO_STATUS_CONVERGED : GTO_STATUS_SLOW; break; }
      lam = fmin(opt->lambda_max, fmax(lam * nu, opt->lambda_reject));
      nu *= 2.0;
      if (lam >= opt->lambda_max) { status = GTO_STATUS_STALLED; break; }
    }
    if (win > 0) {  /* windowed progress test on the accepted cost (acceptable-level termination, same bookkeeping as k_step_cr) */
      if (it >= win && fhist[(it - win) & 15] - F <= opt->slow_ftol * F) { status = GTO_STATUS_SLOW; break; }
      fhist[it & 15] = F;
    }
  }
  /* ---- unpack ---- */
  if (out->Q) memcpy(out->Q + (size_t)b * T * nd, Q, sizeof(double) * T * nd);
  if (out->dQ) {
    double* dQ = out->dQ + (size_t)b * (T - 1) * nd;
    memset(dQ, 0, sizeof(double) * (T - 1) * nd);
    for (int t = 0; t < T - 1; ++t)
      for (int k = 0; k < n; ++k) dQ[(size_t)t * nd + R->opt_qidx[k]] = (X[(t + 1) * n + k] - X[t * n + k]) / in->dt;
  }
  if (out->cost) out->cost[b] = F;
  if (out->iters) out->iters[b] = it;
  if (out->status) out->status[b] = status;
  free(W);
  free(fx);
}

static void* worker(void* arg) {
  job_t* J = (job_t*)arg;
  for (;;) {
    const int b = atomic_fetch_add(&J->next, 1);  /* dynamic schedule: problems differ in iteration count */
    if (b >= J->in->B) break;
    solve_one(J, b);
  }
  return NULL;
}

int oracle_num_threads(void) {
  long n = sysconf(_SC_NPROCESSORS_ONLN);
  return n > 0 ? (int)n : 1;
}

int oracle_solve_batch(const gto_robot_desc* R, const oracle_field* fields, const gto_batch_in* in, const gto_options* opt,
                       gto_batch_out* out, int nthreads) {
  job_t J;
  J.R = R; J.fields = fields; J.in = in; J.opt = opt; J.out = out;
  atomic_init(&J.next, 0);
  atomic_init(&J.failed, 0);
  if (nthreads <= 0) nthreads = oracle_num_threads();
  if (nthreads > in->B) nthreads = in->B;
  if (nthreads <= 1) {
    worker(&J);
  } else {
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * nthreads);
    int started = 0;
    for (int i = 0; i < nthreads; ++i)
      if (pthread_create(&th[i], NULL, worker, &J) == 0) ++started;
      else break;
    if (started == 0) worker(&J);
    for (int i = 0; i < started; ++i) pthread_join(th[i], NULL);
    free(th);
  }
  return atomic_load(&J.failed) ? GTO_ERR_NOMEM : GTO_OK;
}
