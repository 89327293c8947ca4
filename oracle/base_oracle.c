/*
 * base_oracle.c -- C restatement of the base-placement oracle (oracle/base_oracle.py; reference gto/base_planner.py:35-168).
 *
 * TEST INFRASTRUCTURE / CPU BASELINE ONLY: compiled into oracle/libgto_oracle.so, loaded by oracle/c_oracle.py from tests/ and
 * from the CPU-baseline legs of the benchmark tools.  Nothing in grasptrajopt_b200/ may link or load it.
 *
 * Same projected Levenberg-Marquardt iteration as base_oracle.solve_base, float64, one problem per pthread job.  The sum over
 * the gripper points is taken through the 4x4 moment matrix of the point set (cost of a goal = tr(D Mom D^T) with
 * D = [F_R - M_R | F_t - M_t]); tests/test_oracle_c.py checks it against the literal per-point NumPy oracle.
 */
#define _POSIX_C_SOURCE 200809L
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include "../include/gto_b200.h"

#define NV_MAX (GTO_MAX_OPT + 3)
#define NG_MAX 32

static void bmul34(const double* A, const double* B, double* C) {
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 4; ++c) {
      double s = A[r * 4] * B[c] + A[r * 4 + 1] * B[4 + c] + A[r * 4 + 2] * B[8 + c];
      if (c == 3) s += A[r * 4 + 3];
      C[r * 4 + c] = s;
    }
}

static void joint_motion(int type, const double* axis, double q, double* M) {
  const double ax = axis[0], ay = axis[1], az = axis[2];
  if (type == GTO_JOINT_REVOLUTE) {
    const double s = sin(q), c = cos(q), v = 1.0 - c;
    M[0] = 1.0 - v * (ay * ay + az * az); M[1] = -s * az + v * ax * ay; M[2] = s * ay + v * ax * az; M[3] = 0.0;
    M[4] = s * az + v * ax * ay; M[5] = 1.0 - v * (ax * ax + az * az); M[6] = -s * ax + v * ay * az; M[7] = 0.0;
    M[8] = -s * ay + v * ax * az; M[9] = s * ax + v * ay * az; M[10] = 1.0 - v * (ax * ax + ay * ay); M[11] = 0.0;
  } else {
    M[0] = 1; M[1] = 0; M[2] = 0; M[3] = q * ax;
    M[4] = 0; M[5] = 1; M[6] = 0; M[7] = q * ay;
    M[8] = 0; M[9] = 0; M[10] = 1; M[11] = q * az;
  }
}

typedef struct {
  const gto_robot_desc* R;
  const gto_base_in* in;
  const gto_options* opt;
  gto_base_out* out;
  int nchain, chain[GTO_MAX_MOV];
  double mom[16];
  double* wp; /* [npoints][3] robot points at qc */
  int next;
  pthread_mutex_t mu;
} bjob;

/* cost, half gradient g[a] = <E_a, D Mom>, Gram matrix G[a][b] = <E_a Mom, E_b>; parameters: arm joints, then x, y, theta */
static void goal_linearize(const bjob* J, const double* qx, const double* y, const double* A, double* cost, double* g, double (*G)[NV_MAX]) {
  const gto_robot_desc* R = J->R;
  const int nopt = R->nopt, nv = nopt + 3;
  double E[NV_MAX][12], om[GTO_MAX_OPT][3], mm[GTO_MAX_OPT][3];
  int on[GTO_MAX_OPT];
  for (int k = 0; k < nopt; ++k) on[k] = 0;
  double T[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
  for (int c = 0; c < J->nchain; ++c) {
    const int j = J->chain[c];
    double U[12], M[12];
    bmul34(T, R->mov_origin + 12 * j, U);
    const double* ax = R->mov_axis + 3 * j;
    const double z[3] = {U[0] * ax[0] + U[1] * ax[1] + U[2] * ax[2], U[4] * ax[0] + U[5] * ax[1] + U[6] * ax[2],
                         U[8] * ax[0] + U[9] * ax[1] + U[10] * ax[2]};
    const int k = R->mov_opt[j];
    const double qj = k >= 0 ? qx[k] : J->in->qc[R->mov_qidx[j]];
    if (k >= 0) {
      on[k] = 1;
      if (R->mov_type[j] == GTO_JOINT_REVOLUTE) {
        om[k][0] = z[0]; om[k][1] = z[1]; om[k][2] = z[2];
        mm[k][0] = U[7] * z[2] - U[11] * z[1]; mm[k][1] = U[11] * z[0] - U[3] * z[2]; mm[k][2] = U[3] * z[1] - U[7] * z[0];
      } else {
        om[k][0] = om[k][1] = om[k][2] = 0.0;
        mm[k][0] = z[0]; mm[k][1] = z[1]; mm[k][2] = z[2];
      }
    }
    joint_motion(R->mov_type[j], ax, qj, M);
    bmul34(U, M, T);
  }
  double F[12];
  bmul34(T, R->grip_tf, F);
  const double s = sin(y[2]), c = cos(y[2]);
  double D[12], RA[12];
  for (int col = 0; col < 4; ++col) {
    const double a0 = A[col], a1 = A[4 + col], a2 = A[8 + col];
    D[col] = F[col] - (c * a0 - s * a1 + (col == 3 ? y[0] : 0.0));
    D[4 + col] = F[4 + col] - (s * a0 + c * a1 + (col == 3 ? y[1] : 0.0));
    D[8 + col] = F[8 + col] - a2;
    RA[col] = -s * a0 - c * a1;
    RA[4 + col] = c * a0 - s * a1;
    RA[8 + col] = 0.0;
  }
  for (int k = 0; k < nopt; ++k) {
    if (!on[k]) { memset(E[k], 0, sizeof(E[k])); continue; }
    for (int col = 0; col < 4; ++col) {
      const double f0 = F[col], f1 = F[4 + col], f2 = F[8 + col];
      E[k][col] = om[k][1] * f2 - om[k][2] * f1 + (col == 3 ? mm[k][0] : 0.0);
      E[k][4 + col] = om[k][2] * f0 - om[k][0] * f2 + (col == 3 ? mm[k][1] : 0.0);
      E[k][8 + col] = om[k][0] * f1 - om[k][1] * f0 + (col == 3 ? mm[k][2] : 0.0);
    }
  }
  memset(E[nopt], 0, sizeof(E[0]));
  memset(E[nopt + 1], 0, sizeof(E[0]));
  E[nopt][3] = -1.0;
  E[nopt + 1][7] = -1.0;
  for (int e = 0; e < 12; ++e) E[nopt + 2][e] = -RA[e];
  double DM[12], cs = 0.0;
  for (int r = 0; r < 3; ++r)
    for (int col = 0; col < 4; ++col) {
      double v = 0.0;
      for (int m = 0; m < 4; ++m) v += D[4 * r + m] * J->mom[4 * m + col];
      DM[4 * r + col] = v;
      cs += v * D[4 * r + col];
    }
  *cost = cs;
  for (int a = 0; a < nv; ++a) {
    double v = 0.0, EM[12];
    for (int e = 0; e < 12; ++e) v += E[a][e] * DM[e];
    g[a] = v;
    for (int r = 0; r < 3; ++r)
      for (int col = 0; col < 4; ++col) {
        double u = 0.0;
        for (int m = 0; m < 4; ++m) u += E[a][4 * r + m] * J->mom[4 * m + col];
        EM[4 * r + col] = u;
      }
    for (int b = a; b < nv; ++b) {
      double u = 0.0;
      for (int e = 0; e < 12; ++e) u += EM[e] * E[b][e];
      G[a][b] = u;
      G[b][a] = u;
    }
  }
}

static void chol_solve4(double (*M)[GTO_MAX_OPT], double (*Rh)[4], int n) {
  for (int j = 0; j < n; ++j) {
    double d = M[j][j];
    for (int k = 0; k < j; ++k) d -= M[j][k] * M[j][k];
    d = sqrt(d);
    M[j][j] = d;
    for (int i = j + 1; i < n; ++i) {
      double v = M[i][j];
      for (int k = 0; k < j; ++k) v -= M[i][k] * M[j][k];
      M[i][j] = v / d;
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

typedef struct {
  double G[NG_MAX][NV_MAX][NV_MAX], g[NG_MAX][NV_MAX], c[NG_MAX];
} linset;

static double linearize_all(const bjob* J, int b, double (*qx)[GTO_MAX_OPT], const double* y, linset* L) {
  const int n = J->in->n_goals;
  double F = J->in->w_effort * (y[0] * y[0] + y[1] * y[1] + y[2] * y[2]);
  double sum = 0.0;
  for (int i = 0; i < n; ++i) {
    goal_linearize(J, qx[i], y, J->in->goal_tf + ((size_t)b * n + i) * 12, &L->c[i], L->g[i], L->G[i]);
    sum += L->c[i];
  }
  return F + sum;
}

static void solve_base_one(bjob* J, int b, linset* cur, linset* tri) {
  const gto_robot_desc* R = J->R;
  const gto_options* o = J->opt;
  const int n = J->in->n_goals, nopt = R->nopt;
  const double PI = 3.14159265358979323846, BIG = 1e30;
  const double ylo[3] = {-BIG, -BIG, -PI}, yhi[3] = {BIG, BIG, PI};
  double qx[NG_MAX][GTO_MAX_OPT], qn[NG_MAX][GTO_MAX_OPT], dq[NG_MAX][GTO_MAX_OPT], y[3] = {0, 0, 0}, yn[3];
  for (int i = 0; i < n; ++i)
    for (int k = 0; k < nopt; ++k) qx[i][k] = J->in->qc[R->opt_qidx[k]];
  double F = linearize_all(J, b, qx, y, cur);
  double lam = o->lambda0, nu = 2.0;
  int status = GTO_STATUS_MAX_ITER, it = 0;
  while (it < o->max_iter) {
    double S[3][3] = {{0}}, gy[3] = {0, 0, 0};
    for (int i = 0; i < n; ++i)
      for (int a = 0; a < 3; ++a) {
        gy[a] += cur->g[i][nopt + a];
        for (int c = 0; c < 3; ++c) S[a][c] += cur->G[i][nopt + a][nopt + c];
      }
    for (int a = 0; a < 3; ++a) { gy[a] += J->in->w_effort * y[a]; S[a][a] += J->in->w_effort; }
    int fy[3], fq[NG_MAX][GTO_MAX_OPT];
    double pgmax = 0.0;
    for (int a = 0; a < 3; ++a) {
      fy[a] = (y[a] <= ylo[a] + o->bound_eps && gy[a] > 0.0) || (y[a] >= yhi[a] - o->bound_eps && gy[a] < 0.0);
      if (!fy[a]) pgmax = fmax(pgmax, fabs(gy[a]));
    }
    for (int i = 0; i < n; ++i)
      for (int k = 0; k < nopt; ++k) {
        const double gk = cur->g[i][k];
        fq[i][k] = (qx[i][k] <= R->lo[k] + o->bound_eps && gk > 0.0) || (qx[i][k] >= R->hi[k] - o->bound_eps && gk < 0.0) ||
                   !((R->grip_optmask >> k) & 1u);
        if (!fq[i][k]) pgmax = fmax(pgmax, fabs(gk));
      }
    if (2.0 * pgmax <= o->tol_grad) { status = GTO_STATUS_CONVERGED; break; }
    double M[3][4], Z[NG_MAX][GTO_MAX_OPT][4], Cm[NG_MAX][GTO_MAX_OPT][3];
    for (int a = 0; a < 3; ++a) {
      M[a][3] = fy[a] ? 0.0 : -gy[a];
      for (int c = 0; c < 3; ++c) M[a][c] = (fy[a] || fy[c]) ? (a == c ? 1.0 : 0.0) : S[a][c] + (a == c ? lam * S[a][a] : 0.0);
    }
    for (int i = 0; i < n; ++i) {
      double Hd[GTO_MAX_OPT][GTO_MAX_OPT];
      for (int k = 0; k < nopt; ++k) {
        for (int l = 0; l < nopt; ++l) Hd[k][l] = (fq[i][k] || fq[i][l]) ? 0.0 : cur->G[i][k][l];
        Hd[k][k] = fq[i][k] ? 1.0 : cur->G[i][k][k] + lam * cur->G[i][k][k];
        for (int a = 0; a < 3; ++a) {
          Cm[i][k][a] = (fq[i][k] || fy[a]) ? 0.0 : cur->G[i][k][nopt + a];
          Z[i][k][a] = Cm[i][k][a];
        }
        Z[i][k][3] = fq[i][k] ? 0.0 : -cur->g[i][k];
      }
      chol_solve4(Hd, Z[i], nopt);
      for (int a = 0; a < 3; ++a)
        for (int c = 0; c < 4; ++c) {
          double u = 0.0;
          for (int k = 0; k < nopt; ++k) u += Cm[i][k][a] * Z[i][k][c];
          M[a][c] -= u;
        }
    }
    double dy[3];
    for (int p = 0; p < 3; ++p)
      for (int r = p + 1; r < 3; ++r) {
        const double f = M[r][p] / M[p][p];
        for (int c = p; c < 4; ++c) M[r][c] -= f * M[p][c];
      }
    for (int p = 2; p >= 0; --p) {
      double v = M[p][3];
      for (int c = p + 1; c < 3; ++c) v -= M[p][c] * dy[c];
      dy[p] = v / M[p][p];
    }
    double stepm = 0.0;
    for (int i = 0; i < n; ++i)
      for (int k = 0; k < nopt; ++k) {
        const double d = Z[i][k][3] - (Z[i][k][0] * dy[0] + Z[i][k][1] * dy[1] + Z[i][k][2] * dy[2]);
        qn[i][k] = fmin(fmax(qx[i][k] + d, R->lo[k]), R->hi[k]);
        dq[i][k] = qn[i][k] - qx[i][k];
        stepm = fmax(stepm, fabs(dq[i][k]));
      }
    for (int a = 0; a < 3; ++a) {
      yn[a] = fmin(fmax(y[a] + dy[a], ylo[a]), yhi[a]);
      dy[a] = yn[a] - y[a];
      stepm = fmax(stepm, fabs(dy[a]));
    }
    double pq = 0.0, cd[3] = {0, 0, 0};
    for (int i = 0; i < n; ++i)
      for (int k = 0; k < nopt; ++k) {
        double ad = cur->G[i][k][nopt] * dy[0] + cur->G[i][k][nopt + 1] * dy[1] + cur->G[i][k][nopt + 2] * dy[2];
        for (int l = 0; l < nopt; ++l) ad += cur->G[i][k][l] * dq[i][l];
        pq += cur->g[i][k] * dq[i][k] + 0.5 * dq[i][k] * ad;
        for (int a = 0; a < 3; ++a) cd[a] += cur->G[i][k][nopt + a] * dq[i][k];
      }
    double py = 0.0;
    for (int a = 0; a < 3; ++a) {
      const double ady = S[a][0] * dy[0] + S[a][1] * dy[1] + S[a][2] * dy[2] + cd[a];
      py += gy[a] * dy[a] + 0.5 * dy[a] * ady;
    }
    const double pred = -(pq + py);
    ++it;
    const double Ft = linearize_all(J, b, qn, yn, tri);
    if (!isfinite(Ft)) { status = GTO_STATUS_NAN; break; }
    const double ared = 0.5 * (F - Ft);
    if (pred > 0.0 && ared >= o->eta * pred) {
      const double rho = ared / pred, t = 2.0 * fmin(rho, 1.0) - 1.0;
      memcpy(qx, qn, sizeof(qx));
      memcpy(y, yn, sizeof(y));
      linset* tmp = cur; cur = tri; tri = tmp;
      F = Ft;
      lam = fmax(o->lambda_min, lam * fmax(1.0 / 3.0, 1.0 - t * t * t));
      nu = 2.0;
      if (stepm <= o->tol_step) { status = GTO_STATUS_CONVERGED; break; }
    } else {
      if (pred <= 0.0 && stepm <= o->tol_step) { status = GTO_STATUS_CONVERGED; break; }
      lam = fmin(o->lambda_max, lam * nu);
      nu *= 2.0;
      if (lam >= o->lambda_max) { status = GTO_STATUS_STALLED; break; }
    }
  }
  gto_base_out* out = J->out;
  for (int i = 0; i < n; ++i) {
    double* q = out->Q + ((size_t)b * n + i) * R->ndof;
    for (int j = 0; j < R->ndof; ++j) q[j] = J->in->qc[j];
    for (int k = 0; k < nopt; ++k) q[R->opt_qidx[k]] = qx[i][k];
  }
  for (int a = 0; a < 3; ++a) out->y[3 * b + a] = y[a];
  if (out->cost) out->cost[b] = F;
  if (out->iters) out->iters[b] = it;
  if (out->status) out->status[b] = status;
  if (out->collision) {
    double coll = 0.0;
    if (J->in->occupancy) {
      const double s = sin(y[2]), c = cos(y[2]);
      const int nx = J->in->occ_dims[0], ny = J->in->occ_dims[1];
      for (int i = 0; i < R->npoints; ++i) {
        const double px = J->wp[3 * i] - y[0], py2 = J->wp[3 * i + 1] - y[1];
        const double ux = c * px + s * py2, uy = -s * px + c * py2;
        const double fx = floor((ux - J->in->occ_origin[0]) / J->in->occ_resolution), fyv = floor((uy - J->in->occ_origin[1]) / J->in->occ_resolution);
        const int ix = (int)fmin(fmax(fx, 0.0), (double)(nx - 1)), iy = (int)fmin(fmax(fyv, 0.0), (double)(ny - 1));
        coll += (double)J->in->occupancy[(size_t)ix * ny + iy];
      }
    }
    out->collision[b] = coll;
  }
}

static void* base_worker(void* arg) {
  bjob* J = (bjob*)arg;
  linset* a = (linset*)malloc(sizeof(linset));
  linset* t = (linset*)malloc(sizeof(linset));
  for (;;) {
    pthread_mutex_lock(&J->mu);
    const int b = J->next++;
    pthread_mutex_unlock(&J->mu);
    if (b >= J->in->B) break;
    solve_base_one(J, b, a, t);
  }
  free(a);
  free(t);
  return NULL;
}

int oracle_base_place(const gto_robot_desc* R, const gto_base_in* in, const gto_options* opt, gto_base_out* out, int nthreads) {
  if (!R || !in || !opt || !out || in->B < 1 || in->n_goals < 1 || in->n_goals > NG_MAX || R->nopt > GTO_MAX_OPT) return -1;
  bjob J;
  memset(&J, 0, sizeof(J));
  J.R = R; J.in = in; J.opt = opt; J.out = out;
  int tmp[GTO_MAX_MOV], c = 0;
  for (int j = R->grip_mov; j >= 0; j = R->mov_parent[j]) tmp[c++] = j;
  J.nchain = c;
  for (int i = 0; i < c; ++i) J.chain[i] = tmp[c - 1 - i];
  for (int i = R->grip_pt_start; i < R->grip_pt_start + R->grip_pt_count; ++i) {
    const double v[4] = {R->points[3 * i], R->points[3 * i + 1], R->points[3 * i + 2], 1.0};
    for (int a = 0; a < 4; ++a)
      for (int b = 0; b < 4; ++b) J.mom[4 * a + b] += v[a] * v[b];
  }
  /* robot surface points at qc (occupancy count) */
  J.wp = (double*)malloc(sizeof(double) * 3 * (size_t)R->npoints);
  {
    double* Tm = (double*)malloc(sizeof(double) * 12 * (size_t)R->nmov);
    for (int j = 0; j < R->nmov; ++j) {
      double M[12], A[12];
      joint_motion(R->mov_type[j], R->mov_axis + 3 * j, in->qc[R->mov_qidx[j]], M);
      bmul34(R->mov_origin + 12 * j, M, A);
      if (R->mov_parent[j] < 0) memcpy(Tm + 12 * j, A, sizeof(A));
      else bmul34(Tm + 12 * R->mov_parent[j], A, Tm + 12 * j);
    }
    for (int l = 0; l < R->nlinks; ++l) {
      double Fm[12];
      if (R->link_mov[l] < 0) memcpy(Fm, R->link_tf + 12 * l, sizeof(Fm));
      else bmul34(Tm + 12 * R->link_mov[l], R->link_tf + 12 * l, Fm);
      for (int i = R->link_pt_start[l]; i < R->link_pt_start[l] + R->link_pt_count[l]; ++i) {
        const double x = R->points[3 * i], y = R->points[3 * i + 1], z = R->points[3 * i + 2];
        for (int r = 0; r < 3; ++r) J.wp[3 * i + r] = Fm[4 * r] * x + Fm[4 * r + 1] * y + Fm[4 * r + 2] * z + Fm[4 * r + 3];
      }
    }
    free(Tm);
  }
  if (nthreads <= 0) nthreads = (int)sysconf(_SC_NPROCESSORS_ONLN);
  if (nthreads > in->B) nthreads = in->B;
  if (nthreads < 1) nthreads = 1;
  pthread_mutex_init(&J.mu, NULL);
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)nthreads);
  for (int i = 0; i < nthreads; ++i) pthread_create(&th[i], NULL, base_worker, &J);
  for (int i = 0; i < nthreads; ++i) pthread_join(th[i], NULL);
  pthread_mutex_destroy(&J.mu);
  free(th);
  free(J.wp);
  return nthreads;
}
