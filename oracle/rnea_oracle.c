/* ORACLE / TEST INFRASTRUCTURE ONLY -- never linked into or called by the product library.
 *
 * Plain-C restatement (fp64) of the reference's recursive Newton-Euler inverse dynamics,
 * reference dynamics/dynamics.py:109-157, with the liegroups operations it calls written out as the
 * explicit 6x6 matrices the reference builds (SE3.exp :126, SE3.adjoint :128,:130,:143, SE3.curlywedge :130,:145).
 * It deliberately keeps the reference's dense matrix formulation (adjoint matrices, full mat-vecs) so that it is
 * algorithmically independent of the CUDA kernels' closed-form cross-product formulation.
 *
 * Built by oracle/build_c.py into oracle/_build/librnea_oracle.so; pinned against the golden vectors produced by the
 * reference's own files in tests/test_oracle_c.py.  Used as a fast CPU checker for large batches and as an optional
 * "what would a compiled CPU port do" data point in bench.py (clearly labelled; the reference itself is numpy).
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#define MAXJ 16

static void wedge3(const double *v, double W[3][3]) {
  W[0][0] = 0;     W[0][1] = -v[2]; W[0][2] = v[1];
  W[1][0] = v[2];  W[1][1] = 0;     W[1][2] = -v[0];
  W[2][0] = -v[1]; W[2][1] = v[0];  W[2][2] = 0;
}

/* liegroups SO3.exp and SO3.left_jacobian, including the np.isclose(angle, 0.) first-order branch (atol 1e-8) */
static void so3_exp_jac(const double *phi, double R[3][3], double J[3][3]) {
  double angle = sqrt(phi[0] * phi[0] + phi[1] * phi[1] + phi[2] * phi[2]);
  double W[3][3];
  int i, j;
  if (fabs(angle) <= 1e-8) {
    wedge3(phi, W);
    for (i = 0; i < 3; ++i)
      for (j = 0; j < 3; ++j) {
        R[i][j] = (i == j) + W[i][j];
        J[i][j] = (i == j) + 0.5 * W[i][j];
      }
    return;
  }
  double a[3] = {phi[0] / angle, phi[1] / angle, phi[2] / angle};
  double s = sin(angle), c = cos(angle);
  wedge3(a, W);
  for (i = 0; i < 3; ++i)
    for (j = 0; j < 3; ++j) {
      R[i][j] = c * (i == j) + (1 - c) * a[i] * a[j] + s * W[i][j];
      J[i][j] = (s / angle) * (i == j) + (1 - s / angle) * a[i] * a[j] + ((1 - c) / angle) * W[i][j];
    }
}

/* SE3.adjoint: [[R, [t]x R], [0, R]] */
static void adjoint6(const double R[3][3], const double *t, double Ad[6][6]) {
  double W[3][3];
  int i, j, k;
  wedge3(t, W);
  memset(Ad, 0, 36 * sizeof(double));
  for (i = 0; i < 3; ++i)
    for (j = 0; j < 3; ++j) {
      double s = 0;
      for (k = 0; k < 3; ++k) s += W[i][k] * R[k][j];
      Ad[i][j] = R[i][j];
      Ad[3 + i][3 + j] = R[i][j];
      Ad[i][3 + j] = s;
    }
}

/* SE3.curlywedge: [[ [w]x, [v]x ], [0, [w]x ]] for xi = [v; w] */
static void curlywedge6(const double *xi, double ad[6][6]) {
  double Wv[3][3], Ww[3][3];
  int i, j;
  wedge3(xi, Wv);
  wedge3(xi + 3, Ww);
  memset(ad, 0, 36 * sizeof(double));
  for (i = 0; i < 3; ++i)
    for (j = 0; j < 3; ++j) {
      ad[i][j] = Ww[i][j];
      ad[3 + i][3 + j] = Ww[i][j];
      ad[i][3 + j] = Wv[i][j];
    }
}

static void matvec6(const double M[6][6], const double *x, double *y) {
  int i, k;
  for (i = 0; i < 6; ++i) {
    double s = 0;
    for (k = 0; k < 6; ++k) s += M[i][k] * x[k];
    y[i] = s;
  }
}
static void matTvec6(const double M[6][6], const double *x, double *y) {
  int i, k;
  for (i = 0; i < 6; ++i) {
    double s = 0;
    for (k = 0; k < 6; ++k) s += M[k][i] * x[k];
    y[i] = s;
  }
}

/* One sample.  Layouts as in include/rbm_b200.h: hposes_Rt [(nj+1)][12], simats [(nj+1)][36], uscrews [nj][6],
 * traj [3][nj] (rows q, qd, qdd).  Optional outputs (may be NULL): twists / dtwists [(nj+1)][6], poses [nj][12]. */
int rnea_oracle_one(int nj, const double *hposes_Rt, const double *simats, const double *uscrews, const double *twist_0,
                    const double *dtwist_0, const double *wrench_tip, const double *pose_tip_Rt, const double *traj, double *tau,
                    double *twists, double *dtwists, double *poses) {
  double Ads[MAXJ + 1][6][6], V[MAXJ + 1][6], dV[MAXJ + 1][6];
  int i, j, k;
  if (nj < 1 || nj > MAXJ) return -1;
  memcpy(V[0], twist_0, 6 * sizeof(double));
  memcpy(dV[0], dtwist_0, 6 * sizeof(double));
  for (i = 0; i < nj; ++i) { /* forward sweep, dynamics.py:125-133 */
    const double *S = uscrews + 6 * i, *H = hposes_Rt + 12 * (i + 1);
    double q = traj[i], qd = traj[nj + i], qdd = traj[2 * nj + i];
    double xi[6], Re[3][3], Je[3][3], te[3], R[3][3], t[3], ad[6][6], tmp[6], adS[6];
    for (k = 0; k < 6; ++k) xi[k] = -1 * S[k] * q;
    so3_exp_jac(xi + 3, Re, Je); /* SE3.exp = (SO3.exp(phi), J_l(phi) rho) */
    for (j = 0; j < 3; ++j) te[j] = Je[j][0] * xi[0] + Je[j][1] * xi[1] + Je[j][2] * xi[2];
    for (j = 0; j < 3; ++j) { /* .dot(h_p): (Re Rh, Re th + te) */
      for (k = 0; k < 3; ++k) R[j][k] = Re[j][0] * H[k] + Re[j][1] * H[3 + k] + Re[j][2] * H[6 + k];
      t[j] = Re[j][0] * H[9] + Re[j][1] * H[10] + Re[j][2] * H[11] + te[j];
    }
    adjoint6(R, t, Ads[i]);
    matvec6(Ads[i], V[i], tmp); /* Eq. 8.51 */
    for (k = 0; k < 6; ++k) V[i + 1][k] = tmp[k] + S[k] * qd;
    matvec6(Ads[i], dV[i], tmp); /* Eq. 8.52 */
    curlywedge6(V[i + 1], ad);
    matvec6(ad, S, adS);
    for (k = 0; k < 6; ++k) dV[i + 1][k] = tmp[k] + adS[k] * qd + S[k] * qdd;
    if (poses) {
      for (j = 0; j < 3; ++j)
        for (k = 0; k < 3; ++k) poses[12 * i + 3 * j + k] = R[j][k];
      for (j = 0; j < 3; ++j) poses[12 * i + 9 + j] = t[j];
    }
  }
  { /* tip pose, dynamics.py:137 */
    double R[3][3], t[3] = {0, 0, 0};
    for (j = 0; j < 3; ++j)
      for (k = 0; k < 3; ++k) R[j][k] = pose_tip_Rt ? pose_tip_Rt[3 * j + k] : (j == k);
    if (pose_tip_Rt) { t[0] = pose_tip_Rt[9]; t[1] = pose_tip_Rt[10]; t[2] = pose_tip_Rt[11]; }
    adjoint6(R, t, Ads[nj]);
  }
  {
    double F[6], Fn[6], GdV[6], GV[6], adT[6], ad[6][6];
    for (k = 0; k < 6; ++k) F[k] = wrench_tip ? wrench_tip[k] : 0.0;
    for (i = nj; i >= 1; --i) { /* backward sweep, dynamics.py:140-147 (Eq. 8.53) */
      const double(*G)[6] = (const double(*)[6])(simats + 36 * i);
      matTvec6(Ads[i], F, Fn);
      matvec6(G, dV[i], GdV);
      matvec6(G, V[i], GV);
      curlywedge6(V[i], ad);
      matTvec6(ad, GV, adT);
      for (k = 0; k < 6; ++k) F[k] = Fn[k] + GdV[k] - adT[k];
      {
        double s = 0; /* Eq. 8.54: Hadamard product with the screw, row sum */
        for (k = 0; k < 6; ++k) s += F[k] * uscrews[6 * (i - 1) + k];
        tau[i - 1] = s;
      }
    }
  }
  if (twists) memcpy(twists, V, (size_t)(nj + 1) * 6 * sizeof(double));
  if (dtwists) memcpy(dtwists, dV, (size_t)(nj + 1) * 6 * sizeof(double));
  return 0;
}

/* Batch: traj [n][3][nj] -> tau [n][nj] (+ optional last-link twists [n][6]).  OpenMP over samples when compiled with -fopenmp. */
int rnea_oracle_batch(int nj, const double *hposes_Rt, const double *simats, const double *uscrews, const double *twist_0,
                      const double *dtwist_0, const double *wrench_tip, const double *pose_tip_Rt, const double *traj, double *tau,
                      double *twist_last, double *dtwist_last, int64_t n) {
  int64_t s;
  int bad = 0;
#pragma omp parallel for schedule(static) reduction(| : bad)
  for (s = 0; s < n; ++s) {
    double tw[(MAXJ + 1) * 6], dtw[(MAXJ + 1) * 6];
    int rc = rnea_oracle_one(nj, hposes_Rt, simats, uscrews, twist_0, dtwist_0, wrench_tip, pose_tip_Rt, traj + s * 3 * nj, tau + s * nj, tw, dtw, 0);
    if (rc) bad |= 1;
    if (twist_last) memcpy(twist_last + 6 * s, tw + 6 * nj, 6 * sizeof(double));
    if (dtwist_last) memcpy(dtwist_last + 6 * s, dtw + 6 * nj, 6 * sizeof(double));
  }
  return bad ? -1 : 0;
}
