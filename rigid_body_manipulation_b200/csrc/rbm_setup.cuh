// Per-item arithmetic of the small frame-algebra helpers (__host__ __device__: the kernels of rbm_setup.cu call these one item per
// thread, and tests/host_harness instantiates the SAME functions for the host so that the CPU-only suite -- including the run of the
// reference's own simulate() on top of the drop-in packages -- exercises this code without a GPU).
//
//   transfer_simat              dynamics/dynamics.py:72-106      Ad(T^-1)^T G Ad(T^-1)
//   coordinate_transfer_simat   dynamics/dynamics.py:260-263     Ad(T) G Ad(T)^T
//   coordinate_transfer_imat    dynamics/dynamics.py:252-257     R I R^T + m (|t|^2 1 - t t^T)
//   get_spatial_inertia_matrix  dynamics/dynamics.py:49-69       blkdiag(m 1, diag(I))
//   compose (tq2se3 / tr2se3)   transformations/transformations.py:8-50   (t, quat | R) -> pose, with liegroups' validity checks
//   extract_lin{vel,acc}_frame_transferred   dynamics/dynamics.py:160-212
#pragma once
#include <cmath>

#include "rbm_rnea.cuh"

namespace rbm {

RBM_HD void build_adjoint(const double* R, const double* t, double (&Ad)[6][6]) {
  // [[R, [t]x R], [0, R]]
  const double tx[3][3] = {{0.0, -t[2], t[1]}, {t[2], 0.0, -t[0]}, {-t[1], t[0], 0.0}};
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      Ad[r][c] = R[3 * r + c];
      Ad[3 + r][3 + c] = R[3 * r + c];
      Ad[3 + r][c] = 0.0;
      Ad[r][3 + c] = tx[r][0] * R[c] + tx[r][1] * R[3 + c] + tx[r][2] * R[6 + c];
    }
}

// mode 0: Ad(T^-1)^T G Ad(T^-1)   (transfer_simat)      mode 1: Ad(T) G Ad(T)^T   (coordinate_transfer_simat)
RBM_HD void transfer_simat_item(const double* P /* [R | t] */, const double* G /* 6x6 */, double* O /* 6x6 */, int mode) {
  double R[9], t[3];
  if (mode == 0) {  // inverse pose: (R^T, -R^T t)
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int c = 0; c < 3; ++c) R[3 * r + c] = P[3 * c + r];
#pragma unroll
    for (int r = 0; r < 3; ++r) t[r] = -(R[3 * r] * P[9] + R[3 * r + 1] * P[10] + R[3 * r + 2] * P[11]);
  } else {
#pragma unroll
    for (int k = 0; k < 9; ++k) R[k] = P[k];
    t[0] = P[9]; t[1] = P[10]; t[2] = P[11];
  }
  double Ad[6][6];
  build_adjoint(R, t, Ad);
  double M[6][6];  // mode 0: G Ad ; mode 1: G Ad^T
#pragma unroll
  for (int r = 0; r < 6; ++r)
#pragma unroll
    for (int c = 0; c < 6; ++c) {
      double v = 0.0;
#pragma unroll
      for (int k = 0; k < 6; ++k) v += G[6 * r + k] * (mode == 0 ? Ad[k][c] : Ad[c][k]);
      M[r][c] = v;
    }
#pragma unroll
  for (int r = 0; r < 6; ++r)
#pragma unroll
    for (int c = 0; c < 6; ++c) {
      double v = 0.0;
#pragma unroll
      for (int k = 0; k < 6; ++k) v += (mode == 0 ? Ad[k][r] : Ad[r][k]) * M[k][c];
      O[6 * r + c] = v;
    }
}

RBM_HD void transfer_imat_item(const double* P, const double* I, double m, double* O) {
  double RI[9];
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) RI[3 * r + c] = P[3 * r] * I[c] + P[3 * r + 1] * I[3 + c] + P[3 * r + 2] * I[6 + c];
  const double t[3] = {P[9], P[10], P[11]};
  const double tt = t[0] * t[0] + t[1] * t[1] + t[2] * t[2];
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const double rirt = RI[3 * r] * P[3 * c] + RI[3 * r + 1] * P[3 * c + 1] + RI[3 * r + 2] * P[3 * c + 2];
      O[3 * r + c] = rirt + m * ((r == c ? tt : 0.0) - t[r] * t[c]);
    }
}

RBM_HD void spatial_inertia_item(double mass, const double* diag, double* O) {
#pragma unroll
  for (int k = 0; k < 36; ++k) O[k] = 0.0;
  O[0] = O[7] = O[14] = mass;
  O[21] = diag[0];
  O[28] = diag[1];
  O[35] = diag[2];
}

// rot_len 4: wxyz quaternion (unit norm required, |norm-1| <= 1e-8 + 1e-5 like np.isclose); rot_len 9: rotation matrix
// (det ~ 1 and R^T R ~ 1 required).  Returns 0 ok, 1 = non-unit quaternion, 2 = invalid rotation matrix.
RBM_HD int compose_item(const double* trans, const double* rot, int rot_len, double* O) {
  int st = 0;
  if (rot_len == 4) {
    const double w = rot[0], x = rot[1], y = rot[2], z = rot[3];
    const double nrm = sqrt(w * w + x * x + y * y + z * z);
    if (!(fabs(nrm - 1.0) <= 1e-8 + 1e-5)) st = 1;
    O[0] = 1.0 - 2.0 * (y * y + z * z); O[1] = 2.0 * (x * y - w * z);       O[2] = 2.0 * (w * y + x * z);
    O[3] = 2.0 * (w * z + x * y);       O[4] = 1.0 - 2.0 * (x * x + z * z); O[5] = 2.0 * (y * z - w * x);
    O[6] = 2.0 * (x * z - w * y);       O[7] = 2.0 * (w * x + y * z);       O[8] = 1.0 - 2.0 * (x * x + y * y);
  } else {
    const double* R = rot;
#pragma unroll
    for (int k = 0; k < 9; ++k) O[k] = R[k];
    const double det = R[0] * (R[4] * R[8] - R[5] * R[7]) - R[1] * (R[3] * R[8] - R[5] * R[6]) + R[2] * (R[3] * R[7] - R[4] * R[6]);
    if (!(fabs(det - 1.0) <= 1e-8 + 1e-5)) st = 2;
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const double v = R[r] * R[c] + R[3 + r] * R[3 + c] + R[6 + r] * R[6 + c];
        const double want = r == c ? 1.0 : 0.0;
        if (!(fabs(v - want) <= 1e-8 + 1e-5 * want)) st = 2;
      }
  }
  O[9] = trans[0]; O[10] = trans[1]; O[11] = trans[2];
  return st;
}

// v_p = [V]^ p~ ;  a_p = [dV]^ p~ + [V]^ [V]^ p~   (homogeneous 4-vectors; 4th component 0)
RBM_HD void point_motion_item(const double* tw, const double* dtw, const double* pt, double* linvel, double* linacc) {
  const G3<double> v = g3(tw), w = g3(tw + 3), p = g3(pt);
  const G3<double> lv = gcross(w, p) + v;
  if (linvel) { linvel[0] = lv.x; linvel[1] = lv.y; linvel[2] = lv.z; }
  if (linacc) {
    const G3<double> a = g3(dtw), l = g3(dtw + 3);
    // [V]^ applied to the (homogeneous, 4th = 0) velocity vector: w x lv
    const G3<double> la = gcross(l, p) + a + gcross(w, lv);
    linacc[0] = la.x; linacc[1] = la.y; linacc[2] = la.z;
  }
}

}  // namespace rbm
