// Batched device versions of the reference's small frame-algebra helpers (one item per thread).
// They are setup-time functions in the reference (a handful of calls per run) but become batch operations
// once many objects / poses are processed at once (per-object inertias for all targets, pose registers).
//
//   transfer_simat              dynamics/dynamics.py:72-106      Ad(T^-1)^T G Ad(T^-1)
//   coordinate_transfer_simat   dynamics/dynamics.py:260-263     Ad(T) G Ad(T)^T
//   coordinate_transfer_imat    dynamics/dynamics.py:252-257     R I R^T + m (|t|^2 1 - t t^T)
//   get_spatial_inertia_matrix  dynamics/dynamics.py:49-69       blkdiag(m 1, diag(I))
//   compose (tq2se3 / tr2se3)   transformations/transformations.py:8-50   (t, quat | R) -> pose, with liegroups' validity checks
//   extract_lin{vel,acc}_frame_transferred   dynamics/dynamics.py:160-212
#include "rbm_internal.h"
#include "rbm_rnea.cuh"

namespace rbm {

constexpr int kSetupBlock = 128;

__device__ __forceinline__ void build_adjoint(const double* R, const double* t, double (&Ad)[6][6]) {
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
__global__ void __launch_bounds__(kSetupBlock) k_transfer_simat(const double* __restrict__ poses, const double* __restrict__ simats, double* __restrict__ out,
                                                                int64_t n, int pose_stride, int simat_stride, int mode) {
  const int64_t s = (int64_t)blockIdx.x * kSetupBlock + threadIdx.x;
  if (s >= n) return;
  const double* P = poses + s * pose_stride;
  const double* G = simats + s * simat_stride;
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
  double* O = out + s * 36;
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

__global__ void __launch_bounds__(kSetupBlock) k_transfer_imat(const double* __restrict__ poses, const double* __restrict__ imats, const double* __restrict__ mass,
                                                               double* __restrict__ out, int64_t n) {
  const int64_t s = (int64_t)blockIdx.x * kSetupBlock + threadIdx.x;
  if (s >= n) return;
  const double* P = poses + s * 12;
  const double* I = imats + s * 9;
  const double m = mass[s];
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
      out[s * 9 + 3 * r + c] = rirt + m * ((r == c ? tt : 0.0) - t[r] * t[c]);
    }
}

__global__ void __launch_bounds__(kSetupBlock) k_spatial_inertia(const double* __restrict__ mass, const double* __restrict__ diag, double* __restrict__ out, int64_t n) {
  const int64_t s = (int64_t)blockIdx.x * kSetupBlock + threadIdx.x;
  if (s >= n) return;
  double* O = out + s * 36;
#pragma unroll
  for (int k = 0; k < 36; ++k) O[k] = 0.0;
  O[0] = O[7] = O[14] = mass[s];
  O[21] = diag[3 * s];
  O[28] = diag[3 * s + 1];
  O[35] = diag[3 * s + 2];
}

// rot_len 4: wxyz quaternion (unit norm required, |norm-1| <= 1e-8 + 1e-5 like np.isclose); rot_len 9: rotation matrix
// (det ~ 1 and R^T R ~ 1 required).  status[s] = 0 ok, 1 = non-unit quaternion, 2 = invalid rotation matrix.
__global__ void __launch_bounds__(kSetupBlock) k_compose(const double* __restrict__ trans, const double* __restrict__ rot, int rot_len,
                                                         double* __restrict__ out, int* __restrict__ status, int64_t n) {
  const int64_t s = (int64_t)blockIdx.x * kSetupBlock + threadIdx.x;
  if (s >= n) return;
  double* O = out + s * 12;
  int st = 0;
  if (rot_len == 4) {
    const double w = rot[4 * s], x = rot[4 * s + 1], y = rot[4 * s + 2], z = rot[4 * s + 3];
    const double nrm = sqrt(w * w + x * x + y * y + z * z);
    if (!(fabs(nrm - 1.0) <= 1e-8 + 1e-5)) st = 1;
    O[0] = 1.0 - 2.0 * (y * y + z * z); O[1] = 2.0 * (x * y - w * z);       O[2] = 2.0 * (w * y + x * z);
    O[3] = 2.0 * (w * z + x * y);       O[4] = 1.0 - 2.0 * (x * x + z * z); O[5] = 2.0 * (y * z - w * x);
    O[6] = 2.0 * (x * z - w * y);       O[7] = 2.0 * (w * x + y * z);       O[8] = 1.0 - 2.0 * (x * x + y * y);
  } else {
    const double* R = rot + 9 * s;
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
  O[9] = trans[3 * s]; O[10] = trans[3 * s + 1]; O[11] = trans[3 * s + 2];
  status[s] = st;
}

// v_p = [V]^ p~ ;  a_p = [dV]^ p~ + [V]^ [V]^ p~   (homogeneous 4-vectors; 4th component 0)
__global__ void __launch_bounds__(kSetupBlock) k_point_motion(const double* __restrict__ tw, const double* __restrict__ dtw, const double* __restrict__ pts,
                                                              double* __restrict__ linvel, double* __restrict__ linacc, int64_t n) {
  const int64_t s = (int64_t)blockIdx.x * kSetupBlock + threadIdx.x;
  if (s >= n) return;
  const G3<double> v = g3(tw + 6 * s), w = g3(tw + 6 * s + 3), p = g3(pts + 3 * s);
  const G3<double> lv = gcross(w, p) + v;
  if (linvel) { linvel[3 * s] = lv.x; linvel[3 * s + 1] = lv.y; linvel[3 * s + 2] = lv.z; }
  if (linacc) {
    const G3<double> a = g3(dtw + 6 * s), l = g3(dtw + 6 * s + 3);
    // [V]^ applied to the (homogeneous, 4th = 0) velocity vector: w x lv
    const G3<double> la = gcross(l, p) + a + gcross(w, lv);
    linacc[3 * s] = la.x; linacc[3 * s + 1] = la.y; linacc[3 * s + 2] = la.z;
  }
}

static inline unsigned sgrid(int64_t n) { return (unsigned)((n + kSetupBlock - 1) / kSetupBlock); }

int launch_transfer_simat(const double* poses, const double* simats, double* out, int64_t n, int pose_stride, int simat_stride, int mode, cudaStream_t st) {
  if (n == 0) return RBM_OK;
  k_transfer_simat<<<sgrid(n), kSetupBlock, 0, st>>>(poses, simats, out, n, pose_stride, simat_stride, mode);
  RBM_CUDA_TRY(cudaGetLastError());
  return RBM_OK;
}
int launch_transfer_imat(const double* poses, const double* imats, const double* mass, double* out, int64_t n, cudaStream_t st) {
  if (n == 0) return RBM_OK;
  k_transfer_imat<<<sgrid(n), kSetupBlock, 0, st>>>(poses, imats, mass, out, n);
  RBM_CUDA_TRY(cudaGetLastError());
  return RBM_OK;
}
int launch_spatial_inertia(const double* mass, const double* diag, double* out, int64_t n, cudaStream_t st) {
  if (n == 0) return RBM_OK;
  k_spatial_inertia<<<sgrid(n), kSetupBlock, 0, st>>>(mass, diag, out, n);
  RBM_CUDA_TRY(cudaGetLastError());
  return RBM_OK;
}
int launch_compose(const double* trans, const double* rot, int rot_len, double* out, int* status, int64_t n, cudaStream_t st) {
  if (n == 0) return RBM_OK;
  k_compose<<<sgrid(n), kSetupBlock, 0, st>>>(trans, rot, rot_len, out, status, n);
  RBM_CUDA_TRY(cudaGetLastError());
  return RBM_OK;
}
int launch_point_motion(const double* tw, const double* dtw, const double* pts, double* linvel, double* linacc, int64_t n, cudaStream_t st) {
  if (n == 0) return RBM_OK;
  k_point_motion<<<sgrid(n), kSetupBlock, 0, st>>>(tw, dtw, pts, linvel, linacc, n);
  RBM_CUDA_TRY(cudaGetLastError());
  return RBM_OK;
}

}  // namespace rbm
