// Batched device versions of the reference's small frame-algebra helpers (one item per thread; the arithmetic is in rbm_setup.cuh).
// They are setup-time functions in the reference (a handful of calls per run) but become batch operations
// once many objects / poses are processed at once (per-object inertias for all targets, pose registers).
#include "rbm_internal.h"
#include "rbm_setup.cuh"

namespace rbm {

constexpr int kSetupBlock = 128;

__global__ void __launch_bounds__(kSetupBlock) k_transfer_simat(const double* __restrict__ poses, const double* __restrict__ simats, double* __restrict__ out,
                                                                int64_t n, int pose_stride, int simat_stride, int mode) {
  const int64_t s = (int64_t)blockIdx.x * kSetupBlock + threadIdx.x;
  if (s >= n) return;
  transfer_simat_item(poses + s * pose_stride, simats + s * simat_stride, out + s * 36, mode);
}

__global__ void __launch_bounds__(kSetupBlock) k_transfer_imat(const double* __restrict__ poses, const double* __restrict__ imats, const double* __restrict__ mass,
                                                               double* __restrict__ out, int64_t n) {
  const int64_t s = (int64_t)blockIdx.x * kSetupBlock + threadIdx.x;
  if (s >= n) return;
  transfer_imat_item(poses + s * 12, imats + s * 9, mass[s], out + s * 9);
}

__global__ void __launch_bounds__(kSetupBlock) k_spatial_inertia(const double* __restrict__ mass, const double* __restrict__ diag, double* __restrict__ out, int64_t n) {
  const int64_t s = (int64_t)blockIdx.x * kSetupBlock + threadIdx.x;
  if (s >= n) return;
  spatial_inertia_item(mass[s], diag + 3 * s, out + s * 36);
}

__global__ void __launch_bounds__(kSetupBlock) k_compose(const double* __restrict__ trans, const double* __restrict__ rot, int rot_len,
                                                         double* __restrict__ out, int* __restrict__ status, int64_t n) {
  const int64_t s = (int64_t)blockIdx.x * kSetupBlock + threadIdx.x;
  if (s >= n) return;
  status[s] = compose_item(trans + 3 * s, rot + (int64_t)rot_len * s, rot_len, out + s * 12);
}

__global__ void __launch_bounds__(kSetupBlock) k_point_motion(const double* __restrict__ tw, const double* __restrict__ dtw, const double* __restrict__ pts,
                                                              double* __restrict__ linvel, double* __restrict__ linacc, int64_t n) {
  const int64_t s = (int64_t)blockIdx.x * kSetupBlock + threadIdx.x;
  if (s >= n) return;
  point_motion_item(tw + 6 * s, dtw ? dtw + 6 * s : nullptr, pts + 3 * s, linvel ? linvel + 3 * s : nullptr, linacc ? linacc + 3 * s : nullptr);
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
