// Host-side internals shared by the translation units of librbm_b200.so (not part of the ABI).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/rbm_b200.h"
#include "rbm_model.cuh"

struct rbm_model {
  int nj = 0;
  int path = rbm::PATH_GENERIC;
  int device = 0;
  std::vector<double> gp64;  // generic packed parameters (rbm_model.cuh layout), host copies
  std::vector<float> gp32;
  double* d_gp64 = nullptr;  // device copies, staged into shared memory by every block
  float* d_gp32 = nullptr;
  rbm::FastParams<double> fp64;
  rbm::FastParams<float> fp32;
  // staging for the *_host entry points (allocated on first use, grow-only, guarded by pipe_mu)
  static constexpr int kPipeSlots = 3;
  mutable std::mutex pipe_mu;
  mutable cudaStream_t pipe_st[kPipeSlots] = {nullptr, nullptr, nullptr};
  mutable void* pipe_in[kPipeSlots] = {nullptr, nullptr, nullptr};
  mutable void* pipe_out[kPipeSlots] = {nullptr, nullptr, nullptr};
  mutable size_t pipe_in_bytes = 0, pipe_out_bytes = 0;
};

namespace rbm {

void set_error(const std::string& msg);
int cuda_fail(cudaError_t e, const char* what);

#define RBM_CUDA_TRY(expr)                                   \
  do {                                                       \
    cudaError_t _e = (expr);                                 \
    if (_e != cudaSuccess) return ::rbm::cuda_fail(_e, #expr); \
  } while (0)

template <class T> struct ModelView;
template <> struct ModelView<double> {
  static const double* generic(const rbm_model* m) { return m->d_gp64; }
  static const FastParams<double>& fast(const rbm_model* m) { return m->fp64; }
};
template <> struct ModelView<float> {
  static const float* generic(const rbm_model* m) { return m->d_gp32; }
  static const FastParams<float>& fast(const rbm_model* m) { return m->fp32; }
};

// launchers (rbm_rnea.cu)
template <class T>
int launch_rnea_soa(const rbm_model* m, const T* q, const T* qd, const T* qdd, T* tau, T* V, T* dV, int64_t n, int64_t ld, cudaStream_t st);
template <class T>
int launch_rnea_aos(const rbm_model* m, const T* traj, T* tau, int64_t n, cudaStream_t st);
template <class T>
int launch_rnea_full(const rbm_model* m, const T* traj, T* tau, T* poses, T* twists, T* dtwists, int64_t n, cudaStream_t st);

}  // namespace rbm
