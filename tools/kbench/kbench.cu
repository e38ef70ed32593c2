// Stand-alone A/B timing harness for the C ABI (no Python, no torch: starts in a second on a GPU box).
//   kbench <consts.bin> gram   [n] [iters]      -- fp64 Gram: TMA kernel vs RBM_FLAG_NO_TMA (direct loads), packs compared
//   kbench <consts.bin> gram32 [n] [iters]      -- the same in fp32
//   kbench <consts.bin> lin    [n] [iters]      -- LQR linearisation
//   kbench <consts.bin> rnea   [n] [iters]      -- headline inverse dynamics
// consts.bin (tools/kbench/make_consts.py): doubles [hposes_Rt 7x12 | simats 7x36 | uscrews 6x6 | twist_0 6 | dtwist_0 6 | pose_sen 12]
// Inputs are generated on the device (hash -> uniform), timing with CUDA events around `iters` back-to-back launches after 3 warm-ups.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/rbm_b200.h"

#define CK(x)                                                                              \
  do {                                                                                     \
    cudaError_t e_ = (x);                                                                  \
    if (e_ != cudaSuccess) { std::fprintf(stderr, "CUDA %s: %s\n", #x, cudaGetErrorString(e_)); std::exit(2); } \
  } while (0)
#define RB(x)                                                                              \
  do {                                                                                     \
    int rc_ = (x);                                                                         \
    if (rc_ != 0) { std::fprintf(stderr, "rbm %s -> %d: %s\n", #x, rc_, rbm_last_error_string()); std::exit(3); } \
  } while (0)

__global__ void fill_uniform(double* p, int64_t n, double lo, double hi, uint64_t seed) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    uint64_t z = (uint64_t)i + seed * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    p[i] = lo + (hi - lo) * ((double)(z >> 11) * (1.0 / 9007199254740992.0));
  }
}

__global__ void to_float(const double* a, float* b, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) b[i] = (float)a[i];
}
static float* dev_float_copy(const double* a, int64_t n) {
  float* p;
  CK(cudaMalloc(&p, sizeof(float) * n));
  to_float<<<1184, 256>>>(a, p, n);
  CK(cudaGetLastError());
  return p;
}

static double* dev_uniform(int64_t n, double lo, double hi, uint64_t seed) {
  double* p;
  CK(cudaMalloc(&p, sizeof(double) * n));
  fill_uniform<<<1184, 256>>>(p, n, lo, hi, seed);
  CK(cudaGetLastError());
  return p;
}

template <class F>
static double time_ms(F&& launch, int iters) {
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a));
  CK(cudaEventCreate(&b));
  for (int i = 0; i < 3; ++i) launch();
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(a));
  for (int i = 0; i < iters; ++i) launch();
  CK(cudaEventRecord(b));
  CK(cudaEventSynchronize(b));
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, a, b));
  return ms / iters;
}

int main(int argc, char** argv) {
  if (argc < 3) { std::fprintf(stderr, "usage: kbench consts.bin gram|lin|rnea [n] [iters]\n"); return 1; }
  std::vector<double> c(84 + 252 + 36 + 6 + 6 + 12);
  FILE* fh = std::fopen(argv[1], "rb");
  if (!fh || std::fread(c.data(), sizeof(double), c.size(), fh) != c.size()) { std::fprintf(stderr, "cannot read %s\n", argv[1]); return 1; }
  std::fclose(fh);
  const double *hp = c.data(), *sim = hp + 84, *us = sim + 252, *tw0 = us + 36, *dtw0 = tw0 + 6, *sen = dtw0 + 6;
  const std::string what = argv[2];
  const int64_t n = argc > 3 ? std::atoll(argv[3]) : 12500000;
  const int iters = argc > 4 ? std::atoi(argv[4]) : 20;
  std::printf("%s\n", rbm_version());
  auto make = [&](unsigned flags) {
    rbm_model* m = nullptr;
    RB(rbm_model_create(6, hp, sim, us, tw0, dtw0, nullptr, nullptr, sen, flags, 0, &m));
    return m;
  };
  if (what == "gram") {
    double* q = dev_uniform(6 * n, -2, 2, 1);
    double* qd = dev_uniform(6 * n, -1, 1, 2);
    double* qdd = dev_uniform(6 * n, -3, 3, 3);
    double* f = dev_uniform(6 * n, -5, 5, 4);
    double packs[2][112];
    for (int v = 0; v < 2; ++v) {
      rbm_model* m = make(v == 1 ? RBM_FLAG_NO_TMA : 0);
      const size_t wsb = rbm_gram_workspace_bytes(m, n);
      void* ws;
      double* pack;
      CK(cudaMalloc(&ws, wsb));
      CK(cudaMalloc(&pack, sizeof(double) * 112));
      // ragged sizes first (tail tile, n < one tile per CTA), then the timed size
      const int64_t sizes[] = {n, 256, 257, 1000, 37889, 148 * 256 + 5};
      for (int64_t nn : sizes) {
        if (nn > n) continue;
        RB(rbm_regressor_gram_f64(m, q, qd, qdd, f, pack, ws, wsb, nn, n, nullptr));
        CK(cudaDeviceSynchronize());
        double h[112];
        CK(cudaMemcpy(h, pack, sizeof(h), cudaMemcpyDeviceToHost));
        if (nn == n) std::memcpy(packs[v], h, sizeof(h));
        static double ref_small[6][112];
        static int idx = 0;
        const int slot = idx++ % 6;
        if (v == 0) std::memcpy(ref_small[slot], h, sizeof(h));
        else {
          double worst = 0, scale = 0;
          for (int k = 0; k < 112; ++k) { scale = std::fmax(scale, std::fabs(ref_small[slot][k])); }
          for (int k = 0; k < 112; ++k) worst = std::fmax(worst, std::fabs(h[k] - ref_small[slot][k]));
          std::printf("  n=%lld  max|variant - default| / max|pack| = %.3e  (count %g)\n", (long long)nn, worst / scale, h[111]);
        }
      }
      const double ms = time_ms([&] { RB(rbm_regressor_gram_f64(m, q, qd, qdd, f, pack, ws, wsb, n, n, nullptr)); }, iters);
      std::printf("gram f64 %s: n=%lld  %.4f ms  %.3f G samples/s  %.1f GB/s algorithmic\n", v == 1 ? "direct loads" : "TMA pipeline", (long long)n, ms,
                  n / ms * 1e-6, 192.0 * n / ms * 1e-6);
      rbm_model_destroy(m);
      cudaFree(ws);
      cudaFree(pack);
    }
  } else if (what == "gram32") {
    double* t = dev_uniform(6 * n, -2, 2, 1);
    float* q = dev_float_copy(t, 6 * n);
    fill_uniform<<<1184, 256>>>(t, 6 * n, -1, 1, 2);
    float* qd = dev_float_copy(t, 6 * n);
    fill_uniform<<<1184, 256>>>(t, 6 * n, -3, 3, 3);
    float* qdd = dev_float_copy(t, 6 * n);
    fill_uniform<<<1184, 256>>>(t, 6 * n, -5, 5, 4);
    float* f = dev_float_copy(t, 6 * n);
    double ref[112];
    for (int v = 0; v < 2; ++v) {
      rbm_model* m = make(v ? RBM_FLAG_NO_TMA : 0);
      const size_t wsb = rbm_gram_workspace_bytes(m, n);
      void* ws;
      double* pack;
      CK(cudaMalloc(&ws, wsb));
      CK(cudaMalloc(&pack, sizeof(double) * 112));
      RB(rbm_regressor_gram_f32(m, q, qd, qdd, f, pack, ws, wsb, n, n, nullptr));
      double h[112];
      CK(cudaMemcpy(h, pack, sizeof(h), cudaMemcpyDeviceToHost));
      if (v == 0) std::memcpy(ref, h, sizeof(h));
      else {
        double worst = 0, scale = 0;
        for (int k = 0; k < 112; ++k) scale = std::fmax(scale, std::fabs(ref[k]));
        for (int k = 0; k < 112; ++k) worst = std::fmax(worst, std::fabs(h[k] - ref[k]));
        std::printf("  max|variant - default| / max|pack| = %.3e\n", worst / scale);
      }
      const double ms = time_ms([&] { RB(rbm_regressor_gram_f32(m, q, qd, qdd, f, pack, ws, wsb, n, n, nullptr)); }, iters);
      std::printf("gram f32 %s: n=%lld  %.4f ms  %.3f G samples/s  %.1f GB/s algorithmic\n", v ? "direct loads" : "TMA pipeline", (long long)n, ms, n / ms * 1e-6,
                  96.0 * n / ms * 1e-6);
      rbm_model_destroy(m);
    }
  } else if (what == "lin") {
    double* q = dev_uniform(6 * n, -2, 2, 1);
    double* qd = dev_uniform(6 * n, -1, 1, 2);
    double* u = dev_uniform(6 * n, -3, 3, 3);
    double *A, *B;
    CK(cudaMalloc(&A, sizeof(double) * 144 * n));
    CK(cudaMalloc(&B, sizeof(double) * 72 * n));
    rbm_model* m = make(0);
    const double ms = time_ms([&] { RB(rbm_linearize_f64(m, q, qd, u, 0.002, 1e-6, 1, A, B, nullptr, n, n, nullptr)); }, iters);
    std::printf("linearize f64: n=%lld  %.4f ms  %.3f G states/s  %.1f GB/s algorithmic\n", (long long)n, ms, n / ms * 1e-6, 1824.0 * n / ms * 1e-6);
    std::vector<double> h(144);
    CK(cudaMemcpy(h.data(), A, sizeof(double) * 144, cudaMemcpyDeviceToHost));
    double cs = 0;
    for (double x : h) cs += x;
    std::printf("  checksum A[0] = %.15g\n", cs);
  } else if (what == "rnea") {
    double* q = dev_uniform(6 * n, -2, 2, 1);
    double* qd = dev_uniform(6 * n, -1, 1, 2);
    double* qdd = dev_uniform(6 * n, -3, 3, 3);
    double* tau;
    CK(cudaMalloc(&tau, sizeof(double) * 6 * n));
    rbm_model* m = make(0);
    const double ms = time_ms([&] { RB(rbm_rnea_f64(m, q, qd, qdd, tau, nullptr, nullptr, n, n, nullptr)); }, iters);
    std::printf("rnea f64: n=%lld  %.4f ms  %.3f G samples/s  %.1f GB/s algorithmic\n", (long long)n, ms, n / ms * 1e-6, 192.0 * n / ms * 1e-6);
  } else {
    std::fprintf(stderr, "unknown workload %s\n", what.c_str());
    return 1;
  }
  CK(cudaDeviceSynchronize());
  return 0;
}
