// sin/cos of several independent joint angles at once, written for instruction-level parallelism.
//
// libdevice's sincos() carries a branch per call (the Payne-Hanek slow path for |x| > 105615), which keeps the compiler from
// interleaving the three wrist angles of one sample: three ~35-instruction dependent FP64 chains then run back to back and,
// at the low occupancy of the register-heavy kernels, their latency is exposed.  Here the quadrant reduction (3-constant
// Cody-Waite, exact products through FMA) and the two minimax polynomials are branch-free, so N angles give 2N independent
// chains; ONE test per sample routes the whole group to libdevice when any angle is outside the fast range (or not finite).
// Rounding to the nearest quadrant uses the 1.5*2^52 (1.5*2^23) shift, so no int<->float conversions sit on the chain.
// Accuracy of the fast path: < 2e-16 absolute in double, < 2.5e-7 in float (tests/test_trig_host.py); the parity bars on
// tau are 1e-9 / 1e-4 relative.
#pragma once
#include <cuda_runtime.h>

namespace rbm {

#ifndef RBM_HD
#define RBM_HD __host__ __device__ __forceinline__
#endif

RBM_HD int lo_word(double t) {
#ifdef __CUDA_ARCH__
  return __double2loint(t);
#else
  union { double d; long long i; } u;
  u.d = t;
  return (int)(u.i & 0xffffffffLL);
#endif
}
RBM_HD int float_bits(float t) {
#ifdef __CUDA_ARCH__
  return __float_as_int(t);
#else
  union { float f; int i; } u;
  u.f = t;
  return u.i;
#endif
}

constexpr double kTrigFastMaxF64 = 1.0e5;
constexpr float kTrigFastMaxF32 = 4.0e4f;

// |x| <= kTrigFastMaxF64
RBM_HD void sincos_core(double x, double& s, double& c) {
  const double kShift = 6755399441055744.0;  // 1.5 * 2^52
  const double t = fma(x, 0.63661977236758138, kShift);
  const int j = lo_word(t);                  // round(x * 2/pi), low 32 bits, two's complement
  const double jd = t - kShift;
  double r = fma(-jd, 1.5707963267948966e+00, x);
  r = fma(-jd, 6.1232339957367574e-17, r);
  r = fma(-jd, 8.4784276603688985e-32, r);
  const double z = r * r;
  // fdlibm __kernel_sin / __kernel_cos coefficients on [-pi/4, pi/4]
  double ps = fma(z, 1.58969099521155010221e-10, -2.50507602534068634195e-08);
  ps = fma(z, ps, 2.75573137070700676789e-06);
  ps = fma(z, ps, -1.98412698298579493134e-04);
  ps = fma(z, ps, 8.33333333332248946124e-03);
  ps = fma(z, ps, -1.66666666666666324348e-01);
  const double sr = fma(r * z, ps, r);
  double pc = fma(z, -1.13596475577881948265e-11, 2.08757232129817482790e-09);
  pc = fma(z, pc, -2.75573143513906633035e-07);
  pc = fma(z, pc, 2.48015872894767294178e-05);
  pc = fma(z, pc, -1.38888888888741095749e-03);
  pc = fma(z, pc, 4.16666666666666019037e-02);
  const double cr = fma(z * z, pc, fma(z, -0.5, 1.0));
  const bool swap = (j & 1) != 0;
  const double sa = swap ? cr : sr, ca = swap ? sr : cr;
  s = (j & 2) ? -sa : sa;
  c = ((j + 1) & 2) ? -ca : ca;
}

// |x| <= kTrigFastMaxF32
RBM_HD void sincos_core(float x, float& s, float& c) {
  const float kShift = 12582912.0f;  // 1.5 * 2^23
  const float t = fmaf(x, 0.636619772f, kShift);
  const int j = float_bits(t);       // low bits = round(x * 2/pi) (the shift's own low bits are zero)
  const float jf = t - kShift;
  float r = fmaf(-jf, 1.57079601e+00f, x);
  r = fmaf(-jf, 3.13916473e-07f, r);
  r = fmaf(-jf, 5.39030253e-15f, r);
  const float z = r * r;
  float ps = fmaf(z, 2.86567956e-6f, -1.98559923e-4f);
  ps = fmaf(z, ps, 8.33338592e-3f);
  ps = fmaf(z, ps, -1.66666672e-1f);
  const float sr = fmaf(r * z, ps, r);
  float pc = fmaf(z, 2.44677067e-5f, -1.38877297e-3f);
  pc = fmaf(z, pc, 4.16666567e-2f);
  pc = fmaf(z, pc, -5.00000000e-1f);
  const float cr = fmaf(z, pc, 1.0f);
  const bool swap = (j & 1) != 0;
  const float sa = swap ? cr : sr, ca = swap ? sr : cr;
  s = (j & 2) ? -sa : sa;
  c = ((j + 1) & 2) ? -ca : ca;
}

RBM_HD void sincos_lib(double x, double* s, double* c) { sincos(x, s, c); }
RBM_HD void sincos_lib(float x, float* s, float* c) { sincosf(x, s, c); }
RBM_HD bool trig_fast_ok(double x) { return fabs(x) <= kTrigFastMaxF64; }   // false for NaN
RBM_HD bool trig_fast_ok(float x) { return fabsf(x) <= kTrigFastMaxF32; }

// N independent angles: branch-free fast path for the group, libdevice when any of them is out of range / not finite
template <int N, class T>
RBM_HD void sincos_group(const T (&x)[N], T (&s)[N], T (&c)[N]) {
  bool ok = true;
#pragma unroll
  for (int i = 0; i < N; ++i) ok = ok && trig_fast_ok(x[i]);
  if (ok) {
#pragma unroll
    for (int i = 0; i < N; ++i) sincos_core(x[i], s[i], c[i]);
  } else {
#pragma unroll
    for (int i = 0; i < N; ++i) sincos_lib(x[i], &s[i], &c[i]);
  }
}

template <class T>
RBM_HD void sincos_one(T x, T* s, T* c) {
  if (trig_fast_ok(x)) sincos_core(x, *s, *c);
  else sincos_lib(x, s, c);
}

}  // namespace rbm
