// Compile-time sparse spatial algebra for the structure-specialised RNEA kernels.
//
// Every 3-vector component is either a run-time scalar T or the empty type `Z`
// (a STRUCTURAL zero).  Operator overloads propagate Z through +, -, * so that the
// generic Modern-Robotics recursion (reference dynamics/dynamics.py:125-147), written once,
// instantiates to exactly the multiplications and additions that a given robot structure
// needs (nvcc may not fold x*0.0 or x+0.0 itself: IEEE semantics forbid it).  Signed-
// permutation home rotations are template integers, so applying them costs nothing but
// sign flips that fold into the neighbouring FMAs.
#pragma once
#include <cuda_runtime.h>
#include <type_traits>

namespace rbm {

#define RBM_HD __host__ __device__ __forceinline__

struct Z {};  // structural zero

template <class T>
using if_fp = std::enable_if_t<std::is_floating_point<T>::value, int>;

// ---- scalar algebra over {Z, T} --------------------------------------------------------
RBM_HD Z operator+(Z, Z) { return {}; }
template <class T, if_fp<T> = 0> RBM_HD T operator+(Z, T b) { return b; }
template <class T, if_fp<T> = 0> RBM_HD T operator+(T a, Z) { return a; }
RBM_HD Z operator-(Z, Z) { return {}; }
template <class T, if_fp<T> = 0> RBM_HD T operator-(Z, T b) { return -b; }
template <class T, if_fp<T> = 0> RBM_HD T operator-(T a, Z) { return a; }
RBM_HD Z operator-(Z) { return {}; }
RBM_HD Z operator*(Z, Z) { return {}; }
template <class T, if_fp<T> = 0> RBM_HD Z operator*(Z, T) { return {}; }
template <class T, if_fp<T> = 0> RBM_HD Z operator*(T, Z) { return {}; }

template <class T> RBM_HD T to_scalar(Z) { return T(0); }
template <class T> RBM_HD T to_scalar(T v) { return v; }

template <class A> struct is_zero : std::false_type {};
template <> struct is_zero<Z> : std::true_type {};

// ---- typed 3-vectors -------------------------------------------------------------------
template <class A, class B, class C>
struct V3 {
  A x;
  B y;
  C z;
};
using Z3 = V3<Z, Z, Z>;

template <class A, class B, class C>
RBM_HD V3<A, B, C> mk3(A a, B b, C c) { return V3<A, B, C>{a, b, c}; }

template <int I, class A, class B, class C>
RBM_HD auto get(const V3<A, B, C>& v) {
  static_assert(I >= 0 && I < 3, "component index");
  if constexpr (I == 0) return v.x;
  else if constexpr (I == 1) return v.y;
  else return v.z;
}

template <class A, class B, class C, class D, class E, class F>
RBM_HD auto operator+(const V3<A, B, C>& a, const V3<D, E, F>& b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
template <class A, class B, class C, class D, class E, class F>
RBM_HD auto operator-(const V3<A, B, C>& a, const V3<D, E, F>& b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
template <class A, class B, class C>
RBM_HD auto operator-(const V3<A, B, C>& a) { return mk3(-a.x, -a.y, -a.z); }
template <class S, class A, class B, class C>
RBM_HD auto scale(S s, const V3<A, B, C>& a) { return mk3(s * a.x, s * a.y, s * a.z); }
template <class A, class B, class C, class D, class E, class F>
RBM_HD auto cross(const V3<A, B, C>& a, const V3<D, E, F>& b) {
  return mk3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
// a x e3  and  e3-scaled helpers (joint axes are +z in the joint frame)
template <class A, class B, class C>
RBM_HD auto cross_e3(const V3<A, B, C>& a) { return mk3(a.y, -a.x, Z{}); }

// ---- signed permutations: row i of R has the entry sign(Ri) in column |Ri|-1 ---------------
template <int R0, int R1, int R2>
struct SPerm {
  static constexpr int r0 = R0, r1 = R1, r2 = R2;
};

template <int R, class V>
RBM_HD auto sp_row(const V& v) {
  static_assert(R != 0 && R >= -3 && R <= 3, "signed column index in 1..3");
  if constexpr (R > 0) return get<R - 1>(v);
  else return -get<-R - 1>(v);
}
// y = P x
template <class P, class V>
RBM_HD auto sp_apply(const V& v) { return mk3(sp_row<P::r0>(v), sp_row<P::r1>(v), sp_row<P::r2>(v)); }

template <int J, class P, class V>
RBM_HD auto sp_col(const V& v) {  // (P^T v)_J
  constexpr int a0 = P::r0 < 0 ? -P::r0 : P::r0, a1 = P::r1 < 0 ? -P::r1 : P::r1;
  if constexpr (a0 - 1 == J) { if constexpr (P::r0 > 0) return v.x; else return -v.x; }
  else if constexpr (a1 - 1 == J) { if constexpr (P::r1 > 0) return v.y; else return -v.y; }
  else { if constexpr (P::r2 > 0) return v.z; else return -v.z; }
}
// y = P^T x
template <class P, class V>
RBM_HD auto sp_apply_T(const V& v) { return mk3(sp_col<0, P>(v), sp_col<1, P>(v), sp_col<2, P>(v)); }

// ---- rotations about the joint axis ----------------------------------------------------------
// Rz(-q) x = ( c x + s y, -s x + c y, z )   [exp(-S q) for a revolute +z screw, Eq. 8.50]
template <class T, class A, class B, class C>
RBM_HD auto rotz_neg(T c, T s, const V3<A, B, C>& v) { return mk3(c * v.x + s * v.y, c * v.y - s * v.x, v.z); }
// Rz(+q) x = ( c x - s y,  s x + c y, z )   [its transpose, used by the backward sweep]
template <class T, class A, class B, class C>
RBM_HD auto rotz_pos(T c, T s, const V3<A, B, C>& v) { return mk3(c * v.x - s * v.y, s * v.x + c * v.y, v.z); }

// symmetric 3x3 times vector; I = [xx, yy, zz, xy, yz, zx]
template <class T, class A, class B, class C>
RBM_HD auto sym3_mul(const T* I, const V3<A, B, C>& v) {
  return mk3(I[0] * v.x + I[3] * v.y + I[5] * v.z, I[3] * v.x + I[1] * v.y + I[4] * v.z, I[5] * v.x + I[4] * v.y + I[2] * v.z);
}
template <class T>
RBM_HD V3<T, T, T> ld3(const T* p) { return V3<T, T, T>{p[0], p[1], p[2]}; }

}  // namespace rbm
