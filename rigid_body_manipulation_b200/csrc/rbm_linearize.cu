// Batched discrete-time linearisation x+ = A x + B u of the simulated plant, one state per thread.
//
// Replaces dynamics.StateSpace.update_matrices (reference dynamics/dynamics.py:41-46), i.e. MuJoCo's
// mjd_transitionFD(m, d, eps, centered, A, B, ...) consumed by controllers/lqr.py:43-49, for the reference's
// plant: nv = nu = nj, na = 0, motors with unit gear (ctrl == joint force), no damping / friction / contacts,
// default semi-implicit Euler:   qdd = M(q)^-1 (u - h(q, qd)),   qd+ = qd + dt qdd,   q+ = q + dt qd+ ,
// state x = [q; qd].  Everything is built from the same RNEA used for the feed-forward:
//
//   M(q) e_j   = ID(q, 0, e_j) with the base acceleration (gravity) switched off            (nj evaluations)
//   h(q, qd)   = ID(q, qd, 0)                                                                (1)
//   d qdd / dx = -M^-1  d ID(q, qd, qdd_nom) / dx   by finite differences of the INVERSE dynamics at fixed
//                qdd_nom -- centred (2 * 2nj evaluations) or forward (2nj + 1), step `eps`
//   A = [[1 + dt^2 Q, dt (1 + dt V)], [dt Q, 1 + dt V]],  B = [[dt^2 M^-1], [dt M^-1]],  Q = dqdd/dq, V = dqdd/dqd
//
// (differentiating the inverse dynamics instead of re-solving the forward dynamics at every perturbed state gives
// the same Jacobian with one Cholesky factorisation instead of 4nj+1; the difference is O(eps^2).)
//
// Output layout is element-major so that every store is coalesced: entry (r, c) of A for state s is at
// A[(r * 2nj + c) * ld + s], of B at B[(r * nj + c) * ld + s]  (228 scalars written per state for nj = 6).
#include "rbm_internal.h"
#include "rbm_rnea.cuh"

namespace rbm {

constexpr int kLinBlock = 128;
#ifndef RBM_LIN_MINB
#define RBM_LIN_MINB 3
#endif

// ---- inverse-dynamics evaluators ---------------------------------------------------------------------
template <class T, class D>
struct FastEval {
  static constexpr int NJ = 6;
  static constexpr int MAXJ = 6;
  const FastParams<T>& P;
  __device__ __forceinline__ int nj() const { return 6; }
  __device__ __forceinline__ static bool is_hinge(int j) {
    constexpr bool h[6] = {D::L0::jk == JOINT_RZ, D::L1::jk == JOINT_RZ, D::L2::jk == JOINT_RZ, D::L3::jk == JOINT_RZ, D::L4::jk == JOINT_RZ, D::L5::jk == JOINT_RZ};
    bool r = false;
#pragma unroll
    for (int k = 0; k < 6; ++k) r = (k == j) ? h[k] : r;
    return r;
  }
  __device__ __forceinline__ void trig(const T (&q)[6], T (&c)[6], T (&s)[6]) const { fast_sincos<T, D>(q, c, s); }
  __device__ __forceinline__ static bool q_matters(int j) { return D::q_matters(j); }
  __device__ __forceinline__ void id(const T (&q)[6], const T (&c)[6], const T (&s)[6], const T (&qd)[6], const T (&qdd)[6], T (&tau)[6]) const {
    FastResult<T> r;
    fast_rnea_core<T, D, true>(P, P.g, q, c, s, qd, qdd, r);
#pragma unroll
    for (int k = 0; k < 6; ++k) tau[k] = r.tau[k];
  }
  // acceleration-only evaluation without gravity: column of the joint-space inertia matrix for qdd = e_j
  __device__ __forceinline__ void id_inertia(const T (&q)[6], const T (&c)[6], const T (&s)[6], const T (&qdd)[6], T (&tau)[6]) const {
    FastResult<T> r;
    fast_rnea_core<T, D, true, false, false>(P, P.g, q, c, s, qdd /* unused */, qdd, r);
#pragma unroll
    for (int k = 0; k < 6; ++k) tau[k] = r.tau[k];
  }
};

template <class T>
struct GenericEval {
  static constexpr int NJ = 0;
  static constexpr int MAXJ = RBM_MAX_JOINTS;
  const T* sp;     // staged parameters
  const T* zero;   // 18 zeros: base twist / acceleration / tip wrench switched off
  int nj_;
  __device__ __forceinline__ int nj() const { return nj_; }
  __device__ __forceinline__ static bool is_hinge(int) { return false; }  // generic_rnea evaluates its own trigonometry
  __device__ __forceinline__ void trig(const T (&)[MAXJ], T (&)[MAXJ], T (&)[MAXJ]) const {}
  __device__ __forceinline__ static bool q_matters(int) { return true; }
  __device__ __forceinline__ void id(const T (&q)[MAXJ], const T (&)[MAXJ], const T (&)[MAXJ], const T (&qd)[MAXJ], const T (&qdd)[MAXJ],
                                     T (&tau)[MAXJ]) const {
    generic_rnea<T, 0>(sp, sp, nj_, q, qd, qdd, tau, nullptr, nullptr, nullptr, nullptr, nullptr);
  }
  __device__ __forceinline__ void id_inertia(const T (&q)[MAXJ], const T (&)[MAXJ], const T (&)[MAXJ], const T (&qdd)[MAXJ], T (&tau)[MAXJ]) const {
    T qd0[MAXJ];
#pragma unroll
    for (int k = 0; k < MAXJ; ++k) qd0[k] = T(0);
    generic_rnea<T, 0>(sp, zero, nj_, q, qd0, qdd, tau, nullptr, nullptr, nullptr, nullptr, nullptr);
  }
};

// ---- dense SPD solve (Cholesky, in place in the lower triangle; the diagonal holds 1 / L_ii) ------------
template <class T, int MAXJ>
__device__ __forceinline__ void cholesky(T (&M)[MAXJ][MAXJ], int n) {
#pragma unroll
  for (int j = 0; j < MAXJ; ++j) {
    if (j < n) {
      T d = M[j][j];
#pragma unroll
      for (int k = 0; k < j; ++k) d -= M[j][k] * M[j][k];
      const T inv = T(1) / sqrt(d);
      M[j][j] = inv;
#pragma unroll
      for (int i = j + 1; i < MAXJ; ++i) {
        if (i < n) {
          T v = M[i][j];
#pragma unroll
          for (int k = 0; k < j; ++k) v -= M[i][k] * M[j][k];
          M[i][j] = v * inv;
        }
      }
    }
  }
}
template <class T, int MAXJ>
__device__ __forceinline__ void chol_solve(const T (&L)[MAXJ][MAXJ], int n, T (&b)[MAXJ]) {
#pragma unroll
  for (int i = 0; i < MAXJ; ++i) {
    if (i < n) {
      T v = b[i];
#pragma unroll
      for (int k = 0; k < i; ++k) v -= L[i][k] * b[k];
      b[i] = v * L[i][i];
    }
  }
#pragma unroll
  for (int i = MAXJ - 1; i >= 0; --i) {
    if (i < n) {
      T v = b[i];
#pragma unroll
      for (int k = i + 1; k < MAXJ; ++k)
        if (k < n) v -= L[k][i] * b[k];
      b[i] = v * L[i][i];
    }
  }
}

// ---- the per-state algorithm ---------------------------------------------------------------------------
template <class T, class E>
__device__ __forceinline__ void linearize_state(const E& ev, const T* __restrict__ q_in, const T* __restrict__ qd_in, const T* __restrict__ u_in,
                                                T dt, T eps, bool centered, T* __restrict__ A, T* __restrict__ B, T* __restrict__ qdd_out, int64_t s,
                                                int64_t ld) {
  constexpr int MJ = E::MAXJ;
  const int nj = ev.nj();
  const int ns = 2 * nj;
  T q[MJ], qd[MJ], u[MJ], c[MJ], sn[MJ], zero[MJ];
#pragma unroll
  for (int k = 0; k < MJ; ++k) {
    const bool on = k < nj;
    q[k] = on ? __ldg(q_in + k * ld + s) : T(0);
    qd[k] = on ? __ldg(qd_in + k * ld + s) : T(0);
    u[k] = (on && u_in) ? __ldg(u_in + k * ld + s) : T(0);
    zero[k] = T(0);
    c[k] = T(1);
    sn[k] = T(0);
  }
  ev.trig(q, c, sn);

  // joint-space inertia matrix, column by column
  T M[MJ][MJ];
#pragma unroll 1
  for (int j = 0; j < nj; ++j) {
    T e[MJ], col[MJ];
#pragma unroll
    for (int k = 0; k < MJ; ++k) e[k] = (k == j) ? T(1) : T(0);
    ev.id_inertia(q, c, sn, e, col);
#pragma unroll
    for (int r = 0; r < MJ; ++r)
#pragma unroll
      for (int k = 0; k < MJ; ++k)
        if (k == j) M[r][k] = col[r];
  }
  // bias forces and nominal acceleration
  T h[MJ], qdd[MJ];
  ev.id(q, c, sn, qd, zero, h);
  cholesky<T, MJ>(M, nj);
#pragma unroll
  for (int k = 0; k < MJ; ++k) qdd[k] = u[k] - h[k];
  chol_solve<T, MJ>(M, nj, qdd);
  if (qdd_out) {
#pragma unroll
    for (int k = 0; k < MJ; ++k)
      if (k < nj) qdd_out[k * ld + s] = qdd[k];
  }
  // B = [[dt^2 M^-1], [dt M^-1]]
#pragma unroll 1
  for (int j = 0; j < nj; ++j) {
    T x[MJ];
#pragma unroll
    for (int k = 0; k < MJ; ++k) x[k] = (k == j) ? T(1) : T(0);
    chol_solve<T, MJ>(M, nj, x);
#pragma unroll
    for (int r = 0; r < MJ; ++r) {
      if (r < nj) {
        B[((int64_t)r * nj + j) * ld + s] = dt * dt * x[r];
        B[((int64_t)(nj + r) * nj + j) * ld + s] = dt * x[r];
      }
    }
  }
  // reference value for forward differences: ID at the nominal point (== u up to round-off)
  T tau0[MJ];
  if (!centered) ev.id(q, c, sn, qd, qdd, tau0);
  const T inv_step = centered ? T(1) / (T(2) * eps) : T(1) / eps;

  // position columns, then velocity columns
#pragma unroll 1
  for (int pass = 0; pass < 2; ++pass) {
#pragma unroll 1
    for (int j = 0; j < nj; ++j) {
      T xp[MJ], cp[MJ], sp_[MJ], tp[MJ], tm[MJ];
      const T* src = pass == 0 ? q : qd;
      const int col = pass * nj + j;
      if (pass == 0 && !E::q_matters(j)) {  // d tau / d q_j is structurally zero: the column is that of the identity map
#pragma unroll
        for (int r = 0; r < MJ; ++r) {
          if (r < nj) {
            A[((int64_t)r * ns + col) * ld + s] = (r == j) ? T(1) : T(0);
            A[((int64_t)(nj + r) * ns + col) * ld + s] = T(0);
          }
        }
        continue;
      }
#pragma unroll
      for (int k = 0; k < MJ; ++k) { xp[k] = src[k] + ((k == j) ? eps : T(0)); cp[k] = c[k]; sp_[k] = sn[k]; }
      if (pass == 0 && E::is_hinge(j)) {
        T qj = T(0), cj, sj;
#pragma unroll
        for (int k = 0; k < MJ; ++k) qj = (k == j) ? xp[k] : qj;
        sincos_t(qj, &sj, &cj);
#pragma unroll
        for (int k = 0; k < MJ; ++k) { cp[k] = (k == j) ? cj : cp[k]; sp_[k] = (k == j) ? sj : sp_[k]; }
      }
      if (pass == 0) ev.id(xp, cp, sp_, qd, qdd, tp);
      else ev.id(q, c, sn, xp, qdd, tp);
      if (centered) {
#pragma unroll
        for (int k = 0; k < MJ; ++k) { xp[k] = src[k] - ((k == j) ? eps : T(0)); cp[k] = c[k]; sp_[k] = sn[k]; }
        if (pass == 0 && E::is_hinge(j)) {
          T qj = T(0), cj, sj;
#pragma unroll
          for (int k = 0; k < MJ; ++k) qj = (k == j) ? xp[k] : qj;
          sincos_t(qj, &sj, &cj);
#pragma unroll
          for (int k = 0; k < MJ; ++k) { cp[k] = (k == j) ? cj : cp[k]; sp_[k] = (k == j) ? sj : sp_[k]; }
        }
        if (pass == 0) ev.id(xp, cp, sp_, qd, qdd, tm);
        else ev.id(q, c, sn, xp, qdd, tm);
      } else {
#pragma unroll
        for (int k = 0; k < MJ; ++k) tm[k] = tau0[k];
      }
      T x[MJ];
#pragma unroll
      for (int k = 0; k < MJ; ++k) x[k] = -(tp[k] - tm[k]) * inv_step;
      chol_solve<T, MJ>(M, nj, x);   // column j of Q (pass 0) or V (pass 1)
#pragma unroll
      for (int r = 0; r < MJ; ++r) {
        if (r < nj) {
          const T delta = (r == j) ? T(1) : T(0);
          T top, bot;
          if (pass == 0) { top = delta + dt * dt * x[r]; bot = dt * x[r]; }
          else { top = dt * (delta + dt * x[r]); bot = delta + dt * x[r]; }
          A[((int64_t)r * ns + col) * ld + s] = top;
          A[((int64_t)(nj + r) * ns + col) * ld + s] = bot;
        }
      }
    }
  }
}

template <class T, class D>
__global__ void __launch_bounds__(kLinBlock, RBM_LIN_MINB) k_linearize_fast(const __grid_constant__ FastParams<T> P, const T* __restrict__ q, const T* __restrict__ qd,
                                                              const T* __restrict__ u, T dt, T eps, int centered, T* __restrict__ A, T* __restrict__ B,
                                                              T* __restrict__ qdd_out, int64_t n, int64_t ld) {
  const int64_t s = (int64_t)blockIdx.x * kLinBlock + threadIdx.x;
  if (s >= n) return;
  FastEval<T, D> ev{P};
  linearize_state<T>(ev, q, qd, u, dt, eps, centered != 0, A, B, qdd_out, s, ld);
}

template <class T>
__global__ void __launch_bounds__(kLinBlock) k_linearize_generic(const T* __restrict__ gp, int nj, int nparams, const T* __restrict__ q,
                                                                 const T* __restrict__ qd, const T* __restrict__ u, T dt, T eps, int centered,
                                                                 T* __restrict__ A, T* __restrict__ B, T* __restrict__ qdd_out, int64_t n, int64_t ld) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* sp = reinterpret_cast<T*>(smem_raw);
  T* zero = sp + nparams;
  for (int i = threadIdx.x; i < nparams; i += blockDim.x) sp[i] = gp[i];
  for (int i = threadIdx.x; i < 18; i += blockDim.x) zero[i] = T(0);
  __syncthreads();
  const int64_t s = (int64_t)blockIdx.x * kLinBlock + threadIdx.x;
  if (s >= n) return;
  GenericEval<T> ev{sp, zero, nj};
  linearize_state<T>(ev, q, qd, u, dt, eps, centered != 0, A, B, qdd_out, s, ld);
}

template <class T>
int launch_linearize(const rbm_model* m, const T* q, const T* qd, const T* u, double dt, double eps, int centered, T* A, T* B, T* qdd, int64_t n,
                     int64_t ld, cudaStream_t st) {
  if (n == 0) return RBM_OK;
  const unsigned grid = (unsigned)((n + kLinBlock - 1) / kLinBlock);
  if (m->path == PATH_SEQ_ISO) {
    k_linearize_fast<T, SeqIso><<<grid, kLinBlock, 0, st>>>(ModelView<T>::fast(m), q, qd, u, (T)dt, (T)eps, centered, A, B, qdd, n, ld);
  } else if (m->path == PATH_SEQ_RIGID) {
    k_linearize_fast<T, SeqRigid><<<grid, kLinBlock, 0, st>>>(ModelView<T>::fast(m), q, qd, u, (T)dt, (T)eps, centered, A, B, qdd, n, ld);
  } else {
    const int np = generic_param_count(m->nj);
    k_linearize_generic<T><<<grid, kLinBlock, sizeof(T) * (np + 18), st>>>(ModelView<T>::generic(m), m->nj, np, q, qd, u, (T)dt, (T)eps, centered, A, B, qdd,
                                                                         n, ld);
  }
  RBM_CUDA_TRY(cudaGetLastError());
  return RBM_OK;
}

template int launch_linearize<double>(const rbm_model*, const double*, const double*, const double*, double, double, int, double*, double*, double*, int64_t,
                                      int64_t, cudaStream_t);

}  // namespace rbm
