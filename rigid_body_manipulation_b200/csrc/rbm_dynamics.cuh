// Per-state dynamics algorithms shared by the kernels of rbm_linearize.cu -- inverse-dynamics evaluators over both kernel families,
// the dense 6x6 (<= 16x16) SPD solve, the LQR linearisation of one state, one forward-dynamics transition, and the closed-loop
// rollout of one environment.  Everything is __host__ __device__ so that tests/host_harness can run the SAME code on the CPU
// (`-m "not gpu"` parity against the oracle); the kernels themselves, their launch geometry and the ABI live in rbm_linearize.cu.
#pragma once
#include "rbm_rnea.cuh"

namespace rbm {

// read-only global load: the non-coherent path on the device, a plain load on the host
template <class T>
RBM_HD T ld_ro(const T* p) {
#ifdef __CUDA_ARCH__
  return __ldg(p);
#else
  return *p;
#endif
}

// ---- inverse-dynamics evaluators ---------------------------------------------------------------------
template <class T, class D>
struct FastEval {
  static constexpr int NJ = 6;
  static constexpr int MAXJ = 6;
  const FastParams<T>& P;
  RBM_HD int nj() const { return 6; }
  RBM_HD static bool is_hinge(int j) {
    constexpr bool h[6] = {D::L0::jk == JOINT_RZ, D::L1::jk == JOINT_RZ, D::L2::jk == JOINT_RZ, D::L3::jk == JOINT_RZ, D::L4::jk == JOINT_RZ, D::L5::jk == JOINT_RZ};
    bool r = false;
#pragma unroll
    for (int k = 0; k < 6; ++k) r = (k == j) ? h[k] : r;
    return r;
  }
  RBM_HD void trig(const T (&q)[6], T (&c)[6], T (&s)[6]) const { fast_sincos<T, D>(q, c, s); }
  RBM_HD static bool q_matters(int j) { return D::q_matters(j); }
  RBM_HD static bool qd_matters(int j) { return D::qd_matters(j); }
  RBM_HD void id(const T (&q)[6], const T (&c)[6], const T (&s)[6], const T (&qd)[6], const T (&qdd)[6], T (&tau)[6]) const {
    FastResult<T> r;
    fast_rnea_core<T, D, true>(P, P.g, q, c, s, qd, qdd, r);
#pragma unroll
    for (int k = 0; k < 6; ++k) tau[k] = r.tau[k];
  }
  // twist and twist rate of the last link only (forward sweep)
  RBM_HD void last_twists(const T (&q)[6], const T (&c)[6], const T (&s)[6], const T (&qd)[6], const T (&qdd)[6], T (&V)[6],
                                              T (&dV)[6]) const {
    FastResult<T> r;
    fast_rnea_core<T, D, false>(P, P.g, q, c, s, qd, qdd, r);
#pragma unroll
    for (int k = 0; k < 3; ++k) { V[k] = r.v[k]; V[3 + k] = r.w[k]; dV[k] = r.a[k]; dV[3 + k] = r.l[k]; }
  }
  RBM_HD const T* sensor_pose() const { return P.senR; }  // [R (9) | t (3)]
  // velocity-product term alone, C(q, qd) = ID(q, qd, 0) without gravity.  tau = M qdd + C + g, so finite differences over qd
  // at fixed (q, qdd) only need this part (and it is an exact quadratic form in qd: the centred difference has no truncation)
  RBM_HD void id_velocity(const T (&q)[6], const T (&c)[6], const T (&s)[6], const T (&qd)[6], T (&tau)[6]) const {
    FastResult<T> r;
    fast_rnea_core<T, D, true, true, false, false>(P, P.g, q, c, s, qd, qd /* unused */, r);
#pragma unroll
    for (int k = 0; k < 6; ++k) tau[k] = r.tau[k];
  }
  // acceleration-only evaluation without gravity: column of the joint-space inertia matrix for qdd = e_j
  RBM_HD void id_inertia(const T (&q)[6], const T (&c)[6], const T (&s)[6], const T (&qdd)[6], T (&tau)[6]) const {
    FastResult<T> r;
    fast_rnea_core<T, D, true, false, false>(P, P.g, q, c, s, qdd /* unused */, qdd, r);
#pragma unroll
    for (int k = 0; k < 6; ++k) tau[k] = r.tau[k];
  }
  // bias forces h(q, qd) = ID(q, qd, 0): joint accelerations as structural zeros
  RBM_HD void id_bias(const T (&q)[6], const T (&c)[6], const T (&s)[6], const T (&qd)[6], T (&tau)[6]) const {
    FastResult<T> r;
    fast_rnea_core<T, D, true, true, true, false>(P, P.g, q, c, s, qd, qd /* unused */, r);
#pragma unroll
    for (int k = 0; k < 6; ++k) tau[k] = r.tau[k];
  }
  // column J of the inertia matrix with the other joint accelerations as structural zeros (fast_rnea_core, ONEHOT)
  template <int J>
  RBM_HD void inertia_col(const T (&q)[6], const T (&c)[6], const T (&s)[6], T (&M)[6][6]) const {
    FastResult<T> r;
    fast_rnea_core<T, D, true, false, false, true, J>(P, P.g, q, c, s, q /* unused */, q /* unused */, r);
#pragma unroll
    for (int k = 0; k < 6; ++k) M[k][J] = r.tau[k];
  }
  // joint-space inertia matrix M(q), column by column
  RBM_HD void inertia_matrix(const T (&q)[6], const T (&c)[6], const T (&s)[6], T (&M)[6][6]) const {
    inertia_col<0>(q, c, s, M); inertia_col<1>(q, c, s, M); inertia_col<2>(q, c, s, M);
    inertia_col<3>(q, c, s, M); inertia_col<4>(q, c, s, M); inertia_col<5>(q, c, s, M);
  }
};

template <class T>
struct GenericEval {
  static constexpr int NJ = 0;
  static constexpr int MAXJ = RBM_MAX_JOINTS;
  const T* sp;     // staged parameters
  const T* zero;   // 18 zeros: base twist / acceleration / tip wrench switched off
  int nj_;
  RBM_HD int nj() const { return nj_; }
  // generic_rnea writes tau[0 .. nj): the padding entries of the MAXJ-sized arrays must not carry stack garbage into the masked
  // dense algebra downstream (0 * NaN)
  RBM_HD static void clear(T (&tau)[MAXJ]) {
#pragma unroll
    for (int k = 0; k < MAXJ; ++k) tau[k] = T(0);
  }
  RBM_HD static bool is_hinge(int) { return false; }  // generic_rnea evaluates its own trigonometry
  RBM_HD void trig(const T (&)[MAXJ], T (&)[MAXJ], T (&)[MAXJ]) const {}
  RBM_HD static bool q_matters(int) { return true; }
  RBM_HD static bool qd_matters(int) { return true; }
  RBM_HD void id(const T (&q)[MAXJ], const T (&)[MAXJ], const T (&)[MAXJ], const T (&qd)[MAXJ], const T (&qdd)[MAXJ],
                                     T (&tau)[MAXJ]) const {
    clear(tau);
    generic_rnea<T, 0>(sp, sp, nj_, q, qd, qdd, tau, nullptr, nullptr, nullptr, nullptr, nullptr);
  }
  RBM_HD void last_twists(const T (&q)[MAXJ], const T (&)[MAXJ], const T (&)[MAXJ], const T (&qd)[MAXJ], const T (&qdd)[MAXJ],
                                              T (&V)[6], T (&dV)[6]) const {
    T tau[MAXJ];
    generic_rnea<T, 0>(sp, sp, nj_, q, qd, qdd, tau, nullptr, nullptr, nullptr, V, dV);
  }
  RBM_HD const T* sensor_pose() const { return sp + GP_SENR; }
  RBM_HD void id_velocity(const T (&q)[MAXJ], const T (&)[MAXJ], const T (&)[MAXJ], const T (&qd)[MAXJ], T (&tau)[MAXJ]) const {
    T qdd0[MAXJ];
#pragma unroll
    for (int k = 0; k < MAXJ; ++k) qdd0[k] = T(0);
    // full base (twist_0, dtwist_0, tip wrench): with a moving base d tau / d qd contains V_0 x (S qd) coupling terms, so the base
    // twist must stay on; the gravity / tip-wrench parts are the same in every evaluation and cancel in the finite differences
    clear(tau);
    generic_rnea<T, 0>(sp, sp, nj_, q, qd, qdd0, tau, nullptr, nullptr, nullptr, nullptr, nullptr);
  }
  RBM_HD void id_inertia(const T (&q)[MAXJ], const T (&)[MAXJ], const T (&)[MAXJ], const T (&qdd)[MAXJ], T (&tau)[MAXJ]) const {
    T qd0[MAXJ];
#pragma unroll
    for (int k = 0; k < MAXJ; ++k) qd0[k] = T(0);
    clear(tau);
    generic_rnea<T, 0>(sp, zero, nj_, q, qd0, qdd, tau, nullptr, nullptr, nullptr, nullptr, nullptr);
  }
  RBM_HD void id_bias(const T (&q)[MAXJ], const T (&c)[MAXJ], const T (&s)[MAXJ], const T (&qd)[MAXJ], T (&tau)[MAXJ]) const {
    T qdd0[MAXJ];
#pragma unroll
    for (int k = 0; k < MAXJ; ++k) qdd0[k] = T(0);
    id(q, c, s, qd, qdd0, tau);
  }
  RBM_HD void inertia_matrix(const T (&q)[MAXJ], const T (&c)[MAXJ], const T (&s)[MAXJ], T (&M)[MAXJ][MAXJ]) const {
#pragma unroll 1
    for (int j = 0; j < nj_; ++j) {
      T e[MAXJ], col[MAXJ];
#pragma unroll
      for (int k = 0; k < MAXJ; ++k) e[k] = (k == j) ? T(1) : T(0);
      id_inertia(q, c, s, e, col);
#pragma unroll
      for (int r = 0; r < MAXJ; ++r)
#pragma unroll
        for (int k = 0; k < MAXJ; ++k)
          if (k == j) M[r][k] = col[r];
    }
  }
};

// ---- dense SPD solve (Cholesky, in place in the lower triangle; the diagonal holds 1 / L_ii) ------------
// 1 / sqrt(x): the device's rsqrt is one short sequence instead of a square root followed by a division, and it sits on the serial
// chain of the factorisation (results within 1 ulp of the two-step form; the host build of the harness uses the two-step form)
template <class T>
RBM_HD T inv_sqrt(T x) {
#ifdef __CUDA_ARCH__
  return rsqrt(x);
#else
  return T(1) / std::sqrt(x);
#endif
}
template <class T, int MAXJ>
RBM_HD void cholesky(T (&M)[MAXJ][MAXJ], int n) {
#pragma unroll
  for (int j = 0; j < MAXJ; ++j) {
    if (j < n) {
      T d = M[j][j];
#pragma unroll
      for (int k = 0; k < j; ++k) d -= M[j][k] * M[j][k];
      const T inv = inv_sqrt(d);
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
RBM_HD void chol_solve(const T (&L)[MAXJ][MAXJ], int n, T (&b)[MAXJ]) {
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

// M^-1 from the Cholesky factor, all unit right-hand sides at once: the nj substitutions are independent chains that the
// compiler interleaves, instead of nj * 2 serial triangular sweeps (the serial sweeps were the kernel's main stall source).
template <class T, int MAXJ>
RBM_HD void chol_inverse(const T (&L)[MAXJ][MAXJ], int n, T (&X)[MAXJ][MAXJ]) {
  // forward: L Y = I  (Y lower triangular)
#pragma unroll
  for (int i = 0; i < MAXJ; ++i) {
#pragma unroll
    for (int c = 0; c < MAXJ; ++c) {
      T v = (i == c) ? T(1) : T(0);
      if (c <= i) {
#pragma unroll
        for (int k = 0; k < i; ++k)
          if (k >= c) v -= L[i][k] * X[k][c];
        X[i][c] = (i < n && c < n) ? v * L[i][i] : T(0);
      } else {
        X[i][c] = T(0);
      }
    }
  }
  // backward: L^T Z = Y
#pragma unroll
  for (int i = MAXJ - 1; i >= 0; --i) {
#pragma unroll
    for (int c = 0; c < MAXJ; ++c) {
      T v = X[i][c];
#pragma unroll
      for (int k = i + 1; k < MAXJ; ++k)
        if (k < n) v -= L[k][i] * X[k][c];
      X[i][c] = (i < n && c < n) ? v * L[i][i] : T(0);
    }
  }
}
template <class T, int MAXJ>
RBM_HD void matvec(const T (&A)[MAXJ][MAXJ], int n, const T (&b)[MAXJ], T (&x)[MAXJ]) {
#pragma unroll
  for (int i = 0; i < MAXJ; ++i) {
    T v = T(0);
#pragma unroll
    for (int k = 0; k < MAXJ; ++k)
      if (k < n) v += A[i][k] * b[k];
    x[i] = v;
  }
}

// ---- the per-state algorithm ---------------------------------------------------------------------------
template <class T, class E>
RBM_HD void linearize_state(const E& ev, const T* __restrict__ q_in, const T* __restrict__ qd_in, const T* __restrict__ u_in,
                                                T dt, T eps, bool centered, T* __restrict__ A, T* __restrict__ B, T* __restrict__ qdd_out, int64_t s,
                                                int64_t ld) {
  constexpr int MJ = E::MAXJ;
  const int nj = ev.nj();
  const int ns = 2 * nj;
  T q[MJ], qd[MJ], u[MJ], c[MJ], sn[MJ], zero[MJ];
#pragma unroll
  for (int k = 0; k < MJ; ++k) {
    const bool on = k < nj;
    q[k] = on ? ld_ro(q_in + k * ld + s) : T(0);
    qd[k] = on ? ld_ro(qd_in + k * ld + s) : T(0);
    u[k] = (on && u_in) ? ld_ro(u_in + k * ld + s) : T(0);
    zero[k] = T(0);
    c[k] = T(1);
    sn[k] = T(0);
  }
  ev.trig(q, c, sn);

  // joint-space inertia matrix, column by column
  T M[MJ][MJ];
  ev.inertia_matrix(q, c, sn, M);
  // bias forces and nominal acceleration
  T h[MJ], qdd[MJ];
  ev.id_bias(q, c, sn, qd, h);
  cholesky<T, MJ>(M, nj);
  // fast path (nj fixed at compile time): explicit M^-1 once, then mat-vecs; generic path: one pair of triangular sweeps per
  // right-hand side (a fully unrolled 16 x 16 inverse would not fit in registers)
  constexpr bool kExplicitInverse = E::NJ > 0;
  T Minv[kExplicitInverse ? MJ : 1][kExplicitInverse ? MJ : 1];
  auto apply_inverse = [&](const T (&b)[MJ], T (&x)[MJ]) {
    if constexpr (kExplicitInverse) {
      matvec<T, MJ>(Minv, nj, b, x);
    } else {
#pragma unroll
      for (int k = 0; k < MJ; ++k) x[k] = b[k];
      chol_solve<T, MJ>(M, nj, x);
    }
  };
  if constexpr (kExplicitInverse) chol_inverse<T, MJ>(M, nj, Minv);
  {
    T rhs[MJ];
#pragma unroll
    for (int k = 0; k < MJ; ++k) rhs[k] = u[k] - h[k];
    apply_inverse(rhs, qdd);
  }
  if (qdd_out) {
#pragma unroll
    for (int k = 0; k < MJ; ++k)
      if (k < nj) qdd_out[k * ld + s] = qdd[k];
  }
  // B = [[dt^2 M^-1], [dt M^-1]]
  if constexpr (kExplicitInverse) {
#pragma unroll
    for (int r = 0; r < MJ; ++r) {
#pragma unroll
      for (int j = 0; j < MJ; ++j) {
        B[((int64_t)r * nj + j) * ld + s] = dt * dt * Minv[r][j];
        B[((int64_t)(nj + r) * nj + j) * ld + s] = dt * Minv[r][j];
      }
    }
  } else {
#pragma unroll 1
    for (int j = 0; j < nj; ++j) {
      T e[MJ], x[MJ];
#pragma unroll
      for (int k = 0; k < MJ; ++k) e[k] = (k == j) ? T(1) : T(0);
      apply_inverse(e, x);
#pragma unroll
      for (int r = 0; r < MJ; ++r) {
        if (r < nj) {
          B[((int64_t)r * nj + j) * ld + s] = dt * dt * x[r];
          B[((int64_t)(nj + r) * nj + j) * ld + s] = dt * x[r];
        }
      }
    }
  }
  // reference values for forward differences: ID at the nominal point (== u up to round-off) and its velocity-product part
  T tau0[MJ], vel0[MJ];
  if (!centered) {
    ev.id(q, c, sn, qd, qdd, tau0);
    ev.id_velocity(q, c, sn, qd, vel0);
  }
  const T inv_step = centered ? T(1) / (T(2) * eps) : T(1) / eps;

  // Finite-difference columns.  The +eps and -eps evaluations of a centred difference are issued back to back in one basic block
  // so that their two independent dependency chains interleave (the kernel is latency-bound on FP64 chains at 8-12 warps/SM).
  auto store_column = [&](int pass, int j, const T (&tp)[MJ], const T (&tm)[MJ]) {
    T x[MJ], dtau[MJ];
#pragma unroll
    for (int k = 0; k < MJ; ++k) dtau[k] = -(tp[k] - tm[k]) * inv_step;
    apply_inverse(dtau, x);  // column j of Q (pass 0) or V (pass 1)
    const int col = pass * nj + j;
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
  };

  // ---- position columns ------------------------------------------------------------------------------------------------
#pragma unroll 1
  for (int j = 0; j < nj; ++j) {
    if (!E::q_matters(j)) {  // d tau / d q_j is structurally zero: the column is that of the identity map
#pragma unroll
      for (int r = 0; r < MJ; ++r) {
        if (r < nj) {
          A[((int64_t)r * ns + j) * ld + s] = (r == j) ? T(1) : T(0);
          A[((int64_t)(nj + r) * ns + j) * ld + s] = T(0);
        }
      }
      continue;
    }
    T qp[MJ], qm[MJ], cp[MJ], sp_[MJ], cm[MJ], sm[MJ], tp[MJ], tm[MJ];
#pragma unroll
    for (int k = 0; k < MJ; ++k) {
      const T d = (k == j) ? eps : T(0);
      qp[k] = q[k] + d; qm[k] = q[k] - d;
      cp[k] = cm[k] = c[k];
      sp_[k] = sm[k] = sn[k];
    }
    if (E::is_hinge(j)) {  // only the perturbed joint needs new trigonometry; both signs as one group
      T ang[2] = {T(0), T(0)}, sj[2], cj[2];
#pragma unroll
      for (int k = 0; k < MJ; ++k) { ang[0] = (k == j) ? qp[k] : ang[0]; ang[1] = (k == j) ? qm[k] : ang[1]; }
      sincos_group<2, T>(ang, sj, cj);
#pragma unroll
      for (int k = 0; k < MJ; ++k) {
        cp[k] = (k == j) ? cj[0] : cp[k]; sp_[k] = (k == j) ? sj[0] : sp_[k];
        cm[k] = (k == j) ? cj[1] : cm[k]; sm[k] = (k == j) ? sj[1] : sm[k];
      }
    }
    if (centered) {
      ev.id(qp, cp, sp_, qd, qdd, tp);
      ev.id(qm, cm, sm, qd, qdd, tm);
    } else {
      ev.id(qp, cp, sp_, qd, qdd, tp);
#pragma unroll
      for (int k = 0; k < MJ; ++k) tm[k] = tau0[k];
    }
    store_column(0, j, tp, tm);
  }

  // ---- velocity columns (velocity-product term only) ---------------------------------------------------------------------
#pragma unroll 1
  for (int j = 0; j < nj; ++j) {
    if (!E::qd_matters(j)) {  // d tau / d qd_j == 0 (Galilean invariance): the column of the pure integrator
#pragma unroll
      for (int r = 0; r < MJ; ++r) {
        if (r < nj) {
          A[((int64_t)r * ns + nj + j) * ld + s] = (r == j) ? dt : T(0);
          A[((int64_t)(nj + r) * ns + nj + j) * ld + s] = (r == j) ? T(1) : T(0);
        }
      }
      continue;
    }
    T vp[MJ], vm[MJ], tp[MJ], tm[MJ];
#pragma unroll
    for (int k = 0; k < MJ; ++k) {
      const T d = (k == j) ? eps : T(0);
      vp[k] = qd[k] + d; vm[k] = qd[k] - d;
    }
    if (centered) {
      ev.id_velocity(q, c, sn, vp, tp);
      ev.id_velocity(q, c, sn, vm, tm);
    } else {
      ev.id_velocity(q, c, sn, vp, tp);
#pragma unroll
      for (int k = 0; k < MJ; ++k) tm[k] = vel0[k];
    }
    store_column(1, j, tp, tm);
  }
}

// ---- forward dynamics / one transition step (the map the linearisation differentiates) -------------------------------------
//   qdd = M(q)^-1 (u - h(q, qd));   dt > 0:  qd+ = qd + dt qdd,  q+ = q + dt qd+   (semi-implicit Euler, MuJoCo's default)
template <class T, class E>
RBM_HD void forward_dynamics_state(const E& ev, const T* __restrict__ q_in, const T* __restrict__ qd_in, const T* __restrict__ u_in,
                                                       T dt, T* __restrict__ qdd_out, T* __restrict__ q_next, T* __restrict__ qd_next, int64_t s,
                                                       int64_t ld) {
  constexpr int MJ = E::MAXJ;
  const int nj = ev.nj();
  T q[MJ], qd[MJ], u[MJ], c[MJ], sn[MJ], zero[MJ];
#pragma unroll
  for (int k = 0; k < MJ; ++k) {
    const bool on = k < nj;
    q[k] = on ? ld_ro(q_in + k * ld + s) : T(0);
    qd[k] = on ? ld_ro(qd_in + k * ld + s) : T(0);
    u[k] = (on && u_in) ? ld_ro(u_in + k * ld + s) : T(0);
    zero[k] = T(0);
    c[k] = T(1);
    sn[k] = T(0);
  }
  ev.trig(q, c, sn);
  T M[MJ][MJ];
  ev.inertia_matrix(q, c, sn, M);
  T h[MJ], qdd[MJ];
  ev.id_bias(q, c, sn, qd, h);
  cholesky<T, MJ>(M, nj);
#pragma unroll
  for (int k = 0; k < MJ; ++k) qdd[k] = u[k] - h[k];
  chol_solve<T, MJ>(M, nj, qdd);
#pragma unroll
  for (int k = 0; k < MJ; ++k) {
    if (k < nj) {
      if (qdd_out) qdd_out[k * ld + s] = qdd[k];
      if (q_next) {
        const T v = qd[k] + dt * qdd[k];
        qd_next[k * ld + s] = v;
        q_next[k * ld + s] = q[k] + dt * v;
      }
    }
  }
}

// ---- closed-loop rollout (reference core/simulate.py:185-270), one environment per thread ---------------------------------------
// Every step of the reference's main loop, with MuJoCo's mj_step replaced by the transition above:
//   C(step): tgt = plan(step) (:187), tgt_ctrl = ID(tgt) (:188); act = (qpos, qvel, qacc) where qacc still belongs to the PREVIOUS
//            forward pass (:191-194); on frame steps (`frame_count <= time * fps`, :196) log act, the sensor-frame twists of act
//            (:202-209) and the F/T reading left by the previous forward pass (:218-221);
//            res = [(tgt_q - qpos) / div, tgt_qd - qvel] (mj_differentiatePos with m.nu in the dt slot, :257-265);
//            ctrl = tgt_ctrl - K res (:268)
//   A      : forward pass at (qpos, qvel, ctrl): qacc = M^-1 (ctrl - h); F/T sensor = Newton-Euler wrench of the sensed subtree in the
//            sensor frame = Y(V_s, dV_s) phi  (cfrc_int of body "target/" in the site frame; evaluated only when the next step logs)
//   B      : qvel += dt qacc; qpos += dt qvel; time += dt                                  (:270)
// started by one forward pass at the initial state with ctrl = 0 (what the controller's linearisation leaves in MjData).
// Frame record: [act (3 nj) | V_s (6) | dV_s (6) | wrench (6)], stored value-major [frame][value][env] so that every store is coalesced.
template <class T, class E>
RBM_HD void closed_loop_env(const E& ev, const PlanArg<T>& pl, const T* __restrict__ K, const T* __restrict__ phi_in, T dt, T fps,
                                                T div, int n_steps, int max_frames, const T* __restrict__ q0, const T* __restrict__ qd0,
                                                T* __restrict__ frames, int* __restrict__ frame_steps, int* __restrict__ n_frames,
                                                T* __restrict__ final_state, int64_t s, int64_t ld) {
  constexpr int MJ = E::MAXJ;
  const int nj = ev.nj();
  const int fv = 3 * nj + 18;
  T q[MJ], qd[MJ], qacc[MJ], u[MJ], c[MJ], sn[MJ], zero[MJ], wrench[6], phi[10];
#pragma unroll
  for (int k = 0; k < MJ; ++k) {
    const bool on = k < nj;
    q[k] = on ? ld_ro(q0 + k * ld + s) : T(0);
    qd[k] = (on && qd0) ? ld_ro(qd0 + k * ld + s) : T(0);
    qacc[k] = u[k] = zero[k] = sn[k] = T(0);
    c[k] = T(1);
  }
#pragma unroll
  for (int k = 0; k < 10; ++k) phi[k] = ld_ro(phi_in + k);
#pragma unroll
  for (int k = 0; k < 6; ++k) wrench[k] = T(0);
  const T* senp = ev.sensor_pose();
  T time = T(0);
  int frame_count = 0;
  bool integrate = false;
#pragma unroll 1
  for (int it = 0; it <= n_steps; ++it) {
    // ---- A: forward pass at (q, qd, u)
    ev.trig(q, c, sn);
    {
      T M[MJ][MJ];
      ev.inertia_matrix(q, c, sn, M);
      T h[MJ];
      ev.id_bias(q, c, sn, qd, h);
      cholesky<T, MJ>(M, nj);
#pragma unroll
      for (int k = 0; k < MJ; ++k) qacc[k] = u[k] - h[k];
      chol_solve<T, MJ>(M, nj, qacc);
    }
    const T t_next = integrate ? time + dt : time;
    if (it < n_steps && (T)frame_count <= t_next * fps) {  // the F/T reading the next frame will log
      T V[6], dV[6], Vs[6], dVs[6], top[3][4], bot[3][9];
      ev.last_twists(q, c, sn, qd, qacc, V, dV);
      sensor_twists(senp, senp + 9, V, dV, Vs, dVs);
      regressor_blocks(Vs, dVs, top, bot);
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        T f = T(0), m = T(0);
#pragma unroll
        for (int k = 0; k < 4; ++k) f += top[r][k] * phi[k];
#pragma unroll
        for (int k = 0; k < 9; ++k) m += bot[r][k] * phi[1 + k];
        wrench[r] = f;
        wrench[3 + r] = m;
      }
    }
    // ---- B: integrate
    if (integrate) {
#pragma unroll
      for (int k = 0; k < MJ; ++k) {
        qd[k] = qd[k] + dt * qacc[k];
        q[k] = q[k] + dt * qd[k];
      }
      time = t_next;
    }
    integrate = true;
    if (it == n_steps) break;
    // ---- C: plan, log, control law
    T sp_, sv_, sa_, tq[MJ], tqd[MJ], tqdd[MJ], tctrl[MJ], tc[MJ], ts[MJ];
    plan_profile(pl, (int64_t)it, sp_, sv_, sa_);
#pragma unroll
    for (int k = 0; k < MJ; ++k) {
      tq[k] = pl.disp[k] * sp_ + pl.offset[k];
      tqd[k] = pl.disp[k] * sv_;
      tqdd[k] = pl.disp[k] * sa_;
      tc[k] = T(1);
      ts[k] = T(0);
    }
    ev.trig(tq, tc, ts);
    ev.id(tq, tc, ts, tqd, tqdd, tctrl);
    if ((T)frame_count <= time * fps) {
      if (frame_count < max_frames) {
        T V[6], dV[6], Vs[6], dVs[6];
        ev.trig(q, c, sn);
        ev.last_twists(q, c, sn, qd, qacc, V, dV);
        sensor_twists(senp, senp + 9, V, dV, Vs, dVs);
        T* dst = frames + ((int64_t)frame_count * fv) * ld + s;
#pragma unroll
        for (int k = 0; k < MJ; ++k) {
          if (k < nj) {
            dst[(int64_t)k * ld] = q[k];
            dst[(int64_t)(nj + k) * ld] = qd[k];
            dst[(int64_t)(2 * nj + k) * ld] = qacc[k];
          }
        }
#pragma unroll
        for (int k = 0; k < 6; ++k) {
          dst[(int64_t)(3 * nj + k) * ld] = Vs[k];
          dst[(int64_t)(3 * nj + 6 + k) * ld] = dVs[k];
          dst[(int64_t)(3 * nj + 12 + k) * ld] = wrench[k];
        }
        if (s == 0 && frame_steps) frame_steps[frame_count] = it;
      }
      ++frame_count;
    }
#pragma unroll
    for (int r = 0; r < MJ; ++r) {
      if (r < nj) {
        T v = tctrl[r];
        for (int k = 0; k < nj; ++k) {
          v -= ld_ro(K + r * 2 * nj + k) * ((tq[k] - q[k]) / div);
          v -= ld_ro(K + r * 2 * nj + nj + k) * (tqd[k] - qd[k]);
        }
        u[r] = v;
      }
    }
  }
  if (s == 0 && n_frames) *n_frames = frame_count;
  if (final_state) {
#pragma unroll
    for (int k = 0; k < MJ; ++k) {
      if (k < nj) {
        final_state[(int64_t)k * ld + s] = q[k];
        final_state[(int64_t)(nj + k) * ld + s] = qd[k];
        final_state[(int64_t)(2 * nj + k) * ld + s] = qacc[k];
      }
    }
  }
}

}  // namespace rbm
