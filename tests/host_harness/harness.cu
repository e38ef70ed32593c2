// TEST INFRASTRUCTURE ONLY (never part of librbm_b200.so, never reachable from the package).
//
// The device arithmetic of the kernels lives in __host__ __device__ headers (csrc/rbm_typed.cuh, rbm_trig.cuh, rbm_rnea.cuh,
// rbm_dynamics.cuh, rbm_setup.cuh).
// This translation unit instantiates the SAME per-sample functions for the host and exposes them through a tiny C interface,
// so that the `-m "not gpu"` suite can check the kernels' mathematics -- the structure-specialised typed recursion, the generic
// recursion, the regressor blocks, the reduced (inertia-only / velocity-only) evaluations -- against the reference-generated
// golden vectors on a machine without a GPU.  It proves nothing about launches, memory traffic or TMA pipelines; the -m gpu
// suite does that through the real C ABI.
#include <cstdint>
#include <cstring>
#include <type_traits>

#include "rbm_dynamics.cuh"
#include "rbm_gram.cuh"
#include "rbm_rnea.cuh"
#include "rbm_setup.cuh"

using namespace rbm;

namespace {
template <class T, class D, bool VEL, bool GRAV, bool ACC>
void run_fast(const FastParams<T>& P, const T* traj, T* tau, T* V, T* dV, int64_t n) {
  for (int64_t s = 0; s < n; ++s) {
    T q[6], qd[6], qdd[6], c[6], sn[6];
    for (int j = 0; j < 6; ++j) { q[j] = traj[s * 18 + j]; qd[j] = traj[s * 18 + 6 + j]; qdd[j] = traj[s * 18 + 12 + j]; }
    fast_sincos<T, D>(q, c, sn);
    FastResult<T> r;
    fast_rnea_core<T, D, true, VEL, GRAV, ACC>(P, P.g, q, c, sn, qd, qdd, r);
    for (int j = 0; j < 6; ++j) tau[s * 6 + j] = r.tau[j];
    if (V) {
      for (int k = 0; k < 3; ++k) { V[s * 6 + k] = r.v[k]; V[s * 6 + 3 + k] = r.w[k]; dV[s * 6 + k] = r.a[k]; dV[s * 6 + 3 + k] = r.l[k]; }
    }
  }
}

template <class T>
FastParams<T> convert(const double* fp) {
  FastParams<T> P;
  T* dst = reinterpret_cast<T*>(&P);
  for (size_t i = 0; i < sizeof(FastParams<double>) / sizeof(double); ++i) dst[i] = (T)fp[i];
  return P;
}

// mode: 0 = full ID, 1 = inertia-only (VEL off, GRAV off), 2 = velocity-product only (ACC off, GRAV off)
template <class T>
int fast_dispatch(int path, int mode, const double* fp, const T* traj, T* tau, T* V, T* dV, int64_t n) {
  const FastParams<T> P = convert<T>(fp);
#define RUN(D)                                                                  \
  if (mode == 0) run_fast<T, D, true, true, true>(P, traj, tau, V, dV, n);      \
  else if (mode == 1) run_fast<T, D, false, false, true>(P, traj, tau, V, dV, n); \
  else if (mode == 2) run_fast<T, D, true, false, false>(P, traj, tau, V, dV, n); \
  else return -1;
  if (path == PATH_SEQ_ISO) { RUN(SeqIso) }
  else if (path == PATH_SEQ_RIGID) { RUN(SeqRigid) }
  else return -1;
#undef RUN
  return 0;
}
}  // namespace

extern "C" {

int h_fast_rnea_f64(int path, int mode, const double* fast_params, const double* traj, double* tau, double* V, double* dV, int64_t n) {
  return fast_dispatch<double>(path, mode, fast_params, traj, tau, V, dV, n);
}
int h_fast_rnea_f32(int path, int mode, const double* fast_params, const float* traj, float* tau, float* V, float* dV, int64_t n) {
  return fast_dispatch<float>(path, mode, fast_params, traj, tau, V, dV, n);
}

// generic recursion; gp = packed generic parameters (rbm_model_analyze), unrolled != 0 uses the compile-time nj = 6 instance
int h_generic_rnea_f64(const double* gp, int nj, int unrolled, const double* traj, double* tau, double* poses, double* twists, double* dtwists,
                       double* Vlast, double* dVlast, int64_t n) {
  if (nj < 1 || nj > RBM_MAX_JOINTS || (unrolled && nj != 6)) return -1;
  for (int64_t s = 0; s < n; ++s) {
    double q[RBM_MAX_JOINTS], qd[RBM_MAX_JOINTS], qdd[RBM_MAX_JOINTS], t[RBM_MAX_JOINTS];
    for (int j = 0; j < nj; ++j) { q[j] = traj[s * 3 * nj + j]; qd[j] = traj[s * 3 * nj + nj + j]; qdd[j] = traj[s * 3 * nj + 2 * nj + j]; }
    double* P = poses ? poses + s * nj * 12 : nullptr;
    double* TW = twists ? twists + s * (nj + 1) * 6 : nullptr;
    double* DTW = dtwists ? dtwists + s * (nj + 1) * 6 : nullptr;
    double* VL = Vlast ? Vlast + s * 6 : nullptr;
    double* DVL = dVlast ? dVlast + s * 6 : nullptr;
    if (unrolled) generic_rnea<double, 6>(gp, gp, nj, q, qd, qdd, t, P, TW, DTW, VL, DVL);
    else generic_rnea<double, 0>(gp, gp, nj, q, qd, qdd, t, P, TW, DTW, VL, DVL);
    for (int j = 0; j < nj; ++j) tau[s * nj + j] = t[j];
  }
  return 0;
}

int h_generic_rnea_f32(const double* gp, int nj, const float* traj, float* tau, int64_t n) {
  if (nj < 1 || nj > RBM_MAX_JOINTS) return -1;
  const int np = GP_HEAD + GJ_STRIDE * nj;
  float gpf[GP_HEAD + GJ_STRIDE * RBM_MAX_JOINTS];
  for (int i = 0; i < np; ++i) gpf[i] = (float)gp[i];
  for (int64_t s = 0; s < n; ++s) {
    float q[RBM_MAX_JOINTS], qd[RBM_MAX_JOINTS], qdd[RBM_MAX_JOINTS], t[RBM_MAX_JOINTS];
    for (int j = 0; j < nj; ++j) { q[j] = traj[s * 3 * nj + j]; qd[j] = traj[s * 3 * nj + nj + j]; qdd[j] = traj[s * 3 * nj + 2 * nj + j]; }
    generic_rnea<float, 0>(gpf, gpf, nj, q, qd, qdd, t, nullptr, nullptr, nullptr, nullptr, nullptr);
    for (int j = 0; j < nj; ++j) tau[s * nj + j] = t[j];
  }
  return 0;
}

// sensor-frame twists + regressor rows: V, dV [n][6] -> Y [n][6][10]
int h_sensor_regressor_f64(const double* pose_Rt, const double* V, const double* dV, double* Vs, double* dVs, double* Y, int64_t n) {
  for (int64_t s = 0; s < n; ++s) {
    double vs[6], dvs[6], top[3][4], bot[3][9];
    sensor_twists(pose_Rt, pose_Rt + 9, V + 6 * s, dV + 6 * s, vs, dvs);
    regressor_blocks(vs, dvs, top, bot);
    std::memcpy(Vs + 6 * s, vs, sizeof(vs));
    std::memcpy(dVs + 6 * s, dvs, sizeof(dvs));
    double* y = Y + 60 * s;
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 10; ++c) {
        y[r * 10 + c] = c < 4 ? top[r][c] : 0.0;
        y[(3 + r) * 10 + c] = c == 0 ? 0.0 : bot[r][c - 1];
      }
  }
  return 0;
}

}  // extern "C"

// ---- rbm_dynamics.cuh: LQR linearisation, forward dynamics / step, closed-loop rollout (same SoA layouts as the kernels) ----------
namespace {
template <class F>
int with_evaluator(int path, const double* fast_params, const double* gp, int nj, F&& body) {
  if (path == PATH_SEQ_ISO || path == PATH_SEQ_RIGID) {
    const FastParams<double> P = convert<double>(fast_params);
    if (path == PATH_SEQ_ISO) { FastEval<double, SeqIso> ev{P}; body(ev); }
    else { FastEval<double, SeqRigid> ev{P}; body(ev); }
    return 0;
  }
  if (path != PATH_GENERIC || nj < 1 || nj > RBM_MAX_JOINTS) return -1;
  double zero[18] = {0};
  GenericEval<double> ev{gp, zero, nj};
  body(ev);
  return 0;
}
}  // namespace

extern "C" {

int h_linearize_f64(int path, const double* fast_params, const double* gp, int nj, const double* q, const double* qd, const double* u, double dt, double eps,
                    int centered, double* A, double* B, double* qdd, int64_t n) {
  return with_evaluator(path, fast_params, gp, nj, [&](auto& ev) {
    for (int64_t s = 0; s < n; ++s) linearize_state<double>(ev, q, qd, u, dt, eps, centered != 0, A, B, qdd, s, n);
  });
}

int h_forward_dynamics_f64(int path, const double* fast_params, const double* gp, int nj, const double* q, const double* qd, const double* u, double dt,
                           double* qdd, double* q_next, double* qd_next, int64_t n) {
  return with_evaluator(path, fast_params, gp, nj, [&](auto& ev) {
    for (int64_t s = 0; s < n; ++s) forward_dynamics_state<double>(ev, q, qd, u, dt, qdd, q_next, qd_next, s, n);
  });
}

int h_closed_loop_f64(int path, const double* fast_params, const double* gp, int nj, const double* coeffs, const double* disp, const double* offset,
                      double plan_timestep, double step0, int n_steps, const double* K, const double* phi, double dt, double fps, double div,
                      const double* q0, const double* qd0, double* frames, int max_frames, int* frame_steps, int* n_frames, double* final_state, int64_t n) {
  PlanArg<double> pl;
  for (int k = 0; k < 6; ++k) pl.coeffs[k] = coeffs[k];
  for (int k = 0; k < RBM_MAX_JOINTS; ++k) { pl.disp[k] = k < nj ? disp[k] : 0.0; pl.offset[k] = k < nj ? offset[k] : 0.0; }
  pl.inv_dt = 1.0 / plan_timestep;
  pl.inv_dt2 = 1.0 / (plan_timestep * plan_timestep);
  pl.step0 = step0;
  pl.stride = 1.0;
  return with_evaluator(path, fast_params, gp, nj, [&](auto& ev) {
    for (int64_t s = 0; s < n; ++s)
      closed_loop_env<double>(ev, pl, K, phi, dt, fps, div, n_steps, max_frames, q0, qd0, frames, frame_steps, n_frames, final_state, s, n);
  });
}

}  // extern "C"

// ---- rbm_gram.cuh: the fused regressor + Gram accumulation of k_regressor_gram*, one sample after the other ------------------------
extern "C" int h_regressor_gram_f64(int path, const double* fast_params, const double* gp, int nj, const double* q, const double* qd, const double* qdd,
                                    const double* f, double* pack, int64_t n) {
  double acc[kAcc] = {0};
  int rc = with_evaluator(path, fast_params, gp, nj, [&](auto& ev) {
    using E = std::decay_t<decltype(ev)>;
    constexpr int MJ = E::MAXJ;
    const double* senp = ev.sensor_pose();
    for (int64_t s = 0; s < n; ++s) {
      double rq[MJ] = {0}, rqd[MJ] = {0}, rqdd[MJ] = {0}, c[MJ], sn[MJ], V[6], dV[6], Vs[6], dVs[6], fs[6], top[3][4], bot[3][9];
      for (int k = 0; k < nj; ++k) { rq[k] = q[k * n + s]; rqd[k] = qd[k * n + s]; rqdd[k] = qdd[k * n + s]; }
      for (int k = 0; k < MJ; ++k) { c[k] = 1.0; sn[k] = 0.0; }
      for (int k = 0; k < 6; ++k) fs[k] = f[k * n + s];
      ev.trig(rq, c, sn);
      ev.last_twists(rq, c, sn, rqd, rqdd, V, dV);
      sensor_twists(senp, senp + 9, V, dV, Vs, dVs);
      regressor_blocks(Vs, dVs, top, bot);
      gram_accumulate(acc, top, bot, fs);
    }
  });
  if (rc != 0) return rc;
  for (int t = 0; t < 111; ++t) pack[t] = gram_pack_entry(acc, t);
  pack[111] = (double)n;
  return 0;
}


// ---- rbm_setup.cuh: the small frame-algebra helpers, one item per loop iteration (the kernels run one item per thread) ----------------
extern "C" {
int h_transfer_simat_f64(const double* poses, const double* simats, double* out, int64_t n, int mode) {
  for (int64_t s = 0; s < n; ++s) transfer_simat_item(poses + 12 * s, simats + 36 * s, out + 36 * s, mode);
  return 0;
}
int h_transfer_imat_f64(const double* poses, const double* imats, const double* mass, double* out, int64_t n) {
  for (int64_t s = 0; s < n; ++s) transfer_imat_item(poses + 12 * s, imats + 9 * s, mass[s], out + 9 * s);
  return 0;
}
int h_spatial_inertia_f64(const double* mass, const double* diag, double* out, int64_t n) {
  for (int64_t s = 0; s < n; ++s) spatial_inertia_item(mass[s], diag + 3 * s, out + 36 * s);
  return 0;
}
int h_compose_f64(const double* trans, const double* rot, int rot_len, double* out, int32_t* status, int64_t n) {
  for (int64_t s = 0; s < n; ++s) status[s] = compose_item(trans + 3 * s, rot + (int64_t)rot_len * s, rot_len, out + 12 * s);
  return 0;
}
int h_point_motion_f64(const double* tw, const double* dtw, const double* pts, double* linvel, double* linacc, int64_t n) {
  for (int64_t s = 0; s < n; ++s)
    point_motion_item(tw + 6 * s, dtw ? dtw + 6 * s : nullptr, pts + 3 * s, linvel ? linvel + 3 * s : nullptr, linacc ? linacc + 3 * s : nullptr);
  return 0;
}
int h_regressor_rows_f64(const double* V, const double* dV, double* Y, int64_t n) {
  for (int64_t s = 0; s < n; ++s) {
    double top[3][4], bot[3][9];
    regressor_blocks(V + 6 * s, dV + 6 * s, top, bot);
    double* y = Y + 60 * s;
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 10; ++c) {
        y[r * 10 + c] = c < 4 ? top[r][c] : 0.0;
        y[(3 + r) * 10 + c] = c == 0 ? 0.0 : bot[r][c - 1];
      }
  }
  return 0;
}
}  // extern "C"
