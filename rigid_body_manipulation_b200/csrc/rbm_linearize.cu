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
#include "rbm_dynamics.cuh"
#include "rbm_rnea.cuh"

namespace rbm {

constexpr int kLinBlock = 128;
// Occupancy experiments (tools/kbench `lin`, 2^20 states, 200 launches, G states/s): 2 blocks of 128 threads per SM at 254 registers
// (this setting) 2.76; 3 blocks at 168 registers (~300 B of spills) 2.69; M^-1 parked in shared memory between the column loops
// 2.75 with 2 blocks, 2.52 with 3.  More resident warps do not pay for the spills: the setting stays at 2.
// Round 2: a lane-PAIR version (two lanes share a state: +eps / -eps evaluations, half the M^-1 rows each, 168 or 128 registers, 12 or 16
// warps per SM) was built, passed the parity suite and LOST: 2.10 / 2.14 against 2.80 G states/s (ncu: +17 % instructions, +14 % DRAM
// write traffic from half-warp stores, FP64 pipe 46 % instead of 55 %).  Kept as experiments/k_linearize_pair.cu.inc.
// Also round 2: a persistent variant (one two-stage TMA pipeline of 32-state input sub-tiles per warp, no CTA barrier) LOST 9 %
// (2.76 against 3.03; persistent with plain loads 2.69): the hardware block scheduler's dynamic placement beats a static tile loop
// here.  What did help is an L2 prefetch of a later block's 18 input rows: +2 ... 3 % at a distance of 64 ... 148 blocks (3.06 at 64,
// 3.04 at 100 / 148, 3.00 at 200 / 296, 2.97 without; kbench, 200 launches).
#ifndef RBM_LIN_MINB
#define RBM_LIN_MINB 2
#endif
constexpr int kLinPrefetchBlocks = 64;

// evaluators, Cholesky solve, linearize_state, forward_dynamics_state, closed_loop_env: rbm_dynamics.cuh (shared with the host harness)


template <class T, class D>
__global__ void __launch_bounds__(kLinBlock) k_forward_dynamics_fast(const __grid_constant__ FastParams<T> P, const T* __restrict__ q, const T* __restrict__ qd,
                                                                     const T* __restrict__ u, T dt, T* __restrict__ qdd, T* __restrict__ q_next,
                                                                     T* __restrict__ qd_next, int64_t n, int64_t ld) {
  const int64_t s = (int64_t)blockIdx.x * kLinBlock + threadIdx.x;
  if (s >= n) return;
  FastEval<T, D> ev{P};
  forward_dynamics_state<T>(ev, q, qd, u, dt, qdd, q_next, qd_next, s, ld);
}

template <class T>
__global__ void __launch_bounds__(kLinBlock) k_forward_dynamics_generic(const T* __restrict__ gp, int nj, int nparams, const T* __restrict__ q,
                                                                        const T* __restrict__ qd, const T* __restrict__ u, T dt, T* __restrict__ qdd,
                                                                        T* __restrict__ q_next, T* __restrict__ qd_next, int64_t n, int64_t ld) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* sp = reinterpret_cast<T*>(smem_raw);
  T* zero = sp + nparams;
  for (int i = threadIdx.x; i < nparams; i += blockDim.x) sp[i] = gp[i];
  for (int i = threadIdx.x; i < 18; i += blockDim.x) zero[i] = T(0);
  __syncthreads();
  const int64_t s = (int64_t)blockIdx.x * kLinBlock + threadIdx.x;
  if (s >= n) return;
  GenericEval<T> ev{sp, zero, nj};
  forward_dynamics_state<T>(ev, q, qd, u, dt, qdd, q_next, qd_next, s, ld);
}

template <class T, class D>
__global__ void __launch_bounds__(kLinBlock, RBM_LIN_MINB) k_linearize_fast(const __grid_constant__ FastParams<T> P, const T* __restrict__ q, const T* __restrict__ qd,
                                                              const T* __restrict__ u, T dt, T eps, int centered, T* __restrict__ A, T* __restrict__ B,
                                                              T* __restrict__ qdd_out, int64_t n, int64_t ld) {
  const int64_t s = (int64_t)blockIdx.x * kLinBlock + threadIdx.x;
  if (s >= n) return;
  // One state is ~6 000 instructions behind 18 loads and two warps per scheduler cannot cover their DRAM latency (15 % of the stall
  // samples sat on the first use of q): pull the inputs of the block that starts kLinPrefetchBlocks later into L2 now.
  constexpr int64_t ahead = (int64_t)kLinPrefetchBlocks * kLinBlock;
  if (s + ahead < n) {
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      asm volatile("prefetch.global.L2 [%0];" ::"l"(q + k * ld + s + ahead));
      asm volatile("prefetch.global.L2 [%0];" ::"l"(qd + k * ld + s + ahead));
      if (u) asm volatile("prefetch.global.L2 [%0];" ::"l"(u + k * ld + s + ahead));
    }
  }
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



constexpr int kLoopBlock = 64;  // few environments per block: a rollout is one long serial chain per thread, so spread envs over the SMs

template <class T, class D>
__global__ void __launch_bounds__(kLoopBlock) k_closed_loop_fast(const __grid_constant__ FastParams<T> P, const __grid_constant__ PlanArg<T> pl,
                                                                 const T* __restrict__ K, const T* __restrict__ phi, T dt, T fps, T div, int n_steps,
                                                                 int max_frames, const T* __restrict__ q0, const T* __restrict__ qd0, T* __restrict__ frames,
                                                                 int* __restrict__ frame_steps, int* __restrict__ n_frames, T* __restrict__ final_state,
                                                                 int64_t n, int64_t ld) {
  const int64_t s = (int64_t)blockIdx.x * kLoopBlock + threadIdx.x;
  if (s >= n) return;
  FastEval<T, D> ev{P};
  closed_loop_env<T>(ev, pl, K, phi, dt, fps, div, n_steps, max_frames, q0, qd0, frames, frame_steps, n_frames, final_state, s, ld);
}

template <class T>
__global__ void __launch_bounds__(kLoopBlock) k_closed_loop_generic(const T* __restrict__ gp, int nj, int nparams, const __grid_constant__ PlanArg<T> pl,
                                                                    const T* __restrict__ K, const T* __restrict__ phi, T dt, T fps, T div, int n_steps,
                                                                    int max_frames, const T* __restrict__ q0, const T* __restrict__ qd0,
                                                                    T* __restrict__ frames, int* __restrict__ frame_steps, int* __restrict__ n_frames,
                                                                    T* __restrict__ final_state, int64_t n, int64_t ld) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* sp = reinterpret_cast<T*>(smem_raw);
  T* zero = sp + nparams;
  for (int i = threadIdx.x; i < nparams; i += blockDim.x) sp[i] = gp[i];
  for (int i = threadIdx.x; i < 18; i += blockDim.x) zero[i] = T(0);
  __syncthreads();
  const int64_t s = (int64_t)blockIdx.x * kLoopBlock + threadIdx.x;
  if (s >= n) return;
  GenericEval<T> ev{sp, zero, nj};
  closed_loop_env<T>(ev, pl, K, phi, dt, fps, div, n_steps, max_frames, q0, qd0, frames, frame_steps, n_frames, final_state, s, ld);
}

int launch_closed_loop(const rbm_model* m, const double* coeffs, const double* disp, const double* offset, double plan_timestep, double step0,
                       int n_steps, const double* K, const double* phi, double dt, double fps, double div, const double* q0, const double* qd0,
                       double* frames, int max_frames, int* frame_steps, int* n_frames, double* final_state, int64_t n, int64_t ld, cudaStream_t st) {
  if (n == 0) return RBM_OK;
  using T = double;
  PlanArg<T> pl;
  for (int k = 0; k < 6; ++k) pl.coeffs[k] = coeffs[k];
  for (int k = 0; k < RBM_MAX_JOINTS; ++k) {
    pl.disp[k] = k < m->nj ? disp[k] : 0.0;
    pl.offset[k] = k < m->nj ? offset[k] : 0.0;
  }
  pl.inv_dt = 1.0 / plan_timestep;
  pl.inv_dt2 = 1.0 / (plan_timestep * plan_timestep);
  pl.step0 = step0;
  pl.stride = 1.0;
  const unsigned grid = (unsigned)((n + kLoopBlock - 1) / kLoopBlock);
  if (m->path == PATH_SEQ_ISO) {
    k_closed_loop_fast<T, SeqIso><<<grid, kLoopBlock, 0, st>>>(ModelView<T>::fast(m), pl, K, phi, dt, fps, div, n_steps, max_frames, q0, qd0, frames,
                                                              frame_steps, n_frames, final_state, n, ld);
  } else if (m->path == PATH_SEQ_RIGID) {
    k_closed_loop_fast<T, SeqRigid><<<grid, kLoopBlock, 0, st>>>(ModelView<T>::fast(m), pl, K, phi, dt, fps, div, n_steps, max_frames, q0, qd0, frames,
                                                                frame_steps, n_frames, final_state, n, ld);
  } else {
    const int np = generic_param_count(m->nj);
    k_closed_loop_generic<T><<<grid, kLoopBlock, sizeof(T) * (np + 18), st>>>(ModelView<T>::generic(m), m->nj, np, pl, K, phi, dt, fps, div, n_steps,
                                                                             max_frames, q0, qd0, frames, frame_steps, n_frames, final_state, n, ld);
  }
  RBM_CUDA_TRY(cudaGetLastError());
  return RBM_OK;
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

template <class T>
int launch_forward_dynamics(const rbm_model* m, const T* q, const T* qd, const T* u, double dt, T* qdd, T* q_next, T* qd_next, int64_t n, int64_t ld,
                            cudaStream_t st) {
  if (n == 0) return RBM_OK;
  const unsigned grid = (unsigned)((n + kLinBlock - 1) / kLinBlock);
  if (m->path == PATH_SEQ_ISO) {
    k_forward_dynamics_fast<T, SeqIso><<<grid, kLinBlock, 0, st>>>(ModelView<T>::fast(m), q, qd, u, (T)dt, qdd, q_next, qd_next, n, ld);
  } else if (m->path == PATH_SEQ_RIGID) {
    k_forward_dynamics_fast<T, SeqRigid><<<grid, kLinBlock, 0, st>>>(ModelView<T>::fast(m), q, qd, u, (T)dt, qdd, q_next, qd_next, n, ld);
  } else {
    const int np = generic_param_count(m->nj);
    k_forward_dynamics_generic<T><<<grid, kLinBlock, sizeof(T) * (np + 18), st>>>(ModelView<T>::generic(m), m->nj, np, q, qd, u, (T)dt, qdd, q_next, qd_next,
                                                                                n, ld);
  }
  RBM_CUDA_TRY(cudaGetLastError());
  return RBM_OK;
}

template int launch_forward_dynamics<double>(const rbm_model*, const double*, const double*, const double*, double, double*, double*, double*, int64_t,
                                             int64_t, cudaStream_t);
template int launch_linearize<double>(const rbm_model*, const double*, const double*, const double*, double, double, int, double*, double*, double*, int64_t,
                                      int64_t, cudaStream_t);

}  // namespace rbm
