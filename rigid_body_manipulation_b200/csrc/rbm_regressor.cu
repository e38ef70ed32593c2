// Sensor-frame regressor and its Gram accumulation (inertial-parameter identification).
//
// Replaces, batched and fused:  core/simulate.py:202-209 (twists into the F/T sensor frame),
// dynamics/dynamics.py:215-249 (get_regressor_matrix, 6x10 per sample) and the stacking +
// np.linalg.lstsq of loggers/loggers.py:127-129, which is re-expressed as normal equations: the kernel
// accumulates  [Y f]^T [Y f]  (Y^T Y 10x10, Y^T f 10, f^T f) and never writes Y to HBM.
//
// Gram kernel: persistent grid, one sample per thread per iteration; each thread keeps the 70 distinct
// non-zero entries of the two diagonal blocks in registers, blocks reduce with warp shuffles + shared
// memory, and a second single-block kernel sums the per-block partials in a FIXED order (deterministic
// result for a given grid).  HBM traffic per sample: q, qd, qdd, f = 24 scalars read, nothing written.
#include <atomic>
#include <cstdlib>

#include "rbm_async.cuh"
#include "rbm_internal.h"
#include "rbm_gram.cuh"
#include "rbm_rnea.cuh"

namespace rbm {

constexpr int kRowsBlock = 128;
constexpr int kGramBlock = 256;

// ---- last-link twist for one sample, any kernel path ------------------------------------------------
template <class T, int PATH>
__device__ __forceinline__ void last_link_twists(const FastParams<T>& P, const T* sp, int nj, const T* __restrict__ q, const T* __restrict__ qd,
                                                 const T* __restrict__ qdd, int64_t s, int64_t ld, T* V, T* dV) {
  if constexpr (PATH == PATH_GENERIC) {
    T rq[RBM_MAX_JOINTS], rqd[RBM_MAX_JOINTS], rqdd[RBM_MAX_JOINTS];
    for (int j = 0; j < nj; ++j) { rq[j] = __ldg(q + j * ld + s); rqd[j] = __ldg(qd + j * ld + s); rqdd[j] = __ldg(qdd + j * ld + s); }
    generic_rnea<T, 0>(sp, sp, nj, rq, rqd, rqdd, (T*)nullptr, nullptr, nullptr, nullptr, V, dV);
  } else {
    T rq[6], rqd[6], rqdd[6];
#pragma unroll
    for (int j = 0; j < 6; ++j) { rq[j] = __ldg(q + j * ld + s); rqd[j] = __ldg(qd + j * ld + s); rqdd[j] = __ldg(qdd + j * ld + s); }
    FastResult<T> r;
    if constexpr (PATH == PATH_SEQ_ISO) fast_rnea<T, SeqIso, false>(P, rq, rqd, rqdd, r);
    else fast_rnea<T, SeqRigid, false>(P, rq, rqd, rqdd, r);
#pragma unroll
    for (int k = 0; k < 3; ++k) { V[k] = r.v[k]; V[3 + k] = r.w[k]; dV[k] = r.a[k]; dV[3 + k] = r.l[k]; }
  }
}

template <class T>
__device__ __forceinline__ const T* sensor_R(const FastParams<T>& P, const T* sp, bool generic) { return generic ? sp + GP_SENR : P.senR; }

// ---------------------------------------------------------------------------------------------
// regressor rows from given twists (get_regressor_matrix, dynamics.py:215-249), AoS
// ---------------------------------------------------------------------------------------------
template <class T>
__device__ __forceinline__ void store_rows(const T (&top)[3][4], const T (&bot)[3][9], T* __restrict__ Y) {
#pragma unroll
  for (int r = 0; r < 3; ++r) {
#pragma unroll
    for (int c = 0; c < 10; ++c) {
      Y[r * 10 + c] = c < 4 ? top[r][c] : T(0);
      Y[(3 + r) * 10 + c] = c == 0 ? T(0) : bot[r][c - 1];
    }
  }
}

template <class T>
__global__ void __launch_bounds__(kRowsBlock) k_regressor_rows(const T* __restrict__ tw, const T* __restrict__ dtw, T* __restrict__ Y, int64_t n) {
  const int64_t s = (int64_t)blockIdx.x * kRowsBlock + threadIdx.x;
  if (s >= n) return;
  T V[6], dV[6];
#pragma unroll
  for (int k = 0; k < 6; ++k) { V[k] = tw[s * 6 + k]; dV[k] = dtw[s * 6 + k]; }
  T top[3][4], bot[3][9];
  regressor_blocks(V, dV, top, bot);
  store_rows(top, bot, Y + s * 60);
}

// sensor-frame twists from last-link twists, AoS (core/simulate.py:202-209)
template <class T>
struct PoseArg { T R[9]; T t[3]; };

template <class T>
__global__ void __launch_bounds__(kRowsBlock) k_sensor_twists(const __grid_constant__ PoseArg<T> pose, const T* __restrict__ tw, const T* __restrict__ dtw,
                                                              T* __restrict__ tws, T* __restrict__ dtws, int64_t n) {
  const int64_t s = (int64_t)blockIdx.x * kRowsBlock + threadIdx.x;
  if (s >= n) return;
  T V[6], dV[6], Vs[6], dVs[6];
#pragma unroll
  for (int k = 0; k < 6; ++k) { V[k] = tw[s * 6 + k]; dV[k] = dtw[s * 6 + k]; }
  sensor_twists(pose.R, pose.t, V, dV, Vs, dVs);
#pragma unroll
  for (int k = 0; k < 6; ++k) { tws[s * 6 + k] = Vs[k]; dtws[s * 6 + k] = dVs[k]; }
}

// ---------------------------------------------------------------------------------------------
// fused: (q, qd, qdd) -> sensor-frame twists -> regressor rows / predicted wrench
// ---------------------------------------------------------------------------------------------
template <class T, int PATH>
__global__ void __launch_bounds__(kRowsBlock) k_regressor_from_traj(const __grid_constant__ FastParams<T> P, const T* __restrict__ gp, int nj, int nparams,
                                                                    const T* __restrict__ q, const T* __restrict__ qd, const T* __restrict__ qdd,
                                                                    T* __restrict__ Y /* [n][60] or null */, T* __restrict__ Vs_out /* [6][ld] or null */,
                                                                    T* __restrict__ dVs_out, const T* __restrict__ phi /* [10] device or null */,
                                                                    T* __restrict__ F_out /* [6][ld] or null */, int64_t n, int64_t ld) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* sp = reinterpret_cast<T*>(smem_raw);
  if constexpr (PATH == PATH_GENERIC) {
    for (int i = threadIdx.x; i < nparams; i += blockDim.x) sp[i] = gp[i];
    __syncthreads();
  }
  const int64_t s = (int64_t)blockIdx.x * kRowsBlock + threadIdx.x;
  if (s >= n) return;
  T V[6], dV[6], Vs[6], dVs[6];
  last_link_twists<T, PATH>(P, sp, nj, q, qd, qdd, s, ld, V, dV);
  const T* R = sensor_R(P, sp, PATH == PATH_GENERIC);
  sensor_twists(R, R + 9, V, dV, Vs, dVs);
  if (Vs_out) {
#pragma unroll
    for (int k = 0; k < 6; ++k) { Vs_out[k * ld + s] = Vs[k]; dVs_out[k * ld + s] = dVs[k]; }
  }
  if (!Y && !F_out) return;
  T top[3][4], bot[3][9];
  regressor_blocks(Vs, dVs, top, bot);
  if (Y) store_rows(top, bot, Y + s * 60);
  if (F_out) {  // F = Y phi : the wrench a body with parameters phi would load the sensor with
    T ph[10];
#pragma unroll
    for (int k = 0; k < 10; ++k) ph[k] = __ldg(phi + k);
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      T f = T(0), m = T(0);
#pragma unroll
      for (int c = 0; c < 4; ++c) f += top[r][c] * ph[c];
#pragma unroll
      for (int c = 0; c < 9; ++c) m += bot[r][c] * ph[1 + c];
      F_out[r * ld + s] = f;
      F_out[(3 + r) * ld + s] = m;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Gram accumulation
// ---------------------------------------------------------------------------------------------
// bot_nz, gram_accumulate, gram_pack_entry: rbm_gram.cuh (shared with the host harness)

// Per-sample work shared by both Gram kernels: (q, qd, qdd, f) in registers -> regressor blocks -> accumulate.
template <class T, int PATH>
__device__ __forceinline__ void gram_sample_fast(const FastParams<T>& P, const T (&rq)[6], const T (&rqd)[6], const T (&rqdd)[6], const T (&fs)[6],
                                                 T (&acc)[kAcc]) {
  FastResult<T> r;
  if constexpr (PATH == PATH_SEQ_ISO) fast_rnea<T, SeqIso, false>(P, rq, rqd, rqdd, r);
  else fast_rnea<T, SeqRigid, false>(P, rq, rqd, rqdd, r);
  T V[6], dV[6], Vs[6], dVs[6];
#pragma unroll
  for (int k = 0; k < 3; ++k) { V[k] = r.v[k]; V[3 + k] = r.w[k]; dV[k] = r.a[k]; dV[3 + k] = r.l[k]; }
  if (P.sen_diag != T(0)) {  // warp-uniform: Ad(T) is a component-wise sign pattern
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const T d = P.senR[4 * k];
      Vs[k] = d * V[k]; Vs[3 + k] = d * V[3 + k];
      dVs[k] = d * dV[k]; dVs[3 + k] = d * dV[3 + k];
    }
  } else {
    sensor_twists(P.senR, P.sent, V, dV, Vs, dVs);
  }
  T top[3][4], bot[3][9];
  regressor_blocks(Vs, dVs, top, bot);
  gram_accumulate(acc, top, bot, fs);
}

// fp32 only: warp-reduce the float partial sums and add them to the warp's double accumulators in shared memory
__device__ __forceinline__ void gram_flush_f32(float (&acc)[kAcc], double* red_warp, int lane) {
#pragma unroll
  for (int k = 0; k < kAcc; ++k) {
    float v = acc[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == (k & 31)) red_warp[k] += (double)v;
    acc[k] = 0.f;
  }
}

// block epilogue: registers -> per-warp doubles in shared memory -> fixed-order sum over warps -> partials[blockIdx]
template <class T>
__device__ __forceinline__ void gram_block_epilogue(T (&acc)[kAcc], double (*red)[kAcc], double* __restrict__ partials) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if constexpr (sizeof(T) == 4) {
    gram_flush_f32(acc, red[warp], lane);
  } else {
#pragma unroll
    for (int k = 0; k < kAcc; ++k) {
      double v = acc[k];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == (k & 31)) red[warp][k] += v;
    }
  }
  __syncthreads();
  if (threadIdx.x < kAcc) {
    double v = 0.0;
#pragma unroll
    for (int wq = 0; wq < kGramBlock / 32; ++wq) v += red[wq][threadIdx.x];  // fixed order
    partials[(int64_t)blockIdx.x * kAcc + threadIdx.x] = v;
  }
}

constexpr int kFlush = 64;  // fp32: samples accumulated in float registers between flushes into double (rel. error ~64 * 2^-24)

// ---- simple variant: direct global loads (any kernel path, any alignment) ----------------------------------
template <class T, int PATH>
__global__ void __launch_bounds__(kGramBlock, 1) k_regressor_gram(const __grid_constant__ FastParams<T> P, const T* __restrict__ gp, int nj, int nparams,
                                                                  const T* __restrict__ q, const T* __restrict__ qd, const T* __restrict__ qdd,
                                                                  const T* __restrict__ f, double* __restrict__ partials /* [grid][kAcc] */,
                                                                  int64_t n, int64_t ld) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* sp = reinterpret_cast<T*>(smem_raw);
  if constexpr (PATH == PATH_GENERIC) {
    for (int i = threadIdx.x; i < nparams; i += blockDim.x) sp[i] = gp[i];
  }
  __shared__ double red[kGramBlock / 32][kAcc];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int k = lane; k < kAcc; k += 32) red[warp][k] = 0.0;
  __syncthreads();
  T acc[kAcc];
#pragma unroll
  for (int k = 0; k < kAcc; ++k) acc[k] = T(0);
  const T* R = sensor_R(P, sp, PATH == PATH_GENERIC);
  const int64_t stride = (int64_t)gridDim.x * kGramBlock;
  // every lane of a warp runs the same number of iterations so the periodic warp reduction stays convergent
  const int64_t first = (int64_t)blockIdx.x * kGramBlock + warp * 32;
  int since_flush = 0;
  for (int64_t base = first; base < n; base += stride) {
    const int64_t s = base + lane;
    if (s < n) {
      T V[6], dV[6], Vs[6], dVs[6], fs[6];
#pragma unroll
      for (int k = 0; k < 6; ++k) fs[k] = __ldg(f + k * ld + s);
      last_link_twists<T, PATH>(P, sp, nj, q, qd, qdd, s, ld, V, dV);
      sensor_twists(R, R + 9, V, dV, Vs, dVs);
      T top[3][4], bot[3][9];
      regressor_blocks(Vs, dVs, top, bot);
      gram_accumulate(acc, top, bot, fs);
    }
    if constexpr (sizeof(T) == 4) {
      if (++since_flush == kFlush) {
        since_flush = 0;
        gram_flush_f32(acc, red[warp], lane);
      }
    }
  }
  gram_block_epilogue<T>(acc, red, partials);
}

// ---- pipelined variant (fast paths): TMA bulk copies stage whole 256-sample tiles of the 24 input streams (24 x 2 KB in
// fp64) into shared memory kGramStages tiles ahead of the arithmetic, so the loads never wait on registers or occupancy.
// What was measured on B200 while getting here (round 1, 12.5 M samples, G samples/s fp64 / fp32):
//   direct global loads                                   23.6 / 25.0   (41 % of stall samples on the load scoreboard)
//   TMA, one elected thread issues all 24 copies          21.8 / 34.9   (the issuing warp pays ~24 x UBLKCP per tile)
//   same + "last warp to drain refills" instead of a CTA barrier   20.8 / 34.6   (the barrier was not the limiter)
//   per-warp pipelines with 256-byte copies               17.6 / 26.5   (8x more copies: TMA-issue bound)
//   TMA, the 24 copies spread over the 8 warps (this code) 23.9 / 42.6
// Two later restructurings, both bit-identical and both rejected on measurement (tools/kbench, 12.5 M samples, 200 launches,
// fp64: this code 29.0 G samples/s at burst clocks):
//   per-stage `empty` mbarriers + refill one tile late instead of __syncthreads (warps free to drift 3 tiles)   28.8 (fp32 45.6 vs 47.0)
//   warp-specialised: 8 kinematics warps (setmaxnreg 128) hand [V_s | dV_s | f] through shared memory to
//     4 accumulator warps (setmaxnreg 240, one per scheduler, 2 samples per tile each)                           25.3 ... 26.9
//   fp32 only: accumulators held as float pairs, rank-1 rows as packed FFMA2 with a scalar-broadcast operand
//     (109 instead of 180 FMA-class instructions per sample, bit-identical sums)                             fp32 44.8 vs 47.0
// i.e. neither the CTA barrier nor the lock-step phases are the limiter (and FFMA2 buys no issue bandwidth here); the FP64 pipe is ~60 % busy and what is left is
// dependent-issue latency inside each thread, which only more resident warps (registers!) or less arithmetic would remove.
// In fp64 the kernel is then bound by the FP64 pipe at 8 warps per SM (250 registers: 70 double accumulators).
constexpr int kStreams = 24;  // q(6) qd(6) qdd(6) f(6)
constexpr int kGramStages = 4;

template <class T, int PATH>
__global__ void __launch_bounds__(kGramBlock, sizeof(T) == 4 ? 2 : 1) k_regressor_gram_tma(const __grid_constant__ FastParams<T> P, const T* __restrict__ q,
                                                                                            const T* __restrict__ qd, const T* __restrict__ qdd,
                                                                                            const T* __restrict__ f, double* __restrict__ partials,
                                                                                            int64_t n, int64_t ld) {
  constexpr int S = kGramStages;
  constexpr uint32_t kRowBytes = kGramBlock * sizeof(T);
  constexpr uint32_t kTileBytes = kStreams * kRowBytes;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  T* buf = reinterpret_cast<T*>(smem_raw);  // [S][kStreams][kGramBlock]
  __shared__ __align__(8) uint64_t full[S];
  __shared__ double red[kGramBlock / 32][kAcc];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t nfull = n / kGramBlock;  // tiles moved by bulk copies; a ragged tail tile is read directly
  for (int k = lane; k < kAcc; k += 32) red[warp][k] = 0.0;
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < S; ++s) mbar_init(&full[s], kGramBlock / 32);  // one arrive.expect_tx per warp (its share of the 24 copies)
    mbar_init_fence();
  }
  __syncthreads();

  // Issue cost of a bulk copy is paid by the issuing warp, so the 24 copies of a tile are spread over the 8 warps: lane 0 of
  // warp w brings streams w, w + 8, w + 16 and announces their bytes on the stage's barrier.
  auto issue = [&](int64_t it) {
    const int64_t tile = (int64_t)blockIdx.x + it * gridDim.x;
    if (tile >= nfull) return;
    const int st = (int)(it % S);
    uint64_t* bar = &full[st];
    mbar_arrive_expect_tx(bar, 3 * kRowBytes);
    T* dst = buf + (size_t)st * kStreams * kGramBlock;
    const int64_t s0 = tile * kGramBlock;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const int k = warp + 8 * i;  // stream index 0..23 = array (k / 6), row (k % 6)
      bulk_copy_g2s(dst + k * kGramBlock, (k < 6 ? q : k < 12 ? qd : k < 18 ? qdd : f) + (int64_t)(k % 6) * ld + s0, kRowBytes, bar);
    }
  };
  if (lane == 0) {
    for (int it = 0; it < S; ++it) issue(it);
  }

  T acc[kAcc];
#pragma unroll
  for (int k = 0; k < kAcc; ++k) acc[k] = T(0);
  int since_flush = 0;
  for (int64_t it = 0;; ++it) {
    const int64_t tile = (int64_t)blockIdx.x + it * gridDim.x;
    if (tile >= nfull) break;
    const int st = (int)(it % S);
    mbar_wait(&full[st], (uint32_t)((it / S) & 1));
    const T* src = buf + (size_t)st * kStreams * kGramBlock + tid;
    T rq[6], rqd[6], rqdd[6], fs[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      rq[k] = src[k * kGramBlock];
      rqd[k] = src[(6 + k) * kGramBlock];
      rqdd[k] = src[(12 + k) * kGramBlock];
      fs[k] = src[(18 + k) * kGramBlock];
    }
    __syncthreads();               // every thread holds its sample in registers: the stage can be refilled
    if (lane == 0) issue(it + S);  // each warp re-issues its three streams
    gram_sample_fast<T, PATH>(P, rq, rqd, rqdd, fs, acc);
    if constexpr (sizeof(T) == 4) {
      if (++since_flush == kFlush) {
        since_flush = 0;
        gram_flush_f32(acc, red[warp], lane);
      }
    }
  }
  // ragged tail (n % 256 samples): owned by the CTA that would have received tile `nfull`
  if ((nfull % gridDim.x) == blockIdx.x) {
    const int64_t s = nfull * kGramBlock + tid;
    if (s < n) {
      T rq[6], rqd[6], rqdd[6], fs[6];
#pragma unroll
      for (int k = 0; k < 6; ++k) {
        rq[k] = __ldg(q + k * ld + s);
        rqd[k] = __ldg(qd + k * ld + s);
        rqdd[k] = __ldg(qdd + k * ld + s);
        fs[k] = __ldg(f + k * ld + s);
      }
      gram_sample_fast<T, PATH>(P, rq, rqd, rqdd, fs, acc);
    }
  }
  gram_block_epilogue<T>(acc, red, partials);
}

// ---- software-pipelined variant (round 2; the production kernel of the fast paths) ---------------------------------------------
// Two changes against k_regressor_gram_tma, both from the round-1 ncu reading (FP64 pipe ~60 % busy at 2 warps per scheduler, stalls
// on fixed-latency dependencies; 1.14x the needed DRAM traffic):
//   * only the LIVE streams are staged.  V_6 and dV_6 of the sequential structure do not depend on the three gantry positions
//     (SequentialDesc::q_matters), so q[0..2] are never read: 21 bulk copies per tile instead of 24, 168 instead of 192 B of DRAM
//     traffic per fp64 sample, and the smaller tile buys a fifth stage.
//   * the accumulation of sample i is issued together with the kinematics of sample i + 1.  The kinematic chain (three sincos, six
//     links) is one long dependent sequence and the 180 accumulation FMAs are mutually independent, but inside one sample the second
//     needs the end of the first; carrying (x, w, dw, f) -- 15 scalars -- over one iteration puts both in the same basic block, so
//     each warp fills its own chain bubbles instead of relying on the one other warp of its scheduler.
// The sums are the same terms in the same order as before: results are bit-identical to k_regressor_gram_tma.
constexpr int kLiveStreams = 21;  // q3..5 (3) | qd (6) | qdd (6) | f (6)
constexpr int kPipeStages = 5;

template <class T>
struct GramCarry {
  T x[3], w[3], l[3], f[6];
};

// accumulation half
template <class T>
__device__ __forceinline__ void gram_accumulate_carry(T (&acc)[kAcc], const GramCarry<T>& z) {
  T top[3][4], bot[3][9];
  regressor_blocks_xwl(z.x, z.w, z.l, top, bot);
  gram_accumulate(acc, top, bot, z.f);
}

// kinematics half: (q, qd, qdd) and the joint sines / cosines -> sensor-frame (x = dv + w x v, w, dw)
template <class T, int PATH, bool SEN_DIAG>
__device__ __forceinline__ void gram_kinematics_cs(const FastParams<T>& P, const T (&rq)[6], const T (&c)[6], const T (&s)[6], const T (&rqd)[6],
                                                   const T (&rqdd)[6], GramCarry<T>& z) {
  FastResult<T> r;
  if constexpr (PATH == PATH_SEQ_ISO) fast_rnea_cs<T, SeqIso, false>(P, rq, c, s, rqd, rqdd, r);
  else fast_rnea_cs<T, SeqRigid, false>(P, rq, c, s, rqd, rqdd, r);
  T V[6], dV[6], Vs[6], dVs[6];
#pragma unroll
  for (int k = 0; k < 3; ++k) { V[k] = r.v[k]; V[3 + k] = r.w[k]; dV[k] = r.a[k]; dV[3 + k] = r.l[k]; }
  if constexpr (SEN_DIAG) {  // Ad(T) is a component-wise sign pattern (the reference's F/T site: Rz(180 deg), no offset)
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const T d = P.senR[4 * k];
      Vs[k] = d * V[k]; Vs[3 + k] = d * V[3 + k];
      dVs[k] = d * dV[k]; dVs[3 + k] = d * dV[3 + k];
    }
  } else {
    sensor_twists(P.senR, P.sent, V, dV, Vs, dVs);
  }
  regressor_x(Vs, dVs, z.x);
#pragma unroll
  for (int k = 0; k < 3; ++k) { z.w[k] = Vs[3 + k]; z.l[k] = dVs[3 + k]; }
}

// One pipelined step: kinematics of the current sample (into zn) TOGETHER with the accumulation of the previous one (z).  The
// branch on the trigonometric fast range comes FIRST and each arm holds the whole step, so that in the (always taken) fast arm the
// dependent chain -- range reduction, polynomials, six links -- and the 180 independent accumulation FMAs are ONE basic block for
// ptxas to interleave.  (With the usual "if (ok) fast else lib; then the rest" shape the chain is cut into blocks and the
// accumulation is scheduled after it: verified in the SASS.)
template <class T, int PATH, bool SEN_DIAG>
__device__ __forceinline__ void gram_step(const FastParams<T>& P, const T (&rq)[6], const T (&rqd)[6], const T (&rqdd)[6], GramCarry<T>& zn,
                                          const GramCarry<T>& z, T (&acc)[kAcc]) {
  T c[6], s[6];
#pragma unroll
  for (int i = 0; i < 3; ++i) { c[i] = T(1); s[i] = T(0); }
  if (trig_fast_ok(rq[3]) && trig_fast_ok(rq[4]) && trig_fast_ok(rq[5])) {
#pragma unroll
    for (int i = 3; i < 6; ++i) sincos_core(rq[i], s[i], c[i]);
    gram_kinematics_cs<T, PATH, SEN_DIAG>(P, rq, c, s, rqd, rqdd, zn);
    gram_accumulate_carry(acc, z);
  } else {
#pragma unroll
    for (int i = 3; i < 6; ++i) sincos_lib(rq[i], &s[i], &c[i]);
    gram_kinematics_cs<T, PATH, SEN_DIAG>(P, rq, c, s, rqd, rqdd, zn);
    gram_accumulate_carry(acc, z);
  }
}

template <class T>
__device__ __forceinline__ const T* live_stream(int k, const T* q, const T* qd, const T* qdd, const T* f, int64_t ld) {
  return k < 3 ? q + (int64_t)(3 + k) * ld : k < 9 ? qd + (int64_t)(k - 3) * ld : k < 15 ? qdd + (int64_t)(k - 9) * ld : f + (int64_t)(k - 15) * ld;
}

template <class T, int PATH, bool SEN_DIAG, bool PIPELINE>
__global__ void __launch_bounds__(kGramBlock, sizeof(T) == 4 ? 2 : 1) k_regressor_gram_pipe(const __grid_constant__ FastParams<T> P, const T* __restrict__ q,
                                                                                             const T* __restrict__ qd, const T* __restrict__ qdd,
                                                                                             const T* __restrict__ f, double* __restrict__ partials,
                                                                                             int64_t n, int64_t ld) {
  constexpr int S = kPipeStages;
  constexpr int NW = kGramBlock / 32;
  constexpr uint32_t kRowBytes = kGramBlock * sizeof(T);
  extern __shared__ __align__(128) unsigned char smem_raw[];
  T* buf = reinterpret_cast<T*>(smem_raw);  // [S][kLiveStreams][kGramBlock]
  __shared__ __align__(8) uint64_t full[S];
  __shared__ double red[NW][kAcc];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t nfull = n / kGramBlock;  // tiles moved by bulk copies; a ragged tail tile is read directly
  for (int k = lane; k < kAcc; k += 32) red[warp][k] = 0.0;
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < S; ++s) mbar_init(&full[s], NW);  // one arrive.expect_tx per warp (its share of the 21 copies)
    mbar_init_fence();
  }
  __syncthreads();

  // The issue cost of a bulk copy is paid by the issuing warp: lane 0 of warp w brings streams w, w + 8, w + 16 (< 21).  Their
  // source rows are fixed for the whole kernel, so the pointers are formed once.
  const int my_copies = (warp + 2 * NW < kLiveStreams) ? 3 : 2;
  const T* my_src[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) my_src[i] = live_stream(warp + NW * i < kLiveStreams ? warp + NW * i : warp, q, qd, qdd, f, ld);
  auto issue = [&](int64_t it) {
    const int64_t tile = (int64_t)blockIdx.x + it * gridDim.x;
    if (tile >= nfull) return;
    const int st = (int)(it % S);
    uint64_t* bar = &full[st];
    mbar_arrive_expect_tx(bar, my_copies * kRowBytes);
    T* dst = buf + ((size_t)st * kLiveStreams + warp) * kGramBlock;
    const int64_t s0 = tile * kGramBlock;
    bulk_copy_g2s(dst, my_src[0] + s0, kRowBytes, bar);
    bulk_copy_g2s(dst + NW * kGramBlock, my_src[1] + s0, kRowBytes, bar);
    if (my_copies == 3) bulk_copy_g2s(dst + 2 * NW * kGramBlock, my_src[2] + s0, kRowBytes, bar);
  };
  if (lane == 0) {
    for (int it = 0; it < S; ++it) issue(it);
  }

  T acc[kAcc];
#pragma unroll
  for (int k = 0; k < kAcc; ++k) acc[k] = T(0);
  GramCarry<T> z;  // PIPELINE: the previous sample, not yet accumulated (zeros: accumulating them adds exact zeros)
#pragma unroll
  for (int k = 0; k < 3; ++k) { z.x[k] = T(0); z.w[k] = T(0); z.l[k] = T(0); z.f[k] = T(0); z.f[3 + k] = T(0); }
  int since_flush = PIPELINE ? -1 : 0;  // PIPELINE: the first accumulation is the all-zero carry and does not count
  for (int64_t it = 0;; ++it) {
    const int64_t tile = (int64_t)blockIdx.x + it * gridDim.x;
    if (tile >= nfull) break;
    const int st = (int)(it % S);
    mbar_wait(&full[st], (uint32_t)((it / S) & 1));
    const T* src = buf + (size_t)st * kLiveStreams * kGramBlock + tid;
    T rq[6], rqd[6], rqdd[6];
    GramCarry<T> zn;
    rq[0] = rq[1] = rq[2] = T(0);  // dead inputs of the sequential structure
#pragma unroll
    for (int k = 0; k < 3; ++k) rq[3 + k] = src[k * kGramBlock];
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      rqd[k] = src[(3 + k) * kGramBlock];
      rqdd[k] = src[(9 + k) * kGramBlock];
      zn.f[k] = src[(15 + k) * kGramBlock];
    }
    __syncthreads();               // every thread holds its sample in registers: the stage can be refilled
    if (lane == 0) issue(it + S);  // each warp re-issues its streams
    if constexpr (PIPELINE) {
      gram_step<T, PATH, SEN_DIAG>(P, rq, rqd, rqdd, zn, z, acc);
      z = zn;
    } else {
      T c[6], s[6];
      fast_sincos<T, SeqIso>(rq, c, s);
      gram_kinematics_cs<T, PATH, SEN_DIAG>(P, rq, c, s, rqd, rqdd, zn);
      gram_accumulate_carry(acc, zn);
    }
    if constexpr (sizeof(T) == 4) {
      if (++since_flush == kFlush) {
        since_flush = 0;
        gram_flush_f32(acc, red[warp], lane);
      }
    }
  }
  if constexpr (PIPELINE) gram_accumulate_carry(acc, z);
  // ragged tail (n % 256 samples): owned by the CTA that would have received tile `nfull`
  if ((nfull % gridDim.x) == blockIdx.x) {
    const int64_t s = nfull * kGramBlock + tid;
    if (s < n) {
      T rq[6], rqd[6], rqdd[6], c[6], sn[6];
      rq[0] = rq[1] = rq[2] = T(0);
#pragma unroll
      for (int k = 0; k < 3; ++k) rq[3 + k] = __ldg(q + (3 + k) * ld + s);
#pragma unroll
      for (int k = 0; k < 6; ++k) {
        rqd[k] = __ldg(qd + k * ld + s);
        rqdd[k] = __ldg(qdd + k * ld + s);
        z.f[k] = __ldg(f + k * ld + s);
      }
      fast_sincos<T, SeqIso>(rq, c, sn);
      gram_kinematics_cs<T, PATH, SEN_DIAG>(P, rq, c, sn, rqd, rqdd, z);
      gram_accumulate_carry(acc, z);
    }
  }
  gram_block_epilogue<T>(acc, red, partials);
}

// ---- warp-pair variant: 12 warps per SM instead of 8 by halving the accumulators each thread keeps ---------------------------------
// Round-2 ncu reading of the kernels above (fp64): the accumulation phase is FP64-pipe bound, the kinematic chain phase is
// LATENCY bound (two warps per scheduler issue ~46 % of the cycles), and the warp count is pinned by the 70 double accumulators
// (140 registers) every thread keeps.  Here warps work in pairs on the SAME 64 samples: each warp runs the kinematics of its own
// 32 samples, publishes (x, w, dw) -- 9 scalars per sample -- in shared memory, and after a 64-thread named barrier accumulates
// ONE HALF of the 70 entries (gram_accumulate_set<0 / 1>, rbm_gram.cuh) for both its own and its partner's samples.  36
// accumulators per thread -> <= 168 registers -> 384 threads per SM.  The input stage is released one iteration late (the wrench
// rows of the partner's samples are read from it), so nothing has to be parked in registers across the CTA barrier.
constexpr int kPairWarps = 12;
constexpr int kPairBlock = kPairWarps * 32;
constexpr int kPairStages = 3;
constexpr int kPairZ = 9;  // x(3) w(3) dw(3)

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

template <class T, int N>
__device__ __forceinline__ void gram_flush_f32_n(float (&acc)[N], double* red_warp, int lane) {
#pragma unroll
  for (int k = 0; k < N; ++k) {
    float v = acc[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == (k & 31)) red_warp[k] += (double)v;
    acc[k] = 0.f;
  }
}

template <int SET, class T>
__device__ __forceinline__ void pair_accumulate(T (&acc)[kSetMax], const T (&x)[3], const T (&w)[3], const T (&l)[3], const T (&fs)[6]) {
  T top[3][4], bot[3][9];
  regressor_blocks_xwl(x, w, l, top, bot);
  gram_accumulate_set<SET>(acc, top, bot, fs);
}

template <class T, int PATH, bool SEN_DIAG>
__global__ void __launch_bounds__(kPairBlock, 1) k_regressor_gram_pair(const __grid_constant__ FastParams<T> P, const T* __restrict__ q,
                                                                       const T* __restrict__ qd, const T* __restrict__ qdd, const T* __restrict__ f,
                                                                       double* __restrict__ partials, int64_t n, int64_t ld) {
  constexpr int S = kPairStages, NW = kPairWarps, TILE = kPairBlock, HALF = NW / 2;
  constexpr uint32_t kRowBytes = TILE * sizeof(T);
  extern __shared__ __align__(128) unsigned char smem_raw[];
  T* buf = reinterpret_cast<T*>(smem_raw);                   // [S][kLiveStreams][TILE]
  T* zbuf = buf + (size_t)S * kLiveStreams * TILE;           // [NW][kPairZ][32]
  __shared__ __align__(8) uint64_t full[S];
  __shared__ double red[NW][kSetMax];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int set = warp < HALF ? 0 : 1;
  const int partner = warp < HALF ? warp + HALF : warp - HALF;
  const int pair_bar = 1 + (warp < HALF ? warp : warp - HALF);  // named barriers 1..6 (0 is __syncthreads)
  const int64_t nfull = n / TILE;
  for (int k = lane; k < kSetMax; k += 32) red[warp][k] = 0.0;
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < S; ++s) mbar_init(&full[s], NW);
    mbar_init_fence();
  }
  __syncthreads();

  // 21 copies per tile over 12 warps: warp w brings streams w and (w + 12 < 21) w + 12
  const int my_copies = (warp + NW < kLiveStreams) ? 2 : 1;
  const T* my_src[2];
  my_src[0] = live_stream(warp, q, qd, qdd, f, ld);
  my_src[1] = live_stream(warp + NW < kLiveStreams ? warp + NW : warp, q, qd, qdd, f, ld);
  auto issue = [&](int64_t it) {
    const int64_t tile = (int64_t)blockIdx.x + it * gridDim.x;
    if (tile >= nfull) return;
    const int st = (int)(it % S);
    uint64_t* bar = &full[st];
    mbar_arrive_expect_tx(bar, my_copies * kRowBytes);
    T* dst = buf + ((size_t)st * kLiveStreams + warp) * TILE;
    const int64_t s0 = tile * TILE;
    bulk_copy_g2s(dst, my_src[0] + s0, kRowBytes, bar);
    if (my_copies == 2) bulk_copy_g2s(dst + NW * TILE, my_src[1] + s0, kRowBytes, bar);
  };
  if (lane == 0) {
    for (int it = 0; it < S; ++it) issue(it);
  }

  T acc[kSetMax];
#pragma unroll
  for (int k = 0; k < kSetMax; ++k) acc[k] = T(0);

  // one tile: own kinematics -> exchange -> this warp's half of the entries for both samples of the lane pair
  auto process = [&](const T* stage, bool valid) {
    const T* src = stage + tid;
    T rq[6], rqd[6], rqdd[6], c[6], sn[6];
    rq[0] = rq[1] = rq[2] = T(0);  // dead inputs of the sequential structure
#pragma unroll
    for (int k = 0; k < 3; ++k) rq[3 + k] = src[k * TILE];
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      rqd[k] = src[(3 + k) * TILE];
      rqdd[k] = src[(9 + k) * TILE];
    }
    GramCarry<T> z;
    fast_sincos<T, SeqIso>(rq, c, sn);
    gram_kinematics_cs<T, PATH, SEN_DIAG>(P, rq, c, sn, rqd, rqdd, z);
    if (!valid) {  // padding sample of the ragged tail: contributes exact zeros (its wrench rows are zero too)
#pragma unroll
      for (int k = 0; k < 3; ++k) { z.x[k] = T(0); z.w[k] = T(0); z.l[k] = T(0); }
    }
    T* zb = zbuf + (size_t)warp * kPairZ * 32 + lane;
#pragma unroll
    for (int k = 0; k < 3; ++k) { zb[k * 32] = z.x[k]; zb[(3 + k) * 32] = z.w[k]; zb[(6 + k) * 32] = z.l[k]; }
    named_bar_sync(pair_bar, 64);
    const T* zp = zbuf + (size_t)partner * kPairZ * 32 + lane;
    T px[3], pw[3], pl[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) { px[k] = zp[k * 32]; pw[k] = zp[(3 + k) * 32]; pl[k] = zp[(6 + k) * 32]; }
    if (set == 0) {
      T fo[6], fp[6];
      const T* fsrc = stage + 15 * TILE;
#pragma unroll
      for (int k = 0; k < 6; ++k) { fo[k] = fsrc[k * TILE + tid]; fp[k] = fsrc[k * TILE + partner * 32 + lane]; }
      pair_accumulate<0>(acc, z.x, z.w, z.l, fo);
      pair_accumulate<0>(acc, px, pw, pl, fp);
    } else {
      const T none[6] = {T(0), T(0), T(0), T(0), T(0), T(0)};  // set 1 holds no wrench entry
      pair_accumulate<1>(acc, z.x, z.w, z.l, none);
      pair_accumulate<1>(acc, px, pw, pl, none);
    }
  };

  int since_flush = 0;
  for (int64_t it = 0;; ++it) {
    const int64_t tile = (int64_t)blockIdx.x + it * gridDim.x;
    if (tile >= nfull) break;
    const int st = (int)(it % S);
    mbar_wait(&full[st], (uint32_t)((it / S) & 1));
    __syncthreads();  // every warp is done with iteration it - 1: its stage and the exchange buffers are free
    if (lane == 0 && it > 0) issue(it - 1 + S);  // refill the stage consumed by the PREVIOUS iteration
    process(buf + (size_t)st * kLiveStreams * TILE, true);
    if constexpr (sizeof(T) == 4) {
      if (++since_flush == kFlush / 2) {  // two samples per thread and iteration
        since_flush = 0;
        gram_flush_f32_n<T, kSetMax>(acc, red[warp], lane);
      }
    }
  }
  // ragged tail (n % 384 samples): staged by hand into stage 0 (zeros beyond n) by the CTA that would have received tile `nfull`
  if ((nfull % gridDim.x) == blockIdx.x && nfull * TILE < n) {
    __syncthreads();
    const int64_t s = nfull * TILE + tid;
    const bool valid = s < n;
#pragma unroll
    for (int k = 0; k < kLiveStreams; ++k) buf[(size_t)k * TILE + tid] = valid ? __ldg(live_stream(k, q, qd, qdd, f, ld) + s) : T(0);
    __syncthreads();
    process(buf, valid);
  }
  // epilogue: registers -> per-warp doubles -> fixed-order sum over the six warps of each set -> partials[blockIdx][70]
  if constexpr (sizeof(T) == 4) {
    gram_flush_f32_n<T, kSetMax>(acc, red[warp], lane);
  } else {
#pragma unroll
    for (int k = 0; k < kSetMax; ++k) {
      double v = acc[k];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == (k & 31)) red[warp][k] += v;
    }
  }
  __syncthreads();
  if (tid < kAcc) {
    int kset, local;
    pair_locate(tid, kset, local);
    double v = 0.0;
#pragma unroll
    for (int wq = 0; wq < HALF; ++wq) v += red[kset * HALF + wq][local];  // fixed order
    partials[(int64_t)blockIdx.x * kAcc + tid] = v;
  }
}

// partials [nblocks][70] -> pack [112] = [Y^T Y (100, row-major) | Y^T f (10) | f^T f | n]
__global__ void __launch_bounds__(128) k_gram_finalize(const double* __restrict__ partials, int nblocks, double n_samples, double* __restrict__ pack) {
  __shared__ double tot[kAcc];
  if (threadIdx.x < kAcc) {
    double v = 0.0;
    for (int b = 0; b < nblocks; ++b) v += partials[(int64_t)b * kAcc + threadIdx.x];  // fixed order
    tot[threadIdx.x] = v;
  }
  __syncthreads();
  const int t = threadIdx.x;
  if (t < 111) pack[t] = gram_pack_entry(tot, t);
  else if (t == 111) pack[t] = n_samples;
}

// ---- grouped variant: one Gram pack per GROUP (environment / object), one group per thread ----------------------------------
// For logs laid out [frame][row][group] (the closed-loop rollout's frame log, rbm_linearize.cu): thread g walks its n_frames samples,
// keeps the 70 accumulators in registers and writes its own 112-double pack, value-major [112][ld_out] (coalesced).
template <class T, int PATH>
__global__ void __launch_bounds__(kRowsBlock) k_regressor_gram_grouped(const __grid_constant__ FastParams<T> P, const T* __restrict__ gp, int nj, int nparams,
                                                                       const T* __restrict__ q, const T* __restrict__ qd, const T* __restrict__ qdd,
                                                                       const T* __restrict__ f, int64_t frame_stride, int64_t f_frame_stride,
                                                                       int64_t n_frames, double* __restrict__ packs, int64_t n_groups, int64_t ld, int64_t ld_out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* sp = reinterpret_cast<T*>(smem_raw);
  if constexpr (PATH == PATH_GENERIC) {
    for (int i = threadIdx.x; i < nparams; i += blockDim.x) sp[i] = gp[i];
    __syncthreads();
  }
  const int64_t g = (int64_t)blockIdx.x * kRowsBlock + threadIdx.x;
  if (g >= n_groups) return;
  const T* R = sensor_R(P, sp, PATH == PATH_GENERIC);
  double acc[kAcc];
#pragma unroll
  for (int k = 0; k < kAcc; ++k) acc[k] = 0.0;
#pragma unroll 1
  for (int64_t fr = 0; fr < n_frames; ++fr) {
    const int64_t off = fr * frame_stride;
    T V[6], dV[6], Vs[6], dVs[6], fs[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) fs[k] = __ldg(f + fr * f_frame_stride + k * ld + g);
    last_link_twists<T, PATH>(P, sp, nj, q + off, qd + off, qdd + off, g, ld, V, dV);
    sensor_twists(R, R + 9, V, dV, Vs, dVs);
    T top[3][4], bot[3][9];
    regressor_blocks(Vs, dVs, top, bot);
    gram_accumulate(acc, top, bot, fs);
  }
  double* out = packs + g;
#pragma unroll
  for (int t = 0; t < 111; ++t) out[(int64_t)t * ld_out] = gram_pack_entry(acc, t);
  out[(int64_t)111 * ld_out] = (double)n_frames;
}

// ---------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------
static int sm_count(int device) {
  static int cached[64] = {0};
  if (device >= 0 && device < 64 && cached[device]) return cached[device];
  int v = 148;
  cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device);
  if (device >= 0 && device < 64) cached[device] = v;
  return v;
}

// persistent grid: `per_sm` 256-thread CTAs per SM (fp64: 1, register-limited; fp32: 2)
int gram_grid(const rbm_model* m, int64_t n, int per_sm) {
  int64_t need = (n + kGramBlock - 1) / kGramBlock;
  int64_t cap = (int64_t)sm_count(m->device) * per_sm;
  return (int)(need < cap ? (need > 0 ? need : 1) : cap);
}

template <class T>
int launch_regressor_rows(const T* tw, const T* dtw, T* Y, int64_t n, cudaStream_t st) {
  if (n == 0) return RBM_OK;
  k_regressor_rows<T><<<(unsigned)((n + kRowsBlock - 1) / kRowsBlock), kRowsBlock, 0, st>>>(tw, dtw, Y, n);
  RBM_CUDA_TRY(cudaGetLastError());
  return RBM_OK;
}

template <class T>
int launch_sensor_twists(const double* pose_Rt, const T* tw, const T* dtw, T* tws, T* dtws, int64_t n, cudaStream_t st) {
  if (n == 0) return RBM_OK;
  PoseArg<T> p;
  for (int k = 0; k < 9; ++k) p.R[k] = (T)pose_Rt[k];
  for (int k = 0; k < 3; ++k) p.t[k] = (T)pose_Rt[9 + k];
  k_sensor_twists<T><<<(unsigned)((n + kRowsBlock - 1) / kRowsBlock), kRowsBlock, 0, st>>>(p, tw, dtw, tws, dtws, n);
  RBM_CUDA_TRY(cudaGetLastError());
  return RBM_OK;
}

template <class T>
int launch_regressor_from_traj(const rbm_model* m, const T* q, const T* qd, const T* qdd, T* Y, T* Vs, T* dVs, const T* phi, T* F, int64_t n, int64_t ld,
                               cudaStream_t st) {
  if (n == 0) return RBM_OK;
  const unsigned grid = (unsigned)((n + kRowsBlock - 1) / kRowsBlock);
  const int np = generic_param_count(m->nj);
  const FastParams<T>& P = ModelView<T>::fast(m);
  const T* gp = ModelView<T>::generic(m);
  if (m->path == PATH_SEQ_ISO) k_regressor_from_traj<T, PATH_SEQ_ISO><<<grid, kRowsBlock, 0, st>>>(P, gp, m->nj, np, q, qd, qdd, Y, Vs, dVs, phi, F, n, ld);
  else if (m->path == PATH_SEQ_RIGID) k_regressor_from_traj<T, PATH_SEQ_RIGID><<<grid, kRowsBlock, 0, st>>>(P, gp, m->nj, np, q, qd, qdd, Y, Vs, dVs, phi, F, n, ld);
  else k_regressor_from_traj<T, PATH_GENERIC><<<grid, kRowsBlock, sizeof(T) * np, st>>>(P, gp, m->nj, np, q, qd, qdd, Y, Vs, dVs, phi, F, n, ld);
  RBM_CUDA_TRY(cudaGetLastError());
  return RBM_OK;
}

template <class T>
int launch_regressor_gram_grouped(const rbm_model* m, const T* q, const T* qd, const T* qdd, const T* f, int64_t frame_stride, int64_t f_frame_stride,
                                  int64_t n_frames,
                                  double* packs, int64_t n_groups, int64_t ld, int64_t ld_out, cudaStream_t st) {
  if (n_groups == 0) return RBM_OK;
  const unsigned grid = (unsigned)((n_groups + kRowsBlock - 1) / kRowsBlock);
  const int np = generic_param_count(m->nj);
  const FastParams<T>& P = ModelView<T>::fast(m);
  const T* gp = ModelView<T>::generic(m);
  if (m->path == PATH_SEQ_ISO)
    k_regressor_gram_grouped<T, PATH_SEQ_ISO><<<grid, kRowsBlock, 0, st>>>(P, gp, m->nj, np, q, qd, qdd, f, frame_stride, f_frame_stride, n_frames, packs, n_groups, ld, ld_out);
  else if (m->path == PATH_SEQ_RIGID)
    k_regressor_gram_grouped<T, PATH_SEQ_RIGID><<<grid, kRowsBlock, 0, st>>>(P, gp, m->nj, np, q, qd, qdd, f, frame_stride, f_frame_stride, n_frames, packs, n_groups, ld, ld_out);
  else
    k_regressor_gram_grouped<T, PATH_GENERIC><<<grid, kRowsBlock, sizeof(T) * np, st>>>(P, gp, m->nj, np, q, qd, qdd, f, frame_stride, f_frame_stride, n_frames, packs, n_groups,
                                                                                      ld, ld_out);
  RBM_CUDA_TRY(cudaGetLastError());
  return RBM_OK;
}
template int launch_regressor_gram_grouped<double>(const rbm_model*, const double*, const double*, const double*, const double*, int64_t, int64_t, int64_t, double*,
                                                   int64_t, int64_t, int64_t, cudaStream_t);

// development knob for A/B timing of the TMA-fed Gram kernels: RBM_GRAM_VARIANT = 0 round-1 kernel (24 streams) | 1 live streams +
// software-pipelined accumulation | 2 live streams (default) | 3 warp-pair split
static int gram_variant() {
  static const int v = [] {
    const char* e = getenv("RBM_GRAM_VARIANT");
    return e ? atoi(e) : 2;
  }();
  return v;
}

template <class T>
int launch_regressor_gram(const rbm_model* m, const T* q, const T* qd, const T* qdd, const T* f, double* pack, double* partials, int64_t n, int64_t ld,
                          cudaStream_t st) {
  int grid = gram_grid(m, n, 1);
  const int np = generic_param_count(m->nj);
  const FastParams<T>& P = ModelView<T>::fast(m);
  const T* gp = ModelView<T>::generic(m);
  if (n > 0) {
    // bulk copies need 16-byte aligned rows: base pointers and the row pitch
    auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
    const bool tma_ok = m->path != PATH_GENERIC && al16(q) && al16(qd) && al16(qdd) && al16(f) && ((ld * sizeof(T)) % 16 == 0) && n >= kGramBlock &&
                        !m->no_tma;
    const int variant = gram_variant();
    const int dev = (m->device >= 0 && m->device < 64) ? m->device : 0;
    const bool diag = P.sen_diag != T(0);
    // picks the instantiation for (kernel path, sensor-pose form), raises its dynamic shared-memory limit once per device, launches
#define RBM_LAUNCH_GRAM(KERNEL, BLOCK, SMEM, ...)                                                                                  \
  do {                                                                                                                             \
    auto kfn = m->path == PATH_SEQ_ISO ? (diag ? KERNEL<T, PATH_SEQ_ISO, true __VA_ARGS__> : KERNEL<T, PATH_SEQ_ISO, false __VA_ARGS__>)       \
                                       : (diag ? KERNEL<T, PATH_SEQ_RIGID, true __VA_ARGS__> : KERNEL<T, PATH_SEQ_RIGID, false __VA_ARGS__>);  \
    static std::atomic<const void*> ready[64][4];                                                                                  \
    const int slot = (m->path == PATH_SEQ_ISO ? 0 : 2) + (diag ? 0 : 1);                                                           \
    if (ready[dev][slot].load(std::memory_order_acquire) != (const void*)kfn) {                                                    \
      RBM_CUDA_TRY(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(SMEM)));                           \
      ready[dev][slot].store((const void*)kfn, std::memory_order_release);                                                         \
    }                                                                                                                              \
    kfn<<<grid, BLOCK, SMEM, st>>>(P, q, qd, qdd, f, partials, n, ld);                                                              \
  } while (0)
    if constexpr (sizeof(T) == 4) {
      if (tma_ok && (variant == 4 || m->gram_tc)) {  // fp32 mode on the tensor cores (rbm_gram_tc.cu), opt-in
        grid = tc_gram_grid(sm_count(m->device), n);
        int rc = launch_regressor_gram_tc(m, q, qd, qdd, f, partials, n, ld, grid, st);
        if (rc != RBM_OK) return rc;
        k_gram_finalize<<<1, 128, 0, st>>>(partials, grid, (double)n, pack);
        RBM_CUDA_TRY(cudaGetLastError());
        return RBM_OK;
      }
    }
    if (tma_ok && variant == 0) {  // round-1 kernel, kept for A/B runs (RBM_GRAM_VARIANT=0)
      grid = gram_grid(m, n, sizeof(T) == 4 ? 2 : 1);  // fp32: 119 registers, 96 KB of stages: two CTAs per SM
      constexpr size_t smem = (size_t)kGramStages * kStreams * kGramBlock * sizeof(T);
      static std::atomic<bool> attr_set[64];  // function attributes are per device
      if (!attr_set[dev].load(std::memory_order_acquire)) {
        RBM_CUDA_TRY(cudaFuncSetAttribute(k_regressor_gram_tma<T, PATH_SEQ_ISO>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        RBM_CUDA_TRY(cudaFuncSetAttribute(k_regressor_gram_tma<T, PATH_SEQ_RIGID>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set[dev].store(true, std::memory_order_release);
      }
      if (m->path == PATH_SEQ_ISO) k_regressor_gram_tma<T, PATH_SEQ_ISO><<<grid, kGramBlock, smem, st>>>(P, q, qd, qdd, f, partials, n, ld);
      else k_regressor_gram_tma<T, PATH_SEQ_RIGID><<<grid, kGramBlock, smem, st>>>(P, q, qd, qdd, f, partials, n, ld);
    } else if (tma_ok && variant == 3 && n >= kPairBlock) {  // warp-pair split: 12 warps per SM
      const int64_t tiles = n / kPairBlock;
      const int64_t cap = sm_count(m->device);
      grid = (int)(tiles < cap ? tiles : cap);
      constexpr size_t smem = ((size_t)kPairStages * kLiveStreams * kPairBlock + (size_t)kPairWarps * kPairZ * 32) * sizeof(T);
      RBM_LAUNCH_GRAM(k_regressor_gram_pair, kPairBlock, smem, );
    } else if (tma_ok) {
      grid = gram_grid(m, n, sizeof(T) == 4 ? 2 : 1);  // fp32: two CTAs per SM (105 KB of stages each)
      constexpr size_t smem = (size_t)kPipeStages * kLiveStreams * kGramBlock * sizeof(T);
      if (variant == 1) RBM_LAUNCH_GRAM(k_regressor_gram_pipe, kGramBlock, smem, , true);
      else RBM_LAUNCH_GRAM(k_regressor_gram_pipe, kGramBlock, smem, , false);
#undef RBM_LAUNCH_GRAM
    } else if (m->path == PATH_SEQ_ISO) {
      k_regressor_gram<T, PATH_SEQ_ISO><<<grid, kGramBlock, 0, st>>>(P, gp, m->nj, np, q, qd, qdd, f, partials, n, ld);
    } else if (m->path == PATH_SEQ_RIGID) {
      k_regressor_gram<T, PATH_SEQ_RIGID><<<grid, kGramBlock, 0, st>>>(P, gp, m->nj, np, q, qd, qdd, f, partials, n, ld);
    } else {
      k_regressor_gram<T, PATH_GENERIC><<<grid, kGramBlock, sizeof(T) * np, st>>>(P, gp, m->nj, np, q, qd, qdd, f, partials, n, ld);
    }
    RBM_CUDA_TRY(cudaGetLastError());
  }
  k_gram_finalize<<<1, 128, 0, st>>>(partials, n > 0 ? grid : 0, (double)n, pack);
  RBM_CUDA_TRY(cudaGetLastError());
  return RBM_OK;
}

#define RBM_INST(T)                                                                                                                        \
  template int launch_regressor_rows<T>(const T*, const T*, T*, int64_t, cudaStream_t);                                                    \
  template int launch_sensor_twists<T>(const double*, const T*, const T*, T*, T*, int64_t, cudaStream_t);                                  \
  template int launch_regressor_from_traj<T>(const rbm_model*, const T*, const T*, const T*, T*, T*, T*, const T*, T*, int64_t, int64_t, cudaStream_t); \
  template int launch_regressor_gram<T>(const rbm_model*, const T*, const T*, const T*, const T*, double*, double*, int64_t, int64_t, cudaStream_t);
RBM_INST(double)
RBM_INST(float)

}  // namespace rbm
