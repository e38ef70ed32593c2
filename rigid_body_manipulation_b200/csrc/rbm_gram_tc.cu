// fp32-mode fused regressor + Gram with the accumulation on the 5th-generation tensor cores (tcgen05.mma kind::tf32, accumulator
// in TMEM).  Replaces, batched and fused, the same reference path as rbm_regressor.cu: core/simulate.py:202-209 (sensor-frame
// twists), dynamics/dynamics.py:215-249 (get_regressor_matrix) and the normal equations of loggers/loggers.py:127-129.
//
// Why a GEMM exists here at all.  Every entry of the 6x11 matrix [Y f] of one sample is a 0 / +-1 combination of 18 FEATURES
//     z = (x (3), dw (3), ww (6) = wx wx, wy wy, wz wz, wx wy, wy wz, wz wx, f (6)),       x = dv + w x v   (rbm_rnea.cuh)
// so the 70 distinct entries of sum_samples [Y f]^T [Y f] are a FIXED linear image of the second-moment matrix
//     M2 = sum_samples z z^T     (18 x 18),
// and M2 = Z^T Z is a real dense contraction with K = number of samples: M = N = 18 (+ padding), K = 12.5 M per GPU.  The register
// kernel (rbm_regressor.cu) spends 250 of its 523 instructions per sample on the 180 accumulation FMAs and is issue-bound; here
// each thread only WRITES its 18 features to shared memory and one elected thread feeds the tensor core.
//
// Precision.  kind::tf32 keeps 11 significand bits of each operand, far too few for a Gram that is then inverted, so every feature
// is split exactly as z = hi + lo with hi = z rounded toward zero to tf32 (exactly representable) and lo = z - hi (<= 13 significant
// bits; the tensor core drops at most its last two).  The operand rows are [hi (18) | lo (18) | 4 zero rows] = 40, and
//     D = [hi; lo] [hi; lo]^T   ->   M2 = D_hh + D_hl + D_lh + D_ll
// carries every cross term: products of tf32 values are exact in fp32, so the only errors are |z - hi - tf32(lo)| <= 2^-21 |z| per
// feature and the fp32 accumulation inside TMEM, which is bounded by flushing TMEM into fp64 every kTcFlush tiles (2048 samples).
//
// Structure (one CTA = 8 compute warps + 2 service warps, two CTAs per SM, 256-sample tiles):
//   TMA warp, one elected lane     : bulk copies of the 21 live input rows (3 stages)
//   tensor-core warp, one lane     : the 32 tcgen05.mma of the PREVIOUS tile (8 warp tiles x 4 K-steps of 8 samples, M = 64,
//                                    N = 40), tcgen05.commit -> mbarrier
//   compute warps                  : inputs -> kinematics -> features -> hi / lo -> 36 conflict-free STS into the warp's own
//                                    K-major, 128-byte-swizzled operand tile (40 rows x 32 samples); warps 0..2 flush TMEM
// One CTA barrier per tile hands the finished operand tiles to the tensor-core warp and the consumed input stage to the TMA warp.
//
// STATUS (round 2, measured on B200, 12.5 M samples): parity-correct (Gram within 1.3e-5 of the fp64 normal equations, bar 1e-4) but
// SLOWER than the register kernel -- 34.5 vs 47.3 G samples/s.  ncu: tensor pipe 34 % active, i.e. ~65 cycles of pipe time per
// M64 N40 K8 tcgen05.mma (N = 8 costs the same; M = 128: 74): a K = 8 instruction only carries 8 samples, the per-instruction cost
// is fixed, and 32 of them per 256-sample tile cap the kernel at ~35 G samples/s regardless of the compute side (48.8 with the
// MMAs switched off).  It is therefore NOT the default; RBM_FLAG_GRAM_TENSOR_CORES selects it (tests, smoke, A/B timing).
// Numbers and the ncu captures: experiments/README.md, profiles/r2_gram32_tc_ncu.csv.
#include <atomic>
#include <cstdlib>

#include "rbm_async.cuh"
#include "rbm_gram.cuh"
#include "rbm_internal.h"
#include "rbm_rnea.cuh"
#include "rbm_tcgen05.cuh"

namespace rbm {

constexpr int kTcComputeWarps = 8;
constexpr int kTcBlock = 32 * (kTcComputeWarps + 2);  // 320: eight compute warps, one tensor-core warp, one TMA warp
constexpr int kTcTile = 32 * kTcComputeWarps;          // 256 samples
constexpr int kTcStages = 3;
constexpr int kTcStreams = 21;                         // q3..5 | qd | qdd | f (the live rows, as in rbm_regressor.cu)
constexpr int kTcFeat = 18;
constexpr int kTcRows = 40;                            // hi (18) | lo (18) | 4 padding rows: five 8-row swizzle atoms
constexpr int kTcM = 64, kTcN = 40;                    // UMMA shape.  Rows 40..63 of A alias whatever follows the warp tile in shared memory: their
                                                       // outputs are never read.  (M = 128 was measured too: 74 instead of 65 cycles per MMA.)
constexpr int kTcFlush = 8;                            // tiles accumulated in TMEM (fp32) between flushes into fp64
constexpr int kTcTmemCols = 64;                        // power of two >= kTcN
constexpr int kTcWarpTileBytes = kTcRows * 128;        // 5120
constexpr int kTcOpBytes = kTcComputeWarps * kTcWarpTileBytes;                // 40960
constexpr int kTcStageBytes = kTcStreams * kTcTile * (int)sizeof(float);     // 21504
constexpr int kTcAccLd = 19;                                                  // padded row pitch of the fp64 accumulator (bank spread)
constexpr int kTcAccBytes = 36 * kTcAccLd * (int)sizeof(double);             // 5472
constexpr int kTcSmemBytes = 1008 + kTcOpBytes + kTcStages * kTcStageBytes + kTcAccBytes;  // + slack for the manual 1024-byte alignment

__device__ __forceinline__ const float* tc_live_stream(int k, const float* q, const float* qd, const float* qdd, const float* f, int64_t ld) {
  return k < 3 ? q + (int64_t)(3 + k) * ld : k < 9 ? qd + (int64_t)(k - 3) * ld : k < 15 ? qdd + (int64_t)(k - 9) * ld : f + (int64_t)(k - 15) * ld;
}

template <int PATH, bool SEN_DIAG>
__global__ void __launch_bounds__(kTcBlock, 2) k_regressor_gram_tc(const __grid_constant__ FastParams<float> P, const float* __restrict__ q,
                                                                   const float* __restrict__ qd, const float* __restrict__ qdd,
                                                                   const float* __restrict__ f, double* __restrict__ partials, int64_t n, int64_t ld) {
  constexpr int S = kTcStages;
  extern __shared__ unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t full[S];
  __shared__ __align__(8) uint64_t mma_done;
  __shared__ uint32_t tmem_base_slot;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool service = warp >= kTcComputeWarps;
  const bool mma_warp = warp == kTcComputeWarps, tma_warp = warp == kTcComputeWarps + 1;

  // carve shared memory: operand tiles first (1024-byte aligned for the swizzle), then the input stages, then the fp64 accumulator
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t op_addr = (raw_addr + 1023u) & ~1023u;
  unsigned char* op = smem_raw + (op_addr - raw_addr);
  float* stages = reinterpret_cast<float*>(op + kTcOpBytes);
  double* acc64 = reinterpret_cast<double*>(op + kTcOpBytes + S * kTcStageBytes);  // [36][kTcAccLd]: row r of D, columns folded hi + lo

  // local tile list: full tiles blockIdx.x + j * gridDim.x (TMA), plus the ragged tail when this CTA would have received tile nfull
  const int64_t nfull = n / kTcTile;
  const int64_t Lfull = nfull > (int64_t)blockIdx.x ? (nfull - 1 - blockIdx.x) / gridDim.x + 1 : 0;
  const bool owns_tail = (nfull * kTcTile < n) && ((nfull % gridDim.x) == blockIdx.x);
  const int64_t L = Lfull + (owns_tail ? 1 : 0);

  for (int i = tid; i < 36 * kTcAccLd; i += kTcBlock) acc64[i] = 0.0;
  for (int i = tid; i < kTcOpBytes / 4; i += kTcBlock) reinterpret_cast<float*>(op)[i] = 0.f;  // padding rows stay zero for good
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < S; ++s) mbar_init(&full[s], 1);
    mbar_init(&mma_done, 1);
    mbar_init_fence();
  }
  if (mma_warp) {
    tmem_alloc(&tmem_base_slot, kTcTmemCols);
    tmem_relinquish_alloc_permit();
  }
  fence_proxy_async_smem();  // the zeroed operand region is later read by the tensor core through the async proxy
  tc_fence_before_thread_sync();
  __syncthreads();
  tc_fence_after_thread_sync();
  const uint32_t tmem_d = tmem_base_slot;

  if (service) {
    // =================================================== service warp ===================================================
    // One lane does all the issuing, so its instruction count per tile IS the pace of the CTA (a first version with run-time
    // loops and per-copy pointer selection needed ~1100 instructions per tile and held the kernel at 21 G samples/s even with the
    // MMAs switched off).  Everything that does not change from tile to tile is therefore formed once: the 21 source row pointers,
    // the shared-memory addresses, and the base descriptor; the per-tile code is straight-line.
    const float* src[kTcStreams];
#pragma unroll
    for (int k = 0; k < kTcStreams; ++k) src[k] = tc_live_stream(k, q, qd, qdd, f, ld);
    const uint32_t stage_addr = smem_u32(stages);
    const uint32_t full_addr = smem_u32(&full[0]);
    const uint32_t done_addr = smem_u32(&mma_done);
    const int64_t tile_stride = (int64_t)gridDim.x * kTcTile;
    int64_t next_off = (int64_t)blockIdx.x * kTcTile;  // element offset of the next tile to request
    int64_t next_j = 0;
    int next_st = 0;
    auto issue_tma = [&]() {  // local tile next_j -> stage next_st
      if (next_j >= Lfull) return;
      const uint32_t bar = full_addr + 8u * (uint32_t)next_st;
      const uint32_t dst = stage_addr + (uint32_t)next_st * (uint32_t)kTcStageBytes;
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)kTcStageBytes) : "memory");
#pragma unroll
      for (int k = 0; k < kTcStreams; ++k)
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst + (uint32_t)(k * kTcTile * 4)),
                     "l"(src[k] + next_off), "r"((uint32_t)(kTcTile * 4)), "r"(bar)
                     : "memory");
      next_off += tile_stride;
      ++next_j;
      next_st = next_st == S - 1 ? 0 : next_st + 1;
    };
    if (tma_warp && lane == 0) {
#pragma unroll 1
      for (int j = 0; j < S; ++j) issue_tma();
    }
    const uint32_t idesc = umma_idesc_tf32(kTcM, kTcN);
    const uint64_t desc0 = umma_desc_k_sw128(op_addr, 1024);
    int period = 0;  // local tile index modulo kTcFlush of the tile whose MMAs are issued next
    for (int64_t it = 0; it <= L; ++it) {
      __syncthreads();  // B(it): the operand tiles of local tile it - 1 are complete; the inputs of tile it are in registers
      tc_fence_after_thread_sync();
      if (mma_warp && lane == 0) {
        if (it > 0) {  // the tensor core first: the compute warps wait for this commit before they overwrite their operand tiles
          const uint32_t fresh = period == 0 ? 0u : 1u;  // first tile of a flush period overwrites the accumulator
#pragma unroll
          for (int w = 0; w < kTcComputeWarps; ++w) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              const uint64_t desc = desc0 + (uint64_t)((w * kTcWarpTileBytes + ks * 32) >> 4);  // start-address field only (no carry: < 2^14)
              if (w == 0 && ks == 0) {
                asm volatile(
                    "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d),
                    "l"(desc), "l"(desc), "r"(idesc), "r"(fresh)
                    : "memory");
              } else {
                asm volatile(
                    "{\n.reg .pred p;\nsetp.eq.b32 p, 0, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d),
                    "l"(desc), "l"(desc), "r"(idesc)
                    : "memory");
              }
            }
          }
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(done_addr) : "memory");
          period = period == kTcFlush - 1 ? 0 : period + 1;
        }
      }
      // the stage read at iteration it (all compute threads hold their inputs in registers after B(it)) receives tile it + S
      if (tma_warp && lane == 0 && it < L) issue_tma();
      __syncwarp();
    }
  } else {
    // =================================================== compute warps ===================================================
    // byte offsets of this lane's 4-byte slot inside a 128-byte operand row, for each row-in-atom j (16-byte chunk index XOR j)
    uint32_t slot[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) slot[j] = ((((uint32_t)lane >> 2) ^ (uint32_t)j) << 4) | (((uint32_t)lane & 3u) << 2);
    unsigned char* my_tile = op + warp * kTcWarpTileBytes;

    // TMEM -> fp64.  M = 64 accumulator layout (cta_group::1): row m of D lives in TMEM lane (m % 16) + 32 (m / 16), and warp s may
    // only touch lanes 32 s .. 32 s + 31: lanes 0..15 of warp s hold rows 16 s .. 16 s + 15 (warps 0..2 cover the 36 rows).
    auto flush = [&]() {
      if (warp < 3) {
        tc_fence_after_thread_sync();
        uint32_t r[kTcN];
        const uint32_t taddr = tmem_d + ((uint32_t)(32 * warp) << 16);
#pragma unroll
        for (int c0 = 0; c0 < kTcN; c0 += 8) {
          asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                       : "=r"(r[c0]), "=r"(r[c0 + 1]), "=r"(r[c0 + 2]), "=r"(r[c0 + 3]), "=r"(r[c0 + 4]), "=r"(r[c0 + 5]), "=r"(r[c0 + 6]), "=r"(r[c0 + 7])
                       : "r"(taddr + (uint32_t)c0)
                       : "memory");
        }
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < kTcN; ++c) asm volatile("" : "+r"(r[c]));  // values are defined only after the wait
        const int row = 16 * warp + lane;
        if (lane < 16 && row < 36) {
          double* a = acc64 + row * kTcAccLd;
#pragma unroll
          for (int b = 0; b < kTcFeat; ++b) a[b] += (double)__uint_as_float(r[b]) + (double)__uint_as_float(r[kTcFeat + b]);
        }
        tc_fence_before_thread_sync();
      }
    };

    for (int64_t it = 0; it < L; ++it) {
      const bool tail = it >= Lfull;
      float in[kTcStreams];
      bool valid = true;
      if (!tail) {
        const int st = (int)(it % S);
        mbar_wait(&full[st], (uint32_t)((it / S) & 1));
        const float* src = stages + (size_t)st * kTcStreams * kTcTile + tid;
#pragma unroll
        for (int k = 0; k < kTcStreams; ++k) in[k] = src[k * kTcTile];
      } else {
        const int64_t s = nfull * kTcTile + tid;
        valid = s < n;
#pragma unroll
        for (int k = 0; k < kTcStreams; ++k) in[k] = valid ? __ldg(tc_live_stream(k, q, qd, qdd, f, ld) + s) : 0.f;
      }
      __syncthreads();  // B(it)
      float rq[6], rqd[6], rqdd[6], c[6], sn[6];
      rq[0] = rq[1] = rq[2] = 0.f;  // dead inputs of the sequential structure
#pragma unroll
      for (int k = 0; k < 3; ++k) rq[3 + k] = in[k];
#pragma unroll
      for (int k = 0; k < 6; ++k) { rqd[k] = in[3 + k]; rqdd[k] = in[9 + k]; }
      fast_sincos<float, SeqIso>(rq, c, sn);
      FastResult<float> r;
      if constexpr (PATH == PATH_SEQ_ISO) fast_rnea_cs<float, SeqIso, false>(P, rq, c, sn, rqd, rqdd, r);
      else fast_rnea_cs<float, SeqRigid, false>(P, rq, c, sn, rqd, rqdd, r);
      float V[6], dV[6], Vs[6], dVs[6];
#pragma unroll
      for (int k = 0; k < 3; ++k) { V[k] = r.v[k]; V[3 + k] = r.w[k]; dV[k] = r.a[k]; dV[3 + k] = r.l[k]; }
      if constexpr (SEN_DIAG) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const float d = P.senR[4 * k];
          Vs[k] = d * V[k]; Vs[3 + k] = d * V[3 + k];
          dVs[k] = d * dV[k]; dVs[3 + k] = d * dV[3 + k];
        }
      } else {
        sensor_twists(P.senR, P.sent, V, dV, Vs, dVs);
      }
      float z[kTcFeat], x[3], ww[6];
      regressor_x(Vs, dVs, x);
      regressor_products(Vs + 3, ww);
#pragma unroll
      for (int k = 0; k < 3; ++k) { z[k] = x[k]; z[3 + k] = dVs[3 + k]; }
#pragma unroll
      for (int k = 0; k < 6; ++k) { z[6 + k] = ww[k]; z[12 + k] = in[15 + k]; }
      if (!valid) {
#pragma unroll
        for (int k = 0; k < kTcFeat; ++k) z[k] = 0.f;
      }
      if (it > 0) {
        mbar_wait(&mma_done, (uint32_t)((it - 1) & 1));  // the tensor core has consumed the operand tiles of local tile it - 1
        if ((it % kTcFlush) == 0) flush();                 // TMEM holds local tiles it - kTcFlush .. it - 1; tile it restarts the sum
      }
      // hi / lo split and store: row b = hi, row 18 + b = lo; byte offset of row R = (R >> 3) * 1024 + (R & 7) * 128 + slot[R & 7]
#pragma unroll
      for (int b = 0; b < kTcFeat; ++b) {
        const float hi = __uint_as_float(__float_as_uint(z[b]) & 0xffffe000u);
        const float lo = z[b] - hi;
        const int Rh = b, Rl = kTcFeat + b;
        *reinterpret_cast<float*>(my_tile + (Rh >> 3) * 1024 + (Rh & 7) * 128 + slot[Rh & 7]) = hi;
        *reinterpret_cast<float*>(my_tile + (Rl >> 3) * 1024 + (Rl & 7) * 128 + slot[Rl & 7]) = lo;
      }
      fence_proxy_async_smem();  // generic-proxy stores -> visible to the tensor core's async-proxy reads issued after B(it + 1)
    }
    __syncthreads();  // B(L): the service warp issues the last tile
    if (L > 0) {
      mbar_wait(&mma_done, (uint32_t)((L - 1) & 1));
      flush();  // whatever the last (partial) flush period left in TMEM
    }
  }
  tc_fence_before_thread_sync();
  __syncthreads();
  tc_fence_after_thread_sync();
  if (mma_warp) tmem_dealloc(tmem_d, kTcTmemCols);

  // ---- epilogue: second moments -> the 70-entry layout of rbm_gram.cuh (so k_gram_finalize is shared with the register kernels) ----
  // M18[a][b] = D_hh + D_hl + D_lh + D_ll  (columns were folded at flush time, rows are folded here)
  double* M18 = reinterpret_cast<double*>(op);               // [18][18]   (the operand region is free now)
  double* Tu = M18 + kTcFeat * kTcFeat;                      // [18][3][5]  U_r of the unit feature vector e_a
  double* Tw = Tu + kTcFeat * 3 * 5;                         // [18][3][10] W_r of e_a
  for (int i = tid; i < kTcFeat * kTcFeat; i += kTcBlock) {
    const int a = i / kTcFeat, b = i % kTcFeat;
    M18[i] = acc64[a * kTcAccLd + b] + acc64[(kTcFeat + a) * kTcAccLd + b];
  }
  if (tid < kTcFeat) {
    double x[3] = {0, 0, 0}, l[3] = {0, 0, 0}, ww[6] = {0, 0, 0, 0, 0, 0}, fw[6] = {0, 0, 0, 0, 0, 0};
    const int a = tid;
    if (a < 3) x[a] = 1.0;
    else if (a < 6) l[a - 3] = 1.0;
    else if (a < 12) ww[a - 6] = 1.0;
    else fw[a - 12] = 1.0;
    double top[3][4], bot[3][9];
    regressor_blocks_feat(x, l, ww, top, bot);
    for (int r = 0; r < 3; ++r) {
      for (int c = 0; c < 4; ++c) Tu[(a * 3 + r) * 5 + c] = top[r][c];
      Tu[(a * 3 + r) * 5 + 4] = fw[r];
      for (int c = 0; c < 9; ++c) Tw[(a * 3 + r) * 10 + c] = bot[r][c];
      Tw[(a * 3 + r) * 10 + 9] = fw[3 + r];
    }
  }
  __syncthreads();
  if (tid < kAcc) {
    // entry tid of the 70-layout: top block (5x5 upper triangle) first, then the bottom block (10x10 upper triangle)
    int i = 0, j = 0, width = 5;
    const double* T = Tu;
    int k = tid;
    if (k >= kTop) { k -= kTop; width = 10; T = Tw; }
    for (i = 0; i < width; ++i) {
      if (k < width - i) { j = i + k; break; }
      k -= width - i;
    }
    double v = 0.0;
    for (int r = 0; r < 3; ++r)
      for (int a = 0; a < kTcFeat; ++a) {
        const double ta = T[(a * 3 + r) * width + i];
        if (ta == 0.0) continue;
        for (int b = 0; b < kTcFeat; ++b) v += ta * T[(b * 3 + r) * width + j] * M18[a * kTcFeat + b];
      }
    partials[(int64_t)blockIdx.x * kAcc + tid] = v;
  }
}

// host side -------------------------------------------------------------------------------------------------------------------
int tc_gram_grid(int sms, int64_t n) {
  const int64_t tiles = (n + kTcTile - 1) / kTcTile;
  const int64_t cap = (int64_t)sms * 2;
  return (int)(tiles < cap ? (tiles > 0 ? tiles : 1) : cap);
}

int launch_regressor_gram_tc(const rbm_model* m, const float* q, const float* qd, const float* qdd, const float* f, double* partials, int64_t n, int64_t ld,
                             int grid, cudaStream_t st) {
  const FastParams<float>& P = ModelView<float>::fast(m);
  const bool diag = P.sen_diag != 0.f;
  const int dev = (m->device >= 0 && m->device < 64) ? m->device : 0;
  static std::atomic<bool> attr_set[64];
  if (!attr_set[dev].load(std::memory_order_acquire)) {
    RBM_CUDA_TRY(cudaFuncSetAttribute(k_regressor_gram_tc<PATH_SEQ_ISO, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmemBytes));
    RBM_CUDA_TRY(cudaFuncSetAttribute(k_regressor_gram_tc<PATH_SEQ_ISO, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmemBytes));
    RBM_CUDA_TRY(cudaFuncSetAttribute(k_regressor_gram_tc<PATH_SEQ_RIGID, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmemBytes));
    RBM_CUDA_TRY(cudaFuncSetAttribute(k_regressor_gram_tc<PATH_SEQ_RIGID, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmemBytes));
    attr_set[dev].store(true, std::memory_order_release);
  }
  if (m->path == PATH_SEQ_ISO) {
    if (diag) k_regressor_gram_tc<PATH_SEQ_ISO, true><<<grid, kTcBlock, kTcSmemBytes, st>>>(P, q, qd, qdd, f, partials, n, ld);
    else k_regressor_gram_tc<PATH_SEQ_ISO, false><<<grid, kTcBlock, kTcSmemBytes, st>>>(P, q, qd, qdd, f, partials, n, ld);
  } else {
    if (diag) k_regressor_gram_tc<PATH_SEQ_RIGID, true><<<grid, kTcBlock, kTcSmemBytes, st>>>(P, q, qd, qdd, f, partials, n, ld);
    else k_regressor_gram_tc<PATH_SEQ_RIGID, false><<<grid, kTcBlock, kTcSmemBytes, st>>>(P, q, qd, qdd, f, partials, n, ld);
  }
  RBM_CUDA_TRY(cudaGetLastError());
  return RBM_OK;
}

}  // namespace rbm
