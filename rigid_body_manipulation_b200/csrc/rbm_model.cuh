// Model constants as the kernels see them.
//
// The reference binds these once with functools.partial (core/simulate.py:150-156):
// hposes_body_parent, simats_body, uscrews_body, twist_0, dtwist_0 (+ defaults wrench_tip = 0,
// pose_tip_ee = I of dynamics/dynamics.py:116-117) and the static sensor pose of
// core/simulate.py:202.  rbm_model_create() packs them for two kernel families:
//
//   * GENERIC  : any chain of nj <= RBM_MAX_JOINTS joints, dense 6x6 inertias, arbitrary unit
//                screws, arbitrary home / tip / sensor poses.  One flat array of scalars (layout below): a by-value
//                kernel argument (constant bank) for the unrolled nj = 6 kernel, staged once per block into shared
//                memory for the run-time-nj kernels.
//   * FAST     : chains whose structure matches a compile-time descriptor (rbm_rnea.cuh);
//                only the surviving numeric parameters travel, as a by-value kernel argument
//                (constant bank: operands of the FMAs, no load instructions at all).
#pragma once
#include <cstdint>

#define RBM_MAX_JOINTS 16

namespace rbm {

// ---- generic packed layout (scalar offsets) ------------------------------------------------
enum : int {
  GP_V0 = 0,      // twist_0 (6)
  GP_DV0 = 6,     // dtwist_0 (6)
  GP_FTIP = 12,   // wrench_tip (6)
  GP_TIPR = 18,   // pose_tip_ee R row-major (9)
  GP_TIPT = 27,   // pose_tip_ee t (3)
  GP_SENR = 30,   // sensor pose R (9)   [core/simulate.py:202  pose_sen_llj]
  GP_SENT = 39,   // sensor pose t (3)
  GP_HEAD = 42,
  // per joint block.  T_i(q) = SE3.exp(-S q) M_i (dynamics.py:126) is AFFINE in (cos, sin, q) of the joint once the model constants are
  // folded at rbm_model_create (host, fp64):  with a = w/|w|, r = v/|w|, (s, c) = sincos(-|w| q), M = (hR, ht)
  //   R(q) = c RA + s RB + RC          RA = (1 - a a^T) hR            RB = [a]x hR             RC = a a^T hR
  //   p(q) = c PA + s PB + q PC + PD   PA = (1 - a a^T) ht - a x r    PB = a x ht + r - (a.r) a
  //                                    PC = -(a.v) a                  PD = a a^T ht + a x r
  // (Rodrigues + liegroups' left Jacobian J_l(th a) (th r) = s r + (th - s)(a.r) a + (1 - c) a x r: no division by the angle, so no
  // small-angle branch is needed -- at |th| <= 1e-8, where liegroups switches to first order, the two differ by th^2 <= 1e-16.)
  // Prismatic joint (w == 0): RA = RB = 0, RC = hR, PA = PB = 0, PC = -v, PD = ht, NW = 0.
  GJ_RA = 0,      // (9) row-major
  GJ_RB = 9,      // (9)
  GJ_RC = 18,     // (9)
  GJ_PA = 27,     // (3)
  GJ_PB = 30,     // (3)
  GJ_PC = 33,     // (3)
  GJ_PD = 36,     // (3)
  GJ_S = 39,      // unit screw [v; w] (6)
  GJ_NW = 45,     // -|w|: rotation angle per unit q (0 = prismatic)
  GJ_G = 46,      // spatial inertia, dense row-major (36)   simats_body[i+1]
  GJ_RIGID = 82,  // 1 when G has the rigid-body form [[m 1, -[h]x], [[h]x, Ibar]], Ibar symmetric (then 10 numbers define it), else 0
  GJ_STRIDE = 83
};
static inline int generic_param_count(int nj) { return GP_HEAD + GJ_STRIDE * nj; }

// the same block as a by-value kernel argument for a compile-time joint count (nj = 6: 540 scalars, 4.3 KB in fp64; kernel arguments
// may be up to 32 KB since CUDA 12.1): the unrolled recursion then reads every model constant as a constant-bank FMA operand
template <class T, int NJ>
struct GenericBlock {
  T v[GP_HEAD + GJ_STRIDE * NJ];
};

// ---- fast-path parameters (by-value kernel argument) ---------------------------------------
template <class T>
struct FastParams {
  T g[3];        // linear part of dtwist_0 (angular part and twist_0 are structurally zero)
  T mass[6];     // G[0][0]
  T h[6][3];     // first moment m*c, from the lower-left block [h]x of G
  T I[6][6];     // rotational inertia about the joint-frame origin: xx, yy, zz, xy, yz, zx
  T tm[6][3];    // home translations (only read by descriptors that declare them non-zero)
  T senR[9];     // sensor pose rotation, row-major
  T sent[3];     // sensor pose translation
  T sen_diag;    // 1 when senR is diagonal (+-1 entries) and sent == 0 (the reference's F/T site: Rz(180 deg)), else 0
};

enum KernelPath : int { PATH_GENERIC = 0, PATH_SEQ_ISO = 1, PATH_SEQ_RIGID = 2 };

}  // namespace rbm
