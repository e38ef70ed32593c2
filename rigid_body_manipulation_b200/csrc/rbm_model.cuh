// Model constants as the kernels see them.
//
// The reference binds these once with functools.partial (core/simulate.py:150-156):
// hposes_body_parent, simats_body, uscrews_body, twist_0, dtwist_0 (+ defaults wrench_tip = 0,
// pose_tip_ee = I of dynamics/dynamics.py:116-117) and the static sensor pose of
// core/simulate.py:202.  rbm_model_create() packs them for two kernel families:
//
//   * GENERIC  : any chain of nj <= RBM_MAX_JOINTS joints, dense 6x6 inertias, arbitrary unit
//                screws, arbitrary home / tip / sensor poses.  One flat array of scalars in global
//                memory, staged ONCE PER BLOCK into shared memory (layout below).
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
  // per joint block
  GJ_HR = 0,      // home pose R row-major (9)   hposes_body_parent[i+1]
  GJ_HT = 9,      // home pose t (3)
  GJ_S = 12,      // unit screw [v; w] (6)
  GJ_AXIS = 18,   // w / |w| (3) (zeros when |w| == 0)
  GJ_WN = 21,     // |w| (1)
  GJ_G = 22,      // spatial inertia, dense row-major (36)   simats_body[i+1]
  GJ_RIGID = 58,  // 1 when G has the rigid-body form [[m 1, -[h]x], [[h]x, Ibar]], Ibar symmetric (then 10 numbers define it), else 0
  GJ_STRIDE = 59
};
static inline int generic_param_count(int nj) { return GP_HEAD + GJ_STRIDE * nj; }

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
