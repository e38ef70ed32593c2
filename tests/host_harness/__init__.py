"""TEST INFRASTRUCTURE ONLY: builds tests/host_harness/harness.cu (host instantiation of the kernels' per-sample device functions)
with nvcc and wraps it with ctypes.  See the header of harness.cu."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SRC = os.path.join(HERE, "harness.cu")
OUT = os.path.join(HERE, "_build")
LIB = os.path.join(OUT, "libharness.so")
CSRC = os.path.join(ROOT, "rigid_body_manipulation_b200", "csrc")


def nvcc_path():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    return exe if os.path.exists(exe) else None


def build():
    os.makedirs(OUT, exist_ok=True)
    deps = [SRC] + [os.path.join(CSRC, f) for f in ("rbm_rnea.cuh", "rbm_typed.cuh", "rbm_trig.cuh", "rbm_model.cuh", "rbm_dynamics.cuh", "rbm_gram.cuh", "rbm_setup.cuh")]
    if os.path.exists(LIB) and all(os.path.getmtime(LIB) >= os.path.getmtime(d) for d in deps):
        return LIB
    nvcc = nvcc_path()
    if nvcc is None:
        raise RuntimeError("nvcc not found")
    subprocess.run([nvcc, "-O2", "-std=c++17", "--expt-relaxed-constexpr", "-Wno-deprecated-gpu-targets", "-gencode", "arch=compute_100a,code=sm_100a",
                    "-shared", "-Xcompiler", "-fPIC", "-I", CSRC, "-I", os.path.join(ROOT, "include"), "-o", LIB, SRC], check=True)
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


PATH_ID = {"generic": 0, "seq_iso": 1, "seq_rigid": 2}


def fast_rnea(path, fast_params, traj, mode=0, dtype=np.float64):
    traj = np.ascontiguousarray(traj, dtype=dtype)
    n = traj.shape[0]
    tau, V, dV = np.empty((n, 6), dtype), np.empty((n, 6), dtype), np.empty((n, 6), dtype)
    fn = lib().h_fast_rnea_f64 if dtype == np.float64 else lib().h_fast_rnea_f32
    rc = fn(C.c_int(PATH_ID[path]), C.c_int(mode), _p(np.ascontiguousarray(fast_params, dtype=np.float64)), _p(traj), _p(tau), _p(V), _p(dV), C.c_int64(n))
    assert rc == 0
    return tau, V, dV


def generic_rnea(generic_params, nj, traj, unrolled=False, full=False):
    traj = np.ascontiguousarray(traj, dtype=np.float64)
    n = traj.shape[0]
    tau = np.empty((n, nj))
    poses = np.empty((n, nj, 12)) if full else None
    tw = np.empty((n, nj + 1, 6)) if full else None
    dtw = np.empty((n, nj + 1, 6)) if full else None
    V, dV = np.empty((n, 6)), np.empty((n, 6))
    rc = lib().h_generic_rnea_f64(_p(np.ascontiguousarray(generic_params, dtype=np.float64)), C.c_int(nj), C.c_int(int(unrolled)), _p(traj), _p(tau),
                                  _p(poses), _p(tw), _p(dtw), _p(V), _p(dV), C.c_int64(n))
    assert rc == 0
    return dict(tau=tau, poses=poses, twists=tw, dtwists=dtw, V=V, dV=dV)


def generic_rnea_f32(generic_params, nj, traj):
    traj = np.ascontiguousarray(traj, dtype=np.float32)
    tau = np.empty((traj.shape[0], nj), np.float32)
    rc = lib().h_generic_rnea_f32(_p(np.ascontiguousarray(generic_params, dtype=np.float64)), C.c_int(nj), _p(traj), _p(tau), C.c_int64(traj.shape[0]))
    assert rc == 0
    return tau


def sensor_regressor(pose_Rt, V, dV):
    V, dV = np.ascontiguousarray(V, dtype=np.float64), np.ascontiguousarray(dV, dtype=np.float64)
    n = V.shape[0]
    Vs, dVs, Y = np.empty((n, 6)), np.empty((n, 6)), np.empty((n, 6, 10))
    rc = lib().h_sensor_regressor_f64(_p(np.ascontiguousarray(pose_Rt, dtype=np.float64)), _p(V), _p(dV), _p(Vs), _p(dVs), _p(Y), C.c_int64(n))
    assert rc == 0
    return Vs, dVs, Y


def _lib_main():
    from rigid_body_manipulation_b200 import _lib as rbm_lib

    return rbm_lib.load()


def _model_args(analysis):
    """analysis = engine.analyze_model(...) = (path name, fast params, generic params) -> ctypes-ready (path id, fp, gp, nj)."""
    path, fast, generic = analysis
    fp = np.ascontiguousarray(fast, dtype=np.float64)
    gp = np.ascontiguousarray(generic, dtype=np.float64)
    nj = next(k for k in range(1, 17) if _lib_main().rbm_generic_param_count(k) == len(gp))  # the block's size names the joint count
    return C.c_int(PATH_ID[path]), fp, gp, C.c_int(nj)


def linearize(analysis, q, qd, u=None, dt=0.002, eps=1e-8, centered=True):
    """q, qd, u: (n, nj) -> A (n, 2nj, 2nj), B (n, 2nj, nj), qdd (n, nj) through csrc/rbm_dynamics.cuh::linearize_state on the host."""
    path, fp, gp, nj = _model_args(analysis)
    n, njv = q.shape
    qs, qds = np.ascontiguousarray(q.T, dtype=np.float64), np.ascontiguousarray(qd.T, dtype=np.float64)
    us = None if u is None else np.ascontiguousarray(u.T, dtype=np.float64)
    A, B, qdd = np.zeros((2 * njv, 2 * njv, n)), np.zeros((2 * njv, njv, n)), np.zeros((njv, n))
    rc = lib().h_linearize_f64(path, _p(fp), _p(gp), nj, _p(qs), _p(qds), _p(us), C.c_double(dt), C.c_double(eps), C.c_int(int(centered)), _p(A), _p(B),
                               _p(qdd), C.c_int64(n))
    assert rc == 0
    return A.transpose(2, 0, 1), B.transpose(2, 0, 1), qdd.T


def forward_dynamics(analysis, q, qd, u=None, dt=0.0):
    path, fp, gp, nj = _model_args(analysis)
    n, njv = q.shape
    qs, qds = np.ascontiguousarray(q.T, dtype=np.float64), np.ascontiguousarray(qd.T, dtype=np.float64)
    us = None if u is None else np.ascontiguousarray(u.T, dtype=np.float64)
    qdd = np.zeros((njv, n))
    qn, qdn = (np.zeros((njv, n)), np.zeros((njv, n))) if dt > 0 else (None, None)
    rc = lib().h_forward_dynamics_f64(path, _p(fp), _p(gp), nj, _p(qs), _p(qds), _p(us), C.c_double(dt), _p(qdd), _p(qn), _p(qdn), C.c_int64(n))
    assert rc == 0
    return (qdd.T, qn.T, qdn.T) if dt > 0 else qdd.T


def closed_loop(analysis, plan, K, phi, q0, qd0=None, dt=None, fps=50.0, div=None, max_frames=None):
    """q0 [, qd0]: (n, nj).  Returns dict(frames (F, 3nj+18, n), frame_steps (F,), final (3nj, n)) like engine.Model.closed_loop."""
    path, fp, gp, nj = _model_args(analysis)
    n, njv = q0.shape
    dt = float(plan.timestep if dt is None else dt)
    div = float(njv if div is None else div)
    if max_frames is None:
        max_frames = int(np.ceil(plan.n_steps * dt * fps)) + 2
    q0s = np.ascontiguousarray(q0.T, dtype=np.float64)
    qd0s = None if qd0 is None else np.ascontiguousarray(qd0.T, dtype=np.float64)
    frames = np.zeros((max_frames, 3 * njv + 18, n))
    fsteps = np.full(max_frames, -1, dtype=np.int32)
    nfr = np.zeros(1, dtype=np.int32)
    final = np.zeros((3 * njv, n))
    Kc, ph = np.ascontiguousarray(K, dtype=np.float64), np.ascontiguousarray(phi, dtype=np.float64)
    co, di, of = (np.ascontiguousarray(a, dtype=np.float64) for a in (plan.coeffs, plan.displacement, plan.pos_offset))
    rc = lib().h_closed_loop_f64(path, _p(fp), _p(gp), nj, _p(co), _p(di), _p(of), C.c_double(plan.timestep), C.c_double(plan.init_step),
                                 C.c_int(plan.n_steps), _p(Kc), _p(ph), C.c_double(dt), C.c_double(fps), C.c_double(div), _p(q0s), _p(qd0s), _p(frames),
                                 C.c_int(max_frames), _p(fsteps), _p(nfr), _p(final), C.c_int64(n))
    assert rc == 0
    nf = min(int(nfr[0]), max_frames)
    return dict(frames=frames[:nf], frame_steps=fsteps[:nf], final=final)


def regressor_gram(analysis, q, qd, qdd, f):
    """q, qd, qdd (n, nj), f (n, 6) -> the 112-double pack through gram_accumulate / gram_pack_entry (csrc/rbm_gram.cuh) on the host."""
    path, fp, gp, nj = _model_args(analysis)
    n = q.shape[0]
    arrs = [np.ascontiguousarray(a.T, dtype=np.float64) for a in (q, qd, qdd, f)]
    pack = np.zeros(112)
    rc = lib().h_regressor_gram_f64(path, _p(fp), _p(gp), nj, *[_p(a) for a in arrs], _p(pack), C.c_int64(n))
    assert rc == 0
    return pack


# ---- csrc/rbm_setup.cuh on the host (the batched frame-algebra helpers) ------------------------------------------------------------
def _c64(a, shape_tail):
    a = np.ascontiguousarray(a, dtype=np.float64)
    assert tuple(a.shape[1:]) == tuple(shape_tail), (a.shape, shape_tail)
    return a


def transfer_simat(poses_Rt, simats, adjoint_form=False):
    P, G = _c64(poses_Rt, (12,)), _c64(simats, (6, 6))
    out = np.empty_like(G)
    assert lib().h_transfer_simat_f64(_p(P), _p(G), _p(out), C.c_int64(len(P)), C.c_int(1 if adjoint_form else 0)) == 0
    return out


def coordinate_transfer_imat(poses_Rt, imats, mass):
    P, I, m = _c64(poses_Rt, (12,)), _c64(imats, (3, 3)), _c64(np.atleast_1d(mass), ())
    out = np.empty_like(I)
    assert lib().h_transfer_imat_f64(_p(P), _p(I), _p(m), _p(out), C.c_int64(len(P))) == 0
    return out


def spatial_inertia(mass, diag):
    m, d = _c64(np.atleast_1d(mass), ()), _c64(diag, (3,))
    out = np.empty((len(m), 6, 6))
    assert lib().h_spatial_inertia_f64(_p(m), _p(d), _p(out), C.c_int64(len(m))) == 0
    return out


def compose_poses(trans, rot):
    rot = np.ascontiguousarray(rot, dtype=np.float64)
    t = _c64(trans, (3,))
    out, status = np.empty((len(t), 12)), np.empty(len(t), dtype=np.int32)
    assert lib().h_compose_f64(_p(t), _p(rot), C.c_int(rot.shape[1]), _p(out), _p(status), C.c_int64(len(t))) == 0
    return out, status


def point_motion(twists, dtwists, points, want_acc=True):
    tw, p = _c64(twists, (6,)), _c64(points, (3,))
    dtw = _c64(dtwists, (6,)) if (want_acc and dtwists is not None) else None
    lv = np.empty_like(p)
    la = np.empty_like(p) if dtw is not None else None
    assert lib().h_point_motion_f64(_p(tw), _p(dtw), _p(p), _p(lv), _p(la), C.c_int64(len(p))) == 0
    return lv, la


def regressor_rows(twists, dtwists):
    tw, dtw = _c64(twists, (6,)), _c64(dtwists, (6,))
    Y = np.empty((len(tw), 6, 10))
    assert lib().h_regressor_rows_f64(_p(tw), _p(dtw), _p(Y), C.c_int64(len(tw))) == 0
    return Y
