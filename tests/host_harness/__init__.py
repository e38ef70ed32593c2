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
    deps = [SRC] + [os.path.join(CSRC, f) for f in ("rbm_rnea.cuh", "rbm_typed.cuh", "rbm_trig.cuh", "rbm_model.cuh")]
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
