"""ORACLE / TEST INFRASTRUCTURE ONLY.  gcc build of oracle/rnea_oracle.c -> oracle/_build/librnea_oracle.so and its
ctypes wrapper.  (The reference itself is pure Python: there is nothing under /root/reference to compile into
oracle/_ref, so no `_ref` artefact exists for this repo.)"""
from __future__ import annotations

import ctypes as C
import os
import shutil
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "rnea_oracle.c")
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "librnea_oracle.so")


def build(force: bool = False) -> str:
    os.makedirs(OUT_DIR, exist_ok=True)
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    gcc = shutil.which("gcc")
    if gcc is None:
        raise RuntimeError("gcc not found")
    subprocess.run([gcc, "-O2", "-fPIC", "-shared", "-fopenmp", "-std=c99", "-o", LIB, SRC, "-lm"], check=True)
    return LIB


_lib = None


def load():
    global _lib
    if _lib is None:
        lib = C.CDLL(build())
        vp = C.c_void_p
        lib.rnea_oracle_batch.restype = C.c_int
        lib.rnea_oracle_batch.argtypes = [C.c_int] + [vp] * 11 + [C.c_int64]
        _lib = lib
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def inverse_batched_c(traj, hposes_Rt, simats, uscrews, twist_0, dtwist_0, wrench_tip=None, pose_tip_Rt=None, want_twists=False):
    """traj (n, 3, nj) -> tau (n, nj) [, V_last (n, 6), dV_last (n, 6)] through the C restatement."""
    f = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.float64)
    traj, hposes_Rt, simats, uscrews, twist_0, dtwist_0 = map(f, (traj, hposes_Rt, simats, uscrews, twist_0, dtwist_0))
    wrench_tip, pose_tip_Rt = f(wrench_tip), f(pose_tip_Rt)
    n, _, nj = traj.shape
    tau = np.empty((n, nj))
    V = np.empty((n, 6)) if want_twists else None
    dV = np.empty((n, 6)) if want_twists else None
    rc = load().rnea_oracle_batch(nj, _p(hposes_Rt), _p(simats), _p(uscrews), _p(twist_0), _p(dtwist_0), _p(wrench_tip), _p(pose_tip_Rt),
                                  _p(traj), _p(tau), _p(V), _p(dV), n)
    if rc != 0:
        raise ValueError("rnea_oracle_batch failed (nj out of range)")
    return (tau, V, dV) if want_twists else tau
