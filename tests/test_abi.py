"""`not gpu`: the C-ABI library builds, loads and exports every symbol include/rbm_b200.h declares; the ctypes
signature table covers exactly those symbols; argument validation that needs no device works; and the product package
never reaches into oracle/."""
import ctypes as C
import os
import re

from conftest import ROOT


def test_library_exports_every_declared_symbol():
    from rigid_body_manipulation_b200 import _lib

    lib = _lib.load()
    declared = _lib.header_symbols()
    assert len(declared) >= 28 and len(set(declared)) == len(declared)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in rbm_b200.h but not exported"
    assert set(declared) == set(_lib.SIGNATURES), set(declared) ^ set(_lib.SIGNATURES)
    assert lib.rbm_version().decode().startswith("rbm_b200")


def test_header_cites_the_reference_for_every_compute_entry_point():
    text = open(os.path.join(ROOT, "include", "rbm_b200.h")).read()
    for fn in ["dynamics/dynamics.py:109-157", "dynamics/dynamics.py:72-106", "dynamics/dynamics.py:215-249", "core/simulate.py:202-209",
               "loggers/loggers.py:127-129", "dynamics/dynamics.py:41-46", "transformations/transformations.py:8-50", "dynamics.py:160-212"]:
        assert fn in text, fn


def test_argument_validation_without_a_device():
    from rigid_body_manipulation_b200 import _lib

    lib = _lib.load()
    out = C.c_void_p()
    assert lib.rbm_model_create(0, None, None, None, None, None, None, None, None, 0, 0, C.byref(out)) == _lib.RBM_ERR_INVALID
    assert b"nj" in lib.rbm_last_error_string()
    assert lib.rbm_model_create(17, None, None, None, None, None, None, None, None, 0, 0, C.byref(out)) == _lib.RBM_ERR_UNSUPPORTED
    assert lib.rbm_rnea_f64(None, None, None, None, None, None, None, 4, 4, None) == _lib.RBM_ERR_INVALID
    assert lib.rbm_compose_f64(None, None, 5, None, None, 1, None) == _lib.RBM_ERR_INVALID
    assert lib.rbm_model_num_joints(None) == _lib.RBM_ERR_INVALID


def test_no_silent_cpu_path():
    """Without a CUDA device model creation must fail loudly (RbmCudaError), never compute on the host."""
    import numpy as np
    import pytest
    import torch

    from rigid_body_manipulation_b200 import _lib, model
    from rigid_body_manipulation_b200.engine import Model

    if torch.cuda.is_available():
        pytest.skip("a device is present")
    c = model.load_packaged("sequential", "hammer")
    with pytest.raises(_lib.RbmCudaError):
        Model(c.hposes_Rt, c.simats, c.uscrews, c.twist_0, c.dtwist_0)
    assert _lib.load().rbm_device_count() == 0
    assert np.isfinite(c.simats).all()


def test_product_never_imports_the_oracle():
    pat = re.compile(r"^\s*(from|import)\s+oracle\b", re.M)
    # the package, and everything else that is not tests / smoke / the bench's CPU arm: tools, examples
    for top in ("rigid_body_manipulation_b200", "tools", "examples"):
      for dirpath, _, files in os.walk(os.path.join(ROOT, top)):
        if "_build" in dirpath:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not pat.search(src), f"{f} imports the oracle"
                assert "/root/reference" not in src or f in ("extract_reference_assets.py",), f


def test_nccl_less_machine_reports_unavailable_instead_of_crashing():
    """with libnccl absent rbm_nccl_available() must return 0 (include/rbm_b200.h) and leave a message -- it used to
    call dlerror() twice and build a std::string from NULL.  Fresh process: the binding is resolved once per process."""
    import subprocess
    import sys

    code = (
        "import ctypes, sys\n"
        "from rigid_body_manipulation_b200 import _lib\n"
        "lib = _lib.load()\n"
        "ok = lib.rbm_nccl_available()\n"
        "msg = _lib.last_error()\n"
        "buf = (ctypes.c_ubyte * 128)()\n"
        "rc = lib.rbm_nccl_unique_id(buf)\n"
        "print(ok, rc, msg)\n"
        "sys.exit(0 if (ok == 0 and rc == _lib.RBM_ERR_NCCL and 'dlopen' in msg) else 1)\n"
    )
    env = dict(os.environ, RBM_NCCL_LIB="/nonexistent/libnccl-not-here.so.2")
    r = subprocess.run([sys.executable, "-c", code], cwd=ROOT, env=env, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
