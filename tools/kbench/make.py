"""Builds tools/kbench/_build/{kbench, consts.bin} (git-ignored; both travel to the GPU box with the snapshot).
    python tools/kbench/make.py && gpurun -- 'tools/kbench/_build/kbench tools/kbench/_build/consts.bin gram'"""
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from rigid_body_manipulation_b200 import build, model  # noqa: E402

out = os.path.join(HERE, "_build")
os.makedirs(out, exist_ok=True)
build.build_library()
c = model.load_packaged("sequential", sys.argv[1] if len(sys.argv) > 1 else "hammer")
blob = np.concatenate([np.asarray(a, float).reshape(-1) for a in (c.hposes_Rt, c.simats, c.uscrews, c.twist_0, c.dtwist_0, c.pose_sen_Rt)])
assert blob.size == 84 + 252 + 36 + 6 + 6 + 12
blob.tofile(os.path.join(out, "consts.bin"))
libdir = os.path.join(ROOT, "rigid_body_manipulation_b200")
subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O2", "-std=c++17", "-o", os.path.join(out, "kbench"),
                       os.path.join(HERE, "kbench.cu"), "-L" + libdir, "-lrbm_b200", "-Xlinker", "-rpath", "-Xlinker", "$ORIGIN/../../../rigid_body_manipulation_b200"])
print(os.path.join(out, "kbench"))
