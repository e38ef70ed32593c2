"""ORACLE / TEST INFRASTRUCTURE ONLY.  Generates tests/golden/ref_simulate_<target>.npz by EXECUTING THE REFERENCE'S OWN closed loop --
`core/simulate.py::simulate`, `controllers/lqr.py::LinearQuadraticRegulator`, `planners/joint_position_planner.py::JointPositionPlanner`,
`sensors/sensors.py`, `transformations/poses.py`, unmodified, from /root/reference -- on the functional MuJoCo stand-in of
oracle/mujoco_standin.py (this container only; neither the reference nor this script is needed on the GPU box).

    python -m oracle.gen_golden_simulate [target ...]        # default: hammer

Stored: the LQR gain the reference computed, the per-frame sensor twists / twist rates / (noise-perturbed, simulate.py:279-290) F/T
readings / regressors that simulate() returned, and the stand-in model's constants so that tests replay the same plant."""
import os
import sys
from pathlib import Path
from types import SimpleNamespace

import numpy as np

os.environ.setdefault("TQDM_DISABLE", "1")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")

BASE_YAML = dict(duration=3.0, timestep=-1.0, displacements=[0.2, 1.4, 0.6, 3.141592653589793, 0.0, 18.8495559215],  # configurations/base.yaml
                 input_gain=[10.0, 10.0, 10.0, 1e4, 1e4, 1e4], epsilon=1e-8, centered=True, fps=50)


def run_reference_simulation(target="hammer", duration=None, reference_root=None):
    """Returns (result dict of simulate(), controller, planner, stand-in model)."""
    from oracle import mujoco_standin as ms
    from oracle import reference_loader as rl

    root = reference_root or rl.REFERENCE_ROOT
    ref = rl.load_simulation()
    m = ms.StandinModel(os.path.join(root, "xml_models", "manipulators", "sequential.xml"),
                        os.path.join(root, "xml_models", "targets", target, "object_cad_gt.csv"))
    d = ms.StandinData(m)
    ms.mj_resetDataKeyframe(m, d, 0)                                                  # core/core.py:327
    pcfg = SimpleNamespace(duration=BASE_YAML["duration"] if duration is None else duration, timestep=BASE_YAML["timestep"],
                           pos_offset=d.qpos.copy().tolist(), displacements=list(BASE_YAML["displacements"]))
    planner = ref.planner.JointPositionPlanner(pcfg, m, d)                            # main.py: autoinstantiate(cfg.planner, m, d)
    ccfg = SimpleNamespace(state_space=SimpleNamespace(epsilon=BASE_YAML["epsilon"], centered=BASE_YAML["centered"]), input_gain=list(BASE_YAML["input_gain"]))
    controller = ref.controllers.LinearQuadraticRegulator(ccfg, m, d)
    logger = SimpleNamespace(fps=BASE_YAML["fps"], cam_id=0, complete_image_dir=Path("/nonexistent"), render=lambda d_, f: None)
    result = ref.simulate.simulate(m, d, logger, planner, controller)                 # main.py: simulate(m, d, logger, planner, controller)
    return result, controller, planner, m


def main(targets):
    for target in targets:
        res, controller, planner, m = run_reference_simulation(target)
        fr = res["frames"]
        out = os.path.join(GOLD, f"ref_simulate_{target}.npz")
        np.savez_compressed(
            out, gain_matrix=controller.gain_matrix, n_steps=planner.n_steps, timestep=planner.timestep, displacements=np.array(planner.displacements),
            twist_sen=np.array([f["twist_sen"] for f in fr]), dtwist_sen=np.array([f["dtwist_sen"] for f in fr]), ft_sen=np.array([f["ft_sen"] for f in fr]),
            pose_sen_obj=np.array(fr[0]["pose_sen_obj"]), regressors=np.asarray(res["regressors"]),
            hposes_Rt=m.consts["hposes_Rt"], simats=m.consts["simats"], uscrews=m.consts["uscrews"], twist_0=m.consts["twist_0"], dtwist_0=m.consts["dtwist_0"],
            pose_sen_llj=m.pose_sen_Rt, G_sensed=m.G_sensed, key_qpos=m.key_qpos)
        print(out, len(fr), "frames")


if __name__ == "__main__":
    main(sys.argv[1:] or ["hammer"])
