"""ORACLE / TEST INFRASTRUCTURE ONLY.  Import-time stub of the `mujoco` names that the
reference's `dynamics/dynamics.py:7-8`, `transformations/poses.py:2` and `utilities.py:3-4`
import at module top, so those files can be executed UNMODIFIED from /root/reference in a
container without MuJoCo.  The structs / enums are names only.  The functions in `_functions.py` raise -- unless they are handed the
functional stand-in model of oracle/mujoco_standin.py (used to execute the reference's closed loop, oracle/gen_golden_simulate.py)."""
