"""ORACLE / TEST INFRASTRUCTURE ONLY.  Import-time stub of the `mujoco` names that the
reference's `dynamics/dynamics.py:7-8`, `transformations/poses.py:2` and `utilities.py:3-4`
import at module top, so those files can be executed UNMODIFIED from /root/reference in a
container without MuJoCo.  Nothing here computes anything: every callable raises."""
