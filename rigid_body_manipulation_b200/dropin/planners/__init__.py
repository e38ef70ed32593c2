from .joint_position_planner import JointPositionPlanner, JointPositionPlannerConfig, traj_5th_spline  # noqa: F401
