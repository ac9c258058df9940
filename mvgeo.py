"""Importable alias of the package directory `2025_icra_multi_view_robot_pose_estimation_b200`
(a Python identifier cannot start with a digit): `import mvgeo` gives that package."""
import importlib
import sys

sys.modules[__name__] = importlib.import_module("2025_icra_multi_view_robot_pose_estimation_b200")
