"""mvgeo — B200-native geometry hot path for multi-view robot pose estimation:
belief-map decode -> DLT triangulation -> DH forward kinematics + reprojection loss.

The directory name starts with a digit, so import it as `import mvgeo` (alias module at the
repository root) or `importlib.import_module("2025_icra_multi_view_robot_pose_estimation_b200")`.
"""
from . import _lib
from ._lib import MvgeoError, LIB_PATH
from .robots import Chain, view_rotation, VIEW_EULER_ZYX_DEG
from .rig import CameraRig, ZEDX_FHD1200, load_conf_calibration, rodrigues
from . import ops, compat, sharding
from .graphed import GraphedPipeline
from .ops import (decode_heatmaps, triangulate, pnp_refine, pnp_solve, quat_mean, forward_kinematics, project_points, undistort_points, fk_reproj_loss,
                  encode_gaussian, heatmap_mse_loss, decode_and_mse, pipeline, HostPipeline, alloc_outputs, DecodeResult)

__version__ = "0.1.0"
