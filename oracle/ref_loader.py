"""
oracle/ref_loader.py — import the UNMODIFIED reference functions from /root/reference.

TEST INFRASTRUCTURE ONLY. /root/reference exists only in the build container, never on the
GPU box, so nothing under `-m gpu`, smoke() or bench.py may call this. It is used by
tests/golden/make_golden.py (to freeze golden vectors) and by CPU tests that skip when the
checkout is absent.

The reference has no package: functions live in scripts and notebook cells. Nothing is imported or
run wholesale: only the `def` blocks of the named functions are extracted from a script / notebook cell
and exec'd into a namespace holding exactly the modules they need (numpy, math, cv2, scipy Rotation,
torch), so no module-level reference code ever executes.
"""
from __future__ import annotations

import importlib.machinery
import importlib.util
import json
import math
import os
import sys
import types

REF_ROOT = os.environ.get("MVGEO_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "model", "MvRoPose_FR3.py"))


def _cell_source(nb_rel: str, cell: int) -> str:
    with open(os.path.join(REF_ROOT, nb_rel)) as f:
        nb = json.load(f)
    return "".join(nb["cells"][cell]["source"])


def _extract_defs(src: str, names) -> str:
    """Keep only the top-level `def name(...)` / `class name` blocks asked for, verbatim."""
    lines = src.splitlines(keepends=True)
    out, keep = [], False
    for line in lines:
        top = line and not line[0].isspace() and line.strip() != ""
        if top:
            keep = any(line.startswith(f"def {n}(") or line.startswith(f"class {n}") for n in names)
        if keep:
            out.append(line)
    return "".join(out)


def _base_ns() -> dict:
    import cv2
    import numpy as np
    import torch
    import torch.nn.functional as F
    from scipy.spatial.transform import Rotation as R

    return dict(np=np, math=math, cv2=cv2, R=R, torch=torch, F=F, json=json, os=os)


def load_fr3():
    """model/MvRoPose_FR3.py: create_gt_heatmap, get_modified_dh_matrix,
    angle_to_joint_coordinate, joint_coordinate_to_pixel_plane."""
    src = open(os.path.join(REF_ROOT, "model", "MvRoPose_FR3.py")).read()
    ns = _base_ns()
    exec(_extract_defs(src, ["create_gt_heatmap", "get_modified_dh_matrix", "angle_to_joint_coordinate",
                             "joint_coordinate_to_pixel_plane"]), ns)
    return ns


def load_fr5():
    """model/Fr5_model_train.ipynb cell 2 (FK + projection) and cell 14 (decoder, estimate_camera_pose)."""
    ns = _base_ns()
    exec(_extract_defs(_cell_source("model/Fr5_model_train.ipynb", 2),
                       ["get_dh_matrix", "angle_to_joint_coordinate", "joint_coordinate_to_pixel_plane"]), ns)
    exec(_extract_defs(_cell_source("model/Fr5_model_train.ipynb", 14),
                       ["extract_keypoints_from_heatmaps", "estimate_camera_pose"]), ns)
    return ns


def load_meca500():
    """visualization/Meca500_vis.ipynb cell 0: get_dh_matrix, forward_kinematics, project_to_pixel."""
    ns = _base_ns()
    exec(_extract_defs(_cell_source("visualization/Meca500_vis.ipynb", 0),
                       ["get_dh_matrix", "forward_kinematics", "project_to_pixel"]), ns)
    return ns


def load_mv_model():
    """model/MV-model.ipynb cell 6: ForwardKinematics, project_3d_to_2d, robot_pose_loss."""
    ns = _base_ns()
    exec(_extract_defs(_cell_source("model/MV-model.ipynb", 6),
                       ["ForwardKinematics", "project_3d_to_2d", "robot_pose_loss"]), ns)
    return ns


def load_calib():
    """dataset/4_Calib_cam_save.py:35-59 load_fhd_calibration (ZED-X FHD1200 sections)."""
    import configparser

    src = open(os.path.join(REF_ROOT, "dataset", "4_Calib_cam_save.py")).read()
    ns = dict(configparser=configparser, os=os, json=json)
    exec(_extract_defs(src, ["load_fhd_calibration"]), ns)
    return ns
