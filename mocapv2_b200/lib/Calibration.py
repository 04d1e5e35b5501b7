"""Fundamental matrices for the epipolar matcher -- the producer side of ./jsons/fundamentals.json.

Mirrors (host side, 3x3 FP64 algebra once per calibration -- nothing here is worth a kernel):
  poses_to_fundamental_matrix(pose1, pose2, K1=None, K2=None)      CalculateCameraPoses.py:26-78
  the Fs list written to ./jsons/fundamentals.json                 CalculateCameraPoses.py:189-191, 236-240

The reference only ever writes the camera0 -> camera1 matrix (twice, :190-191), which is why its shipped JSON cannot
serve a rig with more than two cameras; `fundamentals_from_poses` produces what the matcher actually indexes --
Fs[i-1] maps a camera-0 point to its epiline in camera i (lib/Helpers.py:206-207) -- for every camera of the rig.
"""
from __future__ import annotations

import json

import numpy as np


def poses_to_fundamental_matrix(pose1, pose2, K1=None, K2=None):
    """F with x2^T F x1 = 0 in pixel coordinates, or the essential matrix when K1/K2 are not given.
    Same argument meaning as the reference: poses are {"R": (3,3), "t": (3,)|(3,1)} world->camera; F is NOT rescaled."""
    R1 = np.asarray(pose1["R"], dtype=np.float64)
    R2 = np.asarray(pose2["R"], dtype=np.float64)
    t1 = np.asarray(pose1["t"], dtype=np.float64).reshape(3)
    t2 = np.asarray(pose2["t"], dtype=np.float64).reshape(3)
    R_rel = R2 @ R1.T                                   # camera 1 frame -> world -> camera 2 frame
    t_rel = t2 - R_rel @ t1
    E = np.cross(t_rel[:, None], R_rel, axis=0)         # [t]x R, column by column
    if K1 is None or K2 is None:
        return E
    K1 = np.asarray(K1, dtype=np.float64)
    K2 = np.asarray(K2, dtype=np.float64)
    return np.linalg.inv(K2).T @ E @ np.linalg.inv(K1)


def fundamentals_from_poses(camera_poses, camera_params):
    """[F(cam0 -> cam i) for i = 1..C-1] as nested lists -- the value `lib.Helpers.Fs` must hold for a C-camera rig."""
    K = [np.asarray(p["intrinsic_matrix"], dtype=np.float64) for p in camera_params]
    if len(K) < len(camera_poses):
        raise IndexError("list index out of range")     # the reference would fail the same way on camera_params[i]
    return [poses_to_fundamental_matrix(camera_poses[0], camera_poses[i], K[0], K[i]).tolist()
            for i in range(1, len(camera_poses))]


def write_fundamentals(camera_poses, camera_params, path="./jsons/fundamentals.json"):
    """Write the Fs list where read_fundamental_matrix() (lib/Helpers.py:22-28) looks for it.  A two-camera rig gets the
    matrix twice, like the file the reference ships."""
    Fs = fundamentals_from_poses(camera_poses, camera_params)
    if len(Fs) == 1:
        Fs = Fs + Fs
    with open(path, "w") as outfile:
        json.dump(Fs, outfile)
    return Fs
