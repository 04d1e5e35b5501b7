"""CPU baseline arm: the capture path as the reference runs it on a CPU, i.e. through OpenCV / SciPy.

TEST INFRASTRUCTURE ONLY (bench.py's cpu_baseline leg and `--impl reference`).  The reference is pure Python over
third-party wheels, so there is no oracle/_ref to compile and /root/reference does not exist on the GPU box; this
module issues the same library calls in the same order as lib/ImageOperations.py:33-65 (cv.undistort -> box blur ->
cv.threshold -> cv.medianBlur -> cv.findContours -> contourArea / arcLength filter -> cv.moments) and, for the
geometry, uses oracle.restate (numpy SVD + the projectPoints restatement) with the reference's control flow
(lib/Helpers.py:178-280).  One substitution, as in oracle/gen_golden.py: the numba-CUDA blur of the reference
(lib/CudaOperations.py:24-41, needs a GPU) is the integer restatement oracle.restate.blur5_floor, bit-identical to it.
tests/test_oracle_cv2.py checks this port against the restated oracle frame by frame.
"""
import numpy as np

from . import restate as R

try:
    import cv2 as cv
except ImportError:                      # the GPU box image ships opencv-python-headless; keep the import soft anyway
    cv = None


def available():
    return cv is not None


def find_dot(img, K, dist):
    """_find_dot(img)[1] with OpenCV doing what it does in the reference."""
    K = np.asarray(K, dtype=np.float64)
    dist = np.asarray(dist, dtype=np.float64)
    und = cv.undistort(img, K, dist)
    grey = R.blur5_floor(und)
    grey = cv.threshold(grey, 255 * 0.85, 255, cv.THRESH_BINARY)[1]
    grey = cv.medianBlur(grey, 5)
    found, _ = cv.findContours(grey, cv.RETR_TREE, cv.CHAIN_APPROX_SIMPLE)
    pts = []
    for cnt in found:
        area = cv.contourArea(cnt)
        per = cv.arcLength(cnt, True)
        if not per or not (4 * np.pi * area / (per * per) > R.MIN_CIRC and area > R.MIN_AREA):
            continue
        m = cv.moments(cnt)
        if m["m00"] != 0:
            pts.append([int(m["m10"] / m["m00"]), int(m["m01"] / m["m00"])])
    return pts if pts else [[None, None]]


def frame_set(frames, rig, obj_count, max_cand=None, max_groups=None, pool=None):
    """One synchronized frame-set (C, H, W) through detect -> match -> triangulate; returns (points per camera, objects)."""
    K = rig["camera_params"][0]["intrinsic_matrix"]
    dist = rig["camera_params"][0]["distortion_coef"]
    if pool is not None:
        pts = list(pool.map(lambda f: find_dot(f, K, dist), frames))
    else:
        pts = [find_dot(f, K, dist) for f in frames]
    obj, ipa = R.correspond(pts, rig["poses"], rig["camera_params"], rig["Fs"], obj_count, max_cand=max_cand, max_groups=max_groups)
    return pts, obj, ipa
