"""Pins the oracle's restatements against the third-party libraries the reference delegates to (OpenCV, SciPy),
called live on random inputs -- the same calls the reference makes at lib/ImageOperations.py:38-65 and
lib/Helpers.py:77,133-139,207.  Skipped when cv2 is not importable.  CPU only."""
import numpy as np
import pytest

from mocapv2_b200 import synth as S
from oracle import restate as R

cv2 = pytest.importorskip("cv2")
K = np.array(S.SHIPPED_K)
D = np.array(S.SHIPPED_DIST)


def blobs_image(rng, H, W, n=6):
    img = rng.integers(0, 40, (H, W)).astype(np.uint8)
    for _ in range(n):
        c = (int(rng.integers(-10, W + 10)), int(rng.integers(-10, H + 10)))
        r = int(rng.integers(4, 36))
        cv2.circle(img, c, r, 255, -1 if rng.random() < 0.6 else int(rng.integers(4, 12)))
    return cv2.GaussianBlur(img, (0, 0), 1.2)


@pytest.mark.parametrize("shape", [(97, 131), (240, 320), (540, 960)])
def test_undistort_threshold_median(shape):
    rng = np.random.default_rng(shape[0])
    img = rng.integers(0, 256, shape).astype(np.uint8)
    assert np.array_equal(R.undistort(img, K, D), cv2.undistort(img, K, D))
    assert np.array_equal(R.threshold_bin(img), cv2.threshold(img, 255 * 0.85, 255, cv2.THRESH_BINARY)[1])
    b = (rng.random(shape) < 0.5).astype(np.uint8) * 255
    assert np.array_equal(R.majority5(b), cv2.medianBlur(b, 5))


def test_find_contours_order_tree_and_moments():
    rng = np.random.default_rng(5)
    for it in range(60):
        H, W = int(rng.integers(8, 120)), int(rng.integers(8, 160))
        if it % 2:
            b = cv2.medianBlur(((rng.random((H, W)) < rng.choice([0.3, 0.5, 0.7])) * 255).astype(np.uint8), 5)
        else:
            b = ((rng.random((H, W)) < rng.choice([0.1, 0.4, 0.6, 0.9])) * 255).astype(np.uint8)
        ref, hier = cv2.findContours(b.copy(), cv2.RETR_TREE, cv2.CHAIN_APPROX_SIMPLE)
        got, info = R.find_contours(b)
        assert len(got) == len(ref)
        for k, (g, c) in enumerate(zip(got, ref)):
            assert np.array_equal(g, c.reshape(-1, 2))                         # same vertices, same order, same start
            assert int(info[k][1]) == int(hier[0][k][3])                       # same parent
            a00, a10, a01, per = R.contour_stats(g)
            m = cv2.moments(c)
            assert abs(a00) * 0.5 == cv2.contourArea(c) == m["m00"]
            sgn = 1.0 if a00 > 0 else -1.0
            assert a10 * sgn * 0.16666666666666666 == m["m10"] and a01 * sgn * 0.16666666666666666 == m["m01"]
            assert per == cv2.arcLength(c, True)
        n, lab = R.label8(b)
        assert n == cv2.connectedComponents(b, connectivity=8)[0] - 1


def test_cv2_port_equals_restated_find_dot():
    from oracle import cv2_port
    rng = np.random.default_rng(9)
    for _ in range(8):
        img = blobs_image(rng, int(rng.integers(120, 400)), int(rng.integers(120, 500)))
        assert cv2_port.find_dot(img, K, D) == R.find_dot(img, K, D)


def test_epilines_and_project_points():
    rng = np.random.default_rng(3)
    F = np.array(S.SHIPPED_F)
    for _ in range(50):
        p = rng.integers(0, 2400, 2)
        ref = cv2.computeCorrespondEpilines(np.array([p], dtype=np.float32), 1, F)[0, 0].tolist()
        assert R.epiline_f32(p, F) == ref
    rig = S.config_rig("c5")
    X = S.config_markers("c5", rig, rng)[:20]
    for c, pose in enumerate(rig["poses"]):
        cam = rig["camera_params"][c]
        ref, _ = cv2.projectPoints(X.astype(np.float32), np.array(pose["R"]), np.array(pose["t"]), np.array(cam["intrinsic_matrix"]),
                                   np.array(cam["distortion_coef"]))
        got = np.array([R.project_point_f32(x, pose, cam) for x in X], dtype=np.float32)
        assert np.array_equal(got, ref.reshape(-1, 2))


def test_dlt_matches_scipy_svd():
    from scipy import linalg
    rig = S.config_rig("c3")
    rng = np.random.default_rng(4)
    Ps = R.projection_matrices(rig["poses"], rig["camera_params"])
    for _ in range(20):
        pts = rng.integers(100, 1300, (6, 2)).astype(np.float64)
        A = []
        for P, (x, y) in zip(Ps, pts):
            A.append(y * P[2, :] - P[1, :])
            A.append(P[0, :] - x * P[2, :])
        A = np.array(A)
        _, _, Vh = linalg.svd(A.T @ A, full_matrices=False)
        ref = Vh[3, 0:3] / Vh[3, 3]
        assert np.allclose(R.triangulate_point(pts, Ps), ref, rtol=1e-9, atol=1e-12)


@pytest.mark.parametrize("shape", [(3, 5), (4, 4), (20, 27), (21, 26), (480, 640)])
def test_bayer_front_step(shape):
    """RealtimeTracking_FLIR.py:103-104: cvtColor(BAYER_GR2BGR) then cvtColor(BGR2GRAY)."""
    rng = np.random.default_rng(shape[1])
    raw = rng.integers(0, 256, shape).astype(np.uint8)
    bgr = cv2.cvtColor(raw, cv2.COLOR_BAYER_GR2BGR)
    assert np.array_equal(R.bayer_gr_to_bgr(raw), bgr)
    assert np.array_equal(R.bayer_gr_to_gray(raw), cv2.cvtColor(bgr, cv2.COLOR_BGR2GRAY))
