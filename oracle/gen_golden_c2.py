"""Generate tests/golden/c2_wide.npz: BASELINE config 2, a wide sample of the reference's videos/cam1-6.mp4.

TEST INFRASTRUCTURE ONLY.  Run from the repo root:  python oracle/gen_golden_c2.py
Like oracle/gen_golden.py this RUNS THE REFERENCE (lib.ImageOperations._find_dot imported from /root/reference, its numba blur
replaced by the integer restatement that gen_golden.py proves bit-identical under the numba simulator) and stores what it
returns, plus the cv2 contour statistics of its filtered image, for every 10th frame of every video and for all frames in
which the reference keeps a blob (25 of the 1200): 141 frames of 960x540.  Frames are stored as the host-decoded grey
images (cv2.VideoCapture -> cvtColor(BGR2GRAY)), i.e. exactly what the replay feeds _find_dot.
"""
import contextlib
import io
import json
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
OUT = os.path.join(REPO, "tests", "golden")
sys.path.insert(0, REPO)
from oracle import restate as R          # noqa: E402


def main():
    import cv2
    os.chdir(REF)
    sys.path.insert(0, REF)
    with contextlib.redirect_stdout(io.StringIO()):
        import lib.ImageOperations as IO
    IO.fast_cuda_blur = lambda image, kernel_size=5: R.blur5_floor(image)     # see gen_golden.py: proven bit-identical (meta.json)
    cp = json.load(open("jsons/camera-params-in.json"))
    K = np.array(cp[0]["intrinsic_matrix"]); D = np.array(cp[0]["distortion_coef"])
    frames, ids, points, tables = [], [], [], []
    kept = 0
    for cam in range(1, 7):
        cap = cv2.VideoCapture(f"videos/cam{cam}.mp4")
        k = 0
        while True:
            ok, fr = cap.read()
            if not ok:
                break
            g = cv2.cvtColor(fr, cv2.COLOR_BGR2GRAY)
            _, pts = IO._find_dot(g.copy())
            n = 0 if pts == [[None, None]] else len(pts)
            if n > 0 or k % 10 == 0:
                und = cv2.undistort(g, K, D)
                grey = IO.image_filter_gpu(und)
                cs, hier = cv2.findContours(grey, cv2.RETR_TREE, cv2.CHAIN_APPROX_SIMPLE)
                tab = []
                for j, c in enumerate(cs):
                    m = cv2.moments(c)
                    area = cv2.contourArea(c)
                    per = cv2.arcLength(c, True)
                    keep = bool(per and (4 * np.pi * area / (per * per) > 0.5 and area > 500))
                    tab.append([cv2.contourArea(c, oriented=True) * 2, m["m10"], m["m01"], m["m00"], per, int(hier[0][j][3]), int(keep)])
                frames.append(g); ids.append((cam, k)); points.append(pts)
                tables.append(np.array(tab, dtype=np.float64).reshape(-1, 7))
                kept += n > 0
            k += 1
    frames = np.stack(frames)
    counts = np.array([len(t) for t in tables])
    np.savez_compressed(os.path.join(OUT, "c2_wide.npz"), frames=frames, ids=np.array(ids), contour_counts=counts,
                        contours=np.concatenate(tables) if counts.sum() else np.zeros((0, 7)))
    json.dump({"points": points, "frames_with_kept_blob": int(kept), "cv2": cv2.__version__}, open(os.path.join(OUT, "c2_wide.json"), "w"))
    print(len(frames), "frames stored,", kept, "with a kept blob; contours per frame:", np.bincount(counts))


if __name__ == "__main__":
    main()
