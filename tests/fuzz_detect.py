"""Fuzz the detection path against the oracle: random lenses, frame sizes and scenes -- on the CUDA-on-CPU build (default) or, with
--gpu, on the real library (larger frames, row lengths that are multiples of 16 so that the TMA scan / TMA windows run; both the
one-shot and the overlapped call).  Development tool: python tests/fuzz_detect.py [iterations] [seed] [--gpu]; it uses the oracle, so it
lives under tests/."""
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))
sys.path.insert(0, os.path.join(REPO, "tests", "emu"))
import build_emu                                    # noqa: E402
from mocapv2_b200.engine import CaptureEngine      # noqa: E402
from oracle import restate as R                     # noqa: E402
from util import oracle_contour_table               # noqa: E402


def main():
    argv = [a for a in sys.argv[1:] if a != "--gpu"]
    gpu = "--gpu" in sys.argv
    iters = int(argv[0]) if len(argv) > 0 else 40
    rng = np.random.default_rng(int(argv[1]) if len(argv) > 1 else 1)
    if gpu:
        eng = CaptureEngine("cuda:0")
    else:
        from emu_engine import EmuEngine
        eng = EmuEngine(build_emu.build())
    reasons = {}
    for it in range(iters):
        if gpu:
            H, W = int(rng.integers(64, 900)), int(rng.integers(64, 1200))
        else:
            H, W = int(rng.integers(40, 300)), int(rng.integers(40, 360))
        if gpu or rng.random() < 0.5:
            W = max(48, (W // 16) * 16)
        f = float(rng.uniform(0.45, 6.0)) * max(H, W)
        K = np.array([[f, 0, W * rng.uniform(0.3, 0.7)], [0, f * rng.uniform(0.95, 1.05), H * rng.uniform(0.3, 0.7)], [0, 0, 1]])
        D = np.array([rng.uniform(-0.4, 0.4), rng.uniform(-0.1, 0.1), rng.uniform(-0.01, 0.01), rng.uniform(-0.01, 0.01), rng.uniform(-0.05, 0.05)])
        bg = int(rng.choice([0, 40, 200, 216]))
        img = rng.integers(0, bg + 1, (H, W)).astype(np.uint8)
        yy, xx = np.mgrid[:H, :W]
        for _ in range(int(rng.integers(0, 14))):
            cx, cy = int(rng.integers(0, W)), int(rng.integers(0, H))
            kind = rng.integers(0, 6)
            if kind == 0:
                r = int(rng.integers(2, 30)); img[(xx - cx) ** 2 + (yy - cy) ** 2 <= r * r] = 255
            elif kind == 1:
                r = int(rng.integers(6, 30)); d2 = (xx - cx) ** 2 + (yy - cy) ** 2
                img[(d2 <= r * r) & (d2 >= (r // 2) ** 2)] = 255                      # ring: a hole border
            elif kind == 2:
                w, h = int(rng.integers(1, 120)), int(rng.integers(1, 90)); img[cy:cy + h, cx:cx + w] = 255
            elif kind == 3:
                img[max(cy - 1, 0):cy + 2, max(cx - 1, 0):cx + 2] = int(rng.integers(217, 256))
            elif kind == 4:                                                          # flat ellipse, several trace windows wide
                ax, ay = float(rng.uniform(20, 140)), float(rng.uniform(4, 30))
                img[((xx - cx) / ax) ** 2 + ((yy - cy) / ay) ** 2 <= 1.0] = 255
            else:                                                                    # sheared diamond: long diagonal runs
                ax, ay, sk = float(rng.uniform(20, 120)), float(rng.uniform(5, 40)), float(rng.uniform(-0.5, 0.5))
                img[np.abs((xx - cx) / ax) + np.abs((yy - cy - sk * (xx - cx)) / ay) <= 1.0] = 255
        min_area = float(rng.choice([0.0, 30.0, 500.0]))
        _, binimg = R.filter_frame(img, K, D)
        _, pts = oracle_contour_table(binimg, min_area)
        res = eng.detect(torch.from_numpy(img[None].copy()).to(eng.device), K, D, min_area=min_area)
        if gpu:                                                  # the overlapped call on the same frame (three copies, two chunks)
            over = eng.detect_pipelined(torch.from_numpy(np.stack([img] * 3)).to(eng.device), K, D, min_area=min_area, chunk_frames=2)
            for k in range(3):
                if over.points(k) != res.points(0) or int(over.flags[k]) != int(res.flags[0]):
                    np.savez("/tmp/fuzz_fail.npz", img=img, K=K, D=D, min_area=min_area)
                    print(f"MISMATCH one-shot vs overlapped at iteration {it}, copy {k}")
                    return 1
        fl = int(res.flags[0])
        reasons[fl >> 8 if fl & 64 else -1] = reasons.get(fl >> 8 if fl & 64 else -1, 0) + 1
        ok = res.points(0) == (pts if pts else [[None, None]]) and (fl & 63 & ~16) == 0
        if not ok:
            np.savez("/tmp/fuzz_fail.npz", img=img, K=K, D=D, min_area=min_area)
            print(f"MISMATCH at iteration {it}: flags {fl}, {len(pts)} oracle points vs {res.points(0)[:3]}... saved /tmp/fuzz_fail.npz")
            return 1
    print(f"{iters} scenes equal to the oracle; path taken (-1 = cluster path, else general-path reason): {reasons}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
