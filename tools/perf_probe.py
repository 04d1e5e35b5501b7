"""Quick stage timing probe on one GPU (development tool; bench.py is the contract benchmark)."""
import argparse
import sys
import os
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mocapv2_b200 import synth as S
from mocapv2_b200.engine import CaptureEngine


def timed(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c4")
    ap.add_argument("--frames", type=int, default=256)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--geom", type=int, default=1)
    a = ap.parse_args()
    eng = CaptureEngine("cuda:0")
    dev = eng.device
    H, W, M = {"c1": (480, 640, 4), "c3": (1080, 1440, 32), "c4": (2048, 2048, 128)}[a.config]
    n = a.frames
    g = torch.Generator().manual_seed(1)
    centres = torch.stack([torch.randint(40, W - 40, (n, M), generator=g), torch.randint(40, H - 40, (n, M), generator=g)], dim=-1)
    ridx = torch.randint(0, 9, (n, M), generator=g)
    frames = torch.empty((n, H, W), dtype=torch.uint8, device=dev)
    for i in range(0, n, 32):
        S.render_batch_torch(H, W, centres[i:i + 32].to(dev), ridx[i:i + 32].to(dev), 100 + i, dev, out=frames[i:i + 32])
    K, D = S.SHIPPED_K, S.SHIPPED_DIST
    res = eng.detect(frames, K, D)
    torch.cuda.synchronize()
    print(f"{a.config}: {n} frames {W}x{H}, blobs/frame mean {res.count.float().mean().item():.1f}, flags {int(res.flags.max())}")
    t_copy = timed(lambda: frames.clone(), a.reps)
    t_det = timed(lambda: eng.detect(frames, K, D, out=res), a.reps)
    t_fil = timed(lambda: eng.filter(frames, K, D), a.reps)
    gb = n * H * W / 1e9
    print(f"clone      {t_copy:8.3f} ms  ({2 * gb / t_copy * 1e3:.0f} GB/s r+w)")
    print(f"detect     {t_det:8.3f} ms  {n / t_det * 1e3:.0f} frames/s  {gb / t_det * 1e3:.0f} GB/s algorithmic  ({t_det / n * 1e3:.2f} us/frame)")
    print(f"filter+mat {t_fil:8.3f} ms")
    if a.geom:
        rig = S.config_rig("c5")
        cams = eng.cameras(rig["poses"], rig["camera_params"])
        for P in (1_000_000, 10_000_000):
            pts = (torch.rand((P, 8, 2), device=dev) * 2000).contiguous()
            xyz = torch.empty((P, 3), device=dev)
            err = torch.empty((P,), device=dev)
            t = timed(lambda: eng.triangulate(pts, cams, xyz=xyz, err=err), a.reps)
            print(f"triangulate fp32 P={P}: {t:.3f} ms  {P / t * 1e3 / 1e9:.3f} Gpts/s  ~{2729 * P / t * 1e3 / 1e12:.2f} TFLOP/s (F=2729/pt)")
        pts64 = pts[:1_000_000].double().contiguous()
        t = timed(lambda: eng.triangulate(pts64, cams), a.reps)
        print(f"triangulate fp64 P=1e6: {t:.3f} ms")


if __name__ == "__main__":
    main()
