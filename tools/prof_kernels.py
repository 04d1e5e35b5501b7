"""One launch of every hot kernel of the path inside a cudaProfilerStart/Stop window (development / evidence tool).

    python tools/prof_kernels.py                                   plain run (must exit 0 before any profiler run)
    ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/prof python tools/prof_kernels.py

Workload = bench.py's (BASELINE config 4: 16 cameras 2048x2048, 128 markers, 64 frame-sets = 1024 frames), then the
config-5 shape (8-view triangulation) and the Bayer front step.  Nothing here is a benchmark number.
"""
import argparse
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import bench as B  # noqa: E402
from mocapv2_b200 import synth as S  # noqa: E402
from mocapv2_b200.engine import CaptureEngine  # noqa: E402
from mocapv2_b200.pipeline import CapturePipeline  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frame-sets", type=int, default=64)
    ap.add_argument("--points", type=int, default=4_000_000)
    ap.add_argument("--bayer-frames", type=int, default=256)
    ap.add_argument("--skip", default="", help="comma list of: detect,geometry,bayer,overlapped")
    a = ap.parse_args()
    skip = set(a.skip.split(",")) if a.skip else set()
    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    eng = CaptureEngine(dev)
    FS = a.frame_sets
    rig, cen, ridx = B.make_scene(FS)
    pipe = CapturePipeline(eng, rig, max_blobs=B.MAX_BLOBS, obj_count=B.N_MARKERS, max_groups=B.MAX_GROUPS)
    frames = B.render_local(rig, cen, ridx, 0, pipe.cams_local, dev)
    H, W = rig["H"], rig["W"]
    corr = [None]

    flat = frames.view(-1, H, W)

    def step():
        det = pipe.detect(frames, pipelined=False)                 # the stages one after the other: every kernel profiled alone
        xy, count = pipe.exchange(det, FS)
        corr[0] = eng.correspond(xy, count, pipe.Fs, pipe.cams, obj_count=B.N_MARKERS, max_groups=B.MAX_GROUPS, out=corr[0])
        eng.scan_cells(flat, pipe.K0, pipe.dist0, variant=1)       # the TMA ring scan of the overlapped call, alone

    def overlapped():
        pipe.detect(frames, pipelined=True)                        # the product call of the bench: scan beside the other stages

    rig5 = S.config_rig("c5")
    cams5 = eng.cameras(rig5["poses"], rig5["camera_params"])
    pts5 = (torch.rand((a.points, 8, 2), device=dev) * 2000).floor().contiguous()
    xyz5 = torch.empty((a.points, 3), device=dev)
    err5 = torch.empty((a.points,), device=dev)
    pts64 = pts5[:200_000].double().contiguous()
    raw = torch.randint(0, 256, (a.bayer_frames, H, W), dtype=torch.uint8, device=dev)
    grey = torch.empty_like(raw)

    def geometry():
        eng.triangulate(pts5, cams5, xyz=xyz5, err=err5)
        eng.triangulate(pts64, cams5)

    def bayer():
        eng.bayer_gr2gray(raw, out=grey)
        eng.bayer_gr2gray_scan(raw, out=grey)                      # the front step that also delivers the detection's hot cell boxes

    for _ in range(3):
        step(); geometry(); bayer(); overlapped()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    if "detect" not in skip:
        step()
    if "overlapped" not in skip:
        overlapped()
    if "geometry" not in skip:
        geometry()
    if "bayer" not in skip:
        bayer()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print("prof_kernels ok:", int(corr[0].n_valid.sum()), "object points,", eng.launches, "launches")


if __name__ == "__main__":
    main()
