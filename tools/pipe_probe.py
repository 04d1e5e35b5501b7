"""Overlapped detection A/B on one GPU (development probe; bench.py is the contract benchmark).

Workload = bench.py's C4 batch (64 frame-sets x 16 cameras of 2048x2048, 128 markers).  Times the one-shot call, the two scan
kernels alone and the chunked / overlapped call over chunk sizes, worker streams, co-resident filter CTAs and sync modes;
checks every variant's output against the one-shot call.
"""
import argparse
import json
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import bench as B  # noqa: E402
from mocapv2_b200.engine import CaptureEngine  # noqa: E402
from mocapv2_b200.pipeline import CapturePipeline  # noqa: E402


def timed(fn, reps, warm=3):
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
    ap.add_argument("--frame-sets", type=int, default=64)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--grid", default="", help="semicolon list of chunk,workers,filter_ctas,sync,scan,plan,prio tuples (overrides the built-in sweep)")
    ap.add_argument("--timelines", action="store_true", help="print the stage timeline of every variant")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    eng = CaptureEngine(dev)
    FS = a.frame_sets
    rig, cen, ridx = B.make_scene(FS)
    pipe = CapturePipeline(eng, rig, max_blobs=B.MAX_BLOBS, obj_count=B.N_MARKERS, max_groups=B.MAX_GROUPS)
    frames = B.render_local(rig, cen, ridx, 0, pipe.cams_local, dev).view(-1, rig["H"], rig["W"])
    n, H, W = frames.shape
    K, D = pipe.K0, pipe.dist0
    mb = B.MAX_BLOBS
    gb = n * H * W / 1e9
    ref = eng.detect(frames, K, D, max_blobs=mb)
    torch.cuda.synchronize()
    out = {"frames": n, "blobs_per_frame": float(ref.count.float().mean())}
    t = timed(lambda: eng.detect(frames, K, D, max_blobs=mb, out=ref), a.reps)
    out["one_shot_ms"] = t
    print(f"one-shot detect            {t:7.3f} ms   {gb / t * 1e3:6.0f} GB/s algorithmic", flush=True)
    for v in (0, 1):
        t = timed(lambda: eng.scan_cells(frames, K, D, variant=v), a.reps)
        out[f"scan_variant{v}_ms"] = t
        print(f"scan alone, variant {v}      {t:7.3f} ms   {gb / t * 1e3:6.0f} GB/s", flush=True)
    refxy, refc = ref.xy.clone(), ref.count.clone()
    rows = []
    # (chunk_frames, workers, filter_ctas, sync_mode, scan_variant, stream_plan, prio_mode)
    if a.quick:
        grid = [(128, 4, 5, 1, 1, 1, 2)]
    else:
        grid = [(256, 3, 0, 1, 1, 0, 0), (128, 3, 0, 1, 1, 0, 0)]
        grid += [(cf, w, fc, 1, 1, 1, pm) for cf in (64, 128, 256) for w in (3, 4, 5) for fc in (4, 5) for pm in (0, 1, 2)]
        grid += [(128, 4, 3, 1, 1, 1, 2), (128, 4, 6, 1, 1, 1, 2), (128, 4, 5, 0, 1, 1, 2), (128, 4, 5, 0, 0, 1, 2)]
    if a.grid:
        grid = [tuple(int(x) for x in g.split(",")) for g in a.grid.split(";") if g]
    engines = {}
    for (cf, w, fc, sm, sv, plan, pm, *rest) in grid:
        ss = rest[0] if rest else 6
        e = engines.get((w, pm))
        if e is None:
            e = CaptureEngine(dev)
            e._tables = eng._tables
            e.pipe_workers = w
            e.pipe_prio_mode = pm
            engines[(w, pm)] = e
        kw = dict(max_blobs=mb, chunk_frames=cf, sync_mode=sm, scan_variant=sv, filter_ctas_per_sm=fc, stream_plan=plan, scan_stages=ss)
        res = e.detect_pipelined(frames, K, D, **kw)
        torch.cuda.synchronize()
        okk = bool(torch.equal(res.count, refc)) and bool(torch.equal(res.xy[:, :1], refxy[:, :1]))
        full_ok = all(torch.equal(res.xy[i, :int(refc[i])], refxy[i, :int(refc[i])]) for i in range(0, n, 37))
        t = timed(lambda: e.detect_pipelined(frames, K, D, out=res, **kw), a.reps)
        rows.append({"chunk_frames": cf, "workers": w, "filter_ctas": fc, "sync_mode": sm, "scan_variant": sv, "stream_plan": plan, "prio_mode": pm, "scan_stages": ss,
                     "ms": t, "same": okk and full_ok, "info": e.last_pipe_info})
        print(f"chunk {cf:4d} workers {w} filter_ctas {fc} sync {sm} scan {sv} plan {plan} prio {pm} ring {ss}: {t:7.3f} ms  {gb / t * 1e3:6.0f} GB/s  same={okk and full_ok}", flush=True)
        if a.timelines:
            e.detect_pipelined(frames, K, D, out=res, timeline=True, **kw)
            torch.cuda.synchronize()
            tl = e.pipe_timeline()
            rows[-1]["timeline_ms"] = tl
            print("    scan done %.3f join %.3f | per chunk: seen, +group, +filter, +borders(end)" % (tl["scan_done"], tl["join"]))
            for c, r in enumerate(tl["chunks"]):
                print(f"    {c:2d}: {r[0]:.3f}  +{r[1] - r[0]:.3f}  +{r[2] - r[1]:.3f}  +{r[3] - r[2]:.3f} ({r[3]:.3f})")
    out["variants"] = rows
    best = min(rows, key=lambda r: r["ms"])
    e = engines[(best["workers"], best["prio_mode"])]
    e.detect_pipelined(frames, K, D, max_blobs=mb, chunk_frames=best["chunk_frames"], sync_mode=best["sync_mode"], scan_variant=best["scan_variant"],
                       filter_ctas_per_sm=best["filter_ctas"], stream_plan=best["stream_plan"], scan_stages=best["scan_stages"], timeline=True)
    torch.cuda.synchronize()
    tl = e.pipe_timeline()
    out["best"] = best
    out["timeline_ms"] = tl
    print("best:", best)
    print("timeline (ms since fork): scan done %.3f  join %.3f" % (tl["scan_done"], tl["join"]))
    for c, row in enumerate(tl["chunks"]):
        print(f"  chunk {c:2d}: scan seen {row[0]:.3f}  grouped {row[1]:.3f}  filtered {row[2]:.3f}  borders {row[3]:.3f}")
    os.makedirs(os.path.join(REPO, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(REPO, "gpurun_out", "pipe_probe.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
