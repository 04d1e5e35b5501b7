#!/usr/bin/env python
"""Contract benchmark of the capture hot path (see BASELINE.json / SURVEY.md section 8d).

    python bench.py [--gpus N --steps K --warmup W]             this repo (libmocap_b200.so, sm_100a)
    python bench.py --impl reference [...]                       the reference's CPU path (OpenCV/NumPy) on the host cores

Workload (config.workload): BASELINE config 4 -- 16-camera 2048x2048 rig, 128 markers, full detect + match +
triangulate.  One step = one batch of F0*N synchronized frame-sets: every rank detects its 16/N cameras of all
frame-sets, ONE all-gather moves the centroid records, every rank matches + triangulates its F0 frame-sets.
Per-GPU work is constant in N (weak scaling).  value = frames/s over all ranks, inputs resident in HBM; e2e = the
same through the public API from pinned host memory (H2D of the frames and D2H of the results inside the timed region).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

from mocapv2_b200 import synth as S  # noqa: E402

CONFIG = "c4"
N_MARKERS = 128
MAX_BLOBS = 160
MAX_GROUPS = 64            # candidate groups evaluated per root (cap policy, flagged per frame-set)
MAX_CAND = 8               # MOCAP_MAX_CAND
JITTER = 0.01              # metres, per frame-set
SPREAD = 0.95              # half-extent (m) of the marker volume around the rig centre: fills the 2048x2048 views


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--frame-sets", type=int, default=64, help="frame-sets per GPU per step (F0)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--cpu-sample", type=int, default=1, help="frame-sets timed for cpu_baseline (0 = skip)")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------------------------------
# synthetic rig + frames (data plumbing)
# ---------------------------------------------------------------------------------------------------------------------
def make_scene(n_frame_sets, seed=S.SEED0 + 4000):
    """Marker positions per frame-set and their integer pixel centres per camera: (rig, centres [FS, C, M, 2] int64)."""
    rig = S.config_rig(CONFIG)
    rng = np.random.default_rng(seed)
    # markers spread over the whole commonly visible volume, >= 2r+12 px apart in every view where that is achievable
    # (with 16 views and 128 markers some views inevitably show touching blobs; the detector handles them)
    X0 = S.sample_markers(rig, N_MARKERS, rng, spread=SPREAD, min_sep_px=56.0, margin=40.0, tries=300)
    cen = np.empty((n_frame_sets, len(rig["poses"]), N_MARKERS, 2), dtype=np.int64)
    for s in range(n_frame_sets):
        X = X0 + rng.uniform(-JITTER, JITTER, X0.shape)
        cen[s] = np.rint(S.marker_pixels(rig, X)).astype(np.int64)
    radius_idx = rng.integers(0, 9, (n_frame_sets, len(rig["poses"]), N_MARKERS))
    return rig, cen, radius_idx


def render_local(rig, cen, radius_idx, cam_begin, cams_local, device):
    import torch
    FS = cen.shape[0]
    H, W = rig["H"], rig["W"]
    frames = torch.empty((FS, cams_local, H, W), dtype=torch.uint8, device=device)
    stamps = torch.from_numpy(S.disc_stamps()).to(device)
    for s in range(FS):
        c = torch.from_numpy(cen[s, cam_begin:cam_begin + cams_local]).to(device)
        r = torch.from_numpy(radius_idx[s, cam_begin:cam_begin + cams_local]).to(device)
        S.render_batch_torch(H, W, c, r, 1000 + s * 64 + cam_begin, device, stamps=stamps, out=frames[s])
    return frames


# ---------------------------------------------------------------------------------------------------------------------
# clocks (B200_PROFILING.md: sample nvidia-smi DURING the timed region)
# ---------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        self.index = index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50"], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            p = [x.strip() for x in line.split(",")]
            if len(p) < 7:
                continue
            try:
                sm.append(float(p[0]))
                mx.append(float(p[1]))
            except ValueError:
                continue
            for nm, v in zip(names, p[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        os.unlink(self.path)
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm))
        return out


# ---------------------------------------------------------------------------------------------------------------------
# CPU arm (reference's OpenCV/NumPy path on the host cores)
# ---------------------------------------------------------------------------------------------------------------------
def cpu_frame_sets(rig, host_frames, n_sets, threads):
    """Time the CPU path on n_sets frame-sets ([n, C, H, W] uint8 numpy).  Returns (seconds, frames, object points)."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import cv2_port
    if not cv2_port.available():
        return None
    pts_total = 0
    with ThreadPoolExecutor(max_workers=threads) as pool:
        t0 = time.perf_counter()
        for s in range(n_sets):
            _, obj, _ = cv2_port.frame_set(host_frames[s], rig, N_MARKERS, max_cand=MAX_CAND, max_groups=MAX_GROUPS, pool=pool)
            pts_total += len(obj)
        dt = time.perf_counter() - t0
    return dt, n_sets * host_frames.shape[1], pts_total


def host_render(rig, cen, radius_idx, n_sets):
    """CPU rendering of the same frames the device generator makes (bit-identical recipe is not required: both arms
    of a run consume their own copy of the same scene; the CPU arm uses the torch generator on the CPU device)."""
    import torch
    return render_local(rig, cen[:n_sets], radius_idx[:n_sets], 0, len(rig["poses"]), torch.device("cpu")).numpy()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    threads = os.cpu_count() or 1
    rig, cen, ridx = make_scene(1)
    frames = host_render(rig, cen, ridx, 1)
    from oracle import cv2_port
    if not cv2_port.available():
        emit({"impl": "reference", "unavailable": "opencv (cv2) is not importable on this host"})
        return 0
    try:
        import cv2
        cv2.setNumThreads(threads)
    except Exception:
        pass
    for _ in range(args.warmup):
        cpu_frame_sets(rig, frames, 1, threads)
    t = 0.0
    n_frames = 0
    pts = 0
    for _ in range(args.steps):
        dt, nf, np_ = cpu_frame_sets(rig, frames, 1, threads)
        t += dt
        n_frames += nf
        pts += np_
    value = n_frames / t
    line = {
        "impl": "reference", "metric": "frames/s (detect+match+triangulate, 16-camera 2048x2048 rig, 128 markers)",
        "value": value, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8 (detection) / f64 (geometry)", "data": "synthetic",
        "points_per_s": pts / t,
        "config": workload_config(1, 1, sample="1 frame-set (16 frames) per step"),
        "cpu_baseline": {"value": value, "unit": "frames/s", "cores": threads, "kind": "port",
                         "sample": "1 frame-set (16 frames 2048x2048) per step; OpenCV calls of lib/ImageOperations.py:33-65 "
                                   "(numba blur -> integer restatement) + lib/Helpers.py:178-280 control flow in NumPy with the "
                                   f"same candidate cap (max_cand {MAX_CAND}, max_groups {MAX_GROUPS}); frames over a {threads}-thread pool"},
        "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


def workload_config(n, f0, sample=None):
    cfg = {"workload": "BASELINE config 4: 16 cameras 2048x2048 u8, 128 markers, detect+match+triangulate",
           "cameras": 16, "frame": [2048, 2048], "markers": N_MARKERS, "frame_sets_per_gpu_per_step": f0,
           "frames_per_step": 16 * f0 * n, "parallelism": f"cameras sharded x{n} for detection, frame-sets sharded x{n} for geometry, 1 all-gather",
           "group_cap": {"max_cand": MAX_CAND, "max_groups": MAX_GROUPS},
           "l2": "inputs per step (>=1 GB) exceed the 126 MB L2; the same resident batch is re-read every step"}
    if sample:
        cfg["sample"] = sample
    return cfg


# ---------------------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist
    from mocapv2_b200.engine import CaptureEngine
    from mocapv2_b200.pipeline import CapturePipeline

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = torch.device(f"cuda:{local}")
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    N = world
    F0 = args.frame_sets
    FS = F0 * N
    rig, cen, ridx = make_scene(FS)
    eng = CaptureEngine(device)
    pipe = CapturePipeline(eng, rig, max_blobs=MAX_BLOBS, obj_count=N_MARKERS, max_groups=MAX_GROUPS)
    frames = render_local(rig, cen, ridx, pipe.cam_begin, pipe.cams_local, device)
    n_local = FS * pipe.cams_local
    H, W = rig["H"], rig["W"]
    corr_out = [None]

    def step(timer=None):
        det = pipe.detect(frames, timer=timer)
        xy, count = pipe.exchange(det, FS)
        corr_out[0] = eng.correspond(xy, count, pipe.Fs, pipe.cams, obj_count=N_MARKERS, max_groups=MAX_GROUPS, out=corr_out[0])
        return det, corr_out[0]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()                     # nvidia-smi needs ~100 ms to deliver its first sample: start before the warm-up
    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    timers = [eng.stage_timer() for _ in range(args.steps)]
    launches0 = eng.launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for k in range(args.steps):
        det, corr = step(timers[k])
    ev1.record()
    gpu_launches = eng.launches - launches0
    barrier()
    # nvidia-smi delivers a sample every ~50-100 ms and the timed region may be shorter than that: keep the same step loop
    # running (untimed) for about a second more so that the clock record is taken under this very load
    t_sus = time.perf_counter()
    while time.perf_counter() - t_sus < 1.0:
        for _ in range(5):
            step()
        torch.cuda.synchronize()
    clocks = sampler.stop() if rank == 0 else None
    if clocks is not None:
        clocks["window"] = "warm-up + timed region + 1 s of the same step loop (untimed)"
    barrier()
    ms = torch.tensor([ev0.elapsed_time(ev1)], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    stage_ms = {}
    for t in timers:
        for k, v in eng.stage_timer_read(t).items():
            stage_ms.setdefault(k, []).append(v)
    stage_avg = {k: float(np.mean(v)) for k, v in stage_ms.items()}

    # work accounting
    n_pts = torch.tensor([float(corr.n_valid.sum().item()), float((corr.flags & 1).sum().item()), float(det.count.sum().item())],
                         device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(n_pts)
    frames_per_step = 16 * F0 * N
    value = frames_per_step * args.steps / (ms_total * 1e-3)
    points_per_step = float(n_pts[0].item())

    # ---- end to end: pinned host frames -> H2D -> pipeline -> D2H of the results, every step ----------------------------
    e2e = None
    if not args.no_e2e:
        host = torch.empty(frames.shape, dtype=torch.uint8, pin_memory=True)
        host.copy_(frames)
        stage_dev = torch.empty_like(frames)
        host_obj = torch.empty(corr.obj.shape, dtype=corr.obj.dtype, pin_memory=True)
        host_nobj = torch.empty(corr.n_obj.shape, dtype=corr.n_obj.dtype, pin_memory=True)
        host_cnt = torch.empty(det.count.shape, dtype=det.count.dtype, pin_memory=True)
        copy_stream = torch.cuda.Stream(device)
        chunks = max(1, min(8, FS))
        bounds = np.linspace(0, FS, chunks + 1).astype(int)
        eng2 = CaptureEngine(device)
        eng2._tables = eng._tables
        dets = [None] * chunks

        def e2e_step():
            # chunked: the copy of chunk k+1 overlaps detection of chunk k (copy stream + compute stream)
            evs = []
            for c in range(chunks):
                with torch.cuda.stream(copy_stream):
                    stage_dev[bounds[c]:bounds[c + 1]].copy_(host[bounds[c]:bounds[c + 1]], non_blocking=True)
                    e = torch.cuda.Event()
                    e.record(copy_stream)
                    evs.append(e)
            xy_parts, cnt_parts = [], []
            for c in range(chunks):
                torch.cuda.current_stream().wait_event(evs[c])
                fr = stage_dev[bounds[c]:bounds[c + 1]].view(-1, H, W)
                dets[c] = eng2.detect(fr, pipe.K0, pipe.dist0, max_blobs=MAX_BLOBS, out=dets[c])
                xy_parts.append(dets[c].xy)
                cnt_parts.append(dets[c].count)
            full = type(det)(torch.cat(xy_parts), torch.cat(cnt_parts), det.flags)
            xy, count = pipe.exchange(full, FS)
            co = eng.correspond(xy, count, pipe.Fs, pipe.cams, obj_count=N_MARKERS, max_groups=MAX_GROUPS, out=corr_out[0])
            host_obj.copy_(co.obj, non_blocking=True)
            host_nobj.copy_(co.n_obj, non_blocking=True)
            host_cnt.copy_(full.count, non_blocking=True)
            torch.cuda.current_stream().synchronize()          # the caller reads the result on the host
            return int(host_nobj.sum())

        e2e_step()
        barrier()
        l0 = eng.launches + eng2.launches
        t0 = time.perf_counter()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(args.e2e_steps):
            e2e_step()
        b.record()
        barrier()
        wall = time.perf_counter() - t0
        ems = torch.tensor([max(a.elapsed_time(b), 0.0)], device=device, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ems, op=dist.ReduceOp.MAX)
        e2e = {"value": frames_per_step * args.e2e_steps / (float(ems.item()) * 1e-3), "unit": "frames/s",
               "h2d_bytes_per_step": int(frames.numel()) * N, "d2h_bytes_per_step": int(host_obj.numel() * 8 + host_nobj.numel() * 4 + host_cnt.numel() * 4) * N,
               "steps": args.e2e_steps, "wall_s": wall, "gpu_launches": eng.launches + eng2.launches - l0,
               "how": f"pinned host frames -> {chunks} chunked H2D copies on a copy stream overlapped with detection -> all-gather -> "
                      "match+triangulate -> D2H of object points/counts, host sync every step"}
        del host, stage_dev

    # ---- geometry-only sweep (BASELINE config 5 shape): 8-view DLT + reprojection error, FP32 main mode, vs the FP32 pipe ------
    geometry = None
    if rank == 0:
        rig5 = S.config_rig("c5")
        cams5 = eng.cameras(rig5["poses"], rig5["camera_params"])
        P5 = 4_000_000
        g5 = torch.Generator(device=device).manual_seed(5)
        X5 = torch.tensor(np.asarray(rig5["centre"]), device=device) + (torch.rand((P5, 3), generator=g5, device=device, dtype=torch.float64) - 0.5)
        Pm = torch.tensor(cams5.cpu().numpy()[:, :12].reshape(8, 3, 4), device=device)
        proj = torch.einsum("cij,pj->pci", Pm, torch.cat([X5, torch.ones((P5, 1), device=device, dtype=torch.float64)], dim=1))
        pts5 = torch.floor(proj[..., :2] / proj[..., 2:3]).float().contiguous()
        del proj, X5
        xyz5 = torch.empty((P5, 3), device=device)
        err5 = torch.empty((P5,), device=device)
        for _ in range(3):
            eng.triangulate(pts5, cams5, xyz=xyz5, err=err5)
        ga, gb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        ga.record()
        for _ in range(10):
            eng.triangulate(pts5, cams5, xyz=xyz5, err=err5)
        gb.record()
        torch.cuda.synchronize()
        tms = ga.elapsed_time(gb) / 10
        flops = 140 * 8 + 1609                                     # F(N) = 140 N + 1609 per point
        sm_mhz = (clocks or {}).get("sm_max_mhz") or 1965.0
        peak32 = 148 * 128 * 2 * sm_mhz * 1e6 / 1e12
        geometry = {"workload": "BASELINE config 5 shape: 8-view DLT triangulation + reprojection error, FP32", "points": P5, "views": 8,
                    "ms": tms, "points_per_s": P5 / (tms * 1e-3), "flop_per_point": flops,
                    "achieved_tflops": P5 * flops / (tms * 1e-3) / 1e12, "fp32_peak_tflops": peak32,
                    "frac_of_fp32_pipe": P5 * flops / (tms * 1e-3) / 1e12 / peak32,
                    "peak_source": "148 SMs x 128 FP32 lanes x 2 x clocks.max.sm (no tensor cores: tiny independent solves)"}
        del pts5, xyz5, err5

    # ---- the step in front of the path: raw Bayer GR sensor frames -> grey (RealtimeTracking_FLIR.py:103-104) ---------------
    front = None
    if rank == 0:
        nb = 256                                                   # 1.07 GB in + 1.07 GB out: larger than L2
        raw = torch.randint(0, 256, (nb, H, W), dtype=torch.uint8, device=device)
        grey = torch.empty_like(raw)
        for _ in range(3):
            eng.bayer_gr2gray(raw, out=grey)
        fa, fb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        fa.record()
        for _ in range(10):
            eng.bayer_gr2gray(raw, out=grey)
        fb.record()
        torch.cuda.synchronize()
        fms = fa.elapsed_time(fb) / 10
        front = {"workload": f"Bayer GR -> grey, {nb} frames {W}x{H} u8 (bilinear demosaic + BGR2GRAY, bit-identical to OpenCV)",
                 "kernel": "bayer_gr2gray_rows_kernel", "ms": fms, "frames_per_s": nb / (fms * 1e-3),
                 "algorithmic_bytes": 2 * nb * H * W, "achieved_gbs": 2 * nb * H * W / (fms * 1e-3) / 1e9}
        del raw, grey

    # ---- one frame through the drop-in the realtime loop calls: numpy image in, centroid list + undistorted image out ------
    latency = None
    if rank == 0:
        from mocapv2_b200 import engine as E
        from mocapv2_b200.lib import ImageOperations as IO
        E.set_default_engine(eng)
        IO.camera_params = rig["camera_params"]
        IO.ANNOTATE = False                                        # the display-only overlays are host-side cv2 drawing
        host_frame = frames[0, 0].cpu().numpy()
        for _ in range(5):
            IO._find_dot(host_frame)
        lat = []
        for _ in range(30):
            t0 = time.perf_counter()
            _, pts_one = IO._find_dot(host_frame)
            lat.append((time.perf_counter() - t0) * 1e3)
        lat.sort()
        latency = {"call": "lib.ImageOperations._find_dot(img) on one 2048x2048 host frame (H2D + detect + undistorted image D2H)",
                   "median_ms": lat[len(lat) // 2], "min_ms": lat[0], "centroids": len(pts_one)}

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel --------------------------------------------------------------------------------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    dom = max(stage_avg, key=lambda k: stage_avg[k])
    alg_bytes = n_local * H * W + n_local * (4 + 8 * MAX_BLOBS)          # H*W read + (4 + 8 n_blobs) written per frame
    achieved = alg_bytes / (stage_avg[dom] * 1e-3) / 1e9
    traffic = None
    try:
        tr = json.load(open(os.path.join(REPO, "profiles", "traffic.json")))
        if dom in tr:
            traffic = tr[dom]["dram_bytes_per_frame"] * n_local
    except Exception:
        pass
    det_ms = sum(stage_avg.values())
    kernels = {"scan": "scan_hot_vec32_kernel", "group": "form_clusters_kernel", "filter": "piece_filter_kernel",
               "borders": "candidates_kernel + borders_finalize_kernel (traces, filter/centroid/order)",
               "finish": "general path for flagged frames (mark_active/compact_tiles/filter_tiles/blobs)"}
    if front:
        front["peak_gbs"] = peak
        front["frac"] = front["achieved_gbs"] / peak
    roofline = {"bound": "hbm", "kernel": kernels.get(dom, dom), "stage": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes,
                "stage_ms": stage_avg, "detect_ms": det_ms,
                "detect_pipeline": {"achieved": alg_bytes / (det_ms * 1e-3) / 1e9, "frac": alg_bytes / (det_ms * 1e-3) / 1e9 / peak,
                                    "note": "all detection kernels of a step together (scan+group+filter+borders+finish) against the same H*W bytes"}}

    # ---- CPU baseline on this host (N=1 only) -------------------------------------------------------------------------------------
    cpu = None
    if world == 1 and args.cpu_sample > 0:
        threads = os.cpu_count() or 1
        hf = frames[:args.cpu_sample].cpu().numpy()
        r = cpu_frame_sets(rig, hf, args.cpu_sample, threads)
        if r is not None:
            dt, nf, npts = r
            cpu = {"value": nf / dt, "unit": "frames/s", "cores": threads, "kind": "port",
                   "sample": f"{args.cpu_sample} frame-set(s) = {nf} frames of this workload in {dt:.2f} s: OpenCV calls of "
                             "lib/ImageOperations.py:33-65 (numba blur -> its integer restatement) + lib/Helpers.py:178-280 in NumPy "
                             f"with the same candidate cap; frames over a {threads}-thread pool", "points_per_s": npts / dt}

    line = {
        "metric": "frames/s (detect+match+triangulate, 16-camera 2048x2048 rig, 128 markers)",
        "value": value, "unit": "frames/s", "n_gpus": N, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8 (detection, integer exact) / f32 (geometry)", "data": "synthetic",
        "config": workload_config(N, F0),
        "points_per_s": points_per_step * args.steps / (ms_total * 1e-3),
        "frame_sets_with_group_cap": float(n_pts[1].item()), "centroids_per_frame": float(n_pts[2].item()) / (n_local * N),
        "e2e": e2e, "gpu_launches": gpu_launches, "collectives_per_step": 1 if N > 1 else 0,
        "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu, "geometry": geometry, "front_step": front, "find_dot_latency": latency,
    }
    emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


RESULT_OUT = sys.stdout


def emit(line: dict):
    """The ONE JSON line of the contract, on the process's real stdout."""
    RESULT_OUT.write(json.dumps(line) + "\n")
    RESULT_OUT.flush()


def main():
    # Libraries write to fd 1 behind Python's back (NCCL prints its version line there at NCCL_DEBUG=WARN/VERSION): send fd 1
    # to stderr for the whole run and keep a private handle on the real stdout for the result line.
    global RESULT_OUT
    sys.stdout.flush()
    RESULT_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    args = parse()
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
