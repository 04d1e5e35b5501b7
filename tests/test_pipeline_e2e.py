"""End-to-end parity of the batched pipeline (detect -> match -> triangulate) against the CPU restatement of the
reference on the same synthetic frame-sets: BASELINE config 1 everywhere, config 3 (6 cameras 1440x1080, 32 markers)
and a reduced config 4 (16 cameras 2048x2048) on the GPU.  Centroids and matched pairs bit-exact, 3-D points within
1e-4 relative, reprojection error within 1e-3 px (RMS pixels)."""
import numpy as np
import pytest
import torch

from mocapv2_b200 import synth as S
from mocapv2_b200.pipeline import CapturePipeline, StepsInFlight
from oracle import restate as R


def scene(name, n_sets, n_markers, seed, spread):
    rig = S.config_rig(name)
    rng = np.random.default_rng(seed)
    if name == "c1":
        X0 = np.array(S.C1_MARKERS)
    else:
        X0 = S.sample_markers(rig, n_markers, rng, spread=spread, min_sep_px=70.0, tries=60)
    frames = []
    for s in range(n_sets):
        X = X0 + rng.uniform(-0.01, 0.01, X0.shape)
        radii = rng.integers(14, 23, (len(rig["poses"]), len(X)))
        frames.append(S.render_frameset(rig, X, radii, rng))
    return rig, np.stack(frames)


def check(engine, name, n_sets, n_markers, seed, spread, max_groups):
    rig, frames = scene(name, n_sets, n_markers, seed, spread)
    pipe = CapturePipeline(engine, rig, max_blobs=64, obj_count=n_markers, max_groups=max_groups)
    res = pipe.step(torch.from_numpy(frames).to(engine.device))
    K = rig["camera_params"][0]["intrinsic_matrix"]
    D = rig["camera_params"][0]["distortion_coef"]
    C = len(rig["poses"])
    total_pts = 0
    for s in range(n_sets):
        pts = [R.find_dot(frames[s, c], K, D) for c in range(C)]
        for c in range(C):
            assert res.det.points(s * C + c) == pts[c], f"centroids of frame-set {s} camera {c}"
        trace = {}
        obj, ipa = R.correspond(pts, rig["poses"], rig["camera_params"], rig["Fs"], n_markers, trace=trace, max_cand=8, max_groups=max_groups)
        nv, no = int(res.corr.n_valid[s]), int(res.corr.n_obj[s])
        assert nv == (len(ipa) if ipa.size else 0) and no == (len(obj) if obj.size else 0)
        if nv:
            assert np.array_equal(res.corr.img[s, :nv].cpu().numpy(), ipa)                       # matched pairs: bit-exact
            got = res.corr.obj[s, :no].cpu().numpy()
            assert (np.abs(got - obj).max(axis=1) / np.linalg.norm(obj, axis=1)).max() < 1e-4     # 3-D points
            e = res.corr.err[s, :nv].cpu().numpy()
            assert np.abs(np.sqrt(e) - np.sqrt(np.array(trace["errors"]))).max() < 1e-3          # reprojection error, px
        total_pts += no
    return total_pts


def test_config1_two_cameras(engine):
    assert check(engine, "c1", 2, 4, 101, 0.0, 64) > 0


@pytest.mark.gpu
def test_config3_six_cameras_1440x1080(gpu_engine):
    assert check(gpu_engine, "c3", 4, 32, 103, 0.42, 256) > 0


@pytest.mark.gpu
def test_config4_sixteen_cameras_2048_reduced_markers(gpu_engine):
    """Config 4 at full frame size with a marker count the reference's group enumeration can finish."""
    assert check(gpu_engine, "c4", 2, 12, 104, 0.9, 4096) > 0


def test_steps_in_flight_equal_one_step_after_the_other(engine):
    """StepsInFlight: five steps over two lanes (own engine state, buffers and -- on the GPU -- streams) give what pipe.step gives."""
    rig, frames = scene("c1", 2, 4, 105, 0.0)
    batches = [torch.from_numpy(np.roll(frames, k, axis=0).copy()).to(engine.device) for k in range(2)]
    pipe = CapturePipeline(engine, rig, max_blobs=64, obj_count=4, max_groups=64)
    pipe.pipelined_min_frames = 1                                  # the overlapped detection call on every lane
    want = []
    for k in range(2):
        r = pipe.step(batches[k])
        want.append((r.det.count.clone(), r.det.xy.clone(), r.corr.n_obj.clone(), r.corr.obj.clone(), r.corr.img.clone()))
    flight = StepsInFlight(pipe, depth=2)
    flight.fork()
    got = [flight.submit(batches[i % 2]) for i in range(5)]
    flight.join()
    if engine.device.type == "cuda":
        torch.cuda.synchronize()
    assert len(flight.lanes) == 2 and flight.lanes[1].eng is not pipe.eng
    for i in (3, 4):                                               # the last result of each lane is still in its buffers
        r, w = got[i], want[i % 2]
        assert torch.equal(r.det.count, w[0]) and torch.equal(r.corr.n_obj, w[2])
        for f in range(len(w[0])):
            assert torch.equal(r.det.xy[f, :int(w[0][f])], w[1][f, :int(w[0][f])])
        for s_ in range(len(w[2])):
            no = int(w[2][s_])
            assert torch.equal(r.corr.obj[s_, :no], w[3][s_, :no]) and torch.equal(r.corr.img[s_, :no], w[4][s_, :no])
