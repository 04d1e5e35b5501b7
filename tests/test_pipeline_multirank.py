"""The N>1 path: cameras sharded over ranks, one all-gather of centroid records, frame-sets sharded for geometry.
world_size 2 over gloo on the CPU (kernel sources through the emulation build); outputs must be bit-identical to the
1-rank run.  The same check runs over NCCL on real GPUs when the box has two of them
(`-m gpu`, skipped on a one-GPU box)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mocapv2_b200 import synth as S
from util import GOLDEN

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _frames():
    z = np.load(os.path.join(GOLDEN, "c1_frames.npz"))
    return z["frames"][:2]                                   # [FS=2, C=2, 480, 640]


def _ring4_scene():
    """Four cameras on a ring, 480x360, two frame-sets of three markers (deterministic): with two ranks every rank owns TWO
    cameras, so the (rank, local camera) -> camera order of the exchange is exercised."""
    rig = S.ring_rig(4, 480, 360, radius=6.0)
    rng = np.random.default_rng(404)
    frames = []
    for _ in range(2):
        X = S.sample_markers(rig, 3, rng, spread=0.10, min_sep_px=75.0, margin=36.0, tries=400)
        radii = rng.integers(17, 21, (4, len(X)))
        frames.append(S.render_frameset(rig, X, radii, rng))
    return rig, np.stack(frames)                              # [FS=2, C=4, 360, 480]


def _run(rank, world, port, out_dir, backend="gloo", scene="c1", overlapped=False):
    sys.path.insert(0, REPO)
    sys.path.insert(0, os.path.join(REPO, "tests", "emu"))
    from mocapv2_b200.engine import CaptureEngine
    from mocapv2_b200.pipeline import CapturePipeline
    if backend == "nccl":
        torch.cuda.set_device(rank)
        if world > 1:
            dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world,
                                    device_id=torch.device(f"cuda:{rank}"))
        eng = CaptureEngine(f"cuda:{rank}")
    else:
        import build_emu
        if world > 1:
            dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
        from emu_engine import EmuEngine
        eng = EmuEngine(build_emu.build())
    if scene == "ring4":
        rig, fr = _ring4_scene()
    else:
        rig, fr = S.config_rig("c1"), _frames()
    pipe = CapturePipeline(eng, rig, max_blobs=8, obj_count=4, max_groups=16, fp64=False)
    frames = torch.from_numpy(fr)
    local = frames[:, pipe.cam_begin:pipe.cam_begin + pipe.cams_local].contiguous().to(eng.device)
    if overlapped == "lanes":                                 # the bench's timed loop: five steps over two lanes (StepsInFlight)
        from mocapv2_b200.pipeline import StepsInFlight
        pipe.pipelined_min_frames = 1
        flight = StepsInFlight(pipe, depth=2)
        flight.fork()
        results = [flight.submit(local) for _ in range(5)]
        flight.join()
        if eng.device.type == "cuda":
            torch.cuda.synchronize()
        assert world == 1 or eng.device.type != "cuda" or all(l._peer for l in flight.lanes), "store-to-peer exchange not set up on every lane"
        res = results[-1]                                     # (lane 0's third step; lane 1's last result is results[-2])
        other = results[-2]
        assert torch.equal(other.corr.n_obj, res.corr.n_obj) and torch.equal(other.det.count, res.det.count)
        pipe.collectives = min(pipe.collectives, 1)
    elif overlapped:                                          # the batch path: overlapped detection + store-to-peer exchange
        pipe.pipelined_min_frames = 1
        for _ in range(3):                                    # three steps: the two receive buffers alternate
            res = pipe.step(local)
        assert world == 1 or pipe._peer, "the store-to-peer exchange was not set up"
        pipe.collectives = min(pipe.collectives, 1)
    else:
        res = pipe.step(local)
    cpu = lambda t: t.cpu()
    torch.save({"b": res.fs_begin, "e": res.fs_end, "obj": cpu(res.corr.obj), "n_obj": cpu(res.corr.n_obj), "img": cpu(res.corr.img),
                "n_valid": cpu(res.corr.n_valid), "err": cpu(res.corr.err), "count": cpu(res.det.count), "collectives": pipe.collectives},
               os.path.join(out_dir, f"rank{rank}_of{world}.pt"))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def _compare(out, cams_per_rank=1):
    one = torch.load(os.path.join(out, "rank0_of1.pt"))
    assert int(one["n_valid"].sum()) > 0 and one["collectives"] == 0
    for r in range(2):
        two = torch.load(os.path.join(out, f"rank{r}_of2.pt"))
        b, e = two["b"], two["e"]
        assert (b, e) == (r, r + 1) and two["collectives"] == 1          # exactly one exchange per batch
        assert torch.equal(two["n_valid"], one["n_valid"][b:e]) and torch.equal(two["n_obj"], one["n_obj"][b:e])
        for s in range(e - b):
            nv, no = int(two["n_valid"][s]), int(two["n_obj"][s])
            assert torch.equal(two["img"][s, :nv], one["img"][b + s, :nv])
            assert torch.equal(two["obj"][s, :no], one["obj"][b + s, :no])  # bit-identical, no cross-rank reductions
            assert torch.equal(two["err"][s, :nv], one["err"][b + s, :nv])
        # each rank detected only its own cameras
        cl = cams_per_rank
        assert torch.equal(two["count"].view(2, cl), one["count"].view(2, 2 * cl)[:, r * cl:(r + 1) * cl])


def test_two_ranks_equal_one_rank(tmp_path):
    sys.path.insert(0, os.path.join(REPO, "tests", "emu"))
    import build_emu
    build_emu.build()
    out = str(tmp_path)
    _run(0, 1, 0, out)
    port = 29500 + os.getpid() % 2000
    mp.spawn(_run, args=(2, port, out), nprocs=2, join=True)
    _compare(out)


def test_two_ranks_with_two_cameras_each(tmp_path):
    sys.path.insert(0, os.path.join(REPO, "tests", "emu"))
    import build_emu
    build_emu.build()
    out = str(tmp_path)
    _run(0, 1, 0, out, "gloo", "ring4")
    one = torch.load(os.path.join(out, "rank0_of1.pt"))
    assert int(one["n_obj"].min()) >= 2                       # the scene triangulates in every frame-set
    port = 33500 + os.getpid() % 2000
    mp.spawn(_run, args=(2, port, out, "gloo", "ring4"), nprocs=2, join=True)
    _compare(out, cams_per_rank=2)


def test_two_ranks_with_steps_in_flight(tmp_path):
    """StepsInFlight over two ranks (gloo, emulation build: lanes without streams, NCCL-style exchange): same results as one rank."""
    sys.path.insert(0, os.path.join(REPO, "tests", "emu"))
    import build_emu
    build_emu.build()
    out = str(tmp_path)
    _run(0, 1, 0, out, "gloo", "ring4", "lanes")
    port = 34500 + os.getpid() % 2000
    mp.spawn(_run, args=(2, port, out, "gloo", "ring4", "lanes"), nprocs=2, join=True)
    _compare(out, cams_per_rank=2)


@pytest.mark.gpu
def test_two_gpus_steps_in_flight_equal_one_gpu(tmp_path):
    """Two steps in flight on each of two GPUs: every lane has its own symmetric receive buffers and stores to the peer's, the scans of
    the lanes are chained by the scan token; five steps, results equal to one GPU."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    out = str(tmp_path)
    port = 30500 + os.getpid() % 2000
    mp.spawn(_run, args=(1, port, out, "nccl", "ring4", "lanes"), nprocs=1, join=True)
    mp.spawn(_run, args=(2, port, out, "nccl", "ring4", "lanes"), nprocs=2, join=True)
    _compare(out, cams_per_rank=2)


@pytest.mark.gpu
def test_two_gpus_over_nccl_equal_one_gpu(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (the one-GPU suite covers the same logic over gloo)")
    out = str(tmp_path)
    port = 31500 + os.getpid() % 2000
    mp.spawn(_run, args=(1, port, out, "nccl", "ring4"), nprocs=1, join=True)
    mp.spawn(_run, args=(2, port, out, "nccl", "ring4"), nprocs=2, join=True)
    _compare(out, cams_per_rank=2)


@pytest.mark.gpu
def test_two_gpus_store_to_peer_equal_one_gpu(tmp_path):
    """The same through the overlapped detection call with its store-to-peer epilogue (records written straight into the matching
    rank's receive buffer over NVLink, symmetric-memory barrier instead of a collective), three steps in a row."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    out = str(tmp_path)
    port = 32500 + os.getpid() % 2000
    mp.spawn(_run, args=(1, port, out, "nccl", "ring4", True), nprocs=1, join=True)
    mp.spawn(_run, args=(2, port, out, "nccl", "ring4", True), nprocs=2, join=True)
    _compare(out, cams_per_rank=2)
