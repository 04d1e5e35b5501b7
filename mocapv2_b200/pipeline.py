"""Batched capture pipeline over one or several GPUs: detect -> (all-gather centroid lists) -> match + triangulate.

Sharding: detection is independent per (camera, frame) and is sharded by CAMERA, so a rank keeps
its cameras' frames (and their undistortion table) resident; matching needs every camera's centroid list of a
frame-set, so the fixed-stride records [count | xy] are exchanged with ONE all-gather (NCCL over NVLink on the GPU
box, gloo in the CPU tests); matching + triangulation are then sharded by FRAME-SET with no further exchange.
No floating-point reduction crosses ranks, so N-rank outputs are bit-identical to 1-rank outputs.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch
import torch.distributed as dist

from .engine import CaptureEngine, CorrespondResult, DetectResult


@dataclass
class StepResult:
    det: DetectResult                 # this rank's cameras, all frame-sets: [FS * cams_local, ...]
    corr: CorrespondResult            # this rank's frame-set shard, all cameras
    fs_begin: int                     # first frame-set of the shard
    fs_end: int


class CapturePipeline:
    def __init__(self, engine: CaptureEngine, rig: dict, *, max_blobs=160, obj_count=None, max_groups=64, fp64=False,
                 group=None):
        self.eng = engine
        self.rig = rig
        self.C = len(rig["poses"])
        self.H, self.W = rig["H"], rig["W"]
        self.max_blobs = int(max_blobs)
        self.obj_count = self.max_blobs if obj_count is None else int(obj_count)
        self.max_groups = int(max_groups)
        self.fp64 = fp64
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if self.world > 1 else 0
        if self.C % self.world:
            raise ValueError(f"{self.C} cameras do not shard over {self.world} ranks")
        self.cams_local = self.C // self.world
        self.cam_begin = self.rank * self.cams_local
        # _find_dot undistorts with camera 0's calibration for every camera (lib/ImageOperations.py:36-38)
        self.K0 = np.asarray(rig["camera_params"][0]["intrinsic_matrix"], dtype=np.float64)
        self.dist0 = np.asarray(rig["camera_params"][0]["distortion_coef"], dtype=np.float64)
        self.cams = engine.cameras(rig["poses"], rig["camera_params"])
        self.Fs = torch.from_numpy(np.asarray(rig["Fs"], dtype=np.float64).reshape(-1, 3, 3).copy()).to(engine.device)
        self._det = None
        self._flat = None
        self._gath = None
        self.collectives = 0

    def frame_set_shard(self, n_frame_sets: int):
        if n_frame_sets % self.world:
            raise ValueError("frame-sets must divide evenly over the ranks")
        per = n_frame_sets // self.world
        return self.rank * per, (self.rank + 1) * per

    def _buffers(self, n: int):
        """Detection outputs of n frames as two views of ONE flat int32 buffer [xy (n*mb*2) | count (n)]: the exchange
        all-gathers that buffer as it is, nothing is packed."""
        if self._det is None or self._det.xy.shape[0] != n:
            mb = self.max_blobs
            self._flat = torch.zeros(n * (2 * mb + 1), dtype=torch.int32, device=self.eng.device)
            self._det = DetectResult(self._flat[: n * mb * 2].view(n, mb, 2), self._flat[n * mb * 2:],
                                     torch.zeros(n, dtype=torch.int32, device=self.eng.device))
            self._gath = None
        return self._det

    def detect(self, frames: torch.Tensor, timer=None) -> DetectResult:
        """frames [FS, cams_local, H, W] uint8 on the device -> centroid lists of this rank's cameras."""
        FS, cl, H, W = frames.shape
        assert cl == self.cams_local and (H, W) == (self.H, self.W)
        flat = frames.view(FS * cl, H, W)
        self._det = self.eng.detect(flat, self.K0, self.dist0, max_blobs=self.max_blobs, out=self._buffers(FS * cl), timer=timer)
        return self._det

    def exchange(self, det: DetectResult, FS: int):
        """One all-gather of the fixed-stride records; returns (xy [F0, C, max_blobs, 2], count [F0, C]) of this rank's shard."""
        cl, mb = self.cams_local, self.max_blobs
        b, e = self.frame_set_shard(FS)
        if self.world == 1:
            return det.xy.view(FS, cl, mb, 2), det.count.view(FS, cl)
        n = FS * cl
        if det is not self._det:                               # results that do not live in the pipeline's flat buffer
            self._buffers(n)
            self._det.xy.copy_(det.xy)
            self._det.count.copy_(det.count)
        if self._gath is None:
            self._gath = torch.empty((self.world, self._flat.numel()), dtype=torch.int32, device=self.eng.device)
        dist.all_gather_into_tensor(self._gath.view(-1), self._flat, group=self.group)
        self.collectives += 1
        xy = self._gath[:, : n * mb * 2].view(self.world, FS, cl, mb, 2)[:, b:e].permute(1, 0, 2, 3, 4).reshape(e - b, self.C, mb, 2)
        count = self._gath[:, n * mb * 2:].view(self.world, FS, cl)[:, b:e].permute(1, 0, 2).reshape(e - b, self.C)
        return xy.contiguous(), count.contiguous()                      # camera index = rank * cams_local + local

    def step(self, frames: torch.Tensor) -> StepResult:
        FS = frames.shape[0]
        det = self.detect(frames)
        xy, count = self.exchange(det, FS)
        corr = self.eng.correspond(xy, count, self.Fs, self.cams, obj_count=self.obj_count, max_groups=self.max_groups,
                                   fp64=self.fp64)
        b, e = self.frame_set_shard(FS)
        return StepResult(det, corr, b, e)
