"""Small end-to-end run for compute-sanitizer (memcheck): every kernel of the library on small inputs."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mocapv2_b200 import synth as S
from mocapv2_b200.engine import CaptureEngine
eng = CaptureEngine("cuda:0")
K, D = S.SHIPPED_K, S.SHIPPED_DIST
z = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "shapes.npz"))
for key in ("frame_0", "frame_2", "frame_4", "frame_5"):
    fr = torch.from_numpy(z[key][None].copy()).cuda()
    eng.detect(fr, K, D, min_area=0.0)
    eng.detect(fr, K, D, outputs=("bits", "labels", "blob_sums", "contours"), min_area=0.0)
    eng.undistort(fr, K, D); eng.blur5(fr); eng.median5_threshold(fr); eng.filter(fr, K, D)
c1 = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "c1_frames.npz"))["frames"]
res = eng.detect(torch.from_numpy(c1.reshape(-1, 480, 640).copy()).cuda(), K, D)
rig = S.config_rig("c1")
cams = eng.cameras(rig["poses"], rig["camera_params"])
mp = max(1, int(res.count.max()))
xy = res.xy[:, :mp].reshape(3, 2, mp, 2).contiguous()
out = eng.correspond(xy, res.count.reshape(3, 2).contiguous(), torch.tensor(np.array(rig["Fs"]), device="cuda"), cams, obj_count=4)
pts = torch.rand((1000, 2, 2), device="cuda") * 600
eng.triangulate(pts, cams); eng.triangulate(pts.double(), cams)
torch.cuda.synchronize()
print("sanitize probe ok", res.count.tolist(), out.n_obj.tolist())
