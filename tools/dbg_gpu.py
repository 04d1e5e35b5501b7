import sys, os, numpy as np, torch, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from mocapv2_b200.engine import CaptureEngine
eng = CaptureEngine("cuda:0")
rig, cen, ridx = bench.make_scene(4)
frames = bench.render_local(rig, cen, ridx, 0, 16, eng.device).view(-1, 2048, 2048)
K, D = rig["camera_params"][0]["intrinsic_matrix"], rig["camera_params"][0]["distortion_coef"]
lean = eng.detect(frames, K, D, max_blobs=160)
torch.cuda.synchronize()
fl = lean.flags.tolist()
print('frames', len(fl), 'general', sum(1 for x in fl if x & 64), 'reasons', collections.Counter(x >> 8 for x in fl))
print('counts', lean.count.tolist()[:32])
