"""Timing probe of the Bayer GR -> grey front step (development tool)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mocapv2_b200.engine import CaptureEngine

eng = CaptureEngine("cuda:0")
n, H, W = int(sys.argv[1]) if len(sys.argv) > 1 else 256, 2048, 2048
raw = torch.randint(0, 256, (n, H, W), dtype=torch.uint8, device="cuda:0")
out = torch.empty_like(raw)
for _ in range(3):
    eng.bayer_gr2gray(raw, out=out)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
a.record()
for _ in range(10):
    eng.bayer_gr2gray(raw, out=out)
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / 10
print(f"bayer {n} frames: {ms:.3f} ms  {2 * n * H * W / ms / 1e6:.0f} GB/s r+w")
t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
t0.record()
for _ in range(10):
    out.copy_(raw)
t1.record()
torch.cuda.synchronize()
print(f"copy: {t0.elapsed_time(t1) / 10:.3f} ms")

# the per-pixel kernel (rows that are not a multiple of 4 bytes) for comparison
raw2 = torch.randint(0, 256, (n, H, W - 1), dtype=torch.uint8, device="cuda:0")
out2 = torch.empty_like(raw2)
for _ in range(2):
    eng.bayer_gr2gray(raw2, out=out2)
torch.cuda.synchronize()
a.record()
for _ in range(5):
    eng.bayer_gr2gray(raw2, out=out2)
b.record()
torch.cuda.synchronize()
ms2 = a.elapsed_time(b) / 5
print(f"per-pixel kernel, {W - 1} wide: {ms2:.3f} ms  {2 * n * H * (W - 1) / ms2 / 1e6:.0f} GB/s r+w")
