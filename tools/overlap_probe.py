"""Does running the streaming scan of one chunk beside the piece filter of another pay?  (development probe)

Splits a resident C4 batch into chunks that alternate between two engines (own workspace each) on two streams and
compares the batch time with the single-call time.
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mocapv2_b200 import synth as S
from mocapv2_b200.engine import CaptureEngine, DetectResult


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=1024)
    ap.add_argument("--reps", type=int, default=10)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    H = W = 2048
    M = 128
    n = a.frames
    g = torch.Generator().manual_seed(1)
    centres = torch.stack([torch.randint(40, W - 40, (n, M), generator=g), torch.randint(40, H - 40, (n, M), generator=g)], dim=-1)
    ridx = torch.randint(0, 9, (n, M), generator=g)
    frames = torch.empty((n, H, W), dtype=torch.uint8, device=dev)
    for i in range(0, n, 32):
        S.render_batch_torch(H, W, centres[i:i + 32].to(dev), ridx[i:i + 32].to(dev), 100 + i, dev, out=frames[i:i + 32])
    K, D = S.SHIPPED_K, S.SHIPPED_DIST
    mb = 160
    flat_xy = torch.zeros((n, mb, 2), dtype=torch.int32, device=dev)
    cnt = torch.zeros(n, dtype=torch.int32, device=dev)
    flg = torch.zeros(n, dtype=torch.int32, device=dev)

    def run(n_chunks, n_streams):
        engines = [CaptureEngine(dev) for _ in range(n_streams)]
        streams = [torch.cuda.Stream(dev) for _ in range(n_streams)]
        per = n // n_chunks
        outs = [DetectResult(flat_xy[c * per:(c + 1) * per], cnt[c * per:(c + 1) * per], flg[c * per:(c + 1) * per]) for c in range(n_chunks)]

        def step():
            main_s = torch.cuda.current_stream(dev)
            if n_streams == 1 and n_chunks == 1:
                engines[0].detect(frames, K, D, max_blobs=mb, out=outs[0])
                return
            for s in streams:
                s.wait_stream(main_s)
            for c in range(n_chunks):
                with torch.cuda.stream(streams[c % n_streams]):
                    engines[c % n_streams].detect(frames[c * per:(c + 1) * per], K, D, max_blobs=mb, out=outs[c])
            for s in streams:
                main_s.wait_stream(s)

        for _ in range(3):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.reps):
            step()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / a.reps

    base = run(1, 1)
    ref = (flat_xy.clone(), cnt.clone())
    print(f"single call           {base:7.3f} ms  {n / base * 1e3:9.0f} frames/s")
    for nc, ns in ((2, 2), (4, 2), (8, 2), (16, 2), (4, 4), (8, 4), (16, 4), (3, 3), (6, 3)):
        if n % nc:
            continue
        flat_xy.zero_(); cnt.zero_()
        t = run(nc, ns)
        same = torch.equal(cnt, ref[1]) and torch.equal(flat_xy, ref[0])
        print(f"{nc:2d} chunks {ns} streams   {t:7.3f} ms  {n / t * 1e3:9.0f} frames/s  x{base / t:.3f}  same={same}")


if __name__ == "__main__":
    main()
