"""Shared helpers of the parity tests (the oracle is the checker; nothing here is product code)."""
import os

import numpy as np
import torch

from mocapv2_b200 import synth as S
from oracle import restate as R

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
K, D = S.SHIPPED_K, S.SHIPPED_DIST


def pack_bits(b):
    """(H, W) nonzero mask -> (H, ceil(W/32)) int32 words, LSB = leftmost pixel (the C-ABI's packed layout)."""
    H, W = b.shape
    TX = (W + 31) // 32
    p = np.zeros((H, TX * 32), np.uint8)
    p[:, :W] = b != 0
    return np.packbits(p, axis=1, bitorder="little").view(np.int32).reshape(H, TX)


def unpack_bits(words, W):
    w = np.ascontiguousarray(words.cpu().numpy() if isinstance(words, torch.Tensor) else words)
    return np.unpackbits(w.view(np.uint8), axis=-1, bitorder="little")[..., :W]


def oracle_contour_table(binimg, min_area=R.MIN_AREA, min_circ=R.MIN_CIRC):
    """rows [a00, a10, a01, perimeter, is_hole, parent, kept] in the reference's output order + kept centroids."""
    contours, info = R.find_contours(binimg)
    rows, pts = [], []
    for c, inf in zip(contours, info):
        a00, a10, a01, per = R.contour_stats(c)
        area = abs(a00) * 0.5
        keep = bool(per and (4 * np.pi * area / (per * per) > min_circ and area > min_area)) and a00 != 0
        if keep:
            pts.append(R.centroid(a00, a10, a01))
        rows.append([a00, a10, a01, per, int(inf[0]), int(inf[1]), int(keep)])
    return np.array(rows, dtype=np.float64).reshape(-1, 7), pts


def check_blob_outputs(res, i, binimg, min_area=R.MIN_AREA, min_circ=R.MIN_CIRC):
    """Bit-exact comparison of frame i of a DetectResult (with labels / blob_sums / contours) against the oracle."""
    ex = res.extras
    assert int(res.flags[i]) & 63 & ~16 == 0, f"flags {int(res.flags[i])}"  # 16 = deep tree (slow ordering), 64 = general path (informational)
    n, lab = R.label8(binimg)
    if "labels" in ex:
        assert np.array_equal(ex["labels"][i].cpu().numpy(), lab), "blob pixel membership differs"
    if "blob_sums" in ex:
        assert int(ex["blob_count"][i]) == n, "blob count differs"
        m = min(n, ex["blob_sums"].shape[1])
        assert np.array_equal(ex["blob_sums"][i, :m].cpu().numpy(), R.blob_pixel_sums(lab, n)[:m]), "blob pixel sums differ"
    table, pts = oracle_contour_table(binimg, min_area, min_circ)
    if "contours" in ex:
        nc = int(ex["contour_count"][i])
        assert nc == len(table), f"contour count {nc} != {len(table)}"
        got = ex["contours"][i, :nc, :7].cpu().numpy()
        assert np.array_equal(got, table), "contour table (a00,a10,a01,perimeter,hole,parent,kept) differs"
    assert res.points(i) == (pts if pts else [[None, None]]), "centroids / order differ"


def lists_to_arrays(points, max_pts=None):
    """Per-camera centroid lists (reference format, may hold one [None, None]) -> xy [1,C,max_pts,2] int32, count [1,C]."""
    C = len(points)
    clean = [[q for q in p if q[0] is not None] for p in points]
    mp = max(1, max(len(p) for p in clean)) if max_pts is None else max_pts
    xy = np.zeros((1, C, mp, 2), np.int32)
    cnt = np.zeros((1, C), np.int32)
    for c, p in enumerate(clean):
        cnt[0, c] = len(p)
        if p:
            xy[0, c, :len(p)] = np.array(p, dtype=np.int32)
    return xy, cnt
