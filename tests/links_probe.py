"""Design probe for a trace-free border stage (development tool, CPU only; not used by the product).

Claim checked here against cv2.findContours(CHAIN_APPROX_NONE): the pixel-to-pixel steps ("links") of ALL borders of a
binary image are determined by the 2x2 pixel configuration at each lattice corner:
  1 or 4 or 0 foreground pixels : no link
  2 edge-adjacent foreground    : one axis link between them
  3 foreground                  : one diagonal link between the two pixels edge-adjacent to the background pixel
  2 diagonal foreground         : two diagonal links (one each way)
with the direction given by the orientation rule below.  The Green sums (a00, a10, a01 of cv.moments on a contour) are
sums over links.  Every link keeps one background pixel on its right-hand side; the pair (8-connected foreground
component of the link's pixels, 4-connected background component of that pixel) names the border the link belongs to, so
the sums of every border -- outer and hole, any nesting -- can be accumulated with atomics per label pair, without
following any border (second check below: the per-label-pair a00/a10/a01 against those of the cv2 contours).
Third check: the whole per-border record (a00, a10, a01, perimeter of the CHAIN_APPROX_SIMPLE polygon, hole flag, start
pixel) computed without following a border -- step sums per label pair, perimeter = axis steps + float32 sqrt(2 k^2) per
diagonal run (a step's successor is one 3x3 lookup), hole flag = sign of a00, start = smallest west/east edge key of the
label pair -- against the oracle's contour table (oracle/restate.py).
Fourth check: the bit-parallel form a kernel would use -- eight step masks per pair of adjacent bit rows from shifts and
logic (step_masks) -- gives the same links.
python tests/links_probe.py [n_images]
"""
import os
import sys
from collections import Counter

import cv2
import numpy as np
from scipy import ndimage

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import restate as R                      # noqa: E402


def local_links(img):
    """Links from the 2x2 configuration at every lattice corner of the zero-padded image; returns Counter of (x0,y0,x1,y1)."""
    H, W = img.shape
    p = np.zeros((H + 2, W + 2), bool)
    p[1:-1, 1:-1] = img != 0
    out = Counter()
    rights = {}
    for vy in range(H + 1):                 # corner (vx, vy) touches padded pixels a=(vy,vx) NW, b=(vy,vx+1) NE, c=(vy+1,vx) SW, d=(vy+1,vx+1) SE
        for vx in range(W + 1):
            a, b, c, d = p[vy, vx], p[vy, vx + 1], p[vy + 1, vx], p[vy + 1, vx + 1]
            A, B, C, D = (vx - 1, vy - 1), (vx, vy - 1), (vx - 1, vy), (vx, vy)      # unpadded pixel coordinates
            n = int(a) + int(b) + int(c) + int(d)
            def add(s, t):
                out[(s[0], s[1], t[0], t[1])] += 1
                # the background pixel on the right-hand side of the step s -> t, among the corner's four pixels
                hx, hy = t[0] - s[0], t[1] - s[1]
                for q, isfg in ((A, a), (B, b), (C, c), (D, d)):        # n == 3: the corner's only background pixel
                    if not isfg and (n == 3 or (q[0] - s[0]) * hy - (q[1] - s[1]) * hx < 0):
                        rights[(s[0], s[1], t[0], t[1])] = q
            if n == 2:
                if a and b: add(A, B)            # background below  (a border keeps the background on its right-hand side)
                elif c and d: add(D, C)          # background above
                elif a and c: add(C, A)          # background to the east
                elif b and d: add(B, D)          # background to the west
                elif a and d: add(A, D); add(D, A)
                else: add(B, C); add(C, B)
            elif n == 3:
                if not d: add(C, B)              # links the two pixels edge-adjacent to the background pixel
                elif not a: add(B, C)
                elif not b: add(D, A)
                else: add(A, D)
    return out, rights


def step_masks(U, L):
    """Bit rows U (row y) and L (row y + 1) of the zero-padded image, bit x = pixel x (Python ints = arbitrarily wide
    words).  With a = U[x], b = U[x+1], c = L[x], d = L[x+1] the corner between them emits (bit x of each mask):
      E  from (x, y)      a b ~c ~d      W  from (x+1, y+1)  c d ~a ~b      N  from (x, y+1)  a c ~b ~d      S  from (x+1, y)  b d ~a ~c
      SE from (x, y)      a d ~c         NW from (x+1, y+1)  a d ~b         SW from (x+1, y)  b c ~a         NE from (x, y+1)  b c ~d
    (the diagonal masks cover both the two-diagonal-pixel corner and the three-pixel corners)."""
    a, b, c, d = U, U >> 1, L, L >> 1
    n = lambda v: ~v
    return {"E": a & b & n(c) & n(d), "W": c & d & n(a) & n(b), "N": a & c & n(b) & n(d), "S": b & d & n(a) & n(c),
            "SE": a & d & n(c), "NW": a & d & n(b), "SW": b & c & n(a), "NE": b & c & n(d)}


def bitrow_links(img):
    H, W = img.shape
    rows = [0] * (H + 2)                                  # padded rows -1 .. H; bit x + 1 = pixel x (one padding column on the left)
    for y in range(H):
        for x in range(W):
            if img[y, x]:
                rows[y + 1] |= 1 << (x + 1)
    src = {"E": (0, 0, 1, 0), "W": (1, 1, -1, 0), "N": (0, 1, 0, -1), "S": (1, 0, 0, 1),
           "SE": (0, 0, 1, 1), "NW": (1, 1, -1, -1), "SW": (1, 0, -1, 1), "NE": (0, 1, 1, -1)}   # source pixel offset from the corner's NW pixel, step
    out = Counter()
    for yy in range(H + 1):                               # corner row between padded rows yy and yy + 1, i.e. image rows yy - 1 and yy
        for name, m in step_masks(rows[yy], rows[yy + 1]).items():
            m &= (1 << (W + 2)) - 1
            ox, oy, dx, dy = src[name]
            x = 0
            while m:
                if m & 1:
                    sx, sy = x - 1 + ox, yy - 1 + oy
                    out[(sx, sy, sx + dx, sy + dy)] += 1
                    # contribution of the step in closed form, (cx, cy) = the corner's NW pixel:
                    #   dxy: E -cy, W cy+1, N -cx, S cx+1, SE cx-cy, NW cy-cx, SW cx+cy+1, NE -(cx+cy+1)
                    #   a10 += dxy * (x0 + x1), a01 += dxy * (y0 + y1) with x0 + x1 = 2 cx + 1 (N: 2 cx, S: 2 cx + 2),
                    #   y0 + y1 = 2 cy + 1 (E: 2 cy, W: 2 cy + 2)
                    cx, cy = x - 1, yy - 1
                    dxy = {"E": -cy, "W": cy + 1, "N": -cx, "S": cx + 1, "SE": cx - cy, "NW": cy - cx, "SW": cx + cy + 1, "NE": -(cx + cy + 1)}[name]
                    sumx = {"N": 2 * cx, "S": 2 * cx + 2}.get(name, 2 * cx + 1)
                    sumy = {"E": 2 * cy, "W": 2 * cy + 2}.get(name, 2 * cy + 1)
                    assert dxy == sx * (sy + dy) - (sx + dx) * sy and sumx == 2 * sx + dx and sumy == 2 * sy + dy
                m >>= 1; x += 1
    return out


def cv_links(img):
    cs, _ = cv2.findContours((img != 0).astype(np.uint8), cv2.RETR_LIST, cv2.CHAIN_APPROX_NONE)
    out = Counter()
    for c in cs:
        pts = c[:, 0, :]
        if len(pts) == 1:
            continue
        for i in range(len(pts)):
            s, t = pts[i], pts[(i + 1) % len(pts)]
            out[(int(s[0]), int(s[1]), int(t[0]), int(t[1]))] += 1
    return out


def green(pts):
    a00 = a10 = a01 = 0
    for i in range(len(pts)):
        (x0, y0), (x1, y1) = pts[i - 1], pts[i]
        dxy = int(x0) * int(y1) - int(x1) * int(y0)
        a00 += dxy; a10 += dxy * (int(x0) + int(x1)); a01 += dxy * (int(y0) + int(y1))
    return a00, a10, a01


def per_border_check(img, links, rights):
    H, W = img.shape
    fgp = np.zeros((H + 2, W + 2), bool)
    fgp[1:-1, 1:-1] = img != 0
    fl, _ = ndimage.label(fgp, structure=np.ones((3, 3)))                 # 8-connected foreground
    bl, _ = ndimage.label(~fgp)                                           # 4-connected background (padded: one outer region)
    acc = {}
    for (x0, y0, x1, y1), cnt in links.items():
        assert cnt == 1
        r = rights[(x0, y0, x1, y1)]
        key = (int(fl[y0 + 1, x0 + 1]), int(bl[r[1] + 1, r[0] + 1]))
        dxy = x0 * y1 - x1 * y0
        s = acc.setdefault(key, [0, 0, 0])
        s[0] += dxy; s[1] += dxy * (x0 + x1); s[2] += dxy * (y0 + y1)
    mine = sorted(tuple(v) for v in acc.values())
    cs, _ = cv2.findContours((img != 0).astype(np.uint8), cv2.RETR_LIST, cv2.CHAIN_APPROX_SIMPLE)
    ref = sorted(green(c[:, 0, :]) for c in cs if len(c) > 1)
    return mine == ref


DX = (1, 1, 0, -1, -1, -1, 0, 1)                        # 0=E 1=NE 2=N 3=NW 4=W 5=SW 6=S 7=SE (y grows downwards)
DY = (0, -1, -1, -1, 0, 1, 1, 1)


def trace_free_records(img):
    """Sorted [(a00, a10, a01, perimeter, is_hole, start_x, start_y)] of every border with more than one pixel."""
    H, W = img.shape
    links, rights = local_links(img)
    fgp = np.zeros((H + 2, W + 2), bool)
    fgp[1:-1, 1:-1] = img != 0
    fl, _ = ndimage.label(fgp, structure=np.ones((3, 3)))
    bl, _ = ndimage.label(~fgp)
    fg = lambda x, y: bool(fgp[y + 1, x + 1])
    key_of = {L: (int(fl[L[1] + 1, L[0] + 1]), int(bl[rights[L][1] + 1, rights[L][0] + 1])) for L in links}
    dir_of = lambda L: [d for d in range(8) if (DX[d], DY[d]) == (L[2] - L[0], L[3] - L[1])][0]
    succ = {}
    for L in links:                                  # successor: one border-following step at the link's end pixel
        x, y, d = L[2], L[3], dir_of(L)
        back = (d + 4) & 7
        for k in range(1, 9):
            nd = (back + k) & 7
            if fg(x + DX[nd], y + DY[nd]):
                succ[L] = (x, y, x + DX[nd], y + DY[nd])
                break
    assert sorted(succ.values()) == sorted(links)    # a bijection on the links
    pred_dir = {succ[L]: dir_of(L) for L in links}
    rec = {}
    for L in links:
        x0, y0, x1, y1 = L
        r = rec.setdefault(key_of[L], {"a": [0, 0, 0], "axis": 0, "diag": 0.0})
        dxy = x0 * y1 - x1 * y0
        r["a"][0] += dxy; r["a"][1] += dxy * (x0 + x1); r["a"][2] += dxy * (y0 + y1)
        d = dir_of(L)
        assert key_of[succ[L]] == key_of[L]
        if d % 2:
            # a diagonal step is continued by a step of the same direction iff, at its end pixel q, the two neighbours after
            # the right-hand background pixel in the search order are empty and the next pixel on the diagonal is set:
            # SE: SW, S empty; NE: SE, E empty; NW: NE, N empty; SW: NW, W empty  (local: bit logic on three rows)
            e1, e2 = (d + 6) & 7, (d + 7) & 7
            local = not fg(x1 + DX[e1], y1 + DY[e1]) and not fg(x1 + DX[e2], y1 + DY[e2]) and fg(x1 + DX[d], y1 + DY[d])
            assert local == (dir_of(succ[L]) == d)
        if d % 2 == 0:
            r["axis"] += 1
        elif pred_dir[L] != d:                       # start of a diagonal run: follow it
            k, M = 1, L
            while dir_of(succ[M]) == d and succ[M] != L:
                M = succ[M]; k += 1
            r["diag"] += float(np.sqrt(np.float32(2 * k * k)))
    # start pixel: the smallest west / east edge (crack between a foreground pixel and the background pixel beside it)
    start = {}
    for y in range(H):
        for x in range(W):
            if not fg(x, y):
                continue
            for east, nx in ((0, x - 1), (1, x + 1)):
                if not fg(nx, y):
                    k = (int(fl[y + 1, x + 1]), int(bl[y + 1, nx + 1]))
                    start[k] = min(start.get(k, (1 << 62,)), (2 * (y * W + x) + east, x, y))
    out = []
    for k, r in rec.items():
        a00 = r["a"][0]
        hole = int(start[k][0] & 1)                  # a border that starts on an east edge is a hole border
        first = (start[k][1], start[k][2]) if not hole else (-1, -1)
        out.append((a00, r["a"][1], r["a"][2], r["axis"] + r["diag"], hole) + first)
    return sorted(out)


def oracle_records(img):
    contours, info = R.find_contours((img != 0).astype(np.uint8) * 255)
    out = []
    for c, inf in zip(contours, info):
        if int(inf[2]) <= 1 and len(c) <= 1:
            continue
        a00, a10, a01, per = R.contour_stats(c)
        first = (int(c[0][0]), int(c[0][1])) if not inf[0] else (-1, -1)   # cv2's first point of a hole border is not the pixel it was found at
        out.append((int(a00), int(a10), int(a01), float(per), int(inf[0])) + first)
    return sorted(out)


def euler_check(img):
    """Holes of every 8-connected component from corner counts: (Q1 - Q3 - 2 QD) / 4 == 1 - holes, where Q1 / Q3 / QD count the
    2x2 corners holding one / three / two diagonal pixels of the component (a component without holes needs no background
    labels: all its steps belong to its outer border)."""
    H, W = img.shape
    p = np.zeros((H + 2, W + 2), bool)
    p[1:-1, 1:-1] = img != 0
    fl, nf = ndimage.label(p, structure=np.ones((3, 3)))
    q = np.zeros((3, nf + 1), np.int64)
    for y in range(H + 1):
        for x in range(W + 1):
            blk = p[y:y + 2, x:x + 2]
            n = int(blk.sum())
            if n in (1, 3) or (n == 2 and blk[0, 0] == blk[1, 1]):
                q[{1: 0, 3: 1, 2: 2}[n], fl[y:y + 2, x:x + 2][blk][0]] += 1
    holes = np.zeros(nf + 1, np.int64)
    cs, hier = cv2.findContours((img != 0).astype(np.uint8), cv2.RETR_CCOMP, cv2.CHAIN_APPROX_NONE)
    if hier is not None:
        for c, h in zip(cs, hier[0]):
            if h[3] >= 0:                                   # a hole border runs over pixels of the component around it
                holes[fl[c[0][0][1] + 1, c[0][0][0] + 1]] += 1
    return np.array_equal((q[0] - q[1] - 2 * q[2])[1:], 4 * (1 - holes[1:]))


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    rng = np.random.default_rng(5)
    for it in range(n):
        H, W = int(rng.integers(3, 40)), int(rng.integers(3, 40))
        img = rng.random((H, W)) < rng.choice([0.2, 0.5, 0.8])
        if it % 3 == 0:                                       # smoother shapes
            img = cv2.blur(img.astype(np.float32), (5, 5)) > 0.5
        (a, rights), b = local_links(img), cv_links(img)
        if bitrow_links(img) != a:
            print("BIT-ROW MASK MISMATCH at image", it)
            return 1
        if a != b:
            print("MISMATCH at image", it, "only local:", list((a - b).items())[:5], "only cv:", list((b - a).items())[:5])
            np.save("/tmp/links_fail.npy", img)
            return 1
        if not per_border_check(img, a, rights):
            print("PER-BORDER MISMATCH at image", it)
            np.save("/tmp/links_fail.npy", img)
            return 1
        if not euler_check(img):
            print("EULER MISMATCH at image", it)
            return 1
        mine, ref = trace_free_records(img), oracle_records(img)
        if mine != ref:
            print("RECORD MISMATCH at image", it, [x for x in mine if x not in ref][:3], [x for x in ref if x not in mine][:3])
            np.save("/tmp/links_fail.npy", img)
            return 1
    print(n, "images: trace-free border records == oracle contour records; bit-row step masks == corner enumeration; holes per component == 1 - Euler number")
    print(n, "images: local link multiset == cv2 border steps; per-(fg, bg)-label sums == per-contour Green sums")
    return 0


if __name__ == "__main__":
    sys.exit(main())
