// Suzuki-Abe border following on a packed binary image (shared by the per-frame general path, detect_blobs.cu, and the
// per-cluster path, detect_cluster.cu).  Restates what cv.findContours(RETR_TREE, CHAIN_APPROX_SIMPLE) does per border
// (lib/ImageOperations.py:41) plus the Green sums / perimeter of cv.moments / cv.arcLength (:44-45, :59).
#pragma once
#include "common.cuh"

// Walker state for Suzuki-Abe border following started from the west (side 4) or east (side 0) edge of a pixel.
struct Walk {
    int x0, y0;      // start pixel
    int x1, y1;      // its predecessor on the border
    int x, y;        // current pixel
    int s;           // direction from current pixel to the previous one
    bool single;
};

__device__ __forceinline__ void walk_init(const BitImg& im, Walk& w, int x, int y, int side)
{
    w.x0 = x; w.y0 = y; w.x = x; w.y = y;
    // clockwise search (side-1, side-2, ... side-7) for the border's predecessor of the start pixel
    uint32_t m = im.nbr8(x, y);
    uint32_t rev = __brev(m) >> 24;                               // bit i = m[7 - i]
    uint32_t sh = (8 - side) & 7;
    uint32_t r = (((rev | (rev << 8)) >> sh) & 0x7fu);            // bit j = m[(side - 1 - j) & 7], j = 0..6
    w.single = r == 0;
    int s = w.single ? (side + 1) & 7 : (side - 1 - (__ffs(r) - 1)) & 7;
    w.s = s;
    w.x1 = x + dir_dx(s); w.y1 = y + dir_dy(s);
}

// One step: counter-clockwise search from the previous pixel; returns the step direction, `zeros` = bitmask of
// neighbour directions examined and found empty.  Advances the walker.  `done` when back at the start.
__device__ __forceinline__ int walk_step(const BitImg& im, Walk& w, unsigned& zeros, bool& done)
{
    uint32_t m = im.nbr8(w.x, w.y);
    int start = (w.s + 1) & 7;
    uint32_t rot = ((m | (m << 8)) >> start) & 0xffu;             // bit k = m[(start + k) & 7]
    int k = rot ? __ffs(rot) - 1 : 8;
    int d = (start + k) & 7;
    uint32_t z = ((1u << k) - 1u) << start;                       // the k empty directions passed over
    zeros = (z | (z >> 8)) & 0xffu;
    int nx = w.x + dir_dx(d), ny = w.y + dir_dy(d);
    done = (nx == w.x0 && ny == w.y0 && w.x == w.x1 && w.y == w.y1);
    w.x = nx; w.y = ny;
    w.s = (d + 4) & 7;
    return d;
}

#define WALK_BUDGET (1 << 22)

// key of a west/east edge: 2 * pixel index + (east ? 1 : 0)
__device__ __forceinline__ long long edge_key(int x, int y, int W, int east) { return 2LL * ((long long)y * W + x) + east; }

// Walk the border owning edge (x, y, side).  mode 0: stop as soon as a smaller west/east edge of the same border is
// seen (returns 1 if the start edge is the border's smallest edge).  mode 1: full loop, *min_key = smallest edge key.
static __device__ int walk_min_edge(const BitImg& im, int x, int y, int side, int mode, long long* min_key, int* overflow)
{
    const int W = im.W;
    long long key0 = edge_key(x, y, W, side == 0);
    long long best = key0;
    Walk w; walk_init(im, w, x, y, side);
    if (w.single) { if (min_key) *min_key = edge_key(x, y, W, 0); return side == 4; }   // isolated pixel: its west edge is smaller
    for (int step = 0; step < WALK_BUDGET; ++step) {
        int cx = w.x, cy = w.y;
        unsigned zeros; bool done;
        walk_step(im, w, zeros, done);
        if (zeros & (1u << 4)) { long long k = edge_key(cx, cy, W, 0); if (k < best) { best = k; if (!mode) return 0; } }
        if (zeros & (1u << 0)) { long long k = edge_key(cx, cy, W, 1); if (k < best) { best = k; if (!mode) return 0; } }
        if (done) { if (min_key) *min_key = best; return best == key0; }
    }
    *overflow = 1;
    if (min_key) *min_key = best;
    return 0;
}

// Full trace of a border from its start edge: Green sums over the CHAIN_APPROX_SIMPLE vertices, perimeter, length.
// key0 >= 0: also verify that the start edge is the border's smallest west/east edge (i.e. where cv.findContours starts
// it); returns 0 as soon as a smaller one is met.  key0 < 0: no check.  Returns 1 for a completed trace.
// (ox, oy, Wabs): the image `im` may be a window of the frame (cluster path): vertices and edge keys are taken in frame
// coordinates (x + ox, y + oy) of a frame of width Wabs.
static __device__ int trace_contour(const BitImg& im, int x, int y, int side, long long key0, long long* a, double* per, int* n_chain, int* overflow,
                             int ox, int oy, int Wabs, int* bbox = nullptr)
{
    const int W = Wabs;
    Walk w; walk_init(im, w, x, y, side);
    if (bbox) { bbox[0] = bbox[2] = x + ox; bbox[1] = bbox[3] = y + oy; }
    if (w.single) { a[0] = a[1] = a[2] = 0; *per = 0.0; *n_chain = 1; return key0 < 0 || side == 4; }
    long long a00 = 0, a10 = 0, a01 = 0;
    double perim = 0.0;
    int n = 0;
    int prev_dir = w.s ^ 4;             // direction of the closing step (predecessor -> start)
    bool have_v = false;
    int vx = 0, vy = 0, fx = 0, fy = 0; // previous vertex, first vertex
    for (int step = 0; step < WALK_BUDGET; ++step) {
        int cx = w.x + ox, cy = w.y + oy;
        unsigned zeros; bool done;
        int d = walk_step(im, w, zeros, done);
        ++n;
        if (key0 >= 0) {
            if ((zeros & (1u << 4)) && edge_key(cx, cy, W, 0) < key0) return 0;
            if ((zeros & (1u << 0)) && edge_key(cx, cy, W, 1) < key0) return 0;
        }
        if (d != prev_dir) {            // direction change: (cx, cy) is a vertex
            if (have_v) {
                long long dxy = (long long)vx * cy - (long long)cx * vy;
                a00 += dxy; a10 += dxy * (vx + cx); a01 += dxy * (vy + cy);
                float ddx = (float)(cx - vx), ddy = (float)(cy - vy);
                perim += (double)__fsqrt_rn(__fadd_rn(__fmul_rn(ddx, ddx), __fmul_rn(ddy, ddy)));
            } else { fx = cx; fy = cy; have_v = true; }
            vx = cx; vy = cy;
            if (bbox) { bbox[0] = min(bbox[0], cx); bbox[1] = min(bbox[1], cy); bbox[2] = max(bbox[2], cx); bbox[3] = max(bbox[3], cy); }
        }
        prev_dir = d;
        if (done) {
            if (have_v) {               // closing segment last vertex -> first vertex
                long long dxy = (long long)vx * fy - (long long)fx * vy;
                a00 += dxy; a10 += dxy * (vx + fx); a01 += dxy * (vy + fy);
                float ddx = (float)(fx - vx), ddy = (float)(fy - vy);
                float q = __fadd_rn(__fmul_rn(ddx, ddx), __fmul_rn(ddy, ddy));
                if (q > 0.f) perim += (double)__fsqrt_rn(q);
            }
            a[0] = a00; a[1] = a10; a[2] = a01; *per = perim; *n_chain = n;
            return 1;
        }
    }
    *overflow = 1;
    a[0] = a[1] = a[2] = 0; *per = 0.0; *n_chain = n;
    return 0;
}


// trace_contour for boxes at most 64 pixels wide and 1024 high (two 32-bit words per row).  Same results, cheaper steps:
//  * the three rows around the walker live in registers as 64-bit words: a step is shifts and logic, one 8-byte load when
//    the walker changes row;
//  * Green sums over box-local vertex coordinates with 32-bit products (|dxy| < 2^16, factor < 2^11), translated at the end:
//    a10 = a10' + 3 ox a00, a01 = a01' + 3 oy a00 (exact integer identity for a closed polygon);
//  * perimeter: CHAIN_APPROX_SIMPLE segments are axis-parallel (length = an integer, exact in float32) or diagonal
//    (length = float32 sqrt(2 k^2)); the double sum of such float32 values is exact in any order, so axis lengths are summed
//    as integers, unit diagonals are counted, longer diagonals added one by one.
//
// WINDOW = true: the same trace on a sliding 64-pixel wide window of a wider box (wpr words per row, mw > 64 pixels wide;
// a blob of a wide cluster box is usually narrow).  The window starts at column wx0 (0 <= wx0 <= mw - 64); (x, ox) and all
// vertex coordinates are relative to wx0.  When the walker reaches a window column whose outer neighbour is not known, the
// window is re-centred on it (three row loads).  Coordinates then exceed 64, so the a10/a01 products are taken in 64 bits.
template <bool WINDOW>
static __device__ int trace_contour64(const uint32_t* __restrict__ rows, int mh, int wpr, int wx0, int mw, int x, int y, int side, long long key0,
                                      long long* a, double* per, int* n_chain, int* overflow, int ox, int oy, int Wabs, int* bbox)
{
    int wofs = 0;                                   // window origin relative to wx0 (WINDOW only)
    int wi0 = wx0 >> 5, wsh = wx0 & 31;
    auto load = [&](int yy) -> unsigned long long {
        if ((unsigned)yy >= (unsigned)mh) return 0ull;
        if (!WINDOW) return *(const unsigned long long*)(rows + 2 * yy);                            // rows of a two-word box are 8-byte aligned
        const uint32_t* r = rows + (size_t)yy * wpr + wi0;                                          // origin + 63 < mw: words wi0, wi0 + 1 exist
        const uint32_t w0 = r[0], w1 = r[1], w2 = (wsh && wi0 + 2 < wpr) ? r[2] : 0u;
        const unsigned long long lo = (unsigned long long)w0 | ((unsigned long long)w1 << 32);
        return wsh ? (lo >> wsh) | ((unsigned long long)w2 << (64 - wsh)) : lo;
    };
    auto nbr = [&](unsigned long long up, unsigned long long mid, unsigned long long dn, int xx) -> uint32_t {
        uint32_t u, m, d;
        if (xx) { u = (uint32_t)(up >> (xx - 1)) & 7u; m = (uint32_t)(mid >> (xx - 1)) & 7u; d = (uint32_t)(dn >> (xx - 1)) & 7u; }
        else { u = ((uint32_t)up << 1) & 7u; m = ((uint32_t)mid << 1) & 7u; d = ((uint32_t)dn << 1) & 7u; }
        return ((m >> 2) & 1u) | (((u >> 2) & 1u) << 1) | (((u >> 1) & 1u) << 2) | ((u & 1u) << 3) |
               ((m & 1u) << 4) | ((d & 1u) << 5) | (((d >> 1) & 1u) << 6) | (((d >> 2) & 1u) << 7);
    };
    unsigned long long up = load(y - 1), mid = load(y), dn = load(y + 1);
    if (bbox) { bbox[0] = bbox[2] = x + ox; bbox[1] = bbox[3] = y + oy; }
    // start: clockwise search for the predecessor (see walk_init)
    uint32_t m0 = nbr(up, mid, dn, x);
    uint32_t rev = __brev(m0) >> 24, sh = (8 - side) & 7;
    uint32_t r0 = (((rev | (rev << 8)) >> sh) & 0x7fu);
    if (r0 == 0) { a[0] = a[1] = a[2] = 0; *per = 0.0; *n_chain = 1; return key0 < 0 || side == 4; }
    int s = (side - 1 - (__ffs(r0) - 1)) & 7;
    const int x0 = x, y0 = y, x1 = x + dir_dx(s), y1 = y + dir_dy(s);
    // edge keys in 32 bits: key = 2 * (row * Wabs + col) + east; the frame sizes the library accepts keep this below 2^31
    const int kbase = 2 * (oy * Wabs + ox);
    const int key0i = (int)key0;
    long long a00 = 0, a10 = 0, a01 = 0;
    int axis_len = 0, n_diag1 = 0;
    double perim_long = 0.0;
    int n = 0, prev_dir = s ^ 4;
    bool have_v = false;
    int vx = 0, vy = 0, fx = 0, fy = 0;          // previous / first vertex, box-local
    int bx0 = x, bx1 = x, by0 = y, by1 = y;
    int wx = x, wy = y;
    for (int step = 0; step < WALK_BUDGET; ++step) {
        const int cx = wx, cy = wy;
        uint32_t m = nbr(up, mid, dn, wx - wofs);
        int start = (s + 1) & 7;
        uint32_t rot = ((m | (m << 8)) >> start) & 0xffu;
        int k = rot ? __ffs(rot) - 1 : 8;
        int d = (start + k) & 7;
        uint32_t z = ((1u << k) - 1u) << start;
        uint32_t zeros = (z | (z >> 8)) & 0xffu;
        int nx = wx + dir_dx(d), ny = wy + dir_dy(d);
        bool done = (nx == x0 && ny == y0 && wx == x1 && wy == y1);
        bool slide = false;
        if (WINDOW) {
            const int bx = nx - wofs, org = wx0 + wofs;              // window column of the next pixel, window origin in the box
            slide = (bx == 0 && org > 0) || (bx == 63 && org + 64 < mw);
            if (slide) {
                const int norg = min(max(wx0 + nx - 32, 0), mw - 64);
                wofs = norg - wx0; wi0 = norg >> 5; wsh = norg & 31;
                up = load(ny - 1); mid = load(ny); dn = load(ny + 1);
            }
        }
        if (!slide && ny != wy) {
            // one block for both directions: the lanes of a warp that step up and those that step down share the row load
            const bool goup = ny < wy;
            const unsigned long long nr = load(goup ? ny - 1 : ny + 1), om = mid;
            mid = goup ? up : dn;
            up = goup ? nr : om;
            dn = goup ? om : nr;
        }
        wx = nx; wy = ny; s = (d + 4) & 7;
        ++n;
        if (key0 >= 0 && (zeros & 0x11u)) {
            int kk = kbase + 2 * (cy * Wabs + cx);
            if ((zeros & (1u << 4)) && kk < key0i) return 0;
            if ((zeros & (1u << 0)) && kk + 1 < key0i) return 0;
        }
        if (d != prev_dir) {
            if (have_v) {
                int dxy = vx * cy - cx * vy;
                a00 += dxy;
                if (WINDOW) { a10 += (long long)dxy * (vx + cx); a01 += (long long)dxy * (vy + cy); }
                else { a10 += dxy * (vx + cx); a01 += dxy * (vy + cy); }
                int adx = abs(cx - vx), ady = abs(cy - vy);
                if (adx == 0 || ady == 0) axis_len += adx + ady;
                else if (adx == 1) ++n_diag1;
                else perim_long += (double)__fsqrt_rn((float)(2 * adx * adx));
            } else { fx = cx; fy = cy; have_v = true; }
            vx = cx; vy = cy;
            bx0 = min(bx0, cx); bx1 = max(bx1, cx); by0 = min(by0, cy); by1 = max(by1, cy);
        }
        prev_dir = d;
        if (done) {
            if (have_v) {
                int dxy = vx * fy - fx * vy;
                a00 += dxy;
                if (WINDOW) { a10 += (long long)dxy * (vx + fx); a01 += (long long)dxy * (vy + fy); }
                else { a10 += dxy * (vx + fx); a01 += dxy * (vy + fy); }
                int adx = abs(fx - vx), ady = abs(fy - vy);
                if (adx == 0 || ady == 0) axis_len += adx + ady;
                else if (adx == 1) ++n_diag1;
                else perim_long += (double)__fsqrt_rn((float)(2 * adx * adx));
            }
            a[0] = a00; a[1] = a10 + 3LL * ox * a00; a[2] = a01 + 3LL * oy * a00;
            *per = (double)axis_len + (double)n_diag1 * (double)__fsqrt_rn(2.0f) + perim_long;
            *n_chain = n;
            if (bbox) { bbox[0] = bx0 + ox; bbox[1] = by0 + oy; bbox[2] = bx1 + ox; bbox[3] = by1 + oy; }
            return 1;
        }
    }
    *overflow = 1;
    a[0] = a[1] = a[2] = 0; *per = 0.0; *n_chain = n;
    return 0;
}
