// Cluster path of _find_dot (the product path): after the streaming scan, every group of hot 32x32 source cells is
// handled as ONE unit entirely in shared memory -- remap + 5x5 floor-mean threshold + 5x5 majority on the group's
// output box, border starts from the rows' bit masks, Suzuki-Abe trace, Green sums -- and appends one record per outer
// border to the frame's list.  A small per-frame kernel then applies the reference's filter / centroid / output order
// (lib/ImageOperations.py:38-65).  No per-frame serial stage and no round trip of the binary image through HBM.
//
// Exactness: (1) a filtered pixel can only be set within reach of a source pixel > thresh, so every foreground pixel lies
// inside the output box of some hot cell; (2) a cluster is a connected component of the "boxes touch" graph over the hot
// cells, so two 8-adjacent foreground pixels always lie in boxes of the same cluster: a blob is covered by the boxes of
// exactly one cluster and no foreground pixel of another cluster lies inside them.  A cluster therefore filters the
// bounding box of its boxes and reports exactly the borders whose start pixel lies inside one of its own boxes (the
// bounding box may show parts of foreign blobs; they are never owned).  What a unit cannot decide locally (a hole
// border -> contour tree, a group larger than the largest size class, capacity overflows) flags the FRAME for the
// general per-frame path (detect_filter.cu + detect_blobs.cu), which recomputes it from the source frame.
#include "common.cuh"
#include "remap.cuh"
#include "walk.cuh"

#define CL_THREADS 128
#define HOT_MAX 1024            // hot cells per frame on the cluster path
#define CELLS_MAX 8192          // TX*TY limit of the dense cell -> slot map held in shared memory
#define ROOTS_MAX 512
#define HUGE_MAX 512            // largest cluster box edge (third size class, staged through global scratch)
#define HUGE_CTAS 48
#define CELL_EMPTY 0xffffffffu

struct ClusterWs {
    int* need_general;          // [n] frame takes the general path
    int* q_count;               // [0] small queue length, [1] large queue length, [2] small cursor, [3] large cursor
    int* q_small;               // [cap][4]: frame, x0 | y0 << 16, x1 | y1 << 16, member offset | count << 16  (box inclusive)
    int* q_large;
    int* q_huge;
    int q_cap;
    int q_caps[3];              // usable entries per size class (bounded by the bit-row storage)
    short* memb;                // [n][HOT_MAX][4] output boxes of the hot cells, grouped by cluster
    uint8_t* huge_scratch;      // [HUGE_CTAS][huge_scratch_stride]
    size_t huge_stride;
    uint32_t* rows_out;         // filtered bit rows of every cluster box: class 0 at item * 64 * 2 words, class 1 after them, ...
    size_t rows_base[3];        // word offset of each size class in rows_out
    int* cand_list;             // [cand_cap][6]: frame, x0 | y0 << 16, lx | ly << 16 | type << 31, rows word offset, mw | mh << 16, words per row
    int cand_cap;
    int* rec_count;             // [n]
    int* rec_start;             // [n][max_contours] start pixel index y * W + x of an outer border
    long long* rec_a;           // [n][max_contours][3] a00 a10 a01
    double* rec_per;            // [n][max_contours]
};

__device__ __forceinline__ int suf_find(int* parent, int x)
{
    int p = parent[x];
    while (p != x) {
        int g = parent[p];
        if (g != p) parent[x] = g;
        x = p; p = g;
    }
    return x;
}
__device__ __forceinline__ void suf_union(int* parent, int a, int b)
{
    for (;;) {
        a = suf_find(parent, a);
        b = suf_find(parent, b);
        if (a == b) return;
        if (a < b) { int t = a; a = b; b = t; }
        int old = atomicMin(&parent[a], b);
        if (old == a) return;
        a = old;
    }
}

__device__ __forceinline__ bool boxes_touch(const int* a, const int* b)
{
    return a[0] <= b[2] + 1 && b[0] <= a[2] + 1 && a[1] <= b[3] + 1 && b[1] <= a[3] + 1;
}

// ---------------------------------------------------------------------------------------------------------
// per frame: hot cells -> groups with pairwise separated output boxes -> work queues
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(CL_THREADS) form_clusters_kernel(const uint32_t* __restrict__ cellbox, TableView tv, ClusterWs cw)
{
    DYN_SHARED(smraw);
    const int f = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
    const int TX = tv.TX, TY = tv.TY, cells = TX * TY, W = tv.W, H = tv.H;
    if (cw.need_general[f]) return;
    // carve: idx_of[cells] u16 | hot_cell[HOT_MAX] u16 | parent[HOT_MAX] int | box[HOT_MAX][4] short | cbox[HOT_MAX][4] int | roots[ROOTS_MAX] u16
    uint16_t* idx_of = (uint16_t*)smraw;
    uint16_t* hot_cell = idx_of + ((cells + 7) & ~7);
    int* parent = (int*)(hot_cell + HOT_MAX);
    int* cbox = parent + HOT_MAX;
    short* box = (short*)(cbox + 4 * HOT_MAX);
    uint16_t* roots = (uint16_t*)(box + 4 * HOT_MAX);
    __shared__ int s_nhot, s_nroots, s_changed, s_bad;
    if (tid == 0) { s_nhot = 0; s_bad = 0; }
    for (int c = tid; c < cells; c += nt) idx_of[c] = 0xffff;
    __syncthreads();
    const uint32_t* cb = cellbox + (size_t)f * cells;
    for (int c = tid; c < cells; c += nt) {
        if (cb[c] != CELL_EMPTY) {
            int slot = atomicAdd(&s_nhot, 1);
            if (slot < HOT_MAX) { hot_cell[slot] = (uint16_t)c; idx_of[c] = (uint16_t)slot; }
        }
    }
    __syncthreads();
    const int n_hot = s_nhot;
    if (n_hot > HOT_MAX) { if (tid == 0) cw.need_general[f] = 2; return; }
    if (n_hot == 0) return;
    // output box of every hot cell: pixels whose filtered value can depend on the cell's hot pixels
    for (int h = tid; h < n_hot; h += nt) {
        int c = hot_cell[h], cy = c / TX, cx = c - cy * TX;
        uint32_t b = cb[c];
        const int32_t* inv = tv.cellinv + 4 * c;
        int hx0 = cx * 32 + (int)(b & 0xff), hx1 = cx * 32 + (int)((b >> 8) & 0xff);
        int hy0 = cy * 32 + (int)((b >> 16) & 0xff), hy1 = cy * 32 + (int)(b >> 24);
        int x0 = hx0 - 1 - inv[1] - 4, x1 = hx1 - inv[0] + 4, y0 = hy0 - 1 - inv[3] - 4, y1 = hy1 - inv[2] + 4;
        if (inv[0] > inv[1]) { x0 = 1; x1 = 0; }                 // no output pixel samples this cell
        x0 = max(x0, 0); y0 = max(y0, 0); x1 = min(x1, W - 1); y1 = min(y1, H - 1);
        if (x0 > x1 || y0 > y1) { x0 = 1; x1 = 0; y0 = 1; y1 = 0; }
        box[4 * h] = (short)x0; box[4 * h + 1] = (short)y0; box[4 * h + 2] = (short)x1; box[4 * h + 3] = (short)y1;
        parent[h] = h;
    }
    __syncthreads();
    // first grouping: hot 8-neighbour cells whose boxes touch
    for (int h = tid; h < n_hot; h += nt) {
        if (box[4 * h] > box[4 * h + 2]) continue;
        int c = hot_cell[h], cy = c / TX, cx = c - cy * TX;
        int a[4] = {box[4 * h], box[4 * h + 1], box[4 * h + 2], box[4 * h + 3]};
        const int ndx[4] = {-1, -1, 0, 1}, ndy[4] = {0, -1, -1, -1};
        for (int q = 0; q < 4; ++q) {
            int nx = cx + ndx[q], ny = cy + ndy[q];
            if (nx < 0 || ny < 0 || nx >= TX) continue;
            int k = idx_of[ny * TX + nx];
            if (k == 0xffff || box[4 * k] > box[4 * k + 2]) continue;
            int bb[4] = {box[4 * k], box[4 * k + 1], box[4 * k + 2], box[4 * k + 3]};
            if (boxes_touch(a, bb)) suf_union(parent, h, k);
        }
    }
    __syncthreads();
    // clusters = connected components of the "boxes touch" graph; members of a cluster are written contiguously
    for (int h = tid; h < n_hot; h += nt) {
        cbox[4 * h] = 0x7fffffff; cbox[4 * h + 1] = 0x7fffffff; cbox[4 * h + 2] = -1; cbox[4 * h + 3] = 0;   // [3] doubles as member count, see below
    }
    if (tid == 0) s_nroots = 0;
    __syncthreads();
    // bounding box per root: x0, y0, x1 in cbox[0..2]; y1 kept in a second pass to leave cbox[3] for the count
    for (int h = tid; h < n_hot; h += nt) {
        if (box[4 * h] > box[4 * h + 2]) continue;
        int r = suf_find(parent, h);
        parent[h] = r;
        atomicMin(&cbox[4 * r], (int)box[4 * h]); atomicMin(&cbox[4 * r + 1], (int)box[4 * h + 1]);
        atomicMax(&cbox[4 * r + 2], (int)box[4 * h + 2]);
        atomicAdd(&cbox[4 * r + 3], 1);
        if (r == h) { int k = atomicAdd(&s_nroots, 1); if (k < ROOTS_MAX) roots[k] = (uint16_t)h; }
    }
    __syncthreads();
    const int nr = s_nroots;
    if (nr > ROOTS_MAX) { if (tid == 0) cw.need_general[f] = 3; return; }
    // member offsets: exclusive scan of the counts over the root list (few hundred roots at most)
    if (tid == 0) {
        int acc = 0;
        for (int i = 0; i < nr; ++i) { int r = roots[i]; int c = cbox[4 * r + 3]; cbox[4 * r + 3] = acc | (c << 16); acc += c; }
    }
    __syncthreads();
    // one thread per cluster gathers its member boxes (a few hundred hot cells per frame at most) and emits the work item
    short* memb = cw.memb + (size_t)f * HOT_MAX * 4;
    for (int i = tid; i < nr; i += nt) {
        int r = roots[i];
        int off = cbox[4 * r + 3] & 0xffff, k = 0;
        int y1 = -1;
        for (int h = 0; h < n_hot; ++h) {
            if (box[4 * h] > box[4 * h + 2] || parent[h] != r) continue;
            short* m = memb + 4 * (off + k);
            m[0] = box[4 * h]; m[1] = box[4 * h + 1]; m[2] = box[4 * h + 2]; m[3] = box[4 * h + 3];
            y1 = max(y1, (int)box[4 * h + 3]);
            ++k;
        }
        int x0 = cbox[4 * r], y0 = cbox[4 * r + 1], x1 = cbox[4 * r + 2];
        int ew = x1 - x0 + 1, eh = y1 - y0 + 1, e = max(ew, eh);
        int* q; int* cnt;
        if (e <= 64) { q = cw.q_small; cnt = &cw.q_count[0]; }
        else if (e <= 128) { q = cw.q_large; cnt = &cw.q_count[1]; }
        else if (e <= HUGE_MAX) { q = cw.q_huge; cnt = &cw.q_count[4]; }
        else { s_bad = 1; continue; }
        int slot = atomicAdd(cnt, 1);
        if (slot >= cw.q_caps[e <= 64 ? 0 : (e <= 128 ? 1 : 2)]) { s_bad = 1; continue; }
        q[4 * slot] = f; q[4 * slot + 1] = x0 | (y0 << 16); q[4 * slot + 2] = x1 | (y1 << 16); q[4 * slot + 3] = off | (k << 16);
    }
    __syncthreads();
    if (s_bad && tid == 0) cw.need_general[f] = 5;      // a group larger than the largest class (or a full queue): general path
}

// ---------------------------------------------------------------------------------------------------------
// per cluster: filter the output box in shared memory, find and trace the outer borders it owns
// ---------------------------------------------------------------------------------------------------------
template <int MAXE>
struct ClusterDims {
    static constexpr int UW = MAXE + 8, BW = MAXE + 4, WPR = MAXE / 32;
    static constexpr size_t U_BYTES = ((size_t)UW * UW + 15) & ~(size_t)15;       // undistorted pixels, origin (mx0 - 4, my0 - 4)
    static constexpr size_t HS_BYTES = ((size_t)UW * BW * 2 + 15) & ~(size_t)15;  // horizontal 5-sums (U rows x B cols), reused for the majority
    static constexpr size_t B_BYTES = ((size_t)BW * BW + 15) & ~(size_t)15;       // thresholded floor-mean, origin (mx0 - 2, my0 - 2)
    static constexpr size_t ROWS_BYTES = (size_t)MAXE * WPR * 4;                  // filtered binary image of the box, bit x - mx0 of row y - my0
    static constexpr size_t IMG_BYTES = U_BYTES + HS_BYTES + B_BYTES + ROWS_BYTES;
};

// GLOBAL = false: the box arrays live in shared memory (size classes 64 and 128); true: in a per-CTA global scratch (<= 512)
template <int MAXE, bool GLOBAL>
__global__ void __launch_bounds__(CL_THREADS) cluster_proc_kernel(const uint8_t* __restrict__ frames, int64_t fstride, TableView tv, int thresh,
                                                                  ClusterWs cw, int which, int max_contours)
{
    DYN_SHARED(smraw);
    typedef ClusterDims<MAXE> DM;
    constexpr int UW = DM::UW, BW = DM::BW, WPR = DM::WPR;
    uint8_t* img = GLOBAL ? cw.huge_scratch + (size_t)blockIdx.x * cw.huge_stride : (uint8_t*)smraw;
    uint8_t* U = img;
    uint16_t* HS = (uint16_t*)(img + DM::U_BYTES);
    uint8_t* B = img + DM::U_BYTES + DM::HS_BYTES;
    uint32_t* rows = (uint32_t*)(img + DM::U_BYTES + DM::HS_BYTES + DM::B_BYTES);
    __shared__ int s_item, s_skip;
    const int tid = threadIdx.x, lane = tid & 31, wy = tid >> 5, NWARP = CL_THREADS / 32;
    const int H = tv.H, W = tv.W, T = thresh + 1;
    const int* queue = which == 0 ? cw.q_small : (which == 1 ? cw.q_large : cw.q_huge);
    const int i_len = which < 2 ? which : 4, i_cur = which < 2 ? 2 + which : 5;
    const int total = min(cw.q_count[i_len], cw.q_caps[which]);
    for (;;) {
        __syncthreads();
        if (tid == 0) {
            int it = atomicAdd(&cw.q_count[i_cur], 1);
            s_item = it;
            s_skip = it < total ? cw.need_general[queue[4 * it]] : 0;      // frame already handed to the general path
        }
        __syncthreads();
        const int item = s_item;
        if (item >= total) break;
        if (s_skip) continue;
        const int f = queue[4 * item];
        const int mx0 = queue[4 * item + 1] & 0xffff, my0 = queue[4 * item + 1] >> 16;
        const int mx1 = queue[4 * item + 2] & 0xffff, my1 = queue[4 * item + 2] >> 16;
        const int m_off = queue[4 * item + 3] & 0xffff, m_cnt = queue[4 * item + 3] >> 16;
        const int mw = mx1 - mx0 + 1, mh = my1 - my0 + 1;
        const int uw = mw + 8, uh = mh + 8, bw = mw + 4, bh = mh + 4;
        const uint8_t* fr = frames + (size_t)f * fstride;
        // ---- 1. undistorted pixels of the box dilated by 4 (zero outside the frame) ----------------------------
        for (int r = wy; r < uh; r += NWARP) {
            int i = my0 - 4 + r;
            bool rowin = (unsigned)i < (unsigned)H;
            for (int c = lane; c < uw; c += 32) {
                int j = mx0 - 4 + c, u = 0;
                if (rowin && (unsigned)j < (unsigned)W) u = remap_px(fr, W, H, i, j, (uint32_t)tv.map[(size_t)i * W + j]);
                U[r * UW + c] = (uint8_t)u;
            }
        }
        __syncthreads();
        // ---- 2. horizontal 5-sums -------------------------------------------------------------------------------
        for (int r = wy; r < uh; r += NWARP)
            for (int c = lane; c < bw; c += 32) {
                const uint8_t* u = &U[r * UW + c];
                HS[r * BW + c] = (uint16_t)(u[0] + u[1] + u[2] + u[3] + u[4]);
            }
        __syncthreads();
        // ---- 3. floor-mean over the in-frame taps > thresh  <=>  sum >= T * count ------------------------------------
        for (int r = wy; r < bh; r += NWARP) {
            int i = my0 - 2 + r;
            int cnty = min(i + 2, H - 1) - max(i - 2, 0) + 1;
            bool rowin = (unsigned)i < (unsigned)H;
            for (int c = lane; c < bw; c += 32) {
                int j = mx0 - 2 + c, b = 0;
                if (rowin && (unsigned)j < (unsigned)W) {
                    const uint16_t* h = &HS[r * BW + c];
                    int s = h[0] + h[BW] + h[2 * BW] + h[3 * BW] + h[4 * BW];
                    int cnt = cnty * (min(j + 2, W - 1) - max(j - 2, 0) + 1);
                    b = s >= T * cnt;
                }
                B[r * BW + c] = (uint8_t)b;
            }
        }
        __syncthreads();
        // ---- 4. 5x5 majority with replicated frame border: horizontal sums (clamped columns) into HS ----------------------
        for (int r = wy; r < bh; r += NWARP) {
            const uint8_t* b = &B[r * BW] - (mx0 - 2);                 // indexed by frame column
            for (int c = lane; c < mw; c += 32) {
                int j = mx0 + c;
                int s = b[max(j - 2, 0)] + b[max(j - 1, 0)] + b[j] + b[min(j + 1, W - 1)] + b[min(j + 2, W - 1)];
                HS[r * BW + c] = (uint16_t)s;
            }
        }
        __syncthreads();
        //      vertical sums (clamped rows) -> bit rows via ballot
        for (int r = wy; r < mh; r += NWARP) {
            int i = my0 + r;
            int r0 = max(i - 2, 0) - (my0 - 2), r1 = max(i - 1, 0) - (my0 - 2), r2 = i - (my0 - 2);
            int r3 = min(i + 1, H - 1) - (my0 - 2), r4 = min(i + 2, H - 1) - (my0 - 2);
            for (int c0 = 0; c0 < MAXE; c0 += 32) {
                int c = c0 + lane, s = 0;
                if (c < mw) s = HS[r0 * BW + c] + HS[r1 * BW + c] + HS[r2 * BW + c] + HS[r3 * BW + c] + HS[r4 * BW + c];
                unsigned wv = __ballot_sync(0xffffffffu, s >= 13);
                if (lane == 0) rows[r * WPR + (c0 >> 5)] = wv;
            }
        }
        __syncthreads();
        // ---- 5. export the bit rows; border-start candidates, one thread per row: a run with no 8-neighbour above starts
        //         an outer border, a gap between two runs that is completely covered from above starts a hole border
        //         (necessary conditions; the trace kernel verifies them).  Only candidates the cluster owns are kept:
        //         start pixel inside one of its own cell boxes. ----------------------------------------------------------------
        const size_t rows_off = cw.rows_base[which] + (size_t)item * MAXE * WPR;
        for (int k = tid; k < mh * WPR; k += CL_THREADS) cw.rows_out[rows_off + k] = rows[k];
        BitImg im; im.p = rows; im.W = mw; im.H = mh; im.WPR = WPR;
        const short* memb = cw.memb + ((size_t)f * HOT_MAX + m_off) * 4;
        for (int r = tid; r < mh; r += CL_THREADS) {
            int prev = 0, run_start = -1, last_end = -2;
            for (int x = 0; x <= mw; ++x) {
                int cur = x < mw ? (int)((rows[r * WPR + (x >> 5)] >> (x & 31)) & 1u) : 0;
                int cx = -1, cty = 0;
                if (cur && !prev) {
                    run_start = x;
                    if (last_end >= 0 && r > 0) {
                        bool covered = true;
                        for (int g = last_end + 1; g < x; ++g) if (!im.get(g, r - 1)) { covered = false; break; }
                        if (covered) { cx = last_end; cty = 1; }
                    }
                }
                if (!cur && prev) {
                    int xe = x - 1;
                    bool top = true;
                    if (r > 0) for (int g = run_start - 1; g <= xe + 1; ++g) if (im.get(g, r - 1)) { top = false; break; }
                    if (top) { cx = run_start; cty = 0; }
                    last_end = xe;
                }
                prev = cur;
                if (cx < 0) continue;
                int ax = cx + mx0, ay = r + my0;
                bool own = false;
                for (int m = 0; m < m_cnt && !own; ++m)
                    own = ax >= memb[4 * m] && ax <= memb[4 * m + 2] && ay >= memb[4 * m + 1] && ay <= memb[4 * m + 3];
                if (!own) continue;
                int slot = atomicAdd(&cw.q_count[6], 1);
                if (slot >= cw.cand_cap) { cw.need_general[f] = 7; continue; }
                int* e = cw.cand_list + 6 * (size_t)slot;
                e[0] = f; e[1] = mx0 | (my0 << 16); e[2] = cx | (r << 16) | (cty << 31);
                e[3] = (int)rows_off; e[4] = mw | (mh << 16); e[5] = WPR;
            }
        }
    }
}

// one thread per owned border-start candidate: Suzuki-Abe trace on the cluster's exported bit rows; a completed outer
// border becomes a record of its frame, a completed hole border sends the frame to the general path (contour tree)
__global__ void __launch_bounds__(128) trace_candidates_kernel(ClusterWs cw, int W, int max_contours)
{
    const int total = min(cw.q_count[6], cw.cand_cap);
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < total; c += gridDim.x * blockDim.x) {
        const int* e = cw.cand_list + 6 * (size_t)c;
        const int f = e[0];
        if (cw.need_general[f]) continue;
        const int mx0 = e[1] & 0xffff, my0 = e[1] >> 16;
        const int lx = e[2] & 0xffff, ly = (e[2] >> 16) & 0x7fff, ty = (e[2] >> 31) & 1;
        BitImg im; im.p = cw.rows_out + (unsigned)e[3]; im.W = e[4] & 0xffff; im.H = e[4] >> 16; im.WPR = e[5];
        long long st = (long long)(ly + my0) * W + (lx + mx0);
        long long a[3]; double per; int nch, ovf = 0;
        int ok = trace_contour(im, lx, ly, ty ? 0 : 4, 2 * st + ty, a, &per, &nch, &ovf, mx0, my0, W);
        if (ovf || (ok && ty)) { cw.need_general[f] = 8; continue; }           // hole border -> contour tree -> general path
        if (!ok) continue;
        int slot = atomicAdd(&cw.rec_count[f], 1);
        if (slot < max_contours) {
            cw.rec_start[(size_t)f * max_contours + slot] = (int)st;
            long long* ra = cw.rec_a + ((size_t)f * max_contours + slot) * 3;
            ra[0] = a[0]; ra[1] = a[1]; ra[2] = a[2];
            cw.rec_per[(size_t)f * max_contours + slot] = per;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// per frame: the reference's filter, centroid and output order on the frame's border records
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(CL_THREADS) finalize_kernel(ClusterWs cw, int max_contours, int max_blobs, double min_area, double min_circ,
                                                              int32_t* __restrict__ out_xy, int32_t* __restrict__ out_count, int32_t* __restrict__ out_flags,
                                                              double* __restrict__ out_contours, int32_t* __restrict__ out_contour_count)
{
    DYN_SHARED(smraw);
    uint8_t* keepv = (uint8_t*)smraw;                       // [max_contours]
    const int f = blockIdx.x, tid = threadIdx.x;
    if (cw.need_general[f]) return;                         // outputs come from the general path
    const int n = cw.rec_count[f];
    __shared__ int s_kept;
    if (tid == 0) s_kept = 0;
    __syncthreads();
    if (n > max_contours) {
        if (tid == 0) { out_flags[f] |= MOCAP_FLAG_CONTOUR_OVERFLOW; out_count[f] = 0; if (out_contour_count) out_contour_count[f] = 0; }
        return;
    }
    const int* start = cw.rec_start + (size_t)f * max_contours;
    const long long* ra = cw.rec_a + (size_t)f * max_contours * 3;
    const double* rper = cw.rec_per + (size_t)f * max_contours;
    for (int c = tid; c < n; c += CL_THREADS) {
        long long a00 = ra[3 * c];
        double area = (double)(a00 < 0 ? -a00 : a00) * 0.5, per = rper[c];
        int keep = 0;
        if (per != 0.0) {
            double circ = __ddiv_rn(__dmul_rn(12.566370614359172, area), __dmul_rn(per, per));
            keep = (circ > min_circ && area > min_area) ? 1 : 0;
        }
        if (a00 == 0) keep = 0;                              // moments["m00"] == 0 -> no centroid
        keepv[c] = (uint8_t)keep;
    }
    __syncthreads();
    for (int c = tid; c < n; c += CL_THREADS) {
        long long a00 = ra[3 * c], a10 = ra[3 * c + 1], a01 = ra[3 * c + 2];
        double per = rper[c];
        int keep = keepv[c];
        // all records are top-level outer borders: cv.findContours lists them in reverse raster order of their start pixel
        int st = start[c], rank = 0, pos = 0;
        for (int u = 0; u < n; ++u) {
            bool before = start[u] > st;
            rank += before;
            pos += before && keepv[u];
        }
        if (out_contours && rank < max_contours) {
            double* o = out_contours + ((size_t)f * max_contours + rank) * 8;
            o[0] = (double)a00; o[1] = (double)a10; o[2] = (double)a01; o[3] = per;
            o[4] = 0.0; o[5] = -1.0; o[6] = (double)keep; o[7] = (double)st;
        }
        if (!keep) continue;
        atomicAdd(&s_kept, 1);
        if (pos < max_blobs) {
            double sgn = a00 > 0 ? 1.0 : -1.0;
            double m00 = __dmul_rn((double)a00, sgn * 0.5);
            double m10 = __dmul_rn((double)a10, sgn * 0.16666666666666666);
            double m01 = __dmul_rn((double)a01, sgn * 0.16666666666666666);
            out_xy[((size_t)f * max_blobs + pos) * 2 + 0] = (int32_t)__ddiv_rn(m10, m00);
            out_xy[((size_t)f * max_blobs + pos) * 2 + 1] = (int32_t)__ddiv_rn(m01, m00);
        }
    }
    __syncthreads();
    if (tid == 0) {
        int kept = s_kept;
        if (kept > max_blobs) { out_flags[f] |= MOCAP_FLAG_BLOB_OVERFLOW; kept = max_blobs; }
        out_count[f] = kept;
        if (out_contour_count) out_contour_count[f] = n;
    }
}

// ---------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------
// queue capacities: every frame may contribute up to 192 small, 48 large and 4 huge clusters before it is sent to the general path
static void cluster_caps(int n, int* caps) { caps[0] = n * 192 + 256; caps[1] = n * 48 + 64; caps[2] = n * 4 + 16; }
static int cluster_cand_cap(int n) { return n * 512 + 1024; }
static size_t cluster_rows_words(int n, int q_cap, size_t* base)
{
    (void)q_cap;
    int caps[3]; cluster_caps(n, caps);
    size_t b0 = 0, b1 = b0 + (size_t)caps[0] * 64 * 2, b2 = b1 + (size_t)caps[1] * 128 * 4, end = b2 + (size_t)caps[2] * HUGE_MAX * (HUGE_MAX / 32);
    if (base) { base[0] = b0; base[1] = b1; base[2] = b2; }
    return end;
}

size_t cluster_ws_bytes(int n, int max_contours, int q_cap, size_t* offs /*[16]*/)
{
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t r = off; off += align_up(bytes, 256); return r; };
    offs[0] = take((size_t)n * 4);                         // need_general
    offs[1] = take(64);                                    // q_count
    offs[2] = take((size_t)q_cap * 16);                    // q_small
    offs[3] = take((size_t)q_cap * 16);                    // q_large
    offs[4] = take((size_t)n * 4);                         // rec_count
    offs[5] = take((size_t)n * max_contours * 4);          // rec_start
    offs[6] = take((size_t)n * max_contours * 24);         // rec_a
    offs[7] = take((size_t)n * max_contours * 8);          // rec_per
    offs[8] = take((size_t)q_cap * 16);                    // q_huge
    offs[9] = take((size_t)n * HOT_MAX * 8);               // memb
    offs[10] = take(align_up(ClusterDims<HUGE_MAX>::IMG_BYTES, 256) * HUGE_CTAS);   // huge scratch
    offs[11] = take(cluster_rows_words(n, q_cap, nullptr) * 4);                      // rows_out
    offs[12] = take((size_t)cluster_cand_cap(n) * 24);                               // cand_list
    return off;
}

bool cluster_path_supported(int H, int W)
{
    return cdiv(W, TILE) * cdiv(H, TILE) <= CELLS_MAX && W <= 32767 && H <= 32767;
}

int launch_cluster_path(const uint8_t* frames, int n, int H, int W, int64_t fstride, const TableView& tv, int thresh,
                        const uint32_t* cellbox, char* ws_base, const size_t* offs, int q_cap,
                        int max_contours, int max_blobs, double min_area, double min_circ,
                        int32_t* out_xy, int32_t* out_count, int32_t* out_flags, double* out_contours, int32_t* out_contour_count,
                        bool finalize_only, cudaStream_t s, StageTimer* timer)
{
    ClusterWs cw;
    cw.need_general = (int*)(ws_base + offs[0]);
    cw.q_count = (int*)(ws_base + offs[1]);
    cw.q_small = (int*)(ws_base + offs[2]);
    cw.q_large = (int*)(ws_base + offs[3]);
    cw.q_huge = (int*)(ws_base + offs[8]);
    cw.q_cap = q_cap;
    cw.rec_count = (int*)(ws_base + offs[4]);
    cw.rec_start = (int*)(ws_base + offs[5]);
    cw.rec_a = (long long*)(ws_base + offs[6]);
    cw.rec_per = (double*)(ws_base + offs[7]);
    cw.memb = (short*)(ws_base + offs[9]);
    cw.huge_scratch = (uint8_t*)(ws_base + offs[10]);
    cw.huge_stride = align_up(ClusterDims<HUGE_MAX>::IMG_BYTES, 256);
    cw.rows_out = (uint32_t*)(ws_base + offs[11]);
    cluster_rows_words(n, q_cap, cw.rows_base);
    cw.cand_list = (int*)(ws_base + offs[12]);
    cw.cand_cap = cluster_cand_cap(n);
    cluster_caps(n, cw.q_caps);
    if (finalize_only) {
        LAUNCH(finalize_kernel, n, CL_THREADS, (size_t)max_contours + 16, s, cw, max_contours, max_blobs, min_area, min_circ, out_xy, out_count, out_flags,
               out_contours, out_contour_count);
        CUDA_TRY(cudaGetLastError());
        return MOCAP_OK;
    }
    CUDA_TRY(cudaMemsetAsync(cw.q_count, 0, 64, s));
    CUDA_TRY(cudaMemsetAsync(cw.rec_count, 0, (size_t)n * 4, s));
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int cells = tv.TX * tv.TY;
    size_t sm_form = (size_t)((cells + 7) & ~7) * 2 + (size_t)HOT_MAX * (2 + 4 + 16 + 8) + (size_t)ROOTS_MAX * 2 + 64;
    size_t sm_small = ClusterDims<64>::IMG_BYTES;
    size_t sm_large = ClusterDims<128>::IMG_BYTES;
    size_t sm_huge = 16;
    auto k_small = cluster_proc_kernel<64, false>;
    auto k_large = cluster_proc_kernel<128, false>;
    auto k_huge = cluster_proc_kernel<HUGE_MAX, true>;
#ifndef MOCAP_EMU
    static bool attr_done = false;
    if (!attr_done) {
        cudaFuncSetAttribute(k_large, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_large);
        cudaFuncSetAttribute(form_clusters_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
        attr_done = true;
    }
#endif
    stage_begin(timer, 1, s);
    LAUNCH(form_clusters_kernel, n, CL_THREADS, sm_form, s, cellbox, tv, cw);
    stage_end(timer, 1, s);
    stage_begin(timer, 2, s);
    LAUNCH(k_small, sms * 8, CL_THREADS, sm_small, s, frames, fstride, tv, thresh, cw, 0, max_contours);
    LAUNCH(k_large, sms * 2, CL_THREADS, sm_large, s, frames, fstride, tv, thresh, cw, 1, max_contours);
    LAUNCH(k_huge, HUGE_CTAS, CL_THREADS, sm_huge, s, frames, fstride, tv, thresh, cw, 2, max_contours);
    LAUNCH(trace_candidates_kernel, sms * 4, 128, 0, s, cw, W, max_contours);
    stage_end(timer, 2, s);
    CUDA_TRY(cudaGetLastError());
    return MOCAP_OK;
}
