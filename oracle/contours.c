/* CPU oracle (TEST INFRASTRUCTURE ONLY -- never linked into the product).
 *
 * Restates what the reference obtains from OpenCV at lib/ImageOperations.py:41
 *     cv.findContours(grey, cv.RETR_TREE, cv.CHAIN_APPROX_SIMPLE)
 * OpenCV is a third-party dependency of the reference (unpinned there; opencv-python-headless
 * 4.13.0.92 in this image) and is not under /root/reference, so this follows the published
 * algorithm: S. Suzuki, K. Abe, "Topological structural analysis of digitized binary images by
 * border following", CVGIP 30 (1985), Algorithm 1, with the conventions OpenCV documents:
 *   - image treated as if surrounded by a 1-pixel zero frame,
 *   - 8-connected foreground, outer borders start at (f(i,j)=1, f(i,j-1)=0), hole borders at
 *     (f(i,j)>=1, f(i,j+1)=0),
 *   - CHAIN_APPROX_SIMPLE keeps a border pixel as vertex only where the step direction changes,
 *   - the contour list is the pre-order walk of the border tree, each new border linked in as the
 *     FIRST child of its parent (so siblings come out in reverse order of discovery).
 * Pinned against cv2 4.13.0 on random images in tests/test_oracle_cv2.py and on the committed
 * reference-generated fixtures in tests/golden/.
 *
 * Also: oracle_label8 = 8-connected component labelling (the "pixel membership" check).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* direction codes: 0=E 1=NE 2=N 3=NW 4=W 5=SW 6=S 7=SE (y grows downwards) */
static const int DX[8] = {1, 1, 0, -1, -1, -1, 0, 1};
static const int DY[8] = {0, -1, -1, -1, 0, 1, 1, 1};

typedef struct {
    int is_hole;
    int parent;      /* border id (0 = frame) */
    int first_child; /* linked list, most recently discovered first */
    int next_sibling;
    int pt_begin, pt_end;
    int n_chain;
} border_t;

int oracle_find_contours(const uint8_t *bin, int H, int W,
                         int32_t *pts, int cap_pts,
                         int32_t *offsets, int32_t *info, int cap_cnt,
                         int32_t *n_pts_out)
{
    const int PW = W + 2, PH = H + 2;
    int32_t *f = (int32_t *)calloc((size_t)PW * PH, sizeof(int32_t));
    if (!f) return -2;
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x)
            f[(y + 1) * PW + (x + 1)] = bin[y * W + x] ? 1 : 0;

    int cap_b = cap_cnt + 2;
    border_t *B = (border_t *)calloc((size_t)cap_b, sizeof(border_t));
    if (!B) { free(f); return -2; }
    /* border 1 = the frame (acts as a hole border, Suzuki Sect. 3) */
    B[1].is_hole = 1; B[1].parent = 0; B[1].first_child = 0; B[1].next_sibling = 0;
    int nbd = 1;
    int npts = 0;
    int overflow = 0;

    for (int y = 1; y <= H && !overflow; ++y) {
        int lnbd = 1;
        for (int x = 1; x <= W; ++x) {
            int32_t *p0 = &f[y * PW + x];
            int v = *p0;
            if (v == 0) continue;
            int is_hole;
            int from; /* direction of the background pixel the search starts from */
            if (v == 1 && p0[-1] == 0) { is_hole = 0; from = 4; }
            else if (v >= 1 && p0[1] == 0) { is_hole = 1; from = 0; if (v > 1) lnbd = v; }
            else { if (v != 1) lnbd = v < 0 ? -v : v; continue; }

            ++nbd;
            if (nbd >= cap_b) { overflow = 1; break; }
            border_t *b = &B[nbd];
            b->is_hole = is_hole;
            /* parent from the last border met on this row (Suzuki Table 1) */
            if (B[lnbd].is_hole == is_hole) b->parent = B[lnbd].parent; else b->parent = lnbd;
            b->first_child = 0;
            b->next_sibling = B[b->parent].first_child;
            B[b->parent].first_child = nbd;
            b->pt_begin = npts;
            b->n_chain = 0;

            /* (3.1) clockwise search around p0 starting after `from` */
            int s = from, found = 0;
            for (int k = 0; k < 7; ++k) {
                s = (s + 7) & 7;
                if (p0[DY[s] * PW + DX[s]] != 0) { found = 1; break; }
            }
            if (!found) {
                *p0 = -nbd;
                if (npts >= cap_pts) { overflow = 1; break; }
                pts[2 * npts] = x - 1; pts[2 * npts + 1] = y - 1; ++npts;
                b->n_chain = 1;
            } else {
                int32_t *p1 = p0 + DY[s] * PW + DX[s];
                int32_t *p3 = p0;
                int cx = x, cy = y;
                int prev_dir = s ^ 4; /* direction of the closing step p1 -> p0 */
                for (;;) {
                    /* (3.3) counter-clockwise search starting after the previous pixel's direction */
                    int s_start = s, east_zero_seen = 0, d = s;
                    int32_t *p4 = 0;
                    for (int k = 0; k < 8; ++k) {
                        d = (d + 1) & 7;
                        p4 = p3 + DY[d] * PW + DX[d];
                        if (*p4 != 0) break;
                        if (d == 0) east_zero_seen = 1;
                    }
                    (void)s_start;
                    /* (3.4) marks */
                    if (east_zero_seen) *p3 = -nbd;
                    else if (*p3 == 1) *p3 = nbd;
                    /* CHAIN_APPROX_SIMPLE vertex */
                    if (d != prev_dir) {
                        if (npts >= cap_pts) { overflow = 1; break; }
                        pts[2 * npts] = cx - 1; pts[2 * npts + 1] = cy - 1; ++npts;
                    }
                    b->n_chain++;
                    prev_dir = d;
                    cx += DX[d]; cy += DY[d];
                    if (p4 == p0 && p3 == p1) break; /* (3.5) back at the start */
                    p3 = p4;
                    s = (d + 4) & 7;
                }
                if (overflow) break;
            }
            b->pt_end = npts;
            /* (4) */
            v = *p0;
            if (v != 1) lnbd = v < 0 ? -v : v;
        }
    }

    int ret;
    if (overflow) {
        ret = -1;
    } else {
        /* pre-order walk, children in first_child -> next_sibling order */
        int n_out = 0;
        int *stack = (int *)malloc(sizeof(int) * (size_t)(nbd + 2));
        int *out_index = (int *)malloc(sizeof(int) * (size_t)(nbd + 2));
        int *order = (int *)malloc(sizeof(int) * (size_t)(nbd + 2));
        int sp = 0;
        /* push frame's children in reverse so that first_child pops first */
        {
            int cnt = 0;
            for (int c = B[1].first_child; c; c = B[c].next_sibling) order[cnt++] = c;
            for (int k = cnt - 1; k >= 0; --k) stack[sp++] = order[k];
        }
        int total_pts = 0;
        int32_t *tmp = (int32_t *)malloc(sizeof(int32_t) * 2 * (size_t)(npts > 0 ? npts : 1));
        out_index[1] = -1;
        while (sp > 0) {
            int c = stack[--sp];
            out_index[c] = n_out;
            offsets[n_out] = total_pts;
            info[3 * n_out + 0] = B[c].is_hole;
            info[3 * n_out + 1] = out_index[B[c].parent];
            info[3 * n_out + 2] = B[c].n_chain;
            int n = B[c].pt_end - B[c].pt_begin;
            memcpy(tmp + 2 * total_pts, pts + 2 * B[c].pt_begin, sizeof(int32_t) * 2 * (size_t)n);
            total_pts += n;
            ++n_out;
            int cnt = 0;
            int base = sp;
            for (int ch = B[c].first_child; ch; ch = B[ch].next_sibling) { stack[sp++] = ch; ++cnt; }
            /* reverse the just-pushed block so first_child is on top */
            for (int a = base, z = sp - 1; a < z; ++a, --z) { int t = stack[a]; stack[a] = stack[z]; stack[z] = t; }
        }
        offsets[n_out] = total_pts;
        memcpy(pts, tmp, sizeof(int32_t) * 2 * (size_t)total_pts);
        *n_pts_out = total_pts;
        free(tmp); free(stack); free(out_index); free(order);
        ret = n_out;
    }
    free(B);
    free(f);
    return ret;
}

/* 8-connected labelling by flood fill in raster order: label k (1-based) = k-th component by raster-first pixel */
int oracle_label8(const uint8_t *bin, int H, int W, int32_t *lab)
{
    int n = 0;
    int32_t *stack = (int32_t *)malloc(sizeof(int32_t) * (size_t)H * W + 4);
    if (!stack) return -2;
    memset(lab, 0, sizeof(int32_t) * (size_t)H * W);
    for (int i = 0; i < H * W; ++i) {
        if (!bin[i] || lab[i]) continue;
        ++n;
        int sp = 0;
        stack[sp++] = i; lab[i] = n;
        while (sp) {
            int p = stack[--sp];
            int py = p / W, px = p % W;
            for (int d = 0; d < 8; ++d) {
                int qx = px + DX[d], qy = py + DY[d];
                if (qx < 0 || qy < 0 || qx >= W || qy >= H) continue;
                int q = qy * W + qx;
                if (bin[q] && !lab[q]) { lab[q] = n; stack[sp++] = q; }
            }
        }
    }
    free(stack);
    return n;
}
