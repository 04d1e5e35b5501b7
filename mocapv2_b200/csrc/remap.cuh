// cv.remap(INTER_LINEAR, BORDER_CONSTANT 0) with OpenCV's 5-bit fixed-point weights (lib/ImageOperations.py:38), one pixel.
#pragma once
#include "common.cuh"

__device__ __forceinline__ int remap_px(const uint8_t* __restrict__ fr, int W, int H, int i, int j, uint32_t m)
{
    if (m == MAP_OUTSIDE) return 0;
    int iu = 32 * j + (int)(int16_t)(m & 0xffff);
    int iv = 32 * i + (int)(int16_t)(m >> 16);
    int sx = iu >> 5, sy = iv >> 5, fx = iu & 31, fy = iv & 31;
    const uint8_t* p = fr + (ptrdiff_t)sy * W + sx;
    bool x0ok = (unsigned)sx < (unsigned)W, x1ok = (unsigned)(sx + 1) < (unsigned)W;
    bool y0ok = (unsigned)sy < (unsigned)H, y1ok = (unsigned)(sy + 1) < (unsigned)H;
    int p00 = (x0ok && y0ok) ? p[0] : 0;
    int p01 = (x1ok && y0ok) ? p[1] : 0;
    int p10 = (x0ok && y1ok) ? p[W] : 0;
    int p11 = (x1ok && y1ok) ? p[W + 1] : 0;
    int r0 = (32 - fx) * p00 + fx * p01;
    int r1 = (32 - fx) * p10 + fx * p11;
    return ((32 - fy) * r0 + fy * r1 + 512) >> 10;
}

