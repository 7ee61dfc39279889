// Integer triangle coverage rule of the reference's cv2 backend, in closed form.
//
// The reference draws every triangle with cv2.fillConvexPoly(img_f32, pts_int32, shift=0,
// lineType=LINE_AA) (torchdrivesim/rendering/cv2.py:59).  On a float32 image that is:
//   outline : three 8-connected LineIterator segments (clipLine + Bresenham, left-to-right)
//   fill    : 16.16 fixed-point scan conversion, rows [y_top, y_bottom-1] only
// This header enumerates exactly that pixel set for ONE triangle without walking rows from
// the top vertex: each row's span is evaluated directly from the two active edges, so a GPU
// thread (or any subset of rows) can be processed independently.
//
// It is host/device code so the very same functions are unit-tested on the CPU against the
// live cv2 module (tests/test_raster_rule.py) and used by the sm_100a kernel (raster.cu).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define TDS_HD __host__ __device__ __forceinline__
#define TDS_HD_NOINLINE static __host__ __device__ __noinline__
#else
#define TDS_HD static inline
#define TDS_HD_NOINLINE static
#endif

namespace tds {

// OpenCV clipLine on the rectangle [0,W-1]x[0,H-1]; returns false when nothing is left.
// Integer division truncating toward zero reproduces OpenCV's (int64)(double * int / int)
// exactly for |operands| < 2^26 (the quotient of two such integers is never within one
// double ulp of an integer it does not equal).
TDS_HD_NOINLINE bool clip_line(int W, int H, long long& x1, long long& y1, long long& x2, long long& y2) {
    const long long right = W - 1, bottom = H - 1;
    int c1 = (x1 < 0) + (x1 > right) * 2 + (y1 < 0) * 4 + (y1 > bottom) * 8;
    int c2 = (x2 < 0) + (x2 > right) * 2 + (y2 < 0) * 4 + (y2 > bottom) * 8;
    if ((c1 & c2) == 0 && (c1 | c2) != 0) {
        long long a;
        if (c1 & 12) {
            a = c1 < 8 ? 0 : bottom;
            x1 += ((a - y1) * (x2 - x1)) / (y2 - y1);
            y1 = a;
            c1 = (x1 < 0) + (x1 > right) * 2;
        }
        if (c2 & 12) {
            a = c2 < 8 ? 0 : bottom;
            x2 += ((a - y2) * (x2 - x1)) / (y2 - y1);
            y2 = a;
            c2 = (x2 < 0) + (x2 > right) * 2;
        }
        if ((c1 & c2) == 0 && (c1 | c2) != 0) {
            if (c1) {
                a = c1 == 1 ? 0 : right;
                y1 += ((a - x1) * (y2 - y1)) / (x2 - x1);
                x1 = a;
                c1 = 0;
            }
            if (c2) {
                a = c2 == 1 ? 0 : right;
                y2 += ((a - x2) * (y2 - y1)) / (x2 - x1);
                x2 = a;
                c2 = 0;
            }
        }
    }
    return (c1 | c2) == 0;
}

// Truncating division for the clipLine updates below, where the quotient is small: the clipped coordinate `a` lies
// between the two end points, so |a - y1| <= |y2 - y1| and |quotient| <= |x2 - x1| < 2^14.  On the GPU the quotient is
// estimated with the fp32 reciprocal (error far below 1 for such quotients) and fixed with the integer remainder:
// exact, a dozen instructions instead of the ~30 of a 32-bit division (the kernel inlines four of them).
TDS_HD int idiv_small(int num, int den) {
#if defined(__CUDA_ARCH__)
    const int un = num < 0 ? -num : num, ud = den < 0 ? -den : den;
    int q = __float2int_rz(__fdividef((float)un, (float)ud));
    const int r = un - q * ud;
    q += r < 0 ? -1 : (r >= ud ? 1 : 0);
    return (num ^ den) < 0 ? -q : q;
#else
    return num / den;
#endif
}

// Same rule in 32-bit integers, valid while |coordinates| < 8192 (products < 2^28): used by the fast path.
TDS_HD bool clip_line32(int W, int H, int& x1, int& y1, int& x2, int& y2) {
    const int right = W - 1, bottom = H - 1;
    int c1 = (x1 < 0) + (x1 > right) * 2 + (y1 < 0) * 4 + (y1 > bottom) * 8;
    int c2 = (x2 < 0) + (x2 > right) * 2 + (y2 < 0) * 4 + (y2 > bottom) * 8;
    if ((c1 & c2) == 0 && (c1 | c2) != 0) {
        int a;
        if (c1 & 12) {
            a = c1 < 8 ? 0 : bottom;
            x1 += idiv_small((a - y1) * (x2 - x1), y2 - y1);
            y1 = a;
            c1 = (x1 < 0) + (x1 > right) * 2;
        }
        if (c2 & 12) {
            a = c2 < 8 ? 0 : bottom;
            x2 += idiv_small((a - y2) * (x2 - x1), y2 - y1);
            y2 = a;
            c2 = (x2 < 0) + (x2 > right) * 2;
        }
        if ((c1 & c2) == 0 && (c1 | c2) != 0) {
            if (c1) {
                a = c1 == 1 ? 0 : right;
                y1 += idiv_small((a - x1) * (y2 - y1), x2 - x1);
                x1 = a;
                c1 = 0;
            }
            if (c2) {
                a = c2 == 1 ? 0 : right;
                y2 += idiv_small((a - x2) * (y2 - y1), x2 - x1);
                x2 = a;
                c2 = 0;
            }
        }
    }
    return (c1 | c2) == 0;
}

// 8-connected line from (xa,ya) to (xb,yb), both endpoints inclusive.  plot(x, y).
template <class Plot>
TDS_HD void draw_line8(int W, int H, int xa, int ya, int xb, int yb, Plot&& plot) {
    if ((unsigned)xa >= (unsigned)W || (unsigned)xb >= (unsigned)W ||
        (unsigned)ya >= (unsigned)H || (unsigned)yb >= (unsigned)H) {
        long long x1 = xa, y1 = ya, x2 = xb, y2 = yb;
        if (!clip_line(W, H, x1, y1, x2, y2)) return;
        xa = (int)x1; ya = (int)y1; xb = (int)x2; yb = (int)y2;
    }
    int dx = xb - xa, dy = yb - ya;
    int x = xa, y = ya;
    if (dx < 0) { dx = -dx; dy = -dy; x = xb; y = yb; }
    int sy = 1;
    if (dy < 0) { dy = -dy; sy = -1; }
    const bool vert = dy > dx;
    if (vert) { int t = dx; dx = dy; dy = t; }
    int err = dx - (dy + dy);
    const int plus = dx + dx, minus = -(dy + dy);
    for (int i = 0; i <= dx; i++) {
        plot(x, y);
        const bool m = err < 0;
        err += minus + (m ? plus : 0);
        if (vert) { y += sy; x += m ? 1 : 0; }
        else      { x += 1;  y += m ? sy : 0; }
    }
}

// slope of edge a->b (a.y < b.y) in 16.16, C-style truncating division as in FillConvexPoly.
// I = int is exact while |coordinates| < 8192 (|num| < 2^31); I = long long otherwise.
template <class I>
TDS_HD I edge_dx(int ax, int ay, int bx, int by) {
    const I dy = (I)by - ay;
    const I num = ((((I)bx - ax) << 16) * 2) + dy;
    return num / (2 * dy);
}

// the same slope for |coordinates| < 8192 (clipped-face path of the kernel).  On the GPU the quotient comes from one
// double-precision division: |num| < 2^31 and a non-integer quotient is at least 1 / |2 dy| > 2^-15 away from the next
// integer, far more than the rounding error of the double quotient (2^-22 at 2^30), so the truncation is exact.
TDS_HD int edge_dx32(int ax, int ay, int bx, int by) {
    const int dy = by - ay;
    const int num = ((bx - ax) << 16) * 2 + dy;
#if defined(__CUDA_ARCH__)
    return (int)((double)num / (double)(2 * dy));
#else
    return num / (2 * dy);
#endif
}

template <class I, class Span>
TDS_HD void fill_rows(int W, int H, int tx, int ty, int mx, int my, int bx, int by, Span&& span) {
    const int ylo = ty > 0 ? ty : 0;
    const int yhi = (by - 1) < (H - 1) ? (by - 1) : (H - 1);
    if (ylo > yhi) return;
    const I dTB = edge_dx<I>(tx, ty, bx, by);
    const I dTM = my > ty ? edge_dx<I>(tx, ty, mx, my) : 0;
    const I dMB = by > my ? edge_dx<I>(mx, my, bx, by) : 0;
    for (int y = ylo; y <= yhi; y++) {
        const I xa = ((I)tx << 16) + (I)(y - ty) * dTB;
        const I xb = y < my ? ((I)tx << 16) + (I)(y - ty) * dTM : ((I)mx << 16) + (I)(y - my) * dMB;
        const I xl = xa < xb ? xa : xb, xr = xa < xb ? xb : xa;
        int c1 = (int)((xl + 32768) >> 16), c2 = (int)((xr + 32768) >> 16);
        if (c2 >= 0 && c1 < W) {
            if (c1 < 0) c1 = 0;
            if (c2 >= W) c2 = W - 1;
            span(y, c1, c2);
        }
    }
}

// span(y, x_first, x_last) is called once per filled row with clamped inclusive columns.
// plot(x, y) is called for every outline pixel.
template <class Plot, class Span>
TDS_HD void draw_triangle(int W, int H, int x0, int y0, int x1, int y1, int x2, int y2,
                          Plot&& plot, Span&& span) {
    // outline: v2->v0, v0->v1, v1->v2
    draw_line8(W, H, x2, y2, x0, y0, plot);
    draw_line8(W, H, x0, y0, x1, y1, plot);
    draw_line8(W, H, x1, y1, x2, y2, plot);
    // sort by y: (tx,ty) top, (mx,my) middle, (bx,by) bottom
    int tx = x0, ty = y0, mx = x1, my = y1, bx = x2, by = y2, t;
    if (my < ty) { t = tx; tx = mx; mx = t; t = ty; ty = my; my = t; }
    if (by < my) { t = mx; mx = bx; bx = t; t = my; my = by; by = t; }
    if (my < ty) { t = tx; tx = mx; mx = t; t = ty; ty = my; my = t; }
    int xmin = x0 < x1 ? x0 : x1; xmin = xmin < x2 ? xmin : x2;
    int xmax = x0 > x1 ? x0 : x1; xmax = xmax > x2 ? xmax : x2;
    if (xmax < 0 || by < 0 || xmin >= W || ty >= H) return;   // bbox misses the image: outline only
    if (by == ty) return;                                     // no fill rows
    const bool small = xmin > -8192 && xmax < 8192 && ty > -8192 && by < 8192;
    if (small) fill_rows<int>(W, H, tx, ty, mx, my, bx, by, span);
    else fill_rows<long long>(W, H, tx, ty, mx, my, bx, by, span);
}


// ------------------------------------------------------------------------------------------------
// Fast path used by the raster kernel: the same pixel set as draw_triangle for |coordinates| < 8192,
// written for a linear target  index = x * sx + y * sy  (plot(index), span(index_of_first, count, sx)).
//   * bounding box within 2x2 pixels: every edge joins two 8-adjacent pixels, so the outline is the set
//     of vertices that lie inside the image, and there are no further fill pixels;
//   * y_bottom - y_top <= 1: the single fill row is covered by the outline (the span joins two vertices
//     of that row, or is one vertex), so only the outline is drawn;
//   * otherwise outline + closed-form spans, all in 32-bit integers.
template <class Plot>
TDS_HD void draw_line8_fast(int W, int H, int sx, int sy, int xa, int ya, int xb, int yb, Plot&& plot) {
    if ((unsigned)xa >= (unsigned)W || (unsigned)xb >= (unsigned)W ||
        (unsigned)ya >= (unsigned)H || (unsigned)yb >= (unsigned)H) {
        if (!clip_line32(W, H, xa, ya, xb, yb)) return;
    }
    int dx = xb - xa, dy = yb - ya;
    int idx = xa * sx + ya * sy;
    if (dx < 0) { dx = -dx; dy = -dy; idx = xb * sx + yb * sy; }
    int stepy = sy;
    if (dy < 0) { dy = -dy; stepy = -sy; }
    int major = sx, minor = stepy;
    if (dy > dx) { const int t = dx; dx = dy; dy = t; major = stepy; minor = sx; }
    int err = dx - (dy + dy);
    const int plus = dx + dx, minus = -(dy + dy);
    for (int i = dx; i >= 0; i--) {
        plot(idx);
        const bool m = err < 0;
        err += minus + (m ? plus : 0);
        idx += major + (m ? minor : 0);
    }
}

template <class Plot, class Span>
TDS_HD void draw_triangle_fast(int W, int H, int sx, int sy, int x0, int y0, int x1, int y1, int x2, int y2,
                               Plot&& plot, Span&& span) {
    int xmin = x0 < x1 ? x0 : x1; xmin = xmin < x2 ? xmin : x2;
    int xmax = x0 > x1 ? x0 : x1; xmax = xmax > x2 ? xmax : x2;
    int ymin = y0 < y1 ? y0 : y1; ymin = ymin < y2 ? ymin : y2;
    int ymax = y0 > y1 ? y0 : y1; ymax = ymax > y2 ? ymax : y2;
    if (xmax - xmin <= 1 && ymax - ymin <= 1) {
        if ((unsigned)x0 < (unsigned)W && (unsigned)y0 < (unsigned)H) plot(x0 * sx + y0 * sy);
        if ((unsigned)x1 < (unsigned)W && (unsigned)y1 < (unsigned)H) plot(x1 * sx + y1 * sy);
        if ((unsigned)x2 < (unsigned)W && (unsigned)y2 < (unsigned)H) plot(x2 * sx + y2 * sy);
        return;
    }
    {   // outline v2->v0, v0->v1, v1->v2: one rolled loop keeps the code small (instruction cache)
        int ax = x2, ay = y2, bx = x0, by = y0, cx = x1, cy = y1;
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
        for (int e = 0; e < 3; e++) {
            draw_line8_fast(W, H, sx, sy, ax, ay, bx, by, plot);
            const int tx_ = ax, ty_ = ay;
            ax = bx; ay = by; bx = cx; by = cy; cx = tx_; cy = ty_;
        }
    }
    if (ymax - ymin <= 1) return;
    if (xmax < 0 || ymax < 0 || xmin >= W || ymin >= H) return;
    int tx = x0, ty = y0, mx = x1, my = y1, bx = x2, by = y2, t;
    if (my < ty) { t = tx; tx = mx; mx = t; t = ty; ty = my; my = t; }
    if (by < my) { t = mx; mx = bx; bx = t; t = my; my = by; by = t; }
    if (my < ty) { t = tx; tx = mx; mx = t; t = ty; ty = my; my = t; }
    const int ylo = ty > 0 ? ty : 0;
    const int yhi = (by - 1) < (H - 1) ? (by - 1) : (H - 1);
    if (ylo > yhi) return;
    const int dTB = edge_dx<int>(tx, ty, bx, by);
    const int dTM = my > ty ? edge_dx<int>(tx, ty, mx, my) : 0;
    const int dMB = by > my ? edge_dx<int>(mx, my, bx, by) : 0;
    int xa = (tx << 16) + (ylo - ty) * dTB;
    for (int y = ylo; y <= yhi; y++) {
        const int xb = y < my ? (tx << 16) + (y - ty) * dTM : (mx << 16) + (y - my) * dMB;
        const int xl = xa < xb ? xa : xb, xr = xa < xb ? xb : xa;
        int c1 = (xl + 32768) >> 16, c2 = (xr + 32768) >> 16;
        if (c2 >= 0 && c1 < W) {
            if (c1 < 0) c1 = 0;
            if (c2 >= W) c2 = W - 1;
            span(c1 * sx + y * sy, c2 - c1 + 1, sx);
        }
        xa += dTB;
    }
}


// ------------------------------------------------------------------------------------------------
// Thin triangles: all three vertices INSIDE the image and spanning at most two adjacent rows.  Then no edge
// is clipped, the fill adds nothing to the outline (see draw_triangle_fast; this does NOT hold for two
// adjacent columns, where the fixed-point spans can leave the Bresenham outline), and every edge is at most
// two axis-aligned runs, because an 8-connected LineIterator segment that moves by one along its minor axis
// takes that step after  i0 = (d + 2) >> 1  pixels (d = extent along the major axis,
// k_i = floor((2 i + d - 1) / (2 d)) flips at i0), always starting from the LEFT endpoint.
// run(index_of_first, count, step) writes `count` pixels starting at a linear index with the given stride.
TDS_HD bool is_thin_inside(int W, int H, int x0, int y0, int x1, int y1, int x2, int y2) {
    const bool inside = ((unsigned)x0 < (unsigned)W) & ((unsigned)x1 < (unsigned)W) & ((unsigned)x2 < (unsigned)W) &
                        ((unsigned)y0 < (unsigned)H) & ((unsigned)y1 < (unsigned)H) & ((unsigned)y2 < (unsigned)H);
    int ymin = y0 < y1 ? y0 : y1; ymin = ymin < y2 ? ymin : y2;
    int ymax = y0 > y1 ? y0 : y1; ymax = ymax > y2 ? ymax : y2;
    return inside && (ymax - ymin <= 1);
}

template <class Run>
TDS_HD void draw_triangle_thin(int sx, int sy, int x0, int y0, int x1, int y1, int x2, int y2, Run&& run) {
    // edges v2->v0, v0->v1, v1->v2 in one rolled loop (code size), each as two runs of a common stride
    int xa = x2, ya = y2, xb = x0, yb = y0, xc = x1, yc = y1;
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
    for (int e = 0; e < 3; e++) {
        // endpoints left to right (LineIterator leftToRight; a vertical edge keeps its order, which does not
        // change its pixel set)
        const bool swap = xb < xa;
        const int lx = swap ? xb : xa, ly = swap ? yb : ya, rx = swap ? xa : xb, ry = swap ? ya : yb;
        const int dx = rx - lx;
        int dy = ry - ly, stepy = sy;
        if (dy < 0) { dy = -dy; stepy = -sy; }
        const int base = lx * sx + ly * sy;
        // x-major (dy <= 1 <= dx or a horizontal edge): i0 pixels on the left endpoint's row, the rest on the
        // other row; y-major (dx <= 1): i0 pixels in the left endpoint's column, the rest in the next column
        const bool xmajor = dx >= dy;
        const int major = xmajor ? dx : dy, minor = xmajor ? dy : dx;
        const int stride = xmajor ? sx : stepy, jump = xmajor ? stepy : sx;
        const int n = major + 1;
        const int i0 = minor == 0 ? n : ((major + 2) >> 1);
        run(base, i0, stride);
        run(base + i0 * stride + jump, n - i0, stride);
        const int tx_ = xa, ty_ = ya;
        xa = xb; ya = yb; xb = xc; yb = yc; xc = tx_; yc = ty_;
    }
}

}  // namespace tds
