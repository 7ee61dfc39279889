// Row-run form of the reference's triangle coverage rule (cv2.fillConvexPoly on int32 vertices,
// torchdrivesim/rendering/cv2.py:59): for ONE image row y the covered pixels of a triangle are the union
// of at most four intervals of columns —
//   * the 16.16 fixed-point fill span of that row (rows [y_top, y_bottom - 1]);
//   * for each of the three outline edges, the run of its 8-connected LineIterator segment in that row.
// The raster kernel keeps one bit per pixel and class, so a row of a triangle is a handful of integer
// operations and one atomic OR instead of a loop over pixels; painter's order is resolved once at the end.
//
// Closed form of the LineIterator runs (left-to-right Bresenham with err0 = dx - 2 dy, see draw_line8 in
// tds_raster_tri.h).  After clipping and ordering the endpoints left to right, with dx = rx - lx >= 0,
// dy = |ry - ly| and r = |y - ly| the row offset from the LEFT endpoint:
//   x-major (dy <= dx): pixel i (x = lx + i) lies in row offset k_i = floor((2 dy i + dx - 1) / (2 dx)), hence
//       row r holds i in [G(r), G(r+1) - 1],  G(0) = 0,  G(r) = floor((2 dx r - dx + 2 dy) / (2 dy)) for
//       1 <= r <= dy,  G(dy + 1) = dx + 1;
//   y-major (dy >  dx): row r holds the single pixel x = lx + floor((2 dx r + dy - 1) / (2 dy)).
// Both divide by 2 dy: one multiply-high with rcp = ceil(2^32 / (2 dy)) is exact while numerator * 2 dy < 2^32,
// i.e. for images up to 1024 pixels (clipped endpoints lie inside the image).
//
// Host/device code: tests/test_raster_rule.py checks exactly these functions against the live cv2 module.
#pragma once
#include "tds_raster_tri.h"

namespace tds {

TDS_HD uint32_t mulhi_u32(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
    return __umulhi(a, b);
#else
    return (uint32_t)(((unsigned long long)a * b) >> 32);
#endif
}

// ceil(2^32 / (2 dy)) for dy >= 1 (2 dy >= 2, so the result fits 32 bits)
TDS_HD uint32_t row_rcp(int dy) { return dy > 0 ? 0xffffffffu / (uint32_t)(2 * dy) + 1u : 0u; }

// One outline edge in row-evaluable form.  Walked by increasing y: t = y - ylo in [0, dy].
struct RowEdge {
    int lx;          // column of the left endpoint
    int ylo, yhi;    // rows covered (yhi < ylo: the edge was rejected by clipLine)
    int n;           // numerator of the row that is evaluated next (advances by b1 per row)
    int b1;          // +-2 dx (sign = direction of r as y grows)
    int cur;         // x-major: boundary G(.) shared with the previous row
    int gend;        // x-major: boundary after the last row (dx + 1 walking away from the left endpoint, else 0)
    uint32_t rcp;
    int xmajor;
};

// (xa,ya) -> (xb,yb) in the reference's drawing order (clipLine is not symmetric in its endpoints)
template <class RcpFn>
TDS_HD void row_edge_setup(int W, int H, int xa, int ya, int xb, int yb, RowEdge& e, RcpFn&& rcp_of) {
    if ((unsigned)xa >= (unsigned)W || (unsigned)xb >= (unsigned)W ||
        (unsigned)ya >= (unsigned)H || (unsigned)yb >= (unsigned)H) {
        if (!clip_line32(W, H, xa, ya, xb, yb)) { e.ylo = 1; e.yhi = 0; e.lx = 0; e.n = 0; e.b1 = 0; e.cur = 0; e.gend = 0; e.rcp = 0; e.xmajor = 1; return; }
    }
    int dx = xb - xa, dy = yb - ya, lx = xa, ly = ya;
    if (dx < 0) { dx = -dx; dy = -dy; lx = xb; ly = yb; }
    const bool down = dy >= 0;           // r grows with y
    if (dy < 0) dy = -dy;
    e.lx = lx;
    e.ylo = down ? ly : ly - dy;
    e.yhi = down ? ly + dy : ly;
    e.xmajor = dx >= dy;
    e.rcp = rcp_of(dy);
    const int a1 = 2 * dx;
    if (e.xmajor) {
        // boundary between row offsets r-1 and r: G(r) = mulhi(a1 r + (2 dy - dx), rcp), 1 <= r <= dy.
        // Walking y upwards: down -> rows r = 0,1,..: row t needs G(t) (cur) and G(t+1) (next, from n);
        //                    up   -> rows r = dy,dy-1,..: row t needs G(r+1) (cur) and G(r) (next, from n).
        const int a0 = 2 * dy - dx;
        e.cur = down ? 0 : dx + 1;
        e.gend = down ? dx + 1 : 0;
        e.n = down ? a1 + a0 : a1 * dy + a0;
        e.b1 = down ? a1 : -a1;
    } else {
        const int a0 = dy - 1;
        e.cur = 0; e.gend = 0;
        e.n = down ? a0 : a1 * dy + a0;
        e.b1 = down ? a1 : -a1;
    }
}

// Run [lo, hi] of the edge in row y (ylo <= y <= yhi; rows must be visited in increasing order, each once).
TDS_HD void row_edge_step(RowEdge& e, int y, int& lo, int& hi) {
    const int q = (int)mulhi_u32((uint32_t)e.n, e.rcp);
    e.n += e.b1;
    if (e.xmajor) {
        const int nxt = y == e.yhi ? e.gend : q;
        const int a = e.cur < nxt ? e.cur : nxt, b = e.cur < nxt ? nxt : e.cur;
        lo = e.lx + a;
        hi = e.lx + b - 1;
        e.cur = nxt;
    } else {
        lo = hi = e.lx + q;
    }
}

// Fixed-point fill (FillConvexPoly): rows [fylo, fyhi], vertices sorted by y exactly as draw_triangle does.
struct RowFill {
    int fylo, fyhi;       // fyhi < fylo: nothing to fill
    int my;
    int xa, dTB;          // long edge T->B, evaluated incrementally
    int xT, dTM, xM, dMB; // short edges: x = xT + (y - ty) dTM above my, xM + (y - my) dMB from my on
    int ty;
};

TDS_HD void row_fill_setup(int W, int H, int x0, int y0, int x1, int y1, int x2, int y2, RowFill& f) {
    int tx = x0, ty = y0, mx = x1, my = y1, bx = x2, by = y2, t;
    if (my < ty) { t = tx; tx = mx; mx = t; t = ty; ty = my; my = t; }
    if (by < my) { t = mx; mx = bx; bx = t; t = my; my = by; by = t; }
    if (my < ty) { t = tx; tx = mx; mx = t; t = ty; ty = my; my = t; }
    int xmin = x0 < x1 ? x0 : x1; xmin = xmin < x2 ? xmin : x2;
    int xmax = x0 > x1 ? x0 : x1; xmax = xmax > x2 ? xmax : x2;
    f.fylo = 1; f.fyhi = 0;
    f.my = my; f.ty = ty; f.xa = 0; f.dTB = 0; f.xT = 0; f.dTM = 0; f.xM = 0; f.dMB = 0;
    if (xmax < 0 || by < 0 || xmin >= W || ty >= H) return;   // bounding box misses the image: outline only
    if (by == ty) return;                                     // no fill rows
    const int ylo = ty > 0 ? ty : 0;
    const int yhi = (by - 1) < (H - 1) ? (by - 1) : (H - 1);
    if (ylo > yhi) return;
    f.fylo = ylo; f.fyhi = yhi;
    f.dTB = edge_dx32(tx, ty, bx, by);
    f.dTM = my > ty ? edge_dx32(tx, ty, mx, my) : 0;
    f.dMB = by > my ? edge_dx32(mx, my, bx, by) : 0;
    f.xT = tx << 16;
    f.xM = mx << 16;
    f.xa = (tx << 16) + (ylo - ty) * f.dTB;
}

// Span of row y (fylo <= y <= fyhi, increasing order); returns false when the span misses the image.
TDS_HD bool row_fill_step(RowFill& f, int W, int y, int& lo, int& hi) {
    const int xb = y < f.my ? f.xT + (y - f.ty) * f.dTM : f.xM + (y - f.my) * f.dMB;
    const int xa = f.xa;
    f.xa += f.dTB;
    const int xl = xa < xb ? xa : xb, xr = xa < xb ? xb : xa;
    int c1 = (xl + 32768) >> 16, c2 = (xr + 32768) >> 16;
    if (c2 < 0 || c1 >= W) return false;
    lo = c1 < 0 ? 0 : c1;
    hi = c2 >= W ? W - 1 : c2;
    return true;
}

struct RowTri {
    RowEdge e[3];
    RowFill f;
    int ylo, yhi;     // rows of the image that any part of the triangle can touch
};

// Valid for |coordinates| < 8192 (32-bit clipLine and slopes), like draw_triangle_fast.
template <class RcpFn>
TDS_HD void row_tri_setup(int W, int H, int x0, int y0, int x1, int y1, int x2, int y2, RowTri& t, RcpFn&& rcp_of) {
    // outline v2->v0, v0->v1, v1->v2
    row_edge_setup(W, H, x2, y2, x0, y0, t.e[0], rcp_of);
    row_edge_setup(W, H, x0, y0, x1, y1, t.e[1], rcp_of);
    row_edge_setup(W, H, x1, y1, x2, y2, t.e[2], rcp_of);
    row_fill_setup(W, H, x0, y0, x1, y1, x2, y2, t.f);
    int ymin = y0 < y1 ? y0 : y1; ymin = ymin < y2 ? ymin : y2;
    int ymax = y0 > y1 ? y0 : y1; ymax = ymax > y2 ? ymax : y2;
    t.ylo = ymin > 0 ? ymin : 0;
    t.yhi = ymax < H - 1 ? ymax : H - 1;
}

// emit(lo, hi) for every interval of row y; rows must be visited from t.ylo to t.yhi in order.
template <class Emit>
TDS_HD void row_tri_step(RowTri& t, int W, int y, Emit&& emit) {
    int lo, hi;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int k = 0; k < 3; k++)
        if (y >= t.e[k].ylo && y <= t.e[k].yhi) { row_edge_step(t.e[k], y, lo, hi); emit(lo, hi); }
    if (y >= t.f.fylo && y <= t.f.fyhi && row_fill_step(t.f, W, y, lo, hi)) emit(lo, hi);
}

// The same pixel set spread over FOUR cooperating workers (the raster kernel: four lanes per face that crosses the
// image border; a third of the code of row_tri_setup + row_tri_step).  Worker `part` = 0, 1, 2
// A triangle that crosses the image border, drawn by FOUR cooperating workers (part = 0..3; four lanes on the GPU, one
// caller playing them in turn on the host).  Workers 0..2 walk the clipped runs of the outline edges v2->v0, v0->v1,
// v1->v2 and compute the 16.16 fill slope of their edge; share(value, w) -> the value held by worker w (a shuffle on the
// GPU).  Then every worker takes each fourth fill row: the span of a row joins the fill columns of the edges that are
// active in it (edge top -> bottom is active in rows [y_top, y_bottom - 1]: exactly two of them in every fill row),
// evaluated from the vertices and the three slopes alone.  emit(y, lo, hi) per interval: the intervals of a row arrive
// from different workers; a bit-per-pixel target ORs them, so the order does not matter.
template <class RcpFn, class Emit, class Share>
TDS_HD void row_tri_part(int W, int H, int x0, int y0, int x1, int y1, int x2, int y2, int part, RcpFn&& rcp_of, Emit&& emit,
                         Share&& share) {
    int slope = 0;
    if (part < 3) {
        const int ax = part == 0 ? x2 : (part == 1 ? x0 : x1), ay = part == 0 ? y2 : (part == 1 ? y0 : y1);
        const int bx = part == 0 ? x0 : (part == 1 ? x1 : x2), by = part == 0 ? y0 : (part == 1 ? y1 : y2);
        if (ay != by) slope = ay < by ? edge_dx32(ax, ay, bx, by) : edge_dx32(bx, by, ax, ay);
        RowEdge e;
        row_edge_setup(W, H, ax, ay, bx, by, e, rcp_of);
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
        for (int y = e.ylo; y <= e.yhi; y++) {
            int lo, hi;
            row_edge_step(e, y, lo, hi);
            emit(y, lo, hi);
        }
    }
    // the fill (FillConvexPoly): nothing when the bounding box misses the image; rows [max(y_top, 0), min(y_bottom - 1, H - 1)]
    const int d20 = share(slope, 0), d01 = share(slope, 1), d12 = share(slope, 2);
    int xmin = x0 < x1 ? x0 : x1; xmin = xmin < x2 ? xmin : x2;
    int xmax = x0 > x1 ? x0 : x1; xmax = xmax > x2 ? xmax : x2;
    int ty = y0 < y1 ? y0 : y1; ty = ty < y2 ? ty : y2;
    int by = y0 > y1 ? y0 : y1; by = by > y2 ? by : y2;
    if (xmax < 0 || by < 0 || xmin >= W || ty >= H) return;
    const int ylo = ty > 0 ? ty : 0, yhi = (by - 1) < (H - 1) ? (by - 1) : (H - 1);
    // edges oriented top -> bottom: (x at the top) << 16, top row, bottom row
    const int t20 = y2 < y0 ? y2 : y0, b20 = y2 < y0 ? y0 : y2, s20 = (y2 < y0 ? x2 : x0) << 16;
    const int t01 = y0 < y1 ? y0 : y1, b01 = y0 < y1 ? y1 : y0, s01 = (y0 < y1 ? x0 : x1) << 16;
    const int t12 = y1 < y2 ? y1 : y2, b12 = y1 < y2 ? y2 : y1, s12 = (y1 < y2 ? x1 : x2) << 16;
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
    for (int y = ylo + part; y <= yhi; y += 4) {
        int xl = 0x7fffffff, xr = -0x7fffffff - 1;
        if (t20 <= y && y < b20) { const int x = s20 + (y - t20) * d20; xl = xl < x ? xl : x; xr = xr > x ? xr : x; }
        if (t01 <= y && y < b01) { const int x = s01 + (y - t01) * d01; xl = xl < x ? xl : x; xr = xr > x ? xr : x; }
        if (t12 <= y && y < b12) { const int x = s12 + (y - t12) * d12; xl = xl < x ? xl : x; xr = xr > x ? xr : x; }
        const int c1 = (xl + 32768) >> 16, c2 = (xr + 32768) >> 16;
        if (c2 >= 0 && c1 < W) emit(y, c1 < 0 ? 0 : c1, c2 >= W ? W - 1 : c2);
    }
}


// ------------------------------------------------------------------------------------------------
// Fast path: all three vertices INSIDE the image (no clipLine, no clamps).  Then every row of the triangle
// is ONE interval [L, R]: the fill span ends within one pixel of the LineIterator pixels of the two boundary
// edges (both approximate the same line, the span by round-half-up of a 16.16 value, the iterator by
// round-half-down), runs of x-major edges are contiguous with them, and at the rows of the vertices the
// edges meet in the vertex pixel.  So L = min and R = max over the fill span and the runs of the edges
// that are active in the row: the long edge T-B, and T-M above the middle vertex / M-B from it on.
// (Clipped edges do NOT have this property: they are redrawn between moved endpoints.)
struct FastEdge {
    int cur;         // x-major: boundary shared with the previous row, as an absolute column
    int gmax;        // x-major: lx + dx + 1, the boundary after the last pixel
    int n, b1;
    int lx;
    uint32_t rcp;
    int xm;          // ~0 for x-major, 0 for y-major
};

TDS_HD int imax_(int a, int b) { return a > b ? a : b; }
TDS_HD int imin_(int a, int b) { return a < b ? a : b; }

// top endpoint (px,py) -> bottom endpoint (qx,qy), py <= qy.
// The boundary of an x-major edge towards the next row is mulhi(max(n, 0), rcp) clamped to dx + 1: the clamp
// yields G(dy + 1) = dx + 1 after the last row of an edge walked away from its left endpoint (the formula
// gives dx + 1 + floor(dx / 2dy) there), the max yields G(0) = 0 after the last row of an edge walked towards
// its left endpoint (n = 2 dy - dx may be negative there).  A horizontal edge is walked "towards" its left end.
template <class RcpFn>
TDS_HD void fast_edge_setup(int px, int py, int qx, int qy, FastEdge& e, RcpFn&& rcp_of) {
    const int dy = qy - py, dxs = qx - px;
    const bool down = dxs >= 0 && dy > 0;       // the left endpoint is the top one
    const int dx = dxs >= 0 ? dxs : -dxs;
    const int lx = dxs >= 0 ? px : qx;
    const bool xmajor = dx >= dy;
    const int a1 = 2 * dx;
    const int a0 = xmajor ? 2 * dy - dx : dy - 1;
    e.lx = lx;
    e.rcp = rcp_of(dy);
    e.xm = xmajor ? -1 : 0;
    e.gmax = lx + dx + 1;
    e.cur = down ? lx : lx + dx + 1;
    e.b1 = down ? a1 : -a1;
    e.n = down ? a0 + (xmajor ? a1 : 0) : a1 * dy + a0;
}

TDS_HD void fast_edge_step(FastEdge& e, int& lo, int& hi) {
    const int q = e.lx + (int)mulhi_u32((uint32_t)imax_(e.n, 0), e.rcp);
    e.n += e.b1;
    const int nxt = imin_(q, e.gmax);
    const int a = imin_(e.cur, nxt), b = imax_(e.cur, nxt) - 1;
    e.cur = nxt;
    lo = (a & e.xm) | (q & ~e.xm);
    hi = (b & e.xm) | (q & ~e.xm);
}

// 16.16 slope of FillConvexPoly through the reciprocal table: exact while |num| * 2 dy < 2^32 (images <= 128)
TDS_HD int edge_dx_rcp(int dxs, int dy, uint32_t rcp) {
    const int num = dxs * 131072 + dy;
    const int q = (int)mulhi_u32((uint32_t)(num < 0 ? -num : num), rcp);
    return num < 0 ? -q : q;
}

struct FastTri {
    FastEdge tb, s, mb;      // long edge, current short edge (T-M until the middle row), M-B
    int ty, my, by;
    int xa, dTB, xb, dS, xM, dMB;   // fill: both ends carry the +0.5 rounding bias
    int eL, eH;              // run of T-M in its LAST row (the middle row), precomputed: the row loop never steps
                             // T-M there, it only switches to M-B
};

// run of the edge top (px,py) -> bottom (qx,qy) in its last row qy, in closed form
template <class RcpFn>
TDS_HD void fast_edge_last_run(int px, int py, int qx, int qy, int& lo, int& hi, RcpFn&& rcp_of) {
    const int dy = qy - py, dxs = qx - px;
    const int dx = dxs >= 0 ? dxs : -dxs;
    const int lx = dxs >= 0 ? px : qx;
    if (dy == 0) { lo = lx; hi = lx + dx; return; }            // horizontal: the whole edge
    if (dx < dy) { lo = hi = qx; return; }                      // y-major: one pixel per row, the endpoint
    // x-major: row offset r from the LEFT endpoint holds [G(r), G(r+1) - 1], G(r) = mulhi(2 dx r - dx + 2 dy, rcp)
    const uint32_t rcp = rcp_of(dy);
    if (dxs >= 0) {         // the bottom endpoint is the right one: last row r = dy -> [G(dy), dx]
        lo = lx + (int)mulhi_u32((uint32_t)(2 * dx * dy - dx + 2 * dy), rcp);
        hi = lx + dx;
    } else {                // the bottom endpoint is the left one: last row r = 0 -> [0, G(1) - 1]
        lo = lx;
        hi = lx + (int)mulhi_u32((uint32_t)(dx + 2 * dy), rcp) - 1;
    }
}

template <bool SMALL, class RcpFn>
TDS_HD void fast_tri_setup(int x0, int y0, int x1, int y1, int x2, int y2, FastTri& t, RcpFn&& rcp_of) {
    int tx = x0, ty = y0, mx = x1, my = y1, bx = x2, by = y2, w;
    if (my < ty) { w = tx; tx = mx; mx = w; w = ty; ty = my; my = w; }
    if (by < my) { w = mx; mx = bx; bx = w; w = my; my = by; by = w; }
    if (my < ty) { w = tx; tx = mx; mx = w; w = ty; ty = my; my = w; }
    fast_edge_setup(tx, ty, bx, by, t.tb, rcp_of);
    fast_edge_setup(tx, ty, mx, my, t.s, rcp_of);
    fast_edge_setup(mx, my, bx, by, t.mb, rcp_of);
    fast_edge_last_run(tx, ty, mx, my, t.eL, t.eH, rcp_of);
    t.ty = ty; t.my = my; t.by = by;
    if (SMALL) {
        t.dTB = by > ty ? edge_dx_rcp(bx - tx, by - ty, t.tb.rcp) : 0;
        t.dS = my > ty ? edge_dx_rcp(mx - tx, my - ty, t.s.rcp) : 0;
        t.dMB = by > my ? edge_dx_rcp(bx - mx, by - my, t.mb.rcp) : 0;
    } else {
        t.dTB = by > ty ? edge_dx<int>(tx, ty, bx, by) : 0;
        t.dS = my > ty ? edge_dx<int>(tx, ty, mx, my) : 0;
        t.dMB = by > my ? edge_dx<int>(mx, my, bx, by) : 0;
    }
    t.xa = (tx << 16) + 32768;
    t.xb = t.xa;
    t.xM = (mx << 16) + 32768;
}

// emit(y, L, R) once per row ty..by, in order
template <class Emit>
TDS_HD void fast_tri_rows(FastTri& t, Emit&& emit) {
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
    for (int y = t.ty; y <= t.by; y++) {
        const bool mid = y == t.my;
        if (mid) {
            // M-B takes over (its first row is this one); the last row of T-M was computed at set-up
            t.s = t.mb;
            t.xb = t.xM;
            t.dS = t.dMB;
        }
        const int eL = mid ? t.eL : 0x7fffffff, eH = mid ? t.eH : -1;
        int l1, h1, l2, h2;
        fast_edge_step(t.tb, l1, h1);
        fast_edge_step(t.s, l2, h2);
        const int xl = t.xa < t.xb ? t.xa : t.xb, xr = t.xa < t.xb ? t.xb : t.xa;
        t.xa += t.dTB;
        t.xb += t.dS;
        int L = xl >> 16, R = xr >> 16;
        L = L < l1 ? L : l1; L = L < l2 ? L : l2; L = L < eL ? L : eL;
        R = R > h1 ? R : h1; R = R > h2 ? R : h2; R = R > eH ? R : eH;
        emit(y, L, R);
    }
}

// ------------------------------------------------------------------------------------------------
// A triangle INSIDE the image as three line walkers: each edge is an 8-connected line with the 16.16 fill column of
// FillConvexPoly's edge walker, and a row is ONE interval - from the left-most to the right-most of the lines' pixels
// and fill columns in that row.  (The outline is the three lines; the fill span of a row joins the fill columns of the
// two edges that are active in it; all three sets lie in one interval because the outline is convex.)  No
// middle-vertex switch, no vertex sorting: the three walkers are the same code, active in their own row ranges.

struct LineWalk {
    FastEdge e;      // LineIterator runs of the line, stepped once per row
    int xf, d;       // 16.16 fill column (+0.5) of the row that is evaluated next, and its slope
    int t, len;      // rows since the line's first row (negative before it), index of its last row
};

// line (px,py)-(qx,qy) inside the image (up to 128 pixels: slopes through the reciprocal table)
template <class RcpFn>
TDS_HD void line_walk_setup(int px, int py, int qx, int qy, LineWalk& l, RcpFn&& rcp_of) {
    if (qy < py) { int w = px; px = qx; qx = w; w = py; py = qy; qy = w; }
    const int dy = qy - py;
    l.d = dy > 0 ? edge_dx_rcp(qx - px, dy, rcp_of(dy)) : 0;
    l.xf = (px << 16) + 32768;
    l.t = py;            // lines3_setup turns this into (first row of the item - py)
    l.len = dy;
    fast_edge_setup(px, py, qx, qy, l.e, rcp_of);
    if (dy == 0) {
        // a horizontal line is a single run; fast_edge_setup hands it out at the FIRST step, but this walker may be
        // stepped before its row: keep the run for the step whose numerator reaches 2^20 (mulhi(2^20, 2^32 - 1) > dx)
        l.e.rcp = 0xffffffffu;
        l.e.n = 1 << 20;
        l.e.b1 = 1 << 20;
        l.e.cur = l.e.lx;
    }
}

struct Lines3 {
    LineWalk l[3];
    int y0, y1;
};

template <class RcpFn>
TDS_HD void lines3_setup(int x0, int y0_, int x1, int y1_, int x2, int y2_, Lines3& q, RcpFn&& rcp_of) {
    line_walk_setup(x0, y0_, x1, y1_, q.l[0], rcp_of);
    line_walk_setup(x1, y1_, x2, y2_, q.l[1], rcp_of);
    line_walk_setup(x2, y2_, x0, y0_, q.l[2], rcp_of);
    int y0 = q.l[0].t, y1 = q.l[0].t + q.l[0].len;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int k = 1; k < 3; k++) {
        const int ys = q.l[k].t, ye = ys + q.l[k].len;
        y0 = y0 < ys ? y0 : ys; y1 = y1 > ye ? y1 : ye;
    }
    q.y0 = y0; q.y1 = y1;
    // all three walkers start at row y0: a LineIterator walker stepped before its first row stays where it is (the
    // clamps of fast_edge_step), the fill column is linear
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int k = 0; k < 3; k++) {
        const int lead = q.l[k].t - y0;         // rows until the line starts
        q.l[k].e.n -= lead * q.l[k].e.b1;
        q.l[k].xf -= lead * q.l[k].d;
        q.l[k].t = -lead;
    }
}

// emit(y, L, R) once per row y0..y1, in order
template <class Emit>
TDS_HD void lines3_rows(Lines3& q, Emit&& emit) {
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
    for (int y = q.y0; y <= q.y1; y++) {
        int L = 0x7fffffff, R = -1;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int k = 0; k < 3; k++) {
            LineWalk& l = q.l[k];
            int lo, hi;
            fast_edge_step(l.e, lo, hi);
            const int xf = l.xf >> 16;
            l.xf += l.d;
            const bool on = (unsigned)l.t <= (unsigned)l.len;       // 0 <= t <= len
            l.t++;
            if (on) {
                lo = lo < xf ? lo : xf; hi = hi > xf ? hi : xf;
                L = L < lo ? L : lo; R = R > hi ? R : hi;
            }
        }
        emit(y, L, R);
    }
}

}  // namespace tds
