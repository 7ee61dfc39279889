// Stateless form of the row rule of tds_raster_rows.h: the intervals of row y of a triangle from the set-up alone,
// in any order of the rows, so that the (face, row) items of a batch of faces can be spread over the lanes of a
// warp instead of every lane walking the rows of its own face.  NOT used by the kernels yet: the
// identities are checked on the host against cv2 and against the incremental walk (tests/test_raster_rule.py).
#pragma once
#include "tds_raster_rows.h"

namespace tds {

// Run [lo, hi] of the edge in row y, ylo <= y <= yhi, from the edge as row_edge_setup left it (row_edge_step
// advances n by b1 per row and carries the boundary G of the previous row in cur).
TDS_HD void row_edge_at(const RowEdge& e, int y, int& lo, int& hi) {
    const int t = y - e.ylo;
    const int q = (int)mulhi_u32((uint32_t)(e.n + t * e.b1), e.rcp);
    if (e.xmajor) {
        const int cur = t == 0 ? e.cur : (int)mulhi_u32((uint32_t)(e.n + (t - 1) * e.b1), e.rcp);
        const int nxt = y == e.yhi ? e.gend : q;
        const int a = cur < nxt ? cur : nxt, b = cur < nxt ? nxt : cur;
        lo = e.lx + a;
        hi = e.lx + b - 1;
    } else {
        lo = hi = e.lx + q;
    }
}

// Span of row y, fylo <= y <= fyhi, from the fill as row_fill_setup left it; false when it misses the image.
TDS_HD bool row_fill_at(const RowFill& f, int W, int y, int& lo, int& hi) {
    const int xb = y < f.my ? f.xT + (y - f.ty) * f.dTM : f.xM + (y - f.my) * f.dMB;
    const int xa = f.xa + (y - f.fylo) * f.dTB;
    const int xl = xa < xb ? xa : xb, xr = xa < xb ? xb : xa;
    const int c1 = (xl + 32768) >> 16, c2 = (xr + 32768) >> 16;
    if (c2 < 0 || c1 >= W) return false;
    lo = c1 < 0 ? 0 : c1;
    hi = c2 >= W ? W - 1 : c2;
    return true;
}

// emit(lo, hi) for every interval of row y of the triangle (t.ylo <= y <= t.yhi), t untouched
template <class Emit>
TDS_HD void row_tri_at(const RowTri& t, int W, int y, Emit&& emit) {
    int lo, hi;
    for (int k = 0; k < 3; k++)
        if (y >= t.e[k].ylo && y <= t.e[k].yhi) { row_edge_at(t.e[k], y, lo, hi); emit(lo, hi); }
    if (y >= t.f.fylo && y <= t.f.fyhi && row_fill_at(t.f, W, y, lo, hi)) emit(lo, hi);
}

}  // namespace tds
