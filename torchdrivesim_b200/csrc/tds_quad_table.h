// Coverage patterns of sliver quads.
//
// Road surfaces and lane markings are triangle STRIPS (lanelet2.py:253-283 and the road mesh of the map files) whose
// cross-pieces are a pixel apart at the usual zoom: consecutive faces (v0, v1, v2), (v1, v2, v3) form a quad with two long
// sides ("rungs", v0-v1 and v2-v3) whose end points are the same or 8-adjacent pixels ("rails", v0-v2 and v1-v3).  For
// a pair of faces INSIDE the image the reference's rule (cv2.fillConvexPoly on int32 vertices, no clipping) is invariant
// under integer translations, so the pixels of the two triangles depend only on the rung g = v1 - v0 and the two rails
// r1 = v2 - v0, r2 = v3 - v1: a finite set of shapes.  The table holds, for every g within +-15 pixels and every pair of
// rails within +-1 pixel, the union of the two triangles as one 32-bit mask per row (bit j = column xmin + j, row 0 =
// ymin of the four vertices), generated at start-up BY the reference's rule (tds::draw_triangle, pinned to cv2 by
// tests/test_raster_rule.py).  The raster kernel ORs the rows into its bitplanes: ~6 instructions per row instead of
// two edge-walking triangles.  6.2 MB per GPU, resident in L2.
#pragma once
#include <stdint.h>
#include "tds_raster_tri.h"

namespace tds {

constexpr int kQuadR = 15;                                   // rung components within +-kQuadR
constexpr int kQuadSpan = 2 * kQuadR + 1;
constexpr int kQuadRows = 20;                                // rows stored per pattern (height <= kQuadR + 3), 4 per uint4
constexpr int kQuadPatterns = kQuadSpan * kQuadSpan * 81;

// rails within one pixel per axis, rung within the table
TDS_HD bool quad_in_table(int gx, int gy, int r1x, int r1y, int r2x, int r2y) {
    return ((unsigned)(r1x + 1) <= 2u) & ((unsigned)(r1y + 1) <= 2u) & ((unsigned)(r2x + 1) <= 2u) & ((unsigned)(r2y + 1) <= 2u) &
           ((unsigned)(gx + kQuadR) <= 2u * kQuadR) & ((unsigned)(gy + kQuadR) <= 2u * kQuadR);
}

TDS_HD int quad_pattern_index(int gx, int gy, int r1x, int r1y, int r2x, int r2y) {
    return (((gy + kQuadR) * kQuadSpan + (gx + kQuadR)) * 9 + (r1y + 1) * 3 + (r1x + 1)) * 9 + (r2y + 1) * 3 + (r2x + 1);
}


// rows[kQuadRows] of the pattern (g, r1, r2): faces (A, B, C), (B, C, D) with B = A + g, C = A + r1, D = B + r2
static inline void quad_pattern_rows(int gx, int gy, int r1x, int r1y, int r2x, int r2y, uint32_t* rows) {
    constexpr int W = 64, H = 64, ax = 24, ay = 24;          // every vertex within [24 - 17, 24 + 17]: inside, no clipping
    const int bx = ax + gx, by = ay + gy, cx = ax + r1x, cy = ay + r1y, dx = bx + r2x, dy = by + r2y;
    const int xs[4] = {ax, bx, cx, dx}, ys[4] = {ay, by, cy, dy};
    int xmin = ax, ymin = ay;
    for (int k = 1; k < 4; k++) { xmin = xs[k] < xmin ? xs[k] : xmin; ymin = ys[k] < ymin ? ys[k] : ymin; }
    for (int r = 0; r < kQuadRows; r++) rows[r] = 0u;
    auto plot = [&](int x, int y) { rows[y - ymin] |= 1u << (x - xmin); };
    auto span = [&](int y, int xa, int xb) { for (int x = xa; x <= xb; x++) rows[y - ymin] |= 1u << (x - xmin); };
    draw_triangle(W, H, ax, ay, bx, by, cx, cy, plot, span);
    draw_triangle(W, H, bx, by, cx, cy, dx, dy, plot, span);
}

// the whole table: [kQuadPatterns][kQuadRows] words
static inline void quad_table_fill(uint32_t* table) {
    for (int gy = -kQuadR; gy <= kQuadR; gy++)
        for (int gx = -kQuadR; gx <= kQuadR; gx++)
            for (int r1 = 0; r1 < 9; r1++)
                for (int r2 = 0; r2 < 9; r2++) {
                    const int r1x = r1 % 3 - 1, r1y = r1 / 3 - 1, r2x = r2 % 3 - 1, r2y = r2 / 3 - 1;
                    quad_pattern_rows(gx, gy, r1x, r1y, r2x, r2y, table + (size_t)quad_pattern_index(gx, gy, r1x, r1y, r2x, r2y) * kQuadRows);
                }
}

}  // namespace tds
