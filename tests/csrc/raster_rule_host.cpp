// Host build of the product's triangle coverage rule (torchdrivesim_b200/csrc/tds_raster_tri.h)
// so that the exact code the CUDA kernel runs can be checked against cv2 on the CPU.
#include <stdint.h>
#include <string.h>
#include "tds_raster_tri.h"

extern "C" void tds_host_draw_triangle(uint8_t* img, int W, int H, const int32_t* p) {
    tds::draw_triangle(W, H, p[0], p[1], p[2], p[3], p[4], p[5],
        [&](int x, int y) { img[y * W + x] = 1; },
        [&](int y, int xa, int xb) { for (int x = xa; x <= xb; x++) img[y * W + x] = 1; });
}

// fast path (|coordinates| < 8192), x-major target like the kernel's shared-memory tile
extern "C" void tds_host_draw_triangle_fast(uint8_t* img, int W, int H, const int32_t* p) {
    // img is row-major [H][W]: index = x * 1 + y * W
    tds::draw_triangle_fast(W, H, 1, W, p[0], p[1], p[2], p[3], p[4], p[5],
        [&](int idx) { img[idx] = 1; },
        [&](int idx, int n, int step) { for (int i = 0; i < n; i++) img[idx + i * step] = 1; });
}

// thin path: returns 0 when the triangle does not qualify (not drawn), 1 when drawn
extern "C" int tds_host_draw_triangle_thin(uint8_t* img, int W, int H, const int32_t* p) {
    if (!tds::is_thin_inside(W, H, p[0], p[1], p[2], p[3], p[4], p[5])) return 0;
    tds::draw_triangle_thin(1, W, p[0], p[1], p[2], p[3], p[4], p[5],
        [&](int idx, int n, int step) { for (int i = 0; i < n; i++) img[idx + i * step] = 1; });
    return 1;
}

// row-run form used by the bitplane raster kernel (tds_raster_rows.h): rows visited in increasing order
#include "tds_raster_rows.h"
extern "C" void tds_host_draw_triangle_rows(uint8_t* img, int W, int H, const int32_t* p) {
    tds::RowTri t;
    tds::row_tri_setup(W, H, p[0], p[1], p[2], p[3], p[4], p[5], t, [](int dy) { return tds::row_rcp(dy); });
    for (int y = t.ylo; y <= t.yhi; y++)
        tds::row_tri_step(t, W, y, [&](int lo, int hi) {
            if (lo < 0 || hi >= W || lo > hi) { img[0] = 99; return; }      // contract violation: flagged
            for (int x = lo; x <= hi; x++) img[y * W + x] = 1;
        });
}

// the same rule spread over four cooperating workers - three outline edges + a fourth of the fill rows each (what the
// kernel runs, one worker per lane, for faces that cross the image border).  The host plays the workers in turn: the
// fill set-up of worker 3 is computed first and handed to the others through `share`.
extern "C" void tds_host_draw_triangle_by_parts(uint8_t* img, int W, int H, const int32_t* p) {
    auto rcp = [](int dy) { return tds::row_rcp(dy); };
    auto emit = [&](int y, int lo, int hi) {
        if (lo < 0 || hi >= W || lo > hi || y < 0 || y >= H) { img[0] = 99; return; }      // contract violation: flagged
        for (int x = lo; x <= hi; x++) img[y * W + x] = 1;
    };
    // pass 1: workers 0..2 compute what they share (their fill slope; nothing is drawn); pass 2: all four workers with
    // every shared value available
    int held[3] = {0, 0, 0};
    for (int part = 0; part < 3; part++) {
        int mine = 0, calls = 0;
        tds::row_tri_part(W, H, p[0], p[1], p[2], p[3], p[4], p[5], part, rcp, [](int, int, int) {},
                          [&](int v, int) { if (calls++ == 0) mine = v; return 0; });
        held[part] = mine;
    }
    for (int part = 0; part < 4; part++)
        tds::row_tri_part(W, H, p[0], p[1], p[2], p[3], p[4], p[5], part, rcp, emit, [&](int, int w) { return held[w]; });
}

// fast path of the bitplane kernel: all vertices inside the image, one interval per row.
// small != 0 uses the reciprocal-table slopes (images up to 128 pixels).
extern "C" int tds_host_draw_triangle_inside(uint8_t* img, int W, int H, const int32_t* p, int small) {
    for (int k = 0; k < 6; k++) if (p[k] < 0 || p[k] >= ((k & 1) ? H : W)) return 0;
    tds::FastTri t;
    auto rcp = [](int dy) { return tds::row_rcp(dy); };
    if (small) tds::fast_tri_setup<true>(p[0], p[1], p[2], p[3], p[4], p[5], t, rcp);
    else tds::fast_tri_setup<false>(p[0], p[1], p[2], p[3], p[4], p[5], t, rcp);
    tds::fast_tri_rows(t, [&](int y, int lo, int hi) {
        if (lo < 0 || hi >= W || lo > hi || y < 0 || y >= H) { img[0] = 99; return; }
        for (int x = lo; x <= hi; x++) img[y * W + x] = 1;
    });
    return 1;
}

// stateless row rule (tds_raster_rows_at.h): rows evaluated from the set-up alone, here bottom-up and every row twice
#include "tds_raster_rows_at.h"
extern "C" void tds_host_draw_triangle_rows_at(uint8_t* img, int W, int H, const int32_t* p) {
    tds::RowTri t;
    tds::row_tri_setup(W, H, p[0], p[1], p[2], p[3], p[4], p[5], t, [](int dy) { return tds::row_rcp(dy); });
    for (int pass = 0; pass < 2; pass++)
        for (int y = t.yhi; y >= t.ylo; y--)
            tds::row_tri_at(t, W, y, [&](int lo, int hi) {
                if (lo < 0 || hi >= W || lo > hi) { img[0] = 99; return; }
                for (int x = lo; x <= hi; x++) img[y * W + x] = 1;
            });
}

// a triangle inside the image (<= 128 pixels) as three line walkers (tds_raster_rows.h): what the 64x64 bitplane kernel
// runs for faces inside the image; 0 when a vertex is outside (nothing drawn)
extern "C" int tds_host_draw_triangle_lines3(uint8_t* img, int W, int H, const int32_t* p) {
    for (int k = 0; k < 6; k++) if (p[k] < 0 || p[k] >= ((k & 1) ? H : W)) return 0;
    tds::Lines3 q;
    tds::lines3_setup(p[0], p[1], p[2], p[3], p[4], p[5], q, [](int dy) { return tds::row_rcp(dy); });
    tds::lines3_rows(q, [&](int yy, int lo, int hi) {
        if (lo < 0 || hi >= W || lo > hi || yy < 0 || yy >= H) { img[0] = 99; return; }      // contract violation: flagged
        for (int xx = lo; xx <= hi; xx++) img[yy * W + xx] = 1;
    });
    return 1;
}

// every triangle with vertices in a (n x n) window, as three line walkers, against the reference's rule
extern "C" long long tds_host_lines3_triangle_sweep(int n) {
    const int W = 64, H = 64;
    static uint8_t ref[64 * 64], got[64 * 64];
    long long bad = 0;
    if (n > 16) return -1;
    for (int a = 0; a < n * n; a++)
        for (int b = 0; b < n * n; b++)
            for (int c = 0; c < n * n; c++) {
                const int32_t p[6] = {20 + a % n, 20 + a / n, 20 + b % n, 20 + b / n, 20 + c % n, 20 + c / n};
                memset(got, 0, sizeof(got));
                memset(ref, 0, sizeof(ref));
                tds_host_draw_triangle_lines3(got, W, H, p);
                tds::draw_triangle(W, H, p[0], p[1], p[2], p[3], p[4], p[5], [&](int x, int y) { ref[y * W + x] = 1; },
                                   [&](int y, int xa, int xb) { for (int x = xa; x <= xb; x++) ref[y * W + x] = 1; });
                if (memcmp(ref, got, sizeof(ref)) != 0) bad++;
            }
    return bad;
}

// sliver quads (tds_quad_table.h): faces (v0, v1, v2), (v1, v2, v3) inside the image, drawn the way the kernel does it -
// the rows of the coverage pattern shifted to the quad's corner.  `table` = tds_host_quad_table().  Returns 0 when the quad
// is not in the table (nothing drawn), 1 / 2 for the naming of the rungs that matched.
#include "tds_quad_table.h"
#include <vector>
extern "C" const uint32_t* tds_host_quad_table() {
    static std::vector<uint32_t> t;
    if (t.empty()) { t.resize((size_t)tds::kQuadPatterns * tds::kQuadRows); tds::quad_table_fill(t.data()); }
    return t.data();
}
extern "C" int tds_host_quad_table_shape(int* out) { out[0] = tds::kQuadPatterns; out[1] = tds::kQuadRows; out[2] = tds::kQuadR; return 0; }

extern "C" int tds_host_draw_quad(uint8_t* img, int W, int H, const int32_t* p) {
    for (int k = 0; k < 8; k++) if (p[k] < 0 || p[k] >= ((k & 1) ? H : W)) return 0;
    int gx = p[2] - p[0], gy = p[3] - p[1], r1x = p[4] - p[0], r1y = p[5] - p[1], r2x = p[6] - p[2], r2y = p[7] - p[3], naming = 1;
    if (!tds::quad_in_table(gx, gy, r1x, r1y, r2x, r2y)) {
        int t = gx; gx = r1x; r1x = t; t = gy; gy = r1y; r1y = t;
        r2x = p[6] - p[4]; r2y = p[7] - p[5];
        naming = 2;
    }
    if (!tds::quad_in_table(gx, gy, r1x, r1y, r2x, r2y)) return 0;
    int xmin = p[0], ymin = p[1], ymax = p[1];
    for (int k = 1; k < 4; k++) { xmin = p[2 * k] < xmin ? p[2 * k] : xmin; ymin = p[2 * k + 1] < ymin ? p[2 * k + 1] : ymin; ymax = p[2 * k + 1] > ymax ? p[2 * k + 1] : ymax; }
    const uint32_t* rows = tds_host_quad_table() + (size_t)tds::quad_pattern_index(gx, gy, r1x, r1y, r2x, r2y) * tds::kQuadRows;
    const int h = ymax - ymin + 1;
    for (int r = 0; r < ((h + 3) & ~3); r++) {
        if (r >= tds::kQuadRows) { img[0] = 99; break; }
        const unsigned long long m = (unsigned long long)rows[r] << xmin;
        const int y = ymin + r < H - 1 ? ymin + r : H - 1;
        if (W < 64 && (m >> W)) { img[0] = 99; }                                                        // a pattern wider than the image: flagged
        if (r >= h && m) { img[0] = 99; }                                                    // rows past the quad must be empty
        for (int x = 0; x < W; x++) if ((m >> x) & 1ull) img[y * W + x] = 1;
    }
    return naming;
}
