// Host build of the product's triangle coverage rule (torchdrivesim_b200/csrc/tds_raster_tri.h)
// so that the exact code the CUDA kernel runs can be checked against cv2 on the CPU.
#include <stdint.h>
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
    int held[10], n = 0, k = 0;
    // pass 1: worker 3 records what it shares; pass 2..4: workers 0..2 receive it
    tds::row_tri_part(W, H, p[0], p[1], p[2], p[3], p[4], p[5], 3, rcp, emit, [&](int v) { held[n++] = v; return v; });
    for (int part = 2; part >= 0; part--) {
        k = 0;
        tds::row_tri_part(W, H, p[0], p[1], p[2], p[3], p[4], p[5], part, rcp, emit, [&](int) { return held[k++]; });
    }
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
