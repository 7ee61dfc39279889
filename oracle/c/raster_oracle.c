/*
 * ORACLE — TEST INFRASTRUCTURE ONLY (never on the product path).
 *
 * CPU restatement of the reference's cv2 birdview backend for ONE camera:
 *   torchdrivesim/rendering/cv2.py:27-70   CV2Renderer.render_rgb_mesh
 *   torchdrivesim/rendering/base.py:102-130 Cameras.transform_points_screen / reverse_...
 *   torchdrivesim/mesh.py:308-348,506-521  trim (keep a face iff any vertex inside the 1.05x view quad)
 *   torchdrivesim/utils.py:99-122          is_inside_polygon
 * plus a restatement of the third-party raster rule the reference calls at
 * rendering/cv2.py:59: OpenCV cv::fillConvexPoly (opencv-python, unpinned in the
 * reference's pyproject.toml:27; pinned here by tests against the image's
 * opencv-python-headless 4.13.0): clipLine -> 8-connected LineIterator outline ->
 * 16.16 fixed-point scan fill.  On a float32 image LINE_AA degrades to LINE_8.
 *
 * All float arithmetic is plain fp32 with no contraction (build with
 * -ffp-contract=off): this was checked to be bit-identical to the reference's
 * torch CPU ops (separately rounded mul/add, sequential 4-term mean).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef long long i64;

/* ---- OpenCV clipLine(Size, Point&, Point&) restated ---- */
static int clip_line(i64 W, i64 H, i64 *px1, i64 *py1, i64 *px2, i64 *py2)
{
    i64 x1 = *px1, y1 = *py1, x2 = *px2, y2 = *py2;
    i64 right = W - 1, bottom = H - 1;
    int c1, c2;
    if (W <= 0 || H <= 0) return 0;
    c1 = (x1 < 0) + (x1 > right) * 2 + (y1 < 0) * 4 + (y1 > bottom) * 8;
    c2 = (x2 < 0) + (x2 > right) * 2 + (y2 < 0) * 4 + (y2 > bottom) * 8;
    if ((c1 & c2) == 0 && (c1 | c2) != 0) {
        i64 a;
        if (c1 & 12) {
            a = c1 < 8 ? 0 : bottom;
            x1 += (i64)((double)(a - y1) * (double)(x2 - x1) / (double)(y2 - y1));
            y1 = a;
            c1 = (x1 < 0) + (x1 > right) * 2;
        }
        if (c2 & 12) {
            a = c2 < 8 ? 0 : bottom;
            x2 += (i64)((double)(a - y2) * (double)(x2 - x1) / (double)(y2 - y1));
            y2 = a;
            c2 = (x2 < 0) + (x2 > right) * 2;
        }
        if ((c1 & c2) == 0 && (c1 | c2) != 0) {
            if (c1) {
                a = c1 == 1 ? 0 : right;
                y1 += (i64)((double)(a - x1) * (double)(y2 - y1) / (double)(x2 - x1));
                x1 = a;
                c1 = 0;
            }
            if (c2) {
                a = c2 == 1 ? 0 : right;
                y2 += (i64)((double)(a - x2) * (double)(y2 - y1) / (double)(x2 - x1));
                x2 = a;
                c2 = 0;
            }
        }
    }
    *px1 = x1; *py1 = y1; *px2 = x2; *py2 = y2;
    return (c1 | c2) == 0;
}

/* ---- cv::Line with connectivity 8 (LineIterator, leftToRight = true) ---- */
static void line8(int32_t *img, int W, int H, i64 x1, i64 y1, i64 x2, i64 y2, int32_t col)
{
    if ((uint64_t)x1 >= (uint64_t)W || (uint64_t)x2 >= (uint64_t)W ||
        (uint64_t)y1 >= (uint64_t)H || (uint64_t)y2 >= (uint64_t)H) {
        if (!clip_line(W, H, &x1, &y1, &x2, &y2)) return;
    }
    i64 dx = x2 - x1, dy = y2 - y1;
    i64 x = x1, y = y1;
    if (dx < 0) { dx = -dx; dy = -dy; x = x2; y = y2; }
    i64 sy = 1;
    if (dy < 0) { dy = -dy; sy = -1; }
    int vert = dy > dx;
    if (vert) { i64 t = dx; dx = dy; dy = t; }
    i64 err = dx - (dy + dy);
    i64 plus = dx + dx, minus = -(dy + dy);
    i64 count = dx + 1;
    for (i64 i = 0; i < count; i++) {
        img[y * W + x] = col;
        int m = err < 0;
        err += minus + (m ? plus : 0);
        if (vert) { y += sy; if (m) x += 1; }
        else      { x += 1;  if (m) y += sy; }
    }
}

/* ---- cv::FillConvexPoly (shift = 0, line_type = 8) ---- */
#define XY_SHIFT 16
#define XY_ONE (1 << XY_SHIFT)
void oracle_fill_convex_poly(int32_t *img, int W, int H, const int32_t *pts /*[n][2]*/, int npts, int32_t col)
{
    struct { int idx, di; i64 x, dx; int ye; } edge[2];
    int i, y, imin = 0, edges = npts;
    i64 xmin, xmax, ymin, ymax;
    const i64 delta1 = XY_ONE >> 1, delta2 = XY_ONE >> 1;
    i64 p0x = pts[2 * (npts - 1)], p0y = pts[2 * (npts - 1) + 1];
    xmin = xmax = pts[0];
    ymin = ymax = pts[1];
    for (i = 0; i < npts; i++) {
        i64 px = pts[2 * i], py = pts[2 * i + 1];
        if (py < ymin) { ymin = py; imin = i; }
        if (py > ymax) ymax = py;
        if (px > xmax) xmax = px;
        if (px < xmin) xmin = px;
        line8(img, W, H, p0x, p0y, px, py, col);
        p0x = px; p0y = py;
    }
    if (npts < 3 || (int)xmax < 0 || (int)ymax < 0 || (int)xmin >= W || (int)ymin >= H) return;
    if (ymax > H - 1) ymax = H - 1;
    edge[0].idx = edge[1].idx = imin;
    edge[0].ye = edge[1].ye = y = (int)ymin;
    edge[0].di = 1;
    edge[1].di = npts - 1;
    edge[0].x = edge[1].x = -XY_ONE;
    edge[0].dx = edge[1].dx = 0;
    do {
        for (i = 0; i < 2; i++) {
            if (y >= edge[i].ye) {
                int idx0 = edge[i].idx, di = edge[i].di;
                int idx = idx0 + di;
                if (idx >= npts) idx -= npts;
                int ty = 0;
                for (; edges-- > 0;) {
                    ty = pts[2 * idx + 1];
                    if (ty > y) {
                        i64 xs = (i64)pts[2 * idx0] << XY_SHIFT;
                        i64 xe = (i64)pts[2 * idx] << XY_SHIFT;
                        edge[i].ye = ty;
                        edge[i].dx = ((xe - xs) * 2 + (ty - y)) / (2 * (i64)(ty - y));
                        edge[i].x = xs;
                        edge[i].idx = idx;
                        break;
                    }
                    idx0 = idx;
                    idx += di;
                    if (idx >= npts) idx -= npts;
                }
            }
        }
        if (edges < 0) break;
        if (y >= 0) {
            int left = 0, right = 1;
            if (edge[0].x > edge[1].x) { left = 1; right = 0; }
            int xx1 = (int)((edge[left].x + delta1) >> XY_SHIFT);
            int xx2 = (int)((edge[right].x + delta2) >> XY_SHIFT);
            if (xx2 >= 0 && xx1 < W) {
                if (xx1 < 0) xx1 = 0;
                if (xx2 >= W) xx2 = W - 1;
                for (int xx = xx1; xx <= xx2; xx++) img[(i64)y * W + xx] = col;
            }
        }
        edge[0].x += edge[0].dx;
        edge[1].x += edge[1].dx;
    } while (++y <= (int)ymax);
}

/* view quad in camera-translated world space, rendering/cv2.py:34-40 + base.py:117-130 */
static void view_quad(float S, float C, float scale, int H, int W, float qx[4], float qy[4])
{
    const float cx[4] = {0.f, 0.f, (float)W, (float)W};
    const float cy[4] = {0.f, (float)H, (float)H, 0.f};
    const float half_min = (float)((H < W ? H : W) / 2.0);
    const float hw = (float)W / 2.0f, hh = (float)H / 2.0f;
    float mx = 0.f, my = 0.f;
    for (int i = 0; i < 4; i++) {
        float x = cx[i] - hw, y = cy[i] - hh;
        x = x / half_min; y = y / half_min;
        x = (-x) / scale; y = (-y) / scale;
        float a = C * x, b = (-S) * y;
        float c = S * x, d = C * y;
        qx[i] = a + b;           /* rot_mat^T row 0 = [C, -S] */
        qy[i] = c + d;           /* rot_mat^T row 1 = [S,  C] */
        qx[i] = qx[i] + 0.0f;    /* + cameras.xy (zero after the translate) */
        qy[i] = qy[i] + 0.0f;
    }
    mx = ((qx[0] + qx[1]) + qx[2]) + qx[3]; mx = mx / 4.0f;
    my = ((qy[0] + qy[1]) + qy[2]) + qy[3]; my = my / 4.0f;
    for (int i = 0; i < 4; i++) {
        float ex = (qx[i] - mx) * 1.05f, ey = (qy[i] - my) * 1.05f;
        qx[i] = mx + ex;
        qy[i] = my + ey;
    }
}

/*
 * Renders one camera.  verts are WORLD coordinates; face_rank is the painter's
 * draw order key (drawn in ascending rank, stable; rank = descending z with the
 * documented tie-break); face_rgb the uint8 colour; out is [3][H][W] float32
 * already transposed as the reference returns it (out[c][i][j] = img[j][i][c]).
 * Returns the number of faces kept by the cull (for statistics).
 */
int oracle_render_camera(const float *verts, int nv, const int32_t *faces, int nf,
                         const uint8_t *face_rank, const uint8_t *face_rgb,
                         float cam_x, float cam_y, float cam_sin, float cam_cos, float scale,
                         int H, int W, float *out)
{
    float qx[4], qy[4], ea[4], eb[4], ec[4];
    const float S = cam_sin, C = cam_cos;
    view_quad(S, C, scale, H, W, qx, qy);
    for (int i = 0; i < 4; i++) {
        int j = (i + 1) & 3;
        ea[i] = qy[j] - qy[i];
        eb[i] = qx[i] - qx[j];
        float t0 = (-ea[i]) * qx[i], t1 = eb[i] * qy[i];
        ec[i] = t0 - t1;
    }
    float *px = (float *)malloc(sizeof(float) * (size_t)(nv > 0 ? nv : 1));
    float *py = (float *)malloc(sizeof(float) * (size_t)(nv > 0 ? nv : 1));
    uint8_t *inside = (uint8_t *)malloc((size_t)(nv > 0 ? nv : 1));
    const float ncx = -cam_x, ncy = -cam_y;
    for (int v = 0; v < nv; v++) {
        float x = verts[2 * v] + ncx, y = verts[2 * v + 1] + ncy;
        px[v] = x; py[v] = y;
        int nr = 0;
        for (int i = 0; i < 4; i++) {
            float t0 = ea[i] * x, t1 = eb[i] * y;
            float s = (t0 + t1) + ec[i];
            nr += (s >= 0.0f);
        }
        inside[v] = (nr == 4) || (nr == 0);
    }
    /* kept faces, stable counting sort by rank */
    int count[257];
    memset(count, 0, sizeof(count));
    int kept = 0;
    for (int f = 0; f < nf; f++) {
        const int32_t *fc = faces + 3 * f;
        if (inside[fc[0]] | inside[fc[1]] | inside[fc[2]]) { count[face_rank[f] + 1]++; kept++; }
    }
    for (int r = 0; r < 256; r++) count[r + 1] += count[r];
    int *order = (int *)malloc(sizeof(int) * (size_t)(kept > 0 ? kept : 1));
    for (int f = 0; f < nf; f++) {
        const int32_t *fc = faces + 3 * f;
        if (inside[fc[0]] | inside[fc[1]] | inside[fc[2]]) order[count[face_rank[f]]++] = f;
    }
    /* owner image: -1 = background, else the face index of the last painter */
    int32_t *owner = (int32_t *)malloc(sizeof(int32_t) * (size_t)H * W);
    for (int i = 0; i < H * W; i++) owner[i] = -1;
    const float fmin = (float)(H < W ? H : W);
    const float hw = (float)W / 2.0f, hh = (float)H / 2.0f;
    for (int k = 0; k < kept; k++) {
        int f = order[k];
        int32_t pts[6];
        for (int c = 0; c < 3; c++) {
            int v = faces[3 * f + c];
            float a = C * px[v], b = S * py[v];
            float u0 = a + b;
            float c0 = (-S) * px[v], d0 = C * py[v];
            float u1 = c0 + d0;
            u0 = (-u0) * scale; u1 = (-u1) * scale;
            u0 = u0 * fmin;     u1 = u1 * fmin;
            u0 = u0 / 2.0f;     u1 = u1 / 2.0f;
            u0 = u0 + hw;       u1 = u1 + hh;
            pts[2 * c] = (int32_t)u0;
            pts[2 * c + 1] = (int32_t)u1;
        }
        oracle_fill_convex_poly(owner, W, H, pts, 3, f);
    }
    /* transpose: out[c][i][j] = img[row j][col i][c]  (rendering/cv2.py:61) */
    for (int i = 0; i < W; i++)
        for (int j = 0; j < H; j++) {
            int f = owner[j * W + i];
            for (int c = 0; c < 3; c++)
                out[((size_t)c * W + i) * H + j] = f < 0 ? 0.0f : (float)face_rgb[3 * f + c];
        }
    free(px); free(py); free(inside); free(order); free(owner);
    return kept;
}
