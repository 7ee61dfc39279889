// Birdview rasteriser: one CTA per camera, image tile in shared memory, pixel-identical to the
// reference's cv2 backend.
//
// Reference pipeline replaced here (per camera):
//   BirdviewRGBMeshGenerator.generate      torchdrivesim/mesh.py:1053-1157  (scene assembly; 1.86 MB per camera)
//   BirdviewRenderer.render_frame          torchdrivesim/rendering/base.py:167-204
//   CV2Renderer.render_rgb_mesh            torchdrivesim/rendering/cv2.py:27-70
//     translate -> trim to the 1.05x view quad (mesh.py:308-348, utils.py:99-122) -> painter's order by z
//     -> project + truncate to int32 -> cv2.fillConvexPoly per triangle -> transpose
//
// Design
//   * the static mesh is never expanded per camera: the CTA walks the rows of the per-map vertex grid
//     that the view quad touches (one contiguous record range per grid row and class);
//   * painter's order = one pass per colour class in draw order with a block barrier in between, so
//     pixels are plain byte stores into a res x res shared-memory tile (no atomics);
//   * each thread scan-converts whole triangles with the closed-form cv2 rule (tds_raster_tri.h);
//   * the tile is stored x-major, which IS the reference's final transpose (cv2.py:61), then expanded
//     through a colour LUT to 3 float planes with coalesced streaming 128-bit stores.
// HBM traffic per camera: 12 * res^2 bytes of image out (49 KB at 64x64); the map records are L2 hits.
#include "tds_map.cuh"
#include "tds_raster_tri.h"

namespace {

using tds::kMaxRasterRows;
using tds::kMaxSlots;
using tds::MapDev;
using tds::MapSetDev;

struct PaletteDev {
    int32_t n_classes;
    int32_t order[TDS_MAX_CLASSES];               // classes in draw order
    float rgb[TDS_MAX_CLASSES + 1][3];            // [0] = background
    int32_t agent_type_class[TDS_MAX_AGENT_TYPES];
    int32_t direction_class;
    int32_t tl_state_class[TDS_MAX_TL_STATES];
    uint32_t dyn_mask;                            // classes that dynamic primitives may carry
};

// ------------------------------------------------------------------ dynamic primitives (per env)
// workspace layout per environment: float tri[T][6] followed by uint8 cls[Tpad]
__host__ __device__ inline int64_t ws_env_bytes(int T) { return (int64_t)T * 24 + ((T + 15) / 16) * 16; }

__global__ void __launch_bounds__(128) dyn_prep_kernel(int B, int N, int L, int R, const float* __restrict__ agent_state,
                                                       const float* __restrict__ agent_size,
                                                       const int32_t* __restrict__ agent_type,
                                                       const float* __restrict__ tl_corners,
                                                       const int32_t* __restrict__ tl_state,
                                                       const float* __restrict__ rect_corners,
                                                       const int32_t* __restrict__ rect_class, PaletteDev pal,
                                                       uint8_t* __restrict__ ws) {
    const int items = N + L + R;
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= (int64_t)B * items) return;
    const int b = (int)(g / items), it = (int)(g % items);
    const int T = 3 * N + 2 * L + 2 * R;
    float* tri = reinterpret_cast<float*>(ws + (int64_t)b * ws_env_bytes(T));
    uint8_t* cls = reinterpret_cast<uint8_t*>(tri + (int64_t)T * 6);
    if (it < N) {
        // agent rectangle + direction triangle: mesh.py:942-951, 911-940, utils.py:82-96
        const float* st = agent_state + ((int64_t)b * N + it) * 4;
        const float l = agent_size[((int64_t)b * N + it) * 2], w = agent_size[((int64_t)b * N + it) * 2 + 1];
        float s, c;
        tds::sincos_cr(st[2], s, c);
        const float hl = l * 0.5f, hw = w * 0.5f, nhl = (-l) * 0.5f, nhw = (-w) * 0.5f;
        const float off = l * 0.2f;                 // l * (0.5 - 0.3)
        const float lx[7] = {hl, hl, nhl, nhl, l * 0.3f + off, 0.0f + off, 0.0f + off};
        const float ly[7] = {hw, nhw, nhw, hw, 0.0f, hw + 0.0f, nhw + 0.0f};
        float wx[7], wy[7];
#pragma unroll
        for (int k = 0; k < 7; k++) {
            wx[k] = (c * lx[k] + (-s) * ly[k]) + st[0];
            wy[k] = (s * lx[k] + c * ly[k]) + st[1];
        }
        const int fi[3][3] = {{0, 1, 3}, {1, 3, 2}, {4, 5, 6}};
        int ty = agent_type ? agent_type[(int64_t)b * N + it] : 0;
        ty = min(max(ty, 0), TDS_MAX_AGENT_TYPES - 1);
        const int acls = pal.agent_type_class[ty];
#pragma unroll
        for (int f = 0; f < 3; f++) {
            float* o = tri + (int64_t)(3 * it + f) * 6;
#pragma unroll
            for (int k = 0; k < 3; k++) { o[2 * k] = wx[fi[f][k]]; o[2 * k + 1] = wy[fi[f][k]]; }
            const int cc = f < 2 ? acls : pal.direction_class;
            cls[3 * it + f] = cc < 0 ? 255 : (uint8_t)cc;
        }
    } else {
        // rectangles: traffic lights then extra static controls; faces [0,1,3], [1,3,2] (mesh.py:1274-1290)
        const bool is_tl = it < N + L;
        const int j = is_tl ? it - N : it - N - L;
        const float* cr = is_tl ? tl_corners + ((int64_t)b * L + j) * 8 : rect_corners + ((int64_t)b * R + j) * 8;
        int cc;
        if (is_tl) {
            int stt = tl_state[(int64_t)b * L + j];
            stt = min(max(stt, 0), TDS_MAX_TL_STATES - 1);
            cc = pal.tl_state_class[stt];
        } else {
            cc = rect_class[(int64_t)b * R + j];
        }
        const int t0 = 3 * N + (is_tl ? 2 * j : 2 * L + 2 * j);
        const int fi[2][3] = {{0, 1, 3}, {1, 3, 2}};
#pragma unroll
        for (int f = 0; f < 2; f++) {
            float* o = tri + (int64_t)(t0 + f) * 6;
#pragma unroll
            for (int k = 0; k < 3; k++) { o[2 * k] = cr[2 * fi[f][k]]; o[2 * k + 1] = cr[2 * fi[f][k] + 1]; }
            cls[t0 + f] = (cc < 0 || cc >= TDS_MAX_CLASSES) ? 255 : (uint8_t)cc;
        }
    }
}

// ------------------------------------------------------------------ camera
struct Camera {
    float ncx, ncy;            // -camera position
    float S, C, scale, fmin, half;
    const float* edges;        // view-quad edge functions a[4], b[4], c[4] (utils.py:99-122), in shared memory
    int res;
};

__device__ __forceinline__ void make_camera(Camera& cam, float cx, float cy, float S, float C, float scale, int res,
                                            float qx[4], float qy[4], float* edges, bool write_edges) {
    cam.ncx = -cx; cam.ncy = -cy; cam.S = S; cam.C = C; cam.scale = scale; cam.res = res;
    cam.fmin = (float)res;
    cam.half = (float)res / 2.0f;
    // rendering/cv2.py:34-40 with base.py:117-130 (cameras.xy is zero after the translate)
    const float cxs[4] = {0.f, 0.f, (float)res, (float)res};
    const float cys[4] = {0.f, (float)res, (float)res, 0.f};
#pragma unroll
    for (int i = 0; i < 4; i++) {
        float x = cxs[i] - cam.half, y = cys[i] - cam.half;
        x = x / cam.half; y = y / cam.half;
        x = (-x) / scale; y = (-y) / scale;
        qx[i] = (C * x + (-S) * y) + 0.0f;
        qy[i] = (S * x + C * y) + 0.0f;
    }
    const float mx = (((qx[0] + qx[1]) + qx[2]) + qx[3]) / 4.0f;
    const float my = (((qy[0] + qy[1]) + qy[2]) + qy[3]) / 4.0f;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        qx[i] = mx + (qx[i] - mx) * 1.05f;
        qy[i] = my + (qy[i] - my) * 1.05f;
    }
    cam.edges = edges;
    if (write_edges) {
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int j = (i + 1) & 3;
            const float ea = qy[j] - qy[i], eb = qx[i] - qx[j];
            edges[i] = ea;
            edges[4 + i] = eb;
            edges[8 + i] = (-ea) * qx[i] - eb * qy[i];
        }
    }
}

__device__ __forceinline__ bool inside_quad(const Camera& cam, float x, float y) {
    int nr = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) nr += ((cam.edges[i] * x + cam.edges[4 + i] * y) + cam.edges[8 + i]) >= 0.0f;
    return nr == 4 || nr == 0;
}

// ---- stage 1: cull (mesh.py:311-313) + project one world-space triangle --------------------------
// Pixel-space shortcut for the cull: the 1.05x view quad is the image square scaled by 1.05 about its
// centre, i.e. pixel coordinates in [-0.025 res, 1.025 res].  A projected vertex further than ~0.01 px from
// that boundary is decided from its pixel coordinates; the thin band around the boundary falls back to
// the reference's fp32 edge functions, so the decision is always the reference's.
struct Item {
    uint32_t a, b, c;       // (x0,y0), (x1,y1), (x2,y2) as int16 pairs
};

__device__ __forceinline__ int classify(float u, float v, float mid, float r_in, float r_out) {
    // 1 inside, 0 outside, -1 undecided: the quad is the max-norm ball of radius 1.05 res / 2 around the image centre
    const float d = fmaxf(fabsf(u - mid), fabsf(v - mid));
    return d < r_in ? 1 : (d > r_out ? 0 : -1);        // NaN falls through to "undecided"
}

__device__ __forceinline__ void project_f(const Camera& cam, float x, float y, float& u0, float& u1) {
    u0 = cam.C * x + cam.S * y;
    u1 = (-cam.S) * x + cam.C * y;
    u0 = (-u0) * cam.scale; u1 = (-u1) * cam.scale;
    u0 = u0 * cam.fmin;     u1 = u1 * cam.fmin;
    u0 = u0 / 2.0f;         u1 = u1 / 2.0f;
    u0 = u0 + cam.half;     u1 = u1 + cam.half;
}

// returns 0: culled / invisible, 1: thin item, 3: general item (|coords| < 8192), 2: kept but needs the 64-bit
// slow path (ints in xy[])
__device__ __forceinline__ int setup_triangle(const Camera& cam, float x0, float y0, float x1, float y1, float x2,
                                              float y2, int own, Item& it, int xy[6]) {
    const float px0 = x0 + cam.ncx, py0 = y0 + cam.ncy;
    const float px1 = x1 + cam.ncx, py1 = y1 + cam.ncy;
    const float px2 = x2 + cam.ncx, py2 = y2 + cam.ncy;
    float u0, v0, u1, v1, u2, v2;
    project_f(cam, px0, py0, u0, v0);
    project_f(cam, px1, py1, u1, v1);
    project_f(cam, px2, py2, u2, v2);
    // fp32 disagreement between the pixel-space and the edge-function test is < 1e-4 m; band = 0.01 px + 2e-3 m
    const float band = 0.01f + 0.002f * (cam.scale * cam.half);
    const float r_in = 0.525f * cam.fmin - band, r_out = 0.525f * cam.fmin + band;
    int c0 = classify(u0, v0, cam.half, r_in, r_out), c1 = classify(u1, v1, cam.half, r_in, r_out),
        c2 = classify(u2, v2, cam.half, r_in, r_out);
    if ((c0 | c1 | c2) < 0) {       // some vertex is within the band around the quad boundary (or NaN): exact test
        // one rolled loop over the three vertices keeps this rare path small; the edge functions live in smem
        float ex = px0, ey = py0;
#pragma unroll 1
        for (int k = 0; k < 3; k++) {
            const int r = inside_quad(cam, ex, ey) ? 1 : 0;
            if (k == 0) { if (c0 < 0) c0 = r; ex = px1; ey = py1; }
            else if (k == 1) { if (c1 < 0) c1 = r; ex = px2; ey = py2; }
            else { if (c2 < 0) c2 = r; }
        }
    }
    if (!(c0 | c1 | c2)) return 0;
    const int first = c0 ? 0 : (c1 ? 1 : 2);
    if (!((own >> first) & 1)) return 0;          // another cell's copy of this face draws it
    xy[0] = __float2int_rz(u0); xy[1] = __float2int_rz(v0);
    xy[2] = __float2int_rz(u1); xy[3] = __float2int_rz(v1);
    xy[4] = __float2int_rz(u2); xy[5] = __float2int_rz(v2);
    int big = 0;
#pragma unroll
    for (int k = 0; k < 6; k++) big |= (xy[k] <= -8192) | (xy[k] >= 8192);
    if (big) return 2;
    // integer bounding box entirely off the image: clipLine rejects all three edges and the fill returns early
    const int res = cam.res;
    const int xmin = min(min(xy[0], xy[2]), xy[4]), xmax = max(max(xy[0], xy[2]), xy[4]);
    const int ymin = min(min(xy[1], xy[3]), xy[5]), ymax = max(max(xy[1], xy[3]), xy[5]);
    if (xmax < 0 || ymax < 0 || xmin >= res || ymin >= res) return 0;
    it.a = (uint32_t)(xy[0] & 0xffff) | ((uint32_t)xy[1] << 16);
    it.b = (uint32_t)(xy[2] & 0xffff) | ((uint32_t)xy[3] << 16);
    it.c = (uint32_t)(xy[4] & 0xffff) | ((uint32_t)xy[5] << 16);
    // thin = at most two rows and fully inside the image: closed-form runs (tds_raster_tri.h)
    const bool thin = (ymax - ymin <= 1) & (xmin >= 0) & (ymin >= 0) & (xmax < res) & (ymax < res);
    return thin ? 1 : 3;
}

// ---- stage 2: scan-convert.  The tile is x-major (img[x * res + y]) which IS the transpose of cv2.py:61
__device__ __forceinline__ void draw_item(uint8_t* img, int res, uint8_t val, const Item& it) {
    const int x0 = (int16_t)(it.a & 0xffff), y0 = (int32_t)it.a >> 16;
    const int x1 = (int16_t)(it.b & 0xffff), y1 = (int32_t)it.b >> 16;
    const int x2 = (int16_t)(it.c & 0xffff), y2 = (int32_t)it.c >> 16;
    tds::draw_triangle_fast(res, res, res, 1, x0, y0, x1, y1, x2, y2,
        [&](int idx) { img[idx] = val; },
        [&](int idx, int n, int step) { for (; n > 0; n--, idx += step) img[idx] = val; });
}

__device__ __forceinline__ void draw_item_thin(uint8_t* img, int res, uint8_t val, const Item& it) {
    const int x0 = (int16_t)(it.a & 0xffff), y0 = (int32_t)it.a >> 16;
    const int x1 = (int16_t)(it.b & 0xffff), y1 = (int32_t)it.b >> 16;
    const int x2 = (int16_t)(it.c & 0xffff), y2 = (int32_t)it.c >> 16;
    const int xmin = min(min(x0, x1), x2), xmax = max(max(x0, x1), x2);
    if (xmax - xmin <= 1) {      // within 2x2 pixels: the outline is the three vertices
        img[x0 * res + y0] = val;
        img[x1 * res + y1] = val;
        img[x2 * res + y2] = val;
        return;
    }
    tds::draw_triangle_thin(res, 1, x0, y0, x1, y1, x2, y2, [&](int idx, int n, int step) {
#pragma unroll 2
        for (; n > 0; n--, idx += step) img[idx] = val;
    });
}

__device__ __noinline__ void draw_item_slow(uint8_t* img, int res, uint8_t val, const int* xy) {
    tds::draw_triangle(res, res, xy[0], xy[1], xy[2], xy[3], xy[4], xy[5],
        [&](int x, int y) { img[x * res + y] = val; },
        [&](int y, int xa, int xb) { for (int x = xa; x <= xb; x++) img[x * res + y] = val; });
}

struct RasterArgs {
    const int32_t* env_map;
    const float* cam_xy;
    const float* cam_sc;
    const uint8_t* present;
    const uint8_t* ws;
    float* out;
    int32_t B, Nc, N, T, present_per_camera, res;
    int32_t ncam;
    float scale;
};

// per-group item queue: thin items grow from the front, general items from the back; a chunk of kChunk candidates
// can never overflow it.  12 B per item.
template <int G>
struct QCfg {
    static constexpr int total = G == 32 ? 128 : 2 * G;      // = candidates per chunk
};
constexpr int kRows = tds::kMaxRasterRows;
constexpr int kGroupExtra = kRows * 8 + 32 + 48;   // row tables, counters, view-quad edge functions

// G = threads cooperating on one camera: 32 (one warp per camera, 4 cameras per CTA, no block barriers)
// for tiles up to 64x64, or the whole CTA for larger tiles.
template <int G>
__device__ __forceinline__ void group_sync() {
    if (G == 32) __syncwarp();
    else __syncthreads();
}

#ifndef TDS_RASTER_MINB
#define TDS_RASTER_MINB 7
#endif
template <int G, int RES>
__global__ void __launch_bounds__(G == 32 ? 128 : G, G == 32 ? TDS_RASTER_MINB : 1) raster_kernel(MapSetDev maps, RasterArgs a, PaletteDev pal) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    constexpr int GROUPS = G == 32 ? 4 : 1;
    const int res = RES ? RES : a.res;          // RES = 64 is compiled with constant strides
    const int group = G == 32 ? (threadIdx.x >> 5) : 0;
    const int tid = G == 32 ? (threadIdx.x & 31) : threadIdx.x;
    const int camid = blockIdx.x * GROUPS + group;
    const bool active = camid < a.ncam;         // whole group is inactive together

    // per-group shared memory: tile | queue | row tables
    const int tile_bytes = res * res;
    constexpr int kQueue = QCfg<G>::total;
    const int group_bytes = tile_bytes + kQueue * 12 + kGroupExtra;
    uint8_t* base = smem_raw + (size_t)group * group_bytes;
    uint8_t* img = base;
    uint32_t* queue = reinterpret_cast<uint32_t*>(base + tile_bytes);      // [3][kQueue]
    int* s_start = reinterpret_cast<int*>(base + tile_bytes + kQueue * 12);
    int* s_pref = s_start + kRows;                                         // exclusive prefix of the row counts
    int* s_cnt = s_pref + kRows;                   // [0] thin count, [1] general count, [4] total
    float* s_edges = reinterpret_cast<float*>(s_cnt + 6);                   // [12]
    __shared__ float s_lut[(TDS_MAX_CLASSES + 1) * 3];
    for (int i = threadIdx.x; i < (TDS_MAX_CLASSES + 1) * 3; i += blockDim.x) s_lut[i] = (&pal.rgb[0][0])[i];
    __syncthreads();
    if (!active) return;                        // no block barrier is used below when G == 32

    const int b = camid / a.Nc;
    const MapDev& map = maps.m[a.env_map ? a.env_map[b] : 0];
    Camera cam;
    float qx[4], qy[4];
    const float2 cxy = reinterpret_cast<const float2*>(a.cam_xy)[camid];
    const float2 csc = reinterpret_cast<const float2*>(a.cam_sc)[camid];
    make_camera(cam, cxy.x, cxy.y, csc.x, csc.y, a.scale, res, qx, qy, s_edges, tid == 0);

    {   // clear the tile
        uint32_t* w = reinterpret_cast<uint32_t*>(img);
        for (int i = tid; i < tile_bytes / 4; i += G) w[i] = 0u;
        if (tid == 0) { s_cnt[0] = 0; s_cnt[1] = 0; s_cnt[2] = 0; s_cnt[3] = 0; }
    }
    // ---- grid rows touched by the view quad (world coordinates), with a 5 cm safety margin
    const float margin = 0.05f;
    float wqx[4], wqy[4];
#pragma unroll
    for (int i = 0; i < 4; i++) { wqx[i] = qx[i] + cxy.x; wqy[i] = qy[i] + cxy.y; }
    const float ymin = fminf(fminf(wqy[0], wqy[1]), fminf(wqy[2], wqy[3])) - margin;
    const float ymax = fmaxf(fmaxf(wqy[0], wqy[1]), fmaxf(wqy[2], wqy[3])) + margin;
    int r0 = (int)floorf((ymin - map.ry0) * map.rinv), r1 = (int)floorf((ymax - map.ry0) * map.rinv);
    r0 = max(r0, 0);
    r1 = min(r1, map.rgy - 1);
    const int nrows = min(max(r1 - r0 + 1, 0), kRows);
    // column interval of row (r0 + tid), kept in registers of lane/thread tid < nrows
    int my_c0 = 0, my_c1 = -1;
    if (tid < nrows) {
        const int r = r0 + tid;
        const float ylo = map.ry0 + (float)r * map.rcs - margin, yhi = map.ry0 + (float)(r + 1) * map.rcs + margin;
        float xmin = 3.0e38f, xmax = -3.0e38f;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int j = (i + 1) & 3;
            const float y0 = wqy[i], y1 = wqy[j], x0 = wqx[i], x1 = wqx[j];
            if (fmaxf(y0, y1) < ylo || fminf(y0, y1) > yhi) continue;
            float t0 = 0.f, t1 = 1.f;
            const float dy = y1 - y0;
            if (dy != 0.f) {
                float ta = (ylo - y0) / dy, tb = (yhi - y0) / dy;
                if (ta > tb) { const float t = ta; ta = tb; tb = t; }
                t0 = fmaxf(t0, ta);
                t1 = fminf(t1, tb);
            }
            const float xa = x0 + t0 * (x1 - x0), xb = x0 + t1 * (x1 - x0);
            xmin = fminf(xmin, fminf(xa, xb));
            xmax = fmaxf(xmax, fmaxf(xa, xb));
        }
        if (xmin <= xmax) {
            my_c0 = max((int)floorf((xmin - margin - map.rx0) * map.rinv), 0);
            my_c1 = min((int)floorf((xmax + margin - map.rx0) * map.rinv), map.rgx - 1);
        }
    }
    group_sync<G>();

    const int T = a.T;
    const float* dtri = reinterpret_cast<const float*>(a.ws + (int64_t)b * ws_env_bytes(T));
    const uint8_t* dcls = reinterpret_cast<const uint8_t*>(dtri + (int64_t)T * 6);
    const uint8_t* pres = a.present ? (a.present_per_camera ? a.present + (int64_t)camid * a.N : a.present + (int64_t)b * a.N)
                                    : nullptr;
    const int ncell = map.rgx * map.rgy;

    // ---- painter's passes, first drawn = highest z.  Each pass: stage 1 culls + projects candidates and
    // appends the kept ones to the queue (warp-aggregated), stage 2 scan-converts the queue.
    for (int ph = 0; ph < pal.n_classes; ph++) {
        const int c = pal.order[ph];
        const uint8_t val = (uint8_t)(c + 1);
        const int slot = map.slot_of_class[c];
        int total_static = 0;
        if (slot >= 0 && nrows > 0) {
            // record range of every touched grid row for this class
            if (tid < nrows) {
                int st = 0, cnt = 0;
                if (my_c1 >= my_c0) {
                    const int cb = slot * ncell + (r0 + tid) * map.rgx;
                    st = map.rcell[cb + my_c0];
                    cnt = map.rcell[cb + my_c1 + 1] - st;
                }
                s_start[tid] = st;
                s_pref[tid] = cnt;
            }
            group_sync<G>();
            if (tid == 0) {
                int acc = 0;
                for (int r = 0; r < nrows; r++) { const int n = s_pref[r]; s_pref[r] = acc; acc += n; }
                s_cnt[4] = acc;
            }
            group_sync<G>();
            total_static = s_cnt[4];
        }
        const int ndyn = ((pal.dyn_mask >> c) & 1u) ? T : 0;
        const int total = total_static + ndyn;
        int row = 0;                                   // candidates are visited in increasing order
        int n_thin_w = 0, n_gen_w = 0;                 // queue fill levels of a warp group
        // chunks of kQueue candidates: stage 1 (cull + project + classify) appends to the two-ended queue, stage 2
        // scan-converts it.  There is exactly ONE copy of each scan-conversion routine in the kernel: the
        // instruction-cache footprint matters more than anything else here.
        for (int c0 = 0; c0 < total; c0 += kQueue) {
            const int cend = min(c0 + kQueue, total);
            for (int i = c0 + tid; i < cend; i += G) {
                float x0, y0, x1, y1, x2, y2;
                int own = 7;
                bool valid = true;
                if (i < total_static) {
                    while (row + 1 < nrows && i >= s_pref[row + 1]) row++;
                    const int idx = s_start[row] + (i - s_pref[row]);
                    const float4 v01 = __ldg(map.rec + 2 * (int64_t)idx);
                    const float4 v2o = __ldg(map.rec + 2 * (int64_t)idx + 1);
                    x0 = v01.x; y0 = v01.y; x1 = v01.z; y1 = v01.w; x2 = v2o.x; y2 = v2o.y;
                    own = __float_as_int(v2o.z);
                } else {
                    // dynamic primitives (agents, direction triangles, traffic lights, signs)
                    int t = i - total_static;
                    bool degenerate = false;
                    if (t < 3 * a.N && pres && !pres[t / 3]) {
                        // absent agent: faces * 0 -> degenerate triangle at actor vertex 0 with agent 0's class
                        // (mesh.py:1083-1089)
                        t = 0;
                        degenerate = true;
                    }
                    valid = dcls[t] == c;
                    const float* p = dtri + (int64_t)t * 6;
                    x0 = p[0]; y0 = p[1];
                    x1 = degenerate ? x0 : p[2]; y1 = degenerate ? y0 : p[3];
                    x2 = degenerate ? x0 : p[4]; y2 = degenerate ? y0 : p[5];
                }
                Item it;
                int xy[6];
                const int kind = valid ? setup_triangle(cam, x0, y0, x1, y1, x2, y2, own, it, xy) : 0;
                if (kind == 2) draw_item_slow(img, res, val, xy);     // huge triangle: rare, drawn in place
                // thin items grow from the front of the queue, general items from the back (warp-aggregated)
                const unsigned full = __activemask();
                const int lane = threadIdx.x & 31, leader = __ffs(full) - 1;
                const unsigned mt = __ballot_sync(full, kind == 1), mg = __ballot_sync(full, kind == 3);
                if (mt | mg) {
                    int base_t = 0, base_g = 0;
                    if (G == 32) {
                        // one warp owns the queue: the counters live in (warp-uniform) registers
                        base_t = n_thin_w; base_g = n_gen_w;
                        n_thin_w += __popc(mt); n_gen_w += __popc(mg);
                    } else {
                        if (lane == leader) {
                            if (mt) base_t = atomicAdd(&s_cnt[0], __popc(mt));
                            if (mg) base_g = atomicAdd(&s_cnt[1], __popc(mg));
                        }
                        base_t = __shfl_sync(full, base_t, leader);
                        base_g = __shfl_sync(full, base_g, leader);
                    }
                    const unsigned below = (1u << lane) - 1;
                    int pos = -1;
                    if (kind == 1) pos = base_t + __popc(mt & below);
                    else if (kind == 3) pos = kQueue - 1 - (base_g + __popc(mg & below));
                    if (pos >= 0) { queue[pos] = it.a; queue[kQueue + pos] = it.b; queue[2 * kQueue + pos] = it.c; }
                }
            }
            // lanes that ran fewer rounds missed the last ballots: take the counts of the lane that ran them all
            if (G == 32) { n_thin_w = __shfl_sync(0xffffffffu, n_thin_w, 0); n_gen_w = __shfl_sync(0xffffffffu, n_gen_w, 0); }
            group_sync<G>();
            const int n_thin = G == 32 ? n_thin_w : s_cnt[0], n_gen = G == 32 ? n_gen_w : s_cnt[1];
            for (int k = tid; k < n_thin; k += G) {
                Item q;
                q.a = queue[k]; q.b = queue[kQueue + k]; q.c = queue[2 * kQueue + k];
                draw_item_thin(img, res, val, q);
            }
            for (int k = kQueue - 1 - tid; k >= kQueue - n_gen; k -= G) {
                Item q;
                q.a = queue[k]; q.b = queue[kQueue + k]; q.c = queue[2 * kQueue + k];
                draw_item(img, res, val, q);
            }
            group_sync<G>();
            if (G == 32) {
                n_thin_w = 0; n_gen_w = 0;
            } else {
                if (tid == 0) { s_cnt[0] = 0; s_cnt[1] = 0; }
                group_sync<G>();
            }
        }
    }
    group_sync<G>();

    // ---- expand the tile through the colour LUT: out[cam][ch][x][y], 4 pixels per 128-bit store
#ifdef TDS_EXP_NOOUT
    if (img[tid] != 77) return;
#endif
    const int nquad = tile_bytes / 4;
    float* outc = a.out + (int64_t)camid * 3 * tile_bytes;
    const uint32_t* w = reinterpret_cast<const uint32_t*>(img);
    for (int i = tid; i < nquad; i += G) {
        const uint32_t v = w[i];
        const int k0 = (v & 255u) * 3, k1 = ((v >> 8) & 255u) * 3, k2 = ((v >> 16) & 255u) * 3, k3 = (v >> 24) * 3;
#pragma unroll
        for (int ch = 0; ch < 3; ch++) {
            const float4 o = make_float4(s_lut[k0 + ch], s_lut[k1 + ch], s_lut[k2 + ch], s_lut[k3 + ch]);
            tds::st_cs_f4(reinterpret_cast<float4*>(outc + (int64_t)ch * tile_bytes) + i, o);
        }
    }
}

}  // namespace

static thread_local cudaEvent_t g_ev_start = nullptr, g_ev_stop = nullptr;

extern "C" void tds_raster_set_timing_events(void* start_event, void* stop_event) {
    g_ev_start = (cudaEvent_t)start_event;
    g_ev_stop = (cudaEvent_t)stop_event;
}

extern "C" int64_t tds_raster_workspace_bytes(int32_t B, int32_t N, int32_t L, int32_t R) {
    if (B < 0 || N < 0 || L < 0 || R < 0) return -1;
    return (int64_t)B * ws_env_bytes(3 * N + 2 * L + 2 * R) + 16;
}

extern "C" int tds_raster_birdview(const tds_map_t* const* maps, int32_t n_maps, const int32_t* d_env_map,
                                   int32_t B, int32_t Nc, int32_t N,
                                   const float* d_cam_xy, const float* d_cam_sc,
                                   const float* d_agent_state, const float* d_agent_size, const int32_t* d_agent_type,
                                   const uint8_t* d_present, int32_t present_per_camera,
                                   const float* d_tl_corners, const int32_t* d_tl_state, int32_t L,
                                   const float* d_rect_corners, const int32_t* d_rect_class, int32_t R,
                                   const tds_palette_t* palette, float scale, int32_t res,
                                   float* d_out, void* d_workspace, void* stream) {
    TDS_REQUIRE(B >= 0 && Nc >= 0 && N >= 0 && L >= 0 && R >= 0, "raster: negative size");
    if (B == 0 || Nc == 0) return TDS_OK;
    TDS_REQUIRE(d_cam_xy && d_cam_sc && d_out && palette, "raster: null pointer");
    TDS_REQUIRE(N == 0 || (d_agent_state && d_agent_size), "raster: null agent tensors");
    TDS_REQUIRE(L == 0 || (d_tl_corners && d_tl_state), "raster: null traffic light tensors");
    TDS_REQUIRE(R == 0 || (d_rect_corners && d_rect_class), "raster: null rectangle tensors");
    TDS_REQUIRE(res >= 4 && res % 4 == 0 && res <= 448, "raster: res=%d must be a multiple of 4 in [4,448]", res);
    TDS_REQUIRE(scale > 0.0f, "raster: scale must be positive");
    TDS_REQUIRE(palette->n_classes >= 0 && palette->n_classes <= TDS_MAX_CLASSES, "raster: bad palette");
    const int T = 3 * N + 2 * L + 2 * R;
    TDS_REQUIRE(T == 0 || d_workspace, "raster: null workspace");
    MapSetDev set;
    if (int e = tds::gather_maps(maps, n_maps, set)) return e;
    // the view quad's bounding box must fit kMaxRasterRows grid rows of every map
    const float fov = 2.0f / scale;
    for (int i = 0; i < n_maps; i++) {
        const float extent = 1.05f * 1.41422f * fov + 0.2f;
        if (extent / set.m[i].rcs + 2.0f > (float)kMaxRasterRows)
            return tds::fail(TDS_ERR_UNSUPPORTED, "raster: fov %.1f m needs more than %d grid rows of %.1f m; recreate the map with a larger raster_cell",
                             fov, kMaxRasterRows, set.m[i].rcs);
    }
    PaletteDev pal = {};
    // active classes sorted by rank (stable on class id) = painter's passes
    pal.n_classes = 0;
    for (int c = 0; c < palette->n_classes; c++)
        if (palette->active[c]) pal.order[pal.n_classes++] = c;
    for (int i = 1; i < pal.n_classes; i++)
        for (int j = i; j > 0 && palette->rank[pal.order[j]] < palette->rank[pal.order[j - 1]]; j--) {
            const int t = pal.order[j]; pal.order[j] = pal.order[j - 1]; pal.order[j - 1] = t;
        }
    for (int c = 0; c < palette->n_classes; c++)
        for (int k = 0; k < 3; k++) pal.rgb[c + 1][k] = (float)palette->rgb[c][k];
    for (int t = 0; t < TDS_MAX_AGENT_TYPES; t++) pal.agent_type_class[t] = palette->agent_type_class[t];
    for (int t = 0; t < TDS_MAX_TL_STATES; t++) pal.tl_state_class[t] = palette->tl_state_class[t];
    pal.direction_class = palette->direction_class;
    pal.dyn_mask = 0;
    if (N > 0) {
        for (int t = 0; t < TDS_MAX_AGENT_TYPES; t++)
            if (pal.agent_type_class[t] >= 0 && pal.agent_type_class[t] < TDS_MAX_CLASSES) pal.dyn_mask |= 1u << pal.agent_type_class[t];
        if (pal.direction_class >= 0 && pal.direction_class < TDS_MAX_CLASSES) pal.dyn_mask |= 1u << pal.direction_class;
    }
    if (L > 0)
        for (int t = 0; t < TDS_MAX_TL_STATES; t++)
            if (pal.tl_state_class[t] >= 0 && pal.tl_state_class[t] < TDS_MAX_CLASSES) pal.dyn_mask |= 1u << pal.tl_state_class[t];
    if (R > 0) pal.dyn_mask = 0xffffffffu;     // extra rectangles carry arbitrary classes

    cudaStream_t st = (cudaStream_t)stream;
    if (T > 0) {
        const int64_t items = (int64_t)B * (N + L + R);
        dyn_prep_kernel<<<(unsigned)((items + 127) / 128), 128, 0, st>>>(B, N, L, R, d_agent_state, d_agent_size, d_agent_type,
                                                                        d_tl_corners, d_tl_state, d_rect_corners, d_rect_class,
                                                                        pal, (uint8_t*)d_workspace);
        TDS_LAUNCH_OK();
    }
    RasterArgs a;
    a.env_map = d_env_map; a.cam_xy = d_cam_xy; a.cam_sc = d_cam_sc; a.present = d_present;
    a.ws = (const uint8_t*)d_workspace; a.out = d_out;
    a.B = B; a.Nc = Nc; a.N = N; a.T = T; a.present_per_camera = present_per_camera; a.res = res; a.scale = scale;
    const int64_t ncam = (int64_t)B * Nc;
    TDS_REQUIRE(ncam <= 2147483647LL, "raster: too many cameras");
    a.ncam = (int32_t)ncam;
    auto launch = [&](auto kernel, int groups, int threads, int queue_items) -> int {
        const size_t group_bytes = (size_t)res * res + (size_t)queue_items * 12 + kGroupExtra;
        const size_t smem = group_bytes * groups;
        TDS_REQUIRE(smem <= 227 * 1024, "raster: res=%d needs %zu bytes of shared memory", res, smem);
        if (smem > 40 * 1024) TDS_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const unsigned grid = (unsigned)((ncam + groups - 1) / groups);
        if (g_ev_start && g_ev_stop) cudaEventRecord(g_ev_start, st);
        kernel<<<grid, threads, smem, st>>>(set, a, pal);
        if (g_ev_start && g_ev_stop) cudaEventRecord(g_ev_stop, st);
        TDS_LAUNCH_OK();
        return TDS_OK;
    };
    if (res == 64) return launch(raster_kernel<32, 64>, 4, 128, QCfg<32>::total);
    if (res < 64) return launch(raster_kernel<32, 0>, 4, 128, QCfg<32>::total);
    if (res <= 128) return launch(raster_kernel<256, 0>, 1, 256, QCfg<256>::total);
    return launch(raster_kernel<512, 0>, 1, 512, QCfg<512>::total);
}
