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
#include <algorithm>
#include <cstdlib>

#include "tds_map.cuh"
#include "tds_raster_tri.h"
#include "tds_raster_rows.h"

namespace {

using tds::kMaxRasterRows;
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
    float kscale;              // -(scale * res / 2), exact for power-of-two res
    float r_in, r_out;         // max-norm radii (pixels) deciding "inside / outside the 1.05x quad" away from its boundary
    const float* edges;        // view-quad edge functions a[4], b[4], c[4] (utils.py:99-122), in shared memory
    int res;
};

__device__ __forceinline__ void make_camera(Camera& cam, float cx, float cy, float S, float C, float scale, int res,
                                            float qx[4], float qy[4], float* edges, bool write_edges) {
    cam.ncx = -cx; cam.ncy = -cy; cam.S = S; cam.C = C; cam.scale = scale; cam.res = res;
    cam.fmin = (float)res;
    cam.half = (float)res / 2.0f;
    cam.kscale = -(scale * cam.half);
    // fp32 disagreement between the pixel-space and the edge-function test is < 1e-4 m; band = 0.01 px + 2e-3 m
    const float band = 0.01f + 0.002f * (scale * cam.half);
    cam.r_in = 0.525f * cam.fmin - band;
    cam.r_out = 0.525f * cam.fmin + band;
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
// centre, i.e. the max-norm ball of radius 0.525 res around the image centre in pixel coordinates.  A
// projected vertex further than ~0.01 px from that boundary is decided from its pixel coordinates; the
// thin band around the boundary falls back to the reference's fp32 edge functions, so the decision is
// always the reference's.
template <bool POW2>
__device__ __forceinline__ void project_f(const Camera& cam, float x, float y, float& u0, float& u1) {
    u0 = cam.C * x + cam.S * y;
    u1 = (-cam.S) * x + cam.C * y;
    if (POW2) {
        // res is a power of two: the multiplications by res and by 1/2 are exact, so they commute with the
        // rounding of (-u) * scale and fold into one constant (bit-identical to the chain below)
        u0 = u0 * cam.kscale + cam.half;
        u1 = u1 * cam.kscale + cam.half;
    } else {
        u0 = (-u0) * cam.scale; u1 = (-u1) * cam.scale;     // rendering/base.py:102-115, operation by operation
        u0 = u0 * cam.fmin;     u1 = u1 * cam.fmin;
        u0 = u0 / 2.0f;         u1 = u1 / 2.0f;
        u0 = u0 + cam.half;     u1 = u1 + cam.half;
    }
}

// kind of a candidate after cull + projection + truncation (rendering/cv2.py:52-56)
enum { kCulled = 0, kVerts = 1, kHuge = 2, kShort = 3, kTall = 4, kClipped = 5 };
#ifndef TDS_SHORT_ROWS
#define TDS_SHORT_ROWS 5
#endif
constexpr int kShortRows = TDS_SHORT_ROWS;      // inside triangles spanning at most this many row steps go to the "short" queue

template <bool POW2>
__device__ __forceinline__ int setup_triangle(const Camera& cam, float x0, float y0, float x1, float y1, float x2,
                                              float y2, int own, int xy[6]) {
    const float px0 = x0 + cam.ncx, py0 = y0 + cam.ncy;
    const float px1 = x1 + cam.ncx, py1 = y1 + cam.ncy;
    const float px2 = x2 + cam.ncx, py2 = y2 + cam.ncy;
    float u0, v0, u1, v1, u2, v2;
    project_f<POW2>(cam, px0, py0, u0, v0);
    project_f<POW2>(cam, px1, py1, u1, v1);
    project_f<POW2>(cam, px2, py2, u2, v2);
    const float d0 = fmaxf(fabsf(u0 - cam.half), fabsf(v0 - cam.half));
    const float d1 = fmaxf(fabsf(u1 - cam.half), fabsf(v1 - cam.half));
    const float d2 = fmaxf(fabsf(u2 - cam.half), fabsf(v2 - cam.half));
    if (fminf(fminf(d0, d1), d2) > cam.r_out) return kCulled;       // all vertices clearly outside (NaN: falls through)
    bool c0 = d0 < cam.r_in, c1 = d1 < cam.r_in, c2 = d2 < cam.r_in;
    const bool und0 = !c0 && !(d0 > cam.r_out), und1 = !c1 && !(d1 > cam.r_out), und2 = !c2 && !(d2 > cam.r_out);
    if (und0 | und1 | und2) {       // some vertex is within the band around the quad boundary (or NaN): exact test
        // one rolled loop over the three vertices keeps this rare path small; the edge functions live in smem
        float ex = px0, ey = py0;
#pragma unroll 1
        for (int k = 0; k < 3; k++) {
            const bool r = inside_quad(cam, ex, ey);
            if (k == 0) { if (und0) c0 = r; ex = px1; ey = py1; }
            else if (k == 1) { if (und1) c1 = r; ex = px2; ey = py2; }
            else { if (und2) c2 = r; }
        }
        if (!(c0 | c1 | c2)) return kCulled;
    }
    const int first = c0 ? 0 : (c1 ? 1 : 2);
    if (!((own >> first) & 1)) return kCulled;    // another cell's copy of this face draws it
    xy[0] = __float2int_rz(u0); xy[1] = __float2int_rz(v0);
    xy[2] = __float2int_rz(u1); xy[3] = __float2int_rz(v1);
    xy[4] = __float2int_rz(u2); xy[5] = __float2int_rz(v2);
    // |coordinates| >= 8000 (or NaN): 64-bit rule
    if (!(fmaxf(fmaxf(d0, d1), d2) < 8000.0f - cam.half)) return kHuge;
    // integer bounding box entirely off the image: clipLine rejects all three edges and the fill returns early
    const int res = cam.res;
    const int xmin = min(min(xy[0], xy[2]), xy[4]), xmax = max(max(xy[0], xy[2]), xy[4]);
    const int ymin = min(min(xy[1], xy[3]), xy[5]), ymax = max(max(xy[1], xy[3]), xy[5]);
    if (((xmax | ymax) < 0) | (xmin >= res) | (ymin >= res)) return kCulled;
    // bounding box within 2x2 pixels: the coverage is the set of in-image vertices (tds_raster_tri.h)
    if (((xmax - xmin) | (ymax - ymin)) <= 1) return kVerts;
    const bool inside = ((xmin | ymin) >= 0) & (xmax < res) & (ymax < res);
    return !inside ? kClipped : (ymax - ymin <= kShortRows ? kShort : kTall);
}

// ---- shared memory is addressed through 32-bit shared-window addresses and explicit ld/st/red.shared: with
// generic pointers the compiler re-derives the window base (S2R SR_CgaCtaId, LEA, IMAD ...) at every access site.
__device__ __forceinline__ uint32_t smem_addr(const void* p) {
    // the volatile move makes the address a plain register value: it is computed once instead of being
    // rematerialised from the special registers wherever it is used
    uint32_t a = (uint32_t)__cvta_generic_to_shared(p), r;
    asm volatile("mov.u32 %0, %1;" : "=r"(r) : "r"(a));
    return r;
}
__device__ __forceinline__ void sred_or(uint32_t a, uint32_t v) { asm volatile("red.shared.or.b32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t slds(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t slds_const(uint32_t a) {      // tables that never change after the prologue
    uint32_t v;
    asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void ssts(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ uint4 slds4(uint32_t a) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void ssts4(uint32_t a, uint4 v) {
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// ---- bitplanes: one bit per pixel and draw rank; plane p, word column w, row y at word  p * res * W32 + w * res + y.
// `plane` below is the shared-window BYTE address of a plane.
__device__ __forceinline__ void or_bit(uint32_t plane, int res, int x, int y) {
    sred_or(plane + 4u * (uint32_t)((x >> 5) * res + y), 1u << (x & 31));
}

__device__ __forceinline__ void or_span(uint32_t plane, int res, int y, int lo, int hi) {
    for (int w = lo >> 5; w <= (hi >> 5); w++) {
        const int l = max(lo - 32 * w, 0), h = min(hi - 32 * w, 31);
        sred_or(plane + 4u * (uint32_t)(w * res + y), (0xffffffffu >> (31 - (h - l))) << l);
    }
}

// 64-pixel rows: the two 32-bit halves of a row mask.  Both are OR-ed unconditionally: ptxas turns a predicated
// red.shared into a branch around it (ISETP, BSSY, BRA, BSYNC), which costs more issue slots than the idle OR of a
// zero (the LSU pipe is 14 % busy, the issue slots 77 %).
__device__ __forceinline__ void or_mask64(uint32_t plane, int y, unsigned long long m) {
    const uint32_t addr = plane + 4u * (uint32_t)y;
    sred_or(addr, (uint32_t)m);
    sred_or(addr + 256u, (uint32_t)(m >> 32));
}

// stage 2a: a triangle with all vertices inside the image: one interval and one atomic OR per row and word
template <int RES, bool SMALL>
__device__ __forceinline__ void draw_inside(uint32_t plane, int res, uint32_t rcp_sa, int x0, int y0, int x1, int y1,
                                            int x2, int y2) {
    tds::FastTri t;
    tds::fast_tri_setup<SMALL>(x0, y0, x1, y1, x2, y2, t, [&](int dy) { return slds_const(rcp_sa + 4u * (uint32_t)dy); });
    tds::fast_tri_rows(t, [&](int y, int lo, int hi) {
        if (RES == 64) or_mask64(plane, y, (~0ull >> (63 - (hi - lo))) << lo);
        else or_span(plane, res, y, lo, hi);
    });
}

// stage 2b: a triangle that crosses the image border (|coordinates| < 8192): clipped outline runs + clamped spans
template <int RES>
__device__ __forceinline__ void draw_clipped(uint32_t plane, int res, uint32_t rcp_sa, int x0, int y0, int x1, int y1,
                                             int x2, int y2) {
    tds::RowTri t;
    tds::row_tri_setup(res, res, x0, y0, x1, y1, x2, y2, t, [&](int dy) { return slds_const(rcp_sa + 4u * (uint32_t)dy); });
#pragma unroll 1
    for (int y = t.ylo; y <= t.yhi; y++) {
        if (RES == 64) {
            unsigned long long m = 0ull;
            tds::row_tri_step(t, res, y, [&](int lo, int hi) { m |= (~0ull >> (63 - (hi - lo))) << lo; });
            or_mask64(plane, y, m);
        } else {
            tds::row_tri_step(t, res, y, [&](int lo, int hi) { or_span(plane, res, y, lo, hi); });
        }
    }
}

// coordinates beyond +-8000 pixels (extreme zoom / giant rectangles): 64-bit rule, pixel by pixel.  Rare.
__device__ __noinline__ void draw_huge(uint32_t plane, int res, int x0, int y0, int x1, int y1, int x2, int y2) {
    tds::draw_triangle(res, res, x0, y0, x1, y1, x2, y2,
        [&](int x, int y) { or_bit(plane, res, x, y); },
        [&](int y, int xa, int xb) { or_span(plane, res, y, xa, xb); });
}

struct RasterArgs {
    const int32_t* env_map;
    const float* cam_xy;
    const float* cam_sc;
    const uint8_t* present;
    const uint8_t* ws;
    float* out;
    int32_t B, Nc, N, T, present_per_camera, res;
    int32_t LR;                // traffic lights + extra rectangles of an environment (T = 3 N + 2 LR)
    int32_t ncam;
    float scale;
    int32_t* next_cam;         // work counter of the persistent grid (zeroed before the launch)
    const float* cam_tris;     // [B*Nc][Tc][6] world-space triangles of each camera (waypoint discs), or NULL
    const int32_t* cam_cls;    // [B*Nc][Tc] their classes (< 0: skipped)
    int32_t Tc;
};

constexpr int kRows = tds::kMaxRasterRows;
constexpr int kQueues = 3;                          // short inside, tall inside, clipped
constexpr int kCullCap = 128;                       // dynamic primitives in view that a camera can list
constexpr int kGroupExtra = kRows * 8 + 16 + 48 + kCullCap * 2;    // row tables, counters, view-quad edge functions, that list

__host__ __device__ constexpr int raster_group_bytes(int res, int n_planes, int G) {
    // planes of the camera | three queues of 64 faces for every warp of the group | tables
    return n_planes * res * ((res + 31) / 32) * 4 + (G / 32) * kQueues * 64 * 16 + kGroupExtra;
}
// The 64x64 / <= 7 classes variant (the benchmark configuration) reserves 7 planes per camera in STATIC shared
// memory: every address is then a compile-time offset and nothing has to be re-derived from the dynamic base.
constexpr int kStaticPlanes = 7;
__host__ __device__ constexpr bool raster_static_smem(int G, int RES, int NS) { return G == 32 && RES == 64 && NS == 3; }

// G = threads cooperating on one camera: 32 (one warp per camera, 4 cameras in flight per CTA, no block barriers)
// for tiles up to 64x64, or the whole CTA for larger tiles.  The warps of a CTA group share the camera's bitplanes
// (atomic ORs) but nothing else: each takes every (G/32)-th batch of 32 candidates and keeps its own three queues,
// so the only block barriers of a camera are after the set-up and before the resolve.
template <int G>
__device__ __forceinline__ void group_sync() {
    if (G == 32) __syncwarp();
    else __syncthreads();
}

#ifndef TDS_RASTER_MINB
#define TDS_RASTER_MINB 7
#endif
// NS = bits of the per-pixel draw rank (0 = background): 3 for up to 7 active classes, 5 for up to 31
template <int G, int RES, int NS, bool SMALL>
#ifndef TDS_RASTER_MINB_BIG
#define TDS_RASTER_MINB_BIG 2
#endif
__global__ void __launch_bounds__(G == 32 ? 128 : G, G <= 128 ? TDS_RASTER_MINB : (G == 256 ? 2 * TDS_RASTER_MINB_BIG : TDS_RASTER_MINB_BIG)) raster_kernel(MapSetDev maps, RasterArgs a, PaletteDev pal) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    constexpr int GROUPS = G == 32 ? 4 : 1;
    constexpr int QN = 64;                      // queue capacity per warp and kind
    // SMALL: images up to 128 pixels, slopes through the reciprocal table
    constexpr bool POW2 = RES != 0 && (RES & (RES - 1)) == 0;     // compile-time tile size that is a power of two
    const int res = RES ? RES : a.res;          // RES = 64 is compiled with constant strides
    const int W32 = RES ? RES / 32 : (res + 31) >> 5;
    const int group = G == 32 ? (threadIdx.x >> 5) : 0;
    const int tid = G == 32 ? (threadIdx.x & 31) : threadIdx.x;
    const int lane = threadIdx.x & 31;
    const int K = pal.n_classes;                // planes = draw ranks of the active classes

    // CTA-wide tables: colour per draw rank (0 = background), reciprocals of the row runs, class -> plane
    __shared__ float4 s_lut[TDS_MAX_CLASSES + 1];
    __shared__ int8_t s_plane_of_class[TDS_MAX_CLASSES];
    constexpr bool STATIC = raster_static_smem(G, RES, NS);
    constexpr int STATIC_RCP = ((64 + 1) * 4 + 15) & ~15;
    constexpr int STATIC_BYTES = STATIC ? STATIC_RCP + 4 * raster_group_bytes(64, kStaticPlanes, 32) : 16;
    __shared__ __align__(16) uint8_t smem_static[STATIC_BYTES];
    uint8_t* const smem = STATIC ? smem_static : smem_raw;
    const int KS = STATIC ? kStaticPlanes : K;                               // planes reserved per camera
    uint32_t* s_rcp = reinterpret_cast<uint32_t*>(smem);                     // [res + 1]
    const int rcp_bytes = ((res + 1) * 4 + 15) & ~15;
    for (int i = threadIdx.x; i <= K; i += blockDim.x) {
        const float* c = pal.rgb[i == 0 ? 0 : pal.order[i - 1] + 1];
        s_lut[i] = make_float4(c[0], c[1], c[2], 0.f);
    }
    for (int i = threadIdx.x; i < TDS_MAX_CLASSES; i += blockDim.x) {
        int p = -1;
        for (int k = 0; k < K; k++) p = pal.order[k] == i ? k : p;
        s_plane_of_class[i] = (int8_t)p;
    }
    for (int i = threadIdx.x; i <= res; i += blockDim.x) s_rcp[i] = tds::row_rcp(i);
    __syncthreads();
    const int plane_tbl = s_plane_of_class[lane];       // lane c holds the plane of class c (TDS_MAX_CLASSES == 32)

    // per-group shared memory: planes | queues | row tables
    const int plane_words = res * W32;
    const int group_bytes = raster_group_bytes(res, KS, G);
    uint8_t* base = smem + rcp_bytes + group * group_bytes;
    const uint32_t planes_sa = smem_addr(base);                            // [KS][W32][res] words
    const uint32_t rcp_sa = planes_sa - (uint32_t)(rcp_bytes + group * group_bytes);
    const uint32_t plane_bytes = 4u * (uint32_t)plane_words;
    constexpr int WARPS = G / 32;
    const uint32_t queues_sa = planes_sa + (uint32_t)KS * plane_bytes;     // [WARPS][kQueues][QN] x 16 B
    const uint32_t queue_sa = queues_sa + (G == 32 ? 0u : (uint32_t)(threadIdx.x >> 5) * (kQueues * QN * 16));   // this warp's
    const uint32_t start_sa = queues_sa + WARPS * kQueues * QN * 16;       // [kRows] first record of a grid row
    const uint32_t count_sa = start_sa + kRows * 4;                        // [kRows] records of a grid row
    int* s_cnt = reinterpret_cast<int*>(base + KS * plane_words * 4 + WARPS * kQueues * QN * 16 + kRows * 8);   // [3] = next camera (G > 32)
    float* s_edges = reinterpret_cast<float*>(s_cnt + 4);                   // [12]
    uint16_t* s_list = reinterpret_cast<uint16_t*>(s_edges + 12);          // [kCullCap] dynamic primitives in view

    // persistent grid: every group (warp or CTA) pulls the next camera from a global counter, so uneven cameras
    // (a junction full of lane markings next to an empty field) do not leave SMs idle at the end
    while (true) {
        int camid = 0;
        if (G == 32) {
            if (lane == 0) camid = atomicAdd(a.next_cam, 1);
            camid = __shfl_sync(0xffffffffu, camid, 0);
        } else {
            __syncthreads();                        // previous camera completely done (s_cnt is reused below)
            if (tid == 0) s_cnt[3] = atomicAdd(a.next_cam, 1);
            __syncthreads();
            camid = s_cnt[3];
        }
        if (camid >= a.ncam) break;
        const int b = camid / a.Nc;
        const MapDev& map = maps.m[a.env_map ? a.env_map[b] : 0];
        Camera cam;
        float qx[4], qy[4];
        const float2 cxy = reinterpret_cast<const float2*>(a.cam_xy)[camid];
        const float2 csc = reinterpret_cast<const float2*>(a.cam_sc)[camid];
        group_sync<G>();                            // previous camera of this group is completely done
        make_camera(cam, cxy.x, cxy.y, csc.x, csc.y, a.scale, res, qx, qy, s_edges, tid == 0);
        {
            for (int i = tid; i < K * plane_words / 4; i += G) ssts4(planes_sa + 16u * (uint32_t)i, make_uint4(0u, 0u, 0u, 0u));
            if (G != 32 && tid < 3) s_cnt[tid] = 0;
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
        // record ranges of the touched grid rows: the columns of the quad's bounding box (with 16 m cells the quad
        // touches 3-5 rows and columns; candidates of the few extra corner cells are rejected after ~40
        // instructions, which is cheaper than intersecting the quad with every row - and much less code)
        const float xmin = fminf(fminf(wqx[0], wqx[1]), fminf(wqx[2], wqx[3])), xmax = fmaxf(fmaxf(wqx[0], wqx[1]), fmaxf(wqx[2], wqx[3]));
        for (int ri = tid; ri < nrows; ri += G) {
            const int r = r0 + ri;
            int st = 0, cnt = 0;
            if (xmin <= xmax) {
                const int c0 = max((int)floorf((xmin - margin - map.rx0) * map.rinv), 0);
                const int c1 = min((int)floorf((xmax + margin - map.rx0) * map.rinv), map.rgx - 1);
                if (c1 >= c0) {
                    st = map.rcell[r * map.rgx + c0];
                    cnt = map.rcell[r * map.rgx + c1 + 1] - st;
                }
            }
            ssts(start_sa + 4u * ri, (uint32_t)st);
            ssts(count_sa + 4u * ri, (uint32_t)cnt);
        }
        group_sync<G>();

        const int T = a.T;
        const float* dtri = reinterpret_cast<const float*>(a.ws + (int64_t)b * ws_env_bytes(T));
        const uint8_t* dcls = reinterpret_cast<const uint8_t*>(dtri + (int64_t)T * 6);
        const uint8_t* pres = a.present ? (a.present_per_camera ? a.present + (int64_t)camid * a.N : a.present + (int64_t)b * a.N)
                                        : nullptr;

        // ---- dynamic primitives in view.  An agent (rectangle + direction triangle, 3 faces) or a traffic light /
        // sign (2 faces) lies within the circle around its rectangle's diagonal (the corners 1 and 3 of its first
        // face); if that circle misses the quad none of its vertices is inside and all its faces are culled, so
        // only the primitives that pass are listed (in any order: the bitplanes do not depend on it).  Absent
        // agents are not listed: they all draw the SAME degenerate face (mesh.py:1083-1089), added once below.
        const int items = a.N + a.LR;
        int n_view = 0;
        bool any_absent = false;
        for (int i0 = 0; i0 < items; i0 += G) {
            const int i = i0 + tid;
            bool keep = false;
            if (i < items) {
                const bool agent = i < a.N;
                if (agent && pres && !pres[i]) {
                    any_absent = true;
                } else {
                    const float* p = dtri + (int64_t)(agent ? 3 * i : 3 * a.N + 2 * (i - a.N)) * 6;
                    const float mx = 0.5f * (p[2] + p[4]), my = 0.5f * (p[3] + p[5]);
                    const float hx = p[2] - mx, hy = p[3] - my;
                    float u, v;
                    project_f<POW2>(cam, mx + cam.ncx, my + cam.ncy, u, v);
                    const float dd = fmaxf(fabsf(u - cam.half), fabsf(v - cam.half));
                    keep = !(dd > (cam.r_out + 0.05f) + sqrtf(hx * hx + hy * hy) * (a.scale * cam.half) * 1.0001f);   // NaN: kept
                }
            }
            const unsigned m = __ballot_sync(0xffffffffu, keep);
            int base = n_view;
            if (G == 32) {
                n_view += __popc(m);
            } else {
                if (lane == 0 && m) base = atomicAdd(&s_cnt[0], __popc(m));
                base = __shfl_sync(0xffffffffu, base, 0);
            }
            const int pos = base + __popc(m & ((1u << lane) - 1u));
            if (keep && pos < kCullCap) s_list[pos] = (uint16_t)i;
        }
        any_absent = __any_sync(0xffffffffu, any_absent);
        if (G != 32) {
            if (any_absent && lane == 0) s_cnt[1] = 1;
            __syncthreads();
            n_view = s_cnt[0];
            any_absent = s_cnt[1] != 0;
        } else {
            __syncwarp();
        }
        // more primitives in view than the list holds (or more than 65535 of them): take them all, in order
        const bool listed = n_view <= kCullCap && items <= 65535;
        const int n_dyn = 3 * (listed ? n_view : items);
        const int dyn_count = n_dyn + a.Tc + (any_absent ? 1 : 0);

        // ---- ONE pass over the candidates: segment r < nrows = the record range of grid row r, segment nrows =
        // the dynamic primitives of the environment.  Stage 1 (cull + project + truncate) plots the faces that are
        // just their vertices and queues the others by kind; whenever a warp has queued 32 faces of a kind, stage 2 turns
        // them into row intervals, one face per thread, so stage 2 always runs with full warps.  The last
        // iteration (seg > nrows) only drains the queues.
        int seg = 0, j0 = 0, seg_start = 0, seg_count = nrows > 0 ? (int)slds(count_sa) : dyn_count;
        if (nrows > 0) seg_start = (int)slds(start_sa);
        int nq0 = 0, nq1 = 0, nq2 = 0;                 // fill levels of this warp's queues (uniform over the warp)
        while (true) {
            while (j0 >= seg_count && seg <= nrows) {
                seg++;
                j0 = 0;
                seg_count = seg < nrows ? (int)slds(count_sa + 4u * seg) : (seg == nrows ? dyn_count : 0);
                seg_start = seg < nrows ? (int)slds(start_sa + 4u * seg) : 0;
            }
            const bool drain = seg > nrows;
            if (!drain) {
                // uniform control flow: every lane fetches a valid record (the last one of the segment past its end)
                const int j = j0 + tid;
                const bool warp_has_work = j0 + (tid & ~31) < seg_count;
                j0 += G;
                if (G != 32 && !warp_has_work) continue;      // this warp's slice of the batch is past the segment's end
                const bool valid = j < seg_count;
                const int jj = valid ? j : seg_count - 1;
                float x0, y0, x1, y1, x2, y2;
                int own = 7, cls;
                if (seg < nrows) {
                    const float4* rp = map.rec + 2 * (int64_t)(seg_start + jj);
                    const float4 v01 = __ldg(rp);
                    const float4 v2o = __ldg(rp + 1);
                    x0 = v01.x; y0 = v01.y; x1 = v01.z; y1 = v01.w; x2 = v2o.x; y2 = v2o.y;
                    const int meta = __float_as_int(v2o.z);
                    own = meta & 7;
                    cls = meta >> 8;
                } else if (jj >= n_dyn && jj < n_dyn + a.Tc) {
                    // triangles of this camera only (goal-waypoint discs, mesh.py:1120-1145)
                    const int64_t ct = (int64_t)camid * a.Tc + (jj - n_dyn);
                    const int c = a.cam_cls[ct];
                    cls = c < 0 ? 255 : c;
                    const float* p = a.cam_tris + ct * 6;
                    x0 = p[0]; y0 = p[1]; x1 = p[2]; y1 = p[3]; x2 = p[4]; y2 = p[5];
                } else {
                    // dynamic primitives in view: slot 3 k + f = face f of the k-th listed agent / light / sign; the
                    // last slot is the degenerate face of the absent agents: actor vertex 0 with agent 0's class
                    const bool degenerate = jj >= n_dyn;
                    const int k = jj / 3, f = jj - 3 * k;
                    const int item = degenerate ? 0 : (listed ? (int)s_list[k] : k);
                    const bool agent = item < a.N;
                    const int t = degenerate ? 0 : (agent ? 3 * item + f : 3 * a.N + 2 * (item - a.N) + f);
                    // a sign has two faces; without the list an absent agent shows up here as well
                    const bool skip = !degenerate && ((!agent && f == 2) || (!listed && agent && pres && !pres[item]));
                    const int ts = skip ? 0 : t;
                    cls = skip ? 255 : dcls[ts];
                    const float* p = dtri + (int64_t)ts * 6;
                    x0 = p[0]; y0 = p[1];
                    x1 = degenerate ? x0 : p[2]; y1 = degenerate ? y0 : p[3];
                    x2 = degenerate ? x0 : p[4]; y2 = degenerate ? y0 : p[5];
                }
                // class -> plane through the lane-resident table (lane c holds the plane of class c)
                int plane = __shfl_sync(0xffffffffu, plane_tbl, cls & 31);
                plane = (valid && (unsigned)cls < (unsigned)TDS_MAX_CLASSES) ? plane : -1;
                int kind = kCulled;
                int xy[6];
                if (plane >= 0) kind = setup_triangle<POW2>(cam, x0, y0, x1, y1, x2, y2, own, xy);
                if (kind == kVerts) {
                    // a vertex outside the image ORs a zero into pixel (0, 0): no branch around the reduction
                    const uint32_t pl = planes_sa + (uint32_t)plane * plane_bytes;
#pragma unroll
                    for (int k = 0; k < 3; k++) {
                        const bool in = (unsigned)xy[2 * k] < (unsigned)res && (unsigned)xy[2 * k + 1] < (unsigned)res;
                        const int x = in ? xy[2 * k] : 0, y = in ? xy[2 * k + 1] : 0;
                        sred_or(pl + 4u * (uint32_t)((x >> 5) * res + y), in ? 1u << (x & 31) : 0u);
                    }
                } else if (kind == kHuge) {
                    draw_huge(planes_sa + (uint32_t)plane * plane_bytes, res, xy[0], xy[1], xy[2], xy[3], xy[4], xy[5]);
                }
                // queue the other faces by kind (warp-aggregated append)
                const unsigned m0 = __ballot_sync(0xffffffffu, kind == kShort), m1 = __ballot_sync(0xffffffffu, kind == kTall),
                               m2 = __ballot_sync(0xffffffffu, kind == kClipped);
                if (m0 | m1 | m2) {
                    const int b0 = nq0, b1 = nq1, b2 = nq2;
                    nq0 += __popc(m0); nq1 += __popc(m1); nq2 += __popc(m2);
                    if (kind >= kShort) {
                        const unsigned mine = kind == kShort ? m0 : (kind == kTall ? m1 : m2);
                        const int qb = kind == kShort ? b0 : (kind == kTall ? QN + b1 : 2 * QN + b2);
                        const int pos = qb + __popc(mine & ((1u << lane) - 1));
                        ssts4(queue_sa + 16u * (uint32_t)pos,
                              make_uint4((uint32_t)(xy[0] & 0xffff) | ((uint32_t)xy[1] << 16),
                                         (uint32_t)(xy[2] & 0xffff) | ((uint32_t)xy[3] << 16),
                                         (uint32_t)(xy[4] & 0xffff) | ((uint32_t)xy[5] << 16), (uint32_t)plane));
                    }
                }
                __syncwarp();
            }
            if (!(drain | (nq0 >= 32) | (nq1 >= 32) | (nq2 >= 32))) continue;
            // stage 2: a queue is drawn when it holds a full group (or, at the end, whatever is left)
#pragma unroll 1
            for (int which = 0; which < 2; which++) {
                const int nq = which ? nq1 : nq0;
                if (nq >= 32 || (drain && nq > 0)) {
                    const int take = min(nq, 32);
                    if (lane < take) {
                        const uint4 q = slds4(queue_sa + 16u * (uint32_t)(which * QN + nq - take + lane));
                        draw_inside<RES, SMALL>(planes_sa + q.w * plane_bytes, res, rcp_sa,
                                                (int16_t)(q.x & 0xffff), (int32_t)q.x >> 16, (int16_t)(q.y & 0xffff),
                                                (int32_t)q.y >> 16, (int16_t)(q.z & 0xffff), (int32_t)q.z >> 16);
                    }
                    if (which) nq1 -= take; else nq0 -= take;
                    __syncwarp();
                }
            }
            if (nq2 >= 32 || (drain && nq2 > 0)) {
                const int take = min(nq2, 32);
                if (lane < take) {
                    const uint4 q = slds4(queue_sa + 16u * (uint32_t)(2 * QN + nq2 - take + lane));
                    draw_clipped<RES>(planes_sa + q.w * plane_bytes, res, rcp_sa,
                                      (int16_t)(q.x & 0xffff), (int32_t)q.x >> 16, (int16_t)(q.y & 0xffff),
                                      (int32_t)q.y >> 16, (int16_t)(q.z & 0xffff), (int32_t)q.z >> 16);
                }
                nq2 -= take;
                __syncwarp();
            }
            if (drain) {
                if ((nq0 | nq1 | nq2) == 0) break;     // a queue held more than one group: drain again
            }
        }
        group_sync<G>();

        // ---- resolve painter's order and expand through the colour LUT: out[cam][ch][x][y].
        // A thread owns 4 image rows y of one 32-pixel word column: it folds the K planes of those rows into bit
        // slices of the top-most draw rank (registers), then walks the 32 columns, one 128-bit store per channel.
        float* outc = a.out + (int64_t)camid * 3 * res * res;
        const int nyq = res >> 2;
        for (int item = tid; item < W32 * nyq; item += G) {
            const int w = item / nyq, yq = item - w * nyq;
            uint32_t sl[NS][4];
#pragma unroll
            for (int s = 0; s < NS; s++)
#pragma unroll
                for (int k = 0; k < 4; k++) sl[s][k] = 0u;
            uint32_t rem[4] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu};
#pragma unroll 1
            for (int p = K - 1; p >= 0; p--) {          // last drawn = on top
                const uint4 v = slds4(planes_sa + (uint32_t)p * plane_bytes + 4u * (uint32_t)(w * res + 4 * yq));
                const uint32_t pw[4] = {v.x, v.y, v.z, v.w};
                const int id = p + 1;
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const uint32_t e = pw[k] & rem[k];
                    rem[k] &= ~e;
#pragma unroll
                    for (int s = 0; s < NS; s++) sl[s][k] |= ((id >> s) & 1) ? e : 0u;
                }
            }
            const int nx = min(32, res - 32 * w);
            float* o = outc + (int64_t)(32 * w) * res + 4 * yq;
            const int plane_stride = res * res;
#pragma unroll 1
            for (int x = 0; x < nx; x++) {
                float4 c[4];
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    uint32_t v = sl[0][k] & 1u;
#pragma unroll
                    for (int s = 1; s < NS; s++) v |= (sl[s][k] & 1u) << s;
#pragma unroll
                    for (int s = 0; s < NS; s++) sl[s][k] >>= 1;
                    c[k] = s_lut[v];
                }
                tds::st_cs_f4(reinterpret_cast<float4*>(o), make_float4(c[0].x, c[1].x, c[2].x, c[3].x));
                tds::st_cs_f4(reinterpret_cast<float4*>(o + plane_stride), make_float4(c[0].y, c[1].y, c[2].y, c[3].y));
                tds::st_cs_f4(reinterpret_cast<float4*>(o + 2 * plane_stride), make_float4(c[0].z, c[1].z, c[2].z, c[3].z));
                o += res;
            }
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
                                   const float* d_cam_tris, const int32_t* d_cam_tri_class, int32_t Tc,
                                   const tds_palette_t* palette, float scale, int32_t res,
                                   float* d_out, void* d_workspace, void* stream) {
    TDS_REQUIRE(B >= 0 && Nc >= 0 && N >= 0 && L >= 0 && R >= 0, "raster: negative size");
    if (B == 0 || Nc == 0) return TDS_OK;
    TDS_REQUIRE(d_cam_xy && d_cam_sc && d_out && palette, "raster: null pointer");
    TDS_REQUIRE(N == 0 || (d_agent_state && d_agent_size), "raster: null agent tensors");
    TDS_REQUIRE(L == 0 || (d_tl_corners && d_tl_state), "raster: null traffic light tensors");
    TDS_REQUIRE(R == 0 || (d_rect_corners && d_rect_class), "raster: null rectangle tensors");
    TDS_REQUIRE(Tc >= 0 && (Tc == 0 || (d_cam_tris && d_cam_tri_class)), "raster: null per-camera triangle tensors");
    TDS_REQUIRE(res >= 4 && res % 4 == 0 && res <= 448, "raster: res=%d must be a multiple of 4 in [4,448]", res);
    TDS_REQUIRE(scale > 0.0f, "raster: scale must be positive");
    TDS_REQUIRE(palette->n_classes >= 0 && palette->n_classes <= TDS_MAX_CLASSES, "raster: bad palette");
    const int T = 3 * N + 2 * L + 2 * R;
    TDS_REQUIRE(d_workspace, "raster: null workspace (tds_raster_workspace_bytes)");
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
    a.B = B; a.Nc = Nc; a.N = N; a.LR = L + R; a.T = T; a.present_per_camera = present_per_camera; a.res = res; a.scale = scale;
    const int64_t ncam = (int64_t)B * Nc;
    TDS_REQUIRE(ncam <= 2147483647LL, "raster: too many cameras");
    a.ncam = (int32_t)ncam;
    a.cam_tris = d_cam_tris; a.cam_cls = d_cam_tri_class; a.Tc = Tc;
    a.next_cam = reinterpret_cast<int32_t*>((uint8_t*)d_workspace + (int64_t)B * ws_env_bytes(T));
    const int K = pal.n_classes;
    TDS_REQUIRE(K <= 31, "raster: at most 31 active classes (got %d)", K);
    const int sms = tds::sm_count();
    auto launch = [&](auto kernel, int groups, int threads, bool static_smem) -> int {
        const size_t rcp_bytes = (((size_t)res + 1) * 4 + 15) & ~(size_t)15;
        const size_t smem = static_smem ? 0 : rcp_bytes + (size_t)raster_group_bytes(res, K, threads / groups) * groups;
        TDS_REQUIRE(smem <= 227 * 1024, "raster: res=%d with %d active classes needs %zu bytes of shared memory", res, K, smem);
        if (smem > 40 * 1024) TDS_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int per_sm = 0;
        TDS_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem));
        TDS_REQUIRE(per_sm >= 1, "raster: kernel does not fit an SM (res=%d, %d classes)", res, K);
        // persistent grid: every resident CTA slot of the GPU, cameras dealt round-robin
        const int64_t want = (ncam + groups - 1) / groups;
        const unsigned grid = (unsigned)std::min<int64_t>(want, (int64_t)sms * per_sm);
        TDS_CUDA_OK(cudaMemsetAsync(a.next_cam, 0, sizeof(int32_t), st));
        if (g_ev_start && g_ev_stop) cudaEventRecord(g_ev_start, st);
        kernel<<<grid, threads, smem, st>>>(set, a, pal);
        if (g_ev_start && g_ev_stop) cudaEventRecord(g_ev_stop, st);
        TDS_LAUNCH_OK();
        return TDS_OK;
    };
    // Threads per camera.  Tiles up to 96x96: a warp per camera, 4 cameras in flight per CTA (at 128x128 only 16 such
    // warps fit an SM: 4 sets of bitplanes per CTA).  Above: a CTA of 4, 8 or 16 independent warps per camera - the
    // fewest warps that still leave 16 warps resident on an SM, because a warp that sees a smaller share of the
    // camera's faces fills its queues of 32 less often (256x256: 2.34 / 2.56 / 3.02 ms with 4 / 8 / 16 warps).
    // TDS_RASTER_G = 32 / 128 / 256 / 512 forces a variant (profiling aid).
    const size_t rcp_smem = (((size_t)res + 1) * 4 + 15) & ~(size_t)15;
    const size_t warp_smem = rcp_smem + (size_t)raster_group_bytes(res, K, 32) * 4;
    int G = 32;
    if (res > 96 || warp_smem > 227 * 1024) {
        int best_warps = -1;
        const int cand[3] = {128, 256, 512}, reg_cap[3] = {TDS_RASTER_MINB, 2 * TDS_RASTER_MINB_BIG, TDS_RASTER_MINB_BIG};
        for (int i = 0; i < 3; i++) {
            const size_t smem_g = rcp_smem + (size_t)raster_group_bytes(res, K, cand[i]) + 1024;
            const int warps = std::min<int>(reg_cap[i], (int)(233472 / smem_g)) * (cand[i] / 32);
            if (warps > best_warps) { best_warps = warps; G = cand[i]; }
            if (warps >= 16) break;
        }
    }
    if (const char* e = getenv("TDS_RASTER_G")) {
        const int g = atoi(e);
        if (g == 32 || g == 128 || g == 256 || g == 512) G = g;
    }
    if (G == 32 && (warp_smem > 227 * 1024 || res > 128)) G = 128;
    const bool small = res <= 128;
    if (K <= 7) {
        if (G == 32 && res == 64) return launch(raster_kernel<32, 64, 3, true>, 4, 128, raster_static_smem(32, 64, 3));
        if (G == 32) return launch(raster_kernel<32, 0, 3, true>, 4, 128, false);
        if (G == 128 && res == 128) return launch(raster_kernel<128, 128, 3, true>, 1, 128, false);     // configs 3 and 4: constant strides
        if (G == 128 && res == 256) return launch(raster_kernel<128, 256, 3, false>, 1, 128, false);
        if (G == 128) return small ? launch(raster_kernel<128, 0, 3, true>, 1, 128, false) : launch(raster_kernel<128, 0, 3, false>, 1, 128, false);
        if (G == 256) return small ? launch(raster_kernel<256, 0, 3, true>, 1, 256, false) : launch(raster_kernel<256, 0, 3, false>, 1, 256, false);
        return launch(raster_kernel<512, 0, 3, false>, 1, 512, false);
    }
    if (G == 32 && res == 64) return launch(raster_kernel<32, 64, 5, true>, 4, 128, raster_static_smem(32, 64, 5));
    if (G == 32) return launch(raster_kernel<32, 0, 5, true>, 4, 128, false);
    if (G == 128) return small ? launch(raster_kernel<128, 0, 5, true>, 1, 128, false) : launch(raster_kernel<128, 0, 5, false>, 1, 128, false);
    if (G == 256) return small ? launch(raster_kernel<256, 0, 5, true>, 1, 256, false) : launch(raster_kernel<256, 0, 5, false>, 1, 256, false);
    return launch(raster_kernel<512, 0, 5, false>, 1, 512, false);
}
