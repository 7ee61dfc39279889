// Offroad infraction: squared distance from the four box corners to the nearest map face.
//
// Reference: offroad_infraction_loss (torchdrivesim/infractions.py:176-229, pure-torch branch) and
// point_to_mesh_distance_pt (infractions.py:86-173), which expand the mesh to a
// [B*A*4, F, 3, 3] tensor (290 TB at 1024 x 64 agents on Town01).  Here one thread owns one
// corner and walks the per-map uniform grid in expanding rings; the per-face arithmetic is the
// reference's, operation for operation (fp32, no FMA), so the minimum over the visited faces is
// the reference's minimum over all faces.  The grid (~1 MB per map) is L2 resident; HBM traffic is
// the 28 B per agent of state/size in and 4 B out.
#include <algorithm>
#include <math_constants.h>

#include "tds_map.cuh"

namespace {

using tds::MapDev;
using tds::MapSetDev;

struct EdgeHit {
    float d2, qx, qy;   // squared distance and closest point
};

__device__ __forceinline__ EdgeHit edge_dist2(float px, float py, float ax, float ay, float bx, float by) {
    const float abx = bx - ax, aby = by - ay;
    const float l2 = abx * abx + aby * aby;
    const float t = (abx * (px - ax) + aby * (py - ay)) / (l2 + 1e-8f);
    const float tt = fminf(fmaxf(t, 0.f), 1.f);
    EdgeHit h;
    h.qx = ax + tt * abx;
    h.qy = ay + tt * aby;
    if (l2 <= 1e-8f) { h.qx = bx; h.qy = by; }     // degenerate edge: distance to v1 (infractions.py:156-158)
    const float dx = px - h.qx, dy = py - h.qy;
    h.d2 = dx * dx + dy * dy;
    return h;
}

// infractions.py:100-169 for z = 0: 0 inside a non-degenerate triangle, else min edge distance
__device__ __forceinline__ float point_tri_dist2(float px, float py, float x0, float y0, float x1, float y1, float x2,
                                                 float y2) {
    const float ax = x2 - x0, ay = y2 - y0, bx = x1 - x0, by = y1 - y0;
    const float cz = ax * by - ay * bx;
    const float p2x = px - x0, p2y = py - y0;
    const float d00 = bx * bx + by * by;
    const float d01 = bx * ax + by * ay;
    const float d11 = ax * ax + ay * ay;
    const float d20 = p2x * bx + p2y * by;
    const float d21 = p2x * ax + p2y * ay;
    const float denom = d00 * d11 - d01 * d01 + 1e-8f;
    const float w1 = (d11 * d20 - d01 * d21) / denom;
    const float w2 = (d00 * d21 - d01 * d20) / denom;
    const float w0 = 1.0f - w1 - w2;
    const bool inside = (0.f <= w0) & (w0 <= 1.f) & (0.f <= w1) & (w1 <= 1.f) & (0.f <= w2) & (w2 <= 1.f);
    const float area = fabsf(bx * ay - by * ax) / 2.0f;
    const bool cond = inside & !(area < 5e-3f) & (fabsf(cz) > 1e-8f);
    if (cond) return 0.0f;
    const float e01 = edge_dist2(px, py, x0, y0, x1, y1).d2;
    const float e02 = edge_dist2(px, py, x0, y0, x2, y2).d2;
    const float e12 = edge_dist2(px, py, x1, y1, x2, y2).d2;
    return fminf(fminf(e01, e02), e12);
}

// squared distance from p to an axis-aligned box (0 inside)
__device__ __forceinline__ float box_dist2(float px, float py, float x0, float y0, float x1, float y1) {
    const float dx = fmaxf(fmaxf(x0 - px, px - x1), 0.0f), dy = fmaxf(fmaxf(y0 - py, py - y1), 0.0f);
    return dx * dx + dy * dy;
}

// A face (or cell) whose bounding box is further than the current best cannot lower the minimum; the
// 1e-5 relative slack covers the rounding of the reference's own distance formula.
__device__ __forceinline__ bool cannot_improve(float lower_bound2, float best) {
    return lower_bound2 > best * 1.00001f + 1e-9f;
}

// ---- nearest face of ONE point by ONE warp.  The faces of the visited cells are spread over the 32 lanes (cells are
// pruned by their box first, then a prefix sum of the cells' face counts maps a flat face index back to its cell), the
// lanes share their minimum after every round of 32 faces, so the bounding-box pruning of the next round uses it.
__device__ __forceinline__ void warp_min(float& best, int& bf) {
    // squared distances are >= 0 (or +inf) and never NaN here, so they order like their bit patterns: two warp
    // reductions (redux.sync) give the minimum and the lowest face index that attains it (-1 = none, the largest)
    const unsigned mn = __reduce_min_sync(0xffffffffu, __float_as_uint(best));
    const unsigned f = __reduce_min_sync(0xffffffffu, __float_as_uint(best) == mn ? (unsigned)bf : 0xffffffffu);
    best = __uint_as_float(mn);
    bf = (int)f;
}

// lane l proposes cell (cx, cy) (valid or not); all their faces are evaluated, 32 per round.  SINGLE: every lane
// proposes the same cell (the point's own), so a flat face index is an offset into that cell.
template <bool SINGLE>
__device__ __forceinline__ void visit_cells(const MapDev& m, bool valid, int cx, int cy, float px, float py, float stop,
                                            float& best, int& bf, int lane) {
    int e0 = 0, cnt = 0;
    if (valid) {
        // cell bounds (faces are binned with 1 mm of slack, see map.cu)
        const float x0 = m.ox0 + (float)cx * m.ocs, y0 = m.oy0 + (float)cy * m.ocs;
        if (!cannot_improve(box_dist2(px, py, x0 - 2e-3f, y0 - 2e-3f, x0 + m.ocs + 2e-3f, y0 + m.ocs + 2e-3f), best)) {
            const int c = cy * m.ogx + cx;
            e0 = __ldg(m.ocell + c);
            cnt = __ldg(m.ocell + c + 1) - e0;
        }
    }
    int incl = cnt;
    if (!SINGLE) {
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
    }
    const int excl = incl - cnt;
    const int total = SINGLE ? cnt : __shfl_sync(0xffffffffu, incl, 31);
    for (int f0 = 0; f0 < total; f0 += 32) {              // uniform over the warp
        const int f = f0 + lane;
        int e = e0 + f;
        if (!SINGLE) {
            // owner of flat face f = the last lane whose exclusive prefix is <= f
            int lo = 0;
#pragma unroll
            for (int step = 16; step > 0; step >>= 1) {
                const int v = __shfl_sync(0xffffffffu, excl, lo + step);     // lo + step <= 31
                if (v <= f) lo += step;
            }
            e = __shfl_sync(0xffffffffu, e0, lo) + (f - __shfl_sync(0xffffffffu, excl, lo));
        }
        if (f < total) {
            const float4 a = __ldg(m.orec + 2 * (int64_t)e), b = __ldg(m.orec + 2 * (int64_t)e + 1);
            const float bx0 = fminf(fminf(a.x, a.z), b.x), bx1 = fmaxf(fmaxf(a.x, a.z), b.x);
            const float by0 = fminf(fminf(a.y, a.w), b.y), by1 = fmaxf(fmaxf(a.y, a.w), b.y);
            if (!cannot_improve(box_dist2(px, py, bx0, by0, bx1, by1), best)) {
                const float d = point_tri_dist2(px, py, a.x, a.y, a.z, a.w, b.x, b.y);
                const int fi = __float_as_int(b.z);
                if (d < best || (d == best && fi < bf)) { best = d; bf = fi; }
            }
        }
        warp_min(best, bf);
        if (best <= stop) return;
    }
}

// min over all faces of the map; returns the squared distance, *face = argmin (lowest face index among equal
// distances).  The search ends as soon as a face within `stop` (the threshold of the loss, >= 0) is found: the loss
// of this corner and its gradient are then 0 whatever the true minimum is (F.threshold, infractions.py:172), and the
// returned value / face are those of that face.  All 32 lanes of the warp search for the same point and return the
// same result.
__device__ float nearest_face(const MapDev& m, float px, float py, float stop, int* face, int lane) {
    float best = CUDART_INF_F;
    int bf = -1;
    if (m.nf == 0) { *face = -1; return 0.0f; }
    const float fx = floorf((px - m.ox0) * m.oinv), fy = floorf((py - m.oy0) * m.oinv);
    const int cx = (int)fminf(fmaxf(fx, -1.0e6f), 1.0e6f), cy = (int)fminf(fmaxf(fy, -1.0e6f), 1.0e6f);
    const int kmax = max(max(cx, m.ogx - 1 - cx), max(cy, m.ogy - 1 - cy));
    // first ring that touches the grid at all
    int k0 = max(max(-cx, cx - (m.ogx - 1)), max(-cy, cy - (m.ogy - 1)));
    k0 = max(k0, 0);
    // faces are binned by bounding box, so after ring k every unvisited face lies outside the square of visited
    // cells: further away than k cells plus the distance from p to the border of its own cell
    const float ux = px - (m.ox0 + (float)cx * m.ocs), uy = py - (m.oy0 + (float)cy * m.ocs);
    const float inner = fminf(fmaxf(fminf(fminf(ux, m.ocs - ux), fminf(uy, m.ocs - uy)), 0.0f), m.ocs);
    for (int k = k0; k <= kmax; k++) {
        if (k == 0) {
            visit_cells<true>(m, true, cx, cy, px, py, stop, best, bf, lane);
        } else {
            // the in-grid cells of ring k: top row, bottom row, left column, right column
            const int xa = max(cx - k, 0), xb = min(cx + k, m.ogx - 1);
            const int ya = max(cy - k + 1, 0), yb = min(cy + k - 1, m.ogy - 1);
            const int nx = max(xb - xa + 1, 0), ny = max(yb - ya + 1, 0);
            const int nt = cy - k >= 0 ? nx : 0, nb = cy + k < m.ogy ? nx : 0;
            const int nl = cx - k >= 0 ? ny : 0, nr = cx + k < m.ogx ? ny : 0;
            const int n = nt + nb + nl + nr;
            for (int i0 = 0; i0 < n; i0 += 32) {
                const int i = i0 + lane;
                int x, y;
                if (i < nt) { x = xa + i; y = cy - k; }
                else if (i < nt + nb) { x = xa + (i - nt); y = cy + k; }
                else if (i < nt + nb + nl) { x = cx - k; y = ya + (i - nt - nb); }
                else { x = cx + k; y = ya + (i - nt - nb - nl); }
                visit_cells<false>(m, i < n, x, y, px, py, stop, best, bf, lane);
                if (best <= stop) break;
            }
        }
        // (1 mm slack for the cell assignment of p itself)
        if (best <= stop) break;
        const float reach = fmaxf((float)k * m.ocs + inner - 1e-3f, 0.0f);
        if (best <= reach * reach) break;
    }
    if (!(best < CUDART_INF_F)) best = 0.0f;   // NaN position / distance: nan_to_num, infractions.py:171
    *face = bf;
    return best;
}

__device__ __forceinline__ void corner_of(const float* st, const float* lw, int k, float& px, float& py, float& s, float& c) {
    // box2corners_th, _iou_utils.py:270-299: (+,+), (-,+), (-,-), (+,-) rotated by psi
    const float sx = (k == 0 || k == 3) ? 0.5f : -0.5f;
    const float sy = (k < 2) ? 0.5f : -0.5f;
    tds::sincos_cr(st[2], s, c);
    const float ux = sx * lw[0], uy = sy * lw[1];
    px = (ux * c + uy * (-s)) + st[0];
    py = (ux * s + uy * c) + st[1];
}

// one CTA of 4 warps per agent, one warp per box corner
#ifndef TDS_OFFROAD_MINB
#define TDS_OFFROAD_MINB 12
#endif
__global__ void __launch_bounds__(128, TDS_OFFROAD_MINB) offroad_fwd_kernel(MapSetDev maps, const int32_t* __restrict__ env_map,
                                                          const float* __restrict__ state, const float* __restrict__ lenwid,
                                                          const uint8_t* __restrict__ present, int B, int A, float thr,
                                                          float* __restrict__ out, int32_t* __restrict__ face) {
    __shared__ float s_v[4];
    __shared__ float2 s_corner[4];
    const int lane = threadIdx.x & 31, k = threadIdx.x >> 5;
    const unsigned n_agents = (unsigned)B * (unsigned)A;          // < 2^31, checked by the host
    for (unsigned agent = blockIdx.x; agent < n_agents; agent += gridDim.x) {
    float v = 0.0f;
    int bf = -1;
    const int b = (int)(agent / (unsigned)A);
    const bool here = present ? present[agent] != 0 : true;
    if (here && threadIdx.x < 4) {                 // the (double precision) sin / cos is evaluated by 4 lanes of ONE warp
        float px, py, s, c;
        corner_of(state + 4 * (size_t)agent, lenwid + 2 * (size_t)agent, threadIdx.x, px, py, s, c);
        s_corner[threadIdx.x] = make_float2(px, py);
    }
    __syncthreads();
    if (here) {
        const MapDev& m = maps.m[env_map ? env_map[b] : 0];
        const float px = s_corner[k].x, py = s_corner[k].y;
        const float d2 = nearest_face(m, px, py, fmaxf(thr, 0.0f), &bf, lane);
        v = d2 > thr ? d2 : 0.0f;           // F.threshold(d2, thr, 0), infractions.py:172
    }
    if (lane == 0) {
        if (face) face[4 * (size_t)agent + k] = bf;
        s_v[k] = v;
    }
    __syncthreads();
    // sum of the 4 corners, in corner order
    if (threadIdx.x == 0) out[agent] = ((s_v[0] + s_v[1]) + s_v[2]) + s_v[3];
    __syncthreads();
    }
}

__global__ void __launch_bounds__(128) offroad_bwd_kernel(MapSetDev maps, const int32_t* __restrict__ env_map,
                                                          const float* __restrict__ state, const float* __restrict__ lenwid,
                                                          const uint8_t* __restrict__ present, int B, int A, float thr,
                                                          const int32_t* __restrict__ face, const float* __restrict__ gout,
                                                          float* __restrict__ g_state, float* __restrict__ g_lenwid) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t agent = t >> 2;
    const int k = (int)(t & 3);
    const bool valid = agent < (int64_t)B * A;
    float gx = 0.f, gy = 0.f, gpsi = 0.f, gl = 0.f, gw = 0.f;
    if (valid) {
        const int b = (int)(agent / A);
        const bool here = present ? present[agent] != 0 : true;
        const int f = face[t];
        if (here && f >= 0) {
            const MapDev& m = maps.m[env_map ? env_map[b] : 0];
            const float* st = state + 4 * agent;
            const float* lw = lenwid + 2 * agent;
            float px, py, s, c;
            corner_of(st, lw, k, px, py, s, c);
            const float2* tr = reinterpret_cast<const float2*>(m.tri + 6 * (size_t)f);
            const float2 v0 = tr[0], v1 = tr[1], v2 = tr[2];
            const float d2 = point_tri_dist2(px, py, v0.x, v0.y, v1.x, v1.y, v2.x, v2.y);
            if (d2 > thr) {
                // d2 is the smallest of the three edge distances; d d2 / d p = 2 (p - q)
                EdgeHit h = edge_dist2(px, py, v0.x, v0.y, v1.x, v1.y);
                const EdgeHit h2 = edge_dist2(px, py, v0.x, v0.y, v2.x, v2.y);
                const EdgeHit h3 = edge_dist2(px, py, v1.x, v1.y, v2.x, v2.y);
                if (h2.d2 < h.d2) h = h2;
                if (h3.d2 < h.d2) h = h3;
                const float g = gout[agent];
                const float dpx = 2.0f * (px - h.qx) * g, dpy = 2.0f * (py - h.qy) * g;
                const float sx = (k == 0 || k == 3) ? 0.5f : -0.5f;
                const float sy = (k < 2) ? 0.5f : -0.5f;
                const float ux = sx * lw[0], uy = sy * lw[1];
                gx = dpx;
                gy = dpy;
                gpsi = dpx * (-ux * s - uy * c) + dpy * (ux * c - uy * s);
                gl = sx * (dpx * c + dpy * s);
                gw = sy * (dpy * c - dpx * s);
            }
        }
    }
    const unsigned full = 0xffffffffu;
#pragma unroll
    for (int o = 1; o <= 2; o <<= 1) {
        gx += __shfl_xor_sync(full, gx, o);
        gy += __shfl_xor_sync(full, gy, o);
        gpsi += __shfl_xor_sync(full, gpsi, o);
        gl += __shfl_xor_sync(full, gl, o);
        gw += __shfl_xor_sync(full, gw, o);
    }
    if (valid && k == 0) {
        if (g_state) reinterpret_cast<float4*>(g_state)[agent] = make_float4(gx, gy, gpsi, 0.0f);
        if (g_lenwid) reinterpret_cast<float2*>(g_lenwid)[agent] = make_float2(gl, gw);
    }
}

}  // namespace

extern "C" int tds_offroad_fwd(const tds_map_t* const* maps, int32_t n_maps, const int32_t* d_env_map,
                               const float* d_state, const float* d_lenwid, const uint8_t* d_present,
                               int32_t B, int32_t A, float threshold, float* d_out, int32_t* d_face, void* stream) {
    TDS_REQUIRE(B >= 0 && A >= 0, "offroad: negative size");
    if (B == 0 || A == 0) return TDS_OK;
    TDS_REQUIRE(d_state && d_lenwid && d_out, "offroad: null pointer");
    MapSetDev set;
    if (int e = tds::gather_maps(maps, n_maps, set)) return e;
    // agents are dealt round-robin to a grid that fills the GPU a few times over (a CTA per agent would spend a
    // good part of its life being launched)
    TDS_REQUIRE((int64_t)B * A <= 2147483647LL, "offroad: too many agents");
    const int64_t grid = std::min<int64_t>((int64_t)B * A, (int64_t)tds::sm_count() * 64);
    offroad_fwd_kernel<<<(unsigned)grid, 128, 0, (cudaStream_t)stream>>>(
        set, d_env_map, d_state, d_lenwid, d_present, B, A, threshold, d_out, d_face);
    TDS_LAUNCH_OK();
    return TDS_OK;
}

extern "C" int tds_offroad_bwd(const tds_map_t* const* maps, int32_t n_maps, const int32_t* d_env_map,
                               const float* d_state, const float* d_lenwid, const uint8_t* d_present,
                               int32_t B, int32_t A, float threshold, const int32_t* d_face, const float* d_grad_out,
                               float* d_grad_state, float* d_grad_lenwid, void* stream) {
    TDS_REQUIRE(B >= 0 && A >= 0, "offroad_bwd: negative size");
    if (B == 0 || A == 0) return TDS_OK;
    TDS_REQUIRE(d_state && d_lenwid && d_face && d_grad_out, "offroad_bwd: null pointer");
    MapSetDev set;
    if (int e = tds::gather_maps(maps, n_maps, set)) return e;
    const int64_t threads = (int64_t)B * A * 4;
    offroad_bwd_kernel<<<(unsigned)((threads + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
        set, d_env_map, d_state, d_lenwid, d_present, B, A, threshold, d_face, d_grad_out, d_grad_state, d_grad_lenwid);
    TDS_LAUNCH_OK();
    return TDS_OK;
}
