// Host side of the birdview raster: the C ABI, the palette, the per-environment preparation of the dynamic primitives
// and the choice of a kernel variant.  The kernel itself is raster_kernel.cuh; its instantiations are compiled in
// raster_g32.cu / raster_g128.cu / raster_g256.cu (in parallel).
#include <mutex>
#include <vector>

#include "raster_kernel.cuh"

using namespace tds_raster;

namespace {

// ------------------------------------------------------------------ dynamic primitives (per env)

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

}  // namespace

static thread_local cudaEvent_t g_ev_start = nullptr, g_ev_stop = nullptr;

extern "C" void tds_raster_set_timing_events(void* start_event, void* stop_event) {
    g_ev_start = (cudaEvent_t)start_event;
    g_ev_stop = (cudaEvent_t)stop_event;
}

extern "C" int64_t tds_raster_workspace_bytes(int32_t B, int32_t N, int32_t L, int32_t R) {
    if (B < 0 || N < 0 || L < 0 || R < 0) return -1;
    // per-environment primitives | 32 bytes of work counters | the redo list of the LEAN kernels for up to B * N cameras
    return (int64_t)B * ws_env_bytes(3 * N + 2 * L + 2 * R) + 32 + (int64_t)B * N * 4;
}

static void make_palette_dev(const tds_palette_t* palette, int N, int L, int R, PaletteDev& pal) {
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
    for (int c = 0; c < TDS_MAX_CLASSES; c++) pal.plane_of_class[c] = -1;
    for (int k = 0; k < pal.n_classes; k++) pal.plane_of_class[pal.order[k]] = (int8_t)k;
}

// Library-owned scratch per (device, stream) for calls with more cameras than agents (free cameras of Simulator.render):
// the work counters of the persistent grids and the list of cameras a LEAN kernel hands to the general one.  Grown with cudaMalloc when a call has more cameras
// than any before it on that stream - which cannot happen inside a CUDA graph capture: call once before capturing.
static int raster_scratch(int64_t ncam, cudaStream_t st, int32_t** out) {
    struct Entry { int dev; cudaStream_t st; int32_t* ptr; int64_t cap; };
    static std::mutex mu;
    static std::vector<Entry> entries;
    int dev = -1;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lock(mu);
    Entry* hit = nullptr;
    for (auto& e : entries)
        if (e.dev == dev && e.st == st) hit = &e;
    if (!hit) { entries.push_back({dev, st, nullptr, 0}); hit = &entries.back(); }
    if (hit->cap < ncam) {
        cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
        cudaStreamIsCapturing(st, &cs);
        if (cs != cudaStreamCaptureStatusNone)
            return tds::fail(TDS_ERR_UNSUPPORTED, "raster: the first call with %lld cameras on a stream allocates scratch memory and cannot "
                                                  "be part of a CUDA graph capture; run it once before capturing", (long long)ncam);
        int32_t* p = nullptr;
        const int64_t cap = std::max<int64_t>(ncam, 4096);
        TDS_CUDA_OK(cudaMalloc(&p, (size_t)(8 + cap) * sizeof(int32_t)));
        // the old buffer is NOT released: work already enqueued, or a CUDA graph captured earlier on this stream, may still
        // refer to it (it stays valid for them; new calls use the larger one)
        hit->ptr = p;
        hit->cap = cap;
    }
    *out = hit->ptr;
    return TDS_OK;
}

// Hand-over buffers of the two-pass 64x64 kernels (bitplanes + lists of border-crossing faces), library-owned per
// (device, stream) and at most kTwoPassBytes: more cameras than fit are rendered in rounds.  Allocated by the first call
// on a stream; inside a CUDA graph capture a missing buffer is not an error - the call falls back to the one-pass kernel.
constexpr int64_t kTwoPassBytes = 512ll << 20;
static uint8_t* two_pass_scratch(int64_t bytes, cudaStream_t st) {
    struct Entry { int dev; cudaStream_t st; uint8_t* ptr; int64_t cap; };
    static std::mutex mu;
    static std::vector<Entry> entries;
    int dev = -1;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lock(mu);
    Entry* hit = nullptr;
    for (auto& e : entries)
        if (e.dev == dev && e.st == st) hit = &e;
    if (!hit) { entries.push_back({dev, st, nullptr, 0}); hit = &entries.back(); }
    if (hit->cap < bytes) {
        cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
        cudaStreamIsCapturing(st, &cs);
        if (cs != cudaStreamCaptureStatusNone) return nullptr;
        uint8_t* p = nullptr;
        if (cudaMalloc(&p, (size_t)bytes) != cudaSuccess) { cudaGetLastError(); return nullptr; }
        // the old buffer is kept (see raster_scratch): a graph captured with it stays valid
        hit->ptr = p;
        hit->cap = bytes;
    }
    return hit->ptr;
}

extern "C" int32_t tds_raster_rank_table(const tds_palette_t* palette, uint8_t h_rgb[][3], int32_t* h_class) {
    if (!palette || !h_rgb || palette->n_classes < 0 || palette->n_classes > TDS_MAX_CLASSES) return 0;
    PaletteDev pal = {};
    make_palette_dev(palette, 0, 0, 0, pal);
    h_rgb[0][0] = h_rgb[0][1] = h_rgb[0][2] = 0;
    if (h_class) h_class[0] = -1;
    for (int k = 0; k < pal.n_classes; k++) {
        for (int c = 0; c < 3; c++) h_rgb[k + 1][c] = palette->rgb[pal.order[k]][c];
        if (h_class) h_class[k + 1] = pal.order[k];
    }
    return pal.n_classes + 1;
}

extern "C" int tds_raster_birdview_fmt(const tds_map_t* const* maps, int32_t n_maps, const int32_t* d_env_map,
                                       int32_t B, int32_t Nc, int32_t N,
                                       const float* d_cam_xy, const float* d_cam_sc,
                                       const float* d_agent_state, const float* d_agent_size, const int32_t* d_agent_type,
                                       const uint8_t* d_present, int32_t present_per_camera,
                                       const float* d_tl_corners, const int32_t* d_tl_state, int32_t L,
                                       const float* d_rect_corners, const int32_t* d_rect_class, int32_t R,
                                       const float* d_cam_tris, const int32_t* d_cam_tri_class, int32_t Tc,
                                       const tds_palette_t* palette, float scale, int32_t res,
                                       const uint8_t* d_agent_class, int32_t image_format, void* d_out, void* d_workspace,
                                       void* stream) {
    TDS_REQUIRE(B >= 0 && Nc >= 0 && N >= 0 && L >= 0 && R >= 0, "raster: negative size");
    TDS_REQUIRE(B < (1 << 22), "raster: at most 4 194 303 environments per call");
    if (B == 0 || Nc == 0) return TDS_OK;
    TDS_REQUIRE(d_cam_xy && d_cam_sc && d_out && palette, "raster: null pointer");
    TDS_REQUIRE(N == 0 || (d_agent_state && d_agent_size), "raster: null agent tensors");
    TDS_REQUIRE(L == 0 || (d_tl_corners && d_tl_state), "raster: null traffic light tensors");
    TDS_REQUIRE(R == 0 || (d_rect_corners && d_rect_class), "raster: null rectangle tensors");
    TDS_REQUIRE(Tc >= 0 && (Tc == 0 || (d_cam_tris && d_cam_tri_class)), "raster: null per-camera triangle tensors");
    TDS_REQUIRE(res >= 4 && res % 4 == 0 && res <= 1024, "raster: res=%d must be a multiple of 4 in [4,1024]", res);
    TDS_REQUIRE(scale > 0.0f, "raster: scale must be positive");
    TDS_REQUIRE(palette->n_classes >= 0 && palette->n_classes <= TDS_MAX_CLASSES, "raster: bad palette");
    TDS_REQUIRE(image_format == TDS_IMAGE_F32 || image_format == TDS_IMAGE_U8 || image_format == TDS_IMAGE_RANK,
                "raster: unknown image format %d", image_format);
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
    make_palette_dev(palette, N, L, R, pal);

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
    a.ws = (const uint8_t*)d_workspace; a.out = d_out; a.out_format = image_format; a.agent_cls = N > 0 ? d_agent_class : nullptr;
    a.B = B; a.Nc = Nc; a.N = N; a.LR = L + R; a.T = T; a.present_per_camera = present_per_camera; a.res = res; a.scale = scale;
    const int64_t ncam = (int64_t)B * Nc;
    TDS_REQUIRE(ncam <= 2147483647LL, "raster: too many cameras");
    a.ncam = (int32_t)ncam;
    a.cam_tris = d_cam_tris; a.cam_cls = d_cam_tri_class; a.Tc = Tc;
    // Stage 1S plots every in-image vertex of a strip without asking whether its face is kept: that needs the image
    // (truncated coordinates in [0, res), i.e. pixel coordinates in (-1, res)) to lie inside the inner decision
    // radius of the 1.05x view quad, r_in = 0.525 res - band >= res / 2 + 1 (make_camera): tiles of 44 pixels and
    // more qualify at any zoom the band allows.  And it only pays while a typical strip is a handful of pixels - most
    // strips are then finished by plotting their six vertices; at closer zoom every face of a strip is classified on
    // its own anyway and the camera walks the plain face records of ALL faces instead (rec_all).
    {
        const float half = (float)res / 2.0f, band = 0.01f + 0.002f * (scale * half);
        // measured (Town01, 35 m): 64x64 (0.9 px) 2.37 -> 2.11 ms with strips; 128x128 (1.8 px) 3.08 -> 3.33 ms; 256x256 2.04 -> 2.55 ms
        float strip_px = 0.f;
        for (int i = 0; i < n_maps; i++) strip_px = std::max(strip_px, set.m[i].strip_len * scale * half);
        a.strip_mode = (0.525f * (float)res - band >= half + 1.001f && strip_px > 0.f && strip_px <= 1.4f) ? 1 : 0;
        if (const char* e = getenv("TDS_RASTER_STRIPS"))
            a.strip_mode = (atoi(e) != 0 && 0.525f * (float)res - band >= half + 1.001f) ? 1 : 0;
        // sliver quads (pairs of strip faces as one item drawn from the pattern table): 64x64 warp-per-camera kernels
        a.quad_table = (a.strip_mode && res == 64) ? reinterpret_cast<const uint4*>(tds::quad_table_device()) : nullptr;
        if (const char* e = getenv("TDS_RASTER_QUADS")) { if (atoi(e) == 0) a.quad_table = nullptr; }
    }
    const int K = pal.n_classes;
    TDS_REQUIRE(K <= 31, "raster: at most 31 active classes (got %d)", K);
    // The LEAN kernels hold the common case only (see raster_kernel.cuh); whatever they cannot finish exactly they put
    // on a list, and the general kernel of the same shape renders those cameras again right behind them.
    bool lean = Tc == 0 && a.agent_cls == nullptr;
    if (const char* e = getenv("TDS_RASTER_LEAN")) lean = lean && atoi(e) != 0;
    // counters + redo list: the tail of the caller's workspace holds B * N cameras (one per agent, the egocentric case);
    // more cameras than that take library-owned scratch memory
    int32_t* scratch = reinterpret_cast<int32_t*>((uint8_t*)d_workspace + (int64_t)B * ws_env_bytes(T));
    if (ncam > (int64_t)B * N)
        if (int e = raster_scratch(ncam, st, &scratch)) return e;
    TDS_CUDA_OK(cudaMemsetAsync(scratch, 0, 8 * sizeof(int32_t), st));
    a.next_cam = scratch;              // work counter of the first launch
    a.redo = scratch + 4;              // [0] = count, [4 ..] = cameras
    a.cam_list = nullptr;
    a.cam_begin = 0; a.planes_io = nullptr; a.plane_stride = 0; a.clip_list = nullptr; a.clip_count = nullptr; a.clip_cap = kClipCap;
    if (const char* e = getenv("TDS_RASTER_CLIP_CAP")) a.clip_cap = std::min(std::max(atoi(e), 0), kClipCap);     // test hook: exercises the redo list

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

    LaunchCfg c;
    c.set = set; c.a = a; c.pal = pal; c.K = K; c.res = res; c.sms = tds::sm_count(); c.ncam = ncam; c.st = st;
    c.ev_start = g_ev_start; c.ev_stop = g_ev_stop;
    const bool f32 = image_format == TDS_IMAGE_F32;
    auto go = [&](const LaunchCfg& cfg, bool is_lean) -> int {
        if (G == 32) return launch_g32(cfg, f32, is_lean);
        if (G == 128) return launch_g128(cfg, f32, is_lean);
        return launch_g256(cfg, G, f32, is_lean);
    };
    // Two passes for the 64x64 warp-per-camera kernels (LEAN or general, up to 15 active classes): draw (everything but the border-crossing faces) -> finish (those faces +
    // resolve), handing bitplanes and face lists over in library-owned memory.  Each program fits the instruction caches
    // where the one-pass kernel does not.  TDS_RASTER_TWO_PASS=0 forces one pass.
    bool two_pass = G == 32 && g32_two_pass_available(c);
    if (const char* e = getenv("TDS_RASTER_TWO_PASS")) two_pass = two_pass && atoi(e) != 0;
    int64_t round_cams = 0;
    uint8_t* hand = nullptr;
    if (two_pass) {
        const int64_t per_cam = two_pass_bytes_per_camera(K);
        int64_t budget = kTwoPassBytes;
        if (const char* e = getenv("TDS_RASTER_TWO_PASS_KB")) budget = std::max<int64_t>(atoll(e), 8) << 10;     // test hook: small rounds
        round_cams = std::min<int64_t>(ncam, std::max<int64_t>(budget / per_cam, 1));
        hand = two_pass_scratch(round_cams * per_cam, st);
        two_pass = hand != nullptr;
    }
    if (two_pass) {
        const int KS = two_pass_planes(K);
        if (g_ev_start && g_ev_stop) cudaEventRecord(g_ev_start, st);   // the timed region spans both passes (and all rounds)
        for (int64_t begin = 0; begin < ncam; begin += round_cams) {
            const int64_t end = std::min<int64_t>(ncam, begin + round_cams);
            LaunchCfg d = c;
            d.ev_start = d.ev_stop = nullptr;
            d.ncam = end;
            d.a.ncam = (int32_t)end;
            d.a.cam_begin = (int32_t)begin;
            d.a.planes_io = reinterpret_cast<uint32_t*>(hand);
            d.a.plane_stride = KS * 32;
            d.a.clip_list = reinterpret_cast<uint4*>(hand + round_cams * KS * 512);
            d.a.clip_count = reinterpret_cast<int32_t*>(hand + round_cams * (KS * 512 + (int64_t)kClipCap * 16));
            // work counters: [0] draw pass, [2] finish pass (the redo list at [4..] spans all rounds)
            if (begin > 0) TDS_CUDA_OK(cudaMemsetAsync(scratch, 0, 4 * sizeof(int32_t), st));     // round 0: zeroed above
            d.a.next_cam = scratch;
            if (int e = launch_g32_draw(d, lean)) return e;
            LaunchCfg f = d;
            f.a.next_cam = scratch + 2;
            f.ev_start = f.ev_stop = nullptr;
            if (int e = launch_g32_finish(f, f32)) return e;
        }
        if (g_ev_start && g_ev_stop) cudaEventRecord(g_ev_stop, st);
    } else {
        if (int e = go(c, lean)) return e;
    }
    if (lean || two_pass) {
        // the cameras on the list (normally none: every CTA of this launch leaves at once)
        LaunchCfg r = c;
        r.a.next_cam = scratch + 1;
        r.a.cam_list = scratch + 4;
        r.ev_start = r.ev_stop = nullptr;
        if (int e = go(r, false)) return e;
    }
    return TDS_OK;
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
    return tds_raster_birdview_fmt(maps, n_maps, d_env_map, B, Nc, N, d_cam_xy, d_cam_sc, d_agent_state, d_agent_size, d_agent_type,
                                   d_present, present_per_camera, d_tl_corners, d_tl_state, L, d_rect_corners, d_rect_class, R,
                                   d_cam_tris, d_cam_tri_class, Tc, palette, scale, res, nullptr, TDS_IMAGE_F32, d_out, d_workspace, stream);
}
