// Instantiations of the raster kernel with one WARP per camera (tiles up to 96x96); see raster_kernel.cuh.
#include "raster_kernel.cuh"

namespace tds_raster {

template <int RES, int NS, int KS>
static int pick(const LaunchCfg& c, bool f32, bool lean, bool static_smem) {
    if (lean) {
        return f32 ? launch_variant(raster_kernel<32, RES, NS, true, KS, true, true>, c, 4, 128, static_smem)
                   : launch_variant(raster_kernel<32, RES, NS, true, KS, false, true>, c, 4, 128, static_smem);
    }
    return f32 ? launch_variant(raster_kernel<32, RES, NS, true, KS, true, false>, c, 4, 128, static_smem)
               : launch_variant(raster_kernel<32, RES, NS, true, KS, false, false>, c, 4, 128, static_smem);
}

int launch_g32(const LaunchCfg& c, bool f32, bool lean) {
    if (c.K <= 7) {
        if (c.res == 64 && c.K <= 5) return pick<64, 3, 5>(c, f32, lean, true);
        if (c.res == 64) return pick<64, 3, 7>(c, f32, lean, true);
        return pick<0, 3, 0>(c, f32, lean, false);
    }
    if (c.res == 64) return pick<64, 5, 0>(c, f32, lean, false);
    return pick<0, 5, 0>(c, f32, lean, false);
}

// ---- two-pass form of the 64x64 kernels: static shared memory and 3-bit ranks up to 7 active classes, dynamic shared
// memory and 5-bit ranks up to 15 (a scene with traffic lights has 10)
bool g32_two_pass_available(const LaunchCfg& c) { return c.res == 64 && c.K <= 15; }

int launch_g32_draw(const LaunchCfg& c, bool lean) {
    if (lean) {
        if (c.K <= 5) return launch_variant(raster_kernel<32, 64, 3, true, 5, true, true, 1>, c, 4, 128, true);
        if (c.K <= 7) return launch_variant(raster_kernel<32, 64, 3, true, 7, true, true, 1>, c, 4, 128, true);
        return launch_variant(raster_kernel<32, 64, 5, true, 0, true, true, 1>, c, 4, 128, false);
    }
    // scenes with per-camera triangles (goal-waypoint discs) or per-camera agent classes (custom colours)
    if (c.K <= 5) return launch_variant(raster_kernel<32, 64, 3, true, 5, true, false, 1>, c, 4, 128, true);
    if (c.K <= 7) return launch_variant(raster_kernel<32, 64, 3, true, 7, true, false, 1>, c, 4, 128, true);
    return launch_variant(raster_kernel<32, 64, 5, true, 0, true, false, 1>, c, 4, 128, false);
}

int launch_g32_finish(const LaunchCfg& c, bool f32) {
    if (c.K <= 5) return f32 ? launch_finish(raster_finish_kernel<3, 5, true>, c) : launch_finish(raster_finish_kernel<3, 5, false>, c);
    if (c.K <= 7) return f32 ? launch_finish(raster_finish_kernel<3, 7, true>, c) : launch_finish(raster_finish_kernel<3, 7, false>, c);
    return f32 ? launch_finish(raster_finish_kernel<5, 15, true>, c) : launch_finish(raster_finish_kernel<5, 15, false>, c);
}

}  // namespace tds_raster
