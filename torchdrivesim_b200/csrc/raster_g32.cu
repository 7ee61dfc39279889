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

}  // namespace tds_raster
