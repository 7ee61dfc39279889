// Instantiations of the raster kernel with a CTA of 4 independent warps per camera (tiles above 96x96); see raster_kernel.cuh.
#include "raster_kernel.cuh"

namespace tds_raster {

template <int RES, int NS, bool SMALL>
static int pick(const LaunchCfg& c, bool f32, bool lean) {
    if (lean) {
        return f32 ? launch_variant(raster_kernel<128, RES, NS, SMALL, 0, true, true>, c, 1, 128, false)
                   : launch_variant(raster_kernel<128, RES, NS, SMALL, 0, false, true>, c, 1, 128, false);
    }
    return f32 ? launch_variant(raster_kernel<128, RES, NS, SMALL, 0, true, false>, c, 1, 128, false)
               : launch_variant(raster_kernel<128, RES, NS, SMALL, 0, false, false>, c, 1, 128, false);
}

int launch_g128(const LaunchCfg& c, bool f32, bool lean) {
    const bool small = c.res <= 128;
    if (c.K <= 7) {
        if (c.res == 128) return pick<128, 3, true>(c, f32, lean);       // configs 3 and 4: constant strides
        if (c.res == 256) return pick<256, 3, false>(c, f32, lean);
        return small ? pick<0, 3, true>(c, f32, lean) : pick<0, 3, false>(c, f32, lean);
    }
    if (c.res == 128) return pick<128, 5, true>(c, f32, lean);
    return small ? pick<0, 5, true>(c, f32, lean) : pick<0, 5, false>(c, f32, lean);
}

}  // namespace tds_raster
