// Instantiations of the raster kernel with a CTA of 8 or 16 independent warps per camera (the largest tiles / many
// classes: whatever keeps 16 warps resident on an SM); see raster_kernel.cuh.
#include "raster_kernel.cuh"

namespace tds_raster {

template <int G, int NS, bool SMALL>
static int pick(const LaunchCfg& c, bool f32, bool lean) {
    if (lean) {
        return f32 ? launch_variant(raster_kernel<G, 0, NS, SMALL, 0, true, true>, c, 1, G, false)
                   : launch_variant(raster_kernel<G, 0, NS, SMALL, 0, false, true>, c, 1, G, false);
    }
    return f32 ? launch_variant(raster_kernel<G, 0, NS, SMALL, 0, true, false>, c, 1, G, false)
               : launch_variant(raster_kernel<G, 0, NS, SMALL, 0, false, false>, c, 1, G, false);
}

int launch_g256(const LaunchCfg& c, int G, bool f32, bool lean) {
    const bool small = c.res <= 128;
    if (c.K <= 7) {
        if (G == 256) return small ? pick<256, 3, true>(c, f32, lean) : pick<256, 3, false>(c, f32, lean);
        return pick<512, 3, false>(c, f32, lean);
    }
    if (G == 256) return small ? pick<256, 5, true>(c, f32, lean) : pick<256, 5, false>(c, f32, lean);
    return pick<512, 5, false>(c, f32, lean);
}

}  // namespace tds_raster
