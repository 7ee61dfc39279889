// Static map acceleration structures (device view).  Built once per map by tds_map_create.
#pragma once
#include "tds_common.cuh"

namespace tds {

constexpr int kMaxSlots = 8;        // distinct palette classes among the static faces of one map
constexpr int kMaxRasterRows = 64;  // grid rows a single camera may touch

struct MapDev {
    // ---- raster grid: triangles binned by the cell of each VERTEX (the reference keeps a face
    // iff any vertex is inside the view quad, mesh.py:311-313).  A record is two float4:
    //   (x0, y0, x1, y1), (x2, y2, meta, unused); meta = own_bits | class << 8, own bit i set = vertex i lies
    //   in this cell.  Records are sorted by cell (row-major), so one grid row is one contiguous range.
    const float4* rec;
    const int32_t* rcell;           // [rgx * rgy + 1] CSR offsets into rec (in records)
    float rx0, ry0, rcs, rinv;
    int32_t rgx, rgy, n_slots;
    int32_t slot_of_class[TDS_MAX_CLASSES];   // -1: the map has no face of that class
    // ---- offroad grid: triangles binned by bounding box overlap
    const float* tri;               // [nf][6] x0,y0,x1,y1,x2,y2 (indexed by face, for the backward)
    const int32_t* ocell;           // [ogx * ogy + 1] CSR offsets into orec (in records)
    const float4* orec;             // per cell entry: (x0,y0,x1,y1), (x2,y2, face index bits, unused); within a
                                    // cell the faces are sorted by decreasing area (a containing face is found early)
    float ox0, oy0, ocs, oinv;
    int32_t ogx, ogy, nf;
};

struct MapSetDev {
    MapDev m[TDS_MAX_MAPS];
};

}  // namespace tds

struct tds_map {
    tds::MapDev dev;
    tds_map_info_t info;
    void* allocations[5];
    int device;
};

namespace tds {
int gather_maps(const tds_map_t* const* maps, int32_t n_maps, MapSetDev& out);
}
