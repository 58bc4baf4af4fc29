// Device-side view of the spatial hash (K0) and its cell lookup, shared by the kNN kernels (psl_grid.cu) and the map
// maintenance kernels (psl_map.cu).
#pragma once
#include "psl_common.cuh"

namespace psl {

struct GridDev {
    const float4* pts;
    const uint64_t* tkeys;
    const uint2* tvals;
    uint32_t mask;
    int n;
    float inv_cell;
    float r_small;      // radius of the first search pass (0 = single pass with the full query radius)
    const psl_grid_meta* meta;   // device-side (capacity, n, r_small) overriding the launch-time values (graph replays)
};

// kernels call this first: with a device meta block the launch-time copies are stale after an in-place rebuild
__device__ __forceinline__ void grid_resolve(GridDev& g) {
    if (g.meta) {
        const uint4 m = __ldg(reinterpret_cast<const uint4*>(g.meta));
        g.mask = m.x ? m.x - 1u : 0u;
        g.n = (int)m.y;
        g.r_small = __uint_as_float(m.z);
    }
}

__device__ __forceinline__ uint2 grid_lookup(const GridDev& g, uint64_t key) {
    uint32_t slot = (uint32_t)mix64(key) & g.mask;
    for (;;) {
        const uint64_t k = __ldg(g.tkeys + slot);
        const uint2 v = __ldg(g.tvals + slot);       // fetched with the key (one round trip when the first probe hits)
        if (k == key) return v;
        if (k == kEmptyKey) return make_uint2(0u, 0u);
        slot = (slot + 1) & g.mask;
    }
}

static inline int make_grid_dev(const psl_grid* gh, GridDev* g) {
    PSL_REQUIRE(gh != nullptr, "grid is NULL");
    PSL_REQUIRE(gh->n >= 0 && gh->cell > 0.f, "grid not built");
    PSL_REQUIRE(gh->n == 0 || (gh->capacity && (gh->capacity & (gh->capacity - 1)) == 0), "capacity must be 2^k");
    g->pts = reinterpret_cast<const float4*>(gh->sorted_pts);
    g->tkeys = gh->table_keys;
    g->tvals = reinterpret_cast<const uint2*>(gh->table_vals);
    g->mask = gh->capacity ? gh->capacity - 1 : 0;
    g->n = gh->n;
    g->inv_cell = 1.0f / gh->cell;
    g->r_small = gh->r_small;
    g->meta = gh->meta;
    return 0;
}

}  // namespace psl
