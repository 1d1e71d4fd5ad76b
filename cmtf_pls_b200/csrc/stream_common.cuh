// Device helpers shared by the column and row streaming passes.
#pragma once

#include "passes.cuh"

namespace tpls {

template <typename XT>
struct VecOf;
template <>
struct VecOf<float> {
    using type = float4;
    static constexpr int N = 4;
};
template <>
struct VecOf<double> {
    using type = double2;
    static constexpr int N = 2;
};

template <typename XT>
union Pack {
    typename VecOf<XT>::type v;
    XT e[VecOf<XT>::N];
};

// Producer: one elected lane streams this CTA's row tiles into the smem ring.
template <typename XT>
__device__ __forceinline__ void produce_tiles(const PassGeom& g, const XT* __restrict__ x, XT* tiles, uint64_t* full,
                                              uint64_t* empty, int c0, int slab_cols, int srow) {
    const long long n_tiles = (g.n_rows + g.tile_rows - 1) / g.tile_rows;
    const size_t stage_elems = (size_t)g.tile_rows * srow;
    long long it = 0;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        const int s = (int)(it % g.stages);
        const uint32_t ph = (uint32_t)((it / g.stages) & 1);
        if (it >= g.stages) mbar_wait(&empty[s], ph ^ 1u);
        const long long r0 = tile * g.tile_rows;
        const int rows = (int)min((long long)g.tile_rows, g.n_rows - r0);
        XT* dst = tiles + s * stage_elems;
        const XT* src = x + r0 * g.pitch + c0;
        if (g.n_slabs == 1) {
            const uint32_t bytes = (uint32_t)((size_t)rows * g.pitch * sizeof(XT));
            mbar_arrive_expect_tx(&full[s], bytes);
            bulk_g2s(dst, src, bytes, &full[s]);
        } else {
            const uint32_t rb = (uint32_t)(slab_cols * sizeof(XT));
            mbar_arrive_expect_tx(&full[s], rb * rows);
            for (int r = 0; r < rows; ++r) bulk_g2s(dst + (size_t)r * srow, src + (size_t)r * g.pitch, rb, &full[s]);
        }
    }
}


}  // namespace tpls
