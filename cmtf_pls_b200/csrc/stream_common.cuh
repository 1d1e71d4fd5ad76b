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

// Producer: one elected lane streams this CTA's row tiles into the smem ring.  With `y` the matching rows of
// Y (pitch_y doubles each, contiguous in memory) ride on the same barrier into ytiles[stage][tile_rows * pitch_y].
template <typename XT>
__device__ __forceinline__ void produce_tiles(const PassGeom& g, const XT* __restrict__ x, XT* tiles, uint64_t* full,
                                              uint64_t* empty, int c0, int slab_cols, int srow,
                                              const double* __restrict__ y = nullptr, int pitch_y = 0,
                                              double* ytiles = nullptr) {
    const long long n_tiles = (g.n_rows + g.tile_rows - 1) / g.tile_rows;
    const size_t stage_elems = (size_t)g.tile_rows * srow;
    const size_t ystage = (size_t)g.tile_rows * pitch_y;
    long long it = 0;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        const int s = (int)(it % g.stages);
        const uint32_t ph = (uint32_t)((it / g.stages) & 1);
        if (it >= g.stages) mbar_wait(&empty[s], ph ^ 1u);
        const long long r0 = tile * g.tile_rows;
        const int rows = (int)min((long long)g.tile_rows, g.n_rows - r0);
        XT* dst = tiles + s * stage_elems;
        const XT* src = x + r0 * g.pitch + c0;
        const uint32_t ybytes = y != nullptr ? (uint32_t)((size_t)rows * pitch_y * sizeof(double)) : 0u;
        if (g.n_slabs == 1) {
            const uint32_t bytes = (uint32_t)((size_t)rows * g.pitch * sizeof(XT));
            mbar_arrive_expect_tx(&full[s], bytes + ybytes);
            bulk_g2s(dst, src, bytes, &full[s]);
        } else {
            const uint32_t rb = (uint32_t)(slab_cols * sizeof(XT));
            mbar_arrive_expect_tx(&full[s], rb * rows + ybytes);
            for (int r = 0; r < rows; ++r) bulk_g2s(dst + (size_t)r * srow, src + (size_t)r * g.pitch, rb, &full[s]);
        }
        if (y != nullptr) bulk_g2s(ytiles + s * ystage, y + r0 * pitch_y, ybytes, &full[s]);
    }
}

// One 32-column chunk of a second-stage fold (reduce_cols order): 256 threads = 32 columns x 8 part-groups, each
// thread folds every 8th partial of its column in four independent chains.  Returns this thread's share.
__device__ __forceinline__ double fold_share(const FoldSet& S, int c, int q) {
    double t0 = 0.0, t1 = 0.0, t2 = 0.0, t3 = 0.0;
    if (c < S.n_cols) {
        int b = q;
        for (; b + 24 < S.n_parts; b += 32) {
            t0 += S.part[(size_t)(b + 0) * S.stride + c];
            t1 += S.part[(size_t)(b + 8) * S.stride + c];
            t2 += S.part[(size_t)(b + 16) * S.stride + c];
            t3 += S.part[(size_t)(b + 24) * S.stride + c];
        }
        for (; b < S.n_parts; b += 8) t0 += S.part[(size_t)b * S.stride + c];
    }
    return (t0 + t1) + (t2 + t3);
}

// which set does chunk `chunk` belong to (chunks are numbered set by set) and which is its first column
__device__ __forceinline__ int fold_locate(const FoldSet* sets, int n_sets, int chunk, int* c_base) {
    int s = 0;
    while (s + 1 < n_sets) {
        const int nc = (sets[s].n_cols + 31) >> 5;
        if (chunk < nc) break;
        chunk -= nc;
        ++s;
    }
    *c_base = chunk * 32;
    return s;
}

}  // namespace tpls
