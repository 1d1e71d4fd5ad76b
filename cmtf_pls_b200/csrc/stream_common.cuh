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

// Which tiles a CTA walks: tile = first, first + step, ... < end.  Normal builds interleave the CTAs (tile = blockIdx.x + k *
// gridDim.x); probe builds (TPLS_DBG & 8) can give every CTA one contiguous run of tiles instead.
struct TileWalk {
    long long first, end, step;
};
__device__ __forceinline__ TileWalk tile_walk(const PassGeom& g) {
    const long long n_tiles = (g.n_rows + g.tile_rows - 1) / g.tile_rows;
    TileWalk w{(long long)blockIdx.x, n_tiles, (long long)gridDim.x};
#ifdef TPLS_PROBE
    if (g.dbg & 8) {
        const long long per = (n_tiles + gridDim.x - 1) / gridDim.x;
        w.first = blockIdx.x * per;
        w.end = min(n_tiles, w.first + per);
        w.step = 1;
    }
#endif
    return w;
}

// Position in a ring of n mbarrier-guarded buffers, advanced without the 64-bit division and modulo that
// (iteration % n, iteration / n) cost once per tile and warp (they were a third of a consumer's instructions on 2-row tiles).
struct RingPos {
    int idx;         // buffer of the current iteration
    uint32_t phase;  // parity of the number of completed trips around the ring
    bool wrapped;    // at least one trip around the ring is complete (the buffer has been used before)
    __device__ __forceinline__ RingPos() : idx(0), phase(0u), wrapped(false) {}
    __device__ __forceinline__ void advance(int n) {
        if (++idx == n) {
            idx = 0;
            phase ^= 1u;
            wrapped = true;
        }
    }
};

// rows of tile `tile` (the last tile of a shard may be short)
__device__ __forceinline__ int tile_rows_of(const PassGeom& g, long long tile, long long n_tiles) {
    return tile == n_tiles - 1 ? (int)(g.n_rows - tile * g.tile_rows) : g.tile_rows;
}

// Producer: one elected lane streams this CTA's row tiles into the smem ring.  With `y` the matching rows of
// Y (pitch_y doubles each, contiguous in memory) ride on the same barrier into ytiles[stage][tile_rows * pitch_y].
template <typename XT>
__device__ __forceinline__ void produce_tiles(const PassGeom& g, const XT* __restrict__ x, XT* tiles, uint64_t* full,
                                              uint64_t* empty, int c0, int slab_cols, int srow,
                                              const double* __restrict__ y = nullptr, int pitch_y = 0,
                                              double* ytiles = nullptr) {
    const TileWalk tw = tile_walk(g);
    const long long n_tiles = (g.n_rows + g.tile_rows - 1) / g.tile_rows;
    const size_t stage_elems = (size_t)g.tile_rows * srow;
    const size_t ystage = (size_t)g.tile_rows * pitch_y;
    RingPos rp;
    for (long long tile = tw.first; tile < tw.end; tile += tw.step, rp.advance(g.stages)) {
        const int s = rp.idx;
        if (rp.wrapped) mbar_wait(&empty[s], rp.phase ^ 1u);
        const long long r0 = tile * g.tile_rows;
        const int rows = tile_rows_of(g, tile, n_tiles);
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

// One 32-column chunk of a second-stage fold: 256 threads = 32 columns x 8 part-groups, each
// thread folds every 8th partial of its column in four independent chains.  Returns this thread's share.
__device__ __forceinline__ double fold_share(const FoldSet& S, int c, int q) {
    // eight independent chains: ~300 partials per column are 37 loads per thread, i.e. five dependent L2 round trips
    // (with four chains it was ten, and this fold sits on the critical path of every trip)
    double t[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    if (c < S.n_cols) {
        int b = q;
        for (; b + 56 < S.n_parts; b += 64) {
#pragma unroll
            for (int k = 0; k < 8; ++k) t[k] += S.part[(size_t)(b + 8 * k) * S.stride + c];
        }
        for (int k = 0; b < S.n_parts; b += 8, ++k) t[k & 7] += S.part[(size_t)b * S.stride + c];
    }
    return ((t[0] + t[1]) + (t[2] + t[3])) + ((t[4] + t[5]) + (t[6] + t[7]));
}

// The same for a NARROW set (at most 8 columns, e.g. the partials of q = Y't): 256 threads = 8 columns x 32
// part-groups, so that the ~300 partials of a column are folded by 32 threads instead of 8 (a fold by 8 threads is
// ten dependent L2 round trips, ~10 us on the critical path of every trip).  fold32 is [32][9] doubles of shared
// memory; returns the column's sum in the threads with grp == 0 (valid for col < n_cols).
__device__ __forceinline__ double fold_narrow(const FoldSet& S, double (*fold32)[9]) {
    const int col = threadIdx.x & 7, grp = threadIdx.x >> 3;
    double t0 = 0.0, t1 = 0.0;
    if (col < S.n_cols) {
        int b = grp;
        for (; b + 32 < S.n_parts; b += 64) {
            t0 += S.part[(size_t)b * S.stride + col];
            t1 += S.part[(size_t)(b + 32) * S.stride + col];
        }
        if (b < S.n_parts) t0 += S.part[(size_t)b * S.stride + col];
    }
    __syncthreads();
    fold32[grp][col] = t0 + t1;
    __syncthreads();
    double t = 0.0;
    if (grp == 0) {
#pragma unroll
        for (int k = 0; k < 32; ++k) t += fold32[k][col];
    }
    return t;
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
