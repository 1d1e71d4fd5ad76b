// Single-pass multi-component projection and rank-R reconstruction -- the steps either side of the fit
// (reference cmtf_pls/tpls.py:122-165,188-189, cmtf.py:142-237).
//
// multiproj_kernel: for complete data the R projections of transform() are independent dot products of the SAME
// rows, S[i, a] = sum_c x[i, c] * w_a[c], so X is read ONCE with R accumulators per row instead of once per
// component.  That is an (N x P) by (P x R) skinny GEMM in fp64 -- at R = 10 the arithmetic (P*R FMAs per row)
// costs as much as the HBM read, so it runs on the fp64 tensor-core path (mma.sync.m8n8k4.f64; SASS DMMA), the one
// place of this library where that is legitimate: 256 FMAs per instruction for one LDS of A and a reusable LDS of B,
// where plain DFMAs would be bound by shared-memory loads.
//   * a CTA owns one column slab (512 fp32 / 256 fp64 columns, half that beyond 16 components) of all its rows: the slab of W (R padded to a multiple of 8
//     components, fp64) stays in shared memory for the whole kernel, laid out [column][RP + 4] so that the B-fragment
//     loads of a half-warp fall on 16 different bank pairs;
//   * row tiles (kTileRows rows x slab) arrive through the usual ring (producer lane, one 1-D bulk copy per row,
//     mbarrier complete_tx); rows are padded by 16/32 bytes so that the A-fragment loads are conflict-free too;
//   * the 8 consumer warps split the slab's columns (k index); each keeps the 16 x RP accumulators of the tile in
//     its C fragments, the warps meet in shared memory once per tile and write the slab's partial scores;
//   * multiproj_finish sums the slabs and the coupled tensors, divides by L (cmtf.py:120) and runs the deflation
//     recurrence of transform on the scores alone (see tpls_transform); a NaN anywhere in a row surfaces here and
//     sends the caller to the sequential masked path.
//
// reconstruct_kernel: X_hat[i, c] = mean[c] + sum_a T[i, a] * w_a[c]   (util.py:18-20 + the mean), written once.
#include "passes.cuh"
#include "stream_common.cuh"

#include <algorithm>

namespace tpls {

namespace {

// columns per CTA slab: the slab of W, three staged row tiles and the meeting area must fit 227 KB
template <typename XT, int NB>
struct SlabW {
    static constexpr int value = 512 / (int)(sizeof(XT) / 4) / (NB > 2 ? 2 : 1);
};
constexpr int kTileRows = 16;    // rows per staged tile (two 8-row MMA blocks)
constexpr int kMpStages = 3;
constexpr int kMpWarps = 8;      // consumer warps; warp w owns columns [w * SW/8, (w+1) * SW/8) of the slab
constexpr int kMpThreads = kMpWarps * 32 + 32;
constexpr int kRowPad = 4;       // elements of padding per staged row

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

}  // namespace

// NB = blocks of 8 components (R <= 8 * NB)
template <typename XT, int NB>
__global__ void __launch_bounds__(kMpThreads, 1) multiproj_kernel(const __grid_constant__ MultiProjArgs a) {
    pdl_prologue();
    constexpr int RP = NB * 8;
    constexpr int WS = RP + 4;                 // leading dimension of the W slab in shared memory (doubles)
    constexpr int kSlabW = SlabW<XT, NB>::value;
    constexpr int SROW = kSlabW + kRowPad;     // leading dimension of a staged row (elements)
    extern __shared__ __align__(128) unsigned char smem[];
    double* wsm = reinterpret_cast<double*>(smem);                                   // [kSlabW][WS]
    XT* tiles = reinterpret_cast<XT*>(smem + sizeof(double) * kSlabW * WS);          // [stages][kTileRows][SROW]
    constexpr size_t stage_elems = (size_t)kTileRows * SROW;
    double* red = reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(tiles) + kMpStages * stage_elems * sizeof(XT));
    // red: [kMpWarps][kTileRows][RP]
    uint64_t* full = reinterpret_cast<uint64_t*>(red + (size_t)kMpWarps * kTileRows * RP);
    uint64_t* empty = full + kMpStages;

    const int c0 = blockIdx.y * kSlabW;
    const int slab_cols = min(kSlabW, a.pitch - c0);   // multiple of 16 bytes
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long n_tiles = (a.n_rows + kTileRows - 1) / kTileRows;

    if (tid == 0) {
        for (int s = 0; s < kMpStages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], kMpWarps);
        }
        fence_mbar_init();
    }
    // the slab of W: wsm[c][n] = W[n][c0 + c], zero for the pad components and the columns past the slab's end
    for (int e = tid; e < kSlabW * RP; e += kMpThreads) {
        const int n = e / kSlabW, c = e - n * kSlabW;      // consecutive threads read consecutive columns of W
        double v = 0.0;
        if (n < a.n_comp && c < slab_cols && c0 + c < a.p) v = a.w[(size_t)n * a.w_pitch + c0 + c];
        wsm[(size_t)c * WS + n] = v;
    }
    __syncthreads();

    if (warp == kMpWarps) {  // ---- producer ----
        if (lane == 0) {
            const XT* x = reinterpret_cast<const XT*>(a.x);
            const uint32_t rb = (uint32_t)(slab_cols * sizeof(XT));
            long long it = 0;
            for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
                const int s = (int)(it % kMpStages);
                const uint32_t ph = (uint32_t)((it / kMpStages) & 1);
                if (it >= kMpStages) mbar_wait(&empty[s], ph ^ 1u);
                const long long r0 = tile * kTileRows;
                const int rows = (int)min((long long)kTileRows, a.n_rows - r0);
                XT* dst = tiles + s * stage_elems;
                mbar_arrive_expect_tx(&full[s], rb * rows);
                for (int r = 0; r < rows; ++r)
                    bulk_g2s(dst + (size_t)r * SROW, x + (r0 + r) * a.pitch + c0, rb, &full[s]);
            }
        }
        return;
    }

    // ---- consumers ----
    const int g = lane >> 2, t4 = lane & 3;           // MMA fragment coordinates
    const int kw = kSlabW / kMpWarps;                 // columns per warp
    const int k_lo = warp * kw;
    const int k_hi = min(k_lo + kw, (slab_cols + 3) & ~3);
    long long it = 0;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        const int s = (int)(it % kMpStages);
        const uint32_t ph = (uint32_t)((it / kMpStages) & 1);
        const long long r0 = tile * kTileRows;
        const int rows = (int)min((long long)kTileRows, a.n_rows - r0);
        mbar_wait(&full[s], ph);
        const XT* tp = tiles + s * stage_elems;
        double acc[2][NB][2];
#pragma unroll
        for (int rb = 0; rb < 2; ++rb)
#pragma unroll
            for (int nb = 0; nb < NB; ++nb) acc[rb][nb][0] = acc[rb][nb][1] = 0.0;
        // rows past the end of a short last tile hold stale data: their lanes feed zeros
        const bool live0 = g < rows, live1 = 8 + g < rows;
        for (int k = k_lo; k < k_hi; k += 4) {
            const bool kin = k + t4 < slab_cols;
            double bfr[NB];
#pragma unroll
            for (int nb = 0; nb < NB; ++nb) bfr[nb] = wsm[(size_t)(k + t4) * WS + nb * 8 + g];
            const double a0 = (live0 && kin) ? (double)tp[(size_t)g * SROW + k + t4] : 0.0;
            const double a1 = (live1 && kin) ? (double)tp[(size_t)(8 + g) * SROW + k + t4] : 0.0;
#pragma unroll
            for (int nb = 0; nb < NB; ++nb) {
                dmma(acc[0][nb][0], acc[0][nb][1], a0, bfr[nb]);
                dmma(acc[1][nb][0], acc[1][nb][1], a1, bfr[nb]);
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
        // ---- the warps' partial tiles meet in shared memory; thread e sums entry e over the warps ----
        double* mine = red + (size_t)warp * kTileRows * RP;
#pragma unroll
        for (int rb = 0; rb < 2; ++rb)
#pragma unroll
            for (int nb = 0; nb < NB; ++nb)
                *reinterpret_cast<double2*>(mine + (size_t)(rb * 8 + g) * RP + nb * 8 + 2 * t4) =
                    make_double2(acc[rb][nb][0], acc[rb][nb][1]);
        named_bar_sync(1, kMpWarps * 32);
        for (int e = tid; e < kTileRows * RP; e += kMpWarps * 32) {
            const int r = e / RP, n = e - r * RP;
            if (r < rows && n < a.n_comp) {
                double t = 0.0;
#pragma unroll
                for (int w = 0; w < kMpWarps; ++w) t += red[(size_t)w * kTileRows * RP + e];
                a.part[((size_t)blockIdx.y * a.n_rows + r0 + r) * a.n_comp + n] = t;
            }
        }
        named_bar_sync(1, kMpWarps * 32);
    }
}

int multiproj_slabs(int dtype, int n_comp, int pitch) {
    const int nb = (n_comp + 7) / 8;
    const int sw = dtype == 0 ? (nb > 2 ? SlabW<float, 4>::value : SlabW<float, 1>::value)
                              : (nb > 2 ? SlabW<double, 4>::value : SlabW<double, 1>::value);
    return (pitch + sw - 1) / sw;
}

template <typename XT, int NB>
static cudaError_t run_multiproj(const MultiProjArgs& a, int sm_count, cudaStream_t s) {
    auto kern = multiproj_kernel<XT, NB>;
    constexpr int RP = NB * 8;
    constexpr int kSlabW = SlabW<XT, NB>::value;
    const size_t smem = sizeof(double) * kSlabW * (RP + 4) + kMpStages * (size_t)kTileRows * (kSlabW + kRowPad) * sizeof(XT) +
                        sizeof(double) * kMpWarps * kTileRows * RP + 128;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const int n_slabs = (a.pitch + kSlabW - 1) / kSlabW;
    const long long n_tiles = (a.n_rows + kTileRows - 1) / kTileRows;
    const int gx = (int)std::max<long long>(1, std::min<long long>(n_tiles, std::max(1, sm_count / n_slabs)));
    launch_k(kern, dim3(gx, n_slabs), dim3(kMpThreads), smem, s, a);
    return cudaGetLastError();
}

cudaError_t launch_multiproj(int dtype, const MultiProjArgs& a, int sm_count, cudaStream_t s) {
    if (a.n_comp < 1 || a.n_comp > 32) return cudaErrorInvalidValue;
    const int nb = (a.n_comp + 7) / 8;
#define TPLS_MP(T)                                   \
    switch (nb) {                                    \
        case 1: return run_multiproj<T, 1>(a, sm_count, s); \
        case 2: return run_multiproj<T, 2>(a, sm_count, s); \
        case 3: return run_multiproj<T, 3>(a, sm_count, s); \
        default: return run_multiproj<T, 4>(a, sm_count, s); \
    }
    if (dtype == 0) {
        TPLS_MP(float)
    }
    TPLS_MP(double)
#undef TPLS_MP
}

// S[a][i] (column-major, n x R) from the slab partials of every coupled tensor:
//   r_a = (1/L) sum_l sum_slab part_l[slab][i][a];   t_a = r_a - c[a] - sum_{b<a} t_b G[b*R + a]
// flag |= 1 when a raw projection is NaN (the row has missing entries: the caller takes the masked path).
__global__ void __launch_bounds__(256) multiproj_finish_kernel(const __grid_constant__ MultiProjFinishArgs a) {
    pdl_prologue();
    bool bad = false;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.n_rows; i += (long long)gridDim.x * blockDim.x) {
        double t[32];
        for (int c = 0; c < a.n_comp; ++c) {
            double r = 0.0;
            for (int l = 0; l < a.n_tensors; ++l)
                for (int sb = 0; sb < a.n_slabs[l]; ++sb) r += a.part[l][((size_t)sb * a.n_rows + i) * a.n_comp + c];
            r /= (double)a.n_tensors;
            bad |= !(r == r);
            double v = r - a.c[c];
            for (int b = 0; b < c; ++b) v = fma(-t[b], a.G[b * a.n_comp + c], v);
            t[c] = v;
            a.S[(size_t)c * a.n_rows + i] = v;
        }
    }
    if (__syncthreads_or(bad) && threadIdx.x == 0) atomicOr(a.flag, 1);
}

cudaError_t launch_multiproj_finish(const MultiProjFinishArgs& a, cudaStream_t s) {
    if (a.n_comp > 32) return cudaErrorInvalidValue;
    const int blocks = (int)std::max<long long>(1, std::min<long long>(1184, (a.n_rows + 255) / 256));
    launch_k(multiproj_finish_kernel, dim3(blocks), dim3(256), 0, s, a);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// X_hat[i, c] = mean[c] + sum_a T[i*n_comp + a] * w[a*w_pitch + c]     (fp64 out, row pitch p)
// A thread owns two adjacent columns and walks the rows of its CTA's block; the rows' scores sit in shared memory.
// ---------------------------------------------------------------------------
constexpr int kRecRows = 32;

__global__ void __launch_bounds__(256) reconstruct_kernel(const __grid_constant__ ReconstructArgs a) {
    pdl_prologue();
    __shared__ double ts[kRecRows][33];
    const long long r0 = (long long)blockIdx.x * kRecRows;
    const int rows = (int)min((long long)kRecRows, a.n_rows - r0);
    for (int e = threadIdx.x; e < kRecRows * a.n_comp; e += blockDim.x) {
        const int c = e / kRecRows, r = e - c * kRecRows;
        ts[r][c] = r < rows ? a.T[(size_t)(r0 + r) * a.n_comp + c] : 0.0;
    }
    __syncthreads();
    for (int c = threadIdx.x * 2; c < a.p; c += blockDim.x * 2) {
        const bool two = c + 1 < a.p;
        const double m0 = a.mean != nullptr ? a.mean[c] : 0.0, m1 = (two && a.mean != nullptr) ? a.mean[c + 1] : 0.0;
        for (int rb = 0; rb < rows; rb += 8) {
            double acc0[8], acc1[8];
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                acc0[r] = m0;
                acc1[r] = m1;
            }
            for (int k = 0; k < a.n_comp; ++k) {
                const double w0 = __ldg(a.w + (size_t)k * a.w_pitch + c);
                const double w1 = two ? __ldg(a.w + (size_t)k * a.w_pitch + c + 1) : 0.0;
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    const double t = ts[(rb + r) & (kRecRows - 1)][k];
                    acc0[r] = fma(t, w0, acc0[r]);
                    acc1[r] = fma(t, w1, acc1[r]);
                }
            }
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                if (rb + r < rows) {
                    double* o = a.out + (size_t)(r0 + rb + r) * a.p + c;
                    o[0] = acc0[r];
                    if (two) o[1] = acc1[r];
                }
            }
        }
    }
}

cudaError_t launch_reconstruct(const ReconstructArgs& a, cudaStream_t s) {
    if (a.n_comp < 1 || a.n_comp > 32) return cudaErrorInvalidValue;
    const long long blocks = (a.n_rows + kRecRows - 1) / kRecRows;
    if (blocks > 0x7fffffffll) return cudaErrorInvalidValue;
    launch_k(reconstruct_kernel, dim3((unsigned)blocks), dim3(256), 0, s, a);
    return cudaGetLastError();
}

}  // namespace tpls
