// One-shot peer-memory all-reduce fused with the local second-stage reduction -- see xchg.cuh.
#include "xchg.cuh"
#include "stream_common.cuh"

#include <algorithm>

namespace tpls {

__device__ __forceinline__ void st_relaxed_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__device__ __forceinline__ double ld_relaxed_sys(const double* p) {
    double v;
    asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}

// chunk = 32 consecutive elements of the exchanged vector; 256 threads = 32 columns x 8 groups
__global__ void __launch_bounds__(256) xchg_kernel(const __grid_constant__ XchgArgs a) {
    pdl_prologue();
    __shared__ double fold[8][37];  // 8 x 37 >= 32 x 9: doubles as the [32][9] area of fold_narrow
    const unsigned long long t_start = (blockIdx.x == 0 && threadIdx.x == 0) ? global_ns() : 0ull;
    // the sequence number of THIS exchange: the counter is advanced by the last CTA to finish phase 1, i.e. after
    // every CTA of the launch has read it
    const unsigned long long seq = *reinterpret_cast<const volatile unsigned long long*>(a.seq_ctr) + 1ull;
    const int slot = (int)(seq & 1ull);
    const size_t base = (size_t)slot * a.cap;
    double* mine = a.data[a.rank] + base;
    const int cl = threadIdx.x & 31, q = threadIdx.x >> 5;
    const int n_chunks = (a.count + 31) >> 5;

    // ---- phase 1: fold the local partials of my chunks into the local slot ----
    if (a.n_sets == 0) {
        for (int chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x) {
            const int idx = chunk * 32 + cl;
            if (q == 0 && idx < a.count) mine[idx] = a.in[idx];
        }
    } else {
        const int n_fold = fold_chunks(a.sets, a.n_sets);
        for (int chunk = blockIdx.x; chunk < n_fold; chunk += gridDim.x) {
            int c_base = 0;
            const FoldSet& S = a.sets[fold_locate(a.sets, a.n_sets, chunk, &c_base)];
            if (S.n_cols <= 8) {  // narrow set (q = Y't, the stop norm): 32 part-groups per column
                const double t = fold_narrow(S, reinterpret_cast<double(*)[9]>(&fold[0][0]));
                if ((threadIdx.x >> 3) == 0 && (int)(threadIdx.x & 7) < S.n_cols) mine[S.off + (threadIdx.x & 7)] = t;
                continue;
            }
            const int c = c_base + cl;
            const double v = fold_share(S, c, q);
            __syncthreads();
            fold[q][cl] = v;
            __syncthreads();
            if (q == 0 && c < S.n_cols) {
                double t = 0.0;
#pragma unroll
                for (int k = 0; k < 8; ++k) t += fold[k][cl];
                mine[S.off + c] = t;
            }
        }
    }
    __shared__ int is_last;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();  // cumulative: the CTA's slot writes (ordered by the barrier) become visible first
        const unsigned int old = atomicAdd(a.done_ctr, 1u);
        is_last = old == gridDim.x - 1 ? 1 : 0;
        if (is_last) {
            atomicExch(a.done_ctr, 0u);
            *reinterpret_cast<volatile unsigned long long*>(a.seq_ctr) = seq;
            __threadfence_system();  // this thread saw every CTA's arrival: order that before the announcements below
        }
    }
    __syncthreads();
    // The last CTA to finish phase 1 announces `seq` in every rank's flag word -- thread r tells rank r, each after
    // its own system fence (fence + relaxed store = release; one thread storing to all W peers with release stores
    // paid one NVLink round trip per peer, one after the other).
    if (is_last && (int)threadIdx.x < a.world) {
        __threadfence_system();
        st_relaxed_sys(a.flags[threadIdx.x] + a.rank, seq);
    }
    // ---- phase 2: wait for every rank's announcement in the LOCAL flag words, thread r for rank r ----
    const unsigned long long t_wait = (blockIdx.x == 0 && threadIdx.x == 0) ? global_ns() : 0ull;
    if ((int)threadIdx.x < a.world) {
        const long long t0 = clock64();
        while (ld_acquire_sys(a.flags[a.rank] + threadIdx.x) < seq) {
            if (clock64() - t0 > 40000000000ll) {  // ~20 s: a peer died
                atomicExch(a.err, 1);
                break;
            }
        }
    }
    __syncthreads();
    if (blockIdx.x == 0 && threadIdx.x == 0 && a.diag != nullptr) {  // how long this rank waited for the slowest one
        const unsigned long long t_sync = global_ns();
        a.diag[0] += t_sync - t_wait;
        a.diag[1] += t_sync - t_start;
        a.diag[2] += 1ull;
    }
    for (int chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x) {
        const int idx = chunk * 32 + cl;
        double v = 0.0;
        if (idx < a.count)
            for (int r = q; r < a.world; r += 8) v += ld_relaxed_sys(a.data[r] + base + idx);
        __syncthreads();
        fold[q][cl] = v;
        __syncthreads();
        if (q == 0 && idx < a.count) {
            double t = 0.0;
#pragma unroll
            for (int k = 0; k < 8; ++k) t += fold[k][cl];
            a.out[idx] = t;
            if (a.do_stop && idx == 0 && !trip_is_dead(a.ctrl, 0)) ctrl_decide(a.ctrl, t, a.loop_end);
        }
    }
    if (a.do_qstop) {  // count <= 32: this is the only CTA and a.out is complete after the barrier
        __syncthreads();
        if (threadIdx.x == 0 && !trip_is_dead(a.ctrl, 0))
            normalize_q_stop_body(a.out, a.q_m, a.q_pitch, a.qcol, a.qvec, a.gram, a.q_prev, a.ctrl, a.loop_end);
    }
}

cudaError_t launch_xchg(const XchgArgs& a, cudaStream_t s) {
    const int n_chunks = std::max((a.count + 31) / 32, a.n_sets > 0 ? fold_chunks(a.sets, a.n_sets) : 0);
    if (a.do_qstop && (a.count > 32 || a.q_m > 8)) return cudaErrorInvalidValue;
    if (a.n_sets > kMaxFoldSets) return cudaErrorInvalidValue;
    const int blocks = std::max(1, std::min(444, n_chunks));  // all CTAs must be co-resident: 3 per SM is safe
    launch_k(xchg_kernel, dim3(blocks), dim3(256), 0, s, a);
    return cudaGetLastError();
}

}  // namespace tpls
