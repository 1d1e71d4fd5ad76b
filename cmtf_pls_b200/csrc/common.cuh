// Shared device/host helpers for the tensor-PLS kernels (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace tpls {

// ---------------------------------------------------------------------------
// Device-resident control block of one NIPALS component (SURVEY.md §7 hard part
// 2): every kernel of inner trip k returns at once when the component already
// converged at an earlier trip, so the host may enqueue trips ahead of the
// convergence test without a sync, and the trailing contraction of the last trip
// of a device-resident (graph WHILE) loop does no work.
// ---------------------------------------------------------------------------
struct Ctrl {
    int done_trip;    // -1 while iterating, else the trip index that met the stop test
    int trips_taken;  // trips that did real work so far
    double last_d2;   // ||u_old - u_new||^2 of the last executed trip (global after allreduce)
    int trip;         // index of the trip in flight; advanced by the kernel that takes the stop decision
    int stop;         // 1 once the component converged or ran max_iter trips: every later loop kernel returns at once
};

// The inner loop of a component is a fixed sequence of launches (the "trip body") whose arguments do not depend
// on the trip index, so that it can be the body of a CUDA-graph WHILE node or be enqueued ahead by the host.
// `trip` is kept in the argument blocks of the single-operator entry points but no longer read.
__device__ __forceinline__ bool trip_is_dead(const Ctrl* c, int /*trip*/) {
    if (c == nullptr) return false;
    return *reinterpret_cast<const volatile int*>(&c->stop) != 0;
}

// ---------------------------------------------------------------------------
// Programmatic dependent launch (PDL).  Every kernel of the library starts with pdl_prologue(): it lets the
// NEXT kernel in the stream be scheduled early (its CTAs then sit at their own griddepcontrol.wait instead of
// paying the launch latency after this grid has drained) and waits until the PREVIOUS grid has completed and
// flushed its writes.  Both instructions do nothing for a kernel launched without the PDL attribute; launch_k()
// sets it unless TPLS_PDL=0.  Nothing before the prologue may read or write global memory.
// ---------------------------------------------------------------------------
__device__ __forceinline__ void pdl_prologue() {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
}

bool pdl_enabled();  // small.cu: on unless TPLS_PDL=0 in the environment (read once)
// The next launch of this thread is a plain one even with PDL on (the first kernel after a graph WHILE node: a
// conditional node cannot be the source of a programmatic edge).
void pdl_hold_next();
bool pdl_take_hold();

// kernel<<<grid, block, smem, stream>>>(args...) with the optional PDL attribute
template <typename... KArgs, typename... Args>
cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = (pdl_enabled() && !pdl_take_hold()) ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ---------------------------------------------------------------------------
// mbarrier + 1-D bulk async copy (TMA engine; SASS: UBLKCP / SYNCS)
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}

__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "LAB_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra LAB_DONE_%=;\n"
        "bra LAB_WAIT_%=;\n"
        "LAB_DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// global -> shared bulk copy; bytes and both addresses must be multiples of 16.
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(smem_dst)),
        "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ---------------------------------------------------------------------------
// reductions
// ---------------------------------------------------------------------------
__device__ __forceinline__ double shfl_xor_d(double v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) v += shfl_xor_d(v, m);
    return v;
}

// sum over all threads of a block (blockDim.x multiple of 32, <= 1024); result valid in every thread
__device__ __forceinline__ double block_sum(double v, double* red /* >= 33 doubles of smem */) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    if (w == 0) {
        double t = lane < nw ? red[lane] : 0.0;
        t = warp_sum(t);
        if (lane == 0) red[32] = t;
    }
    __syncthreads();
    return red[32];
}

}  // namespace tpls
