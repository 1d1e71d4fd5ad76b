// Micro-benchmarks that size the single-CTA rank-1 kernel: latency / issue cost of the fp64 and
// synchronisation primitives it is made of, measured with clock64 on one SM.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/micro/fp64_lat.cu -o /tmp/fp64_lat && /tmp/fp64_lat
#include <cstdio>
#include <cuda_runtime.h>

#define N 512

__global__ void k_dfma_chain(double* out, long long* cyc, double a, double b) {
    double x = a;
    long long t0 = clock64();
#pragma unroll
    for (int i = 0; i < N; ++i) x = fma(x, b, a);
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
    out[threadIdx.x] = x;
}

__global__ void k_dadd_chain(double* out, long long* cyc, double a, double b) {
    double x = a;
    long long t0 = clock64();
#pragma unroll
    for (int i = 0; i < N; ++i) x = x + b;
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
    out[threadIdx.x] = x;
}

// 16 independent chains per thread: issue-bound.  iters is a run-time value so that nothing can be hoisted.
__global__ void k_dfma_tput(double* out, long long* cyc, double a, double b, int iters) {
    double x[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) x[j] = a + j;
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < iters; ++i)
#pragma unroll
        for (int j = 0; j < 16; ++j) x[j] = fma(x[j], b, a);
    long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
    double s = 0;
#pragma unroll
    for (int j = 0; j < 16; ++j) s += x[j];
    out[threadIdx.x] = s;
}

// the inner loop of the symmetric rank-k update: 4 x LDS.128 + 16 DFMA per step
__global__ void k_syrk_loop(double* out, long long* cyc, int iters, int ld) {
    extern __shared__ double sm[];
    for (int i = threadIdx.x; i < 64 * ld; i += blockDim.x) sm[i] = 1e-3 * i;
    __syncthreads();
    const int t = threadIdx.x >> 1, ks = threadIdx.x & 1;
    const int ti = (t / 11) % 16, tj = t % 16;
    double acc[4][4];
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int r = 0; r < 4; ++r) acc[q][r] = 0.0;
    const double* pa = sm + 4 * ti;
    const double* pb = sm + 4 * tj;
    long long t0 = clock64();
#pragma unroll 2
    for (int k = ks; k < iters; k += 2) {
        const int kk = k & 63;
        const double2 a01 = *reinterpret_cast<const double2*>(pa + kk * ld);
        const double2 a23 = *reinterpret_cast<const double2*>(pa + kk * ld + 2);
        const double2 b01 = *reinterpret_cast<const double2*>(pb + kk * ld);
        const double2 b23 = *reinterpret_cast<const double2*>(pb + kk * ld + 2);
        const double ai[4] = {a01.x, a01.y, a23.x, a23.y};
        const double bj[4] = {b01.x, b01.y, b23.x, b23.y};
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int r = 0; r < 4; ++r) acc[q][r] = fma(ai[q], bj[r], acc[q][r]);
    }
    long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
    double s = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int r = 0; r < 4; ++r) s += acc[q][r];
    out[threadIdx.x] = s;
}

// fp64 tensor-core path: mma.sync m8n8k4, NI independent accumulator pairs per warp
template <int NI>
__global__ void k_dmma_tput(double* out, long long* cyc, double a, double b, int iters) {
    double c0[NI], c1[NI];
#pragma unroll
    for (int j = 0; j < NI; ++j) {
        c0[j] = j;
        c1[j] = -j;
    }
    double av = a + threadIdx.x, bv = b - threadIdx.x;
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < iters; ++i)
#pragma unroll
        for (int j = 0; j < NI; ++j)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c0[j]), "+d"(c1[j])
                         : "d"(av), "d"(bv));
    long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
    double s = 0;
#pragma unroll
    for (int j = 0; j < NI; ++j) s += c0[j] + c1[j];
    out[threadIdx.x] = s;
}

__global__ void k_shfl_chain(double* out, long long* cyc, double a) {
    double x = a + threadIdx.x;
    long long t0 = clock64();
#pragma unroll
    for (int i = 0; i < 128; ++i) x += __shfl_xor_sync(0xffffffffu, x, 1 + (i & 15));
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
    out[threadIdx.x] = x;
}

__global__ void k_barrier(double* out, long long* cyc) {
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 256; ++i) __syncthreads();
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
    out[threadIdx.x] = 1.0;
}

__global__ void k_div_chain(double* out, long long* cyc, double a, double b) {
    double x = a;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 64; ++i) x = b / x + a;
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
    out[threadIdx.x] = x;
}

__global__ void k_sqrt_chain(double* out, long long* cyc, double a, double b) {
    double x = a;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 64; ++i) x = sqrt(x) + b;
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
    out[threadIdx.x] = x;
}

__global__ void k_lds_chain(double* out, long long* cyc) {
    __shared__ int nxt[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) nxt[i] = (i * 17 + 5) & 1023;
    __syncthreads();
    int p = threadIdx.x;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 256; ++i) p = nxt[p];
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
    out[threadIdx.x] = p;
}

// block-wide sum the way the rank-1 kernel does it: warp shuffle tree, one barrier, every thread folds the partials
__global__ void k_bsum(double* out, long long* cyc, double a) {
    __shared__ double red[2][32];
    double v = a + threadIdx.x;
    const int nw = blockDim.x >> 5;
    int flip = 0;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 64; ++i) {
        double w = v;
        for (int m = 16; m >= 1; m >>= 1) w += __shfl_xor_sync(0xffffffffu, w, m);
        if ((threadIdx.x & 31) == 0) red[flip][threadIdx.x >> 5] = w;
        __syncthreads();
        double t = 0.0;
        for (int q = 0; q < nw; ++q) t += red[flip][q];
        flip ^= 1;
        v = t * 1e-3;
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
    out[threadIdx.x] = v;
}

int main() {
    double* out;
    long long* cyc;
    cudaMalloc(&out, 1024 * sizeof(double));
    cudaMalloc(&cyc, 8 * sizeof(long long));
    long long h;
#define RUN(name, per, threads, ...)                                            \
    name<<<1, threads>>>(__VA_ARGS__);                                          \
    name<<<1, threads>>>(__VA_ARGS__);                                          \
    cudaDeviceSynchronize();                                                    \
    cudaMemcpy(&h, cyc, sizeof h, cudaMemcpyDeviceToHost);                      \
    printf("%-14s threads %4d : %8.1f cycles per %s\n", #name, threads, (double)h / (per), #per);
    RUN(k_dfma_chain, N, 32, out, cyc, 1.0, 0.5)
    RUN(k_dadd_chain, N, 32, out, cyc, 1.0, 0.5)
    for (int th : {32, 64, 128, 256, 512}) { RUN(k_dfma_tput, 1000 * 16, th, out, cyc, 1.0, 0.5, 1000) }
    for (int th : {32, 128, 256, 512}) { RUN(k_dmma_tput<8>, 1000 * 8, th, out, cyc, 1.0, 0.5, 1000) }
    RUN(k_dmma_tput<1>, 1000, 32, out, cyc, 1.0, 0.5, 1000)
    cudaFuncSetAttribute(k_syrk_loop, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 72 * 8);
    for (int th : {32, 128, 256, 288, 512}) {
        k_syrk_loop<<<1, th, 64 * 72 * 8>>>(out, cyc, 2048, 66);
        cudaDeviceSynchronize();
        cudaMemcpy(&h, cyc, sizeof h, cudaMemcpyDeviceToHost);
        printf("k_syrk_loop    threads %4d : %8.1f cycles per step (4 LDS.128 + 16 DFMA), ld 66\n", th, (double)h / 1024);
    }
    RUN(k_shfl_chain, 128, 32, out, cyc, 1.0)
    for (int th : {128, 256, 512}) { RUN(k_barrier, 256, th, out, cyc) }
    RUN(k_div_chain, 64, 32, out, cyc, 1.5, 0.7)
    RUN(k_sqrt_chain, 64, 32, out, cyc, 1.5, 0.7)
    RUN(k_lds_chain, 256, 32, out, cyc)
    for (int th : {160, 512}) { RUN(k_bsum, 64, th, out, cyc, 1.0) }
    printf("err: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
