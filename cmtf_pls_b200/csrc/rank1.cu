// Single-CTA rank-1 step with on-device convergence -- see rank1.cuh.
//
//   1 mode : w = Z / ||Z||
//   >=2    : HOSVD start (leading left singular vector of every unfolding, sign
//            fixed so the largest-|entry| is positive, mode 0 carrying sigma),
//            then rank-1 ALS sweeps with tensorly's stopping rule
//            |err_{s-1} - err_s| < tol from the second sweep on, <= 100 sweeps,
//            factors renormalised at the end of every sweep that did not stop.
//
// The leading eigenvector of each (small) Gram matrix is found by repeated
// squaring of the normalised matrix (logarithmic in the spectral gap, no
// data-dependent trip count worth speaking of) and polished against the
// original Gram matrix.  Everything is fp64.
#include "rank1.cuh"

#include <algorithm>

namespace tpls {

namespace {

constexpr int NTH = kRank1Threads;

struct Geo {  // mode-k unfolding geometry
    int dk, mk, ik;
};

__device__ __forceinline__ Geo mode_geo(const Rank1Task& T, int k) {
    Geo g;
    g.dk = T.dims[k];
    g.ik = 1;
    for (int m = k + 1; m < T.nmodes; ++m) g.ik *= T.dims[m];
    g.mk = T.p / g.dk;
    return g;
}

// flat index of element (a, j) of the mode-k unfolding (tensorly convention: remaining modes in C order)
__device__ __forceinline__ int unf_index(const Geo& g, int a, int j) {
    const int o = j / g.ik;
    const int in = j - o * g.ik;
    return (o * g.dk + a) * g.ik + in;
}

// C[i][j] = sum_k Mt[k*ld + i] * Mt[k*ld + j], i,j < n; C has leading dimension n.
__device__ void syrk_t(double* C, const double* Mt, int krows, int n, int ld) {
    const int nt = (n + 3) >> 2;
    for (int t = threadIdx.x; t < nt * nt; t += NTH) {
        const int i0 = (t / nt) * 4, j0 = (t % nt) * 4;
        double acc[4][4];
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int r = 0; r < 4; ++r) acc[q][r] = 0.0;
        for (int k = 0; k < krows; ++k) {
            const double* row = Mt + (size_t)k * ld;
            double ai[4], bj[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                ai[q] = (i0 + q < n) ? row[i0 + q] : 0.0;
                bj[q] = (j0 + q < n) ? row[j0 + q] : 0.0;
            }
#pragma unroll
            for (int q = 0; q < 4; ++q)
#pragma unroll
                for (int r = 0; r < 4; ++r) acc[q][r] = fma(ai[q], bj[r], acc[q][r]);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int r = 0; r < 4; ++r)
                if (i0 + q < n && j0 + r < n) C[(size_t)(i0 + q) * n + j0 + r] = acc[q][r];
    }
    __syncthreads();
}

// y[i] = sum_j M[i*ld + j] * x[j]   (warp per row)
__device__ void matvec_rows(double* y, const double* M, const double* x, int rows, int cols, int ld) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int i = w; i < rows; i += NTH / 32) {
        double s = 0.0;
        for (int j = lane; j < cols; j += 32) s = fma(M[(size_t)i * ld + j], x[j], s);
        s = warp_sum(s);
        if (lane == 0) y[i] = s;
    }
    __syncthreads();
}

__device__ double vec_dot(const double* a, const double* b, int n, double* red) {
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += NTH) s = fma(a[i], b[i], s);
    return block_sum(s, red);
}

__device__ void vec_scale(double* a, double f, int n) {
    for (int i = threadIdx.x; i < n; i += NTH) a[i] *= f;
    __syncthreads();
}

__device__ void vec_div(double* a, double d, int n) {
    for (int i = threadIdx.x; i < n; i += NTH) a[i] /= d;
    __syncthreads();
}

// index of the first entry with the largest |a[i]|
__device__ int argmax_abs(const double* a, int n, double* red) {
    double bv = -1.0;
    int bi = 0x7fffffff;
    for (int i = threadIdx.x; i < n; i += NTH) {
        const double v = fabs(a[i]);
        if (v > bv) {
            bv = v;
            bi = i;
        }
    }
    for (int m = 16; m >= 1; m >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, bv, m);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, m);
        if (ov > bv || (ov == bv && oi < bi)) {
            bv = ov;
            bi = oi;
        }
    }
    int* ired = reinterpret_cast<int*>(red + 40);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) {
        red[w] = bv;
        ired[w] = bi;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int q = 1; q < NTH / 32; ++q)
            if (red[q] > bv || (red[q] == bv && ired[q] < bi)) {
                bv = red[q];
                bi = ired[q];
            }
        ired[NTH / 32] = bi;
    }
    __syncthreads();
    const int out = ired[NTH / 32];
    __syncthreads();
    return out;
}

// Leading eigenpair of the symmetric PSD matrix G (n x n, ld n).  A, B: n*n work buffers.
// v (n) receives the unit eigenvector; returns the eigenvalue.
__device__ double lead_eig(const double* G, double* A, double* B, double* v, double* tmp, int n, double* red) {
    const int nn = n * n;
    double s = 0.0;
    for (int i = threadIdx.x; i < nn; i += NTH) s = fma(G[i], G[i], s);
    const double fro = sqrt(block_sum(s, red));
    if (!(fro > 0.0)) {
        for (int i = threadIdx.x; i < n; i += NTH) v[i] = 0.0;
        __syncthreads();
        return 0.0;
    }
    for (int i = threadIdx.x; i < nn; i += NTH) A[i] = G[i] / fro;
    __syncthreads();
    for (int it = 0; it < 64; ++it) {
        syrk_t(B, A, n, n, n);  // B = A*A (A symmetric)
        s = 0.0;
        for (int i = threadIdx.x; i < nn; i += NTH) s = fma(B[i], B[i], s);
        const double nb = sqrt(block_sum(s, red));
        double d = 0.0;
        for (int i = threadIdx.x; i < nn; i += NTH) {
            const double b = B[i] / nb;
            const double e = b - A[i];
            B[i] = b;
            d = fma(e, e, d);
        }
        d = block_sum(d, red);
        double* sw = A;
        A = B;
        B = sw;
        if (sqrt(d) < 1e-13) break;
    }
    // A ~ v v^T: take the column with the largest diagonal entry
    for (int i = threadIdx.x; i < n; i += NTH) tmp[i] = A[(size_t)i * n + i];
    __syncthreads();
    const int bi = argmax_abs(tmp, n, red);
    for (int i = threadIdx.x; i < n; i += NTH) v[i] = A[(size_t)i * n + bi];
    __syncthreads();
    double nv = sqrt(vec_dot(v, v, n, red));
    vec_div(v, nv, n);
    // polish against the original matrix
    for (int q = 0; q < 2; ++q) {
        matvec_rows(tmp, G, v, n, n, n);
        nv = sqrt(vec_dot(tmp, tmp, n, red));
        for (int i = threadIdx.x; i < n; i += NTH) v[i] = tmp[i] / nv;
        __syncthreads();
    }
    matvec_rows(tmp, G, v, n, n, n);
    return vec_dot(tmp, v, n, red);
}

// f[a] = sum_j Zk(a, j) * x[j]
__device__ void unf_matvec(double* f, const double* zs, const Geo& g, const double* x) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int a = w; a < g.dk; a += NTH / 32) {
        double s = 0.0;
        for (int j = lane; j < g.mk; j += 32) s = fma(zs[unf_index(g, a, j)], x[j], s);
        s = warp_sum(s);
        if (lane == 0) f[a] = s;
    }
    __syncthreads();
}

__device__ void flip_to_positive_peak(double* f, int n, double* red) {
    const int i = argmax_abs(f, n, red);
    if (f[i] < 0.0) {
        __syncthreads();
        for (int q = threadIdx.x; q < n; q += NTH) f[q] = -f[q];
    }
    __syncthreads();
}

// kr[j] = prod_{m != k} f_m[i_m(j)], j enumerating the other modes in C order
__device__ void other_modes_product(double* kr, const Rank1Task& T, int k, double* const* f, int mk) {
    for (int j = threadIdx.x; j < mk; j += NTH) {
        int rem = j;
        double pr = 1.0;
        for (int m = T.nmodes - 1; m >= 0; --m) {
            if (m == k) continue;
            const int d = T.dims[m];
            const int i = rem % d;
            rem /= d;
            pr *= f[m][i];
        }
        kr[j] = pr;
    }
    __syncthreads();
}

__device__ void rank1_task(const Rank1Task& T, double tol, int normalize_on_break, double* ws) {
    __shared__ double red[64];
    __shared__ double* f[kMaxZModes];
    const int p = T.p;
    const int nm = T.nmodes;

    // workspace carve-up (doubles)
    double* zs = ws;
    double* mt = zs + p;
    double* G = mt + p;
    const size_t n2 = (size_t)T.nmax * T.nmax;
    double* A = G + n2;
    double* B = A + n2;
    double* fac = B + n2;
    int sumd = 0, maxd = 0;
    for (int m = 0; m < nm; ++m) {
        sumd += T.dims[m];
        maxd = max(maxd, T.dims[m]);
    }
    double* tmp = fac + sumd;
    double* tmp2 = tmp + max(maxd, T.nmax);
    if (threadIdx.x == 0) {
        int off = 0;
        for (int m = 0; m < nm; ++m) {
            f[m] = fac + off;
            off += T.dims[m];
        }
    }

    // Z, with the observed-count rescaling of missingvals.py:18 when masked
    double s = 0.0;
    for (int i = threadIdx.x; i < p; i += NTH) {
        double z = T.z[i];
        if (T.colcnt != nullptr) {
            const double c = T.colcnt[i];
            z = c > 0.0 ? z / c * T.n_total : 0.0;
        }
        zs[i] = z;
        s = fma(z, z, s);
    }
    const double normz2 = block_sum(s, red);
    const double normz = sqrt(normz2);

    if (nm == 1) {
        for (int i = threadIdx.x; i < T.pitch; i += NTH) {
            const double w = i < p ? zs[i] / normz : 0.0;
            if (i < p) T.w[0][i] = w;
            T.wkron[i] = w;
        }
        if (threadIdx.x == 0 && T.sweeps) *T.sweeps = 0;
        return;
    }

    // ---- HOSVD start ----
    double weight = 1.0;
    if (nm == 2) {
        // one eigenproblem on the short side; the other vector follows from Z
        const int ks = T.dims[0] <= T.dims[1] ? 0 : 1;
        const int ko = 1 - ks;
        const Geo gs = mode_geo(T, ks), go = mode_geo(T, ko);
        // Mt[j][a] = Z_(ks)(a, j): for ks == 1 that is Z itself, for ks == 0 its transpose
        if (ks == 1) {
            syrk_t(G, zs, gs.mk, gs.dk, gs.dk);
        } else {
            for (int i = threadIdx.x; i < p; i += NTH) {
                const int a = i / gs.mk, j = i - a * gs.mk;
                mt[(size_t)j * gs.dk + a] = zs[i];
            }
            __syncthreads();
            syrk_t(G, mt, gs.mk, gs.dk, gs.dk);
        }
        lead_eig(G, A, B, f[ks], tmp, gs.dk, red);
        unf_matvec(f[ko], zs, go, f[ks]);  // Z_(ko) f_ks = sigma * other vector
        const double sigma = sqrt(vec_dot(f[ko], f[ko], go.dk, red));
        vec_div(f[ko], sigma, go.dk);
        weight = sigma;
        flip_to_positive_peak(f[0], T.dims[0], red);
        flip_to_positive_peak(f[1], T.dims[1], red);
    } else {
        for (int k = 0; k < nm; ++k) {
            const Geo g = mode_geo(T, k);
            double lam;
            if (g.dk <= g.mk) {
                for (int i = threadIdx.x; i < p; i += NTH) {
                    const int a = i / g.mk, j = i - a * g.mk;
                    mt[(size_t)j * g.dk + a] = zs[unf_index(g, a, j)];
                }
                __syncthreads();
                syrk_t(G, mt, g.mk, g.dk, g.dk);
                lam = lead_eig(G, A, B, f[k], tmp, g.dk, red);
                if (k == 0) weight = sqrt(lam);
            } else {
                for (int i = threadIdx.x; i < p; i += NTH) {
                    const int a = i / g.mk, j = i - a * g.mk;
                    mt[i] = zs[unf_index(g, a, j)];
                }
                __syncthreads();
                syrk_t(G, mt, g.dk, g.mk, g.mk);
                lead_eig(G, A, B, tmp2, tmp, g.mk, red);
                unf_matvec(f[k], zs, g, tmp2);
                const double sigma = sqrt(vec_dot(f[k], f[k], g.dk, red));
                vec_div(f[k], sigma, g.dk);
                if (k == 0) weight = sigma;
            }
            flip_to_positive_peak(f[k], g.dk, red);
        }
    }

    // ---- ALS sweeps (tensorly parafac, rank 1) ----
    double nrm2[kMaxZModes];
    for (int m = 0; m < nm; ++m) nrm2[m] = vec_dot(f[m], f[m], T.dims[m], red);
    double err_prev = 0.0;
    int sweeps = 0;
    double* kr = mt;
    for (int it = 0; it < 100; ++it) {
        ++sweeps;
        double iprod = 0.0;
        for (int k = 0; k < nm; ++k) {
            const Geo g = mode_geo(T, k);
            other_modes_product(kr, T, k, f, g.mk);
            unf_matvec(tmp, zs, g, kr);
            double gram = weight * weight;
            for (int m = 0; m < nm; ++m)
                if (m != k) gram *= nrm2[m];
            // factor = (weight * Z x_others f) / gram
            for (int i = threadIdx.x; i < g.dk; i += NTH) {
                const double mt_i = tmp[i] * weight;
                tmp[i] = mt_i;
                f[k][i] = mt_i / gram;
            }
            __syncthreads();
            nrm2[k] = vec_dot(f[k], f[k], g.dk, red);
            if (k == nm - 1) iprod = vec_dot(tmp, f[k], g.dk, red);
        }
        double fn2 = weight * weight;
        for (int m = 0; m < nm; ++m) fn2 *= nrm2[m];
        const double err = sqrt(fabs(normz2 + fn2 - 2.0 * iprod)) / normz;
        const bool stop = it >= 1 && fabs(err_prev - err) < tol;
        err_prev = err;
        if (stop && !normalize_on_break) break;
        // cp_normalize: weights into factor 0, then every column norm back into the weights
        vec_scale(f[0], weight, T.dims[0]);
        nrm2[0] *= weight * weight;
        weight = 1.0;
        for (int m = 0; m < nm; ++m) {
            const double sc = sqrt(vec_dot(f[m], f[m], T.dims[m], red));
            weight *= sc;
            vec_div(f[m], sc == 0.0 ? 1.0 : sc, T.dims[m]);
            nrm2[m] = vec_dot(f[m], f[m], T.dims[m], red);
        }
        if (stop) break;
    }

    // ---- publish ----
    for (int m = 0; m < nm; ++m)
        for (int i = threadIdx.x; i < T.dims[m]; i += NTH) T.w[m][i] = f[m][i];
    for (int i = threadIdx.x; i < T.pitch; i += NTH) {
        double pr = 0.0;
        if (i < p) {
            int rem = i;
            pr = 1.0;
            // kron(w_0, w_1, ...) built the way numpy.kron nests it: ((w0 * w1) * w2) ...
            int idx[kMaxZModes];
            for (int m = nm - 1; m >= 0; --m) {
                idx[m] = rem % T.dims[m];
                rem /= T.dims[m];
            }
            pr = f[0][idx[0]];
            for (int m = 1; m < nm; ++m) pr *= f[m][idx[m]];
        }
        T.wkron[i] = pr;
    }
    if (threadIdx.x == 0 && T.sweeps) *T.sweeps = sweeps;
}

__global__ void __launch_bounds__(kRank1Threads, 1) rank1_kernel(const __grid_constant__ Rank1Args a) {
    if (trip_is_dead(a.ctrl, a.trip)) return;
    extern __shared__ __align__(16) double dyn[];
    const Rank1Task& T = a.t[blockIdx.x];
    rank1_task(T, a.tol, a.normalize_on_break, T.use_smem ? dyn : T.scratch);
}

}  // namespace

size_t rank1_workspace_doubles(int nmodes, const int* dims, int* nmax_out) {
    long long p = 1;
    int sumd = 0, maxd = 0;
    for (int m = 0; m < nmodes; ++m) {
        p *= dims[m];
        sumd += dims[m];
        maxd = std::max(maxd, dims[m]);
    }
    int nmax = 1;
    if (nmodes == 2) {
        nmax = std::min(dims[0], dims[1]);
    } else if (nmodes >= 3) {
        for (int m = 0; m < nmodes; ++m) nmax = std::max<long long>(nmax, std::min<long long>(dims[m], p / dims[m]));
    }
    if (nmax_out) *nmax_out = nmax;
    return (size_t)(2 * p + 3ll * nmax * nmax + sumd + 2 * std::max(maxd, nmax) + 16);
}

cudaError_t launch_rank1(const Rank1Args& a, size_t smem_bytes, cudaStream_t s) {
    cudaError_t e = cudaFuncSetAttribute(rank1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
    if (e != cudaSuccess) return e;
    rank1_kernel<<<a.n_tasks, kRank1Threads, smem_bytes, s>>>(a);
    return cudaGetLastError();
}

}  // namespace tpls
