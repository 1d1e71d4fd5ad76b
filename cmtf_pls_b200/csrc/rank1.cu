// Single-CTA rank-1 step with on-device convergence -- see rank1.cuh.
//
//   1 mode : w = Z / ||Z||
//   >=2    : HOSVD start (leading left singular vector of every unfolding, sign
//            fixed so the largest-|entry| is positive, mode 0 carrying sigma),
//            then rank-1 ALS sweeps with tensorly's stopping rule
//            |err_{s-1} - err_s| < tol from the second sweep on, <= 100 sweeps,
//            factors renormalised at the end of every sweep that did not stop.
//
// Leading eigenvector of each (small) Gram matrix: repeated squaring of the
// trace-normalised matrix, A <- A*A / tr(A*A).  tr(A*A) -> 1 exactly when A is
// rank one, so "1 - trace < 1e-7" means the iterate just formed has a
// second-to-first eigenvalue ratio below ~1e-14: logarithmic in the spectral
// gap, one tiny reduction per step, then two polish steps against the original
// Gram matrix.  Everything is fp64.  Matrices are stored with leading dimension
// rounded up to 4 and zero pads so that the 4x4 register tiles need no bounds
// checks and load 16 bytes at a time.
#include "rank1.cuh"

#include <algorithm>

namespace tpls {

namespace {

constexpr int NTH = kRank1Threads;
constexpr int NWARP = NTH / 32;

__host__ __device__ inline int up4(int v) { return (v + 3) & ~3; }

// block-wide sums with ONE barrier each: warp partials go to one of two alternating
// buffers, every thread then folds the 16 partials itself
struct Blk {
    double* red;  // [2][2*NWARP]
    int* ired;    // [2][NWARP]
    int flip;
};

__device__ __forceinline__ double bsum(double v, Blk& b) {
    v = warp_sum(v);
    double* r = b.red + b.flip * 2 * NWARP;
    if ((threadIdx.x & 31) == 0) r[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
#pragma unroll
    for (int q = 0; q < NWARP; ++q) t += r[q];
    b.flip ^= 1;
    return t;
}

__device__ __forceinline__ void bsum2(double& v1, double& v2, Blk& b) {
    v1 = warp_sum(v1);
    v2 = warp_sum(v2);
    double* r = b.red + b.flip * 2 * NWARP;
    if ((threadIdx.x & 31) == 0) {
        r[threadIdx.x >> 5] = v1;
        r[NWARP + (threadIdx.x >> 5)] = v2;
    }
    __syncthreads();
    double t1 = 0.0, t2 = 0.0;
#pragma unroll
    for (int q = 0; q < NWARP; ++q) {
        t1 += r[q];
        t2 += r[NWARP + q];
    }
    b.flip ^= 1;
    v1 = t1;
    v2 = t2;
}

// index of the first entry with the largest |a[i]|
__device__ __forceinline__ int argmax_abs(const double* a, int n, Blk& b) {
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
    double* r = b.red + b.flip * 2 * NWARP;
    int* ir = b.ired + b.flip * NWARP;
    if ((threadIdx.x & 31) == 0) {
        r[threadIdx.x >> 5] = bv;
        ir[threadIdx.x >> 5] = bi;
    }
    __syncthreads();
    bv = r[0];
    bi = ir[0];
#pragma unroll
    for (int q = 1; q < NWARP; ++q) {
        const double ov = r[q];
        const int oi = ir[q];
        if (ov > bv || (ov == bv && oi < bi)) {
            bv = ov;
            bi = oi;
        }
    }
    b.flip ^= 1;
    return bi;
}

// C[i*n4 + j] = scale * sum_k Mt[k*n4 + i] * Mt[k*n4 + j]   (n4 % 4 == 0, pads of Mt are zero)
__device__ __forceinline__ void syrk_fast(double* __restrict__ C, const double* __restrict__ Mt, int krows, int n4, double scale) {
    // a thread owns rows i0..i0+3 (contiguous: two broadcast 16-byte loads) and columns tj, tj+nt, tj+2nt,
    // tj+3nt (strided: the 16 lanes of a tile row read 16 consecutive doubles -- conflict-free)
    const int nt = n4 >> 2;
    for (int t = threadIdx.x; t < nt * nt; t += NTH) {
        const int i0 = (t / nt) << 2, tj = t % nt;
        double acc[4][4];
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int r = 0; r < 4; ++r) acc[q][r] = 0.0;
        const double* pa = Mt + i0;
        const double* pb = Mt + tj;
#pragma unroll 4
        for (int k = 0; k < krows; ++k) {
            const double2 a01 = *reinterpret_cast<const double2*>(pa + (size_t)k * n4);
            const double2 a23 = *reinterpret_cast<const double2*>(pa + (size_t)k * n4 + 2);
            const double ai[4] = {a01.x, a01.y, a23.x, a23.y};
            double bj[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) bj[r] = pb[(size_t)k * n4 + r * nt];
#pragma unroll
            for (int q = 0; q < 4; ++q)
#pragma unroll
                for (int r = 0; r < 4; ++r) acc[q][r] = fma(ai[q], bj[r], acc[q][r]);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int r = 0; r < 4; ++r) C[(size_t)(i0 + q) * n4 + tj + r * nt] = acc[q][r] * scale;
    }
    __syncthreads();
}

// y[i] = sum_{j < cols} M[i*ld + j] * x[j], i < rows: 8 lanes per row
__device__ __forceinline__ void matvec8(double* y, const double* M, int ld, const double* x, int rows, int cols) {
    const int part = threadIdx.x & 7, r0 = threadIdx.x >> 3;
    for (int rb = 0; rb < rows; rb += NTH / 8) {
        const int i = rb + r0;
        double s = 0.0;
        if (i < rows)
            for (int j = part; j < cols; j += 8) s = fma(M[(size_t)i * ld + j], x[j], s);
        s += shfl_xor_d(s, 4);
        s += shfl_xor_d(s, 2);
        s += shfl_xor_d(s, 1);
        if (i < rows && part == 0) y[i] = s;
    }
    __syncthreads();
}

__device__ __forceinline__ double vec_dot(const double* a, const double* b, int n, Blk& blk) {
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += NTH) s = fma(a[i], b[i], s);
    return bsum(s, blk);
}

__device__ __forceinline__ void vec_div(double* a, double d, int n) {
    for (int i = threadIdx.x; i < n; i += NTH) a[i] /= d;
    __syncthreads();
}

// dst = src / ||src||
__device__ __forceinline__ void normalize_into(double* dst, const double* src, int n, Blk& blk) {
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += NTH) s = fma(src[i], src[i], s);
    const double nv = sqrt(bsum(s, blk));
    for (int i = threadIdx.x; i < n; i += NTH) dst[i] = src[i] / nv;
    __syncthreads();
}

// Leading eigenpair of the symmetric PSD matrix G (n x n stored n4 x n4 with zero pads).
// A, B: n4*n4 work buffers.  v (n) receives the unit eigenvector; returns the eigenvalue.
__device__ __forceinline__ double lead_eig(const double* G, double* A, double* B, double* v, double* tmp, int n, Blk& blk, int polish,
                                           bool want_lambda) {
    const int n4 = up4(n);
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += NTH) s += G[(size_t)i * n4 + i];
    const double tr = bsum(s, blk);
    if (!(tr > 0.0)) {
        for (int i = threadIdx.x; i < n; i += NTH) v[i] = 0.0;
        __syncthreads();
        return 0.0;
    }
    const double* src = G;
    double* dst = A;
    double* oth = B;
    double scale = (1.0 / tr) * (1.0 / tr);
    for (int it = 0; it < 64; ++it) {
        syrk_fast(dst, src, n4, n4, scale);  // dst = (src / tr(src))^2
        s = 0.0;
        for (int i = threadIdx.x; i < n; i += NTH) s += dst[(size_t)i * n4 + i];
        const double tau = bsum(s, blk);  // sum of squared normalised eigenvalues, -> 1 at rank one
        src = dst;
        double* sw = dst;
        dst = oth;
        oth = sw;
        if (1.0 - tau < 1e-7) break;
        scale = (1.0 / tau) * (1.0 / tau);
    }
    // src ~ v v^T: take the column with the largest diagonal entry
    for (int i = threadIdx.x; i < n; i += NTH) tmp[i] = src[(size_t)i * n4 + i];
    __syncthreads();
    const int bi = argmax_abs(tmp, n, blk);
    for (int i = threadIdx.x; i < n; i += NTH) v[i] = src[(size_t)i * n4 + bi];
    __syncthreads();
    double nv = sqrt(vec_dot(v, v, n, blk));
    vec_div(v, nv, n);
    // polish against the original matrix
    for (int q = 0; q < polish; ++q) {
        matvec8(tmp, G, n4, v, n, n);
        nv = sqrt(vec_dot(tmp, tmp, n, blk));
        for (int i = threadIdx.x; i < n; i += NTH) v[i] = tmp[i] / nv;
        __syncthreads();
    }
    if (!want_lambda) return 0.0;
    matvec8(tmp, G, n4, v, n, n);
    return vec_dot(tmp, v, n, blk);
}

__device__ __forceinline__ void flip_to_positive_peak(double* f, int n, Blk& blk) {
    const int i = argmax_abs(f, n, blk);
    const bool neg = f[i] < 0.0;
    __syncthreads();
    if (neg)
        for (int q = threadIdx.x; q < n; q += NTH) f[q] = -f[q];
    __syncthreads();
}

// the factor vectors live back to back in the workspace; offsets stay in registers so that the compiler
// keeps seeing workspace-derived (shared-memory) addresses
struct Facs {
    double* base;
    int off[kMaxZModes];
    __device__ __forceinline__ double* operator[](int m) const { return base + off[m]; }
};

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

// f[a] = sum_j Zk(a, j) * x[j]   (general mode, strided)
__device__ __forceinline__ void unf_matvec(double* f, const double* zs, const Geo& g, const double* x) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int a = w; a < g.dk; a += NWARP) {
        double s = 0.0;
        for (int j = lane; j < g.mk; j += 32) s = fma(zs[unf_index(g, a, j)], x[j], s);
        s = warp_sum(s);
        if (lane == 0) f[a] = s;
    }
    __syncthreads();
}

// kr[j] = prod_{m != k} f_m[i_m(j)], j enumerating the other modes in C order
__device__ __forceinline__ void other_modes_product(double* kr, const Rank1Task& T, int k, const Facs& f, int mk) {
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

__device__ __forceinline__ void rank1_task(const Rank1Task& T, double tol, int normalize_on_break, double* ws) {
    __shared__ double red[4 * NWARP];
    __shared__ int ired[2 * NWARP];
    Blk blk{red, ired, 0};
    const int p = T.p;
    const int nm = T.nmodes;
#define TPLS_STAMP(i) \
    if (T.stamps != nullptr && threadIdx.x == 0) T.stamps[i] = clock64()

    // workspace carve-up (doubles); see rank1_workspace_doubles
    int sumd = 0, maxd = 0;
    for (int m = 0; m < nm; ++m) {
        sumd += T.dims[m];
        maxd = max(maxd, T.dims[m]);
    }
    const int n4max = up4(T.nmax);
    const size_t n2 = (size_t)n4max * n4max;
    double* zs = ws;             // Z (nm == 2: row-padded d0 x up4(d1))
    double* mt = zs + T.zs_len;  // unfolding copy / Z^T padded / Khatri-Rao vector
    double* G = mt + T.mt_len;
    double* A = G + n2;
    double* B = A + n2;
    double* fac = B + n2;
    double* tmp = fac + up4(sumd);
    double* tmp2 = tmp + up4(max(maxd, T.nmax));
    Facs f;
    f.base = fac;
    {
        int off = 0;
        for (int m = 0; m < kMaxZModes; ++m) {
            f.off[m] = off;
            if (m < nm) off += T.dims[m];
        }
    }

    TPLS_STAMP(0);
    if (nm == 1) {
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
        const double normz = sqrt(bsum(s, blk));
        for (int i = threadIdx.x; i < T.pitch; i += NTH) {
            const double w = i < p ? zs[i] / normz : 0.0;
            if (i < p) T.w[0][i] = w;
            T.wkron[i] = w;
        }
        if (threadIdx.x == 0 && T.sweeps) *T.sweeps = 0;
        return;
    }

    // ---- load Z (with the observed-count rescaling of missingvals.py:18 when masked) ----
    const int d0 = T.dims[0], d1 = nm == 2 ? T.dims[1] : 0;
    const int ld0 = up4(d0), ld1 = up4(d1);
    double s = 0.0;
    if (nm == 2) {
        // padded row-major Z in zs and padded Z^T in mt
        if (ld1 != d1)
            for (int i = threadIdx.x; i < d0 * ld1; i += NTH) zs[i] = 0.0;
        if (ld0 != d0)
            for (int i = threadIdx.x; i < d1 * ld0; i += NTH) mt[i] = 0.0;
        if (ld1 != d1 || ld0 != d0) __syncthreads();
    }
    for (int i = threadIdx.x; i < p; i += NTH) {
        double z = T.z[i];  // plain load: the covariance loop rewrites Z between calls
        if (T.colcnt != nullptr) {
            const double c = __ldg(T.colcnt + i);
            z = c > 0.0 ? z / c * T.n_total : 0.0;
        }
        s = fma(z, z, s);
        if (nm == 2) {
            const int a = i / d1, j = i - a * d1;
            zs[(size_t)a * ld1 + j] = z;
            mt[(size_t)j * ld0 + a] = z;
        } else {
            zs[i] = z;
        }
    }
    const double normz2 = bsum(s, blk);
    const double normz = sqrt(normz2);

    TPLS_STAMP(1);
    int sweeps = 0;
    if (nm == 2) {
        // ---- matrix Z: the rank-1 CP is the leading singular pair ----
        // tensorly starts ALS from the exact SVD, which is already the fixed point: its two sweeps leave
        // the pair where it is and only settle the sign -- the LAST mode keeps the "largest-|entry|
        // positive" convention of the start, mode 0 follows from Z.  So: one eigenproblem on the short
        // side, the sign rule on f1, then one power step each way (= one ALS sweep) as a polish.
        const int ks = d0 <= d1 ? 0 : 1;
        const int n = T.dims[ks], no = T.dims[1 - ks];
        // rows of Mt are the columns of the mode-ks unfolding: Z^T for ks == 0, Z for ks == 1
        const double* Mt = ks == 0 ? mt : zs;
        syrk_fast(G, Mt, no, up4(n), 1.0);
        lead_eig(G, A, B, f[ks], tmp, n, blk, /*polish=*/1, /*want_lambda=*/false);
        TPLS_STAMP(2);
        if (ks == 0) {
            matvec8(tmp, mt, ld0, f[0], d1, d0);
            normalize_into(f[1], tmp, d1, blk);
        }
        flip_to_positive_peak(f[1], d1, blk);
        matvec8(tmp, zs, ld1, f[1], d0, d1);
        normalize_into(f[0], tmp, d0, blk);
        matvec8(tmp, mt, ld0, f[0], d1, d0);
        normalize_into(f[1], tmp, d1, blk);
        sweeps = 2;
        TPLS_STAMP(3);
    } else {
        // ---- HOSVD start ----
        double weight = 1.0;
        for (int k = 0; k < nm; ++k) {
            const Geo g = mode_geo(T, k);
            double lam;
            if (g.dk <= g.mk) {
                const int n4 = up4(g.dk);
                for (int i = threadIdx.x; i < g.mk * n4; i += NTH) {
                    const int j = i / n4, a = i - j * n4;
                    mt[i] = a < g.dk ? zs[unf_index(g, a, j)] : 0.0;
                }
                __syncthreads();
                syrk_fast(G, mt, g.mk, n4, 1.0);
                lam = lead_eig(G, A, B, f[k], tmp, g.dk, blk, 2, true);
                if (k == 0) weight = sqrt(lam);
            } else {
                const int n4 = up4(g.mk);
                for (int i = threadIdx.x; i < g.dk * n4; i += NTH) {
                    const int a = i / n4, j = i - a * n4;
                    mt[i] = j < g.mk ? zs[unf_index(g, a, j)] : 0.0;
                }
                __syncthreads();
                syrk_fast(G, mt, g.dk, n4, 1.0);
                lead_eig(G, A, B, tmp2, tmp, g.mk, blk, 2, false);
                unf_matvec(f[k], zs, g, tmp2);
                const double sigma = sqrt(vec_dot(f[k], f[k], g.dk, blk));
                vec_div(f[k], sigma, g.dk);
                if (k == 0) weight = sigma;
            }
            flip_to_positive_peak(f[k], g.dk, blk);
        }
        TPLS_STAMP(2);
        TPLS_STAMP(3);

        // ---- ALS sweeps (tensorly parafac, rank 1) ----
        double nrm2[kMaxZModes];
        for (int m = 0; m < nm; ++m) nrm2[m] = vec_dot(f[m], f[m], T.dims[m], blk);
        double err_prev = 0.0;
        double* kr = mt;
        for (int it = 0; it < 100; ++it) {
            ++sweeps;
            double iprod = 0.0;
            for (int k = 0; k < nm; ++k) {
                const int dk = T.dims[k];
                const Geo g = mode_geo(T, k);
                other_modes_product(kr, T, k, f, g.mk);
                unf_matvec(tmp, zs, g, kr);
                double gram = weight * weight;
                for (int m = 0; m < nm; ++m)
                    if (m != k) gram *= nrm2[m];
                // factor = (weight * Z x_others f) / gram; iprod = <mttkrp, factor> (last mode only)
                double s1 = 0.0, s2 = 0.0;
                for (int i = threadIdx.x; i < dk; i += NTH) {
                    const double mt_i = tmp[i] * weight;
                    const double fi = mt_i / gram;
                    f[k][i] = fi;
                    s1 = fma(fi, fi, s1);
                    s2 = fma(mt_i, fi, s2);
                }
                bsum2(s1, s2, blk);
                nrm2[k] = s1;
                if (k == nm - 1) iprod = s2;
            }
            double fn2 = weight * weight;
            for (int m = 0; m < nm; ++m) fn2 *= nrm2[m];
            const double err = sqrt(fabs(normz2 + fn2 - 2.0 * iprod)) / normz;
            const bool stop = it >= 1 && fabs(err_prev - err) < tol;
            err_prev = err;
            if (stop && !normalize_on_break) break;
            // cp_normalize: weights into factor 0, then every column norm back into the weights
            const double w_in = weight;
            weight = 1.0;
            for (int m = 0; m < nm; ++m) {
                const double sc = sqrt(nrm2[m]) * (m == 0 ? fabs(w_in) : 1.0);
                const double dv = sc == 0.0 ? 1.0 : sc;
                double s1 = 0.0;
                for (int i = threadIdx.x; i < T.dims[m]; i += NTH) {
                    const double v = (m == 0 ? f[m][i] * w_in : f[m][i]) / dv;
                    f[m][i] = v;
                    s1 = fma(v, v, s1);
                }
                weight *= sc;
                nrm2[m] = bsum(s1, blk);
            }
            if (stop) break;
        }
    }

    TPLS_STAMP(4);
    // ---- publish ----
    for (int m = 0; m < nm; ++m)
        for (int i = threadIdx.x; i < T.dims[m]; i += NTH) T.w[m][i] = f[m][i];
    for (int i = threadIdx.x; i < T.pitch; i += NTH) {
        double pr = 0.0;
        if (i < p && nm == 2) {
            const int a = i / d1;
            pr = f[0][a] * f[1][i - a * d1];
        } else if (i < p) {
            int rem = i;
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
    TPLS_STAMP(5);
}

// SMEM = true: the workspace is the dynamic shared memory of the CTA (the compiler then emits LDS/STS
// instead of generic loads); false: it is the task's global scratch buffer.
template <bool SMEM>
__global__ void __launch_bounds__(kRank1Threads, 1) rank1_kernel(const __grid_constant__ Rank1Args a) {
    if (trip_is_dead(a.ctrl, a.trip)) return;
    extern __shared__ __align__(16) double dyn[];
    const Rank1Task& T = a.t[blockIdx.x];
    if (SMEM)
        rank1_task(T, a.tol, a.normalize_on_break, dyn);
    else
        rank1_task(T, a.tol, a.normalize_on_break, T.scratch);
}

template <bool SMEM>
__global__ void __launch_bounds__(kRank1Threads, 1) cov_loop_kernel(const __grid_constant__ CovLoopArgs a) {
    extern __shared__ __align__(16) double dyn[];
    __shared__ double q_last[8], q_new[8], r_acc[8], lred[4 * NWARP];
    __shared__ int lired[2 * NWARP];
    Blk blk{lred, lired, 0};
    const int M = a.m, L = a.n_tasks;
    if (threadIdx.x < 8) q_last[threadIdx.x] = threadIdx.x == 0 ? 1.0 : 0.0;  // u_0 = Y[:, 0] = Y e_0 (tpls.py:78)
    __syncthreads();
    int trips = 0;
    for (int trip = 0; trip < a.max_iter; ++trip) {
        trips = trip + 1;
        if (threadIdx.x < 8) r_acc[threadIdx.x] = 0.0;
        for (int l = 0; l < L; ++l) {
            const Rank1Task& T = a.t[l];
            const double* Cl = a.C[l];
            // Z = X x_1 u with u = Y q_last  ==  sum_m q_last[m] * C[m][:]
            double* z = const_cast<double*>(T.z);
            for (int i = threadIdx.x; i < T.p; i += NTH) {
                double v = 0.0;
                for (int m = 0; m < M; ++m) v = fma(q_last[m], Cl[(size_t)m * T.pitch + i], v);
                z[i] = v;
            }
            __syncthreads();
            if (SMEM)
                rank1_task(T, a.tol, a.normalize_on_break, dyn);
            else
                rank1_task(T, a.tol, a.normalize_on_break, T.scratch);
            __syncthreads();
            // Y't for this tensor = C'^T kron(w): the row-rescaled block when masked (missingvals.py:23-38)
            const double* Cq = Cl + (size_t)a.masked[l] * T.pitch;
            for (int m = 0; m < M; ++m) {
                double s = 0.0;
                for (int i = threadIdx.x; i < T.p; i += NTH) s = fma(Cq[(size_t)m * T.pitch + i], T.wkron[i], s);
                s = bsum(s, blk);
                if (threadIdx.x == 0) r_acc[m] += s;
            }
            __syncthreads();
        }
        // q = Y't / ||Y't||, t = average of the tensors' projections (cmtf.py:120-122)
        if (threadIdx.x == 0) {
            double nrm = 0.0;
            for (int m = 0; m < M; ++m) {
                r_acc[m] /= (double)L;
                nrm = fma(r_acc[m], r_acc[m], nrm);
            }
            nrm = sqrt(nrm);
            for (int m = 0; m < M; ++m) q_new[m] = r_acc[m] / nrm;
        }
        __syncthreads();
        bool stop = false;
        if (trip >= 1) {
            double d2 = 0.0;
            for (int i = 0; i < M; ++i) {
                const double di = q_last[i] - q_new[i];
                for (int j = 0; j < M; ++j) d2 = fma(di * a.gram_y[i * M + j], q_last[j] - q_new[j], d2);
            }
            stop = sqrt(fabs(d2)) < a.tol;
        }
        __syncthreads();
        if (threadIdx.x < 8) q_last[threadIdx.x] = threadIdx.x < M ? q_new[threadIdx.x] : 0.0;
        __syncthreads();
        if (stop) break;
    }
    for (int i = threadIdx.x; i < a.pitch_y; i += NTH) {
        const double v = i < M ? q_last[i] : 0.0;
        a.qvec[i] = v;
        if (i < M) a.q_out[i] = v;
    }
    if (threadIdx.x == 0) *a.trips_out = trips;
}

}  // namespace

cudaError_t launch_cov_loop(const CovLoopArgs& a, size_t smem_bytes, bool use_smem, cudaStream_t s) {
    if (use_smem && smem_bytes > 0) {
        cudaError_t e =
            cudaFuncSetAttribute(cov_loop_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
        if (e != cudaSuccess) return e;
        cov_loop_kernel<true><<<1, kRank1Threads, smem_bytes, s>>>(a);
    } else {
        cov_loop_kernel<false><<<1, kRank1Threads, 0, s>>>(a);
    }
    return cudaGetLastError();
}

size_t rank1_workspace_doubles(int nmodes, const int* dims, int* nmax_out, int* zs_len_out, int* mt_len_out) {
    long long p = 1;
    int sumd = 0, maxd = 0;
    for (int m = 0; m < nmodes; ++m) {
        p *= dims[m];
        sumd += dims[m];
        maxd = std::max(maxd, dims[m]);
    }
    long long nmax = 1, zs_len = p, mt_len = p;
    if (nmodes == 2) {
        nmax = std::min(dims[0], dims[1]);
        zs_len = (long long)dims[0] * up4(dims[1]);
        mt_len = (long long)dims[1] * up4(dims[0]);
    } else if (nmodes >= 3) {
        mt_len = 0;
        for (int m = 0; m < nmodes; ++m) {
            const long long dk = dims[m], mk = p / dims[m];
            nmax = std::max(nmax, std::min(dk, mk));
            mt_len = std::max(mt_len, dk <= mk ? mk * up4((int)dk) : dk * up4((int)mk));
        }
        mt_len = std::max(mt_len, p);
    }
    zs_len = (zs_len + 3) & ~3ll;
    mt_len = (mt_len + 3) & ~3ll;
    if (nmax_out) *nmax_out = (int)nmax;
    if (zs_len_out) *zs_len_out = (int)zs_len;
    if (mt_len_out) *mt_len_out = (int)mt_len;
    const long long n4 = up4((int)nmax);
    return (size_t)(zs_len + mt_len + 3 * n4 * n4 + up4(sumd) + 2 * up4((int)std::max<long long>(maxd, nmax)) + 16);
}

cudaError_t launch_rank1(const Rank1Args& a, size_t smem_bytes, cudaStream_t s) {
    if (smem_bytes > 0 && a.t[0].use_smem) {
        cudaError_t e =
            cudaFuncSetAttribute(rank1_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
        if (e != cudaSuccess) return e;
        rank1_kernel<true><<<a.n_tasks, kRank1Threads, smem_bytes, s>>>(a);
    } else {
        rank1_kernel<false><<<a.n_tasks, kRank1Threads, 0, s>>>(a);
    }
    return cudaGetLastError();
}

}  // namespace tpls
