// Small (latency-bound) kernels of the fit: means, Y-side normalisation, stop
// test, score Gram / regression, linear combinations.
#pragma once

#include "common.cuh"

namespace tpls {

// mean[c] = colsum[c] / colcnt[c], rounded to the storage type of X (the reference's
// np.nanmean keeps X's dtype, tpls.py:66); pads -> 0.  miss_flag |= any column with
// colcnt < n_total, or when *nmiss > 0 (nmiss = all-reduced number of unobserved entries, counted whatever the row
// weights: a cross-validation fold whose NaNs all sit in held-out rows still needs the masked kernels).
// native_out receives the p means in the storage type.
cudaError_t launch_finalize_mean(int dtype, const double* colsum, const double* colcnt, const double* n_total, int p,
                                 int pitch, double* mean_d, void* native_out, int* miss_flag, const double* nmiss,
                                 cudaStream_t s);

// dst[i] = src[i*pitch + col]
cudaError_t launch_gather_col(const double* src, long long n, int pitch, int col, double* dst, cudaStream_t s);

// q = qraw / ||qraw||  (tpls.py:100-101) -> qcol[m] and qvec[pitch] (pads 0)
cudaError_t launch_normalize_q(const double* qraw, int m, int pitch, double* qcol, double* qvec, const Ctrl* ctrl,
                               int trip, cudaStream_t s);

// How a trip ends (tpls.py:103-107).  The kernel that holds ||u_old - u_new||^2 of the trip calls ctrl_decide from
// ONE thread: it records the trip, takes the stop decision (never at trip 0: the reference compares against +inf,
// tpls.py:77), advances the trip counter and -- when the loop is the body of a CUDA-graph WHILE node -- tells the
// node whether to run the body again.
struct LoopEnd {
    double tol;
    int max_iter;
    int has_cond;                        // 0: host-driven loop
    unsigned long long cond;             // cudaGraphConditionalHandle of the WHILE node
};

__device__ __forceinline__ void ctrl_decide(Ctrl* ctrl, double d2, const LoopEnd& e) {
    const int trip = ctrl->trip;
    ctrl->trips_taken = trip + 1;
    ctrl->last_d2 = d2;
    bool stop = false;
    if (trip >= 1 && sqrt(fabs(d2)) < e.tol) {
        ctrl->done_trip = trip;
        stop = true;
    }
    if (trip + 1 >= e.max_iter) stop = true;
    ctrl->trip = trip + 1;
    if (stop) ctrl->stop = 1;
#if defined(__CUDA_ARCH__)
    if (e.has_cond) cudaGraphSetConditional((cudaGraphConditionalHandle)e.cond, stop ? 0u : 1u);
#endif
}

// The same normalisation plus the stop test of tpls.py:103 without a pass over the samples: u = Y q, so
// ||u_old - u_new||^2 = dq^T (Y'Y) dq with dq = q_prev - q (gram = all-reduced Y'Y, m <= 8); q_prev <- q.
// One device routine (thread 0 of a CTA), shared by the kernels that fold it into the reduction that produces
// qraw (reduce_q_stop below, xchg_kernel).
__device__ __forceinline__ void normalize_q_stop_body(const double* qraw, int m, int pitch, double* qcol, double* qvec,
                                                      const double* gram, double* q_prev, Ctrl* ctrl, const LoopEnd& e) {
    double q[8], dq[8];
    double nrm = 0.0;
    for (int i = 0; i < m; ++i) nrm = fma(qraw[i], qraw[i], nrm);
    nrm = sqrt(nrm);
    for (int i = 0; i < m; ++i) {
        q[i] = qraw[i] / nrm;
        dq[i] = q_prev[i] - q[i];
        qcol[i] = q[i];
        q_prev[i] = q[i];
    }
    for (int i = 0; i < pitch; ++i) qvec[i] = i < m ? q[i] : 0.0;
    double d2 = 0.0;
    for (int i = 0; i < m; ++i)
        for (int j = 0; j < m; ++j) d2 = fma(dq[i] * gram[i * m + j], dq[j], d2);
    ctrl_decide(ctrl, d2, e);
}

// Second stage of the q = Y't reduction (same association order as reduce_cols), the normalisation and the stop
// test in ONE launch: qraw[c] = sum_b part[b*stride + c], c < pitch (<= 32).
cudaError_t launch_reduce_q_stop(const double* part, int n_parts, int stride, double* qraw, int m, int pitch, double* qcol,
                                 double* qvec, const double* gram, double* q_prev, Ctrl* ctrl, const LoopEnd& e,
                                 cudaStream_t s);

// stop test of tpls.py:103 on d2 = sum parts (more than kMaxFusedResp responses: explicit ||u_old - u_new||^2)
cudaError_t launch_stop(Ctrl* ctrl, const double* parts, int n, const LoopEnd& e, cudaStream_t s);
cudaError_t launch_reset_ctrl(Ctrl* ctrl, cudaStream_t s);

// part[bx*npairs + j] = partial dot(a[j], b[j]) over this CTA's rows
struct DotPairs {
    const double* a[64];
    const double* b[64];
    const double* w;  // optional per-row weights (nullptr => 1)
    int npairs;
    long long n;
};
cudaError_t launch_multi_dot(const DotPairs& d, double* part, int* grid_out, cudaStream_t s);

// Regression of u_a on the scores so far (np.linalg.lstsq over the non-zero columns,
// tpls.py:110-112): dots = [T_b.T_a (b<=a), T_b.u_a (b<=a)]; updates the persistent
// Gram matrix, solves by Cholesky, writes coef[b*R + a]; also trips_out[a] = ctrl->trips_taken and
// conv_out[a] = the component met the stop test (it may do so on the last allowed trip).
cudaError_t launch_solve_coef(const double* dots, double* gram, double* coef, int R, int a, const Ctrl* ctrl,
                              int* trips_out, int* conv_out, cudaStream_t s);

// s[i] = sum_{b<=a} T_b[i] * coef[b*R + a]      (T column-major, column stride ldt)
cudaError_t launch_lincomb(const double* T, long long n, long long ldt, const double* coef, int R, int a,
                           const double* row_w, double* out, cudaStream_t s);

// y[i, :] *= w[i]   (zeroes the held-out rows of the centred Y in a cross-validation fold)
cudaError_t launch_scale_rows(double* y, long long n, int pitch, const double* w, cudaStream_t s);

// out (n x cols, C order) = in^T where `in` is column-major with column stride ld
cudaError_t launch_transpose_out(const double* in, long long n, long long ld, int cols, double* out, cudaStream_t s);

cudaError_t launch_fill(double* p, long long n, double v, cudaStream_t s);

// part[bx*m*m + i*m + j] = partial of sum_rows y[row,i] * y[row,j]   (m <= 8)
cudaError_t launch_gram_rows(const double* y, long long n, int pitch, int m, double* part, int* grid_out, cudaStream_t s);

// z[c] = cnt[c] > 0 ? z[c] / cnt[c] * n_total : 0     (missingvals.py:18)
cudaError_t launch_count_rescale(double* z, const double* cnt, double n_total, int p, cudaStream_t s);

// dst[i] = (double)src[i], src in the storage type
cudaError_t launch_widen(int dtype, const void* src, double* dst, int n, cudaStream_t s);

}  // namespace tpls
