// Rank-1 factorisation of the (small) covariance tensor Z, one CTA per coupled
// tensor, convergence test on the device -- the step the reference delegates to
//   Z / norm(Z)                                            (cmtf_pls/tpls.py:84)
//   tensorly parafac(Z, 1, tol, init="svd", normalize_factors=True)[1]
//                                                          (tpls.py:85-88, cmtf.py:98-102)
// whose algorithm is restated in oracle/tensorly_standin/.../_cp.py.
#pragma once

#include "common.cuh"

namespace tpls {

constexpr int kMaxZModes = 7;   // Z modes (X has one more)
constexpr int kMaxTensors = 8;  // coupled tensors per fit
constexpr int kRank1Threads = 512;

struct Rank1Task {
    const double* z;        // [p] sum over samples (already all-reduced)
    const double* colcnt;   // masked fit: observed rows per column (global); nullptr when dense
    double n_total;         // masked fit: global sample count
    int p;
    int pitch;              // length of wkron (pads are zeroed)
    int nmodes;
    int dims[kMaxZModes];
    double* w[kMaxZModes];  // out: unit weight vector of every mode (contiguous)
    double* wkron;          // out: kron(w_0, w_1, ...) in the row layout of X
    double* scratch;        // global workspace (used when it does not fit in shared memory)
    int use_smem;
    int nmax;               // largest Gram order needed
    int zs_len, mt_len;     // workspace segment lengths (doubles), from rank1_workspace_doubles
    int tab_cols;           // >= 3 modes: sum over the modes of the unfolding widths (index-table length)
    int* sweeps;            // out (optional): ALS sweeps taken
    long long* stamps;      // out (optional, diagnostics): clock64 at phase boundaries [0..7], cycle accumulators [8..15]
};

struct Rank1Args {
    Rank1Task t[kMaxTensors];
    int n_tasks;
    double tol;
    int normalize_on_break;  // see oracle/.../_cp.py NORMALIZE_ON_BREAK
    const Ctrl* ctrl;
    int trip;
};

// Covariance-mode inner iteration: the whole NIPALS loop of one component on
// (P x M)-sized data, in ONE single-CTA launch with the convergence test on the device:
//   Z_l = C_l q_prev;  w_l = rank-1(Z_l);  q = normalise(mean_l C'_l^T kron(w_l));
//   stop when trip >= 1 and sqrt(dq^T (Y'Y) dq) < tol   (= ||u_old - u_new|| of tpls.py:103, u = Y q)
struct CovLoopArgs {
    Rank1Task t[kMaxTensors];          // t[l].z must point at a scratch vector [p] this kernel fills
    const double* C[kMaxTensors];      // all-reduced cross-covariance, layout [mr][pitch]
    int masked[kMaxTensors];           // 0 when dense, else the first row of the row-rescaled block of C (= mr / 2)
    int n_tasks;
    int m;                             // response columns (<= 8)
    const double* gram_y;              // [m][m] all-reduced Y'Y of the current (deflated) Y
    double tol;
    int max_iter;
    int normalize_on_break;
    double* q_out;                     // [m] unit q of the last trip
    double* qvec;                      // [pitch_y] the same, zero-padded
    int pitch_y;
    int* trips_out;                    // inner trips taken
    int* conv_out;                     // optional: 1 when the stop test was met
};
cudaError_t launch_cov_loop(const CovLoopArgs& a, size_t smem_bytes, bool use_smem, cudaStream_t s);

// doubles of workspace a task needs, and the Gram order it implies
size_t rank1_workspace_doubles(int nmodes, const int* dims, int* nmax_out, int* zs_len_out, int* mt_len_out,
                               int* tab_cols_out);
cudaError_t launch_rank1(const Rank1Args& a, size_t smem_bytes, cudaStream_t s);

}  // namespace tpls
