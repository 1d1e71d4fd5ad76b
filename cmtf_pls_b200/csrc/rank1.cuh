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

// ---------------------------------------------------------------------------------------------------------------
// Resident trip loop ("one persistent kernel per trip for L2-resident problems"): ALL inner trips
// of one component (tpls.py:79-107, cmtf.py:90-128) in ONE launch of one CTA per SM.  The phases of a trip --
// fold of the Z partials, rank-1 step, projection + partials of q = Y't, stop test, contraction for the next trip --
// are separated by grid-wide barriers instead of kernel boundaries (four per trip, ~1.5 us each).  Every CTA owns a
// contiguous block of samples; X does not change during the trips of a component, so the first rows of the block
// (as many as fit in the ~200 KB of shared memory the rank-1 workspace leaves, with their rows of Y) are copied into
// shared memory ONCE per launch by bulk async copies that land behind the first fold and rank-1 step; the rest of the
// block is read from L2 with 16-byte loads in both passes of every trip.  Meant for working sets that stay in the
// 126 MB L2, where a trip of the streaming kernels (7 launches) is bound by launch ramps rather than bytes.  One GPU,
// fused Y side (<= 8 responses); larger problems, several ranks and profiled fits keep the streaming kernels.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kResidentKc = 8;  // 16-byte column groups per thread in the contraction (<= 8 * 512 groups per row)

struct ResidentTensor {
    const void* x;          // working copy [n_rows][pitch] in the storage type
    int dtype;              // 0 = float32, 1 = float64
    int p, pitch;
    int masked;             // NaNs present: zero them, rescale projections by p / rowcnt[row] (missingvals.py:23-38)
    const double* rowcnt;   // masked: observed entries per row
    double* zpart;          // [max(parts0, CTAs)][pitch] per-CTA partials of Z
    int parts0;             // rows of zpart that hold trip 0's partials (written by the fused centring/deflation pass)
    double* z;              // folded Z [pitch] (what the rank-1 task reads)
    const double* wkron;    // kron of this component's weight vectors [pitch] (what the rank-1 task writes)
    int cache_rows;         // set by the launcher: rows of every CTA's block of this tensor that are copied into shared memory
};

struct ResidentArgs {
    ResidentTensor x[kMaxTensors];
    Rank1Task r1[kMaxTensors];
    int n_tensors;
    long long n_rows;
    const double* y;        // centred / deflated responses [n_rows][pitch_y]
    int pitch_y, m;
    double* t_out;          // scores of this component [n_rows]
    double* qpart;          // [CTAs][8] partials of q = Y't
    double* qcol;           // out: unit q of the last trip [m]
    double* qvec;           // out: the same, zero-padded [pitch_y]
    double* q_prev;         // in/out [m] (what the stop test of the streaming loop keeps between launches)
    double* gram;           // out: Y'Y [m][m] of the current Y (formed in the launch: per-CTA partials, folded by every CTA)
    double* grampart;       // [CTAs][64] per-CTA partials of Y'Y
    double* u_out;          // out: u = Y q of the last trip [n_rows] (tpls.py:102; inside the loop it only exists row by row)
    Ctrl* ctrl;
    double tol;
    int max_iter;
    int normalize_on_break;
    int r1_in_smem;         // rank-1 workspace in dynamic shared memory (else the tasks' global scratch)
    unsigned int* bar;      // grid barrier: [0] generation, [32 * (c + 1)] flag of CTA c; 32 * (CTAs + 1) words, zero-initialised once
    // ---- the tail of the component in the same launch (tpls.py:110-113): regression of u on the scores so far, Y
    //      deflation, residual norm of Y; off when tail == 0 (the host then launches the five kernels that do it) ----
    int tail;
    int comp, n_comp;       // index of this component, components of the fit (<= 32)
    const double* T;        // scores so far [n_comp][n_rows] (row comp = t_out)
    const double* row_w;    // optional 0/1 sample weights [n_rows] (cross-validation folds), else nullptr
    double* y_rw;           // = y: deflated in place, y -= (T coef_a) q^T
    double* dotpart;        // [CTAs][64] per-CTA partials of T_b . T_a and T_b . u
    double* gram_t;         // T'T [n_comp][n_comp]: row and column comp are filled here
    double* coef;           // [n_comp][n_comp]: column comp out
    int* trips_out;         // [n_comp]: entry comp out
    int* conv_out;          // [n_comp]: entry comp out (1 = met the stop test)
    double* sspart_y;       // [CTAs] out: per-CTA residual norm of the deflated Y (weighted)
    unsigned w_off;         // set by the launcher: byte offset of the shared-memory copy of kron(w) of all tensors (0: none)
    int r1_everywhere;      // set by the launcher: every CTA runs the rank-1 step (no barrier after it), CTA 0 publishes
    int cache_rows;         // set by the launcher: rows of every CTA's block that EVERY tensor has in shared memory,
    int y_cache_rows;       //   the rows of Y there, and the byte offset of that cache in dynamic shared memory (the copies
    unsigned cache_off;     //   are made once per launch)
    long long* stamps;      // optional diagnostics (TPLS_RESIDENT_STAMPS=1): ns CTA 0 spent per phase, summed over the trips
                            //   [0] fold [1] rank-1 [2] projection [3] q / stop [4] contraction [5..8] the four barriers [9] trips
};
// smem_bytes: rank-1 workspace (0 when it lives in global memory); the launcher adds what the other phases need
cudaError_t launch_resident_loop(const ResidentArgs& a, int n_ctas, size_t r1_smem_bytes, cudaStream_t s);
size_t resident_min_smem();
int resident_fine_stamps(long long* out32);  // probe builds (-DTPLS_PROBE): cycles per sub-phase, summed; 0 otherwise

// doubles of workspace a task needs, and the Gram order it implies
size_t rank1_workspace_doubles(int nmodes, const int* dims, int* nmax_out, int* zs_len_out, int* mt_len_out,
                               int* tab_cols_out);
cudaError_t launch_rank1(const Rank1Args& a, size_t smem_bytes, cudaStream_t s);

}  // namespace tpls
