// Streaming passes over one X (or Y) shard: argument blocks and launchers.
//
// X is viewed as n_rows x p, row-major, row pitch `pitch` elements with
// pitch*sizeof(XT) a multiple of 16 (numpy C order, last mode
// fastest).  Two kernel families cover every pass of the fit:
//
//   colpass  : per-COLUMN accumulation over the sample mode
//              (contraction Z = X x_1 u  -- reference cmtf_pls/tpls.py:83,
//               missingvals.py:7-20; column sums for nanmean tpls.py:66;
//               rank-1 deflation tpls.py:109 fused with the NEXT contraction
//               and with the residual norm that gives R2X, util.py:7-15)
//   rowpass  : per-ROW reduction over the feature modes
//              (projection t = X x_2 w2 x_3 w3 ... -- tpls.py:97-99,
//               missingvals.py:23-38; also u = Y q, tpls.py:102)
//
// Both stage row tiles in shared memory with 1-D bulk async copies (TMA engine,
// mbarrier complete_tx) issued by a dedicated producer warp, and accumulate in
// fp64 whatever the storage type.
#pragma once

#include "common.cuh"

namespace tpls {

constexpr int kConsumers = 256;             // consumer threads per CTA
constexpr int kThreads = kConsumers + 32;   // + one producer warp
constexpr int kMaxStages = 6;
constexpr int kMaxCpt = 4;                  // 16-byte column groups per thread

enum PassFlags : int {
    PF_DEFLATE = 1,    // xn = x - a[row] * w[col]
    PF_WRITE = 2,      // store xn (rounded to the storage type)
    PF_CONTRACT = 4,   // zacc[col] += xn * u[row]
    PF_SUMSQ = 8,      // ss += xn^2
    PF_COLSTAT = 16,   // zacc[col] += x over observed rows, cnt[col] += observed
};

struct PassGeom {
    long long n_rows;
    int p;          // valid columns
    int pitch;      // elements per row (multiple of 16/sizeof(XT))
    int slab_w;     // columns per CTA column slab (multiple of 16/sizeof(XT))
    int n_slabs;
    int lpr;        // lanes per row (power of two <= kConsumers)
    int rpt;        // row lanes = kConsumers / lpr
    int cpt;        // 16-byte groups per lane actually needed (1..kMaxCpt)
    int tile_rows;  // rows per staged tile
    int stages;
    int grid_x;     // CTAs along the row-tile axis
    int elem_size;  // 4 or 8
    int dbg;        // probe builds only (-DTPLS_PROBE, tools/probe_build.sh): experiment switches from TPLS_DBG; 0 otherwise
};

// Chooses slab width, thread layout, tile size and grid for a shard.  y_row_bytes > 0: contractions over this
// shard stage the matching rows of Y (pitch_y doubles each) beside every X tile (ColPassArgs::y).
PassGeom make_geom(long long n_rows, int p, int pitch, int elem_size, int sm_count, int y_row_bytes = 0);
size_t colpass_smem(const PassGeom& g, int pitch_y);
constexpr int kMaxFusedResp = 8;            // responses (padded) up to which the Y side of a trip is fused into the X passes
int tune_env(const char* name, int dflt);
// Row passes use their own layout (fewer lanes per row, a slot ring for the reducer warp).
PassGeom make_row_geom(long long n_rows, int p, int pitch, int elem_size, int sm_count, bool masked);
size_t rowpass_smem(const PassGeom& g, bool masked);

struct ColPassArgs {
    PassGeom g;
    const void* x_in;
    void* x_out;            // PF_WRITE (may alias x_in)
    const double* row_a;    // PF_DEFLATE per-row scalar, nullptr => 1
    const double* col_w;    // PF_DEFLATE per-column vector [pitch]
    const double* row_u;    // PF_CONTRACT per-row weights (used when y == nullptr)
    const double* y;        // PF_CONTRACT, optional: responses [n_rows][pitch_y]; then u[row] = y[row,:] . q is formed
    const double* q;        //   in the kernel (tpls.py:102 fused into tpls.py:83) and row_u is not read
    int pitch_y;            //   pitch_y <= kMaxFusedResp, even; the rows of Y ride in the same bulk-copy ring as X
    const double* row_sw;   // optional 0/1 sample weights for PF_COLSTAT / PF_SUMSQ (nullptr => 1)
    double* zpart;          // [grid_x][pitch] per-CTA column partials
    double* cntpart;        // PF_COLSTAT [grid_x][pitch]
    double* sspart;         // PF_SUMSQ [n_slabs * grid_x]; PF_COLSTAT (optional): unobserved entries seen, UNWEIGHTED
    const Ctrl* ctrl;
    int trip;
};

struct RowPassArgs {
    PassGeom g;
    const void* x_in;
    const double* col_w;    // [pitch]
    double* t_out;          // [n_rows]
    double* tpart;          // n_slabs > 1: [n_slabs][n_rows] partial dots
    double* cpart;          // n_slabs > 1 and counting: [n_slabs][n_rows] observed counts
    double* rowcnt;         // masked: [n_rows] observed entries per row (read in mode 1, written in mode 2)
    int epi;                // 0: t = v   1: t += v   2: t = (t + v) / div
    double div;
    double inv_div;         // 1 / div when that is exact (div a power of two: the product rounds like the quotient), else 0
    double* d2part;         // optional [grid_x]: sum over rows of (t_old - t_new)^2
    const double* y;        // optional (with qpart): responses [n_rows][pitch_y], pitch_y <= kMaxFusedResp
    int pitch_y;
    double* qpart;          // optional [grid_x or row_finish grid][kMaxFusedResp]: partials of q = Y't over the FINAL t
                            //   of this CTA's rows (tpls.py:100 fused into the projection's epilogue)
    const Ctrl* ctrl;
    int trip;
};

// dtype: 0 = float32, 1 = float64
cudaError_t launch_colpass(int dtype, bool masked, int flags, const ColPassArgs& a, cudaStream_t s);
// mode: 0 dense, 1 masked with known row counts, 2 masked and counting (writes a.rowcnt)
cudaError_t launch_rowpass(int dtype, int mode, const RowPassArgs& a, cudaStream_t s);

// Cross-covariance pass (covpass.cu): centre/deflate + write + C[m][col] = sum_rows xn * Y[row, m] + ||xn||^2.
struct CovPassArgs {
    PassGeom g;             // make_cov_geom
    const void* x_in;
    void* x_out;            // may alias x_in
    const double* row_a;    // per-row deflation scalar, nullptr => 1 (centring)
    const double* col_w;    // per-column vector [pitch]
    const double* row_sw;   // optional 0/1 sample weights for the residual norm
    const double* y;        // responses [n_rows][pitch_y] (centred / deflated work copy)
    int pitch_y;
    int m;                  // response columns
    const double* rowcnt;   // masked: observed entries per row
    double* cpart;          // [grid_x][c_stride] per-CTA partials, layout [mr][pitch]
    size_t c_stride;
    double* sspart;         // [n_slabs * grid_x]
};
PassGeom make_cov_geom(long long n_rows, int p, int pitch, int elem_size, int sm_count);
size_t covpass_smem(const PassGeom& g, int pitch_y, int mr);
// mr = accumulators per column: next power of two >= m (dense) or >= 2m (masked: second block row-rescaled)
cudaError_t launch_covpass(int dtype, bool masked, int mr, const CovPassArgs& a, cudaStream_t s);

// Single-pass multi-component projection (multiproj.cu): part[slab][row][a] = sum over the slab's columns of
// x[row, c] * w[a][c], a < n_comp <= 32, on the fp64 tensor-core path.  Complete data only (a NaN poisons its row,
// which the finishing kernel reports).
struct MultiProjArgs {
    const void* x;          // [n_rows][pitch] in the storage type
    long long n_rows;
    int p, pitch;
    const double* w;        // [n_comp][w_pitch]
    int w_pitch, n_comp;
    double* part;           // [n_slabs][n_rows][n_comp]
};
int multiproj_slabs(int dtype, int n_comp, int pitch);
cudaError_t launch_multiproj(int dtype, const MultiProjArgs& a, int sm_count, cudaStream_t s);

struct MultiProjFinishArgs {
    const double* part[8];  // per coupled tensor
    int n_slabs[8];
    int n_tensors, n_comp;
    long long n_rows;
    const double* c;        // [n_comp] offsets  mean_l <mean_l, w_l[a]>
    const double* G;        // [n_comp][n_comp]  mean_l <w_l[b], w_l[a]>
    double* S;              // out: scores, column-major n_rows x n_comp
    int* flag;              // |= 1 when a row holds NaNs
};
cudaError_t launch_multiproj_finish(const MultiProjFinishArgs& a, cudaStream_t s);

// X_hat = T W + mean (util.py:18-20): out [n_rows][p] fp64
struct ReconstructArgs {
    const double* T;        // [n_rows][n_comp], C order
    const double* w;        // [n_comp][w_pitch]
    const double* mean;     // [p] or nullptr
    long long n_rows;
    int p, w_pitch, n_comp;
    double* out;
};
cudaError_t launch_reconstruct(const ReconstructArgs& a, cudaStream_t s);

// out[c] = sum_b part[b*stride + c]  (fixed order => bit-reproducible); optionally
// ss_out[0] = sum of sspart[0..n_ss).
struct ReduceArgs {
    const double* part;
    double* out;
    int n_cols;
    int stride;
    int n_parts;
    const double* sspart;
    double* ss_out;
    int n_ss;
    const Ctrl* ctrl;
    int trip;
};
cudaError_t launch_reduce_cols(const ReduceArgs& a, cudaStream_t s);

// Several second-stage folds in ONE launch: out[set.off + c] = sum_b set.part[b*set.stride + c], c < set.n_cols, in the
// association order of reduce_cols (8 interleaved groups of 4 chains).  A vector of per-CTA scalars (sum of squares) is a
// set with stride 1 and one column.  Also the first phase of the peer-memory exchange (xchg.cuh).
constexpr int kMaxFoldSets = 32;
struct FoldSet {
    const double* part;
    int n_parts, stride, n_cols, off;
};
struct FoldArgs {
    FoldSet sets[kMaxFoldSets];
    int n_sets;
    double* out;
    const Ctrl* ctrl;
};
cudaError_t launch_fold_sets(const FoldArgs& a, cudaStream_t s);

// chunk (32 consecutive columns of one set) -> set index and first column; chunks are numbered set by set
__host__ __device__ inline int fold_chunks(const FoldSet* sets, int n_sets) {
    int n = 0;
    for (int s = 0; s < n_sets; ++s) n += (sets[s].n_cols + 31) >> 5;
    return n;
}

// n_slabs > 1 only: t = epilogue(sum over slabs of tpart), masked scaling included.
struct RowFinishArgs {
    long long n_rows;
    int n_slabs;
    const double* tpart;
    const double* cpart;    // counting pass: per-slab observed counts (else nullptr)
    double* rowcnt;         // masked: per-row observed counts (written when cpart != nullptr, else read); nullptr when dense
    double pads;            // zero-filled pad columns per row (they look observed)
    double p_total;
    double* t_out;
    int epi;
    double div;
    double* d2part;         // [gridDim.x]
    const double* y;        // see RowPassArgs
    int pitch_y;
    double* qpart;          // [gridDim.x][kMaxFusedResp]
    const Ctrl* ctrl;
    int trip;
};
int row_finish_grid(long long n_rows);
cudaError_t launch_row_finish(const RowFinishArgs& a, int* grid_out, cudaStream_t s);

}  // namespace tpls
