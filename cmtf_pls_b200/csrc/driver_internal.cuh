// Internal to the host side of the C ABI: the device state behind a tpls_handle and the helpers shared by
// driver.cu (fit), transform.cu (transform / predict) and ops.cu (single-operator entry points for tests and
// tuning).  Not part of the public interface (include/tpls_b200.h).
#pragma once

#include "../../include/tpls_b200.h"

#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <string>
#include <vector>

#include "nccl_dl.h"
#include "passes.cuh"
#include "rank1.cuh"
#include "small.cuh"
#include "xchg.cuh"

using namespace tpls;

namespace tpls_drv {

extern std::string g_error;  // error text of calls without a handle
extern NcclApi g_nccl;

struct Tensor {
    bool set = false;
    int dtype = 0, elem = 4, ndim = 0;
    long long shape[TPLS_MAX_MODES] = {0};
    long long n = 0;
    int p = 0, pitch = 0;
    const void* src = nullptr;  // centred from here ...
    void* work = nullptr;       // ... into here (may alias src)
    void* owned = nullptr;      // library-owned staging / work allocation(s)
    void* owned2 = nullptr;
    bool masked = false;
    PassGeom g{};   // column passes
    PassGeom gr{};      // row passes
    PassGeom gr_cnt{};  // the one counting row pass of a masked fit (needs twice the slot space)
    PassGeom gc{};      // cross-covariance passes (covariance mode)
    double *covpart = nullptr, *zscratch = nullptr, *sspart_cov = nullptr;
    int cov_mr = 0;     // accumulators per column of the cross-covariance pass
    size_t off_c = 0;   // arena offset of C [cov_mr][pitch]
    // per-fit device buffers
    double *zpart = nullptr, *cntpart = nullptr, *sspart = nullptr;
    double *mean_d = nullptr, *wkron = nullptr, *tpart = nullptr, *cpart = nullptr, *r1_scratch = nullptr;
    double* rowcnt = nullptr;   // masked: observed entries per row (filled by the first projection of a fit)
    bool rowcnt_ready = false;
    void* mean_native = nullptr;
    double* W[TPLS_MAX_MODES] = {nullptr};
    int* miss_flag = nullptr;
    int* sweeps = nullptr;
    size_t r1_ws = 0;
    int r1_nmax = 1, r1_zs = 0, r1_mt = 0, r1_tab = 0;
    // arena offsets (doubles)
    size_t off_colsum = 0, off_colcnt = 0, off_z = 0, off_ss = 0;
};

}  // namespace tpls_drv

using namespace tpls_drv;

struct tpls_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t private_stream = nullptr;  // created on demand; used when the caller gives no stream
    int sm_count = 148;
    std::string error;
    // comm
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1;
    // peer-memory exchange (xchg.cuh): own buffer + the peers' mappings
    void* xchg_buf = nullptr;
    void* xchg_peer[kXchgMaxRanks] = {nullptr};
    int xchg_cap = 0;
    bool xchg_ready = false;
    // data
    Tensor x[TPLS_MAX_TENSORS];
    long long n = 0;
    int m = 0, pitch_y = 0;
    double *y_src = nullptr, *y_work = nullptr;
    double* row_w = nullptr;  // optional 0/1 sample weights of the next fit (cross-validation folds)
    PassGeom gy{}, gy_row{};
    double h2d_bytes = 0;
    // fit state
    int L = 0, R = 0;
    bool fitted = false;
    std::vector<void*> fit_allocs;
    double *T = nullptr, *U = nullptr, *Q = nullptr, *coef = nullptr, *gram = nullptr;
    double *arena = nullptr, *qvec = nullptr, *svec = nullptr, *ymean_d = nullptr;
    double *zpart_y = nullptr, *cntpart_y = nullptr, *sspart_y = nullptr, *d2part = nullptr, *dotpart = nullptr;
    int sspart_y_n = 0;  // partials of ||Y||^2 the last Y deflation left in sspart_y (its pass's grid, or the resident loop's CTAs)
    double* scratch_ss = nullptr;
    int* trips_dev = nullptr;
    int* ymiss_flag = nullptr;
    Ctrl* ctrl = nullptr;
    int* h_done = nullptr;  // pinned
    // cache of the large X-sized device buffers, reused across fits (cudaMalloc/cudaFree of tens of GB
    // costs hundreds of ms); tpls_trim() returns them to the driver
    struct PoolBuf {
        void* p;
        size_t bytes;
        bool used;
    };
    std::vector<PoolBuf> pool;
    // slab of the per-fit buffers (see dev_alloc) and a grow-only bounce buffer for the getters
    char* slab = nullptr;
    size_t slab_cap = 0, slab_off = 0, slab_need = 0;
    bool slab_dry = false;
    void* tmp_buf = nullptr;
    size_t tmp_cap = 0;
    void* bounce[2] = {nullptr, nullptr};          // pinned host buffers for block-wise results (tpls_reconstruct)
    cudaEvent_t bounce_ev[2] = {nullptr, nullptr};
    size_t arena_doubles = 0, off_ysum = 0, off_ycnt = 0, off_n = 0, off_stats_end = 0, off_zcat = 0, zcat_len = 0,
           off_q = 0, off_d2 = 0, off_dots = 0, off_ss = 0, ss_len = 0;
    std::vector<double> r2x[TPLS_MAX_TENSORS];
    std::vector<double> r2y;
    std::vector<int> trips;
    double n_total = 0;
    bool cov_alloc = false;  // the last alloc_fit reserved the covariance-mode buffers
    size_t off_cov = 0, cov_len = 0, off_gram_y = 0;
    double *grampart = nullptr, *q_prev = nullptr;
    double *qpart = nullptr, *e0vec = nullptr, *nloc = nullptr;
    unsigned int* res_bar = nullptr;  // grid barrier of the resident trip loop (rank1.cuh)
    long long* res_stamps = nullptr;  // its per-phase diagnostics (TPLS_RESIDENT_STAMPS=1), else nullptr
    size_t off_nmiss = 0;
    int* conv_dev = nullptr;
    std::vector<int> converged;
    // device-resident fit (CUDA graph, one WHILE node per component): the instantiated graph of the last fit is
    // kept and relaunched when the next fit has the same key (same buffers, shapes, options)
    cudaStream_t cap_stream = nullptr, body_stream = nullptr;
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t graph_exec = nullptr;
    unsigned long long graph_key = 0;
    bool capturing = false;
    bool graph_broken = false;  // capturing / instantiating failed once on this handle: host-enqueued trips from then on
    // launch accounting of the trip body of every component (launches / collectives / streamed bytes of ONE body,
    // bytes of the trailing contraction that the last body skips, bodies the host enqueued)
    struct BodyCount {
        long long launches = 0, collectives = 0, enqueued = 0;
        double streamed = 0, tail_streamed = 0;
        bool resident = false;  // the component's trips ran in one resident-loop launch
    };
    BodyCount body[32], g_body[32];
    tpls_stats g_static{};  // launches / collectives / bytes of one pass through the captured graph
    int fit_mode = 0;  // 0 host-driven loop, 1 graph
    tpls_stats stats{};
    // optional per-kernel-class timing (TPLS_FIT_PROFILE): event pairs on the launching stream
    bool profile = false;
    struct ProfRec {
        int cls;
        cudaEvent_t a, b;
        double bytes;
    };
    std::vector<ProfRec> prof;
    std::vector<cudaEvent_t> ev_pool;
    tpls_profile prof_sum{};
    cudaEvent_t ev_start = nullptr, ev_stop = nullptr, ev_trip[4] = {nullptr, nullptr, nullptr, nullptr};
};

namespace tpls_drv {

// NVTX range over a scope of host code (fit stages, components and host-enqueued trips)
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};

int fail(tpls_handle h, const char* fmt, ...);

#define CK(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e__ = (call);                                                                        \
        if (e__ != cudaSuccess)                                                                          \
            return fail(h, "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__));        \
    } while (0)

#define CKN(call)                                                                                        \
    do {                                                                                                 \
        int r__ = (call);                                                                                \
        if (r__ != 0) return fail(h, "%s:%d %s -> %s", __FILE__, __LINE__, #call, g_nccl.GetErrorString(r__)); \
    } while (0)

#define TRY(call)                  \
    do {                           \
        int r__ = (call);          \
        if (r__ != 0) return r__;  \
    } while (0)

bool is_device_ptr(const void* p);
int pool_get(tpls_handle h, void** out, size_t bytes);
void pool_put(tpls_handle h, void* p);
void pool_trim(tpls_handle h);
int dev_alloc(tpls_handle h, void** out, size_t bytes, std::vector<void*>* track);
void free_fit(tpls_handle h);
void drop_graph(tpls_handle h);  // forget the captured fit (its buffers or peers are about to change)
void free_tensor(tpls_handle h, Tensor& t);

// ---- per-class timing (TPLS_FIT_PROFILE): an event pair around every launch ----
cudaEvent_t prof_event(tpls_handle h);
struct ProfScope {
    tpls_handle h;
    bool on;
    tpls_ctx::ProfRec r{};
    ProfScope(tpls_handle h_, int cls, double bytes) : h(h_), on(h_->profile) {
        if (!on) return;
        r.cls = cls;
        r.bytes = bytes;
        r.a = prof_event(h);
        r.b = prof_event(h);
        cudaEventRecord(r.a, h->stream);
    }
    ~ProfScope() {
        if (!on) return;
        cudaEventRecord(r.b, h->stream);
        h->prof.push_back(r);
    }
};
void prof_collect(tpls_handle h);

int allreduce_nccl(tpls_handle h, double* buf, size_t count);
int xchg_launch(tpls_handle h, XchgArgs& a);
int allreduce(tpls_handle h, double* buf, size_t count);

// Several per-CTA partial blocks -> their sums in the arena, summed over the ranks: ONE launch on one GPU
// (fold_sets) and ONE launch per rank with the peer-memory exchange; fold + ncclAllReduce otherwise.
// Set offsets are relative to arena + base_off; [base_off, base_off + count) is what crosses the ranks.
struct SetList {
    FoldSet s[kMaxFoldSets];
    int n = 0;
    void add(const double* part, int n_parts, int stride, int n_cols, size_t off) {
        s[n++] = FoldSet{part, n_parts, stride, n_cols, (int)off};
    }
};
int fold_sum(tpls_handle h, const SetList& sl, size_t base_off, size_t count, const Ctrl* ctrl, bool local_only = false);

// pass wrappers that keep the launch / byte counters
int col_pass(tpls_handle h, int dtype, bool masked, int flags, ColPassArgs& a, int cls = -1);
int row_pass(tpls_handle h, int dtype, int mode, RowPassArgs& a, int cls = TPLS_K_PROJECT);
int reduce_cols(tpls_handle h, const double* part, double* out, int n_cols, int stride, int n_parts, const double* sspart,
                double* ss_out, int n_ss, const Ctrl* ctrl, int trip);
int d2_grid(const PassGeom& g);

// (rows x cols, C order) host or device copy of a column-major device matrix
int copy_out_transposed(tpls_handle h, const double* colmajor, long long rows, int cols, double* out);

}  // namespace tpls_drv
