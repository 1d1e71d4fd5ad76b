// Host side of the C ABI (include/tpls_b200.h): device state of one fit, the
// NIPALS driver that enqueues the passes, and NCCL all-reduces of the small
// replicated quantities.  No CPU arithmetic on the data path.
// transform / predict live in transform.cu, the single-operator entry points in ops.cu.
#include "driver_internal.cuh"

namespace tpls_drv {

std::string g_error;
NcclApi g_nccl;

int fail(tpls_handle h, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (h)
        h->error = buf;
    else
        g_error = buf;
    return 1;
}

bool is_device_ptr(const void* p) {
    cudaPointerAttributes at{};
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}


// Per-fit buffers are carved out of ONE cached slab (cudaMalloc / cudaFree cost milliseconds each once
// NCCL has enabled peer access; a fit needs ~60 buffers).  `track == &h->fit_allocs` selects the slab:
// a dry run adds up the sizes, the real run hands out 256-byte aligned pieces.  Any other `track` is a
// list of pool buffers the caller returns with pool_put.
int dev_alloc(tpls_handle h, void** out, size_t bytes, std::vector<void*>* track) {
    *out = nullptr;
    if (track == &h->fit_allocs) {
        const size_t al = (std::max<size_t>(bytes, 16) + 255) & ~(size_t)255;
        if (h->slab_dry) {
            h->slab_need += al;
            *out = reinterpret_cast<void*>(256);
            return 0;
        }
        if (h->slab_off + al > h->slab_cap) return fail(h, "fit slab overflow (internal error)");
        *out = h->slab + h->slab_off;
        h->slab_off += al;
        return 0;
    }
    TRY(pool_get(h, out, bytes));
    if (track) track->push_back(*out);
    return 0;
}

void pool_trim(tpls_handle h) {
    std::vector<tpls_ctx::PoolBuf> keep;
    for (auto& b : h->pool) {
        if (b.used)
            keep.push_back(b);
        else
            cudaFree(b.p);
    }
    h->pool.swap(keep);
}

int pool_get(tpls_handle h, void** out, size_t bytes) {
    bytes = std::max<size_t>(bytes, 256);
    for (auto& b : h->pool)
        if (!b.used && b.bytes >= bytes && b.bytes <= bytes + bytes / 4 + (1u << 20)) {
            b.used = true;
            *out = b.p;
            return 0;
        }
    cudaError_t e = cudaMalloc(out, bytes);
    if (e != cudaSuccess) {
        cudaGetLastError();
        pool_trim(h);
        CK(cudaMalloc(out, bytes));
    }
    h->pool.push_back({*out, bytes, true});
    return 0;
}

void pool_put(tpls_handle h, void* p) {
    for (auto& b : h->pool)
        if (b.p == p) b.used = false;
}

void free_fit(tpls_handle h) {
    h->fit_allocs.clear();
    h->slab_off = 0;
    h->fitted = false;
}

void free_tensor(tpls_handle h, Tensor& t) {
    if (t.owned) pool_put(h, t.owned);
    if (t.owned2) pool_put(h, t.owned2);
    t = Tensor();
}


// ---- per-class timing ----
cudaEvent_t prof_event(tpls_handle h) {
    cudaEvent_t e = nullptr;
    if (!h->ev_pool.empty()) {
        e = h->ev_pool.back();
        h->ev_pool.pop_back();
    } else {
        cudaEventCreate(&e);
    }
    return e;
}


// Sums the launch records of every profiled fit since the last collection (event queries cost a few
// microseconds each, so they are made when the profile is asked for, not inside the fit).
void prof_collect(tpls_handle h) {
    h->prof_sum = tpls_profile{};
    for (auto& r : h->prof) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, r.a, r.b);
        h->prof_sum.ms[r.cls] += ms;
        h->prof_sum.launches[r.cls] += 1;
        h->prof_sum.bytes[r.cls] += r.bytes;
        h->ev_pool.push_back(r.a);
        h->ev_pool.push_back(r.b);
    }
    h->prof.clear();
}

int allreduce_nccl(tpls_handle h, double* buf, size_t count) {
    if (h->world <= 1 || count == 0) return 0;
    ProfScope ps(h, TPLS_K_NCCL, 0.0);
    CKN(g_nccl.AllReduce(buf, buf, count, kNcclFloat64, kNcclSum, h->comm, h->stream));
    h->stats.collectives++;
    return 0;
}

// fills the communicator part of an exchange and launches it (never skipped: see xchg.cuh)
int xchg_launch(tpls_handle h, XchgArgs& a) {
    ProfScope ps(h, TPLS_K_NCCL, 0.0);
    a.cap = h->xchg_cap;
    a.rank = h->rank;
    a.world = h->world;
    a.seq = ++h->xchg_seq;
    for (int r = 0; r < h->world; ++r) {
        char* base = static_cast<char*>(h->xchg_peer[r]);
        a.flags[r] = reinterpret_cast<unsigned long long*>(base);
        a.data[r] = reinterpret_cast<double*>(base + kXchgHeaderBytes);
    }
    char* mine = static_cast<char*>(h->xchg_buf);
    a.done_ctr = reinterpret_cast<unsigned int*>(mine + 512);
    a.err = reinterpret_cast<int*>(mine + 520);
    CK(launch_xchg(a, h->stream));
    h->stats.kernel_launches++;
    h->stats.collectives++;
    return 0;
}

// Sum of a small replicated vector over the ranks: the one-shot peer-memory exchange when it is set up
// (and the vector fits its slots), NCCL otherwise.
int allreduce(tpls_handle h, double* buf, size_t count) {
    if (h->world <= 1 || count == 0) return 0;
    if (!h->xchg_ready || count > (size_t)h->xchg_cap) return allreduce_nccl(h, buf, count);
    XchgArgs a{};
    a.in = buf;
    a.out = buf;
    a.count = (int)count;
    return xchg_launch(h, a);
}

// ---- pass wrappers that keep the launch / byte counters ----
int col_pass(tpls_handle h, int dtype, bool masked, int flags, ColPassArgs& a, int cls) {
    const double bytes = (double)a.g.n_rows * a.g.pitch * a.g.elem_size * ((flags & PF_WRITE) ? 2.0 : 1.0);
    if (cls < 0) {
        cls = TPLS_K_OTHER;
        if (flags == PF_COLSTAT) cls = TPLS_K_COLSTAT;
        if (flags == PF_CONTRACT) cls = TPLS_K_CONTRACT;
        if (flags == (PF_DEFLATE | PF_WRITE | PF_CONTRACT | PF_SUMSQ)) cls = TPLS_K_DEFLATE_CONTRACT;
        if (flags == (PF_DEFLATE | PF_SUMSQ)) cls = TPLS_K_RESIDUAL;
    }
    ProfScope ps(h, cls, bytes);
    CK(launch_colpass(dtype, masked, flags, a, h->stream));
    h->stats.kernel_launches++;
    if (cls != TPLS_K_YSIDE) h->stats.streamed_bytes += bytes;
    return 0;
}

// mode: 0 dense, 1 masked with known row counts (a.rowcnt), 2 masked and counting (fills a.rowcnt)
int row_pass(tpls_handle h, int dtype, int mode, RowPassArgs& a, int cls) {
    const double bytes = (double)a.g.n_rows * a.g.pitch * a.g.elem_size;
    ProfScope ps(h, cls, bytes);
    CK(launch_rowpass(dtype, mode, a, h->stream));
    h->stats.kernel_launches++;
    if (cls != TPLS_K_YSIDE) h->stats.streamed_bytes += bytes;
    if (a.g.n_slabs > 1) {
        RowFinishArgs f{};
        f.n_rows = a.g.n_rows;
        f.n_slabs = a.g.n_slabs;
        f.tpart = a.tpart;
        f.cpart = mode == 2 ? a.cpart : nullptr;
        f.rowcnt = mode != 0 ? a.rowcnt : nullptr;
        f.pads = (double)(a.g.pitch - a.g.p);
        f.p_total = (double)a.g.p;
        f.t_out = a.t_out;
        f.epi = a.epi;
        f.div = a.div;
        f.d2part = a.d2part;
        f.ctrl = a.ctrl;
        f.trip = a.trip;
        CK(launch_row_finish(f, nullptr, h->stream));
        h->stats.kernel_launches++;
    }
    return 0;
}

int reduce_cols(tpls_handle h, const double* part, double* out, int n_cols, int stride, int n_parts,
                const double* sspart, double* ss_out, int n_ss, const Ctrl* ctrl, int trip) {
    ReduceArgs r{};
    r.part = part;
    r.out = out;
    r.n_cols = n_cols;
    r.stride = stride;
    r.n_parts = n_parts;
    r.sspart = sspart;
    r.ss_out = ss_out;
    r.n_ss = n_ss;
    r.ctrl = ctrl;
    r.trip = trip;
    ProfScope ps(h, TPLS_K_OTHER, 0.0);
    CK(launch_reduce_cols(r, h->stream));
    h->stats.kernel_launches++;
    return 0;
}

int d2_grid(const PassGeom& g) { return g.n_slabs > 1 ? row_finish_grid(g.n_rows) : g.grid_x; }

int copy_out_transposed(tpls_handle h, const double* colmajor, long long rows, int cols, double* out) {
    CK(cudaSetDevice(h->device));
    if (is_device_ptr(out)) {
        CK(launch_transpose_out(colmajor, rows, rows, cols, out, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        return 0;
    }
    const size_t need = sizeof(double) * std::max<long long>(1, rows * cols);
    if (need > h->tmp_cap) {
        if (h->tmp_buf) pool_put(h, h->tmp_buf);
        h->tmp_buf = nullptr;
        h->tmp_cap = 0;
        TRY(pool_get(h, &h->tmp_buf, need));
        h->tmp_cap = need;
    }
    double* tmp = static_cast<double*>(h->tmp_buf);
    CK(launch_transpose_out(colmajor, rows, rows, cols, tmp, h->stream));
    CK(cudaMemcpyAsync(out, tmp, sizeof(double) * rows * cols, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}

}  // namespace tpls_drv

// ===========================================================================
// C ABI
// ===========================================================================
extern "C" {

int tpls_version(void) { return 100; }

const char* tpls_last_error(tpls_handle h) { return h ? h->error.c_str() : g_error.c_str(); }

int tpls_create(tpls_handle* out, int device, void* cuda_stream) {
    tpls_handle h = nullptr;
    if (!out) return fail(nullptr, "tpls_create: out is NULL");
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
        cudaGetLastError();
        return fail(nullptr, "tpls_create: no CUDA device is visible (this library has no CPU path)");
    }
    if (device < 0 || device >= count) return fail(nullptr, "tpls_create: device %d out of range (%d visible)", device, count);
    cudaDeviceProp prop{};
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return fail(nullptr, "cudaGetDeviceProperties failed");
    if (prop.major != 10)
        return fail(nullptr, "tpls_create: device %d is sm_%d%d; this build carries sm_100a code only", device, prop.major,
                    prop.minor);
    h = new tpls_ctx();
    h->device = device;
    h->sm_count = prop.multiProcessorCount;
    cudaSetDevice(device);
    if (cuda_stream) {
        h->stream = (cudaStream_t)cuda_stream;
    } else {
        if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) {
            delete h;
            return fail(nullptr, "cudaStreamCreate failed");
        }
        h->own_stream = true;
    }
    cudaEventCreate(&h->ev_start);
    cudaEventCreate(&h->ev_stop);
    for (auto& e : h->ev_trip) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
    cudaMallocHost((void**)&h->h_done, sizeof(int) * 4096);
    *out = h;
    return 0;
}

int tpls_destroy(tpls_handle h) {
    if (!h) return 0;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    free_fit(h);
    for (auto& t : h->x) free_tensor(h, t);
    if (h->y_src) pool_put(h, h->y_src);
    if (h->y_work) pool_put(h, h->y_work);
    if (h->row_w) pool_put(h, h->row_w);
    if (h->slab) pool_put(h, h->slab);
    if (h->tmp_buf) pool_put(h, h->tmp_buf);
    pool_trim(h);
    if (h->xchg_buf) {
        for (int r = 0; r < h->world; ++r)
            if (r != h->rank && h->xchg_peer[r]) cudaIpcCloseMemHandle(h->xchg_peer[r]);
        cudaFree(h->xchg_buf);
    }
    if (h->comm) g_nccl.CommDestroy(h->comm);
    if (h->h_done) cudaFreeHost(h->h_done);
    cudaEventDestroy(h->ev_start);
    cudaEventDestroy(h->ev_stop);
    for (auto& e : h->ev_trip) cudaEventDestroy(e);
    for (auto& e : h->ev_pool) cudaEventDestroy(e);
    for (auto& r : h->prof) {  // launch records of profiled fits that were never collected
        cudaEventDestroy(r.a);
        cudaEventDestroy(r.b);
    }
    if (h->own_stream) cudaStreamDestroy(h->stream);
    delete h;
    return 0;
}

int tpls_comm_unique_id(void* id128) {
    const char* why = "";
    if (!g_nccl.load(&why)) return fail(nullptr, "tpls_comm_unique_id: %s", why);
    NcclUniqueId id;
    int r = g_nccl.GetUniqueId(&id);
    if (r != 0) return fail(nullptr, "ncclGetUniqueId -> %s", g_nccl.GetErrorString(r));
    memcpy(id128, &id, sizeof id);
    return 0;
}

int tpls_comm_init(tpls_handle h, const void* id128, int rank, int world) {
    if (!h) return fail(nullptr, "NULL handle");
    if (world <= 1) {
        h->rank = 0;
        h->world = 1;
        return 0;
    }
    const char* why = "";
    if (!g_nccl.load(&why)) return fail(h, "tpls_comm_init: %s", why);
    CK(cudaSetDevice(h->device));
    NcclUniqueId id;
    memcpy(&id, id128, sizeof id);
    CKN(g_nccl.CommInitRank(&h->comm, world, id, rank));
    h->rank = rank;
    h->world = world;
    return 0;
}

int tpls_comm_xchg_handle(tpls_handle h, void* handle64) {
    if (!h) return fail(nullptr, "NULL handle");
    if (h->world <= 1) return fail(h, "tpls_comm_xchg_handle: no communicator (call tpls_comm_init first)");
    if (h->world > kXchgMaxRanks) return fail(h, "tpls_comm_xchg_handle: at most %d ranks", kXchgMaxRanks);
    CK(cudaSetDevice(h->device));
    if (!h->xchg_buf) {
        h->xchg_cap = 1 << 17;  // doubles per slot (1 MiB): Z of all coupled tensors fits
        const size_t bytes = kXchgHeaderBytes + 2 * sizeof(double) * (size_t)h->xchg_cap;
        CK(cudaMalloc(&h->xchg_buf, bytes));
        CK(cudaMemset(h->xchg_buf, 0, bytes));
        CK(cudaDeviceSynchronize());
    }
    cudaIpcMemHandle_t hnd;
    CK(cudaIpcGetMemHandle(&hnd, h->xchg_buf));
    static_assert(sizeof(hnd) == 64, "cudaIpcMemHandle_t is 64 bytes");
    memcpy(handle64, &hnd, sizeof hnd);
    return 0;
}

int tpls_comm_xchg_open(tpls_handle h, const void* handles) {
    if (!h) return fail(nullptr, "NULL handle");
    if (handles == nullptr) {  // switch the exchange off again (every rank must do the same)
        h->xchg_ready = false;
        return 0;
    }
    if (!h->xchg_buf) return fail(h, "tpls_comm_xchg_open: call tpls_comm_xchg_handle first");
    CK(cudaSetDevice(h->device));
    for (int r = 0; r < h->world; ++r) {
        if (r == h->rank) {
            h->xchg_peer[r] = h->xchg_buf;
            continue;
        }
        cudaIpcMemHandle_t hnd;
        memcpy(&hnd, static_cast<const char*>(handles) + 64 * (size_t)r, sizeof hnd);
        CK(cudaIpcOpenMemHandle(&h->xchg_peer[r], hnd, cudaIpcMemLazyEnablePeerAccess));
    }
    h->xchg_seq = 0;
    h->xchg_ready = true;
    return 0;
}

int tpls_set_x(tpls_handle h, int index, const void* x, int dtype, int ndim, const int64_t* shape, int flags) {
    if (!h) return fail(nullptr, "NULL handle");
    if (index < 0 || index >= TPLS_MAX_TENSORS) return fail(h, "tpls_set_x: index %d out of range", index);
    if (ndim < 2 || ndim > TPLS_MAX_MODES) return fail(h, "tpls_set_x: X must have 2..%d modes, got %d", TPLS_MAX_MODES, ndim);
    if (dtype != TPLS_F32 && dtype != TPLS_F64) return fail(h, "tpls_set_x: dtype must be TPLS_F32 or TPLS_F64");
    CK(cudaSetDevice(h->device));
    Tensor& t = h->x[index];
    free_tensor(h, t);
    t.dtype = dtype;
    t.elem = dtype == TPLS_F32 ? 4 : 8;
    t.ndim = ndim;
    long long p = 1;
    for (int k = 0; k < ndim; ++k) {
        if (shape[k] <= 0) return fail(h, "tpls_set_x: empty mode %d", k);
        if (k && ndim > 3 && shape[k] > 65535)  // the rank-1 kernel keeps per-mode indices of a >= 3-way Z in 16 bits
            return fail(h, "tpls_set_x: mode %d of a %d-way X has %lld entries, at most 65535 are supported", k, ndim, (long long)shape[k]);
        t.shape[k] = shape[k];
        if (k) p *= shape[k];
    }
    if (p > (1ll << 30)) return fail(h, "tpls_set_x: trailing size %lld too large", p);
    t.n = shape[0];
    t.p = (int)p;
    const int vec = 16 / t.elem;
    t.pitch = (int)((p + vec - 1) / vec * vec);
    const size_t bytes = (size_t)t.n * t.pitch * t.elem;
    const bool dev = is_device_ptr(x);
    if (dev && t.pitch == t.p && ((uintptr_t)x % 16 == 0)) {
        t.src = x;
        if (flags & TPLS_X_MAY_OVERWRITE) {
            t.work = const_cast<void*>(x);
        } else {
            TRY(pool_get(h, &t.owned, bytes));
            t.work = t.owned;
        }
    } else {
        TRY(pool_get(h, &t.owned, bytes));
        if (t.pitch != t.p) CK(cudaMemsetAsync(t.owned, 0, bytes, h->stream));
        CK(cudaMemcpy2DAsync(t.owned, (size_t)t.pitch * t.elem, x, (size_t)t.p * t.elem, (size_t)t.p * t.elem, t.n,
                             cudaMemcpyDefault, h->stream));
        if (!dev) h->h2d_bytes += (double)t.n * t.p * t.elem;
        t.src = t.owned;
        t.work = t.owned;
    }
    t.set = true;
    h->fitted = false;
    return 0;
}

int tpls_set_y(tpls_handle h, const double* y, int64_t n, int64_t m) {
    if (!h) return fail(nullptr, "NULL handle");
    if (n <= 0 || m <= 0) return fail(h, "tpls_set_y: empty Y");
    CK(cudaSetDevice(h->device));
    if (h->y_src) pool_put(h, h->y_src);
    if (h->y_work) pool_put(h, h->y_work);
    h->y_src = h->y_work = nullptr;
    h->n = n;
    h->m = (int)m;
    h->pitch_y = (int)((m + 1) / 2 * 2);
    const size_t bytes = (size_t)n * h->pitch_y * sizeof(double);
    TRY(pool_get(h, (void**)&h->y_src, bytes));
    TRY(pool_get(h, (void**)&h->y_work, bytes));
    if (h->pitch_y != m) CK(cudaMemsetAsync(h->y_src, 0, bytes, h->stream));
    CK(cudaMemcpy2DAsync(h->y_src, (size_t)h->pitch_y * 8, y, (size_t)m * 8, (size_t)m * 8, n, cudaMemcpyDefault,
                         h->stream));
    if (!is_device_ptr(y)) h->h2d_bytes += (double)n * m * 8;
    h->fitted = false;
    return 0;
}

// ---------------------------------------------------------------------------
// fit
// ---------------------------------------------------------------------------
static int layout_fit(tpls_handle h, int L, int R);

static int alloc_fit(tpls_handle h, int L, int R) {
    free_fit(h);
    h->slab_dry = true;
    h->slab_need = 0;
    TRY(layout_fit(h, L, R));
    h->slab_dry = false;
    if (h->slab_need > h->slab_cap) {
        if (h->slab) pool_put(h, h->slab);
        h->slab = nullptr;
        h->slab_cap = 0;
        void* p = nullptr;
        TRY(pool_get(h, &p, h->slab_need));
        h->slab = static_cast<char*>(p);
        h->slab_cap = h->slab_need;
    }
    CK(cudaMemsetAsync(h->slab, 0, h->slab_need, h->stream));
    h->slab_off = 0;
    return layout_fit(h, L, R);
}

static int layout_fit(tpls_handle h, int L, int R) {
    auto* tr = &h->fit_allocs;
    const long long n = h->n;
    h->L = L;
    h->R = R;
    h->gy = make_geom(n, h->m, h->pitch_y, 8, h->sm_count);
    h->gy_row = make_row_geom(n, h->m, h->pitch_y, 8, h->sm_count, false);
    // arena layout
    size_t off = 0;
    for (int l = 0; l < L; ++l) {
        Tensor& t = h->x[l];
        t.g = make_geom(n, t.p, t.pitch, t.elem, h->sm_count);
        t.off_colsum = off;
        off += t.pitch;
        t.off_colcnt = off;
        off += t.pitch;
    }
    h->off_ysum = off;
    off += h->pitch_y;
    h->off_ycnt = off;
    off += h->pitch_y;
    h->off_n = off;
    off += 1;
    h->off_stats_end = off;
    off = (off + 1) / 2 * 2;
    h->off_zcat = off;
    for (int l = 0; l < L; ++l) {
        h->x[l].off_z = off;
        off += h->x[l].pitch;
    }
    h->zcat_len = off - h->off_zcat;
    h->off_q = off;
    off += h->pitch_y;
    h->off_d2 = off;
    off += 2;
    h->off_dots = off;
    off += 64;
    h->off_ss = off;
    for (int l = 0; l < L; ++l) {
        h->x[l].off_ss = off;
        off += R + 1;
    }
    off += R + 1;  // Y
    h->ss_len = off - h->off_ss;
    // covariance mode: C of every tensor followed by Y'Y, one all-reduce per component
    off = (off + 1) / 2 * 2;
    h->off_cov = off;
    if (h->cov_alloc) {
        const int mr_alloc = h->m <= 1 ? 2 : (h->m <= 2 ? 4 : 8);  // room for the masked (row-rescaled) block when m <= 4
        for (int l = 0; l < L; ++l) {
            h->x[l].off_c = off;
            off += (size_t)mr_alloc * h->x[l].pitch;
        }
    }
    h->off_gram_y = off;
    off += 64;
    h->cov_len = off - h->off_cov;
    h->arena_doubles = off;
    TRY(dev_alloc(h, (void**)&h->arena, off * sizeof(double), tr));

    TRY(dev_alloc(h, (void**)&h->T, sizeof(double) * n * R, tr));
    TRY(dev_alloc(h, (void**)&h->U, sizeof(double) * n * R, tr));
    TRY(dev_alloc(h, (void**)&h->Q, sizeof(double) * h->m * R, tr));
    TRY(dev_alloc(h, (void**)&h->coef, sizeof(double) * R * R, tr));
    TRY(dev_alloc(h, (void**)&h->gram, sizeof(double) * R * R, tr));
    TRY(dev_alloc(h, (void**)&h->qvec, sizeof(double) * h->pitch_y, tr));
    TRY(dev_alloc(h, (void**)&h->svec, sizeof(double) * n, tr));
    TRY(dev_alloc(h, (void**)&h->ymean_d, sizeof(double) * h->pitch_y, tr));
    TRY(dev_alloc(h, (void**)&h->zpart_y, sizeof(double) * h->gy.grid_x * h->pitch_y, tr));
    TRY(dev_alloc(h, (void**)&h->cntpart_y, sizeof(double) * h->gy.grid_x * h->pitch_y, tr));
    TRY(dev_alloc(h, (void**)&h->sspart_y, sizeof(double) * h->gy.grid_x * h->gy.n_slabs, tr));
    TRY(dev_alloc(h, (void**)&h->d2part, sizeof(double) * 2048, tr));
    TRY(dev_alloc(h, (void**)&h->dotpart, sizeof(double) * 148 * 64, tr));
    TRY(dev_alloc(h, (void**)&h->grampart, sizeof(double) * 148 * 64, tr));
    TRY(dev_alloc(h, (void**)&h->q_prev, sizeof(double) * 8, tr));
    TRY(dev_alloc(h, (void**)&h->scratch_ss, sizeof(double) * 8, tr));
    TRY(dev_alloc(h, (void**)&h->trips_dev, sizeof(int) * R, tr));
    TRY(dev_alloc(h, (void**)&h->ymiss_flag, sizeof(int) * 4, tr));
    TRY(dev_alloc(h, (void**)&h->ctrl, sizeof(Ctrl), tr));

    for (int l = 0; l < L; ++l) {
        Tensor& t = h->x[l];
        const size_t gp = (size_t)t.g.grid_x * t.pitch;
        TRY(dev_alloc(h, (void**)&t.zpart, sizeof(double) * gp, tr));
        TRY(dev_alloc(h, (void**)&t.cntpart, sizeof(double) * gp, tr));
        TRY(dev_alloc(h, (void**)&t.sspart, sizeof(double) * t.g.grid_x * t.g.n_slabs, tr));
        TRY(dev_alloc(h, (void**)&t.mean_d, sizeof(double) * t.pitch, tr));
        TRY(dev_alloc(h, &t.mean_native, (size_t)t.elem * t.pitch, tr));
        TRY(dev_alloc(h, (void**)&t.wkron, sizeof(double) * t.pitch * R, tr));
        TRY(dev_alloc(h, (void**)&t.rowcnt, sizeof(double) * n, tr));
        t.rowcnt_ready = false;
        if (h->cov_alloc) {
            const int mr_alloc = h->m <= 1 ? 2 : (h->m <= 2 ? 4 : 8);
            t.gc = make_cov_geom(n, t.p, t.pitch, t.elem, h->sm_count);
            TRY(dev_alloc(h, (void**)&t.covpart, sizeof(double) * (size_t)t.gc.grid_x * mr_alloc * t.pitch, tr));
            TRY(dev_alloc(h, (void**)&t.sspart_cov, sizeof(double) * (size_t)t.gc.grid_x * t.gc.n_slabs, tr));
            TRY(dev_alloc(h, (void**)&t.zscratch, sizeof(double) * t.pitch, tr));
        }
        if (t.g.n_slabs > 1) {
            TRY(dev_alloc(h, (void**)&t.tpart, sizeof(double) * n * t.g.n_slabs, tr));
            TRY(dev_alloc(h, (void**)&t.cpart, sizeof(double) * n * t.g.n_slabs, tr));
        }
        for (int k = 1; k < t.ndim; ++k) {
            TRY(dev_alloc(h, (void**)&t.W[k], sizeof(double) * t.shape[k] * R, tr));
        }
        TRY(dev_alloc(h, (void**)&t.miss_flag, sizeof(int) * 4, tr));
        TRY(dev_alloc(h, (void**)&t.sweeps, sizeof(int) * 4, tr));
        int dims[kMaxZModes];
        for (int k = 1; k < t.ndim; ++k) dims[k - 1] = (int)t.shape[k];
        t.r1_ws = rank1_workspace_doubles(t.ndim - 1, dims, &t.r1_nmax, &t.r1_zs, &t.r1_mt, &t.r1_tab);
        TRY(dev_alloc(h, (void**)&t.r1_scratch, sizeof(double) * t.r1_ws, tr));
    }
    return 0;
}

static void fill_rank1_task(tpls_handle h, Tensor& t, int a, Rank1Task& k, bool use_smem) {
    k.z = h->arena + t.off_z;
    k.colcnt = t.masked ? h->arena + t.off_colcnt : nullptr;
    k.n_total = h->n_total;
    k.p = t.p;
    k.pitch = t.pitch;
    k.nmodes = t.ndim - 1;
    for (int m = 1; m < t.ndim; ++m) {
        k.dims[m - 1] = (int)t.shape[m];
        k.w[m - 1] = t.W[m] + (size_t)a * t.shape[m];
    }
    k.wkron = t.wkron + (size_t)a * t.pitch;
    k.scratch = t.r1_scratch;
    k.use_smem = use_smem ? 1 : 0;
    k.nmax = t.r1_nmax;
    k.zs_len = t.r1_zs;
    k.mt_len = t.r1_mt;
    k.tab_cols = t.r1_tab;
    k.sweeps = t.sweeps;
}


// Regression of u_a on the scores so far and the Y deflation that follows (tpls.py:110-113), shared by the
// streaming and the covariance component loops.  stream_mode: the trip count comes from the control block.
static int component_tail(tpls_handle h, int R, int a, double* ss_y, bool stream_mode) {
    cudaStream_t st = h->stream;
    const long long n = h->n;
    double* A = h->arena;
    double* Ta = h->T + (size_t)a * n;
    double* Ua = h->U + (size_t)a * n;
    // ---- regression on the scores so far (tpls.py:110-112) ----
    {
        DotPairs d{};
        d.n = n;
        d.npairs = 2 * (a + 1);
        d.w = h->row_w;
        for (int b = 0; b <= a; ++b) {
            d.a[b] = h->T + (size_t)b * n;
            d.b[b] = Ta;
            d.a[a + 1 + b] = h->T + (size_t)b * n;
            d.b[a + 1 + b] = Ua;
        }
        int gx = 1;
        {
            ProfScope ps_small(h, TPLS_K_OTHER, 0.0);
            CK(launch_multi_dot(d, h->dotpart, &gx, st));
            h->stats.kernel_launches++;
        }
        TRY(reduce_cols(h, h->dotpart, A + h->off_dots, d.npairs, d.npairs, gx, nullptr, nullptr, 0, nullptr, 0));
        TRY(allreduce(h, A + h->off_dots, d.npairs));
        {
            ProfScope ps_small(h, TPLS_K_OTHER, 0.0);
            CK(launch_solve_coef(A + h->off_dots, h->gram, h->coef, R, a, h->ctrl, stream_mode ? h->trips_dev : nullptr, st));
            h->stats.kernel_launches++;
        }
    }
    // ---- Y deflation (tpls.py:113) + ||Y||^2 for R2Y ----
    {
        {
            ProfScope ps_small(h, TPLS_K_OTHER, 0.0);
            CK(launch_lincomb(h->T, n, n, h->coef, R, a, h->row_w, h->svec, st));
            h->stats.kernel_launches++;
        }
        ColPassArgs c{};
        c.g = h->gy;
        c.x_in = h->y_work;
        c.x_out = h->y_work;
        c.row_a = h->svec;
        c.col_w = h->qvec;
        c.sspart = h->sspart_y;
        c.row_sw = h->row_w;
        TRY(col_pass(h, TPLS_F64, false, PF_DEFLATE | PF_WRITE | PF_SUMSQ, c, TPLS_K_YSIDE));
        TRY(reduce_cols(h, nullptr, nullptr, 0, 0, 0, h->sspart_y, ss_y + a + 1, h->gy.grid_x * h->gy.n_slabs, nullptr, 0));
    }
    return 0;
}

// ---------------------------------------------------------------------------
// covariance-mode component loop: per component one cross-covariance pass
// (centring / deflation fused in), the inner iteration on (P x M)-sized data in a single kernel,
// one projection pass; no per-trip pass over X and no per-trip collective.
// ---------------------------------------------------------------------------
static int pow2_at_least(int v) {
    int p = 1;
    while (p < v) p <<= 1;
    return p;
}

static bool cov_supported(tpls_handle h, int L) {
    if (!h->cov_alloc || h->m > 8) return false;
    for (int l = 0; l < L; ++l)
        if (h->x[l].masked && h->m > 4) return false;
    return true;
}

static int fit_covariance(tpls_handle h, int L, int R, double tol, int max_iter, int flags, double* ss_y) {
    cudaStream_t st = h->stream;
    const long long n = h->n;
    double* A = h->arena;
    const int M = h->m;

    size_t r1_smem = 0;
    bool r1_use_smem = true;
    for (int l = 0; l < L; ++l) r1_smem = std::max(r1_smem, h->x[l].r1_ws * sizeof(double));
    if (r1_smem > 200 * 1024) {
        r1_use_smem = false;
        r1_smem = 0;
    }
    for (int l = 0; l < L; ++l) {
        Tensor& t = h->x[l];
        t.cov_mr = t.masked ? pow2_at_least(2 * M) : pow2_at_least(M);
        if (t.masked) {
            // observed entries per row (constant during the fit): one counting row pass with zero weights
            RowPassArgs r{};
            r.g = t.gr_cnt;
            r.x_in = t.src;
            r.col_w = t.wkron;  // still all zero
            r.t_out = h->svec;
            r.tpart = t.tpart;
            r.cpart = t.cpart;
            r.rowcnt = t.rowcnt;
            r.epi = 0;
            r.div = 1.0;
            TRY(row_pass(h, t.dtype, 2, r));
            t.rowcnt_ready = true;
        }
    }

    for (int a = 0; a < R; ++a) {
        double* Ta = h->T + (size_t)a * n;
        double* Ua = h->U + (size_t)a * n;
        // ---- cross-covariance pass: centring (a == 0) or the deflation by component a-1 fused in ----
        for (int l = 0; l < L; ++l) {
            Tensor& t = h->x[l];
            CovPassArgs c{};
            c.g = t.gc;
            c.x_in = a == 0 ? t.src : t.work;
            c.x_out = t.work;
            c.row_a = a == 0 ? nullptr : h->T + (size_t)(a - 1) * n;
            c.col_w = a == 0 ? t.mean_d : t.wkron + (size_t)(a - 1) * t.pitch;
            c.row_sw = h->row_w;
            c.y = h->y_work;
            c.pitch_y = h->pitch_y;
            c.m = M;
            c.rowcnt = t.rowcnt;
            c.cpart = t.covpart;
            c.c_stride = (size_t)t.cov_mr * t.pitch;
            c.sspart = t.sspart_cov;
            {
                const double bytes = 2.0 * (double)n * t.pitch * t.elem;
                ProfScope ps(h, TPLS_K_DEFLATE_CONTRACT, bytes);
                CK(launch_covpass(t.dtype, t.masked, t.cov_mr, c, st));
                h->stats.kernel_launches++;
                h->stats.streamed_bytes += bytes;
            }
            TRY(reduce_cols(h, t.covpart, A + t.off_c, (int)c.c_stride, (int)c.c_stride, t.gc.grid_x, t.sspart_cov,
                            A + t.off_ss + a, t.gc.grid_x * t.gc.n_slabs, nullptr, 0));
        }
        // ---- Y'Y of the current Y for the stop test ----
        {
            ProfScope ps_small(h, TPLS_K_OTHER, 0.0);
            int gx = 1;
            CK(launch_gram_rows(h->y_work, n, h->pitch_y, M, h->grampart, &gx, st));
            h->stats.kernel_launches++;
            TRY(reduce_cols(h, h->grampart, A + h->off_gram_y, M * M, M * M, gx, nullptr, nullptr, 0, nullptr, 0));
        }
        TRY(allreduce(h, A + h->off_cov, h->cov_len));
        // ---- the whole inner iteration of this component, on the device ----
        {
            CovLoopArgs la{};
            la.n_tasks = L;
            la.m = M;
            la.gram_y = A + h->off_gram_y;
            la.tol = tol;
            la.max_iter = max_iter;
            la.normalize_on_break = (flags & TPLS_FIT_NORMALIZE_ON_BREAK) ? 1 : 0;
            la.q_out = h->Q + (size_t)a * M;
            la.qvec = h->qvec;
            la.pitch_y = h->pitch_y;
            la.trips_out = h->trips_dev + a;
            for (int l = 0; l < L; ++l) {
                Tensor& t = h->x[l];
                fill_rank1_task(h, t, a, la.t[l], r1_use_smem);
                la.t[l].z = t.zscratch;
                la.C[l] = A + t.off_c;
                la.masked[l] = t.masked ? t.cov_mr / 2 : 0;
            }
            ProfScope ps(h, TPLS_K_RANK1, 0.0);
            CK(launch_cov_loop(la, r1_smem, r1_use_smem, st));
            h->stats.kernel_launches++;
        }
        // ---- scores of the converged component: t = X x_2 w2 ..., u = Y q ----
        for (int l = 0; l < L; ++l) {
            Tensor& t = h->x[l];
            RowPassArgs r{};
            r.g = t.gr;
            r.x_in = t.work;
            r.col_w = t.wkron + (size_t)a * t.pitch;
            r.t_out = Ta;
            r.tpart = t.tpart;
            r.cpart = t.cpart;
            r.rowcnt = t.rowcnt;
            r.epi = l == 0 ? 0 : (l == L - 1 ? 2 : 1);
            r.div = (double)L;
            TRY(row_pass(h, t.dtype, t.masked ? 1 : 0, r));
        }
        {
            RowPassArgs r{};
            r.g = h->gy_row;
            r.x_in = h->y_work;
            r.col_w = h->qvec;
            r.t_out = Ua;
            r.epi = 0;
            r.div = 1.0;
            TRY(row_pass(h, TPLS_F64, 0, r, TPLS_K_YSIDE));
        }
        TRY(component_tail(h, R, a, ss_y, false));
    }
    // ---- residual norm after the last component ----
    for (int l = 0; l < L; ++l) {
        Tensor& t = h->x[l];
        ColPassArgs c{};
        c.g = t.g;
        c.x_in = t.work;
        c.x_out = t.work;
        c.row_a = h->T + (size_t)(R - 1) * n;
        c.col_w = t.wkron + (size_t)(R - 1) * t.pitch;
        c.sspart = t.sspart;
        c.row_sw = h->row_w;
        TRY(col_pass(h, t.dtype, t.masked, PF_DEFLATE | PF_SUMSQ, c));
        TRY(reduce_cols(h, nullptr, nullptr, 0, 0, 0, t.sspart, A + t.off_ss + R, t.g.grid_x * t.g.n_slabs, nullptr, 0));
    }
    return 0;
}

int tpls_fit(tpls_handle h, int n_tensors, int n_components, double tol, int max_iter, int flags) {
    if (!h) return fail(nullptr, "NULL handle");
    const int L = n_tensors, R = n_components;
    if (L < 1 || L > TPLS_MAX_TENSORS) return fail(h, "tpls_fit: n_tensors must be 1..%d", TPLS_MAX_TENSORS);
    if (R < 1 || R > 32) return fail(h, "tpls_fit: n_components must be 1..32");
    if (max_iter < 1 || max_iter > 4096) return fail(h, "tpls_fit: max_iter must be 1..4096");
    if (!h->y_src) return fail(h, "tpls_fit: Y not set");
    for (int l = 0; l < L; ++l) {
        if (!h->x[l].set) return fail(h, "tpls_fit: X[%d] not set", l);
        if (h->x[l].n != h->n)
            return fail(h, "tpls_fit: X[%d] has %lld samples, Y has %lld", l, h->x[l].n, h->n);
    }
    CK(cudaSetDevice(h->device));
    cudaStream_t st = h->stream;
    const double h2d = h->h2d_bytes;
    h->stats = tpls_stats{};
    h->stats.h2d_bytes = h2d;
    h->h2d_bytes = 0;
    h->profile = (flags & TPLS_FIT_PROFILE) != 0;
    h->cov_alloc = (flags & TPLS_FIT_COVARIANCE) != 0 && h->m <= 8;
    TRY(alloc_fit(h, L, R));
    CK(cudaEventRecord(h->ev_start, st));
    const long long n = h->n;
    double* A = h->arena;

    // ---- column statistics (np.nanmean, tpls.py:66-67) ----
    for (int l = 0; l < L; ++l) {
        Tensor& t = h->x[l];
        ColPassArgs c{};
        c.g = t.g;
        c.x_in = t.src;
        c.zpart = t.zpart;
        c.cntpart = t.cntpart;
        c.row_sw = h->row_w;
        TRY(col_pass(h, t.dtype, true, PF_COLSTAT, c));
        TRY(reduce_cols(h, t.zpart, A + t.off_colsum, t.pitch, t.pitch, t.g.grid_x, nullptr, nullptr, 0, nullptr, 0));
        TRY(reduce_cols(h, t.cntpart, A + t.off_colcnt, t.pitch, t.pitch, t.g.grid_x, nullptr, nullptr, 0, nullptr, 0));
    }
    {
        ColPassArgs c{};
        c.g = h->gy;
        c.x_in = h->y_src;
        c.zpart = h->zpart_y;
        c.cntpart = h->cntpart_y;
        c.row_sw = h->row_w;
        TRY(col_pass(h, TPLS_F64, true, PF_COLSTAT, c, TPLS_K_YSIDE));
        TRY(reduce_cols(h, h->zpart_y, A + h->off_ysum, h->pitch_y, h->pitch_y, h->gy.grid_x, nullptr, nullptr, 0, nullptr, 0));
        TRY(reduce_cols(h, h->cntpart_y, A + h->off_ycnt, h->pitch_y, h->pitch_y, h->gy.grid_x, nullptr, nullptr, 0, nullptr, 0));
        if (h->row_w == nullptr) {
            ProfScope ps_small(h, TPLS_K_OTHER, 0.0);
            CK(launch_fill(A + h->off_n, 1, (double)n, st));
            h->stats.kernel_launches++;
        } else {
            // sample count of the fold = sum of the 0/1 weights
            DotPairs d{};
            d.n = n;
            d.npairs = 1;
            d.a[0] = h->row_w;
            d.b[0] = h->row_w;
            int gx = 1;
            CK(launch_multi_dot(d, h->dotpart, &gx, st));
            h->stats.kernel_launches++;
            TRY(reduce_cols(h, h->dotpart, A + h->off_n, 1, 1, gx, nullptr, nullptr, 0, nullptr, 0));
        }
    }
    TRY(allreduce(h, A, h->off_stats_end));
    for (int l = 0; l < L; ++l) {
        Tensor& t = h->x[l];
        {
            ProfScope ps_small(h, TPLS_K_OTHER, 0.0);
            CK(launch_finalize_mean(t.dtype, A + t.off_colsum, A + t.off_colcnt, A + h->off_n, t.p, t.pitch, t.mean_d,
                                t.mean_native, t.miss_flag, st));
            h->stats.kernel_launches++;
        }
    }
    {
        ProfScope ps_small(h, TPLS_K_OTHER, 0.0);
        CK(launch_finalize_mean(TPLS_F64, A + h->off_ysum, A + h->off_ycnt, A + h->off_n, h->m, h->pitch_y, h->ymean_d,
                            nullptr, h->ymiss_flag, st));
        h->stats.kernel_launches++;
    }
    {
        int flagsh[TPLS_MAX_TENSORS] = {0};
        for (int l = 0; l < L; ++l)
            CK(cudaMemcpyAsync(&flagsh[l], h->x[l].miss_flag, sizeof(int), cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(&h->n_total, A + h->off_n, sizeof(double), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        for (int l = 0; l < L; ++l) {
            Tensor& t = h->x[l];
            t.masked = flagsh[l] != 0;
            t.gr = make_row_geom(n, t.p, t.pitch, t.elem, h->sm_count, false);
            t.gr_cnt = make_row_geom(n, t.p, t.pitch, t.elem, h->sm_count, true);
        }
    }

    // ---- centre Y (tpls.py:71), u0 = first column (tpls.py:78) ----
    double* ss_y = A + h->off_ss + (size_t)L * (R + 1);
    {
        ColPassArgs c{};
        c.g = h->gy;
        c.x_in = h->y_src;
        c.x_out = h->y_work;
        c.col_w = h->ymean_d;
        c.sspart = h->sspart_y;
        c.row_sw = h->row_w;
        TRY(col_pass(h, TPLS_F64, false, PF_DEFLATE | PF_WRITE | PF_SUMSQ, c, TPLS_K_YSIDE));
        TRY(reduce_cols(h, nullptr, nullptr, 0, 0, 0, h->sspart_y, ss_y, h->gy.grid_x * h->gy.n_slabs, nullptr, 0));
        if (h->row_w != nullptr) {
            // held-out rows of the centred Y are zeroed: they then drop out of u, Z, q and the stop test
            CK(launch_scale_rows(h->y_work, n, h->pitch_y, h->row_w, st));
            h->stats.kernel_launches++;
        }
        {
            ProfScope ps_small(h, TPLS_K_OTHER, 0.0);
            CK(launch_gather_col(h->y_work, n, h->pitch_y, 0, h->U, st));
            h->stats.kernel_launches++;
        }
    }
    const bool cov_mode = (flags & TPLS_FIT_COVARIANCE) != 0 && cov_supported(h, L);
    h->stats.covariance_mode = cov_mode ? 1 : 0;
    if (cov_mode) {
        TRY(fit_covariance(h, L, R, tol, max_iter, flags, ss_y));
    } else {
    // ---- centre X fused with the first contraction ----
    for (int l = 0; l < L; ++l) {
        Tensor& t = h->x[l];
        ColPassArgs c{};
        c.g = t.g;
        c.x_in = t.src;
        c.x_out = t.work;
        c.col_w = t.mean_d;
        c.row_u = h->U;
        c.zpart = t.zpart;
        c.sspart = t.sspart;
        c.row_sw = h->row_w;
        TRY(col_pass(h, t.dtype, t.masked, PF_DEFLATE | PF_WRITE | PF_CONTRACT | PF_SUMSQ, c));
        TRY(reduce_cols(h, t.zpart, A + t.off_z, t.pitch, t.pitch, t.g.grid_x, t.sspart, A + t.off_ss,
                        t.g.grid_x * t.g.n_slabs, nullptr, 0));
    }

    // rank-1 launch configuration
    size_t r1_smem = 0;
    bool r1_use_smem = true;
    for (int l = 0; l < L; ++l) r1_smem = std::max(r1_smem, h->x[l].r1_ws * sizeof(double));
    if (r1_smem > 200 * 1024) {
        r1_use_smem = false;
        r1_smem = 0;
    }

    const int LOOK = 1;
    const bool fused_xchg = h->world > 1 && h->xchg_ready && h->zcat_len <= (size_t)h->xchg_cap;
    const bool gram_stop = h->m <= 8 && getenv("TPLS_NO_GRAM_STOP") == nullptr;
    for (int a = 0; a < R; ++a) {
        double* Ta = h->T + (size_t)a * n;
        double* Ua = h->U + (size_t)a * n;
        {
            ProfScope ps_small(h, TPLS_K_OTHER, 0.0);
            CK(launch_reset_ctrl(h->ctrl, st));
            h->stats.kernel_launches++;
        }
        if (gram_stop) {
            // Y'Y of the current (deflated) Y: the stop test then needs no pass over the samples and no
            // collective of its own (||u_old - u_new||^2 = dq^T Y'Y dq, u = Y q)
            int gx = 1;
            {
                ProfScope ps_small(h, TPLS_K_OTHER, 0.0);
                CK(launch_gram_rows(h->y_work, n, h->pitch_y, h->m, h->grampart, &gx, st));
                h->stats.kernel_launches++;
            }
            TRY(reduce_cols(h, h->grampart, A + h->off_gram_y, h->m * h->m, h->m * h->m, gx, nullptr, nullptr, 0, nullptr, 0));
            TRY(allreduce(h, A + h->off_gram_y, (size_t)h->m * h->m));
        }
        for (int trip = 0; trip < max_iter; ++trip) {
            if (trip >= 1 + LOOK) {
                const int back = trip - 1 - LOOK;
                CK(cudaEventSynchronize(h->ev_trip[back & 3]));
                if (h->h_done[back] >= 0) break;
            }
            // K1: Z = X x_1 u (trip 0 got it from the fused centring / deflation pass)
            if (trip > 0) {
                for (int l = 0; l < L; ++l) {
                    Tensor& t = h->x[l];
                    ColPassArgs c{};
                    c.g = t.g;
                    c.x_in = t.work;
                    c.row_u = Ua;
                    c.zpart = t.zpart;
                    c.ctrl = h->ctrl;
                    c.trip = trip;
                    TRY(col_pass(h, t.dtype, t.masked, PF_CONTRACT, c));
                    if (!fused_xchg)
                        TRY(reduce_cols(h, t.zpart, A + t.off_z, t.pitch, t.pitch, t.g.grid_x, nullptr, nullptr, 0, h->ctrl, trip));
                }
            }
            if (fused_xchg && trip > 0) {
                // second reduction stage of every tensor + the cross-GPU sum in ONE kernel
                XchgArgs xa{};
                xa.n_sets = L;
                for (int l = 0; l < L; ++l) {
                    Tensor& t = h->x[l];
                    xa.sets[l] = XchgSet{t.zpart, t.g.grid_x, t.pitch, t.pitch, (int)(t.off_z - h->off_zcat)};
                }
                xa.out = A + h->off_zcat;
                xa.count = (int)h->zcat_len;
                TRY(xchg_launch(h, xa));
            } else {
                TRY(allreduce(h, A + h->off_zcat, h->zcat_len));
            }
            // K3: rank-1 weight vectors, one CTA per tensor
            {
                Rank1Args ra{};
                ra.n_tasks = L;
                ra.tol = tol;
                ra.normalize_on_break = (flags & TPLS_FIT_NORMALIZE_ON_BREAK) ? 1 : 0;
                ra.ctrl = h->ctrl;
                ra.trip = trip;
                for (int l = 0; l < L; ++l) fill_rank1_task(h, h->x[l], a, ra.t[l], r1_use_smem);
                ProfScope ps(h, TPLS_K_RANK1, 0.0);
                CK(launch_rank1(ra, r1_smem, st));
                h->stats.kernel_launches++;
            }
            // K2: t = X x_2 w2 x_3 w3 ..., averaged over the coupled tensors (cmtf.py:120)
            for (int l = 0; l < L; ++l) {
                Tensor& t = h->x[l];
                RowPassArgs r{};
                const int rmode = !t.masked ? 0 : (t.rowcnt_ready ? 1 : 2);
                r.g = rmode == 2 ? t.gr_cnt : t.gr;
                r.x_in = t.work;
                r.col_w = t.wkron + (size_t)a * t.pitch;
                r.t_out = Ta;
                r.tpart = t.tpart;
                r.cpart = t.cpart;
                r.rowcnt = t.rowcnt;
                r.epi = l == 0 ? 0 : (l == L - 1 ? 2 : 1);
                r.div = (double)L;
                r.ctrl = h->ctrl;
                r.trip = trip;
                // the first projection of a masked tensor also counts the observed entries of every row
                // (trip 0 of component 0 always executes, so the counts are there for every later trip)
                TRY(row_pass(h, t.dtype, rmode, r));
                t.rowcnt_ready = true;
            }
            // K4: q = Y't / ||.||, u = Y q, stop test (tpls.py:100-107)
            {
                ColPassArgs c{};
                c.g = h->gy;
                c.x_in = h->y_work;
                c.row_u = Ta;
                c.zpart = h->zpart_y;
                c.ctrl = h->ctrl;
                c.trip = trip;
                TRY(col_pass(h, TPLS_F64, false, PF_CONTRACT, c, TPLS_K_YSIDE));
                // with the Y'Y stop test the normalisation of q and the stop decision ride on the kernel that
                // finishes the q reduction (the exchange kernel, or the second reduction stage on one GPU)
                const bool q_fused = gram_stop && h->pitch_y <= 32 && (fused_xchg || h->world == 1);
                if (fused_xchg) {
                    XchgArgs xa{};
                    xa.n_sets = 1;
                    xa.sets[0] = XchgSet{h->zpart_y, h->gy.grid_x, h->pitch_y, h->pitch_y, 0};
                    xa.out = A + h->off_q;
                    xa.count = h->pitch_y;
                    if (q_fused) {
                        xa.do_qstop = 1;
                        xa.q_m = h->m;
                        xa.q_pitch = h->pitch_y;
                        xa.qcol = h->Q + (size_t)a * h->m;
                        xa.qvec = h->qvec;
                        xa.gram = A + h->off_gram_y;
                        xa.q_prev = h->q_prev;
                        xa.ctrl = h->ctrl;
                        xa.trip = trip;
                        xa.tol = tol;
                    }
                    TRY(xchg_launch(h, xa));
                } else if (q_fused) {
                    ProfScope ps_small(h, TPLS_K_OTHER, 0.0);
                    CK(launch_reduce_q_stop(h->zpart_y, h->gy.grid_x, h->pitch_y, A + h->off_q, h->m, h->pitch_y,
                                            h->Q + (size_t)a * h->m, h->qvec, A + h->off_gram_y, h->q_prev, h->ctrl, trip, tol, st));
                    h->stats.kernel_launches++;
                } else {
                    TRY(reduce_cols(h, h->zpart_y, A + h->off_q, h->pitch_y, h->pitch_y, h->gy.grid_x, nullptr, nullptr, 0, h->ctrl, trip));
                    TRY(allreduce(h, A + h->off_q, h->pitch_y));
                }
                if (!q_fused) {
                    ProfScope ps_small(h, TPLS_K_OTHER, 0.0);
                    if (gram_stop)
                        CK(launch_normalize_q_stop(A + h->off_q, h->m, h->pitch_y, h->Q + (size_t)a * h->m, h->qvec,
                                                   A + h->off_gram_y, h->q_prev, h->ctrl, trip, tol, st));
                    else
                        CK(launch_normalize_q(A + h->off_q, h->m, h->pitch_y, h->Q + (size_t)a * h->m, h->qvec, h->ctrl, trip, st));
                    h->stats.kernel_launches++;
                }
                RowPassArgs r{};
                r.g = h->gy_row;
                r.x_in = h->y_work;
                r.col_w = h->qvec;
                r.t_out = Ua;
                r.epi = 0;
                r.div = 1.0;
                r.d2part = gram_stop ? nullptr : h->d2part;
                r.ctrl = h->ctrl;
                r.trip = trip;
                TRY(row_pass(h, TPLS_F64, 0, r, TPLS_K_YSIDE));
                const int nd2 = d2_grid(h->gy_row);
                if (gram_stop) {
                    // the stop test already ran inside normalize_q_stop
                } else if (fused_xchg) {
                    // partial sums of ||u_old - u_new||^2 -> global sum -> stop test, one kernel
                    XchgArgs xa{};
                    xa.n_sets = 1;
                    xa.sets[0] = XchgSet{h->d2part, nd2, 1, 1, 0};
                    xa.out = A + h->off_d2;
                    xa.count = 1;
                    xa.ctrl = h->ctrl;
                    xa.trip = trip;
                    xa.tol = tol;
                    xa.do_stop = 1;
                    TRY(xchg_launch(h, xa));
                } else if (h->world > 1) {
                    {
                        ProfScope ps_small(h, TPLS_K_OTHER, 0.0);
                        CK(launch_sum_small(h->d2part, nd2, A + h->off_d2, h->ctrl, trip, st));
                        h->stats.kernel_launches++;
                    }
                    TRY(allreduce(h, A + h->off_d2, 1));
                    CK(launch_stop(h->ctrl, trip, A + h->off_d2, 1, tol, st));
                    h->stats.kernel_launches++;
                } else {
                    CK(launch_stop(h->ctrl, trip, h->d2part, nd2, tol, st));
                    h->stats.kernel_launches++;
                }
            }
            CK(cudaMemcpyAsync(&h->h_done[trip], &h->ctrl->done_trip, sizeof(int), cudaMemcpyDeviceToHost, st));
            CK(cudaEventRecord(h->ev_trip[trip & 3], st));
        }

        TRY(component_tail(h, R, a, ss_y, true));
        // ---- X deflation (tpls.py:109) fused with the next component's first contraction ----
        if (a + 1 < R) {
            {
                ProfScope ps_small(h, TPLS_K_OTHER, 0.0);
                CK(launch_gather_col(h->y_work, n, h->pitch_y, 0, h->U + (size_t)(a + 1) * n, st));
                h->stats.kernel_launches++;
            }
        }
        for (int l = 0; l < L; ++l) {
            Tensor& t = h->x[l];
            ColPassArgs c{};
            c.g = t.g;
            c.x_in = t.work;
            c.x_out = t.work;
            c.row_a = Ta;
            c.col_w = t.wkron + (size_t)a * t.pitch;
            c.sspart = t.sspart;
            c.row_sw = h->row_w;
            if (a + 1 < R) {
                c.row_u = h->U + (size_t)(a + 1) * n;
                c.zpart = t.zpart;
                TRY(col_pass(h, t.dtype, t.masked, PF_DEFLATE | PF_WRITE | PF_CONTRACT | PF_SUMSQ, c));
                TRY(reduce_cols(h, t.zpart, A + t.off_z, t.pitch, t.pitch, t.g.grid_x, t.sspart, A + t.off_ss + a + 1,
                                t.g.grid_x * t.g.n_slabs, nullptr, 0));
            } else {
                TRY(col_pass(h, t.dtype, t.masked, PF_DEFLATE | PF_SUMSQ, c));
                TRY(reduce_cols(h, nullptr, nullptr, 0, 0, 0, t.sspart, A + t.off_ss + a + 1, t.g.grid_x * t.g.n_slabs, nullptr, 0));
            }
        }
    }

    }  // streaming mode

    // ---- R2X / R2Y from the residual norms ----
    // through NCCL on purpose: it cannot complete before every peer has finished all earlier exchanges,
    // so no rank can leave the fit (and possibly free its exchange buffer) while a peer still reads it
    TRY(allreduce_nccl(h, A + h->off_ss, h->ss_len));
    std::vector<double> ss(h->ss_len);
    h->trips.assign(R, 0);
    CK(cudaMemcpyAsync(ss.data(), A + h->off_ss, sizeof(double) * h->ss_len, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(h->trips.data(), h->trips_dev, sizeof(int) * R, cudaMemcpyDeviceToHost, st));
    CK(cudaEventRecord(h->ev_stop, st));
    CK(cudaStreamSynchronize(st));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, h->ev_start, h->ev_stop));
    h->stats.fit_ms = ms;
    if (h->xchg_ready) {
        int xerr = 0;
        CK(cudaMemcpy(&xerr, static_cast<char*>(h->xchg_buf) + 520, sizeof(int), cudaMemcpyDeviceToHost));
        if (xerr) return fail(h, "tpls_fit: a peer-memory exchange timed out (a rank died or fell out of step)");
    }
    for (int l = 0; l < L; ++l) {
        h->r2x[l].assign(R, 0.0);
        const double* s = ss.data() + (h->x[l].off_ss - h->off_ss);
        for (int a = 0; a < R; ++a) h->r2x[l][a] = 1.0 - s[a + 1] / s[0];
    }
    h->r2y.assign(R, 0.0);
    {
        const double* s = ss.data() + (size_t)L * (R + 1);
        for (int a = 0; a < R; ++a) h->r2y[a] = 1.0 - s[a + 1] / s[0];
    }
    long long total = 0;
    for (int a = 0; a < R; ++a) total += h->trips[a];
    h->stats.total_trips = total;
    double alg = 0.0;
    for (int l = 0; l < L; ++l) alg += (double)h->x[l].elem * n * h->x[l].p * (2.0 * total + R + 2);
    h->stats.alg_bytes = alg;
    h->fitted = true;
    return 0;
}

// ---------------------------------------------------------------------------
// getters
// ---------------------------------------------------------------------------

#define NEED_FIT()                                     \
    if (!h) return fail(nullptr, "NULL handle");       \
    if (!h->fitted) return fail(h, "not fitted")

int tpls_get_x_factor(tpls_handle h, int index, int mode, double* out) {
    NEED_FIT();
    if (index < 0 || index >= h->L) return fail(h, "tensor index %d out of range", index);
    Tensor& t = h->x[index];
    if (mode < 0 || mode >= t.ndim) return fail(h, "mode %d out of range", mode);
    if (mode == 0) return copy_out_transposed(h, h->T, h->n, h->R, out);
    return copy_out_transposed(h, t.W[mode], t.shape[mode], h->R, out);
}

int tpls_get_y_factor(tpls_handle h, int which, double* out) {
    NEED_FIT();
    if (which == 0) return copy_out_transposed(h, h->U, h->n, h->R, out);
    if (which == 1) return copy_out_transposed(h, h->Q, h->m, h->R, out);
    return fail(h, "which must be 0 (U) or 1 (Q)");
}

int tpls_get_coef(tpls_handle h, double* out) {
    NEED_FIT();
    CK(cudaMemcpyAsync(out, h->coef, sizeof(double) * h->R * h->R, cudaMemcpyDefault, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}

int tpls_get_r2x(tpls_handle h, int index, double* out) {
    NEED_FIT();
    if (index < 0 || index >= h->L) return fail(h, "tensor index %d out of range", index);
    memcpy(out, h->r2x[index].data(), sizeof(double) * h->R);
    return 0;
}

int tpls_get_r2y(tpls_handle h, double* out) {
    NEED_FIT();
    memcpy(out, h->r2y.data(), sizeof(double) * h->R);
    return 0;
}

int tpls_get_x_mean(tpls_handle h, int index, void* out) {
    NEED_FIT();
    if (index < 0 || index >= h->L) return fail(h, "tensor index %d out of range", index);
    Tensor& t = h->x[index];
    CK(cudaMemcpyAsync(out, t.mean_native, (size_t)t.elem * t.p, cudaMemcpyDefault, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}

int tpls_get_y_mean(tpls_handle h, double* out) {
    NEED_FIT();
    CK(cudaMemcpyAsync(out, h->ymean_d, sizeof(double) * h->m, cudaMemcpyDefault, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}

int tpls_get_has_missing(tpls_handle h, int index, int* out) {
    NEED_FIT();
    if (index < 0 || index >= h->L) return fail(h, "tensor index %d out of range", index);
    *out = h->x[index].masked ? 1 : 0;
    return 0;
}

int tpls_get_trips(tpls_handle h, int* out) {
    NEED_FIT();
    memcpy(out, h->trips.data(), sizeof(int) * h->R);
    return 0;
}

int tpls_get_profile(tpls_handle h, tpls_profile* out) {
    if (h && !h->prof.empty()) {
        CK(cudaSetDevice(h->device));
        CK(cudaStreamSynchronize(h->stream));
        prof_collect(h);
    }
    if (!h) return fail(nullptr, "NULL handle");
    *out = h->prof_sum;
    return 0;
}

int tpls_get_stats(tpls_handle h, tpls_stats* out) {
    if (!h) return fail(nullptr, "NULL handle");
    *out = h->stats;
    return 0;
}

// ---------------------------------------------------------------------------
// transform
// ---------------------------------------------------------------------------
int tpls_release_data(tpls_handle h) {
    if (!h) return fail(nullptr, "NULL handle");
    CK(cudaSetDevice(h->device));
    CK(cudaStreamSynchronize(h->stream));
    for (auto& t : h->x) {
        if (t.owned) pool_put(h, t.owned);
        if (t.owned2) pool_put(h, t.owned2);
        t.owned = t.owned2 = nullptr;
        t.src = nullptr;
        t.work = nullptr;
        t.set = false;
    }
    if (h->y_src) pool_put(h, h->y_src);
    if (h->y_work) pool_put(h, h->y_work);
    if (h->row_w) pool_put(h, h->row_w);
    h->y_src = h->y_work = h->row_w = nullptr;
    return 0;
}

int tpls_set_row_weights(tpls_handle h, const double* w, int64_t n) {
    if (!h) return fail(nullptr, "NULL handle");
    CK(cudaSetDevice(h->device));
    if (h->row_w) pool_put(h, h->row_w);
    h->row_w = nullptr;
    if (w == nullptr) return 0;
    if (n != h->n) return fail(h, "tpls_set_row_weights: %lld weights for %lld samples (call tpls_set_y first)", (long long)n, h->n);
    TRY(pool_get(h, (void**)&h->row_w, sizeof(double) * n));
    CK(cudaMemcpyAsync(h->row_w, w, sizeof(double) * n, cudaMemcpyDefault, h->stream));
    h->fitted = false;
    return 0;
}

int tpls_trim(tpls_handle h) {
    if (!h) return fail(nullptr, "NULL handle");
    CK(cudaSetDevice(h->device));
    CK(cudaStreamSynchronize(h->stream));
    pool_trim(h);
    return 0;
}

}  // extern "C"
