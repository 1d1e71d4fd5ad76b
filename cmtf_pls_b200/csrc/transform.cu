// transform / predict on new data (reference cmtf_pls/tpls.py:122-186, cmtf.py:142-231): the scores of new
// samples through the stored loadings -- see include/tpls_b200.h (tpls_transform).
#include "driver_internal.cuh"

extern "C" {

int tpls_transform(tpls_handle h, int n_tensors, int n_components, const void* const* xs, const int* dtypes,
                   int64_t n_new, const int64_t* ps, const void* const* means, const double* const* wkrons,
                   const double* proj_offset, const double* proj_gram, double* scores_out) {
    if (!h) return fail(nullptr, "NULL handle");
    const int L = n_tensors, R = n_components;
    if (L < 1 || L > TPLS_MAX_TENSORS || R < 1 || R > 64 || n_new <= 0) return fail(h, "tpls_transform: bad sizes");
    CK(cudaSetDevice(h->device));
    NvtxRange nvtx_tr("tpls_transform");
    cudaStream_t st = h->stream;
    h->stats.kernel_launches = 0;
    h->stats.streamed_bytes = 0;
    std::vector<void*> tmp;
    double *S = nullptr, *sspart = nullptr, *pc = nullptr, *pg = nullptr;
    int* flag = nullptr;
    int rc = 0;
    std::vector<const void*> src(L, nullptr);  // the new data where it can be read in place, else its staged copy
    std::vector<void*> xw(L, nullptr);         // writable working copy (only the sequential path needs it)
    std::vector<PassGeom> gs(L), grs(L), grs_cnt(L);
    std::vector<double*> tpart(L, nullptr), cpart(L, nullptr), mean_d(L, nullptr), wk(L, nullptr), rowcnt(L, nullptr);
    std::vector<int> pitch(L), elem(L);
    auto stage = [&](int l) -> int {  // pitched private copy of tensor l
        if (xw[l]) return 0;
        const size_t bytes = (size_t)n_new * pitch[l] * elem[l];
        TRY(dev_alloc(h, &xw[l], bytes, &tmp));
        if (pitch[l] != (int)ps[l]) CK(cudaMemsetAsync(xw[l], 0, bytes, st));
        CK(cudaMemcpy2DAsync(xw[l], (size_t)pitch[l] * elem[l], xs[l], (size_t)ps[l] * elem[l], (size_t)ps[l] * elem[l], n_new,
                             cudaMemcpyDefault, st));
        return 0;
    };
    do {
        if ((rc = dev_alloc(h, (void**)&S, sizeof(double) * n_new * R, &tmp))) break;
        if ((rc = dev_alloc(h, (void**)&sspart, sizeof(double) * 4096, &tmp))) break;
        if ((rc = dev_alloc(h, (void**)&pc, sizeof(double) * (R + R * R), &tmp))) break;
        if ((rc = dev_alloc(h, (void**)&flag, sizeof(int) * 4, &tmp))) break;
        pg = pc + R;
        for (int l = 0; l < L && !rc; ++l) {
            if (dtypes[l] != TPLS_F32 && dtypes[l] != TPLS_F64) {
                rc = fail(h, "tpls_transform: bad dtype for X[%d]", l);
                break;
            }
            elem[l] = dtypes[l] == TPLS_F32 ? 4 : 8;
            const int vec = 16 / elem[l];
            const int p = (int)ps[l];
            pitch[l] = (p + vec - 1) / vec * vec;
            if ((rc = dev_alloc(h, (void**)&mean_d[l], sizeof(double) * pitch[l], &tmp))) break;
            if ((rc = dev_alloc(h, (void**)&wk[l], sizeof(double) * pitch[l] * R, &tmp))) break;
            if ((rc = dev_alloc(h, (void**)&rowcnt[l], sizeof(double) * n_new, &tmp))) break;
            void* mean_nat = nullptr;
            if ((rc = dev_alloc(h, &mean_nat, (size_t)elem[l] * pitch[l], &tmp))) break;
            cudaError_t e = cudaMemsetAsync(wk[l], 0, sizeof(double) * pitch[l] * R, st);
            if (e == cudaSuccess) e = cudaMemsetAsync(mean_nat, 0, (size_t)elem[l] * pitch[l], st);
            if (e == cudaSuccess) e = cudaMemcpyAsync(mean_nat, means[l], (size_t)elem[l] * p, cudaMemcpyDefault, st);
            if (e == cudaSuccess)
                e = cudaMemcpy2DAsync(wk[l], sizeof(double) * pitch[l], wkrons[l], sizeof(double) * p, sizeof(double) * p, R,
                                      cudaMemcpyDefault, st);
            if (e == cudaSuccess) e = launch_widen(dtypes[l], mean_nat, mean_d[l], pitch[l], st);
            if (e != cudaSuccess) {
                rc = fail(h, "tpls_transform: staging X[%d] -> %s", l, cudaGetErrorString(e));
                break;
            }
            if (is_device_ptr(xs[l]) && pitch[l] == p && ((uintptr_t)xs[l] % 16 == 0)) {
                src[l] = xs[l];
            } else {
                if ((rc = stage(l))) break;
                src[l] = xw[l];
            }
            gs[l] = make_geom(n_new, p, pitch[l], elem[l], h->sm_count);
            grs[l] = make_row_geom(n_new, p, pitch[l], elem[l], h->sm_count, false);
            grs_cnt[l] = make_row_geom(n_new, p, pitch[l], elem[l], h->sm_count, true);
            if (gs[l].n_slabs > 1) {
                if ((rc = dev_alloc(h, (void**)&tpart[l], sizeof(double) * n_new * gs[l].n_slabs, &tmp))) break;
                if ((rc = dev_alloc(h, (void**)&cpart[l], sizeof(double) * n_new * gs[l].n_slabs, &tmp))) break;
            }
        }
        if (rc) break;

        // ---- single-pass path for complete data: ALL R raw projections of the untouched rows in ONE read of X
        //      (multiproj.cu, fp64 tensor-core path), then the deflation recurrence on the scores alone.  With
        //      x_c = x - mean and no NaN,
        //        t_a = mean_l (x_c - sum_{b<a} t_b w_lb) . w_la = r_a - c_a - sum_{b<a} t_b G_ba,
        //      r_a = mean_l x . w_la,  c_a = mean_l mean_l . w_la,  G_ba = mean_l w_lb . w_la  (passed in).
        //      A NaN poisons the raw projections of its row; the finishing kernel reports it and the call falls
        //      through to the sequential masked path.
        bool done = false;
        if (proj_offset != nullptr && proj_gram != nullptr && R <= 32) {
            cudaError_t e = cudaMemcpyAsync(pc, proj_offset, sizeof(double) * R, cudaMemcpyDefault, st);
            if (e == cudaSuccess) e = cudaMemcpyAsync(pg, proj_gram, sizeof(double) * R * R, cudaMemcpyDefault, st);
            if (e == cudaSuccess) e = cudaMemsetAsync(flag, 0, sizeof(int) * 4, st);
            if (e != cudaSuccess) {
                rc = fail(h, "tpls_transform: %s", cudaGetErrorString(e));
                break;
            }
            MultiProjFinishArgs f{};
            f.n_tensors = L;
            f.n_comp = R;
            f.n_rows = n_new;
            f.c = pc;
            f.G = pg;
            f.S = S;
            f.flag = flag;
            for (int l = 0; l < L && !rc; ++l) {
                const int ns = multiproj_slabs(dtypes[l], R, pitch[l]);
                double* part = nullptr;
                if ((rc = dev_alloc(h, (void**)&part, sizeof(double) * (size_t)ns * n_new * R, &tmp))) break;
                MultiProjArgs m{};
                m.x = src[l];
                m.n_rows = n_new;
                m.p = (int)ps[l];
                m.pitch = pitch[l];
                m.w = wk[l];
                m.w_pitch = pitch[l];
                m.n_comp = R;
                m.part = part;
                const double bytes = (double)n_new * pitch[l] * elem[l];
                {
                    ProfScope psx(h, TPLS_K_PROJECT, bytes);
                    cudaError_t e2 = launch_multiproj(dtypes[l], m, h->sm_count, st);
                    if (e2 != cudaSuccess) rc = fail(h, "multiproj -> %s", cudaGetErrorString(e2));
                }
                h->stats.kernel_launches++;
                h->stats.streamed_bytes += bytes;
                f.part[l] = part;
                f.n_slabs[l] = ns;
            }
            if (rc) break;
            cudaError_t e4 = launch_multiproj_finish(f, st);
            if (e4 != cudaSuccess) {
                rc = fail(h, "multiproj_finish -> %s", cudaGetErrorString(e4));
                break;
            }
            h->stats.kernel_launches++;
            int incomplete = 1;
            cudaError_t e3 = cudaMemcpyAsync(&incomplete, flag, sizeof(int), cudaMemcpyDeviceToHost, st);
            if (e3 == cudaSuccess) e3 = cudaStreamSynchronize(st);
            if (e3 != cudaSuccess) {
                rc = fail(h, "tpls_transform: %s", cudaGetErrorString(e3));
                break;
            }
            done = !incomplete;
        }

        // ---- sequential path (NaNs present, or no projection constants): centre a private copy, then per
        //      component project (masked) and deflate with the stored loadings (tpls.py:151-165) ----
        if (!done) {
            for (int l = 0; l < L && !rc; ++l) {
                if ((rc = stage(l))) break;
                ColPassArgs c{};
                c.g = gs[l];
                c.x_in = xw[l];
                c.x_out = xw[l];
                c.col_w = mean_d[l];
                c.sspart = sspart;
                rc = col_pass(h, dtypes[l], true, PF_DEFLATE | PF_WRITE | PF_SUMSQ, c);
            }
            for (int a = 0; a < R && !rc; ++a) {
                double* Sa = S + (size_t)a * n_new;
                for (int l = 0; l < L && !rc; ++l) {
                    RowPassArgs r{};
                    r.g = a == 0 ? grs_cnt[l] : grs[l];
                    r.x_in = xw[l];
                    r.col_w = wk[l] + (size_t)a * pitch[l];
                    r.t_out = Sa;
                    r.tpart = tpart[l];
                    r.cpart = cpart[l];
                    r.rowcnt = rowcnt[l];
                    r.epi = l == 0 ? 0 : (l == L - 1 ? 2 : 1);
                    r.div = (double)L;
                    rc = row_pass(h, dtypes[l], a == 0 ? 2 : 1, r);
                }
                for (int l = 0; l < L && !rc && a + 1 < R; ++l) {
                    ColPassArgs c{};
                    c.g = gs[l];
                    c.x_in = xw[l];
                    c.x_out = xw[l];
                    c.row_a = Sa;
                    c.col_w = wk[l] + (size_t)a * pitch[l];
                    c.sspart = sspart;
                    rc = col_pass(h, dtypes[l], true, PF_DEFLATE | PF_WRITE | PF_SUMSQ, c);
                }
            }
        }
        if (rc) break;
        h->stats.last_transform_path = done ? 1 : 2;
        rc = copy_out_transposed(h, S, n_new, R, scores_out);
    } while (0);
    cudaStreamSynchronize(st);
    for (void* p : tmp) pool_put(h, p);
    return rc;
}

int tpls_reconstruct(tpls_handle h, int n_components, const double* scores, int64_t n, int64_t p, const double* wkron,
                     const void* mean, int mean_dtype, double* out) {
    if (!h) return fail(nullptr, "NULL handle");
    const int R = n_components;
    if (R < 1 || R > 32 || n <= 0 || p <= 0 || p > (1ll << 30)) return fail(h, "tpls_reconstruct: bad sizes");
    if (mean != nullptr && mean_dtype != TPLS_F32 && mean_dtype != TPLS_F64) return fail(h, "tpls_reconstruct: bad mean dtype");
    CK(cudaSetDevice(h->device));
    NvtxRange nvtx_rc("tpls_reconstruct");
    cudaStream_t st = h->stream;
    std::vector<void*> tmp;
    int rc = 0;
    do {
        double *T = nullptr, *W = nullptr, *mean_d = nullptr, *blk = nullptr;
        void* mean_nat = nullptr;
        if ((rc = dev_alloc(h, (void**)&T, sizeof(double) * n * R, &tmp))) break;
        if ((rc = dev_alloc(h, (void**)&W, sizeof(double) * p * R, &tmp))) break;
        cudaError_t e = cudaMemcpyAsync(T, scores, sizeof(double) * n * R, cudaMemcpyDefault, st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(W, wkron, sizeof(double) * p * R, cudaMemcpyDefault, st);
        if (e == cudaSuccess && mean != nullptr) {
            const size_t me = mean_dtype == TPLS_F32 ? 4 : 8;
            if ((rc = dev_alloc(h, (void**)&mean_d, sizeof(double) * p, &tmp))) break;
            if ((rc = dev_alloc(h, &mean_nat, me * p, &tmp))) break;
            e = cudaMemcpyAsync(mean_nat, mean, me * p, cudaMemcpyDefault, st);
            if (e == cudaSuccess) e = launch_widen(mean_dtype, mean_nat, mean_d, (int)p, st);
        }
        if (e != cudaSuccess) {
            rc = fail(h, "tpls_reconstruct: staging -> %s", cudaGetErrorString(e));
            break;
        }
        const bool out_dev = is_device_ptr(out);
        // A host result is produced block by block: two device blocks and two pinned bounce buffers (kept in the
        // handle), so that the writer of block k+1, the DMA of block k and the host copy of block k-1 overlap --
        // a pageable destination would otherwise be fed through the driver's own small staging buffers.
        const size_t kBounce = 32u << 20;
        const long long blk_rows = out_dev ? n : std::max<long long>(1, std::min<long long>(n, (long long)(kBounce / (8 * p))));
        if (!out_dev) {
            if ((size_t)blk_rows * p * 8 > kBounce) {  // a single row block larger than a bounce buffer (huge p)
                rc = fail(h, "tpls_reconstruct: p = %lld too large for a host result", (long long)p);
                break;
            }
            if ((rc = dev_alloc(h, (void**)&blk, sizeof(double) * 2 * blk_rows * p, &tmp))) break;
            for (int q = 0; q < 2 && !rc; ++q) {
                if (!h->bounce[q] && cudaMallocHost(&h->bounce[q], kBounce) != cudaSuccess) rc = fail(h, "cudaMallocHost failed");
                if (!h->bounce_ev[q] && cudaEventCreateWithFlags(&h->bounce_ev[q], cudaEventDisableTiming) != cudaSuccess)
                    rc = fail(h, "cudaEventCreate failed");
            }
            if (rc) break;
        }
        long long pend_r0[2] = {-1, -1}, pend_rows[2] = {0, 0};
        auto drain = [&](int q) -> cudaError_t {   // host copy of the block that last used bounce buffer q
            if (pend_r0[q] < 0) return cudaSuccess;
            cudaError_t e3 = cudaEventSynchronize(h->bounce_ev[q]);
            if (e3 == cudaSuccess) memcpy(out + pend_r0[q] * p, h->bounce[q], sizeof(double) * pend_rows[q] * p);
            pend_r0[q] = -1;
            return e3;
        };
        int q = 0;
        for (long long r0 = 0; r0 < n && !rc; r0 += blk_rows, q ^= 1) {
            const long long rows = std::min<long long>(blk_rows, n - r0);
            cudaError_t e2 = out_dev ? cudaSuccess : drain(q);
            ReconstructArgs a{};
            a.T = T + r0 * R;
            a.w = W;
            a.mean = mean_d;
            a.n_rows = rows;
            a.p = (int)p;
            a.w_pitch = (int)p;
            a.n_comp = R;
            a.out = out_dev ? out + r0 * p : blk + (size_t)q * blk_rows * p;
            if (e2 == cudaSuccess) e2 = launch_reconstruct(a, st);
            h->stats.kernel_launches++;
            if (e2 == cudaSuccess && !out_dev) {
                e2 = cudaMemcpyAsync(h->bounce[q], a.out, sizeof(double) * rows * p, cudaMemcpyDeviceToHost, st);
                if (e2 == cudaSuccess) e2 = cudaEventRecord(h->bounce_ev[q], st);
                pend_r0[q] = r0;
                pend_rows[q] = rows;
            }
            if (e2 != cudaSuccess) rc = fail(h, "tpls_reconstruct -> %s", cudaGetErrorString(e2));
        }
        for (int k = 0; k < 2 && !rc && !out_dev; ++k)
            if (drain(k) != cudaSuccess) rc = fail(h, "tpls_reconstruct: device-to-host copy failed");
    } while (0);
    cudaStreamSynchronize(st);
    for (void* q : tmp) pool_put(h, q);
    return rc;
}

}  // extern "C"
