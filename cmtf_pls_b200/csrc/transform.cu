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
    cudaStream_t st = h->stream;
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

        // ---- read-only path for complete data: R raw projections of the UNTOUCHED rows, then the
        //      deflation recurrence on the scores alone.  With x_c = x - mean and no NaN,
        //        t_a = mean_l (x_c - sum_{b<a} t_b w_lb) . w_la = r_a - c_a - sum_{b<a} t_b G_ba,
        //      r_a = mean_l x . w_la,  c_a = mean_l mean_l . w_la,  G_ba = mean_l w_lb . w_la  (passed in).
        bool done = false;
        if (proj_offset != nullptr && proj_gram != nullptr) {
            cudaError_t e = cudaMemcpyAsync(pc, proj_offset, sizeof(double) * R, cudaMemcpyDefault, st);
            if (e == cudaSuccess) e = cudaMemcpyAsync(pg, proj_gram, sizeof(double) * R * R, cudaMemcpyDefault, st);
            if (e == cudaSuccess) e = cudaMemsetAsync(flag, 0, sizeof(int) * 4, st);
            if (e != cudaSuccess) {
                rc = fail(h, "tpls_transform: %s", cudaGetErrorString(e));
                break;
            }
            // component 0 doubles as the NaN census: a counting masked pass fills the per-row observed counts
            for (int l = 0; l < L && !rc; ++l) {
                RowPassArgs r{};
                r.g = grs_cnt[l];
                r.x_in = src[l];
                r.col_w = wk[l];
                r.t_out = S;
                r.tpart = tpart[l];
                r.cpart = cpart[l];
                r.rowcnt = rowcnt[l];
                r.epi = l == 0 ? 0 : (l == L - 1 ? 2 : 1);
                r.div = (double)L;
                rc = row_pass(h, dtypes[l], 2, r);
                if (!rc) {
                    cudaError_t e2 = launch_rows_complete(rowcnt[l], n_new, (double)ps[l], flag, st);
                    if (e2 != cudaSuccess) rc = fail(h, "rows_complete -> %s", cudaGetErrorString(e2));
                    h->stats.kernel_launches++;
                }
            }
            if (rc) break;
            int incomplete = 1;
            cudaError_t e3 = cudaMemcpyAsync(&incomplete, flag, sizeof(int), cudaMemcpyDeviceToHost, st);
            if (e3 == cudaSuccess) e3 = cudaStreamSynchronize(st);
            if (e3 != cudaSuccess) {
                rc = fail(h, "tpls_transform: %s", cudaGetErrorString(e3));
                break;
            }
            if (!incomplete) {
                for (int a = 1; a < R && !rc; ++a)
                    for (int l = 0; l < L && !rc; ++l) {
                        RowPassArgs r{};
                        r.g = grs[l];
                        r.x_in = src[l];
                        r.col_w = wk[l] + (size_t)a * pitch[l];
                        r.t_out = S + (size_t)a * n_new;
                        r.tpart = tpart[l];
                        r.cpart = cpart[l];
                        r.epi = l == 0 ? 0 : (l == L - 1 ? 2 : 1);
                        r.div = (double)L;
                        rc = row_pass(h, dtypes[l], 0, r);
                    }
                if (rc) break;
                cudaError_t e4 = launch_score_recurrence(S, n_new, R, pc, pg, st);
                if (e4 != cudaSuccess) {
                    rc = fail(h, "score_recurrence -> %s", cudaGetErrorString(e4));
                    break;
                }
                h->stats.kernel_launches++;
                done = true;
            }
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

}  // extern "C"
