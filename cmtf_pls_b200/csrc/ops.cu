// Single-operator entry points of the C ABI (tpls_op_*): one streaming pass or one rank-1 step on caller
// buffers, timed with CUDA events on the handle's stream -- for the operator-level parity tests
// (tests/test_gpu_ops.py) and tuning (tools/opbench.py), not used by the fit.
#include "driver_internal.cuh"


static int time_loop(tpls_handle h, int repeats, float* ms_out, const std::function<int()>& body) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    repeats = std::max(1, repeats);
    int rc = 0;
    if (repeats > 1) rc = body();  // warm-up
    CK(cudaEventRecord(e0, h->stream));
    for (int i = 0; i < repeats && !rc; ++i) rc = body();
    CK(cudaEventRecord(e1, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms_out) *ms_out = ms / repeats;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return rc;
}

extern "C" {

static int op_check(tpls_handle h, int dtype, int64_t n, int64_t p) {
    if (!h) return fail(nullptr, "NULL handle");
    if (dtype != TPLS_F32 && dtype != TPLS_F64) return fail(h, "bad dtype");
    const int elem = dtype == TPLS_F32 ? 4 : 8;
    if (n <= 0 || p <= 0 || (p * elem) % 16 != 0) return fail(h, "operator needs p*elem %% 16 == 0 (p=%lld)", (long long)p);
    return 0;
}

int tpls_op_contract(tpls_handle h, const void* x, int dtype, int64_t n, int64_t p, const double* u, int masked,
                     double* z_out, float* ms_out, int repeats) {
    TRY(op_check(h, dtype, n, p));
    CK(cudaSetDevice(h->device));
    const int elem = dtype == TPLS_F32 ? 4 : 8;
    PassGeom g = make_geom(n, (int)p, (int)p, elem, h->sm_count);
    double *zpart = nullptr, *cntpart = nullptr, *cnt = nullptr;
#ifdef TPLS_PROBE
    // TPLS_DBG & 64: the contraction as the fit runs it -- u[row] = Y[row,:] . q from rows of Y staged beside the X tiles
    double *fake_y = nullptr, *fake_q = nullptr;
    const bool y_variant = (tpls::tune_env("TPLS_DBG", 0) & 64) != 0;
    if (y_variant) {
        g = make_geom(n, (int)p, (int)p, elem, h->sm_count, 32);
        CK(cudaMalloc((void**)&fake_y, sizeof(double) * n * 4));
        CK(cudaMalloc((void**)&fake_q, sizeof(double) * 8));
        CK(cudaMemset(fake_y, 0, sizeof(double) * n * 4));
        CK(cudaMemset(fake_q, 0, sizeof(double) * 8));
    }
#endif
    CK(cudaMalloc((void**)&zpart, sizeof(double) * g.grid_x * p));
    CK(cudaMalloc((void**)&cntpart, sizeof(double) * g.grid_x * p));
    CK(cudaMalloc((void**)&cnt, sizeof(double) * p));
    int rc = 0;
    if (masked) {
        ColPassArgs c{};
        c.g = g;
        c.x_in = x;
        c.zpart = zpart;
        c.cntpart = cntpart;
        rc = col_pass(h, dtype, true, PF_COLSTAT, c);
        if (!rc) rc = reduce_cols(h, cntpart, cnt, (int)p, (int)p, g.grid_x, nullptr, nullptr, 0, nullptr, 0);
    }
    if (!rc)
        rc = time_loop(h, repeats, ms_out, [&]() -> int {
            ColPassArgs c{};
            c.g = g;
            c.x_in = x;
            c.row_u = u;
            c.zpart = zpart;
#ifdef TPLS_PROBE
            if (y_variant) {
                c.y = fake_y;
                c.q = fake_q;
                c.pitch_y = 4;
            }
#endif
            TRY(col_pass(h, dtype, masked != 0, PF_CONTRACT, c));
            return reduce_cols(h, zpart, z_out, (int)p, (int)p, g.grid_x, nullptr, nullptr, 0, nullptr, 0);
        });
    // during a fit the observed-count rescaling (missingvals.py:18) is applied by the rank-1 kernel
    if (!rc && masked) {
        cudaError_t e = launch_count_rescale(z_out, cnt, (double)n, (int)p, h->stream);
        if (e != cudaSuccess) rc = fail(h, "count_rescale -> %s", cudaGetErrorString(e));
    }
    cudaStreamSynchronize(h->stream);
    cudaFree(zpart);
    cudaFree(cntpart);
    cudaFree(cnt);
#ifdef TPLS_PROBE
    cudaFree(fake_y);
    cudaFree(fake_q);
#endif
    return rc;
}

int tpls_op_project(tpls_handle h, const void* x, int dtype, int64_t n, int64_t p, const double* w, int masked,
                    double* t_out, float* ms_out, int repeats) {
    TRY(op_check(h, dtype, n, p));
    CK(cudaSetDevice(h->device));
    const int elem = dtype == TPLS_F32 ? 4 : 8;
    PassGeom g = make_row_geom(n, (int)p, (int)p, elem, h->sm_count, false);
    PassGeom g_cnt = make_row_geom(n, (int)p, (int)p, elem, h->sm_count, true);
    double *tpart = nullptr, *cpart = nullptr, *rowcnt = nullptr;
    bool counted = false;
    CK(cudaMalloc((void**)&rowcnt, sizeof(double) * n));
    if (g.n_slabs > 1) {
        CK(cudaMalloc((void**)&tpart, sizeof(double) * n * g.n_slabs));
        CK(cudaMalloc((void**)&cpart, sizeof(double) * n * g.n_slabs));
    }
    int rc = time_loop(h, repeats, ms_out, [&]() -> int {
        RowPassArgs r{};
        r.g = (masked && !counted) ? g_cnt : g;
        r.x_in = x;
        r.col_w = w;
        r.t_out = t_out;
        r.tpart = tpart;
        r.cpart = cpart;
        r.rowcnt = rowcnt;
        r.epi = 0;
        r.div = 1.0;
        const int mode = !masked ? 0 : (counted ? 1 : 2);
        counted = true;
        return row_pass(h, dtype, mode, r);
    });
    cudaStreamSynchronize(h->stream);
    if (tpart) cudaFree(tpart);
    if (cpart) cudaFree(cpart);
    cudaFree(rowcnt);
    return rc;
}

int tpls_op_deflate_contract(tpls_handle h, void* x, int dtype, int64_t n, int64_t p, const double* t,
                             const double* w, const double* u, int masked, double* z_out, double* ss_out,
                             float* ms_out, int repeats) {
    TRY(op_check(h, dtype, n, p));
    CK(cudaSetDevice(h->device));
    const int elem = dtype == TPLS_F32 ? 4 : 8;
    PassGeom g = make_geom(n, (int)p, (int)p, elem, h->sm_count);
    double *zpart = nullptr, *sspart = nullptr;
    CK(cudaMalloc((void**)&zpart, sizeof(double) * g.grid_x * p));
    CK(cudaMalloc((void**)&sspart, sizeof(double) * g.grid_x * g.n_slabs));
    int rc = time_loop(h, repeats, ms_out, [&]() -> int {
        ColPassArgs c{};
        c.g = g;
        c.x_in = x;
        c.x_out = x;
        c.row_a = t;
        c.col_w = w;
        c.row_u = u;
        c.zpart = zpart;
        c.sspart = sspart;
        TRY(col_pass(h, dtype, masked != 0, PF_DEFLATE | PF_WRITE | PF_CONTRACT | PF_SUMSQ, c));
        return reduce_cols(h, zpart, z_out, (int)p, (int)p, g.grid_x, sspart, ss_out, g.grid_x * g.n_slabs, nullptr, 0);
    });
    cudaStreamSynchronize(h->stream);
    cudaFree(zpart);
    cudaFree(sspart);
    return rc;
}

int tpls_op_rank1(tpls_handle h, const double* z, int nmodes, const int* dims, double tol, int flags, double* w_out,
                  double* wkron_out, int* sweeps_out, float* ms_out, int repeats) {
    if (!h) return fail(nullptr, "NULL handle");
    if (nmodes < 1 || nmodes > kMaxZModes) return fail(h, "tpls_op_rank1: nmodes must be 1..%d", kMaxZModes);
    CK(cudaSetDevice(h->device));
    Rank1Args ra{};
    ra.n_tasks = 1;
    ra.tol = tol;
    ra.normalize_on_break = (flags & TPLS_FIT_NORMALIZE_ON_BREAK) ? 1 : 0;
    Rank1Task& k = ra.t[0];
    long long p = 1;
    size_t off = 0;
    for (int m = 0; m < nmodes; ++m) {
        k.dims[m] = dims[m];
        k.w[m] = w_out + off;
        off += dims[m];
        p *= dims[m];
    }
    k.z = z;
    k.p = (int)p;
    k.pitch = (int)p;
    k.nmodes = nmodes;
    k.wkron = wkron_out;
    int nmax = 1, zs_len = 0, mt_len = 0, tab_cols = 0;
    const size_t ws = rank1_workspace_doubles(nmodes, dims, &nmax, &zs_len, &mt_len, &tab_cols);
    k.nmax = nmax;
    k.zs_len = zs_len;
    k.mt_len = mt_len;
    k.tab_cols = tab_cols;
    double* scratch = nullptr;
    int* sweeps = nullptr;
    CK(cudaMalloc((void**)&scratch, sizeof(double) * ws));
    CK(cudaMalloc((void**)&sweeps, sizeof(int) * 4));
    k.scratch = scratch;
    k.sweeps = sweeps;
    long long* stamps = nullptr;
    const bool want_stamps = getenv("TPLS_RANK1_STAMPS") != nullptr;
    if (want_stamps) {
        CK(cudaMalloc((void**)&stamps, sizeof(long long) * 16));
        CK(cudaMemset(stamps, 0, sizeof(long long) * 16));
        k.stamps = stamps;
    }
    size_t smem = ws * sizeof(double);
    k.use_smem = smem <= 200 * 1024;
    if (!k.use_smem) smem = 0;
    int rc = time_loop(h, repeats, ms_out, [&]() -> int {
        CK(launch_rank1(ra, smem, h->stream));
        h->stats.kernel_launches++;
        return 0;
    });
    cudaStreamSynchronize(h->stream);
    if (!rc && sweeps_out) cudaMemcpy(sweeps_out, sweeps, sizeof(int), cudaMemcpyDeviceToHost);
    if (want_stamps) {
        long long hs[16];
        cudaMemcpy(hs, stamps, sizeof hs, cudaMemcpyDeviceToHost);
        fprintf(stderr, "rank1 stamps (cycles): load->gram %lld  eig %lld  init-rest %lld  als %lld  publish %lld\n", hs[1] - hs[0],
                hs[2] - hs[1], hs[3] - hs[2], hs[4] - hs[3], hs[5] - hs[4]);
        fprintf(stderr, "   last run: als rows %lld  als sums %lld  als renorm %lld | squarings %lld  in %lld  eig tail %lld\n", hs[8],
                hs[9], hs[10], hs[12], hs[13], hs[14]);
        cudaFree(stamps);
    }
    cudaFree(scratch);
    cudaFree(sweeps);
    return rc;
}

}  // extern "C"
