// C ABI of libvlgp_b200.so (include/vlgp_b200.h): context, model/trial-set buffers in HBM, entry points, and the
// small elementwise / reduction kernels of the vem bookkeeping (vlgp/core.py:300-305,350-354,366-416).
#include <stdarg.h>

#include <algorithm>
#include <atomic>
#include <map>
#include <thread>

#include "common.cuh"
#include "p2p.cuh"
#include "hostpack.h"
#include "linalg.cuh"

// An M-step begun with vlgp_mstep_begin may still be running on stream_m: calls that read or overwrite what it writes
// (a, b, noise, da, db) or what it reads (mu, v, y) wait for it first.  vlgp_mstep_end still reports its counter.
static int settle_mstep(vlgp_ctx *ctx) {
    if (!ctx->mstep_pending) return VLGP_OK;
    int rc = vlgp_mstep_pump(ctx, 1 << 30);
    if (rc) return rc;
    CK(cudaStreamSynchronize(ctx->stream_m));
    return VLGP_OK;
}
#define SETTLE()                          \
    do {                                  \
        int rs_ = settle_mstep(ctx);      \
        if (rs_) return rs_;              \
    } while (0)

// kernels.cu files
int vlgp_launch_ichol(vlgp_ctx *ctx, PriorFactor &pf, const double *d_omega, const double *d_sigma, double *d_work);
int vlgp_launch_estep_generic(vlgp_ctx *ctx, TrialSet *ts, int mode, int n_iter, double dmu_bound, int method_vb,
                              const int32_t *d_subset = nullptr, int n_subset = 0);
int vlgp_launch_estep_segments(vlgp_ctx *ctx, TrialSet *ts, int n_iter, double dmu_bound, int method_vb, bool *handled,
                               const int32_t *d_subset = nullptr, int n_subset = 0);
int vlgp_launch_estep_long(vlgp_ctx *ctx, TrialSet *ts, int n_iter, double dmu_bound, int method_vb, bool *handled);
int vlgp_launch_mstep(vlgp_ctx *ctx, TrialSet *ts, int n_iter, int use_hessian, double eps, double lr,
                      double da_bound, double db_bound);
int vlgp_mstep_job_setup(vlgp_ctx *ctx, TrialSet *ts, int n_iter, int use_hessian, double eps, double lr,
                         double da_bound, double db_bound);
int vlgp_mstep_job_pump(vlgp_ctx *ctx, int max_iters);
int vlgp_mstep_job_remaining(vlgp_ctx *ctx);
void vlgp_mstep_job_free(vlgp_ctx *ctx);
int vlgp_launch_hstep_prepare(vlgp_ctx *ctx, TrialSet *ts);
int vlgp_launch_hstep_objective(vlgp_ctx *ctx, TrialSet *ts, const HEvalBatch &eb, double *ll, double *dll, int *info);
void vlgp_comm_destroy(vlgp_ctx *ctx);

static std::string g_create_error;

int vlgp_fail(vlgp_ctx *ctx, int code, const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (ctx) ctx->err = buf;
    else g_create_error = buf;
    return code;
}

namespace {

// rows != null: the map is applied to the n listed bins, one application per LIST ENTRY (a bin listed twice is mapped
// twice -- by different threads, so a list must not repeat a bin; the host's lists of shared bins never do).
__global__ void latent_affine_kernel(int64_t nbin, int L, double *mu, const double *__restrict__ shiftM, int has_shift,
                                     int has_M, const int64_t *__restrict__ rows) {
    // shiftM: [0,L) shift, [L, L+L*L) M row-major
    int64_t bin = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (bin >= nbin) return;
    if (rows) bin = rows[bin];
    double x[VLGP_MAX_L], o[VLGP_MAX_L];
    for (int l = 0; l < L; ++l) x[l] = mu[bin * L + l] - (has_shift ? shiftM[l] : 0.0);
    if (has_M) {
        for (int k = 0; k < L; ++k) {
            double s = 0.0;
            for (int l = 0; l < L; ++l) s = fma(x[l], shiftM[L + l * L + k], s);
            o[k] = s;
        }
        for (int k = 0; k < L; ++k) mu[bin * L + k] = o[k];
    } else {
        for (int l = 0; l < L; ++l) mu[bin * L + l] = x[l];
    }
}

// dst row <- src row for n (src, dst) pairs of bins, in each of the listed per-bin arrays (L doubles per bin).  All reads
// of a pair's source happen in the same thread as its write; the host never lists a bin both as a source and as a
// destination in one call (tail rows of one segment -> head rows of the next), so the pairs are independent.
__global__ void copy_rows_kernel(int64_t n, int L, const int64_t *__restrict__ src, const int64_t *__restrict__ dst,
                                 double *a0, double *a1, double *a2, double *a3) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * L) return;
    const int64_t pair = i / L;
    const int l = (int)(i - pair * L);
    const int64_t s = src[pair] * L + l, d = dst[pair] * L + l;
    if (a0) a0[d] = a0[s];
    if (a1) a1[d] = a1[s];
    if (a2) a2[d] = a2[s];
    if (a3) a3[d] = a3[s];
}

// part[b][0..2L+2): per-latent sum mu, per-latent sum mu^2, sum dmu^2 total, (unused)
__global__ void __launch_bounds__(256) moments_kernel(int64_t nbin, int L, const double *__restrict__ mu,
                                                      const double *__restrict__ dmu, double *part) {
    __shared__ double red[32];
    double sm[VLGP_MAX_L], sq[VLGP_MAX_L];
    for (int l = 0; l < L; ++l) sm[l] = sq[l] = 0.0;
    double dd = 0.0;
    for (int64_t bin = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; bin < nbin; bin += (int64_t)gridDim.x * blockDim.x)
        for (int l = 0; l < L; ++l) {
            const double m = mu[bin * L + l], d = dmu[bin * L + l];
            sm[l] += m;
            sq[l] = fma(m, m, sq[l]);
            dd = fma(d, d, dd);
        }
    const int K = 2 * L + 1;
    for (int l = 0; l < L; ++l) {
        const double a = block_sum(sm[l], red);
        const double b = block_sum(sq[l], red);
        if (threadIdx.x == 0) {
            part[(size_t)blockIdx.x * K + l] = a;
            part[(size_t)blockIdx.x * K + L + l] = b;
        }
    }
    dd = block_sum(dd, red);
    if (threadIdx.x == 0) part[(size_t)blockIdx.x * K + 2 * L] = dd;
}

__global__ void reduce_parts_kernel3(const double *__restrict__ part, int G, int K, double *__restrict__ out) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= K) return;
    double s = 0.0;
    for (int g = 0; g < G; ++g) s += part[(size_t)g * K + k];
    out[k] = s;
}

// ---- peak probes ----------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) dfma_peak_kernel(int iters, double *out) {
    double a[16];
    const double x = 1.0 + 1e-9 * threadIdx.x, y = 1e-9;
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) a[i] = fma(a[i], x, y);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i];
    if (s == 123.456) out[0] = s;
}

__global__ void __launch_bounds__(256) dmma_peak_kernel(int iters, double *out) {
    double c[8][2];
    const double a = 1.0 + 1e-9 * threadIdx.x, b = 1e-9;
#pragma unroll
    for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1])
                         : "d"(a), "d"(b));
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    if (s == 123.456) out[0] = s;
}

__global__ void copy_kernel(const double4 *__restrict__ src, double4 *__restrict__ dst, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        dst[i] = src[i];
}

__global__ void fill_kernel(double4 *dst, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        dst[i] = make_double4(0.0, 0.0, 0.0, 0.0);
}

void free_set(TrialSet &ts, cudaStream_t stream = nullptr) {
    auto F = [&](auto *&p) {
        if (p) {
            if (stream) cudaFreeAsync((void *)p, stream);
            else cudaFree((void *)p);
        }
        p = nullptr;
    };
    F(ts.d_len); F(ts.d_start); F(ts.d_fidx); F(ts.d_Gptr); F(ts.d_ncolptr); F(ts.d_y);
    F(ts.d_mu); F(ts.d_v); F(ts.d_w); F(ts.d_dmu); F(ts.d_ra); F(ts.d_u); F(ts.d_minv);
    F(ts.d_k3_tab); F(ts.d_k3_buf); F(ts.d_k3_bad);
    F(ts.d_M); F(ts.d_K); F(ts.d_hpart); F(ts.d_hout); F(ts.d_mompart); F(ts.d_x); F(ts.d_xb);
    for (auto &pf : ts.factors) {
        F(pf.d_G); F(pf.d_ncol); F(pf.d_piv);
    }
    ts = TrialSet();
}

}   // namespace

extern "C" {

int vlgp_create(int device, vlgp_ctx **out) {
    vlgp_ctx *ctx = nullptr;
    if (!out) return vlgp_fail(nullptr, VLGP_ERR_ARG, "vlgp_create: out is NULL");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return vlgp_fail(nullptr, VLGP_ERR_CUDA, "vlgp_create: no CUDA device (%s)", cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return vlgp_fail(nullptr, VLGP_ERR_ARG, "vlgp_create: bad device %d", device);
    e = cudaSetDevice(device);
    if (e != cudaSuccess) return vlgp_fail(nullptr, VLGP_ERR_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
    ctx = new vlgp_ctx();
    ctx->device = device;
    cudaGetDeviceProperties(&ctx->prop, device);
    if (ctx->prop.major < 10) {
        int rc = vlgp_fail(nullptr, VLGP_ERR_UNSUPPORTED, "vlgp_create: device %s is sm_%d%d; this library is sm_100a only",
                           ctx->prop.name, ctx->prop.major, ctx->prop.minor);
        delete ctx;
        return rc;
    }
    // The overlapped M-step is the background job: its stream has the lowest priority so that the short kernels of the
    // host-driven H-step rounds on the main stream are scheduled ahead of its queued CTAs.
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    bool ok = cudaStreamCreateWithPriority(&ctx->stream, cudaStreamNonBlocking, prio_hi) == cudaSuccess;
    ok = ok && cudaStreamCreateWithPriority(&ctx->stream2, cudaStreamNonBlocking, prio_hi) == cudaSuccess;
    ok = ok && cudaStreamCreateWithPriority(&ctx->stream_m, cudaStreamNonBlocking, prio_lo) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&ctx->ev_m_start, cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&ctx->ev_m_done, cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaEventCreate(&ctx->ev0) == cudaSuccess && cudaEventCreate(&ctx->ev1) == cudaSuccess;
    ok = ok && cudaEventCreate(&ctx->pev0) == cudaSuccess && cudaEventCreate(&ctx->pev1) == cudaSuccess;
    ok = ok && cudaMalloc(&ctx->d_flags, 16 * sizeof(int)) == cudaSuccess;
    ok = ok && cudaMallocHost(&ctx->h_flags, 16 * sizeof(int)) == cudaSuccess;
    ok = ok && cudaMallocHost(&ctx->h_pin, 4096) == cudaSuccess;
    ok = ok && cudaMalloc(&ctx->d_small, 4096) == cudaSuccess;
    ok = ok && cudaMemset(ctx->d_flags, 0, 16 * sizeof(int)) == cudaSuccess;
    {   // trial-set buffers come from the stream-ordered pool: keep freed blocks cached instead of returning them to
        // the OS at every synchronisation, so creating / freeing a trial set costs microseconds, not milliseconds
        cudaMemPool_t pool = nullptr;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess && pool) {
            uint64_t thr = UINT64_MAX;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
        }
    }
    if (!ok) {
        int rc = vlgp_fail(nullptr, VLGP_ERR_CUDA, "vlgp_create: %s", cudaGetErrorString(cudaGetLastError()));
        delete ctx;
        return rc;
    }
    *out = ctx;
    return VLGP_OK;
}

int vlgp_destroy(vlgp_ctx *ctx) {
    if (!ctx) return VLGP_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    cudaStreamSynchronize(ctx->stream_m);
    vlgp_mstep_job_free(ctx);
    vlgp_comm_destroy(ctx);
    for (auto &ts : ctx->sets)
        if (ts.used) free_set(ts, ctx->stream);
    auto F = [](auto *&p) {
        if (p) cudaFree((void *)p);
        p = nullptr;
    };
    F(ctx->d_poisson); F(ctx->d_a); F(ctx->d_b); F(ctx->d_noise); F(ctx->d_da); F(ctx->d_db);
    F(ctx->d_bpart); F(ctx->d_bstat);
    F(ctx->d_mpart); F(ctx->d_mstat); F(ctx->d_ymom); F(ctx->d_ppack); F(ctx->d_gshared); F(ctx->d_flags); F(ctx->d_smslots); F(ctx->d_small); F(ctx->d_flush);
    if (ctx->stream_copy) { cudaStreamSynchronize(ctx->stream_copy); cudaStreamDestroy(ctx->stream_copy); }
    if (ctx->h_prefetch) cudaFreeHost(ctx->h_prefetch);
    if (ctx->ev_prefetch_go) cudaEventDestroy(ctx->ev_prefetch_go);
    if (ctx->ev_prefetch_done) cudaEventDestroy(ctx->ev_prefetch_done);
    if (ctx->h_flags) cudaFreeHost(ctx->h_flags);
    if (ctx->h_pin) cudaFreeHost(ctx->h_pin);
    for (int i = 0; i < 2; ++i) {
        if (ctx->h_stage[i]) cudaFreeHost(ctx->h_stage[i]);
        if (ctx->stage_ev[i]) cudaEventDestroy(ctx->stage_ev[i]);
    }
    cudaEventDestroy(ctx->ev0); cudaEventDestroy(ctx->ev1); cudaEventDestroy(ctx->pev0); cudaEventDestroy(ctx->pev1);
    cudaEventDestroy(ctx->ev_fork); cudaEventDestroy(ctx->ev_join);
    cudaEventDestroy(ctx->ev_m_start); cudaEventDestroy(ctx->ev_m_done);
    cudaStreamDestroy(ctx->stream_m);
    cudaStreamDestroy(ctx->stream2);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
    return VLGP_OK;
}

const char *vlgp_last_error(const vlgp_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int vlgp_device_info(vlgp_ctx *ctx, int *sm_count, int *cc_major, int *cc_minor, uint64_t *total_mem, char name[128]) {
    if (!ctx) return VLGP_ERR_ARG;
    if (sm_count) *sm_count = ctx->prop.multiProcessorCount;
    if (cc_major) *cc_major = ctx->prop.major;
    if (cc_minor) *cc_minor = ctx->prop.minor;
    if (total_mem) *total_mem = ctx->prop.totalGlobalMem;
    if (name) {
        strncpy(name, ctx->prop.name, 127);
        name[127] = 0;
    }
    return VLGP_OK;
}

int vlgp_sync(vlgp_ctx *ctx) {
    if (!ctx) return VLGP_ERR_ARG;
    CK(cudaStreamSynchronize(ctx->stream));
    return VLGP_OK;
}

int vlgp_timer_start(vlgp_ctx *ctx) {
    if (!ctx) return VLGP_ERR_ARG;
    CK(cudaEventRecord(ctx->ev0, ctx->stream));
    return VLGP_OK;
}

int vlgp_timer_stop(vlgp_ctx *ctx, float *ms) {
    if (!ctx || !ms) return VLGP_ERR_ARG;
    CK(cudaEventRecord(ctx->ev1, ctx->stream));
    CK(cudaEventSynchronize(ctx->ev1));
    CK(cudaEventElapsedTime(ms, ctx->ev0, ctx->ev1));
    return VLGP_OK;
}

int vlgp_counters(vlgp_ctx *ctx, int64_t c[4]) {
    if (!ctx || !c) return VLGP_ERR_ARG;
    for (int i = 0; i < 4; ++i) c[i] = ctx->counters[i];
    return VLGP_OK;
}

// ---- model ------------------------------------------------------------------------------------------------------------
int vlgp_set_model(vlgp_ctx *ctx, int N, int L, int rank, const uint8_t *poisson_mask, double gp_noise, double dt) {
    if (!ctx) return VLGP_ERR_ARG;
    REQUIRE(N >= 1 && L >= 1 && L <= 12, "set_model: need 1 <= n_latents <= 12 and n_neurons >= 1 (got %d, %d)", L, N);
    REQUIRE(rank >= 1 && rank <= VLGP_MAX_RANK, "set_model: rank %d outside [1, %d]", rank, VLGP_MAX_RANK);
    REQUIRE(poisson_mask != nullptr, "set_model: poisson_mask is NULL");
    CK(cudaSetDevice(ctx->device));
    SETTLE();
    CK(cudaStreamSynchronize(ctx->stream));
    for (auto &ts : ctx->sets)
        if (ts.used) free_set(ts, ctx->stream);
    auto F = [](auto *&p) {
        if (p) cudaFree((void *)p);
        p = nullptr;
    };
    F(ctx->d_poisson); F(ctx->d_a); F(ctx->d_b); F(ctx->d_noise); F(ctx->d_da); F(ctx->d_db); F(ctx->d_mstat);
    F(ctx->d_mpart); F(ctx->d_ymom); F(ctx->d_ppack); F(ctx->d_bpart); F(ctx->d_bstat);
    ctx->mpart_grid = 0;
    ctx->bpart_len = 0;
    ctx->xdim = 1;
    ctx->N = N; ctx->L = L; ctx->rank = rank; ctx->gp_noise = gp_noise; ctx->dt = dt;
    CK(cudaMalloc(&ctx->d_poisson, N));
    CK(cudaMalloc(&ctx->d_a, (size_t)L * N * sizeof(double)));
    CK(cudaMalloc(&ctx->d_da, (size_t)L * N * sizeof(double)));
    CK(cudaMalloc(&ctx->d_b, (size_t)N * sizeof(double)));
    CK(cudaMalloc(&ctx->d_db, (size_t)N * sizeof(double)));
    CK(cudaMalloc(&ctx->d_noise, (size_t)N * sizeof(double)));
    const int nstat = L + L * (L + 1) / 2 + 4;
    CK(cudaMalloc(&ctx->d_mstat, ((size_t)nstat * N + 8) * sizeof(double)));
    CK(cudaMemcpy(ctx->d_poisson, poisson_mask, N, cudaMemcpyHostToDevice));
    CK(cudaMemset(ctx->d_a, 0, (size_t)L * N * sizeof(double)));
    CK(cudaMemset(ctx->d_da, 0, (size_t)L * N * sizeof(double)));
    CK(cudaMemset(ctx->d_b, 0, (size_t)N * sizeof(double)));
    CK(cudaMemset(ctx->d_db, 0, (size_t)N * sizeof(double)));
    std::vector<double> ones(N, 1.0);
    CK(cudaMemcpy(ctx->d_noise, ones.data(), (size_t)N * sizeof(double), cudaMemcpyHostToDevice));
    ctx->any_gauss = false;
    for (int n = 0; n < N; ++n)
        if (!poisson_mask[n]) ctx->any_gauss = true;
    ctx->h_sigma.assign(L, 1.0);
    ctx->h_omega.assign(L, 1e-3);
    return VLGP_OK;
}

int vlgp_set_regressors(vlgp_ctx *ctx, int xdim) {
    if (!ctx) return VLGP_ERR_ARG;
    REQUIRE(ctx->N > 0, "set_regressors before set_model");
    REQUIRE(xdim >= 1 && xdim <= VLGP_MAX_XDIM, "set_regressors: xdim %d outside [1, %d]", xdim, VLGP_MAX_XDIM);
    if (xdim == ctx->xdim) return VLGP_OK;
    CK(cudaSetDevice(ctx->device));
    SETTLE();
    CK(cudaStreamSynchronize(ctx->stream));
    for (auto &ts : ctx->sets)
        REQUIRE(!ts.used, "set_regressors: free the trial sets of the previous model first");
    if (ctx->d_b) CK(cudaFree(ctx->d_b));
    if (ctx->d_db) CK(cudaFree(ctx->d_db));
    ctx->d_b = ctx->d_db = nullptr;
    const size_t bytes = (size_t)xdim * ctx->N * sizeof(double);
    CK(cudaMalloc(&ctx->d_b, bytes));
    CK(cudaMalloc(&ctx->d_db, bytes));
    CK(cudaMemset(ctx->d_b, 0, bytes));
    CK(cudaMemset(ctx->d_db, 0, bytes));
    ctx->xdim = xdim;
    return VLGP_OK;
}

int vlgp_trials_set_x(vlgp_ctx *ctx, int set_id, const double *x) {
    TrialSet *ts = get_set(ctx, set_id);
    REQUIRE(ts && x, "trials_set_x: bad arguments");
    CK(cudaSetDevice(ctx->device));
    SETTLE();
    const size_t n = (size_t)ts->nbin * ctx->xdim * ctx->N;
    if (!ts->d_x) CK(vlgp_dalloc(ctx, &ts->d_x, n * sizeof(double)));
    if (!ts->d_xb) CK(vlgp_dalloc(ctx, &ts->d_xb, (size_t)ts->nbin * ctx->N * sizeof(double)));
    CK(cudaMemcpyAsync(ts->d_x, x, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return VLGP_OK;
}

int vlgp_set_params(vlgp_ctx *ctx, const double *a, const double *b, const double *noise, const double *sigma,
                    const double *omega) {
    if (!ctx) return VLGP_ERR_ARG;
    REQUIRE(ctx->N > 0, "set_params before set_model");
    const size_t N = ctx->N, L = ctx->L;
    CK(cudaSetDevice(ctx->device));
    if (a || b || noise) SETTLE();
    if (a) CK(cudaMemcpyAsync(ctx->d_a, a, L * N * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    if (b) CK(cudaMemcpyAsync(ctx->d_b, b, ctx->xdim * N * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    if (noise) CK(cudaMemcpyAsync(ctx->d_noise, noise, N * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    if (sigma) ctx->h_sigma.assign(sigma, sigma + L);
    if (omega) ctx->h_omega.assign(omega, omega + L);
    CK(cudaStreamSynchronize(ctx->stream));     // host buffers are borrowed only for the duration of the call
    return VLGP_OK;
}

int vlgp_get_params(vlgp_ctx *ctx, double *a, double *b, double *noise, double *da, double *db, double *sigma,
                    double *omega) {
    if (!ctx) return VLGP_ERR_ARG;
    REQUIRE(ctx->N > 0, "get_params before set_model");
    const size_t N = ctx->N, L = ctx->L;
    CK(cudaSetDevice(ctx->device));
    SETTLE();
    if (a) CK(cudaMemcpyAsync(a, ctx->d_a, L * N * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if (da) CK(cudaMemcpyAsync(da, ctx->d_da, L * N * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if (b) CK(cudaMemcpyAsync(b, ctx->d_b, ctx->xdim * N * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if (db) CK(cudaMemcpyAsync(db, ctx->d_db, ctx->xdim * N * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if (noise) CK(cudaMemcpyAsync(noise, ctx->d_noise, N * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (sigma) std::copy(ctx->h_sigma.begin(), ctx->h_sigma.end(), sigma);
    if (omega) std::copy(ctx->h_omega.begin(), ctx->h_omega.end(), omega);
    return VLGP_OK;
}

// ---- trial sets -------------------------------------------------------------------------------------------------------
int vlgp_trials_create(vlgp_ctx *ctx, int n_trials, const int32_t *lengths, int *set_id) {
    if (!ctx) return VLGP_ERR_ARG;
    REQUIRE(ctx->N > 0, "trials_create before set_model");
    REQUIRE(n_trials >= 1 && lengths && set_id, "trials_create: bad arguments");
    CK(cudaSetDevice(ctx->device));
    int id = -1;
    for (size_t i = 0; i < ctx->sets.size(); ++i)
        if (!ctx->sets[i].used) { id = (int)i; break; }
    if (id < 0) {
        ctx->sets.emplace_back();
        id = (int)ctx->sets.size() - 1;
    }
    REQUIRE(id < 0x1000, "trials_create: too many live trial sets");
    TrialSet &ts = ctx->sets[id];
    ts = TrialSet();
    ts.used = true;
    ts.gen = (++ctx->set_gen) & 0x7ffff;
    ts.n_trials = n_trials;
    ts.h_len.assign(lengths, lengths + n_trials);
    ts.h_start.resize(n_trials);
    std::map<int, int> uniq;
    int64_t off = 0;
    ts.max_len = 0;
    ts.min_len = 1 << 30;
    for (int i = 0; i < n_trials; ++i) {
        if (lengths[i] < 1) {
            ts.used = false;
            return vlgp_fail(ctx, VLGP_ERR_ARG, "trials_create: trial %d has length %d", i, lengths[i]);
        }
        ts.h_start[i] = off;
        off += lengths[i];
        ts.max_len = std::max(ts.max_len, (int)lengths[i]);
        ts.min_len = std::min(ts.min_len, (int)lengths[i]);
        uniq.emplace(lengths[i], 0);
    }
    ts.nbin = off;
    int k = 0;
    for (auto &kv : uniq) kv.second = k++;
    ts.factors.resize(uniq.size());
    for (auto &kv : uniq) ts.factors[kv.second].length = kv.first;
    ts.h_fidx.resize(n_trials);
    for (int i = 0; i < n_trials; ++i) ts.h_fidx[i] = uniq[lengths[i]];

    const size_t L = ctx->L, N = ctx->N, R = ctx->rank;
    CK(vlgp_dalloc(ctx, &ts.d_len, n_trials * sizeof(int)));
    CK(vlgp_dalloc(ctx, &ts.d_start, n_trials * sizeof(int64_t)));
    CK(vlgp_dalloc(ctx, &ts.d_fidx, n_trials * sizeof(int)));
    CK(cudaMemcpyAsync(ts.d_len, ts.h_len.data(), n_trials * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ts.d_start, ts.h_start.data(), n_trials * sizeof(int64_t), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ts.d_fidx, ts.h_fidx.data(), n_trials * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    std::vector<double *> gp(ts.factors.size());
    std::vector<int *> np_(ts.factors.size());
    for (size_t f = 0; f < ts.factors.size(); ++f) {
        PriorFactor &pf = ts.factors[f];
        CK(vlgp_dalloc(ctx, &pf.d_G, L * pf.length * R * sizeof(double)));
        CK(cudaMemsetAsync(pf.d_G, 0, L * pf.length * R * sizeof(double), ctx->stream));
        CK(vlgp_dalloc(ctx, &pf.d_ncol, L * sizeof(int)));
        CK(cudaMemsetAsync(pf.d_ncol, 0, L * sizeof(int), ctx->stream));
        CK(vlgp_dalloc(ctx, &pf.d_piv, L * R * sizeof(int)));
        CK(cudaMemsetAsync(pf.d_piv, 0xff, L * R * sizeof(int), ctx->stream));
        pf.h_ncol.assign(L, 0);
        gp[f] = pf.d_G;
        np_[f] = pf.d_ncol;
    }
    CK(vlgp_dalloc(ctx, &ts.d_Gptr, gp.size() * sizeof(double *)));
    CK(vlgp_dalloc(ctx, &ts.d_ncolptr, np_.size() * sizeof(int *)));
    CK(cudaMemcpyAsync(ts.d_Gptr, gp.data(), gp.size() * sizeof(double *), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ts.d_ncolptr, np_.data(), np_.size() * sizeof(int *), cudaMemcpyHostToDevice, ctx->stream));
    const size_t nb = (size_t)ts.nbin;
    CK(vlgp_dalloc(ctx, &ts.d_mu, nb * L * sizeof(double)));
    CK(vlgp_dalloc(ctx, &ts.d_v, nb * L * sizeof(double)));
    CK(vlgp_dalloc(ctx, &ts.d_w, nb * L * sizeof(double)));
    CK(vlgp_dalloc(ctx, &ts.d_dmu, nb * L * sizeof(double)));
    CK(vlgp_dalloc(ctx, &ts.d_ra, nb * L * sizeof(double)));
    CK(vlgp_dalloc(ctx, &ts.d_u, nb * sizeof(double)));
    CK(cudaMemsetAsync(ts.d_mu, 0, nb * L * sizeof(double), ctx->stream));
    CK(cudaMemsetAsync(ts.d_v, 0, nb * L * sizeof(double), ctx->stream));
    CK(cudaMemsetAsync(ts.d_w, 0, nb * L * sizeof(double), ctx->stream));
    CK(cudaMemsetAsync(ts.d_dmu, 0, nb * L * sizeof(double), ctx->stream));
    (void)N;
    CK(cudaStreamSynchronize(ctx->stream));     // the host tables above are borrowed by the async copies
    *set_id = id | (ts.gen << 12);
    return VLGP_OK;
}

int vlgp_trials_free(vlgp_ctx *ctx, int set_id) {
    TrialSet *ts = get_set(ctx, set_id);
    if (!ts) return vlgp_fail(ctx, VLGP_ERR_ARG, "trials_free: bad set %d", set_id);
    CK(cudaSetDevice(ctx->device));
    SETTLE();
    if (ctx->prefetch_set == set_id) {            // a prefetch of this set's state may still be in flight
        if (ctx->stream_copy) CK(cudaStreamSynchronize(ctx->stream_copy));
        ctx->prefetch_set = -1;
    }
    free_set(*ts, ctx->stream);
    return VLGP_OK;
}

int vlgp_trials_set_y(vlgp_ctx *ctx, int set_id, const void *y, int ydtype) {
    TrialSet *ts = get_set(ctx, set_id);
    REQUIRE(ts && y, "trials_set_y: bad set %d or NULL y", set_id);
    REQUIRE(ydtype == VLGP_Y_F64 || ydtype == VLGP_Y_U8, "trials_set_y: bad dtype %d", ydtype);
    CK(cudaSetDevice(ctx->device));
    SETTLE();
    const size_t bytes = (size_t)ts->nbin * ctx->N * (ydtype == VLGP_Y_U8 ? 1 : sizeof(double));
    if (ts->d_y && ts->ydtype != ydtype) {
        CK(vlgp_dfree(ctx, ts->d_y));
        ts->d_y = nullptr;
    }
    if (!ts->d_y) CK(vlgp_dalloc(ctx, &ts->d_y, bytes));
    ts->ydtype = ydtype;
    CK(cudaMemcpyAsync(ts->d_y, y, bytes, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return VLGP_OK;
}

// Gather per-trial observation blocks (each rows[i] x N, C-contiguous) into the set's device buffer through two pinned
// staging buffers: host threads convert/copy block k+1 while block k is in flight over PCIe.  float64 sources holding
// only integer counts in [0, 255] are stored as uint8 (8x less HBM traffic in every pass over y); anything else stays
// float64.  stored_dtype returns what was chosen.
int vlgp_trials_set_y_parts(vlgp_ctx *ctx, int set_id, int n_parts, const void *const *parts, const int64_t *rows,
                            int src_dtype, int *stored_dtype) {
    TrialSet *ts = get_set(ctx, set_id);
    REQUIRE(ts && parts && rows && n_parts >= 1, "trials_set_y_parts: bad arguments");
    REQUIRE(src_dtype == VLGP_Y_F64 || src_dtype == VLGP_Y_U8, "trials_set_y_parts: bad dtype %d", src_dtype);
    CK(cudaSetDevice(ctx->device));
    SETTLE();
    const size_t N = ctx->N;
    int64_t total = 0;
    for (int i = 0; i < n_parts; ++i) total += rows[i];
    REQUIRE(total == ts->nbin, "trials_set_y_parts: parts hold %lld rows, the set has %lld bins", (long long)total,
            (long long)ts->nbin);
    static const size_t STAGE = [] {              // bytes per pinned staging chunk (the buffers hold 8 MiB)
        const char *e = getenv("VLGP_YSTAGE_KB");
        size_t kb = e ? (size_t)atol(e) : 0;
        return (kb >= 64 && kb <= 8192) ? kb << 10 : (size_t)8 << 20;
    }();
    if (!ctx->h_stage[0]) {
        CK(cudaMallocHost(&ctx->h_stage[0], (size_t)8 << 20));
        CK(cudaMallocHost(&ctx->h_stage[1], (size_t)8 << 20));
        CK(cudaEventCreateWithFlags(&ctx->stage_ev[0], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&ctx->stage_ev[1], cudaEventDisableTiming));
    }
    // the float64 -> uint8 conversion streams 8 bytes per entry at ~6 GB/s per thread (hostpack.cpp): up to 16 threads
    const unsigned hw = std::thread::hardware_concurrency();
    static const int thr_env = getenv("VLGP_HOST_THREADS") ? atoi(getenv("VLGP_HOST_THREADS")) : 0;
    const int nthreads = thr_env > 0 ? thr_env : (int)std::min<unsigned>(hw ? hw : 4, src_dtype == VLGP_Y_F64 ? 16 : 8);

    // Flatten the parts into one virtual element range [0, total * N) so that chunks need not align with parts.
    std::vector<int64_t> part_off(n_parts + 1, 0);
    for (int i = 0; i < n_parts; ++i) part_off[i + 1] = part_off[i] + rows[i] * (int64_t)N;
    const int64_t nelem = part_off[n_parts];

    for (int attempt = 0; attempt < 2; ++attempt) {
        // attempt 0: store uint8 (convert + verify when the source is float64); attempt 1: store float64 as is
        const int dst_dtype = (attempt == 0) ? VLGP_Y_U8 : VLGP_Y_F64;
        if (attempt == 1 && src_dtype == VLGP_Y_U8) break;
        const size_t esz = dst_dtype == VLGP_Y_U8 ? 1 : sizeof(double);
        if (ts->d_y && ts->ydtype != dst_dtype) {
            CK(vlgp_dfree(ctx, ts->d_y));
            ts->d_y = nullptr;
        }
        if (!ts->d_y) CK(vlgp_dalloc(ctx, &ts->d_y, (size_t)nelem * esz));
        ts->ydtype = dst_dtype;
        const int64_t chunk = (int64_t)(STAGE / esz);
        std::atomic<bool> exact(true);
        int buf = 0;
        int ip = 0;                                  // part containing the chunk start
        for (int64_t e0 = 0; e0 < nelem && exact.load(); e0 += chunk, buf ^= 1) {
            const int64_t e1 = std::min(nelem, e0 + chunk);
            CK(cudaEventSynchronize(ctx->stage_ev[buf]));        // previous copy out of this buffer has finished
            while (part_off[ip + 1] <= e0) ++ip;
            // split [e0, e1) among the threads
            auto work = [&](int t) {
                const int64_t per = (e1 - e0 + nthreads - 1) / nthreads;
                int64_t a = e0 + t * per, b = std::min(e1, a + per);
                if (a >= b) return;
                int p = ip;
                while (part_off[p + 1] <= a) ++p;
                while (a < b) {
                    const int64_t stop = std::min(b, part_off[p + 1]);
                    const int64_t loc = a - part_off[p];
                    const int64_t cnt = stop - a;
                    unsigned char *dst = (unsigned char *)ctx->h_stage[buf] + (size_t)(a - e0) * esz;
                    if (dst_dtype == VLGP_Y_F64) {
                        memcpy(dst, (const double *)parts[p] + loc, (size_t)cnt * sizeof(double));
                    } else if (src_dtype == VLGP_Y_U8) {
                        memcpy(dst, (const unsigned char *)parts[p] + loc, (size_t)cnt);
                    } else {
                        // exact only for integer counts in [0, 255] (hostpack.cpp: AVX2 body chosen at run time)
                        if (!vlgp_host_f64_to_u8((const double *)parts[p] + loc, dst, cnt)) exact.store(false);
                    }
                    a = stop;
                    ++p;
                }
            };
            vlgp_host_parallel_for(nthreads, work);          // persistent pool (hostpack.cpp)
            if (!exact.load()) break;
            CK(cudaMemcpyAsync((unsigned char *)ts->d_y + (size_t)e0 * esz, ctx->h_stage[buf], (size_t)(e1 - e0) * esz,
                               cudaMemcpyHostToDevice, ctx->stream));
            CK(cudaEventRecord(ctx->stage_ev[buf], ctx->stream));
        }
        // (no synchronisation: the sources were consumed by the host threads above; see pipeline_copy)
        if (exact.load()) {
            if (stored_dtype) *stored_dtype = dst_dtype;
            return VLGP_OK;
        }
    }
    return vlgp_fail(ctx, VLGP_ERR_ARG, "trials_set_y_parts: uint8 source could not be stored");
}

// Gather (to_device) or scatter (from device) a device array of nelem elements of esz bytes from/to host blocks whose
// element offsets are part_off[0..n_parts], through the two pinned staging buffers, with host threads doing the
// block copies while the previous chunk is in flight over PCIe.
static int pipeline_copy(vlgp_ctx *ctx, void *dev, size_t esz, int n_parts, void *const *parts,
                         const std::vector<int64_t> &part_off, bool to_device) {
    const size_t STAGE_MAX = (size_t)8 << 20;
    static const size_t STAGE = [] {
        const char *e = getenv("VLGP_STAGE_KB");
        size_t kb = e ? (size_t)atol(e) : 0;
        return (kb >= 64 && kb <= 8192) ? kb << 10 : (size_t)2 << 20;      // 2 MiB chunks: measured on B200, vem() with host
                                                                            // arrays 34.8 -> 29.9 ms against 8 MiB (gather / scatter
                                                                            // of a chunk overlaps the transfer of its neighbour)
    }();
    if (!ctx->h_stage[0]) {
        CK(cudaMallocHost(&ctx->h_stage[0], STAGE_MAX));
        CK(cudaMallocHost(&ctx->h_stage[1], STAGE_MAX));
        CK(cudaEventCreateWithFlags(&ctx->stage_ev[0], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&ctx->stage_ev[1], cudaEventDisableTiming));
    }
    const int64_t nelem = part_off[n_parts];
    const int64_t chunk = (int64_t)(STAGE / esz);
    const unsigned hw = std::thread::hardware_concurrency();
    static const int thr_env = getenv("VLGP_HOST_THREADS") ? atoi(getenv("VLGP_HOST_THREADS")) : 0;
    int nthreads = thr_env > 0 ? thr_env : (int)std::min<unsigned>(hw ? hw : 4, 8);
    if (!to_device) {
        // Overlapping destination blocks (segments of a trial whose length is not a multiple of the window are
        // overlapping views): scatter with one thread so that the blocks are written in order and the last one wins,
        // like the reference's sequential loop over segments.
        for (int i = 0; i + 1 < n_parts && nthreads > 1; ++i) {
            const unsigned char *a0 = (const unsigned char *)parts[i];
            const unsigned char *a1 = a0 + (size_t)(part_off[i + 1] - part_off[i]) * esz;
            const unsigned char *b0 = (const unsigned char *)parts[i + 1];
            const unsigned char *b1 = b0 + (size_t)(part_off[i + 2] - part_off[i + 1]) * esz;
            if (a0 < b1 && b0 < a1) nthreads = 1;
        }
    }
    auto copy_chunk = [&](int buf, int64_t e0, int64_t e1) {
        int ip = 0;
        {   // binary search of the part containing e0
            int lo = 0, hi = n_parts - 1;
            while (lo < hi) {
                const int mid = (lo + hi + 1) / 2;
                if (part_off[mid] <= e0) lo = mid; else hi = mid - 1;
            }
            ip = lo;
        }
        auto work = [&](int t) {
            const int64_t per = (e1 - e0 + nthreads - 1) / nthreads;
            int64_t a = e0 + t * per;
            const int64_t b = std::min(e1, a + per);
            if (a >= b) return;
            int p = ip;
            while (part_off[p + 1] <= a) ++p;
            while (a < b) {
                const int64_t stop = std::min(b, part_off[p + 1]);
                unsigned char *st = (unsigned char *)ctx->h_stage[buf] + (size_t)(a - e0) * esz;
                unsigned char *hp = (unsigned char *)parts[p] + (size_t)(a - part_off[p]) * esz;
                if (to_device) memcpy(st, hp, (size_t)(stop - a) * esz);
                else memcpy(hp, st, (size_t)(stop - a) * esz);
                a = stop;
                ++p;
            }
        };
        if (e1 - e0 < (int64_t)(1 << 16)) {
            for (int t = 0; t < nthreads; ++t) work(t);
        } else {
            vlgp_host_parallel_for(nthreads, work);
        }
    };
    if (to_device) {
        int buf = 0;
        for (int64_t e0 = 0; e0 < nelem; e0 += chunk, buf ^= 1) {
            const int64_t e1 = std::min(nelem, e0 + chunk);
            CK(cudaEventSynchronize(ctx->stage_ev[buf]));
            copy_chunk(buf, e0, e1);
            CK(cudaMemcpyAsync((unsigned char *)dev + (size_t)e0 * esz, ctx->h_stage[buf], (size_t)(e1 - e0) * esz,
                               cudaMemcpyHostToDevice, ctx->stream));
            CK(cudaEventRecord(ctx->stage_ev[buf], ctx->stream));
        }
        // no synchronisation here: the caller's blocks have been copied into the pinned staging buffers (they are no
        // longer needed), the staging buffers are protected by their events, and whatever uses the device array is
        // enqueued on the same stream -- so the next upload's gather overlaps this one's transfer
    } else {
        // issue chunk k+1's D2H before scattering chunk k
        int buf = 0;
        int64_t e0 = 0;
        if (nelem > 0) {
            const int64_t e1 = std::min(nelem, chunk);
            CK(cudaMemcpyAsync(ctx->h_stage[0], dev, (size_t)e1 * esz, cudaMemcpyDeviceToHost, ctx->stream));
            CK(cudaEventRecord(ctx->stage_ev[0], ctx->stream));
        }
        while (e0 < nelem) {
            const int64_t e1 = std::min(nelem, e0 + chunk);
            if (e1 < nelem) {
                const int64_t f1 = std::min(nelem, e1 + chunk);
                CK(cudaMemcpyAsync(ctx->h_stage[buf ^ 1], (const unsigned char *)dev + (size_t)e1 * esz,
                                   (size_t)(f1 - e1) * esz, cudaMemcpyDeviceToHost, ctx->stream));
                CK(cudaEventRecord(ctx->stage_ev[buf ^ 1], ctx->stream));
            }
            CK(cudaEventSynchronize(ctx->stage_ev[buf]));
            copy_chunk(buf, e0, e1);
            e0 = e1;
            buf ^= 1;
        }
        CK(cudaStreamSynchronize(ctx->stream));
    }
    return VLGP_OK;
}

static double *state_array(TrialSet *ts, int which) {
    switch (which) {
        case 0: return ts->d_mu;
        case 1: return ts->d_v;
        case 2: return ts->d_w;
        case 3: return ts->d_dmu;
        default: return nullptr;
    }
}

// Scatter of a prefetched array (pinned, already on the host) into the caller's blocks.
static int scatter_prefetched(vlgp_ctx *ctx, const double *src, int n_parts, void *const *parts, const std::vector<int64_t> &off) {
    const unsigned hw = std::thread::hardware_concurrency();
    int nthreads = (int)std::min<unsigned>(hw ? hw : 4, 8);
    for (int i = 0; i + 1 < n_parts && nthreads > 1; ++i) {          // overlapping blocks: in order, last one wins
        const unsigned char *a0 = (const unsigned char *)parts[i], *a1 = a0 + (size_t)(off[i + 1] - off[i]) * 8;
        const unsigned char *b0 = (const unsigned char *)parts[i + 1], *b1 = b0 + (size_t)(off[i + 2] - off[i + 1]) * 8;
        if (a0 < b1 && b0 < a1) nthreads = 1;
    }
    auto work = [&](int t) {
        const int p0 = (int)((int64_t)n_parts * t / nthreads), p1 = (int)((int64_t)n_parts * (t + 1) / nthreads);
        for (int p = p0; p < p1; ++p) memcpy(parts[p], src + off[p], (size_t)(off[p + 1] - off[p]) * sizeof(double));
    };
    if (nthreads == 1) work(0);
    else vlgp_host_parallel_for(nthreads, work);
    return VLGP_OK;
}

static int state_parts(vlgp_ctx *ctx, int set_id, int which, int n_parts, void *const *parts, const int64_t *rows,
                       bool to_device) {
    TrialSet *ts = get_set(ctx, set_id);
    REQUIRE(ts && parts && rows && n_parts >= 1, "trials_state_parts: bad arguments");
    double *dev = state_array(ts, which);
    REQUIRE(dev != nullptr && (which != 3 || !to_device), "trials_state_parts: bad array selector %d", which);
    CK(cudaSetDevice(ctx->device));
    if (to_device) SETTLE();
    if (to_device) ts->state_version++;
    std::vector<int64_t> off(n_parts + 1, 0);
    for (int i = 0; i < n_parts; ++i) off[i + 1] = off[i] + rows[i] * (int64_t)ctx->L;
    REQUIRE(off[n_parts] == ts->nbin * (int64_t)ctx->L, "trials_state_parts: blocks hold %lld rows, the set has %lld bins",
            (long long)(off[n_parts] / ctx->L), (long long)ts->nbin);
    if (!to_device && ctx->prefetch_set == set_id && ts->prefetch_version == ts->state_version &&
        (ts->prefetch_mask & (1 << which)) && ctx->h_prefetch) {
        CK(cudaEventSynchronize(ctx->ev_prefetch_done));
        const double *src = ts->prefetch_ext[which] ? ts->prefetch_ext[which] :
            (const double *)((unsigned char *)ctx->h_prefetch + (size_t)which * ts->nbin * ctx->L * sizeof(double));
        return scatter_prefetched(ctx, src, n_parts, parts, off);
    }
    return pipeline_copy(ctx, dev, sizeof(double), n_parts, parts, off, to_device);
}

// Starts copying the listed state arrays (bit 0 mu, 1 v, 2 w, 3 dmu) to pinned host memory on a copy stream, behind
// everything enqueued so far on the main stream.  A later vlgp_trials_get_state_parts of one of them is then served
// from that copy -- provided no entry point has written the set's state in between (otherwise it is ignored).  vem()
// issues this after the E-step of its last iteration: the transfer runs under the M- and H-step.
int vlgp_trials_prefetch_state_into(vlgp_ctx *ctx, int set_id, int which_mask, double *const *dst) {
    TrialSet *ts = get_set(ctx, set_id);
    REQUIRE(ts && which_mask > 0 && which_mask < 16, "trials_prefetch_state: bad arguments");
    CK(cudaSetDevice(ctx->device));
    const size_t bytes = (size_t)ts->nbin * ctx->L * sizeof(double);
    if (ctx->prefetch_cap < 4 * bytes) {
        if (ctx->h_prefetch) {
            CK(cudaStreamSynchronize(ctx->stream_copy));
            CK(cudaFreeHost(ctx->h_prefetch));
            ctx->h_prefetch = nullptr;
        }
        CK(cudaMallocHost(&ctx->h_prefetch, 4 * bytes));
        ctx->prefetch_cap = 4 * bytes;
    }
    if (!ctx->stream_copy) {
        CK(cudaStreamCreateWithFlags(&ctx->stream_copy, cudaStreamNonBlocking));
        CK(cudaEventCreateWithFlags(&ctx->ev_prefetch_go, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&ctx->ev_prefetch_done, cudaEventDisableTiming));
    }
    CK(cudaEventRecord(ctx->ev_prefetch_go, ctx->stream));
    CK(cudaStreamWaitEvent(ctx->stream_copy, ctx->ev_prefetch_go, 0));
    for (int k = 0; k < 4; ++k) {
        ts->prefetch_ext[k] = nullptr;
        if (!(which_mask & (1 << k))) continue;
        void *to = (unsigned char *)ctx->h_prefetch + k * bytes;
        if (dst && dst[k]) to = ts->prefetch_ext[k] = dst[k];
        CK(cudaMemcpyAsync(to, state_array(ts, k), bytes, cudaMemcpyDeviceToHost, ctx->stream_copy));
    }
    CK(cudaEventRecord(ctx->ev_prefetch_done, ctx->stream_copy));
    ctx->prefetch_set = set_id;
    ts->prefetch_mask = which_mask;
    ts->prefetch_version = ts->state_version;
    return VLGP_OK;
}

int vlgp_trials_prefetch_state(vlgp_ctx *ctx, int set_id, int which_mask) {
    return vlgp_trials_prefetch_state_into(ctx, set_id, which_mask, nullptr);
}

// Waits for the prefetch; *valid = 1 when array `which` was copied into the caller's own block (dst[which] of
// vlgp_trials_prefetch_state_into) and the set's state has not been written since: the block then IS the download.
int vlgp_trials_prefetch_take(vlgp_ctx *ctx, int set_id, int which, int *valid) {
    TrialSet *ts = get_set(ctx, set_id);
    REQUIRE(ts && valid && which >= 0 && which < 4, "trials_prefetch_take: bad arguments");
    *valid = 0;
    if (ctx->prefetch_set != set_id || ts->prefetch_version != ts->state_version || !(ts->prefetch_mask & (1 << which)) ||
        !ts->prefetch_ext[which])
        return VLGP_OK;
    CK(cudaSetDevice(ctx->device));
    CK(cudaEventSynchronize(ctx->ev_prefetch_done));
    *valid = 1;
    return VLGP_OK;
}

// Blocks until the set's prefetch (if any) has landed: its destination blocks may then be reused.
int vlgp_trials_prefetch_wait(vlgp_ctx *ctx, int set_id) {
    TrialSet *ts = get_set(ctx, set_id);
    REQUIRE(ts, "trials_prefetch_wait: bad set %d", set_id);
    if (ctx->prefetch_set != set_id || !ctx->ev_prefetch_done) return VLGP_OK;
    CK(cudaSetDevice(ctx->device));
    CK(cudaEventSynchronize(ctx->ev_prefetch_done));
    for (int k = 0; k < 4; ++k) ts->prefetch_ext[k] = nullptr;
    ts->prefetch_mask = 0;
    return VLGP_OK;
}

// Page-locked host memory for blocks that a prefetch fills directly (the Python side hands them out as the arrays of
// trial["w"] / trial["dmu"], vlgp_b200/engine.py::PinnedPool).
int vlgp_host_alloc(void **p, size_t bytes) {
    if (!p || bytes == 0) return VLGP_ERR_ARG;
    return cudaHostAlloc(p, bytes, cudaHostAllocPortable) == cudaSuccess ? VLGP_OK : VLGP_ERR_CUDA;
}

int vlgp_host_free(void *p) {
    return (!p || cudaFreeHost(p) == cudaSuccess) ? VLGP_OK : VLGP_ERR_CUDA;
}

int vlgp_trials_set_state_parts(vlgp_ctx *ctx, int set_id, int which, int n_parts, const double *const *parts,
                                const int64_t *rows) {
    return state_parts(ctx, set_id, which, n_parts, (void *const *)parts, rows, true);
}

int vlgp_trials_get_state_parts(vlgp_ctx *ctx, int set_id, int which, int n_parts, double *const *parts,
                                const int64_t *rows) {
    return state_parts(ctx, set_id, which, n_parts, (void *const *)parts, rows, false);
}

int vlgp_trials_set_state(vlgp_ctx *ctx, int set_id, const double *mu, const double *v, const double *w) {
    TrialSet *ts = get_set(ctx, set_id);
    REQUIRE(ts, "trials_set_state: bad set %d", set_id);
    CK(cudaSetDevice(ctx->device));
    SETTLE();
    const size_t bytes = (size_t)ts->nbin * ctx->L * sizeof(double);
    ts->state_version++;
    if (mu) CK(cudaMemcpyAsync(ts->d_mu, mu, bytes, cudaMemcpyHostToDevice, ctx->stream));
    if (v) CK(cudaMemcpyAsync(ts->d_v, v, bytes, cudaMemcpyHostToDevice, ctx->stream));
    if (w) CK(cudaMemcpyAsync(ts->d_w, w, bytes, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return VLGP_OK;
}

int vlgp_trials_get_state(vlgp_ctx *ctx, int set_id, double *mu, double *v, double *w, double *dmu) {
    TrialSet *ts = get_set(ctx, set_id);
    REQUIRE(ts, "trials_get_state: bad set %d", set_id);
    CK(cudaSetDevice(ctx->device));
    SETTLE();
    const size_t bytes = (size_t)ts->nbin * ctx->L * sizeof(double);
    if (mu) CK(cudaMemcpyAsync(mu, ts->d_mu, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    if (v) CK(cudaMemcpyAsync(v, ts->d_v, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    if (w) CK(cudaMemcpyAsync(w, ts->d_w, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    if (dmu) CK(cudaMemcpyAsync(dmu, ts->d_dmu, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return VLGP_OK;
}

// ---- prior factor -----------------------------------------------------------------------------------------------------
int vlgp_make_cholesky(vlgp_ctx *ctx, int set_id) {
    TrialSet *ts = get_set(ctx, set_id);
    REQUIRE(ts, "make_cholesky: bad set %d", set_id);
    CK(cudaSetDevice(ctx->device));
    const int L = ctx->L, R = ctx->rank;
    for (int l = 0; l < L; ++l) {
        ctx->h_pin[l] = ctx->h_omega[l];
        ctx->h_pin[VLGP_MAX_L + l] = ctx->h_sigma[l];
    }
    CK(cudaMemcpyAsync(ctx->d_small, ctx->h_pin, 2 * VLGP_MAX_L * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    double *work = nullptr;
    CK(vlgp_dalloc(ctx, &work, (size_t)L * R * ts->max_len * sizeof(double)));
    int rc = VLGP_OK;
    for (auto &pf : ts->factors) {
        rc = vlgp_launch_ichol(ctx, pf, ctx->d_small, ctx->d_small + VLGP_MAX_L, work);
        if (rc) break;
    }
    if (!rc) {
        for (auto &pf : ts->factors) {
            cudaError_t e = cudaMemcpyAsync(pf.h_ncol.data(), pf.d_ncol, L * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
            if (e != cudaSuccess) rc = vlgp_fail(ctx, VLGP_ERR_CUDA, "make_cholesky: %s", cudaGetErrorString(e));
        }
    }
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    vlgp_dfree(ctx, work);
    if (!rc && e != cudaSuccess) rc = vlgp_fail(ctx, VLGP_ERR_CUDA, "make_cholesky: %s", cudaGetErrorString(e));
    return rc;
}

static PriorFactor *find_factor(TrialSet *ts, int length) {
    for (auto &pf : ts->factors)
        if (pf.length == length) return &pf;
    return nullptr;
}

int vlgp_get_cholesky(vlgp_ctx *ctx, int set_id, int length, double *G, int32_t *pivots, int32_t *ncol) {
    TrialSet *ts = get_set(ctx, set_id);
    REQUIRE(ts, "get_cholesky: bad set %d", set_id);
    PriorFactor *pf = find_factor(ts, length);
    REQUIRE(pf, "get_cholesky: no trial of length %d in set %d", length, set_id);
    CK(cudaSetDevice(ctx->device));
    const size_t L = ctx->L, R = ctx->rank;
    if (G) CK(cudaMemcpyAsync(G, pf->d_G, L * length * R * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if (pivots) CK(cudaMemcpyAsync(pivots, pf->d_piv, L * R * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    if (ncol) CK(cudaMemcpyAsync(ncol, pf->d_ncol, L * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return VLGP_OK;
}

int vlgp_set_cholesky(vlgp_ctx *ctx, int set_id, int length, const double *G) {
    TrialSet *ts = get_set(ctx, set_id);
    REQUIRE(ts && G, "set_cholesky: bad set %d or NULL G", set_id);
    PriorFactor *pf = find_factor(ts, length);
    REQUIRE(pf, "set_cholesky: no trial of length %d in set %d", length, set_id);
    CK(cudaSetDevice(ctx->device));
    const size_t L = ctx->L, R = ctx->rank;
    // number of leading non-zero columns (trailing columns of an early-stopped factor are exactly zero)
    for (size_t l = 0; l < L; ++l) {
        int nc = 0;
        for (size_t t = 0; t < (size_t)length; ++t)
            for (int m = (int)R - 1; m >= nc; --m)
                if (G[(l * length + t) * R + m] != 0.0) { nc = m + 1; break; }
        pf->h_ncol[l] = nc;
    }
    CK(cudaMemcpyAsync(pf->d_G, G, L * length * R * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(pf->d_ncol, pf->h_ncol.data(), L * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return VLGP_OK;
}

// ---- E-step ---------------------------------------------------------------------------------------------------------
static int read_flag(vlgp_ctx *ctx, int idx, int *out) {
    CK(cudaMemcpyAsync(ctx->h_flags, ctx->d_flags, 16 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (out) *out = ctx->h_flags[idx];
    return VLGP_OK;
}

static int check_ready(vlgp_ctx *ctx, TrialSet *ts, const char *who) {
    REQUIRE(ts, "%s: bad trial set", who);
    REQUIRE(ts->d_y, "%s: y not set", who);
    return VLGP_OK;
}

}   // extern "C"

// Host index lists (segments, bins) travel through a stream-ordered scratch allocation: validated here, so that a kernel
// never sees an index outside the set.
template <typename T>
static int upload_indices(vlgp_ctx *ctx, const T *h, int64_t n, int64_t bound, const char *who, T **d_out) {
    REQUIRE(h && n >= 1, "%s: empty index list", who);
    for (int64_t i = 0; i < n; ++i)
        REQUIRE(h[i] >= 0 && (int64_t)h[i] < bound, "%s: index %lld at position %lld outside [0, %lld)", who,
                (long long)h[i], (long long)i, (long long)bound);
    T *d = nullptr;
    CK(vlgp_dalloc(ctx, &d, (size_t)n * sizeof(T)));
    cudaError_t e = cudaMemcpyAsync(d, h, (size_t)n * sizeof(T), cudaMemcpyHostToDevice, ctx->stream);
    if (e != cudaSuccess) {
        vlgp_dfree(ctx, d);
        CK(e);
    }
    *d_out = d;
    return VLGP_OK;
}

extern "C" {

static int estep_impl(vlgp_ctx *ctx, int set_id, int n_iter, double dmu_bound, int method_vb, const int32_t *segs,
                      int n_segs, bool subset, int *n_failed) {
    TrialSet *ts = get_set(ctx, set_id);
    int rc = check_ready(ctx, ts, "estep");
    if (rc) return rc;
    if (n_failed) *n_failed = 0;
    if (n_iter < 1) return VLGP_OK;
    if (subset && n_segs == 0) return VLGP_OK;
    CK(cudaSetDevice(ctx->device));
    SETTLE();
    int32_t *d_sub = nullptr;
    if (subset) {
        REQUIRE(n_segs > 0 && n_segs <= ts->n_trials, "estep_subset: %d segments listed, the set has %d", n_segs,
                ts->n_trials);
        // a segment listed twice would be processed by two CTAs at once
        std::vector<uint8_t> seen((size_t)ts->n_trials, 0);
        for (int i = 0; i < n_segs && segs; ++i) {
            if (segs[i] < 0 || segs[i] >= ts->n_trials) break;      // reported by upload_indices
            REQUIRE(!seen[segs[i]], "estep_subset: segment %d listed twice", segs[i]);
            seen[segs[i]] = 1;
        }
        rc = upload_indices<int32_t>(ctx, segs, n_segs, ts->n_trials, "estep_subset", &d_sub);
        if (rc) return rc;
    }
    CK(cudaMemsetAsync(ctx->d_flags, 0, sizeof(int), ctx->stream));
    ts->state_version++;
    {
        ProfScope ps(ctx, 0);
        bool handled = false;
        rc = vlgp_launch_estep_segments(ctx, ts, n_iter, dmu_bound, method_vb, &handled, d_sub, n_segs);
        if (!rc && !handled && !d_sub) rc = vlgp_launch_estep_long(ctx, ts, n_iter, dmu_bound, method_vb, &handled);
        if (!rc && !handled) rc = vlgp_launch_estep_generic(ctx, ts, 0, n_iter, dmu_bound, method_vb, d_sub, n_segs);
    }
    if (d_sub) vlgp_dfree(ctx, d_sub);
    if (rc) return rc;
    ctx->counters[1] += (int64_t)(subset ? n_segs : ts->n_trials) * ctx->L * n_iter * 2;
    return read_flag(ctx, 0, n_failed);
}

int vlgp_estep(vlgp_ctx *ctx, int set_id, int n_iter, double dmu_bound, int method_vb, int *n_failed) {
    return estep_impl(ctx, set_id, n_iter, dmu_bound, method_vb, nullptr, 0, false, n_failed);
}

int vlgp_estep_subset(vlgp_ctx *ctx, int set_id, int n_iter, double dmu_bound, int method_vb, const int32_t *segments,
                      int n_segments, int *n_failed) {
    REQUIRE(ctx && n_segments >= 0 && (segments || n_segments == 0), "estep_subset: bad arguments");
    return estep_impl(ctx, set_id, n_iter, dmu_bound, method_vb, segments, n_segments, true, n_failed);
}

int vlgp_trials_copy_rows(vlgp_ctx *ctx, int set_id, int which_mask, const int64_t *src, const int64_t *dst, int64_t n) {
    TrialSet *ts = get_set(ctx, set_id);
    REQUIRE(ts, "trials_copy_rows: bad set %d", set_id);
    REQUIRE(which_mask > 0 && which_mask < 16, "trials_copy_rows: which_mask %d (bit 0 mu, 1 v, 2 w, 3 dmu)", which_mask);
    REQUIRE(n >= 0 && (n == 0 || (src && dst)), "trials_copy_rows: bad arguments");
    if (n == 0) return VLGP_OK;
    {   // the pairs must be independent: no bin both read and written, no bin written twice
        std::vector<uint8_t> mark((size_t)ts->nbin, 0);
        for (int64_t i = 0; i < n; ++i) {
            REQUIRE(src[i] >= 0 && src[i] < ts->nbin && dst[i] >= 0 && dst[i] < ts->nbin,
                    "trials_copy_rows: pair %lld (%lld -> %lld) outside [0, %lld)", (long long)i, (long long)src[i],
                    (long long)dst[i], (long long)ts->nbin);
            REQUIRE(!(mark[dst[i]] & 2), "trials_copy_rows: bin %lld is written twice", (long long)dst[i]);
            mark[dst[i]] |= 2;
        }
        for (int64_t i = 0; i < n; ++i)
            REQUIRE(!(mark[src[i]] & 2), "trials_copy_rows: bin %lld is both a source and a destination",
                    (long long)src[i]);
    }
    CK(cudaSetDevice(ctx->device));
    SETTLE();
    ts->state_version++;
    int64_t *d_idx = nullptr;
    std::vector<int64_t> both((size_t)2 * n);
    std::copy(src, src + n, both.begin());
    std::copy(dst, dst + n, both.begin() + n);
    int rc = upload_indices<int64_t>(ctx, both.data(), 2 * n, ts->nbin, "trials_copy_rows", &d_idx);
    if (rc) return rc;
    const int L = ctx->L, nt = 256;
    copy_rows_kernel<<<(unsigned)((n * L + nt - 1) / nt), nt, 0, ctx->stream>>>(
        n, L, d_idx, d_idx + n, (which_mask & 1) ? ts->d_mu : nullptr, (which_mask & 2) ? ts->d_v : nullptr,
        (which_mask & 4) ? ts->d_w : nullptr, (which_mask & 8) ? ts->d_dmu : nullptr);
    cudaError_t e = cudaGetLastError();
    vlgp_dfree(ctx, d_idx);
    CK(e);
    ctx->counters[0]++;
    CK(cudaStreamSynchronize(ctx->stream));      // `both` is pageable host memory read by the asynchronous copy
    return VLGP_OK;
}

int vlgp_update_w(vlgp_ctx *ctx, int set_id) {
    TrialSet *ts = get_set(ctx, set_id);
    int rc = check_ready(ctx, ts, "update_w");
    if (rc) return rc;
    CK(cudaSetDevice(ctx->device));
    SETTLE();
    ts->state_version++;
    rc = vlgp_launch_estep_generic(ctx, ts, 1, 1, 0.0, 0);
    if (rc) return rc;
    CK(cudaStreamSynchronize(ctx->stream));
    return VLGP_OK;
}

int vlgp_update_v(vlgp_ctx *ctx, int set_id, int *n_failed) {
    TrialSet *ts = get_set(ctx, set_id);
    int rc = check_ready(ctx, ts, "update_v");
    if (rc) return rc;
    CK(cudaSetDevice(ctx->device));
    SETTLE();
    CK(cudaMemsetAsync(ctx->d_flags, 0, sizeof(int), ctx->stream));
    ts->state_version++;
    rc = vlgp_launch_estep_generic(ctx, ts, 2, 1, 0.0, 1);
    if (rc) return rc;
    ctx->counters[1] += (int64_t)ts->n_trials * ctx->L;
    return read_flag(ctx, 0, n_failed);
}

// ---- M-step ---------------------------------------------------------------------------------------------------------
int vlgp_mstep(vlgp_ctx *ctx, int set_id, int n_iter, int use_hessian, double eps, double lr, double da_bound,
               double db_bound, int *n_fallback) {
    TrialSet *ts = get_set(ctx, set_id);
    int rc = check_ready(ctx, ts, "mstep");
    if (rc) return rc;
    if (n_fallback) *n_fallback = 0;
    if (n_iter < 1) return VLGP_OK;
    CK(cudaSetDevice(ctx->device));
    SETTLE();
    CK(cudaMemsetAsync(ctx->d_flags + 1, 0, sizeof(int), ctx->stream));
    rc = vlgp_launch_mstep(ctx, ts, n_iter, use_hessian, eps, lr, da_bound, db_bound);
    if (rc) return rc;
    return read_flag(ctx, 1, n_fallback);
}

// The M-step on its own stream: statistics, reductions, allreduces and solves go to stream_m behind an event that orders
// them after the E-step, so the host can drive the H-step on the main stream in the meantime.  With several ranks the
// allreduces use the duplicate communicator comm_m (two streams must not share a communicator); without one the
// M-step simply runs on the main stream here and vlgp_mstep_end has nothing left to do.
// Only the first Newton iteration is enqueued here; vlgp_mstep_pump adds the rest a few at a time from inside the
// H-step objective calls (and vlgp_mstep_end / any call that needs the result adds whatever is left).
struct StreamSwap {      // run a piece of host code with ctx->stream / ctx->comm pointing at the M-step's
    vlgp_ctx *ctx;
    cudaStream_t s;
    void *c;
    explicit StreamSwap(vlgp_ctx *x) : ctx(x), s(x->stream), c(x->comm) {
        ctx->stream = ctx->stream_m;
        if (ctx->comm_m) ctx->comm = ctx->comm_m;
        ctx->p2p_chan = 1;
    }
    ~StreamSwap() {
        ctx->stream = s;
        ctx->comm = c;
        ctx->p2p_chan = 0;
    }
};

int vlgp_mstep_begin(vlgp_ctx *ctx, int set_id, int n_iter, int use_hessian, double eps, double lr, double da_bound,
                     double db_bound) {
    TrialSet *ts = get_set(ctx, set_id);
    int rc = check_ready(ctx, ts, "mstep_begin");
    if (rc) return rc;
    REQUIRE(!ctx->mstep_pending, "mstep_begin: the previous M-step has not been ended");
    if (n_iter < 1) return VLGP_OK;
    CK(cudaSetDevice(ctx->device));
    if (ctx->n_ranks > 1 && !ctx->comm_m && !ctx->p2p) {   // no second communicator / channel: in order, on the main stream
        CK(cudaMemsetAsync(ctx->d_flags + 1, 0, sizeof(int), ctx->stream));
        rc = vlgp_launch_mstep(ctx, ts, n_iter, use_hessian, eps, lr, da_bound, db_bound);
        if (rc) return rc;
        ctx->mstep_pending = true;
        return VLGP_OK;
    }
    CK(cudaEventRecord(ctx->ev_m_start, ctx->stream));
    CK(cudaStreamWaitEvent(ctx->stream_m, ctx->ev_m_start, 0));
    {
        StreamSwap swap(ctx);
        CK(cudaMemsetAsync(ctx->d_flags + 1, 0, sizeof(int), ctx->stream));
        rc = vlgp_mstep_job_setup(ctx, ts, n_iter, use_hessian, eps, lr, da_bound, db_bound);
        if (rc) return rc;
        rc = vlgp_mstep_job_pump(ctx, 1);
        if (rc) return rc;
    }
    ctx->mstep_pending = true;
    return VLGP_OK;
}

}   // extern "C"  (vlgp_mstep_pump is internal: C++ linkage, declared in common.cuh)

int vlgp_mstep_pump(vlgp_ctx *ctx, int max_iters) {
    if (!ctx->mstep_pending || vlgp_mstep_job_remaining(ctx) <= 0) return VLGP_OK;
    StreamSwap swap(ctx);
    return vlgp_mstep_job_pump(ctx, max_iters);
}

extern "C" {

int vlgp_mstep_end(vlgp_ctx *ctx, int *n_fallback) {
    if (!ctx) return VLGP_ERR_ARG;
    if (n_fallback) *n_fallback = 0;
    if (!ctx->mstep_pending) return VLGP_OK;
    CK(cudaSetDevice(ctx->device));
    int rc = settle_mstep(ctx);      // enqueue what is left, wait: whatever the host enqueues from here on sees a, b, noise
    ctx->mstep_pending = false;
    if (rc) return rc;
    rc = vlgp_p2p_check(ctx);
    if (rc) return rc;
    return read_flag(ctx, 1, n_fallback);
}

// ---- H-step ---------------------------------------------------------------------------------------------------------
int vlgp_hstep_prepare(vlgp_ctx *ctx, int set_id) {
    TrialSet *ts = get_set(ctx, set_id);
    REQUIRE(ts, "hstep_prepare: bad set %d", set_id);
    REQUIRE(ts->min_len == ts->max_len, "hstep_prepare: all segments must have the same length (vlgp/gp.py:77-80)");
    REQUIRE(ts->max_len <= VLGP_MAX_W_H, "hstep_prepare: window %d > %d", ts->max_len, VLGP_MAX_W_H);
    CK(cudaSetDevice(ctx->device));
    return vlgp_launch_hstep_prepare(ctx, ts);
}

int vlgp_hstep_objective_batch(vlgp_ctx *ctx, int set_id, int n, const int32_t *latents, const double *hypers,
                               double *ll, double *dll, int32_t *info) {
    TrialSet *ts = get_set(ctx, set_id);
    REQUIRE(ts && ts->h_prepared, "hstep_objective: call hstep_prepare first");
    REQUIRE(n >= 1 && n <= VLGP_MAX_L && latents && hypers && ll && dll && info, "hstep_objective: bad arguments");
    HEvalBatch eb{};
    eb.n = n;
    for (int e = 0; e < n; ++e) {
        REQUIRE(latents[e] >= 0 && latents[e] < ctx->L, "hstep_objective: latent %d out of range", latents[e]);
        eb.latent[e] = latents[e];
        eb.sigmasq[e] = hypers[3 * e];
        eb.omega[e] = hypers[3 * e + 1];
        eb.eps[e] = hypers[3 * e + 2];
    }
    CK(cudaSetDevice(ctx->device));
    int inf[VLGP_MAX_L];
    int rc = vlgp_launch_hstep_objective(ctx, ts, eb, ll, dll, inf);
    if (rc) return rc;
    for (int e = 0; e < n; ++e) info[e] = inf[e];
    return VLGP_OK;
}

int vlgp_hstep_objective(vlgp_ctx *ctx, int set_id, int latent, const double hyper[3], double *ll, double *dll,
                         int *info) {
    REQUIRE(hyper && ll && dll && info, "hstep_objective: bad arguments");
    int32_t l = latent, inf = 0;
    int rc = vlgp_hstep_objective_batch(ctx, set_id, 1, &l, hyper, ll, dll, &inf);
    *info = inf;
    return rc;
}

// ---- initial posterior means: FactorAnalysis.transform of every bin (vlgp/preprocess.py:36-41) ----------------------
// mu[bin] = ((y[bin] - mean) P) Cz, P = Wpsi' (N x L), Cz = cov_z (L x L) of the fitted factor model.
static __global__ void project_y_kernel(int64_t nbin, int N, int L, const void *__restrict__ y, int ydtype,
                                        const double *__restrict__ mean, const double *__restrict__ P,
                                        const double *__restrict__ Cz, double *__restrict__ mu) {
    const int lane = threadIdx.x & 31;
    const int64_t bin = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (bin >= nbin) return;
    double acc[VLGP_MAX_L];
    for (int l = 0; l < L; ++l) acc[l] = 0.0;
    for (int n = lane; n < N; n += 32) {
        const double c = load_y(y, ydtype, bin * N + n) - mean[n];
        for (int l = 0; l < L; ++l) acc[l] = fma(c, P[(size_t)n * L + l], acc[l]);
    }
    for (int l = 0; l < L; ++l) acc[l] = warp_sum(acc[l]);
    if (lane < L) {
        double s = 0.0;
        for (int k = 0; k < L; ++k) s = fma(acc[k], Cz[k * L + lane], s);
        mu[bin * L + lane] = s;
    }
}

extern "C" int vlgp_trials_project_y(vlgp_ctx *ctx, int set_id, const double *mean, const double *P, const double *Cz) {
    TrialSet *ts = get_set(ctx, set_id);
    REQUIRE(ts && ts->d_y && mean && P && Cz, "trials_project_y: bad arguments (y must be set)");
    CK(cudaSetDevice(ctx->device));
    SETTLE();
    ts->state_version++;
    const size_t N = ctx->N, L = ctx->L;
    double *d = nullptr;
    CK(vlgp_dalloc(ctx, &d, (N + N * L + L * L) * sizeof(double)));
    CK(cudaMemcpyAsync(d, mean, N * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(d + N, P, N * L * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(d + N + N * L, Cz, L * L * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    const int64_t threads = ts->nbin * 32;
    project_y_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, ctx->stream>>>(ts->nbin, (int)N, (int)L, ts->d_y, ts->ydtype,
                                                                                d, d + N, d + N + N * L, ts->d_mu);
    CKL();
    CK(vlgp_dfree(ctx, d));
    CK(cudaStreamSynchronize(ctx->stream));
    return VLGP_OK;
}

// ---- constraints / bookkeeping ---------------------------------------------------------------------------------------
static int latent_affine_impl(vlgp_ctx *ctx, int set_id, const double *shift, const double *M, const int64_t *rows,
                              int64_t n_rows, bool listed) {
    TrialSet *ts = get_set(ctx, set_id);
    REQUIRE(ts, "latent_affine: bad set %d", set_id);
    if (!shift && !M) return VLGP_OK;
    ts->state_version++;
    if (listed && n_rows == 0) return VLGP_OK;
    if (listed) {
        REQUIRE(rows && n_rows > 0, "latent_affine_rows: bad arguments");
        std::vector<uint8_t> seen((size_t)ts->nbin, 0);
        for (int64_t i = 0; i < n_rows; ++i) {
            if (rows[i] < 0 || rows[i] >= ts->nbin) break;          // reported by upload_indices
            REQUIRE(!seen[rows[i]], "latent_affine_rows: bin %lld listed twice", (long long)rows[i]);
            seen[rows[i]] = 1;
        }
    }
    CK(cudaSetDevice(ctx->device));
    SETTLE();
    const int L = ctx->L;
    int64_t *d_rows = nullptr;
    if (listed) {
        int rc = upload_indices<int64_t>(ctx, rows, n_rows, ts->nbin, "latent_affine_rows", &d_rows);
        if (rc) return rc;
    }
    for (int l = 0; l < L; ++l) ctx->h_pin[l] = shift ? shift[l] : 0.0;
    for (int i = 0; i < L * L; ++i) ctx->h_pin[L + i] = M ? M[i] : 0.0;
    CK(cudaMemcpyAsync(ctx->d_small, ctx->h_pin, (L + L * L) * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    const int nt = 256;
    const int64_t n = listed ? n_rows : ts->nbin;
    latent_affine_kernel<<<(unsigned)((n + nt - 1) / nt), nt, 0, ctx->stream>>>(n, L, ts->d_mu, ctx->d_small,
                                                                               shift != nullptr, M != nullptr, d_rows);
    cudaError_t e = cudaGetLastError();
    if (d_rows) vlgp_dfree(ctx, d_rows);
    CK(e);
    ctx->counters[0]++;
    CK(cudaStreamSynchronize(ctx->stream));
    return VLGP_OK;
}

int vlgp_latent_affine(vlgp_ctx *ctx, int set_id, const double *shift, const double *M) {
    return latent_affine_impl(ctx, set_id, shift, M, nullptr, 0, false);
}

int vlgp_latent_affine_rows(vlgp_ctx *ctx, int set_id, const double *shift, const double *M, const int64_t *rows,
                            int64_t n_rows) {
    REQUIRE(ctx && n_rows >= 0, "latent_affine_rows: bad arguments");
    return latent_affine_impl(ctx, set_id, shift, M, rows, n_rows, true);
}

static int moments(vlgp_ctx *ctx, TrialSet *ts, std::vector<double> &out) {
    const int L = ctx->L, K = 2 * L + 1;
    int grid = 2 * ctx->prop.multiProcessorCount;
    if ((int64_t)grid * 256 > ts->nbin) grid = (int)((ts->nbin + 255) / 256);
    double *part = nullptr;
    CK(vlgp_dalloc(ctx, &part, (size_t)(grid + 1) * K * sizeof(double)));
    moments_kernel<<<grid, 256, 0, ctx->stream>>>(ts->nbin, L, ts->d_mu, ts->d_dmu, part);
    CKL();
    double *res = part + (size_t)grid * K;
    reduce_parts_kernel3<<<1, 64, 0, ctx->stream>>>(part, grid, K, res);
    CKL();
    int rc = ctx->shm ? VLGP_OK : vlgp_allreduce_dev(ctx, res, K, 0);
    if (rc) { vlgp_dfree(ctx, part); return rc; }
    CK(cudaMemcpyAsync(ctx->h_pin, res, K * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    out.assign(ctx->h_pin, ctx->h_pin + K);
    if (ctx->shm) {      // the host consumes these: sum them on the host
        rc = vlgp_comm_allreduce(ctx, out.data(), K, 0);
        if (rc) { vlgp_dfree(ctx, part); return rc; }
    }
    CK(vlgp_dfree(ctx, part));
    return VLGP_OK;
}

int vlgp_norms(vlgp_ctx *ctx, int set_id, double out[2]) {
    TrialSet *ts = get_set(ctx, set_id);
    REQUIRE(ts && out, "norms: bad arguments");
    CK(cudaSetDevice(ctx->device));
    std::vector<double> m;
    int rc = moments(ctx, ts, m);
    if (rc) return rc;
    const int L = ctx->L;
    double s = 0.0;
    for (int l = 0; l < L; ++l) s += m[L + l];
    out[0] = s;
    out[1] = m[2 * L];
    return VLGP_OK;
}

int vlgp_latent_moments(vlgp_ctx *ctx, int set_id, double *sum, double *sumsq, int64_t *count) {
    TrialSet *ts = get_set(ctx, set_id);
    REQUIRE(ts, "latent_moments: bad set %d", set_id);
    CK(cudaSetDevice(ctx->device));
    std::vector<double> m;
    int rc = moments(ctx, ts, m);
    if (rc) return rc;
    const int L = ctx->L;
    for (int l = 0; l < L; ++l) {
        if (sum) sum[l] = m[l];
        if (sumsq) sumsq[l] = m[L + l];
    }
    if (count) {
        double c = (double)ts->nbin;
        if (ctx->n_ranks > 1) {
            rc = vlgp_comm_allreduce(ctx, &c, 1, 0);
            if (rc) return rc;
        }
        *count = (int64_t)(c + 0.5);
    }
    return VLGP_OK;
}

// ---- measurement helpers -------------------------------------------------------------------------------------------
int vlgp_peak_fp64(vlgp_ctx *ctx, double *dfma_tflops, double *dmma_tflops) {
    if (!ctx) return VLGP_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    const int grid = ctx->prop.multiProcessorCount * 8, nt = 256, iters = 4096;
    float ms = 0.f;
    for (int which = 0; which < 2; ++which) {
        double best = 0.0;
        for (int rep = 0; rep < 4; ++rep) {
            CK(cudaEventRecord(ctx->ev0, ctx->stream));
            if (which == 0) dfma_peak_kernel<<<grid, nt, 0, ctx->stream>>>(iters, ctx->d_small);
            else dmma_peak_kernel<<<grid, nt, 0, ctx->stream>>>(iters, ctx->d_small);
            CKL();
            CK(cudaEventRecord(ctx->ev1, ctx->stream));
            CK(cudaEventSynchronize(ctx->ev1));
            CK(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
            // DFMA: 16 FMA per thread per iteration; DMMA: 8 mma of 8x8x4 per warp per iteration = 8*256 FMA per warp
            const double fma_count = which == 0 ? (double)grid * nt * iters * 16.0
                                                : (double)grid * (nt / 32) * iters * 8.0 * 256.0;
            const double tf = 2.0 * fma_count / (ms * 1e-3) / 1e12;
            if (rep > 0 && tf > best) best = tf;
        }
        if (which == 0 && dfma_tflops) *dfma_tflops = best;
        if (which == 1 && dmma_tflops) *dmma_tflops = best;
    }
    return VLGP_OK;
}

int vlgp_peak_hbm(vlgp_ctx *ctx, uint64_t nbytes, double *gbs) {
    if (!ctx || !gbs) return VLGP_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    nbytes = (nbytes / 32) * 32;
    REQUIRE(nbytes >= 32, "peak_hbm: nbytes too small");
    void *src = nullptr, *dst = nullptr;
    CK(cudaMalloc(&src, nbytes));
    CK(cudaMalloc(&dst, nbytes));
    CK(cudaMemsetAsync(src, 1, nbytes, ctx->stream));
    const int grid = ctx->prop.multiProcessorCount * 16;
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
        float ms = 0.f;
        CK(cudaEventRecord(ctx->ev0, ctx->stream));
        copy_kernel<<<grid, 512, 0, ctx->stream>>>((const double4 *)src, (double4 *)dst, nbytes / 32);
        CKL();
        CK(cudaEventRecord(ctx->ev1, ctx->stream));
        CK(cudaEventSynchronize(ctx->ev1));
        CK(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
        const double g = 2.0 * (double)nbytes / (ms * 1e-3) / 1e9;
        if (rep > 0 && g > best) best = g;
    }
    CK(cudaFree(src));
    CK(cudaFree(dst));
    *gbs = best;
    return VLGP_OK;
}

int vlgp_flush_l2(vlgp_ctx *ctx) {
    if (!ctx) return VLGP_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    if (!ctx->d_flush) {
        ctx->flush_bytes = (size_t)256 << 20;     // 256 MiB > 126 MB L2
        CK(cudaMalloc(&ctx->d_flush, ctx->flush_bytes));
    }
    fill_kernel<<<ctx->prop.multiProcessorCount * 8, 512, 0, ctx->stream>>>((double4 *)ctx->d_flush, ctx->flush_bytes / 32);
    CKL();
    ctx->counters[0]--;   // not one of the engine's kernels
    return VLGP_OK;
}

int vlgp_set_precision(vlgp_ctx *ctx, int bits) {
    if (!ctx) return VLGP_ERR_ARG;
    REQUIRE(bits == 32 || bits == 64, "set_precision: bits must be 32 or 64, not %d", bits);
    ctx->estep_f32 = bits == 32 ? 1 : 0;
    return VLGP_OK;
}

int vlgp_profile_enable(vlgp_ctx *ctx, int mask) {
    if (!ctx) return VLGP_ERR_ARG;
    ctx->profile = mask;
    for (int i = 0; i < 4; ++i) {
        ctx->prof_ms[i] = 0.0;
        ctx->prof_n[i] = 0;
    }
    return VLGP_OK;
}

int vlgp_profile_get(vlgp_ctx *ctx, int which, double *total_ms, int64_t *n) {
    if (!ctx || which < 0 || which > 3) return VLGP_ERR_ARG;
    if (total_ms) *total_ms = ctx->prof_ms[which];
    if (n) *n = ctx->prof_n[which];
    return VLGP_OK;
}

}   // extern "C"
