// General regressors x (bins x xdim x neurons): the linear predictor eta = mu a + einsum(x, b) of the E-step
// (vlgp/core.py:66-69) and the regression part of the M-step (:205-220 Poisson Newton step on b with design x[..., n],
// :221-235 Gaussian least squares with b[1:, n] = 0).  The reference's default regressor is the all-ones bias column
// (xdim = 1, vlgp/preprocess.py:43-44), which the tuned kernels treat as a per-neuron constant; any other x --
// xdim = max(history, 1) > 1, or a user-supplied design -- takes the path in this file:
//   * xb[bin][n] = sum_k x[bin][k][n] b[k][n] is formed once per parameter change and replaces b[n] as the offset of the
//     linear predictor in the E-step / update_w kernels and in the M-step statistics kernel;
//   * one extra statistics pass per Newton iteration accumulates, per neuron, x'(y - r), x' diag(r) x (Poisson) or
//     x'y, x'x, x'mu, mu'xb (Gaussian), reduced over CTAs and ranks like the loading statistics;
//   * one thread per neuron solves the xdim x xdim system.
// Correctness first: this path is rare (BASELINE's configurations all have xdim = 1) and is not tuned.
#include "common.cuh"

namespace {

constexpr int XMAX = VLGP_MAX_XDIM;

__global__ void xb_kernel(int64_t nbin, int N, int xd, const double *__restrict__ x, const double *__restrict__ b,
                          double *__restrict__ xb) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nbin * N) return;
    const int64_t bin = i / N;
    const int n = (int)(i - bin * N);
    double s = 0.0;
    for (int k = 0; k < xd; ++k) s = fma(x[(bin * xd + k) * N + n], b[(size_t)k * N + n], s);
    xb[i] = s;
}

// slots per neuron: [0, xd) g ; [xd, xd + xd(xd+1)/2) H (packed lower) ; then xd * L: x'mu ; then L: mu'xb
__host__ __device__ inline int nbstat(int xd, int L) { return xd + xd * (xd + 1) / 2 + xd * L + L; }

template <int LT>
__global__ void __launch_bounds__(128) mstep_bstats_kernel(int64_t nbin, int N, int xd, const void *__restrict__ y,
                                                           int ydtype, const double *__restrict__ x,
                                                           const double *__restrict__ xb, const double *__restrict__ mu,
                                                           const double *__restrict__ v, const double *__restrict__ a,
                                                           const uint8_t *__restrict__ poisson, double *__restrict__ part) {
    const int n = blockIdx.y * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const int KB = nbstat(xd, LT);
    const int64_t per = (nbin + gridDim.x - 1) / gridDim.x;
    const int64_t b0 = (int64_t)blockIdx.x * per, b1 = b0 + per < nbin ? b0 + per : nbin;
    double acc[XMAX + XMAX * (XMAX + 1) / 2 + XMAX * LT + LT];
    for (int s = 0; s < KB; ++s) acc[s] = 0.0;
    double al[LT];
#pragma unroll
    for (int l = 0; l < LT; ++l) al[l] = a[l * N + n];
    const bool pois = poisson[n] != 0;
    for (int64_t t = b0; t < b1; ++t) {
        double eta = xb[t * N + n], h = 0.0, m[LT];
#pragma unroll
        for (int l = 0; l < LT; ++l) {
            m[l] = mu[t * LT + l];
            eta = fma(m[l], al[l], eta);
            h = fma(v[t * LT + l], al[l] * al[l], h);
        }
        const double yv = load_y(y, ydtype, t * N + n);
        const double r = pois ? trunc_exp(eta + 0.5 * h) : 1.0;
        const double res = pois ? yv - r : yv;
        double xk[XMAX];
        for (int k = 0; k < xd; ++k) xk[k] = x[(t * xd + k) * N + n];
        int q = xd;
        for (int k = 0; k < xd; ++k) {
            acc[k] = fma(xk[k], res, acc[k]);
            const double rx = r * xk[k];
            for (int j = 0; j <= k; ++j) {
                acc[q] = fma(rx, xk[j], acc[q]);
                ++q;
            }
        }
        if (!pois) {
            for (int k = 0; k < xd; ++k)
#pragma unroll
                for (int l = 0; l < LT; ++l) acc[q + k * LT + l] = fma(xk[k], m[l], acc[q + k * LT + l]);
            q += xd * LT;
            const double xbv = xb[t * N + n];
#pragma unroll
            for (int l = 0; l < LT; ++l) acc[q + l] = fma(m[l], xbv, acc[q + l]);
        }
    }
    for (int s = 0; s < KB; ++s) part[((size_t)blockIdx.x * KB + s) * N + n] = acc[s];
}

__global__ void reduce_parts_kernel3(const double *__restrict__ part, int G, int K, double *__restrict__ out) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= K) return;
    double s = 0.0;
    for (int g = 0; g < G; ++g) s += part[(size_t)g * K + k];
    out[k] = s;
}

// Cholesky solve of an n x n SPD system held in local arrays (n <= XMAX); false if not positive definite
__device__ inline bool chol_solve_dyn(double (*H)[XMAX], double *g, int n) {
    for (int k = 0; k < n; ++k) {
        double d = H[k][k];
        for (int m = 0; m < k; ++m) d = fma(-H[k][m], H[k][m], d);
        if (!(d > 0.0)) return false;
        const double lkk = sqrt(d);
        H[k][k] = lkk;
        for (int i = k + 1; i < n; ++i) {
            double s = H[i][k];
            for (int m = 0; m < k; ++m) s = fma(-H[i][m], H[k][m], s);
            H[i][k] = s / lkk;
        }
    }
    for (int i = 0; i < n; ++i) {
        double s = g[i];
        for (int m = 0; m < i; ++m) s = fma(-H[i][m], g[m], s);
        g[i] = s / H[i][i];
    }
    for (int i = n - 1; i >= 0; --i) {
        double s = g[i];
        for (int m = i + 1; m < n; ++m) s = fma(-H[m][i], g[m], s);
        g[i] = s / H[i][i];
    }
    return true;
}

// b[:, n] update (after the loading update of the same iteration, so Gaussian channels see the NEW a like the
// reference's sequential code, vlgp/core.py:226-234)
template <int LT>
__global__ void mstep_bsolve_kernel(int N, int xd, const double *__restrict__ stat, const uint8_t *__restrict__ poisson,
                                    const double *__restrict__ a, double *__restrict__ b, double *__restrict__ db,
                                    int use_hessian, double eps, double lr, double db_bound, int *flags) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    auto S = [&](int s) { return stat[(size_t)s * N + n]; };
    double g[XMAX], H[XMAX][XMAX];
    int q = xd;
    for (int k = 0; k < xd; ++k) {
        g[k] = S(k);
        for (int j = 0; j <= k; ++j) {
            const double hv = S(q++);
            H[k][j] = hv;
            H[j][k] = hv;
        }
    }
    if (poisson[n]) {
        double step[XMAX];
        bool newton = use_hessian != 0;
        if (newton) {
            for (int k = 0; k < xd; ++k) {
                H[k][k] += eps;
                step[k] = g[k];
            }
            if (!chol_solve_dyn(H, step, xd)) {
                newton = false;
                atomicAdd(flags + 1, 1);
            }
        }
        if (!newton)
            for (int k = 0; k < xd; ++k) step[k] = lr * g[k];
        for (int k = 0; k < xd; ++k) {
            const double d = clipd(step[k], db_bound);
            db[(size_t)k * N + n] = d;
            b[(size_t)k * N + n] += d;
        }
    } else {
        // (x'x)^-1 x'(y - mu a_new), then every regression weight but the first is zeroed (vlgp/core.py:229-235)
        for (int k = 0; k < xd; ++k) {
            double s = g[k];
#pragma unroll
            for (int l = 0; l < LT; ++l) s = fma(-S(q + k * LT + l), a[l * N + n], s);
            g[k] = s;
        }
        if (chol_solve_dyn(H, g, xd)) {
            b[n] = g[0];
            for (int k = 1; k < xd; ++k) b[(size_t)k * N + n] = 0.0;
        } else {
            atomicAdd(flags + 1, 1);
        }
    }
}

}   // namespace

int vlgp_launch_xb(vlgp_ctx *ctx, TrialSet *ts) {
    if (!ts->d_x) return VLGP_OK;
    const int64_t tot = ts->nbin * ctx->N;
    xb_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, ctx->stream>>>(ts->nbin, ctx->N, ctx->xdim, ts->d_x, ctx->d_b, ts->d_xb);
    CKL();
    return VLGP_OK;
}

int vlgp_bstat_count(vlgp_ctx *ctx) { return nbstat(ctx->xdim, ctx->L); }

// offset (in statistics slots) of the Gaussian channels' mu'xb block inside the reduced b-statistics
int vlgp_bstat_muxb_offset(vlgp_ctx *ctx) {
    const int xd = ctx->xdim;
    return xd + xd * (xd + 1) / 2 + xd * ctx->L;
}

template <int LT>
static int bstats_t(vlgp_ctx *ctx, TrialSet *ts) {
    const int N = ctx->N, xd = ctx->xdim, KB = nbstat(xd, LT);
    int gx = 2 * ctx->prop.multiProcessorCount / ((N + 127) / 128);
    if (gx < 1) gx = 1;
    if (gx > ts->nbin) gx = (int)ts->nbin;
    const size_t need = (size_t)gx * KB * N;
    if (ctx->bpart_len < need) {
        if (ctx->d_bpart) CK(cudaFree(ctx->d_bpart));
        ctx->d_bpart = nullptr;
        CK(cudaMalloc(&ctx->d_bpart, need * sizeof(double)));
        ctx->bpart_len = need;
    }
    if (!ctx->d_bstat) CK(cudaMalloc(&ctx->d_bstat, (size_t)nbstat(XMAX, VLGP_MAX_L) * N * sizeof(double)));
    mstep_bstats_kernel<LT><<<dim3(gx, (N + 127) / 128), 128, 0, ctx->stream>>>(ts->nbin, N, xd, ts->d_y, ts->ydtype, ts->d_x,
                                                                               ts->d_xb, ts->d_mu, ts->d_v, ctx->d_a,
                                                                               ctx->d_poisson, ctx->d_bpart);
    CKL();
    reduce_parts_kernel3<<<(KB * N + 127) / 128, 128, 0, ctx->stream>>>(ctx->d_bpart, gx, KB * N, ctx->d_bstat);
    CKL();
    return vlgp_allreduce_dev(ctx, ctx->d_bstat, (size_t)KB * N, 0);
}

template <int LT>
static int bsolve_t(vlgp_ctx *ctx, int use_hessian, double eps, double lr, double db_bound) {
    const int N = ctx->N;
    mstep_bsolve_kernel<LT><<<(N + 63) / 64, 64, 0, ctx->stream>>>(N, ctx->xdim, ctx->d_bstat, ctx->d_poisson, ctx->d_a,
                                                                    ctx->d_b, ctx->d_db, use_hessian, eps, lr, db_bound,
                                                                    ctx->d_flags);
    CKL();
    return VLGP_OK;
}

#define DISPATCH_LX(L, CALL)                                                            \
    switch (L) {                                                                        \
        case 1: { constexpr int LT = 1; CALL; } break;                                  \
        case 2: { constexpr int LT = 2; CALL; } break;                                  \
        case 3: { constexpr int LT = 3; CALL; } break;                                  \
        case 4: { constexpr int LT = 4; CALL; } break;                                  \
        case 5: { constexpr int LT = 5; CALL; } break;                                  \
        case 6: { constexpr int LT = 6; CALL; } break;                                  \
        case 7: { constexpr int LT = 7; CALL; } break;                                  \
        case 8: { constexpr int LT = 8; CALL; } break;                                  \
        case 9: { constexpr int LT = 9; CALL; } break;                                  \
        case 10: { constexpr int LT = 10; CALL; } break;                                \
        case 11: { constexpr int LT = 11; CALL; } break;                                \
        case 12: { constexpr int LT = 12; CALL; } break;                                \
        default: return vlgp_fail(ctx, VLGP_ERR_UNSUPPORTED, "n_latents %d > 12", L);   \
    }

int vlgp_launch_bstats(vlgp_ctx *ctx, TrialSet *ts) {
    int rc = VLGP_OK;
    DISPATCH_LX(ctx->L, rc = bstats_t<LT>(ctx, ts));
    return rc;
}

int vlgp_launch_bsolve(vlgp_ctx *ctx, int use_hessian, double eps, double lr, double db_bound) {
    int rc = VLGP_OK;
    DISPATCH_LX(ctx->L, rc = bsolve_t<LT>(ctx, use_hessian, eps, lr, db_bound));
    return rc;
}
