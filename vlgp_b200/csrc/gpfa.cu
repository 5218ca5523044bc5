// GPFA branch: the Gaussian-likelihood EM of vlgp/gpfa.py:20-56 on equal-length segments.
//
// Reference E-step: z = A B^-1 (y - d) with bigC = kron(C', I_n), A = bigK bigC', B = bigC A + bigR -- an (n ydim)^2 dense
// solve per iteration.  With the (diagonal) noise moved to the other side (push-through identity, checked against the
// reference to 1e-14 in oracle/make_golden.py's prototype):
//     z = P h,   h = bigC' bigR^-1 (y - d)   (L n values per segment),   P = (I + bigK S)^-1 bigK,   S = bigC' bigR^-1 bigC
// S is block-diagonal in time, P is (L n) x (L n) and the same for every segment: the host forms it (250 x 250 at the
// reference's window), the device applies it to all segments.  Two quirks of the reference are kept, because a drop-in
// must give its numbers: bigR = kron(I_n, R) is indexed time-major while everything else is neuron-major, so the noise
// applied to (neuron j, bin t) is R[(j n + t) mod ydim] -- the host passes that table as rho[t][j] = 1 / R[...] -- and bigR
// is built ONCE before the loop, so the E-step never sees the M-step's R.
//   gpfa_project_kernel : h[seg][l n + t] = sum_j C[l][j] rho[t][j] (y[seg][t][j] - d[j])
//   gpfa_apply_kernel   : mu[seg][t][l]   = sum_k P[l n + t][k] h[seg][k]           (P passed transposed: coalesced)
//   gpfa_stats_kernel   : Z1'Z1, Z1'Y, sum y^2 with Z1 = [mu, 1]: the normal equations of the reference's lstsq M-step
#include "common.cuh"

namespace {

constexpr int NT = 256;
constexpr int SEGB = 8;      // segments per CTA in the apply kernel

__global__ void __launch_bounds__(NT) gpfa_project_kernel(int n_seg, int W, int N, int L, const void *__restrict__ y, int ydtype,
                                                          const double *__restrict__ C, const double *__restrict__ d,
                                                          const double *__restrict__ rho, double *__restrict__ h) {
    extern __shared__ double sm[];                 // W x N weighted residuals of this segment
    const int seg = blockIdx.x;
    const int64_t bin0 = (int64_t)seg * W;
    for (int i = threadIdx.x; i < W * N; i += NT) {
        const int j = i % N;
        sm[i] = rho[i] * (load_y(y, ydtype, bin0 * N + i) - d[j]);
    }
    __syncthreads();
    for (int o = threadIdx.x; o < L * W; o += NT) {
        const int l = o / W, t = o - l * W;
        const double *r = sm + t * N, *c = C + (size_t)l * N;
        double a = 0.0;
        for (int j = 0; j < N; ++j) a = fma(c[j], r[j], a);
        h[(size_t)seg * L * W + o] = a;
    }
}

__global__ void __launch_bounds__(NT) gpfa_apply_kernel(int n_seg, int W, int L, const double *__restrict__ PT,
                                                        const double *__restrict__ h, double *__restrict__ mu) {
    extern __shared__ double sh[];                 // SEGB x (L W)
    const int LW = L * W, s0 = blockIdx.x * SEGB, ns = min(SEGB, n_seg - s0);
    for (int i = threadIdx.x; i < SEGB * LW; i += NT) sh[i] = i < ns * LW ? h[(size_t)s0 * LW + i] : 0.0;
    __syncthreads();
    for (int o = threadIdx.x; o < LW; o += NT) {
        double acc[SEGB];
#pragma unroll
        for (int s = 0; s < SEGB; ++s) acc[s] = 0.0;
        for (int k = 0; k < LW; ++k) {
            const double p = PT[(size_t)k * LW + o];
#pragma unroll
            for (int s = 0; s < SEGB; ++s) acc[s] = fma(p, sh[s * LW + k], acc[s]);
        }
        const int l = o / W, t = o - l * W;
#pragma unroll
        for (int s = 0; s < SEGB; ++s)
            if (s < ns) mu[((size_t)(s0 + s) * W + t) * L + l] = acc[s];
    }
}

// part[cta][ (L+1) x N | N | (L+1)^2 ]
__global__ void __launch_bounds__(NT) gpfa_stats_kernel(int64_t nbin, int N, int L, const void *__restrict__ y, int ydtype,
                                                        const double *__restrict__ mu, double *__restrict__ part) {
    __shared__ double zs[64][VLGP_MAX_L + 1];
    const int L1 = L + 1;
    const int64_t per = (nbin + gridDim.x - 1) / gridDim.x;
    const int64_t b0 = (int64_t)blockIdx.x * per, b1 = min(nbin, b0 + per);
    double *out = part + (size_t)blockIdx.x * ((size_t)L1 * N + N + L1 * L1);
    double gacc = 0.0;                              // thread a * L1 + b < L1^2 accumulates Z1'Z1[a][b]
    const int ga = threadIdx.x / L1, gb = threadIdx.x - ga * L1;
    for (int n0 = 0; n0 < N; n0 += NT) {            // neuron slots of 256
        const int n = n0 + threadIdx.x;
        double acc[VLGP_MAX_L + 1], yy = 0.0;
#pragma unroll
        for (int a = 0; a <= VLGP_MAX_L; ++a) acc[a] = 0.0;
        for (int64_t t0 = b0; t0 < b1; t0 += 64) {
            const int nb = (int)min((int64_t)64, b1 - t0);
            __syncthreads();
            for (int i = threadIdx.x; i < nb * L1; i += NT) {
                const int t = i / L1, a = i - t * L1;
                zs[t][a] = a < L ? mu[(t0 + t) * L + a] : 1.0;
            }
            __syncthreads();
            if (n < N) {
                for (int t = 0; t < nb; ++t) {
                    const double yv = load_y(y, ydtype, (t0 + t) * N + n);
                    yy = fma(yv, yv, yy);
#pragma unroll
                    for (int a = 0; a <= VLGP_MAX_L; ++a)
                        if (a < L1) acc[a] = fma(zs[t][a], yv, acc[a]);
                }
            }
            if (n0 == 0 && threadIdx.x < L1 * L1)
                for (int t = 0; t < nb; ++t) gacc = fma(zs[t][ga], zs[t][gb], gacc);
        }
        if (n < N) {
            for (int a = 0; a < L1; ++a) out[(size_t)a * N + n] = acc[a];
            out[(size_t)L1 * N + n] = yy;
        }
    }
    if (threadIdx.x < L1 * L1) out[(size_t)L1 * N + N + threadIdx.x] = gacc;
}

__global__ void gpfa_reduce_kernel(const double *__restrict__ part, int G, int K, double *__restrict__ out) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= K) return;
    double s = 0.0;
    for (int g = 0; g < G; ++g) s += part[(size_t)g * K + k];
    out[k] = s;
}

}   // namespace

extern "C" {

// mu <- E-step of vlgp/gpfa.py:37-46 (before the mean subtraction) for every segment of the set.  C: L x N, d: N,
// rho: W x N (1 / noise applied to (bin t, neuron j)), PT: (L W) x (L W) = P transposed, vec index l W + t.
int vlgp_gpfa_estep(vlgp_ctx *ctx, int set_id, const double *C, const double *d, const double *rho, const double *PT) {
    TrialSet *ts = get_set(ctx, set_id);
    REQUIRE(ts && ts->d_y && C && d && rho && PT, "gpfa_estep: bad arguments (y must be set)");
    REQUIRE(ts->min_len == ts->max_len, "gpfa_estep: all segments must have the same length");
    CK(cudaSetDevice(ctx->device));
    const int W = ts->max_len, N = ctx->N, L = ctx->L, S = ts->n_trials, LW = L * W;
    const size_t smem_p = (size_t)W * N * sizeof(double), smem_a = (size_t)SEGB * LW * sizeof(double);
    REQUIRE(smem_p <= (size_t)ctx->prop.sharedMemPerBlockOptin && smem_a <= (size_t)ctx->prop.sharedMemPerBlockOptin,
            "gpfa_estep: window %d x %d neurons / %d latents does not fit in shared memory", W, N, L);
    ts->state_version++;
    double *buf = nullptr;
    const size_t nd = (size_t)L * N + N + (size_t)W * N + (size_t)LW * LW + (size_t)S * LW;
    CK(vlgp_dalloc(ctx, &buf, nd * sizeof(double)));
    double *dC = buf, *dd = dC + (size_t)L * N, *drho = dd + N, *dP = drho + (size_t)W * N, *dh = dP + (size_t)LW * LW;
    CK(cudaMemcpyAsync(dC, C, (size_t)L * N * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(dd, d, (size_t)N * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(drho, rho, (size_t)W * N * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(dP, PT, (size_t)LW * LW * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    if (smem_p > 48 * 1024)
        CK(cudaFuncSetAttribute(gpfa_project_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_p));
    if (smem_a > 48 * 1024)
        CK(cudaFuncSetAttribute(gpfa_apply_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_a));
    gpfa_project_kernel<<<S, NT, smem_p, ctx->stream>>>(S, W, N, L, ts->d_y, ts->ydtype, dC, dd, drho, dh);
    CKL();
    gpfa_apply_kernel<<<(S + SEGB - 1) / SEGB, NT, smem_a, ctx->stream>>>(S, W, L, dP, dh, ts->d_mu);
    CKL();
    CK(cudaStreamSynchronize(ctx->stream));          // the host arrays may be released by the caller
    CK(vlgp_dfree(ctx, buf));
    return VLGP_OK;
}

// Normal equations of the least-squares M-step (vlgp/gpfa.py:49-53,83-88) with Z1 = [mu, 1]:
// ZtZ (L+1) x (L+1), ZtY (L+1) x N, yy[n] = sum y^2.
int vlgp_gpfa_stats(vlgp_ctx *ctx, int set_id, double *ZtZ, double *ZtY, double *yy) {
    TrialSet *ts = get_set(ctx, set_id);
    REQUIRE(ts && ts->d_y && ZtZ && ZtY && yy, "gpfa_stats: bad arguments (y must be set)");
    CK(cudaSetDevice(ctx->device));
    const int N = ctx->N, L = ctx->L, L1 = L + 1;
    REQUIRE(L1 * L1 <= NT, "gpfa_stats: too many latents");
    const int K = L1 * N + N + L1 * L1;
    int grid = 2 * ctx->prop.multiProcessorCount;
    if ((int64_t)grid > (ts->nbin + 63) / 64) grid = (int)((ts->nbin + 63) / 64);
    if (grid < 1) grid = 1;
    double *part = nullptr;
    CK(vlgp_dalloc(ctx, &part, ((size_t)grid + 1) * K * sizeof(double)));
    double *red = part + (size_t)grid * K;
    gpfa_stats_kernel<<<grid, NT, 0, ctx->stream>>>(ts->nbin, N, L, ts->d_y, ts->ydtype, ts->d_mu, part);
    CKL();
    gpfa_reduce_kernel<<<(K + 255) / 256, 256, 0, ctx->stream>>>(part, grid, K, red);
    CKL();
    std::vector<double> host((size_t)K);
    CK(cudaMemcpyAsync(host.data(), red, (size_t)K * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    CK(vlgp_dfree(ctx, part));
    memcpy(ZtY, host.data(), (size_t)L1 * N * sizeof(double));
    memcpy(yy, host.data() + (size_t)L1 * N, (size_t)N * sizeof(double));
    memcpy(ZtZ, host.data() + (size_t)L1 * N + N, (size_t)L1 * L1 * sizeof(double));
    return VLGP_OK;
}

}   // extern "C"
