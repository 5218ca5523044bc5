// Full posterior covariance of one latent of one trial: replaces the dense T x T algebra of api.sample_posterior
// (vlgp/api.py:160-166: inv(inv(K + reg I) + W) with K = G G') and util.posterior_cov (vlgp/util.py:541-547:
// K - K (1/W + K)^-1 K, the reg = 0 case).  The reference inverts two T x T matrices per latent (K + reg I has condition
// number ~1e8 at the default reg = 1e-6); with the rank-r factor G (T x r) of the prior the same matrix is, exactly,
//     cov = reg E^-1 + F Q^-1 F',   E = I + reg W (diagonal),  F = E^-1 G,  Q = I_r + G' (W E^-1) G     (r x r, SPD)
// (Woodbury twice; checked against the reference's expression to 3e-10, its own asymmetry, tests/test_gpu_parity.py),
// i.e. one r x r inverse and a T x r by r x T product: O(T r^2 + T^2 r) instead of O(T^3).
//   pcov_factor_kernel : one CTA: Q from the factor and the weights, symmetric sweep, H = F Q^-1 (T x r) to scratch
//   pcov_outer_kernel  : cov[i][j] = H_i . F_j + delta_ij reg / E_i, 16 x 16 output tiles
#include "common.cuh"
#include "linalg.cuh"

namespace {

constexpr int NT = 256;

__global__ void __launch_bounds__(NT) pcov_factor_kernel(int T, int rank, int L, int latent, double reg,
                                                         const double *__restrict__ G,      // T x rank of this latent
                                                         const double *__restrict__ w,      // T x L (this trial's rows)
                                                         double *__restrict__ H, int *flag) {
    extern __shared__ double sm[];
    const int ld = rank | 1;
    double *Q = sm;                    // rank x ld
    double *ck = Q + rank * ld;        // 2 x 64
    const int tid = threadIdx.x;
    // Q = I + G' diag(w / (1 + reg w)) G
    for (int e = tid; e < rank * rank; e += NT) {
        const int i = e / rank, j = e - i * rank;
        if (j > i) continue;
        double acc = 0.0;
        for (int t = 0; t < T; ++t) {
            const double wt = w[(size_t)t * L + latent];
            acc = fma(G[(size_t)t * rank + i] * (wt / (1.0 + reg * wt)), G[(size_t)t * rank + j], acc);
        }
        if (i == j) acc += 1.0;
        Q[i * ld + j] = acc;
        Q[j * ld + i] = acc;
    }
    __syncthreads();
    const bool ok = block_sweep_spd(Q, ld, rank, ck, nullptr);       // Q <- -Q^-1
    if (!ok) {
        if (tid == 0) atomicAdd(flag, 1);
        return;
    }
    // H = F Q^-1,  F_t = G_t / (1 + reg w_t)
    for (int e = tid; e < T * rank; e += NT) {
        const int t = e / rank, j = e - t * rank;
        const double wt = w[(size_t)t * L + latent];
        const double sc = 1.0 / (1.0 + reg * wt);
        double acc = 0.0;
        for (int k = 0; k < rank; ++k) acc = fma(G[(size_t)t * rank + k], Q[k * ld + j], acc);
        H[e] = -acc * sc;
    }
}

__global__ void __launch_bounds__(256) pcov_outer_kernel(int T, int rank, int L, int latent, double reg,
                                                         const double *__restrict__ G, const double *__restrict__ w,
                                                         const double *__restrict__ H, double *__restrict__ cov) {
    __shared__ double Hs[16][65], Fs[16][65];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int i0 = blockIdx.y * 16, j0 = blockIdx.x * 16;
    for (int e = threadIdx.x; e < 16 * rank; e += 256) {
        const int r = e / rank, k = e - r * rank;
        const int i = i0 + r, j = j0 + r;
        Hs[r][k] = i < T ? H[(size_t)i * rank + k] : 0.0;
        double f = 0.0;
        if (j < T) {
            const double wt = w[(size_t)j * L + latent];
            f = G[(size_t)j * rank + k] / (1.0 + reg * wt);
        }
        Fs[r][k] = f;
    }
    __syncthreads();
    const int i = i0 + ty, j = j0 + tx;
    if (i < T && j < T) {
        double acc = 0.0;
        for (int k = 0; k < rank; ++k) acc = fma(Hs[ty][k], Fs[tx][k], acc);
        if (i == j) acc += reg / (1.0 + reg * w[(size_t)i * L + latent]);
        cov[(size_t)i * T + j] = acc;
    }
}

}   // namespace

extern "C" int vlgp_posterior_cov(vlgp_ctx *ctx, int set_id, int trial, int latent, double reg, double *cov) {
    TrialSet *ts = get_set(ctx, set_id);
    REQUIRE(ts && cov, "posterior_cov: bad arguments");
    REQUIRE(trial >= 0 && trial < ts->n_trials && latent >= 0 && latent < ctx->L, "posterior_cov: trial %d / latent %d out of range",
            trial, latent);
    REQUIRE(reg >= 0.0, "posterior_cov: reg must be >= 0");
    REQUIRE(ctx->rank <= 64, "posterior_cov: rank %d > 64", ctx->rank);
    CK(cudaSetDevice(ctx->device));
    const int T = ts->h_len[trial], rank = ctx->rank, L = ctx->L;
    const PriorFactor &pf = ts->factors[ts->h_fidx[trial]];
    const double *G = pf.d_G + (size_t)latent * T * rank;
    const double *w = ts->d_w + (size_t)ts->h_start[trial] * L;
    double *H = nullptr, *dcov = nullptr;
    CK(vlgp_dalloc(ctx, &H, (size_t)T * rank * sizeof(double)));
    CK(vlgp_dalloc(ctx, &dcov, (size_t)T * T * sizeof(double)));
    CK(cudaMemsetAsync(ctx->d_flags + 2, 0, sizeof(int), ctx->stream));
    const size_t smem = ((size_t)rank * (rank | 1) + 128) * sizeof(double);
    pcov_factor_kernel<<<1, NT, smem, ctx->stream>>>(T, rank, L, latent, reg, G, w, H, ctx->d_flags + 2);
    CKL();
    pcov_outer_kernel<<<dim3((T + 15) / 16, (T + 15) / 16), 256, 0, ctx->stream>>>(T, rank, L, latent, reg, G, w, H, dcov);
    CKL();
    CK(cudaMemcpyAsync(cov, dcov, (size_t)T * T * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(ctx->h_flags + 2, ctx->d_flags + 2, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    CK(vlgp_dfree(ctx, H));
    CK(vlgp_dfree(ctx, dcov));
    if (ctx->h_flags[2]) return vlgp_fail(ctx, VLGP_ERR_ARG, "posterior_cov: I + G'WG is not positive definite (negative weights?)");
    return VLGP_OK;
}
