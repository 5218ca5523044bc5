// K7 for WIDE windows (64 < W <= VLGP_MAX_W_H): the H-step objective of hstep.cu with every W x W matrix held in shared
// memory instead of registers.  The reference accepts any window (vlgp/gp.py:65-123, fit(..., window=100) is a valid
// call); the tuned kernels of hstep.cu / hstep_dmma.cu hold a window of at most 64 bins.  Same algebra as there:
//     B_i = I + d K d (d = sqrt(w_i)),  tr(K^-1 S_i) = tr(B_i^-1),  K^-1 S_i K^-1 - K^-1 = -d B_i^-1 d,
//     ll = -1/2 tr(K^-1 M) - 1/2 sum_i tr(B_i^-1) - S sum log diag chol K,   dll = 1/2 [(K^-1 M K^-1):dK - sum_i (d B_i^-1 d):dK]
// One CTA per matrix, symmetric sweep (Gauss-Jordan without pivoting: pivot <= 0 is LAPACK's "not positive definite")
// over the full W x W array in shared memory, 2-D thread layout (16 x 16) so that no step divides.  Built for
// correctness on an uncommon path, not for speed: a W = 100 evaluation costs ~W^3 FMAs per segment on one CTA.
#include "common.cuh"
#include "linalg.cuh"

namespace {

constexpr int NT = 256;

// A (n x n, leading dimension ld, both triangles) <- -A^-1 in place.  col: n doubles of SMEM; piv (optional): n pivots.
// Returns false (uniformly) on a non-positive pivot.  All NT threads must call.
__device__ bool smem_sweep(double *A, int ld, int n, double *col, double *piv) {
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    for (int k = 0; k < n; ++k) {
        for (int i = tid; i < n; i += NT) col[i] = A[i * ld + k];
        __syncthreads();
        const double d = col[k];
        if (!(d > 0.0)) return false;
        if (piv && tid == 0) piv[k] = d;
        const double pinv = 1.0 / d;
        for (int i = ty; i < n; i += 16) {
            const double ci = col[i] * pinv;
            for (int j = tx; j < n; j += 16) {
                double v;
                if (i == k) v = (j == k) ? -pinv : col[j] * pinv;
                else if (j == k) v = ci;
                else v = fma(-ci, col[j], A[i * ld + j]);
                A[i * ld + j] = v;
            }
        }
        __syncthreads();
    }
    return true;
}

// second moments of mu over segments: part[chunk][l][a][b] = sum_{s in chunk} mu[s][a][l] mu[s][b][l]
__global__ void __launch_bounds__(NT) hstep_moment_wide_kernel(int nseg, int W, int L, const double *__restrict__ mu,
                                                               double *__restrict__ part) {
    const int idx = blockIdx.x * NT + threadIdx.x, chunk = blockIdx.y, l = blockIdx.z;
    if (idx >= W * W) return;
    const int a = idx / W, b = idx - a * W;
    const int per = (nseg + gridDim.y - 1) / gridDim.y;
    const int s0 = chunk * per, s1 = min(nseg, s0 + per);
    double x = 0.0;
    for (int s = s0; s < s1; ++s) x = fma(mu[((size_t)s * W + a) * L + l], mu[((size_t)s * W + b) * L + l], x);
    part[((size_t)chunk * L + l) * W * W + idx] = x;
}

// one CTA per evaluation: K, dK (written to Kall for the per-segment kernel), -K^-1 by the sweep, then
// out[e][0] = tr(K^-1 M), [1] = sum log diag chol(K), [2] = (K^-1 M K^-1):dK, [5] = 1 if K is not PD
__global__ void __launch_bounds__(NT) hstep_global_wide_kernel(HEvalBatch eb, int W, double dt,
                                                               const double *__restrict__ Mall, double *__restrict__ Kall,
                                                               double *__restrict__ outall) {
    extern __shared__ double sm[];
    const int e = blockIdx.x, ld = W + 1;
    const double sigmasq = eb.sigmasq[e], omega = eb.omega[e], eps = eb.eps[e];
    const double *M = Mall + (size_t)eb.latent[e] * W * W;
    double *Kout = Kall + (size_t)e * 2 * W * W, *dKout = Kout + (size_t)W * W;
    double *out = outall + e * 8;
    double *A = sm;                    // W x ld
    double *col = A + (size_t)W * ld;  // W
    double *piv = col + W;             // W
    double *red = piv + W;             // 32
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    for (int i = ty; i < W; i += 16)
        for (int j = tx; j < W; j += 16) {
            const double dx = (double)(i - j) * dt, d2 = dx * dx;
            const double ks = sigmasq * exp(-omega * d2);
            const double k = ks + (i == j ? eps : 0.0);
            A[i * ld + j] = k;
            Kout[i * W + j] = k;
            dKout[i * W + j] = -ks * d2 * omega;
        }
    __syncthreads();
    const bool ok = smem_sweep(A, ld, W, col, piv);
    if (!ok) {
        if (tid == 0) {
            out[0] = out[1] = out[2] = 0.0;
            out[5] = 1.0;
        }
        return;
    }
    double lg = 0.0;
    for (int i = tid; i < W; i += NT) lg += log(piv[i]);
    const double logdet = block_sum(lg, red);
    // T' = (-K^-1) M and U' = dK (-K^-1): T' o U' = T o U, tr(T) = -tr(T')  (the dK just written is read back through L2)
    __threadfence_block();
    double t1 = 0.0, gr = 0.0;
    for (int i = ty; i < W; i += 16)
        for (int j = tx; j < W; j += 16) {
            double t = 0.0, u = 0.0;
            for (int k = 0; k < W; ++k) {
                t = fma(A[i * ld + k], M[(size_t)k * W + j], t);
                u = fma(dKout[(size_t)i * W + k], A[k * ld + j], u);
            }
            gr = fma(t, u, gr);
            if (i == j) t1 -= t;
        }
    t1 = block_sum(t1, red);
    gr = block_sum(gr, red);
    if (tid == 0) {
        out[0] = t1;
        out[1] = 0.5 * logdet;
        out[2] = gr;
        out[5] = 0.0;
    }
}

// one CTA per segment at a time (blockIdx.y = evaluation): part[e][seg] = tr(B^-1), part[e][nseg + seg] = (d B^-1 d):dK
__global__ void __launch_bounds__(NT) hstep_segment_wide_kernel(HEvalBatch eb, int nseg, int W, int L,
                                                                const double *__restrict__ w,
                                                                const double *__restrict__ Kall,
                                                                double *__restrict__ partall) {
    extern __shared__ double sm[];
    const int e = blockIdx.y, l = eb.latent[e], ld = W + 1;
    const double *K = Kall + (size_t)e * 2 * W * W, *dK = K + (size_t)W * W;
    double *part = partall + (size_t)e * 2 * nseg;
    double *A = sm;                    // W x ld
    double *col = A + (size_t)W * ld;  // W
    double *dv = col + W;              // W : sqrt(w)
    double *red = dv + W;              // 32
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    for (int seg = blockIdx.x; seg < nseg; seg += gridDim.x) {
        __syncthreads();
        for (int t = tid; t < W; t += NT) dv[t] = sqrt(fmax(w[((size_t)seg * W + t) * L + l], 0.0));
        __syncthreads();
        for (int i = ty; i < W; i += 16)
            for (int j = tx; j < W; j += 16)
                A[i * ld + j] = dv[i] * K[(size_t)i * W + j] * dv[j] + (i == j ? 1.0 : 0.0);
        __syncthreads();
        const bool ok = smem_sweep(A, ld, W, col, nullptr);
        double tr = 0.0, pd = 0.0;
        if (ok) {
            for (int i = ty; i < W; i += 16)
                for (int j = tx; j < W; j += 16) {
                    const double binv = -A[i * ld + j];
                    if (i == j) tr += binv;
                    pd = fma(binv * dv[i] * dv[j], dK[(size_t)i * W + j], pd);
                }
        }
        tr = block_sum(tr, red);
        pd = block_sum(pd, red);
        if (tid == 0) {
            const double nan = __longlong_as_double(0x7ff8000000000000LL);
            part[seg] = ok ? tr : nan;
            part[nseg + seg] = ok ? pd : nan;
        }
    }
}

size_t wide_smem(int W) { return ((size_t)W * (W + 1) + 2 * (size_t)W + 32) * sizeof(double); }

}   // namespace

// Moments of the H-step for a wide window: Mpart[chunk][l] on ctx->stream (the caller reduces the chunks).
int vlgp_launch_hstep_moments_wide(vlgp_ctx *ctx, TrialSet *ts, int chunks, double *part) {
    const int W = ts->max_len, L = ctx->L, S = ts->n_trials;
    hstep_moment_wide_kernel<<<dim3((W * W + NT - 1) / NT, chunks, L), NT, 0, ctx->stream>>>(S, W, L, ts->d_mu, part);
    CKL();
    return VLGP_OK;
}

// K^-1 terms and per-segment terms of eb.n evaluations on ctx->stream (the caller adds the final reduction).
int vlgp_launch_hstep_wide(vlgp_ctx *ctx, TrialSet *ts, const HEvalBatch &eb) {
    const int W = ts->max_len, S = ts->n_trials;
    const size_t smem = wide_smem(W);
    static bool attr_done = false;
    if (!attr_done) {
        CK(cudaFuncSetAttribute(hstep_global_wide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        CK(cudaFuncSetAttribute(hstep_segment_wide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_done = true;
    }
    hstep_global_wide_kernel<<<eb.n, NT, smem, ctx->stream>>>(eb, W, ctx->dt, ts->d_M, ts->d_K, ts->d_hout);
    CKL();
    int per_sm = 1;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, hstep_segment_wide_kernel, NT, smem));
    int grid = (per_sm < 1 ? 1 : per_sm) * ctx->prop.multiProcessorCount;
    if (grid > S) grid = S;
    hstep_segment_wide_kernel<<<dim3(grid, eb.n), NT, smem, ctx->stream>>>(eb, S, W, ctx->L, ts->d_w, ts->d_K, ts->d_hpart);
    CKL();
    return VLGP_OK;
}
