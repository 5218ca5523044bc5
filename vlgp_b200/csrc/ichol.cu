// K1: batched pivoted incomplete Cholesky of the squared-exponential kernel, one CTA per latent.
//
// Replaces math.ichol_gauss (vlgp/math.py:76-126) as called by gp.make_cholesky (vlgp/gp.py:150-162).
// Parity notes (hard part 1): the pivot is the FIRST arg-max of the residual diagonal in the
// current permuted order, so every row carries its position and ties are broken on the smallest position; the residual
// diagonal is recomputed from scratch after every column as 1 - sum(F^2) with NumPy's pairwise-summation order and
// without FMA contraction, so the values that decide exact ties are bit-identical to NumPy's whenever the factor entries
// are.  Rows are never moved: the factor is built in natural row order (the reference permutes and un-permutes).
#include "common.cuh"

namespace {

struct PivotCand {
    double val;
    int pos;
    int row;
};

__device__ __forceinline__ PivotCand better(PivotCand a, PivotCand b) {
    // larger residual wins; ties -> smaller position (np.argmax returns the first maximum)
    if (b.row < 0) return a;
    if (a.row < 0) return b;
    if (b.val > a.val || (b.val == a.val && b.pos < a.pos)) return b;
    return a;
}

// Work layout: Fw is column-major (rank x n) so that consecutive threads (rows) read consecutive addresses.
__global__ void __launch_bounds__(1024)
ichol_gauss_kernel(int n, int rank, double dt, double tol, const double *__restrict__ omega,
                   const double *__restrict__ sigma, double *__restrict__ Fwork, double *__restrict__ Gout,
                   int *__restrict__ piv_out, int *__restrict__ ncol_out) {
    extern __shared__ unsigned char smem_raw[];
    double *resid = (double *)smem_raw;              // n   residual diagonal, by natural row
    int *posn = (int *)(resid + n);                  // n   position of each natural row in the permuted order
    int *perm = posn + n;                            // n   row at each position
    __shared__ double prow[VLGP_MAX_RANK];           // entries of the pivot row
    __shared__ double red_d[32];
    __shared__ PivotCand red_c[32];
    __shared__ double s_sum;
    __shared__ PivotCand s_piv;

    const int l = blockIdx.x;
    const int tid = threadIdx.x, nt = blockDim.x;
    const int lane = tid & 31, wid = tid >> 5, nw = nt >> 5;
    const double om = omega[l];
    double *F = Fwork + (size_t)l * rank * n;

    for (int j = tid; j < n; j += nt) {
        resid[j] = 1.0;
        posn[j] = j;
        perm[j] = j;
    }
    for (size_t i = tid; i < (size_t)rank * n; i += nt) F[i] = 0.0;
    for (int i = tid; i < rank; i += nt) piv_out[l * rank + i] = -1;
    __syncthreads();

    int k = 0;
    for (; k < rank; ++k) {
        // ---- residual trace over the active rows (positions >= k) and first-arg-max pivot --------------------------
        double part = 0.0;
        PivotCand cand{0.0, 0, -1};
        for (int j = tid; j < n; j += nt) {
            if (posn[j] >= k) {
                part += resid[j];
                cand = better(cand, PivotCand{resid[j], posn[j], j});
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            part += __shfl_xor_sync(0xffffffffu, part, o);
            PivotCand oth;
            oth.val = __shfl_xor_sync(0xffffffffu, cand.val, o);
            oth.pos = __shfl_xor_sync(0xffffffffu, cand.pos, o);
            oth.row = __shfl_xor_sync(0xffffffffu, cand.row, o);
            cand = better(cand, oth);
        }
        if (lane == 0) {
            red_d[wid] = part;
            red_c[wid] = cand;
        }
        __syncthreads();
        if (wid == 0) {
            part = lane < nw ? red_d[lane] : 0.0;
            cand = lane < nw ? red_c[lane] : PivotCand{0.0, 0, -1};
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                part += __shfl_xor_sync(0xffffffffu, part, o);
                PivotCand oth;
                oth.val = __shfl_xor_sync(0xffffffffu, cand.val, o);
                oth.pos = __shfl_xor_sync(0xffffffffu, cand.pos, o);
                oth.row = __shfl_xor_sync(0xffffffffu, cand.row, o);
                cand = better(cand, oth);
            }
            if (lane == 0) {
                s_sum = part;
                if (k == 0) cand = PivotCand{resid[perm[0]], 0, perm[0]};   // the reference takes position 0 first
                s_piv = cand;
            }
        }
        __syncthreads();
        if (!(s_sum > tol * (double)n)) break;      // while ... np.sum(d[i:]) > tol * n
        const int prow_id = s_piv.row;
        const int ppos = s_piv.pos;
        const double pivot = sqrt(s_piv.val);
        // ---- swap positions k <-> ppos, publish the pivot row ------------------------------------------------------
        if (tid == 0) {
            const int q = perm[k];
            perm[k] = prow_id;
            perm[ppos] = q;
            posn[q] = ppos;
            posn[prow_id] = k;
            piv_out[l * rank + k] = prow_id;
            F[(size_t)k * n + prow_id] = pivot;
        }
        for (int m = tid; m < k; m += nt) prow[m] = F[(size_t)m * n + prow_id];
        __syncthreads();
        // ---- next column and the residual diagonal of the still-active rows ----------------------------------------
        const double xp = (double)prow_id * dt;
        const int M = k + 1;
        for (int j = tid; j < n; j += nt) {
            if (posn[j] <= k) continue;
            const double dx = (double)j * dt - xp;
            const double col = exp(__dmul_rn(-om, __dmul_rn(dx, dx)));
            double dot = 0.0;
            double r8[8];
            double res = 0.0;
            const double *Fj = F + j;
            // pass over the k existing entries: dot product with the pivot row + NumPy-ordered sum of squares
            if (M < 8) {
                for (int m = 0; m < k; ++m) {
                    const double f = Fj[(size_t)m * n];
                    dot = fma(f, prow[m], dot);
                    res = __dadd_rn(res, __dmul_rn(f, f));
                }
                const double val = (col - dot) / pivot;
                res = __dadd_rn(res, __dmul_rn(val, val));
                F[(size_t)k * n + j] = val;
            } else {
                const int Mblk = M - (M % 8);        // entries [0, Mblk) go through the 8 strided accumulators
                // entries 0..k-1 are known; entry k (the new one) needs the finished dot product first
                for (int m = 0; m < k; ++m) dot = fma(Fj[(size_t)m * n], prow[m], dot);
                const double val = (col - dot) / pivot;
                F[(size_t)k * n + j] = val;
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const double f = Fj[(size_t)q * n];   // M >= 8 -> indices 0..7 exist (index 7 may be k)
                    const double g = (q == k) ? val : f;
                    r8[q] = __dmul_rn(g, g);
                }
                for (int i = 8; i < Mblk; i += 8) {
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const int m = i + q;
                        const double f = (m == k) ? val : Fj[(size_t)m * n];
                        r8[q] = __dadd_rn(r8[q], __dmul_rn(f, f));
                    }
                }
                res = __dadd_rn(__dadd_rn(__dadd_rn(r8[0], r8[1]), __dadd_rn(r8[2], r8[3])),
                                __dadd_rn(__dadd_rn(r8[4], r8[5]), __dadd_rn(r8[6], r8[7])));
                for (int m = Mblk; m < M; ++m) {
                    const double f = (m == k) ? val : Fj[(size_t)m * n];
                    res = __dadd_rn(res, __dmul_rn(f, f));
                }
            }
            resid[j] = __dsub_rn(1.0, res);
        }
        __syncthreads();
    }
    // ---- emit G = sigma * F in natural row order (row-major n x rank) ------------------------------------------------
    const double sg = sigma[l];
    double *G = Gout + (size_t)l * n * rank;
    for (size_t i = tid; i < (size_t)n * rank; i += nt) {
        const int j = (int)(i / rank), m = (int)(i % rank);
        G[i] = m < k ? F[(size_t)m * n + j] * sg : 0.0;
    }
    if (tid == 0) ncol_out[l] = k;
}

}   // namespace

// Launch for one unique length; d_omega/d_sigma are device arrays of L doubles.
int vlgp_launch_ichol(vlgp_ctx *ctx, PriorFactor &pf, const double *d_omega, const double *d_sigma, double *d_work) {
    const int n = pf.length;
    size_t smem = (size_t)n * (sizeof(double) + 2 * sizeof(int));
    if (smem > 200 * 1024) return vlgp_fail(ctx, VLGP_ERR_UNSUPPORTED, "trial length %d too long for ichol kernel", n);
    int nt = ((n + 31) / 32) * 32;
    if (nt > 1024) nt = 1024;
    if (nt < 64) nt = 64;
    if (smem > 48 * 1024)
        CK(cudaFuncSetAttribute(ichol_gauss_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ProfScope ps(ctx, 3);
    ichol_gauss_kernel<<<ctx->L, nt, smem, ctx->stream>>>(n, ctx->rank, ctx->dt, 1e-6, d_omega, d_sigma, d_work,
                                                          pf.d_G, pf.d_piv, pf.d_ncol);
    CKL();
    return VLGP_OK;
}
