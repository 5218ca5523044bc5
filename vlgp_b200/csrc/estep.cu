// K3/K4: variational E-step for trials of ANY length (one CTA per trial, per-bin state in HBM/L2, r x r work in SMEM).
//
// Replaces core.infer_single_trial / core.estep (vlgp/core.py:22-126), core.update_w (:419-442) and core.update_v
// (:445-471).  The SMEM-resident kernel specialised for window-length segments lives in estep_seg.cu; this one is the
// general path used for the uncut trials before/after vem and by infer (vlgp/api.py:52-54,66-71).
//
// Algebra (verified against the reference forms to 1e-15, tests/test_reformulation.py): with A = G' diag(w_l) G and
// Minv = (I + A)^-1,
//   mean step      delta = u - G Minv G'(w_l o u),  u = G G'(resid a_l) - mu_l      (reference: u - Gc + G A solve(I+A,c))
//   variance       v_t   = G_t Minv G_t'                                            (reference: rowsum(G o (G - GA + GAM)))
// and the factor built from w after iteration i serves both the variance of iteration i and the mean step of i+1
// (the reference rebuilds and re-solves the same system twice).  Minv comes from an in-place symmetric sweep
// (Gauss-Jordan on an SPD matrix; pivot <= 0  <=>  LAPACK posv would report "not positive definite").
// Only the first ncol columns of G are non-zero (ichol stops early), so every r x r object is ncol x ncol.
#include "common.cuh"
#include "linalg.cuh"

int vlgp_launch_xb(vlgp_ctx *ctx, TrialSet *ts);      // regress.cu

namespace {

struct EstepArgs {
    int n_trials, N, L, rank;
    const int32_t *subset;       // n_trials trial indices to process (vlgp_estep_subset), or null: trials 0..n_trials-1
    const int *len;
    const int64_t *start;
    const int *fidx;
    double *const *Gptr;
    int *const *ncolptr;
    const void *y;
    int ydtype;
    double *mu, *v, *w, *dmu, *ra, *u, *minv;
    const double *a, *b, *noise;
    const double *xb;            // nbin x N offsets einsum(x, b) for general regressors (regress.cu), or null: b[n]
    const uint8_t *poisson;
    int n_iter;
    double dmu_bound;
    int do_mean, do_w, do_v;     // estep: 1,1,method_vb ; update_w: 0,1,0 ; update_v: 0,0,1
    int *flags;                  // flags[0] += number of non-PD systems
};

constexpr int NT = 256;
constexpr int NWARP = NT / 32;

// resid @ a_l (STAGE 1, vlgp/core.py:69-70,81-87) or U @ a_l^2 (STAGE 2, :100-104) for every bin of the trial.
template <int LT, int STAGE>
__device__ __forceinline__ void rate_stage(const EstepArgs &p, int64_t s0, int T, double *out) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int N = p.N;
    for (int t = wid; t < T; t += NWARP) {
        const int64_t bin = s0 + t;
        double mu_t[LT], v_t[LT], acc[LT];
#pragma unroll
        for (int l = 0; l < LT; ++l) {
            mu_t[l] = p.mu[bin * LT + l];
            v_t[l] = p.v[bin * LT + l];
            acc[l] = 0.0;
        }
        for (int n = lane; n < N; n += 32) {
            double al[LT];
            double eta = p.xb ? p.xb[bin * N + n] : p.b[n], h = 0.0;
#pragma unroll
            for (int l = 0; l < LT; ++l) {
                al[l] = p.a[l * N + n];
                eta = fma(mu_t[l], al[l], eta);
                h = fma(v_t[l], al[l] * al[l], h);
            }
            const bool pois = p.poisson[n] != 0;
            double coef;
            if (STAGE == 1) {
                const double yv = load_y(p.y, p.ydtype, bin * N + n);
                coef = pois ? yv - trunc_exp(eta + 0.5 * h) : (yv - eta) / p.noise[n];
#pragma unroll
                for (int l = 0; l < LT; ++l) acc[l] = fma(coef, al[l], acc[l]);
            } else {
                coef = pois ? trunc_exp(eta + 0.5 * h) : 1.0 / p.noise[n];
#pragma unroll
                for (int l = 0; l < LT; ++l) acc[l] = fma(coef, al[l] * al[l], acc[l]);
            }
        }
#pragma unroll
        for (int l = 0; l < LT; ++l) {
            const double s = warp_sum(acc[l]);
            if (lane == 0) out[bin * LT + l] = s;
        }
    }
}

// out[i] = sum_t G[t][i] * x(t), i < nc.  part: SMEM 4 x 64.
template <typename XF>
__device__ __forceinline__ void gt_vec(const double *__restrict__ G, int T, int rank, int nc, XF x, double *part,
                                       double *out) {
    const int i = threadIdx.x & 63, tg = threadIdx.x >> 6;
    double s = 0.0;
    if (i < nc)
        for (int t = tg; t < T; t += NT / 64) s = fma(G[(size_t)t * rank + i], x(t), s);
    part[tg * 64 + i] = s;
    __syncthreads();
    if (threadIdx.x < nc) {
        double r = 0.0;
#pragma unroll
        for (int g = 0; g < NT / 64; ++g) r += part[g * 64 + threadIdx.x];
        out[threadIdx.x] = r;
    }
    __syncthreads();
}

// G[t] . vec for one row by one warp (vec in SMEM, nc <= 64)
__device__ __forceinline__ double row_dot(const double *__restrict__ Grow, const double *vec, int nc, int lane) {
    double s = 0.0;
    if (lane < nc) s = Grow[lane] * vec[lane];
    if (lane + 32 < nc) s = fma(Grow[lane + 32], vec[lane + 32], s);
    return warp_sum(s);
}

// Aw (SMEM, ld) <- -(I + G' diag(w_l) G)^-1 restricted to nc x nc; also written (negated, i.e. +Minv) to minv_out
// (global, ld = rank).  Returns false if the matrix is not positive definite.
__device__ __forceinline__ bool build_minv(const double *__restrict__ G, int T, int rank, int nc, const double *wcol,
                                           int wstride, double *Aw, int ld, double *ck, double *minv_out) {
    const int tid = threadIdx.x;
    // ---- A = I + G' W G, 2 x 2 register tiles over the lower triangle ---------------------------------------------
    const int nb = (nc + 1) >> 1;
    const int ntile = nb * (nb + 1) / 2;
    for (int idx = tid; idx < ntile; idx += NT) {
        int bi, bj;
        tri_decode(idx, bi, bj);
        const int i0 = 2 * bi, j0 = 2 * bj;
        const int i1 = min(i0 + 1, rank - 1), j1 = min(j0 + 1, rank - 1);
        double c00 = 0, c01 = 0, c10 = 0, c11 = 0;
        for (int t = 0; t < T; ++t) {
            const double *g = G + (size_t)t * rank;
            const double wt = wcol[(size_t)t * wstride];
            const double gi0 = g[i0] * wt, gi1 = g[i1] * wt, gj0 = g[j0], gj1 = g[j1];
            c00 = fma(gi0, gj0, c00);
            c01 = fma(gi0, gj1, c01);
            c10 = fma(gi1, gj0, c10);
            c11 = fma(gi1, gj1, c11);
        }
        const bool vi1 = i0 + 1 < nc, vj1 = j0 + 1 < nc;
        Aw[i0 * ld + j0] = c00 + (i0 == j0 ? 1.0 : 0.0);
        Aw[j0 * ld + i0] = c00 + (i0 == j0 ? 1.0 : 0.0);
        if (vj1) { Aw[i0 * ld + j0 + 1] = c01; Aw[(j0 + 1) * ld + i0] = c01; }
        if (vi1) { Aw[(i0 + 1) * ld + j0] = c10; Aw[j0 * ld + i0 + 1] = c10; }
        if (vi1 && vj1) {
            Aw[(i0 + 1) * ld + j0 + 1] = c11 + (i0 == j0 ? 1.0 : 0.0);
            Aw[(j0 + 1) * ld + i0 + 1] = c11 + (i0 == j0 ? 1.0 : 0.0);
        }
    }
    __syncthreads();
    // On the diagonal tiles (bi == bj) c01 and c10 are both sum_t g_i0 w g_i1 up to rounding; the two mirrored writes
    // above store c01 into [i0][i0+1] and [i0+1][i0] first and c10 into the same two cells second, so the matrix
    // stays exactly symmetric.
    // ---- symmetric sweep -> Aw = -(I + A)^-1 -----------------------------------------------------------------------
    const int ty = tid >> 4, tx = tid & 15;
    const bool ok = block_sweep_spd(Aw, ld, nc, ck, nullptr);
    if (ok && minv_out != nullptr) {
        for (int i = ty; i < nc; i += 16)
            for (int j = tx; j < nc; j += 16) minv_out[i * rank + j] = -Aw[i * ld + j];
    }
    __syncthreads();
    return ok;
}

template <int LT>
__global__ void __launch_bounds__(NT) estep_generic_kernel(EstepArgs p) {
    extern __shared__ double sm[];
    const int rank = p.rank;
    const int ld = rank | 1;                 // odd leading dimension: conflict-free row-per-lane access
    double *Aw = sm;                         // rank x ld
    double *ck = Aw + rank * ld;             // 64
    double *pv = ck + 128;                   // 64   (ck: 2 x 64, double-buffered pivot column)
    double *cv = pv + 64;                    // 64
    double *mv = cv + 64;                    // 64
    double *part = mv + 64;                  // 4 x 64
    __shared__ int bad[VLGP_MAX_L];          // factor of latent l unusable (non-PD)

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    double *minv_cta = p.minv + (size_t)blockIdx.x * LT * rank * rank;

    for (int item = blockIdx.x; item < p.n_trials; item += gridDim.x) {
        const int trial = p.subset ? p.subset[item] : item;
        const int T = p.len[trial];
        const int64_t s0 = p.start[trial];
        const double *G = p.Gptr[p.fidx[trial]];
        const int *ncol = p.ncolptr[p.fidx[trial]];
        if (tid < LT) bad[tid] = 0;
        __syncthreads();

        for (int it = 0; it < p.n_iter; ++it) {
            if (p.do_mean) {
                rate_stage<LT, 1>(p, s0, T, p.ra);
                __syncthreads();
                for (int l = 0; l < LT; ++l) {
                    const int nc = ncol[l];
                    const double *Gl = G + (size_t)l * T * rank;
                    double *minv_l = minv_cta + (size_t)l * rank * rank;
                    if (it == 0) {
                        const bool ok = build_minv(Gl, T, rank, nc, p.w + s0 * LT + l, LT, Aw, ld, ck, minv_l);
                        if (!ok && tid == 0) { bad[l] = 1; atomicAdd(p.flags, 1); }
                        __syncthreads();
                    }
                    if (bad[l]) {            // failed solve: delta_mu = 0 (vlgp/core.py:92-94)
                        for (int t = tid; t < T; t += NT) p.dmu[(s0 + t) * LT + l] = 0.0;
                        continue;
                    }
                    const double *ra = p.ra;
                    gt_vec(Gl, T, rank, nc, [&](int t) { return ra[(s0 + t) * LT + l]; }, part, pv);
                    for (int t = wid; t < T; t += NWARP) {
                        const double s = row_dot(Gl + (size_t)t * rank, pv, nc, lane);
                        if (lane == 0) p.u[s0 + t] = s - p.mu[(s0 + t) * LT + l];
                    }
                    __syncthreads();
                    const double *uu = p.u;
                    const double *ww = p.w;
                    gt_vec(Gl, T, rank, nc, [&](int t) { return ww[(s0 + t) * LT + l] * uu[s0 + t]; }, part, cv);
                    if (tid < nc) {
                        double s = 0.0;
                        for (int j = 0; j < nc; ++j) s = fma(minv_l[j * rank + tid], cv[j], s);
                        mv[tid] = s;
                    }
                    __syncthreads();
                    for (int t = wid; t < T; t += NWARP) {
                        const double s = row_dot(Gl + (size_t)t * rank, mv, nc, lane);
                        if (lane == 0) {
                            const double d = clipd(p.u[s0 + t] - s, p.dmu_bound);
                            p.dmu[(s0 + t) * LT + l] = d;
                            p.mu[(s0 + t) * LT + l] += d;
                        }
                    }
                    __syncthreads();
                }
            }
            if (p.do_w) {
                __syncthreads();
                rate_stage<LT, 2>(p, s0, T, p.w);
                __syncthreads();
            }
            const bool need_factor = p.do_v || (p.do_mean && it + 1 < p.n_iter);
            if (need_factor) {
                for (int l = 0; l < LT; ++l) {
                    const int nc = ncol[l];
                    const double *Gl = G + (size_t)l * T * rank;
                    double *minv_l = minv_cta + (size_t)l * rank * rank;
                    const bool ok = build_minv(Gl, T, rank, nc, p.w + s0 * LT + l, LT, Aw, ld, ck, minv_l);
                    if (tid == 0) {
                        bad[l] = ok ? 0 : 1;
                        if (!ok) atomicAdd(p.flags, 1);
                    }
                    if (ok && p.do_v) {
                        // v_t = G_t Minv G_t'   (Aw holds -Minv)
                        for (int t = wid; t < T; t += NWARP) {
                            const double *g = Gl + (size_t)t * rank;
                            double s0a = 0.0, s1a = 0.0;
                            const int i0 = lane, i1 = lane + 32;
                            for (int j = 0; j < nc; ++j) {
                                const double gj = g[j];
                                if (i0 < nc) s0a = fma(Aw[i0 * ld + j], gj, s0a);
                                if (i1 < nc) s1a = fma(Aw[i1 * ld + j], gj, s1a);
                            }
                            double s = 0.0;
                            if (i0 < nc) s = g[i0] * s0a;
                            if (i1 < nc) s = fma(g[i1], s1a, s);
                            s = warp_sum(s);
                            if (lane == 0) p.v[(s0 + t) * LT + l] = -s;
                        }
                    }
                    __syncthreads();
                }
            }
        }
        __syncthreads();
    }
}

template <int LT>
int launch_estep_t(vlgp_ctx *ctx, TrialSet *ts, EstepArgs &p) {
    const int rank = p.rank;
    const size_t smem = ((size_t)rank * (rank | 1) + 64 * 9) * sizeof(double);
    if (smem > 48 * 1024)
        CK(cudaFuncSetAttribute(estep_generic_kernel<LT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 1;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, estep_generic_kernel<LT>, NT, smem));
    if (per_sm < 1) per_sm = 1;
    int grid = per_sm * ctx->prop.multiProcessorCount;
    if (grid > p.n_trials) grid = p.n_trials;
    if (grid < 1) grid = 1;
    if (ts->minv_grid < grid) {
        if (ts->d_minv) CK(vlgp_dfree(ctx, ts->d_minv));
        ts->d_minv = nullptr;
        CK(vlgp_dalloc(ctx, &ts->d_minv, (size_t)grid * LT * rank * rank * sizeof(double)));
        ts->minv_grid = grid;
    }
    p.minv = ts->d_minv;
    estep_generic_kernel<LT><<<grid, NT, smem, ctx->stream>>>(p);
    CKL();
    return VLGP_OK;
}

}   // namespace

#define DISPATCH_L(L, CALL)                                                            \
    switch (L) {                                                                       \
        case 1: { constexpr int LT = 1; CALL; } break;                                 \
        case 2: { constexpr int LT = 2; CALL; } break;                                 \
        case 3: { constexpr int LT = 3; CALL; } break;                                 \
        case 4: { constexpr int LT = 4; CALL; } break;                                 \
        case 5: { constexpr int LT = 5; CALL; } break;                                 \
        case 6: { constexpr int LT = 6; CALL; } break;                                 \
        case 7: { constexpr int LT = 7; CALL; } break;                                 \
        case 8: { constexpr int LT = 8; CALL; } break;                                 \
        case 9: { constexpr int LT = 9; CALL; } break;                                 \
        case 10: { constexpr int LT = 10; CALL; } break;                               \
        case 11: { constexpr int LT = 11; CALL; } break;                               \
        case 12: { constexpr int LT = 12; CALL; } break;                               \
        default: return vlgp_fail(ctx, VLGP_ERR_UNSUPPORTED, "n_latents %d > 12", L);  \
    }

// mode: 0 = estep, 1 = update_w, 2 = update_v
int vlgp_launch_estep_generic(vlgp_ctx *ctx, TrialSet *ts, int mode, int n_iter, double dmu_bound, int method_vb,
                              const int32_t *d_subset, int n_subset) {
    EstepArgs p{};
    p.n_trials = d_subset ? n_subset : ts->n_trials;
    p.subset = d_subset;
    p.N = ctx->N;
    p.L = ctx->L;
    p.rank = ctx->rank;
    p.len = ts->d_len;
    p.start = ts->d_start;
    p.fidx = ts->d_fidx;
    p.Gptr = ts->d_Gptr;
    p.ncolptr = ts->d_ncolptr;
    p.y = ts->d_y;
    p.ydtype = ts->ydtype;
    p.mu = ts->d_mu; p.v = ts->d_v; p.w = ts->d_w; p.dmu = ts->d_dmu; p.ra = ts->d_ra; p.u = ts->d_u;
    p.a = ctx->d_a; p.b = ctx->d_b; p.noise = ctx->d_noise; p.poisson = ctx->d_poisson;
    p.n_iter = mode == 0 ? n_iter : 1;
    p.dmu_bound = dmu_bound;
    p.do_mean = mode == 0;
    p.do_w = mode == 0 || mode == 1;
    p.do_v = (mode == 0 && method_vb) || mode == 2;
    p.flags = ctx->d_flags;
    int rc = VLGP_OK;
    if (ts->d_x && p.do_w) {                  // general regressors: refresh einsum(x, b) with the current b
        rc = vlgp_launch_xb(ctx, ts);
        if (rc) return rc;
        p.xb = ts->d_xb;
    }
    DISPATCH_L(ctx->L, rc = launch_estep_t<LT>(ctx, ts, p));
    return rc;
}
