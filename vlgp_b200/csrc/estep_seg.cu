// K2: SMEM-resident E-step for window-length segments (W <= 64 bins): all Eniter Newton iterations of a segment run
// inside one CTA without touching HBM between the first load and the final store of (mu, v, w, dmu).
//
// Replaces core.infer_single_trial (vlgp/core.py:22-120) for the segments vem works on (vlgp/api.py:56,
// vlgp/util.py:457-499).  Same algebra as estep.cu (mean step through Minv = (I + G'WG)^-1, variance as the quadratic
// form G_t Minv G_t', one factorisation per latent per iteration shared by the variance and the next mean step).
//
// Mapping (256 threads, persistent CTAs looping over segments):
//   * rate passes (exp link, the residual and weight contractions over neurons) use ALL threads: thread = (bin,
//     neuron-chunk); the loading a, a^2, bias and the uint8 count tile of the segment are staged in shared memory;
//   * the per-latent r x r work (Gram, symmetric sweep, variance, mean step) runs in ONE WARP PER LATENT with only
//     __syncwarp between its stages, so the L latents proceed concurrently and an iteration needs six block barriers;
//   * only the nc leading non-zero columns of each latent's prior factor are kept (compact copy in shared memory,
//     loaded once per CTA): nc = 6..29 for the reference's omega bounds at W = 50.
#include "common.cuh"
#include "linalg.cuh"

namespace {

constexpr int NT = 256;
constexpr int NWARP = NT / 32;

struct SegArgs {
    int n_seg, W, N, rank;
    const double *G;              // L x W x rank
    int nc[VLGP_MAX_L];           // leading non-zero columns per latent
    int goff[VLGP_MAX_L];         // offsets (doubles) of the compact factor / Minv of latent l inside their regions
    int moff[VLGP_MAX_L];
    int g_total, m_total;         // region sizes (doubles)
    const void *y;
    int ydtype;
    double *mu, *v, *w, *dmu;
    const double *a, *b, *noise;
    const uint8_t *poisson;
    int n_iter;
    double dmu_bound;
    int method_vb;
    int *flags;
    int tpb, chunk;               // threads per bin and neurons per thread in the rate passes
};

__device__ __forceinline__ int ldodd(int n) { return n | 1; }

template <int LT>
struct Smem {
    double *a, *a2, *b, *inv_noise, *Gs, *Mi, *mu, *v, *w, *ra, *dmu, *part, *vec;
    uint8_t *pois, *ys;
    __device__ Smem(unsigned char *base, const SegArgs &p) {
        double *d = (double *)base;
        const int N = p.N, W = p.W;
        a = d; d += 2 * LT * N;                  // interleaved (a, a^2) pairs: one 128-bit load per (latent, neuron)
        a2 = a;
        b = d; d += N;
        inv_noise = d; d += N;
        Gs = d; d += p.g_total;
        Mi = d; d += p.m_total;
        mu = d; d += W * LT;
        v = d; d += W * LT;
        w = d; d += W * LT;
        ra = d; d += W * LT;
        dmu = d; d += W * LT;
        part = d;                                 // rate passes: tpb x W x LT partial sums ...
        vec = d;                                  // ... aliased with the per-warp vectors of the latent phases
        d += max(p.tpb * W * LT, NWARP * 4 * 64);
        pois = (uint8_t *)d;
        ys = pois + ((N + 15) / 16) * 16;
    }
};

__host__ __device__ inline size_t seg_smem_bytes(int LT, int N, int W, int g_total, int m_total, int tpb, bool y_u8) {
    const size_t un = (size_t)tpb * W * LT > (size_t)NWARP * 4 * 64 ? (size_t)tpb * W * LT : (size_t)NWARP * 4 * 64;
    size_t d = (size_t)2 * LT * N + 2 * N + g_total + m_total + (size_t)5 * W * LT + un;
    size_t bytes = d * sizeof(double) + ((N + 15) / 16) * 16;
    if (y_u8) bytes += ((size_t)W * N + 15) / 16 * 16;
    return bytes;
}

// One rate pass over the segment.  STAGE 1: part <- partial sums of resid * a_l ; STAGE 2: of U * a_l^2.
template <int LT, int STAGE>
__device__ __forceinline__ void rate_pass(const SegArgs &p, const Smem<LT> &s, int64_t bin0) {
    const int tid = threadIdx.x;
    const int t = tid / p.tpb, k = tid - t * p.tpb;
    if (t < p.W) {
        const int N = p.N;
        double mu_t[LT], v_t[LT], acc[LT];
#pragma unroll
        for (int l = 0; l < LT; ++l) {
            mu_t[l] = s.mu[t * LT + l];
            v_t[l] = s.v[t * LT + l];
            acc[l] = 0.0;
        }
        const int n0 = k * p.chunk, n1 = min(N, n0 + p.chunk);
        const double2 *aa = (const double2 *)s.a;
#pragma unroll 2
        for (int n = n0; n < n1; ++n) {
            double al[LT], eta = s.b[n], h = 0.0;
#pragma unroll
            for (int l = 0; l < LT; ++l) {
                const double2 p2 = aa[l * N + n];         // (a, a^2)
                eta = fma(mu_t[l], p2.x, eta);
                h = fma(v_t[l], p2.y, h);
                al[l] = (STAGE == 1) ? p2.x : p2.y;
            }
            const bool pois = s.pois[n] != 0;
            double coef;
            if (STAGE == 1) {
                const double yv = p.ydtype == VLGP_Y_U8 ? (double)s.ys[t * N + n]
                                                       : ((const double *)p.y)[(bin0 + t) * N + n];
                coef = pois ? yv - trunc_exp(eta + 0.5 * h) : (yv - eta) * s.inv_noise[n];
            } else {
                coef = pois ? trunc_exp(eta + 0.5 * h) : s.inv_noise[n];
            }
#pragma unroll
            for (int l = 0; l < LT; ++l) acc[l] = fma(coef, al[l], acc[l]);
        }
#pragma unroll
        for (int l = 0; l < LT; ++l) s.part[(k * p.W + t) * LT + l] = acc[l];
    }
    __syncthreads();
    double *out = (STAGE == 1) ? s.ra : s.w;
    for (int idx = tid; idx < p.W * LT; idx += NT) {
        double r = 0.0;
        for (int kk = 0; kk < p.tpb; ++kk) r += s.part[kk * p.W * LT + idx];
        out[idx] = r;
    }
    __syncthreads();
}

// ---- warp-level routines of one latent (lane owns columns lane, lane+32 and bins lane, lane+32) ---------------------
// M (ld odd) <- -(I + G' diag(w_l) G)^-1.  Returns false (warp-uniform) if not positive definite.
template <int LT>
__device__ __forceinline__ bool warp_build_minv(const double *Gl, int ldg, int nc, int W, const double *wl, double *M,
                                                int ldm, double *colk) {
    const int lane = threadIdx.x & 31;
    // Gram matrix: the nc (nc + 1) / 2 lower-triangle entries are spread over the 32 lanes, two per lane per pass
    const int npair = nc * (nc + 1) / 2;
    for (int e0 = lane; e0 < npair; e0 += 64) {
        const bool has1 = e0 + 32 < npair;
        int i0, j0, i1, j1;
        tri_decode(e0, i0, j0);
        tri_decode(has1 ? e0 + 32 : e0, i1, j1);
        double c0 = 0.0, c1 = 0.0;
        for (int t = 0; t < W; ++t) {
            const double *g = Gl + t * ldg;
            const double wt = wl[t * LT];
            c0 = fma(g[i0] * wt, g[j0], c0);
            c1 = fma(g[i1] * wt, g[j1], c1);
        }
        c0 += (i0 == j0) ? 1.0 : 0.0;
        M[i0 * ldm + j0] = c0;
        M[j0 * ldm + i0] = c0;
        if (has1) {
            c1 += (i1 == j1) ? 1.0 : 0.0;
            M[i1 * ldm + j1] = c1;
            M[j1 * ldm + i1] = c1;
        }
    }
    __syncwarp();
    // symmetric sweep, one pivot per step; the nc x nc entries are spread over the lanes
    const float inv_nc = 1.0f / (float)nc;
    const int nn = nc * nc;
    for (int k = 0; k < nc; ++k) {
        for (int i = lane; i < nc; i += 32) colk[i] = M[i * ldm + k];
        __syncwarp();
        const double d = colk[k];
        if (!(d > 0.0)) return false;
        const double pinv = 1.0 / d;
        for (int e = lane; e < nn; e += 32) {
            const int i = (int)(((float)e + 0.5f) * inv_nc);      // exact floor(e / nc) for e < 4096, nc <= 64
            const int j = e - i * nc;
            const double ci = colk[i], cj = colk[j];
            double val;
            if (i == k) val = (j == k) ? -pinv : cj * pinv;
            else if (j == k) val = ci * pinv;
            else val = fma(-ci * pinv, cj, M[i * ldm + j]);
            M[i * ldm + j] = val;
        }
        __syncwarp();
    }
    return true;
}

template <int LT>
__device__ __forceinline__ void warp_variance(const double *Gl, int ldg, int nc, int W, const double *M, int ldm,
                                              double *vl) {
    const int lane = threadIdx.x & 31;
    for (int t = lane; t < W; t += 32) {
        const double *g = Gl + t * ldg;
        double s = 0.0;
        for (int i = 0; i < nc; ++i) {
            double inner = 0.0;
            for (int j = 0; j < nc; ++j) inner = fma(M[i * ldm + j], g[j], inner);
            s = fma(g[i], inner, s);
        }
        vl[t * LT] = -s;          // M holds -Minv
    }
}

// Newton step of the posterior mean of one latent (vlgp/core.py:81-97); vec: 4 x 64 doubles of per-warp SMEM.
template <int LT>
__device__ __forceinline__ void warp_mean_step(const double *Gl, int ldg, int nc, int W, const double *M, int ldm,
                                               const double *ral, const double *wl, double *mul, double *dmul,
                                               double *vec, double bound) {
    const int lane = threadIdx.x & 31;
    double *pv = vec + 64, *cv = vec + 128, *uv = vec + 192;
    for (int j = lane; j < nc; j += 32) {                       // p = G' (resid a_l)
        double s = 0.0;
        for (int t = 0; t < W; ++t) s = fma(Gl[t * ldg + j], ral[t * LT], s);
        pv[j] = s;
    }
    __syncwarp();
    for (int t = lane; t < W; t += 32) {                        // u = G p - mu_l
        double s = 0.0;
        for (int j = 0; j < nc; ++j) s = fma(Gl[t * ldg + j], pv[j], s);
        uv[t] = s - mul[t * LT];
    }
    __syncwarp();
    for (int j = lane; j < nc; j += 32) {                       // c = G' (w_l o u)
        double s = 0.0;
        for (int t = 0; t < W; ++t) s = fma(Gl[t * ldg + j], wl[t * LT] * uv[t], s);
        cv[j] = s;
    }
    __syncwarp();
    for (int i = lane; i < nc; i += 32) {                       // m = Minv c   (M = -Minv, symmetric)
        double s = 0.0;
        for (int j = 0; j < nc; ++j) s = fma(M[j * ldm + i], cv[j], s);
        pv[i] = -s;
    }
    __syncwarp();
    for (int t = lane; t < W; t += 32) {                        // delta = clip(u - G m)
        double s = 0.0;
        for (int j = 0; j < nc; ++j) s = fma(Gl[t * ldg + j], pv[j], s);
        const double d = clipd(uv[t] - s, bound);
        dmul[t * LT] = d;
        mul[t * LT] += d;
    }
    __syncwarp();
}

template <int LT>
__global__ void __launch_bounds__(NT) estep_seg_kernel(SegArgs p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Smem<LT> s(smem_raw, p);
    __shared__ int bad[VLGP_MAX_L];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int W = p.W, N = p.N;

    // ---- once per CTA: parameters and the compact prior factors ------------------------------------------------------
    for (int i = tid; i < LT * N; i += NT) {
        const double x = p.a[i];
        s.a[2 * i] = x;
        s.a[2 * i + 1] = x * x;
    }
    for (int n = tid; n < N; n += NT) {
        s.b[n] = p.b[n];
        s.inv_noise[n] = 1.0 / p.noise[n];
        s.pois[n] = p.poisson[n];
    }
    for (int l = 0; l < LT; ++l) {
        const int nc = p.nc[l], ldg = ldodd(nc);
        const double *Gsrc = p.G + (size_t)l * W * p.rank;
        double *Gd = s.Gs + p.goff[l];
        for (int i = tid; i < W * nc; i += NT) {
            const int t = i / nc, c = i - t * nc;
            Gd[t * ldg + c] = Gsrc[(size_t)t * p.rank + c];
        }
    }
    __syncthreads();

    for (int seg = blockIdx.x; seg < p.n_seg; seg += gridDim.x) {
        const int64_t bin0 = (int64_t)seg * W;
        for (int i = tid; i < W * LT; i += NT) {
            s.mu[i] = p.mu[bin0 * LT + i];
            s.v[i] = p.v[bin0 * LT + i];
            s.w[i] = p.w[bin0 * LT + i];
            s.dmu[i] = 0.0;
        }
        if (p.ydtype == VLGP_Y_U8) {
            const uint8_t *ysrc = (const uint8_t *)p.y + bin0 * N;
            for (int i = tid; i < W * N; i += NT) s.ys[i] = ysrc[i];
        }
        if (tid < LT) bad[tid] = 0;
        __syncthreads();

        for (int it = 0; it < p.n_iter; ++it) {
            rate_pass<LT, 1>(p, s, bin0);
            for (int l = wid; l < LT; l += NWARP) {
                const int nc = p.nc[l], ldg = ldodd(nc), ldm = ldodd(nc);
                const double *Gl = s.Gs + p.goff[l];
                double *M = s.Mi + p.moff[l];
                double *vec = s.vec + wid * 256;
                if (it == 0) {
                    const bool ok = warp_build_minv<LT>(Gl, ldg, nc, W, s.w + l, M, ldm, vec);
                    if (lane == 0) {
                        bad[l] = ok ? 0 : 1;
                        if (!ok) atomicAdd(p.flags, 1);
                    }
                    __syncwarp();
                }
                if (bad[l]) {
                    for (int t = lane; t < W; t += 32) s.dmu[t * LT + l] = 0.0;
                } else {
                    warp_mean_step<LT>(Gl, ldg, nc, W, M, ldm, s.ra + l, s.w + l, s.mu + l, s.dmu + l, vec,
                                       p.dmu_bound);
                }
            }
            __syncthreads();
            rate_pass<LT, 2>(p, s, bin0);
            if (p.method_vb || it + 1 < p.n_iter) {
                for (int l = wid; l < LT; l += NWARP) {
                    const int nc = p.nc[l], ldg = ldodd(nc), ldm = ldodd(nc);
                    const double *Gl = s.Gs + p.goff[l];
                    double *M = s.Mi + p.moff[l];
                    double *vec = s.vec + wid * 256;
                    const bool ok = warp_build_minv<LT>(Gl, ldg, nc, W, s.w + l, M, ldm, vec);
                    if (lane == 0) {
                        bad[l] = ok ? 0 : 1;
                        if (!ok) atomicAdd(p.flags, 1);
                    }
                    if (ok && p.method_vb) warp_variance<LT>(Gl, ldg, nc, W, M, ldm, s.v + l);
                }
            }
            __syncthreads();
        }
        for (int i = tid; i < W * LT; i += NT) {
            p.mu[bin0 * LT + i] = s.mu[i];
            p.v[bin0 * LT + i] = s.v[i];
            p.w[bin0 * LT + i] = s.w[i];
            p.dmu[bin0 * LT + i] = s.dmu[i];
        }
        __syncthreads();
    }
}

template <int LT>
int launch_seg_t(vlgp_ctx *ctx, TrialSet *ts, SegArgs &p, size_t smem, bool *handled) {
    CK(cudaFuncSetAttribute(estep_seg_kernel<LT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, estep_seg_kernel<LT>, NT, smem));
    if (per_sm < 1) return VLGP_OK;       // does not fit: let the general kernel handle it
    int grid = per_sm * ctx->prop.multiProcessorCount;
    if (grid > p.n_seg) grid = p.n_seg;
    estep_seg_kernel<LT><<<grid, NT, smem, ctx->stream>>>(p);
    CKL();
    *handled = true;
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

int vlgp_launch_estep_segments(vlgp_ctx *ctx, TrialSet *ts, int n_iter, double dmu_bound, int method_vb, bool *handled) {
    *handled = false;
    if (getenv("VLGP_FORCE_GENERIC_ESTEP")) return VLGP_OK;
    if (ts->min_len != ts->max_len || ts->max_len > VLGP_MAX_W || ts->factors.size() != 1) return VLGP_OK;
    const int W = ts->max_len, L = ctx->L, N = ctx->N;
    SegArgs p{};
    p.n_seg = ts->n_trials; p.W = W; p.N = N; p.rank = ctx->rank;
    p.G = ts->factors[0].d_G;
    int goff = 0, moff = 0;
    for (int l = 0; l < L; ++l) {
        const int nc = ts->factors[0].h_ncol[l];
        p.nc[l] = nc;
        p.goff[l] = goff;
        p.moff[l] = moff;
        goff += W * (nc | 1);
        moff += (nc > 0 ? nc : 1) * (nc | 1);
    }
    p.g_total = goff; p.m_total = moff;
    p.y = ts->d_y; p.ydtype = ts->ydtype;
    p.mu = ts->d_mu; p.v = ts->d_v; p.w = ts->d_w; p.dmu = ts->d_dmu;
    p.a = ctx->d_a; p.b = ctx->d_b; p.noise = ctx->d_noise; p.poisson = ctx->d_poisson;
    p.n_iter = n_iter; p.dmu_bound = dmu_bound; p.method_vb = method_vb; p.flags = ctx->d_flags;
    p.tpb = NT / W;
    if (p.tpb < 1) return VLGP_OK;
    if (p.tpb > N) p.tpb = N;
    p.chunk = (N + p.tpb - 1) / p.tpb;
    const size_t smem = seg_smem_bytes(L, N, W, p.g_total, p.m_total, p.tpb, ts->ydtype == VLGP_Y_U8);
    if (smem > (size_t)ctx->prop.sharedMemPerBlockOptin) return VLGP_OK;
    int rc = VLGP_OK;
    DISPATCH_L(L, rc = launch_seg_t<LT>(ctx, ts, p, smem, handled));
    return rc;
}
