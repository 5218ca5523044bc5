// K3: variational E-step for LONG trials (any length, prior factor of up to 56 columns) -- the full-trial inference that
// follows vem (core.infer, vlgp/core.py:260-266, vlgp/api.py:66-71) and transform; replaces the one-CTA-per-trial
// kernel of estep.cu for Poisson channels with uint8 counts (everything else still goes there).
//
// A trial of T bins is far too long for one CTA to own (T x rank factor, T x N counts), and too few trials exist to fill
// 148 SMs with one CTA each.  So the iteration is cut into phases that are each embarrassingly parallel over ITEMS of
// 256 bins (4 chunks of 64 bins = 8 row tiles of the FP64 tensor path), and the small per-(trial, latent) objects that
// couple the items travel through the 126 MB L2 as partial sums:
//   k3_gram   (item)            w-weighted Gram partials  G_l' W_l G_l  over the item's bins, by DMMA       -> apart
//   k3_factor (trial, latent)   A = I + sum of partials, blocked symmetric sweep in registers (one warp)    -> Minv
//   k3_var    (item)            v_t = G_t Minv G_t' for the item's bins; Minv_l staged in shared memory by TMA
//                               (cp.async.bulk + mbarrier, double-buffered over the latents)
//   k3_a      (item)            rate pass 1, z = (y - rate) a_l' + w o mu, partials of s_l = G_l' z_l       -> spart
//   k3_solve  (trial, latent)   m_l = Minv_l (sum of partials)                                              -> mvec
//   k3_c      (item)            mu += clip(G_l m_l - mu), rate pass 2 -> w, then this item's Gram partials  -> apart
// Same algebra as the fused segment pipeline (estep_seg_impl.cuh: three-product mean step, one factorisation per latent
// and iteration shared by the variance and the next mean step), same rate-pass tile loop (rate_tiles_core).  Sums over
// partials are taken in item order: results are deterministic.
#include "estep_seg_impl.cuh"
#include "tma.cuh"

namespace k3 {

constexpr int NT = 256;
constexpr int NWARP = NT / 32;
constexpr int CH = 64;               // bins per chunk: one row tile per warp
constexpr int SCN = 4;               // chunks per item
constexpr int IB = CH * SCN;         // bins per item

struct Args {
    int n_items, n_trials, N, rank, np, kp;
    const int *item_trial, *item_t0, *trial_item0;
    const int *len;
    const int64_t *start;
    const int *fidx;
    double *const *Gptr;
    const uint8_t *y;
    double *mu, *v, *w, *dmu, *ya;
    const double2 *pa, *pb;
    double *spart, *mvec, *apart, *Minv;
    int *bad;
    double dmu_bound;
    int *flags;
    int first, gram_only;
    int dbg;                     // timing experiments only (VLGP_K3_DEBUG): 1 = Gram without its bulk copies
};

// operand of the tensor-path rate passes (as in estep_seg_kernel): rows a_l | a_l^2 / 2 | b | 0, columns padded with 0
template <int LT>
__device__ __forceinline__ void stage_bx(const Args &p, double *Bx, double *etab) {
    const int tid = threadIdx.x, N = p.N;
    for (int i = tid; i < p.kp * p.np; i += NT) {
        const int k = i / p.np, n = i - k * p.np;
        double val = 0.0;
        if (n < N) {
            if (k < LT) val = p.pa[k * N + n].x;
            else if (k < 2 * LT) val = 0.5 * p.pa[(k - LT) * N + n].y;
            else if (k == 2 * LT) val = p.pb[n].x;
        }
        Bx[i] = val;
    }
    if (tid < 32) etab[tid] = VLGP_EXP_T[tid];
}

// ---- rate pass 1 + partial projections ------------------------------------------------------------------------------
template <int LT>
__global__ void __launch_bounds__(NT, 3) k3_a_kernel(Args p) {
    constexpr int NOT = (LT + 7) / 8;
    extern __shared__ __align__(16) unsigned char raw[];
    double *Bx = (double *)raw;
    double *etab = Bx + p.kp * p.np;
    double *smu = etab + 32, *sv = smu + CH * LT, *sz = sv + CH * LT;
    double *spw = sz + CH * LT;                        // NWARP x LT x 64
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, r = lane >> 2, q = lane & 3;
    const int item = blockIdx.x, trial = p.item_trial[item], t0 = p.item_t0[item];
    const int T = p.len[trial], rank = p.rank, N = p.N;
    const int64_t s0 = p.start[trial];
    const double *G = p.Gptr[p.fidx[trial]];
    stage_bx<LT>(p, Bx, etab);
    for (int i = lane; i < LT * 64; i += 32) spw[wid * LT * 64 + i] = 0.0;
    __syncthreads();
    for (int c = 0; c < SCN; ++c) {
        const int tc = t0 + CH * c;
        if (tc >= T) break;
        const int nbv = min(CH, T - tc);
        if (8 * wid >= nbv) continue;
        const int64_t bin0 = s0 + tc;
        __syncwarp();
        for (int i = lane; i < 8 * LT; i += 32) {
            const bool in = 8 * wid + i / LT < nbv;
            smu[8 * wid * LT + i] = in ? p.mu[(bin0 + 8 * wid) * LT + i] : 0.0;
            sv[8 * wid * LT + i] = in ? p.v[(bin0 + 8 * wid) * LT + i] : 0.0;
        }
        __syncwarp();
        const int t = 8 * wid + r;
        const bool tin = t < nbv;
        if (p.first) {                                 // y a_l' for this row tile (constant during the launch sequence)
            const uint8_t *yrow = p.y + (bin0 + (tin ? t : 0)) * N;
            Tile ya[NOT];
#pragma unroll
            for (int o = 0; o < NOT; ++o) ya[o].x = ya[o].y = 0.0;
            for (int j = 0; j < (p.np >> 3); ++j) {
                const int n0 = 8 * j + 2 * q;
                const double y0 = (tin && n0 < N) ? (double)yrow[n0] : 0.0;
                const double y1 = (tin && n0 + 1 < N) ? (double)yrow[n0 + 1] : 0.0;
#pragma unroll
                for (int o = 0; o < NOT; ++o) {
                    double2 b2 = make_double2(0.0, 0.0);
                    if (8 * o + r < LT) b2 = *reinterpret_cast<const double2 *>(Bx + (r + 8 * o) * p.np + n0);
                    dmma(ya[o], y0, b2.x);
                    dmma(ya[o], y1, b2.y);
                }
            }
            if (tin) {
#pragma unroll
                for (int o = 0; o < NOT; ++o) {
                    const int l0 = 8 * o + 2 * q;
                    if (l0 < LT) p.ya[(bin0 + t) * LT + l0] = ya[o].x;
                    if (l0 + 1 < LT) p.ya[(bin0 + t) * LT + l0 + 1] = ya[o].y;
                }
            }
        }
        Tile acc[NOT];
        segk::rate_tiles_core<LT, 1, true>(nbv, N, p.np, Bx, smu, sv, etab, nullptr, acc);
        if (tin) {
#pragma unroll
            for (int o = 0; o < NOT; ++o) {
                const int l0 = 8 * o + 2 * q;
                const int64_t g = (bin0 + t) * LT + l0;
                if (l0 < LT) sz[t * LT + l0] = fma(p.w[g], smu[t * LT + l0], p.ya[g] - acc[o].x);
                if (l0 + 1 < LT) sz[t * LT + l0 + 1] = fma(p.w[g + 1], smu[t * LT + l0 + 1], p.ya[g + 1] - acc[o].y);
            }
        }
        __syncwarp();
        const int rows = min(8, nbv - 8 * wid);
        for (int idx = lane; idx < LT * 64; idx += 32) {
            const int l = idx >> 6, j = idx & 63;
            if (j >= rank) continue;
            const double *g = G + ((size_t)l * T + tc + 8 * wid) * rank + j;
            const double *z = sz + 8 * wid * LT + l;
            double a = 0.0;
            for (int tt = 0; tt < rows; ++tt) a = fma(g[(size_t)tt * rank], z[tt * LT], a);
            spw[wid * LT * 64 + idx] += a;
        }
    }
    __syncthreads();
    for (int idx = tid; idx < LT * 64; idx += NT) {
        double a = 0.0;
#pragma unroll
        for (int w = 0; w < NWARP; ++w) a += spw[w * LT * 64 + idx];
        p.spart[(size_t)item * LT * 64 + idx] = a;
    }
}

// ---- m_l = Minv_l s_l : one warp per (trial, latent) -----------------------------------------------------------------
template <int NB>
__global__ void __launch_bounds__(128) k3_solve_kernel(Args p, int LT) {
    constexpr int LDM = 8 * NB + 4;
    __shared__ double sv[4][64];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int task = blockIdx.x * 4 + wid;
    if (task >= p.n_trials * LT) return;
    const int trial = task / LT, l = task - trial * LT;
    const int i0 = p.trial_item0[trial], i1 = p.trial_item0[trial + 1];
    for (int j = lane; j < 64; j += 32) {
        double a = 0.0;
        for (int it = i0; it < i1; ++it) a += p.spart[((size_t)it * LT + l) * 64 + j];
        sv[wid][j] = a;
    }
    __syncwarp();
    const double *M = p.Minv + (size_t)task * 8 * NB * LDM;           // holds -Minv
    for (int i = lane; i < 64; i += 32) {
        double a = 0.0;
        if (i < 8 * NB)
            for (int j = 0; j < 8 * NB; ++j) a = fma(M[i * LDM + j], sv[wid][j], a);
        p.mvec[(size_t)task * 64 + i] = -a;
    }
}

// ---- mean update + rate pass 2 + Gram partials ----------------------------------------------------------------------
template <int LT, int NB>
__global__ void __launch_bounds__(NT, 2) k3_c_kernel(Args p) {
    constexpr int NOT = (LT + 7) / 8;
    constexpr int NTL = NB * (NB + 1) / 2;
    extern __shared__ __align__(16) unsigned char raw[];
    double *Bx = (double *)raw;
    double *etab = Bx + p.kp * p.np;
    double *smu = etab + 32, *sv = smu + CH * LT;
    double *wsc = sv + CH * LT;                        // IB x LT : the item's weights
    double *mv = wsc + IB * LT;                        // LT x 64
    double *Gs = mv + LT * 64;                         // 2 x CH x rank : staged rows of G_l (TMA destination)
    __shared__ int sbad[VLGP_MAX_L];
    __shared__ __align__(8) uint64_t gbar[2];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, r = lane >> 2, q = lane & 3;
    const int item = blockIdx.x, trial = p.item_trial[item], t0 = p.item_t0[item];
    const int T = p.len[trial], rank = p.rank, N = p.N;
    const int64_t s0 = p.start[trial];
    const double *G = p.Gptr[p.fidx[trial]];
    const int nb_item = min(IB, T - t0);
    if (tid == 0) {
        mbar_init(&gbar[0], 1);
        mbar_init(&gbar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (!p.gram_only) {
        stage_bx<LT>(p, Bx, etab);
        for (int i = tid; i < LT * 64; i += NT) mv[i] = p.mvec[(size_t)trial * LT * 64 + i];
        if (tid < LT) sbad[tid] = p.bad[trial * LT + tid];
    }
    __syncthreads();
    for (int c = 0; c < SCN; ++c) {
        const int tc = t0 + CH * c;
        if (tc >= T) break;
        const int nbv = min(CH, T - tc);
        if (8 * wid >= nbv) continue;
        const int64_t bin0 = s0 + tc;
        if (p.gram_only) {
            for (int i = lane; i < 8 * LT; i += 32)
                if (8 * wid + i / LT < nbv) wsc[(CH * c + 8 * wid) * LT + i] = p.w[(bin0 + 8 * wid) * LT + i];
            continue;
        }
        __syncwarp();
        for (int i = lane; i < 8 * LT; i += 32) {
            const bool in = 8 * wid + i / LT < nbv;
            smu[8 * wid * LT + i] = in ? p.mu[(bin0 + 8 * wid) * LT + i] : 0.0;
            sv[8 * wid * LT + i] = in ? p.v[(bin0 + 8 * wid) * LT + i] : 0.0;
        }
        __syncwarp();
        for (int idx = lane; idx < 8 * LT; idx += 32) {         // delta = clip(G m - mu); a failed factorisation zeroes it
            const int l = idx >> 3, t = 8 * wid + (idx & 7);
            if (t < nbv) {
                double d = 0.0;
                if (!sbad[l]) {
                    const double *g = G + ((size_t)l * T + tc + t) * rank;
                    const double *m = mv + l * 64;
                    double a = 0.0;
                    for (int j = 0; j < rank; ++j) a = fma(g[j], m[j], a);
                    d = clipd(a - smu[t * LT + l], p.dmu_bound);
                }
                const double munew = smu[t * LT + l] + d;
                smu[t * LT + l] = munew;
                p.mu[(bin0 + t) * LT + l] = munew;
                p.dmu[(bin0 + t) * LT + l] = d;
            }
        }
        __syncwarp();
        Tile acc[NOT];
        segk::rate_tiles_core<LT, 2, true>(nbv, N, p.np, Bx, smu, sv, etab, nullptr, acc);
        const int t = 8 * wid + r;
        if (t < nbv) {
#pragma unroll
            for (int o = 0; o < NOT; ++o) {
                const int l0 = 8 * o + 2 * q;
                if (l0 < LT) {
                    const double wv = 2.0 * acc[o].x;                        // Bx holds a^2 / 2
                    wsc[(CH * c + t) * LT + l0] = wv;
                    p.w[(bin0 + t) * LT + l0] = wv;
                }
                if (l0 + 1 < LT) {
                    const double wv = 2.0 * acc[o].y;
                    wsc[(CH * c + t) * LT + l0 + 1] = wv;
                    p.w[(bin0 + t) * LT + l0 + 1] = wv;
                }
            }
        }
    }
    __syncthreads();
    // Gram partials.  The item's rows of G_l arrive in shared memory chunk by chunk (64 rows x rank doubles, contiguous in
    // HBM / L2: one TMA bulk copy each, double-buffered over the (latent, chunk) steps); the NTL tiles of a latent are
    // dealt round-robin to the 8 warps and stay in registers over the item's k4 steps.
    constexpr int TPW = (NTL + NWARP - 1) / NWARP;
    int ti[TPW], tj[TPW];
#pragma unroll
    for (int m = 0; m < TPW; ++m) {
        ti[m] = tj[m] = -1;
        if (wid + NWARP * m < NTL) tri_decode(wid + NWARP * m, ti[m], tj[m]);
    }
    const int nch = (nb_item + CH - 1) / CH, nsteps = LT * nch;
    auto issue = [&](int st) {
        const int l = st / nch, c = st - l * nch;
        const unsigned bytes = (unsigned)(min(CH, nb_item - CH * c) * rank * sizeof(double));
        mbar_expect_tx(&gbar[st & 1], bytes);
        tma_load_1d(Gs + (st & 1) * CH * rank, G + ((size_t)l * T + t0 + CH * c) * rank, bytes, &gbar[st & 1]);
    };
    if (tid == 0 && !(p.dbg & 1)) issue(0);
    Tile A[TPW];
    for (int st = 0; st < nsteps; ++st) {
        const int l = st / nch, c = st - l * nch;
        if (tid == 0 && st + 1 < nsteps && !(p.dbg & 1)) issue(st + 1);   // the other buffer was released by the barrier below
        if (!(p.dbg & 1)) mbar_wait(&gbar[st & 1], (st >> 1) & 1);
        const double *Gb = Gs + (st & 1) * CH * rank;
        if (c == 0) {
#pragma unroll
            for (int m = 0; m < TPW; ++m) A[m].x = A[m].y = 0.0;
        }
        const int nbc = min(CH, nb_item - CH * c);
        // k4 step k takes the bins 16 (k / 4) + k % 4 + {0, 4, 8, 12}: rows 4 apart are 200 doubles apart (rank = 50), so
        // the four rows a warp reads fall on disjoint halves of the banks -- the natural choice t = 4 k + q is a 4-way
        // conflict.  (A contraction does not care in which order its bins are taken.)
#pragma unroll 4
        for (int k = 0; 16 * (k >> 2) < nbc; ++k) {
            const int t = 16 * (k >> 2) + (k & 3) + 4 * q;
            const bool tin = t < nbc;
            const double wt = tin ? wsc[(CH * c + t) * LT + l] : 0.0;
            const double *grow = Gb + (tin ? t : 0) * rank;
#pragma unroll
            for (int m = 0; m < TPW; ++m) {
                if (ti[m] < 0) continue;
                const int ci = 8 * ti[m] + r, cj = 8 * tj[m] + r;
                const double gi = (tin && ci < rank) ? grow[ci] * wt : 0.0;
                const double gj = (tin && cj < rank) ? grow[cj] : 0.0;
                dmma(A[m], gi, gj);
            }
        }
        if (c == nch - 1) {
            double2 *out = (double2 *)p.apart + ((size_t)item * LT + l) * NTL * 32 + lane;
#pragma unroll
            for (int m = 0; m < TPW; ++m)
                if (ti[m] >= 0) out[(wid + NWARP * m) * 32] = make_double2(A[m].x, A[m].y);
        }
        __syncthreads();
    }
}

// ---- A = I + sum of partials, sweep, -Minv to global: one warp per (trial, latent) ----------------------------------
template <int NB>
__global__ void __launch_bounds__(128) k3_factor_kernel(Args p, int LT) {
    constexpr int NTL = NB * (NB + 1) / 2;
    constexpr int LDM = 8 * NB + 4;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, r = lane >> 2, c0 = 2 * (lane & 3);
    const int task = blockIdx.x * 4 + wid;
    if (task >= p.n_trials * LT) return;
    const int trial = task / LT, l = task - trial * LT;
    const int i0 = p.trial_item0[trial], i1 = p.trial_item0[trial + 1];
    Tile A[NTL];
#pragma unroll
    for (int t = 0; t < NTL; ++t) A[t].x = A[t].y = 0.0;
    for (int it = i0; it < i1; ++it) {
        const double2 *src = (const double2 *)p.apart + ((size_t)it * LT + l) * NTL * 32 + lane;
#pragma unroll
        for (int t = 0; t < NTL; ++t) {
            const double2 vv = src[t * 32];
            A[t].x += vv.x;
            A[t].y += vv.y;
        }
    }
#pragma unroll
    for (int i = 0; i < NB; ++i) {
        A[tix(i, i)].x += (r == c0) ? 1.0 : 0.0;
        A[tix(i, i)].y += (r == c0 + 1) ? 1.0 : 0.0;
    }
    const bool ok = tile_sweep<NB>(A, lane);
    double *M = p.Minv + (size_t)task * 8 * NB * LDM;
#pragma unroll
    for (int i = 0; i < NB; ++i)
#pragma unroll
        for (int j = 0; j <= i; ++j) {
            const Tile t = A[tix(i, j)];
            M[(8 * i + r) * LDM + 8 * j + c0] = t.x;
            M[(8 * i + r) * LDM + 8 * j + c0 + 1] = t.y;
            if (i != j) {
                M[(8 * j + c0) * LDM + 8 * i + r] = t.x;
                M[(8 * j + c0 + 1) * LDM + 8 * i + r] = t.y;
            }
        }
    if (lane == 0) {
        p.bad[task] = ok ? 0 : 1;
        if (!ok) atomicAdd(p.flags, 1);
    }
}

// ---- variances: Minv_l staged in shared memory by TMA bulk copies, double-buffered over the latents -------------------
template <int NB>
__global__ void __launch_bounds__(NT, 2) k3_var_kernel(Args p, int LT) {
    constexpr int LDM = 8 * NB + 4;
    constexpr unsigned MBYTES = 8 * NB * LDM * sizeof(double);
    extern __shared__ __align__(128) unsigned char raw[];
    double *Ms = (double *)raw;                        // 2 x (8 NB x LDM)
    __shared__ __align__(8) uint64_t bar[2];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, r = lane >> 2, q = lane & 3, c0 = 2 * q;
    const int item = blockIdx.x, trial = p.item_trial[item], t0 = p.item_t0[item];
    const int T = p.len[trial], rank = p.rank;
    const int64_t s0 = p.start[trial];
    const double *G = p.Gptr[p.fidx[trial]];
    const int nb_item = min(IB, T - t0);
    const double *Mg = p.Minv + (size_t)trial * LT * 8 * NB * LDM;
    if (tid == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
        mbar_expect_tx(&bar[0], MBYTES);
        tma_load_1d(Ms, Mg, MBYTES, &bar[0]);
    }
    for (int l = 0; l < LT; ++l) {
        const int st = l & 1;
        if (tid == 0 && l + 1 < LT) {                  // prefetch the next latent's inverse into the other buffer
            mbar_expect_tx(&bar[st ^ 1], MBYTES);
            tma_load_1d(Ms + (st ^ 1) * 8 * NB * LDM, Mg + (size_t)(l + 1) * 8 * NB * LDM, MBYTES, &bar[st ^ 1]);
        }
        mbar_wait(&bar[st], (l >> 1) & 1);
        const double *M = Ms + st * 8 * NB * LDM;
        if (!p.bad[trial * LT + l]) {                  // a failed solve keeps v (vlgp/core.py:112)
            const double *Gl = G + ((size_t)l * T + t0) * rank;
            for (int rt = wid; 8 * rt < nb_item; rt += NWARP) {
                const int trow = 8 * rt + r;
                const bool tin = trow < nb_item;
                const double *grow = Gl + (size_t)(tin ? trow : 0) * rank;
                double aop[2 * NB];
#pragma unroll
                for (int k = 0; k < 2 * NB; ++k) {
                    const int c = 4 * k + q;
                    aop[k] = (tin && c < rank) ? grow[c] : 0.0;
                }
                // g' M g over the lower block triangle only (M is symmetric): per column block jt the diagonal block
                // once and the blocks below it twice -- NB (NB + 1) DMMA per row tile instead of 2 NB^2
                double acc = 0.0;
#pragma unroll
                for (int jt = 0; jt < NB; ++jt) {
                    Tile Td{0.0, 0.0}, To{0.0, 0.0};
#pragma unroll
                    for (int k = 2 * jt; k < 2 * jt + 2; ++k)
                        if (4 * k < rank) dmma(Td, aop[k], M[(4 * k + q) * LDM + 8 * jt + r]);
#pragma unroll
                    for (int k = 2 * jt + 2; k < 2 * NB; ++k)
                        if (4 * k < rank) dmma(To, aop[k], M[(4 * k + q) * LDM + 8 * jt + r]);
                    const int c = 8 * jt + c0;
                    const double g0 = (tin && c < rank) ? grow[c] : 0.0;
                    const double g1 = (tin && c + 1 < rank) ? grow[c + 1] : 0.0;
                    acc = fma(fma(2.0, To.x, Td.x), g0, acc);
                    acc = fma(fma(2.0, To.y, Td.y), g1, acc);
                }
                acc += __shfl_xor_sync(FULL, acc, 1);
                acc += __shfl_xor_sync(FULL, acc, 2);
                if (q == 0 && tin) p.v[(s0 + t0 + trow) * LT + l] = -acc;
            }
        }
        __syncthreads();                               // everyone is done with buffer st before it is refilled
    }
}

template <int LT, int NB>
int launch(vlgp_ctx *ctx, TrialSet *ts, Args &p, int n_iter, int method_vb) {
    constexpr int NTL = NB * (NB + 1) / 2;
    constexpr int LDM = 8 * NB + 4;
    const int tasks = p.n_trials * LT;
    const size_t smem_a = ((size_t)p.kp * p.np + 32 + 3 * CH * LT + NWARP * LT * 64) * sizeof(double);
    const size_t smem_c = ((size_t)p.kp * p.np + 32 + 2 * CH * LT + IB * LT + LT * 64 + 2 * CH * p.rank) * sizeof(double);
    const size_t smem_v = (size_t)2 * 8 * NB * LDM * sizeof(double);
    CK(cudaFuncSetAttribute(k3_a_kernel<LT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_a));
    CK(cudaFuncSetAttribute(k3_c_kernel<LT, NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_c));
    CK(cudaFuncSetAttribute(k3_var_kernel<NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_v));
    (void)NTL;
    const int tgrid = (tasks + 3) / 4;
    p.gram_only = 1;
    k3_c_kernel<LT, NB><<<p.n_items, NT, smem_c, ctx->stream>>>(p);           // the first mean step uses the incoming w
    CKL();
    k3_factor_kernel<NB><<<tgrid, 128, 0, ctx->stream>>>(p, LT);
    CKL();
    p.gram_only = 0;
    for (int it = 0; it < n_iter; ++it) {
        if (it > 0 && method_vb) {
            k3_var_kernel<NB><<<p.n_items, NT, smem_v, ctx->stream>>>(p, LT);
            CKL();
        }
        p.first = it == 0;
        k3_a_kernel<LT><<<p.n_items, NT, smem_a, ctx->stream>>>(p);
        CKL();
        k3_solve_kernel<NB><<<tgrid, 128, 0, ctx->stream>>>(p, LT);
        CKL();
        k3_c_kernel<LT, NB><<<p.n_items, NT, smem_c, ctx->stream>>>(p);
        CKL();
        if (method_vb || it + 1 < n_iter) {
            k3_factor_kernel<NB><<<tgrid, 128, 0, ctx->stream>>>(p, LT);
            CKL();
        }
    }
    if (method_vb && n_iter > 0) {
        k3_var_kernel<NB><<<p.n_items, NT, smem_v, ctx->stream>>>(p, LT);
        CKL();
    }
    return VLGP_OK;
}

template <int NB>
int launch_nb(vlgp_ctx *ctx, TrialSet *ts, Args &p, int n_iter, int method_vb) {
    int rc = VLGP_OK;
    DISPATCH_L(ctx->L, (rc = launch<LT, NB>(ctx, ts, p, n_iter, method_vb)));
    return rc;
}

}   // namespace k3

// Sets *handled when the long-trial path took the call: all channels Poisson, uint8 counts, bias-only regressors, every
// trial (not a subset), rank <= 56.  Scratch (item tables, partials, inverses) is cached on the trial set.
int vlgp_launch_estep_long(vlgp_ctx *ctx, TrialSet *ts, int n_iter, double dmu_bound, int method_vb, bool *handled) {
    using namespace k3;
    *handled = false;
    if (getenv("VLGP_NO_LONG_ESTEP") || n_iter < 1) return VLGP_OK;
    if (ctx->any_gauss || ts->ydtype != VLGP_Y_U8 || ts->d_x || ctx->rank > 56 || ctx->L > 12) return VLGP_OK;
    const int L = ctx->L, N = ctx->N;
    const int NB = ctx->rank <= 32 ? 4 : 7;
    const int NTL = NB * (NB + 1) / 2, LDM = 8 * NB + 4;
    if (!ts->k3_ready) {
        std::vector<int> it_trial, it_t0, tr_item0(ts->n_trials + 1, 0);
        for (int i = 0; i < ts->n_trials; ++i) {
            tr_item0[i] = (int)it_trial.size();
            for (int t0 = 0; t0 < ts->h_len[i]; t0 += IB) {
                it_trial.push_back(i);
                it_t0.push_back(t0);
            }
        }
        tr_item0[ts->n_trials] = (int)it_trial.size();
        ts->k3_items = (int)it_trial.size();
        const size_t ni = it_trial.size();
        CK(vlgp_dalloc(ctx, &ts->d_k3_tab, (2 * ni + tr_item0.size()) * sizeof(int)));
        CK(cudaMemcpyAsync(ts->d_k3_tab, it_trial.data(), ni * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(ts->d_k3_tab + ni, it_t0.data(), ni * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(ts->d_k3_tab + 2 * ni, tr_item0.data(), tr_item0.size() * sizeof(int), cudaMemcpyHostToDevice,
                           ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));        // the host vectors go out of scope
        const size_t tasks = (size_t)ts->n_trials * L;
        const size_t nd = (size_t)ts->nbin * L + ni * L * 64 + tasks * 64 + ni * L * NTL * 64 + tasks * 8 * NB * LDM;
        CK(vlgp_dalloc(ctx, &ts->d_k3_buf, nd * sizeof(double)));
        CK(vlgp_dalloc(ctx, &ts->d_k3_bad, tasks * sizeof(int)));
        ts->k3_ready = true;
    }
    Args p{};
    const size_t ni = (size_t)ts->k3_items, tasks = (size_t)ts->n_trials * L;
    p.n_items = ts->k3_items; p.n_trials = ts->n_trials; p.N = N; p.rank = ctx->rank;
    p.np = 8 * ((N + 7) / 8);
    if (p.np % 16 == 0) p.np += 8;
    p.kp = 4 * ((2 * L + 1 + 3) / 4);
    p.item_trial = ts->d_k3_tab; p.item_t0 = ts->d_k3_tab + ni; p.trial_item0 = ts->d_k3_tab + 2 * ni;
    p.len = ts->d_len; p.start = ts->d_start; p.fidx = ts->d_fidx; p.Gptr = ts->d_Gptr;
    p.y = (const uint8_t *)ts->d_y;
    p.mu = ts->d_mu; p.v = ts->d_v; p.w = ts->d_w; p.dmu = ts->d_dmu;
    p.Minv = ts->d_k3_buf;                              // first: the bulk copies need 16-byte aligned sources
    p.apart = p.Minv + tasks * 8 * NB * LDM;
    p.spart = p.apart + ni * L * NTL * 64;
    p.mvec = p.spart + ni * L * 64;
    p.ya = p.mvec + tasks * 64;
    p.bad = ts->d_k3_bad;
    p.dmu_bound = dmu_bound; p.flags = ctx->d_flags;
    p.dbg = getenv("VLGP_K3_DEBUG") ? atoi(getenv("VLGP_K3_DEBUG")) : 0;
    // (a, a^2) and (b, 1 / noise) pairs, as for the segment kernel
    if (!ctx->d_ppack) CK(cudaMalloc(&ctx->d_ppack, (size_t)(VLGP_MAX_L + 1) * N * sizeof(double2)));
    p.pa = (const double2 *)ctx->d_ppack;
    p.pb = p.pa + L * N;
    segk::pack_params_kernel<<<(L * N + 255) / 256, 256, 0, ctx->stream>>>(L * N, N, ctx->d_a, ctx->d_b, ctx->d_noise,
                                                                           (double2 *)p.pa, (double2 *)p.pb);
    CKL();
    const size_t smem_c = ((size_t)p.kp * p.np + 32 + 2 * CH * L + IB * L + L * 64 + 2 * CH * p.rank) * sizeof(double);
    if (smem_c > (size_t)ctx->prop.sharedMemPerBlockOptin) return VLGP_OK;
    int rc = NB == 4 ? launch_nb<4>(ctx, ts, p, n_iter, method_vb) : launch_nb<7>(ctx, ts, p, n_iter, method_vb);
    if (rc == VLGP_OK) *handled = true;
    return rc;
}
