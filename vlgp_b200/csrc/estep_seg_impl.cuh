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
#pragma once
#include "common.cuh"
#include "dmma.cuh"
#include "linalg.cuh"

namespace segk {

constexpr int NT = 256;
constexpr int NWARP = NT / 32;

struct SegArgs {
    int n_seg, W, N, rank;
    const int32_t *subset;        // n_seg segment indices to process (vlgp_estep_subset), or null: segments 0..n_seg-1
    const double *G;              // L x W x rank
    int nc[VLGP_MAX_L];           // leading non-zero columns per latent
    int goff[VLGP_MAX_L];         // offsets (doubles) of the compact factor / Minv of latent l inside their regions
    int moff[VLGP_MAX_L];
    int g_total, m_total;         // region sizes (doubles)
    int pairoff[VLGP_MAX_L + 1];  // prefix sums of nc (nc + 1) / 2 : (latent, Gram entry) items
    int coloff[VLGP_MAX_L + 1];   // prefix sums of nc : (latent, column) items
    int pair_total, col_total;
    int ldm[VLGP_MAX_L];          // leading dimension of latent l's r x r matrix in SMEM
    int ldg[VLGP_MAX_L];          // leading dimension of latent l's factor: nc | 1 (compact copy in SMEM) or rank (in place)
    int g_global;                 // 1: the factors do not fit in SMEM next to everything else and are read in place from
                                  //    HBM / L2 (NBMAX = 4 variants only; goff then indexes p.G, g_total = 0)
    int use_dmma;                 // every nc <= 32: Gram / inverse / variance on the FP64 tensor path
    int fused;                    // tensor-path rate passes and use_dmma: the fused pipeline (fused_a .. fused_c)
    int f32;                      // fused pipeline with single-precision rate passes (FAST == 2 instantiations)
    const void *y;
    int ydtype;
    double *mu, *v, *w, *dmu;
    const double *a, *b, *noise;
    const uint8_t *poisson;
    const double2 *pa;            // L x N (a, a^2) pairs and N (b, 1/noise) pairs, packed by pack_params_kernel
    const double2 *pb;
    int n_iter;
    double dmu_bound;
    int method_vb;
    int *flags;
    int tpb, chunk;               // threads per bin and neurons per thread in the rate passes
    int np, kp;                   // tensor-path rate passes: padded neuron count (= 8 mod 16) and rows of the operand Bx
    int *sm_slots;                // per-SM arrival counters (zeroed before the launch), see estep_seg_kernel
    int stagger;                  // cycles by which the k-th CTA to arrive on an SM delays its start
    int skip;                     // debug/timing only (VLGP_DEBUG_SKIP): 1 rate passes, 2 mean step, 4 factor, 8 variance
};

__device__ __forceinline__ int ldodd(int n) { return n | 1; }

template <int LT>
struct Smem {
    double *a, *a2, *b, *inv_noise, *Mi, *mu, *v, *w, *ra, *dmu, *part, *vec, *ya, *etab;
    const double *Gs;             // compact factors in SMEM, or p.G itself (SegArgs::g_global)
    uint8_t *pois, *ys;
    __device__ Smem(unsigned char *base, const SegArgs &p) {
        double *d = (double *)base;
        const int N = p.N, W = p.W;
        a = d;                                   // interleaved (a, a^2) pairs: one 128-bit load per (latent, neuron)
        a2 = a;                                  // (tensor-path rate passes: the kp x np operand Bx lives here instead)
        b = d + 2 * LT * N;                      // interleaved (bias, 1 / noise) pairs
        inv_noise = b;
        d += p.f32 ? LT * N + (N + 2) / 2 : max(2 * LT * N + 2 * N, p.kp * p.np);
        Gs = d; d += p.g_total;                  // (g_total = 0 and Gs re-pointed by the kernel when g_global)
        Mi = d; d += p.m_total;
        mu = d; d += W * LT;
        v = d; d += W * LT;
        w = d; d += W * LT;
        ra = d; d += W * LT;
        dmu = d; d += W * LT;
        part = d;                                 // rate passes: tpb x W x LT partial sums ...
        vec = d;                                  // ... aliased with the per-latent vectors (3 x 64) of the r x r phases
                                                  // (fused path: NWARP x col_total partial projections + col_total)
        d += max(max(max(p.tpb * W * LT, LT * 192), (NWARP + 1) * p.col_total), p.fused ? (W * N + 55) / 8 : 0);
        ya = d; d += W * LT;                      // fused path: y a_l' per (bin, latent), constant during the launch's iterations
        etab = d; d += 32;                        // 2^(j/32) (common.cuh: VLGP_EXP_T)
        pois = (uint8_t *)d;
        ys = pois + ((N + 15) / 16) * 16;
    }
};

__host__ __device__ inline size_t seg_smem_bytes(int LT, int N, int W, int g_total, int m_total, int tpb, bool y_u8,
                                                 int kp, int np, int col_total, bool f32 = false, bool fused = false) {
    size_t un = (size_t)tpb * W * LT > (size_t)LT * 192 ? (size_t)tpb * W * LT : (size_t)LT * 192;
    if ((size_t)(NWARP + 1) * col_total > un) un = (size_t)(NWARP + 1) * col_total;
    if (fused && (size_t)(W * N + 55) / 8 > un) un = (size_t)(W * N + 55) / 8;      // fused path: count tile staged here
    size_t par = (size_t)2 * LT * N + 2 * N > (size_t)kp * np ? (size_t)2 * LT * N + 2 * N : (size_t)kp * np;
    if (f32) par = (size_t)LT * N + (N + 2) / 2;
    size_t d = par + g_total + m_total + (size_t)6 * W * LT + 32 + un;
    size_t bytes = d * sizeof(double) + ((N + 15) / 16) * 16;
    if (y_u8) bytes += ((size_t)W * N + 15) / 16 * 16;
    return bytes;
}

// (a, a^2) of one (latent, neuron) entry of the packed loading in shared memory.  The rate passes are bound by
// shared-memory wavefronts, not by the FP64 pipe (ncu, config 2: 1.68e9 LSU wavefronts in a 9.4 ms launch = 61 % of the
// SM-cycles, FP64 pipe 35 %): every lane of a warp needs the pair of ITS neuron, a 128-bit load costs four wavefronts
// however few distinct addresses the warp touches, and there are L of them per (bin, neuron).  Reading only a (64-bit,
// two wavefronts) and squaring it in registers halves that traffic for one DMUL; a * a is the very value
// pack_params_kernel stored, so results are unchanged bit for bit.  -DVLGP_ESTEP_A2_FROM_SMEM restores the 128-bit load.
__device__ __forceinline__ double2 load_a(const double2 *aa, int idx) {
#ifdef VLGP_ESTEP_A2_FROM_SMEM
    return aa[idx];
#else
    const double x = reinterpret_cast<const double *>(aa)[2 * idx];
    return make_double2(x, x * x);
#endif
}

// One rate pass over the segment.  STAGE 1: part <- partial sums of resid * a_l ; STAGE 2: of U * a_l^2.
#ifdef VLGP_ESTEP_TWO_BINS
// Build option (not the default; never run on a GPU yet): every thread works on TWO consecutive bins, so that each
// fetch of the loading from shared memory is used twice and the two bins give the scheduler two independent Horner
// chains.  Static facts (ptxas, L = 5): at the 128-register budget this needs no spills and the chains stay
// interleaved (longest run of dependent DFMA: 5 / 2 in the two passes, against 13 in the one-bin form at 80 registers),
// 8 LDS + 77 FP64-pipe instructions per two (bin, neuron) items instead of 14 + 82 -- at 2 CTAs per SM instead of 3.
// Here p.tpb is the number of threads per bin PAIR (estep_seg.cu).
template <int LT, int STAGE>
__device__ __forceinline__ void rate_pass_two_bins(const SegArgs &p, const Smem<LT> &s) {
    const int tid = threadIdx.x;
    const int tp = tid / p.tpb, k = tid - tp * p.tpb;
    const int t = 2 * tp;
    if (t < p.W) {
        const int N = p.N;
        const bool two = t + 1 < p.W;
        const int t1 = two ? t + 1 : t;            // an odd last bin is computed twice and stored once
        double muA[LT], vA[LT], accA[LT], muB[LT], vB[LT], accB[LT];
#pragma unroll
        for (int l = 0; l < LT; ++l) {
            muA[l] = s.mu[t * LT + l];
            vA[l] = s.v[t * LT + l];
            muB[l] = s.mu[t1 * LT + l];
            vB[l] = s.v[t1 * LT + l];
            accA[l] = accB[l] = 0.0;
        }
        const double2 *aa = (const double2 *)s.a;
        const double2 *bb = (const double2 *)s.b;
        const uint8_t *yA = s.ys + t * N, *yB = s.ys + t1 * N;
        for (int n = k; n < N; n += p.tpb) {
            double al[LT], sq[LT];
            const double bn = bb[n].x;
            double etaA = bn, etaB = bn, hA = 0.0, hB = 0.0;
#pragma unroll
            for (int l = 0; l < LT; ++l) {
                const double2 pp = load_a(aa, l * N + n);
                al[l] = pp.x;
                sq[l] = pp.y;
                etaA = fma(muA[l], pp.x, etaA);
                etaB = fma(muB[l], pp.x, etaB);
                hA = fma(vA[l], pp.y, hA);
                hB = fma(vB[l], pp.y, hB);
            }
            double rA, rB;
            trunc_exp2(fma(0.5, hA, etaA), fma(0.5, hB, etaB), rA, rB);
            const double cA = (STAGE == 1) ? (double)yA[n] - rA : rA;
            const double cB = (STAGE == 1) ? (double)yB[n] - rB : rB;
#pragma unroll
            for (int l = 0; l < LT; ++l) {
                const double x = (STAGE == 1) ? al[l] : sq[l];
                accA[l] = fma(cA, x, accA[l]);
                accB[l] = fma(cB, x, accB[l]);
            }
        }
#pragma unroll
        for (int l = 0; l < LT; ++l) {
            s.part[(k * p.W + t) * LT + l] = accA[l];
            if (two) s.part[(k * p.W + t + 1) * LT + l] = accB[l];
        }
    }
}
#endif

// Rate pass on the FP64 tensor path (all channels Poisson, counts staged as uint8): warp i owns the 8 bins of row tile
// i and walks over the column tiles of 8 neurons.  Per tile pair:
//   x[t][n] = b_n + sum_l mu_tl a_ln + 1/2 sum_l v_tl a_ln^2   as ONE contraction over k = (mu | v | 1) against the
//             operand Bx = (a ; a^2 / 2 ; b), KS = ceil((2L + 1) / 4) DMMA; the accumulator tile hands every lane two
//             (bin, neuron) entries -> two interleaved exponentials;
//   acc[t][l] += sum_n coef[t][n] Bx[l or L + l][n]  (coef = y - rate or rate): the accumulator entries ARE the A
//             operand of this second contraction when its k index is taken as n = 8 j + 2 (lane % 4) + {0, 1} -- the
//             sum over neurons does not care about the order, so no shuffle is needed; the B operand is one 128-bit load.
// 3 + 2 DMMA and 22 FP64-pipe instructions per 64 (bin, neuron) entries at L = 5 against ~70 FP64-pipe instructions
// per entry pair in the scalar form: the contractions move to the tensor pipe and run beside the exponentials.  The
// row tile's output is complete inside the warp: no partial sums in shared memory, one barrier per pass instead of two.
// trunc_exp2 (common.cuh) for the tensor-path rate passes: bit-identical values, fewer issue slots.  On sm_100a an FP64
// instruction holds the issue port of its sub-partition for two cycles and a DMMA for sixteen, every other instruction
// for one, and nothing overlaps (scripts/mb/mb_rate_pass.cu, mb_fp64_ops.cu: a column-tile step costs the SUM of its
// slots), so the integer and select instructions around the polynomial are not free:
//   * the clamps test the high word (x >= 10 <=> hi(x) >= hi(10.0) as signed integers, x < -708 <=> hi(x) > hi(-708.0)
//     as unsigned ones, up to values within 1e-13 below -708 which need no clamp): 3 slots instead of 4 each;
//   * the table 2^(j/32) is read from shared memory (2 slots instead of 5 for the generic-address __ldg);
//   * no lower bound on the exponent: x >= -708 already implies (m >> 5) >= -1022.
__device__ __forceinline__ void trunc_exp2_tile(double x0, double x1, const double *tab, double &e0, double &e1) {
    const int h0 = __double2hiint(x0), h1 = __double2hiint(x1);
    x0 = h0 >= 0x40240000 ? 10.0 : x0;
    x1 = h1 >= 0x40240000 ? 10.0 : x1;
    x0 = (unsigned)h0 > 0xC0862000u ? -708.0 : x0;
    x1 = (unsigned)h1 > 0xC0862000u ? -708.0 : x1;
    const double shift = 6755399441055744.0;
    const double m0 = fma(x0, VLGP_EXP_INV, shift), m1 = fma(x1, VLGP_EXP_INV, shift);
    const int i0 = __double2loint(m0), i1 = __double2loint(m1);
    const double tj0 = tab[i0 & 31], tj1 = tab[i1 & 31];
    const double t0 = m0 - shift, t1 = m1 - shift;
    double r0 = fma(t0, -VLGP_EXP_HI, x0), r1 = fma(t1, -VLGP_EXP_HI, x1);
    r0 = fma(t0, -VLGP_EXP_LO, r0);
    r1 = fma(t1, -VLGP_EXP_LO, r1);
    double p0 = VLGP_EXP_C[0], p1 = VLGP_EXP_C[0];
#pragma unroll
    for (int k = 1; k < 6; ++k) {
        p0 = fma(p0, r0, VLGP_EXP_C[k]);
        p1 = fma(p1, r1, VLGP_EXP_C[k]);
    }
    p0 *= r0;
    p1 *= r1;
    const double q0 = fma(tj0, p0, tj0), q1 = fma(tj1, p1, tj1);
    e0 = __hiloint2double(__double2hiint(q0) + ((i0 >> 5) << 20), __double2loint(q0));
    e1 = __hiloint2double(__double2hiint(q1) + ((i1 >> 5) << 20), __double2loint(q1));
}

// The column-tile loop of one row tile (warp wid, 8 wid < W): on return lane (r, q) holds, per output tile o, the sums over
// neurons for bin 8 wid + r and latents 8 o + 2 q, 8 o + 2 q + 1 (STAGE 2: of a^2 / 2, to be doubled by the caller).
// FUSED: the y term of stage 1 is left out (the fused pipeline keeps y a_l' in s.ya and subtracts the sum of rate x a_l),
// the exponential reads its table from shared memory.
template <int LT, int STAGE, bool FUSED = false>
__device__ __forceinline__ void rate_tiles_core(int W, int N, int NP, const double *Bx, const double *smu, const double *sv,
                                                const double *etab, const uint8_t *ys, Tile (&acc)[(LT + 7) / 8]) {
    constexpr int KS = (2 * LT + 1 + 3) / 4;        // k4 steps of the x contraction
    constexpr int NOT = (LT + 7) / 8;               // output tiles of 8 latents
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, r = lane >> 2, q = lane & 3;
    {
        const int t = 8 * wid + r;
        const bool tin = t < W;
        double afr[KS];
#pragma unroll
        for (int kk = 0; kk < KS; ++kk) {
            const int k = 4 * kk + q;
            double val = 0.0;
            if (tin) {
                if (k < LT) val = smu[t * LT + k];
                else if (k < 2 * LT) val = sv[t * LT + k - LT];
                else if (k == 2 * LT) val = 1.0;
            }
            afr[kk] = val;
        }
#pragma unroll
        for (int o = 0; o < NOT; ++o) acc[o].x = acc[o].y = 0.0;
        const uint8_t *yrow = ys + (tin ? t : 0) * N;
        const int brow = (STAGE == 1 ? 0 : LT) + r;
        // (build option VLGP_ESTEP_RATE_ILP2: two column tiles per step, four exponentials in flight per lane)
        const int nct = NP >> 3;
        auto tile_pair = [&](int j, double ea, double eb) {
            const int n0 = 8 * j + 2 * q;
            double c0 = ea, c1 = eb;
            if (STAGE == 1 && !FUSED) {
                const double y0 = (tin && n0 < N) ? (double)yrow[n0] : 0.0;
                const double y1 = (tin && n0 + 1 < N) ? (double)yrow[n0 + 1] : 0.0;
                c0 = y0 - ea;
                c1 = y1 - eb;
            }
#pragma unroll
            for (int o = 0; o < NOT; ++o) {
                double2 b2 = make_double2(0.0, 0.0);
                if (8 * o + r < LT) b2 = *reinterpret_cast<const double2 *>(Bx + (brow + 8 * o) * NP + n0);
                dmma(acc[o], c0, b2.x);
                dmma(acc[o], c1, b2.y);
            }
        };
        int j = 0;
#ifdef VLGP_ESTEP_RATE_ILP2      // measured on B200 (profiles/ab_estep_variants_r2.txt): 0.5 ms SLOWER per launch than one tile at a time
        for (; j + 1 < nct; j += 2) {
            Tile xa{0.0, 0.0}, xb{0.0, 0.0};
#pragma unroll
            for (int kk = 0; kk < KS; ++kk) {
                dmma(xa, afr[kk], Bx[(4 * kk + q) * NP + 8 * j + r]);
                dmma(xb, afr[kk], Bx[(4 * kk + q) * NP + 8 * j + 8 + r]);
            }
            double e0, e1, e2, e3;
            trunc_exp4(xa.x, xa.y, xb.x, xb.y, e0, e1, e2, e3);
            tile_pair(j, e0, e1);
            tile_pair(j + 1, e2, e3);
        }
#endif
        for (; j < nct; ++j) {
            Tile x{0.0, 0.0};
#pragma unroll
            for (int kk = 0; kk < KS; ++kk) dmma(x, afr[kk], Bx[(4 * kk + q) * NP + 8 * j + r]);
            double e0, e1;
            if (FUSED) trunc_exp2_tile(x.x, x.y, etab, e0, e1);
            else trunc_exp2(x.x, x.y, e0, e1);
            tile_pair(j, e0, e1);
        }
    }
}

template <int LT, int STAGE, bool FUSED = false>
__device__ __forceinline__ void rate_tiles(const SegArgs &p, const Smem<LT> &s, Tile (&acc)[(LT + 7) / 8]) {
    rate_tiles_core<LT, STAGE, FUSED>(p.W, p.N, p.np, s.a, s.mu, s.v, s.etab, s.ys, acc);
}

// Single-precision rate pass for the fused pipeline (FAST == 2, BASELINE.json configs[2] "fp32"): the linear predictor,
// the exponential and the sums over neurons in FP32 on the FMA pipe (one issue slot per instruction instead of two, no
// tensor-path padding); everything that is ill-conditioned -- the Gram matrices, their inverses, the variances, the mean
// step -- stays FP64, and so do the state arrays and y a_l'.  Lane (r, q) of a row-tile warp takes bin 8 wid + r and the
// neurons n = q (mod 4): the four lanes of a bin read consecutive (a, a^2 / 2) pairs, the eight bins of the warp the
// same ones (shared-memory broadcast).  Returns the sums in the accumulator layout of rate_tiles.
template <int LT, int STAGE>
__device__ __forceinline__ void rate_tiles_f32(const SegArgs &p, const Smem<LT> &s, Tile (&acc)[(LT + 7) / 8]) {
    constexpr int NOT = (LT + 7) / 8;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, r = lane >> 2, q = lane & 3;
    const int N = p.N, t = 8 * wid + r;
    const bool tin = t < p.W;
    float mu[LT], hv[LT], sum[LT];
#pragma unroll
    for (int l = 0; l < LT; ++l) {
        mu[l] = tin ? (float)s.mu[t * LT + l] : 0.f;
        hv[l] = tin ? (float)s.v[t * LT + l] : 0.f;
        sum[l] = 0.f;
    }
    const float2 *af = (const float2 *)s.a;            // [l * N + n] : (a, a^2 / 2)
    const float *bf = (const float *)(af + LT * N);    // [n] : bias
    int n = q;
    for (; n + 4 < N; n += 8) {                        // two neurons in flight
        float x0 = bf[n], x1 = bf[n + 4];
        float2 c0[LT], c1[LT];
#pragma unroll
        for (int l = 0; l < LT; ++l) {
            c0[l] = af[l * N + n];
            c1[l] = af[l * N + n + 4];
            x0 = fmaf(mu[l], c0[l].x, x0);
            x1 = fmaf(mu[l], c1[l].x, x1);
            x0 = fmaf(hv[l], c0[l].y, x0);
            x1 = fmaf(hv[l], c1[l].y, x1);
        }
        const float e0 = expf(fminf(x0, 10.f)), e1 = expf(fminf(x1, 10.f));
#pragma unroll
        for (int l = 0; l < LT; ++l) {
            sum[l] = fmaf(e0, STAGE == 1 ? c0[l].x : c0[l].y, sum[l]);
            sum[l] = fmaf(e1, STAGE == 1 ? c1[l].x : c1[l].y, sum[l]);
        }
    }
    if (n < N) {
        float x0 = bf[n];
        float2 c0[LT];
#pragma unroll
        for (int l = 0; l < LT; ++l) {
            c0[l] = af[l * N + n];
            x0 = fmaf(mu[l], c0[l].x, x0);
            x0 = fmaf(hv[l], c0[l].y, x0);
        }
        const float e0 = expf(fminf(x0, 10.f));
#pragma unroll
        for (int l = 0; l < LT; ++l) sum[l] = fmaf(e0, STAGE == 1 ? c0[l].x : c0[l].y, sum[l]);
    }
#pragma unroll
    for (int l = 0; l < LT; ++l) {
        sum[l] += __shfl_xor_sync(FULL, sum[l], 1);
        sum[l] += __shfl_xor_sync(FULL, sum[l], 2);
    }
#pragma unroll
    for (int o = 0; o < NOT; ++o) {
        float ax = 0.f, ay = 0.f;
#pragma unroll
        for (int l = 0; l < LT; ++l) {
            if (l == 8 * o + 2 * q) ax = sum[l];
            if (l == 8 * o + 2 * q + 1) ay = sum[l];
        }
        acc[o].x = (double)ax;
        acc[o].y = (double)ay;
    }
}

// s.ya[t][l] = sum_n y[t][n] a[l][n] for this warp's row tile (fused path, once per segment; counts read in place)
template <int LT, bool F32>
__device__ __forceinline__ void ya_tiles(const SegArgs &p, const Smem<LT> &s, const uint8_t *yseg) {
    constexpr int NOT = (LT + 7) / 8;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, r = lane >> 2, q = lane & 3;
    const int W = p.W, N = p.N, NP = p.np;
    const int t = 8 * wid + r;
    const bool tin = t < W;
    const uint8_t *yrow = yseg + (size_t)(tin ? t : 0) * N;
    Tile acc[NOT];
#pragma unroll
    for (int o = 0; o < NOT; ++o) acc[o].x = acc[o].y = 0.0;
    for (int j = 0; j < (NP >> 3); ++j) {
        const int n0 = 8 * j + 2 * q;
        const double y0 = (tin && n0 < N) ? (double)yrow[n0] : 0.0;
        const double y1 = (tin && n0 + 1 < N) ? (double)yrow[n0 + 1] : 0.0;
#pragma unroll
        for (int o = 0; o < NOT; ++o) {
            double2 b2 = make_double2(0.0, 0.0);
            if (8 * o + r < LT) {
                if (F32) {                             // FP32 mode keeps no FP64 loading in SMEM: read it in place
                    if (n0 < N) b2.x = p.pa[(r + 8 * o) * N + n0].x;
                    if (n0 + 1 < N) b2.y = p.pa[(r + 8 * o) * N + n0 + 1].x;
                } else {
                    b2 = *reinterpret_cast<const double2 *>(s.a + (r + 8 * o) * NP + n0);
                }
            }
            dmma(acc[o], y0, b2.x);
            dmma(acc[o], y1, b2.y);
        }
    }
    if (tin) {
#pragma unroll
        for (int o = 0; o < NOT; ++o) {
            const int l0 = 8 * o + 2 * q;
            if (l0 < LT) s.ya[t * LT + l0] = acc[o].x;
            if (l0 + 1 < LT) s.ya[t * LT + l0 + 1] = acc[o].y;
        }
    }
}

template <int LT, int STAGE>
__device__ __forceinline__ void rate_pass_dmma(const SegArgs &p, const Smem<LT> &s) {
    constexpr int NOT = (LT + 7) / 8;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, r = lane >> 2, q = lane & 3;
    const int W = p.W;
    if (8 * wid < W) {
        const int t = 8 * wid + r;
        const bool tin = t < W;
        Tile acc[NOT];
        rate_tiles<LT, STAGE>(p, s, acc);
        double *out = (STAGE == 1) ? s.ra : s.w;
        const double sc = (STAGE == 1) ? 1.0 : 2.0;          // Bx holds a^2 / 2
        if (tin) {
#pragma unroll
            for (int o = 0; o < NOT; ++o) {
                const int l0 = 8 * o + 2 * q;
                if (l0 < LT) out[t * LT + l0] = sc * acc[o].x;
                if (l0 + 1 < LT) out[t * LT + l0 + 1] = sc * acc[o].y;
            }
        }
    }
    __syncthreads();
}

template <int LT, int STAGE, int FAST>
__device__ __forceinline__ void rate_pass(const SegArgs &p, const Smem<LT> &s, int64_t bin0) {
#ifndef VLGP_ESTEP_SCALAR_RATE_PASS
    if (FAST) {
        rate_pass_dmma<LT, STAGE>(p, s);
        return;
    }
#endif
    const int tid = threadIdx.x;
#ifdef VLGP_ESTEP_TWO_BINS
    if (FAST) {
        rate_pass_two_bins<LT, STAGE>(p, s);
    } else
    for (int t = 2 * (tid / p.tpb), k = tid % p.tpb, tend = min(t + 2, p.W); t < tend; ++t) {
#else
    const int t = tid / p.tpb, k = tid - t * p.tpb;
    if (t < p.W) {
#endif
        const int N = p.N;
        double mu_t[LT], v_t[LT], acc[LT];
#pragma unroll
        for (int l = 0; l < LT; ++l) {
            mu_t[l] = s.mu[t * LT + l];
            v_t[l] = s.v[t * LT + l];
            acc[l] = 0.0;
        }
        const double2 *aa = (const double2 *)s.a;
        const double2 *bb = (const double2 *)s.b;
        // neurons are interleaved over the tpb threads of a bin (n = k, k + tpb, ...): at every step the threads of a
        // warp touch tpb CONSECUTIVE neurons, so the (a, a^2), bias and count loads are bank-conflict-free broadcasts
        if (FAST) {
            // all channels Poisson, counts staged as uint8: straight-line code, two neurons in flight per iteration
            const uint8_t *yrow = s.ys + t * N;
            int n = k;
            for (; n + p.tpb < N; n += 2 * p.tpb) {
                const int n2 = n + p.tpb;
                double al0[LT], al1[LT];
                double eta0 = bb[n].x, eta1 = bb[n2].x, h0 = 0.0, h1 = 0.0;
#pragma unroll
                for (int l = 0; l < LT; ++l) {
                    const double2 p0 = load_a(aa, l * N + n), p1 = load_a(aa, l * N + n2);
                    eta0 = fma(mu_t[l], p0.x, eta0);
                    eta1 = fma(mu_t[l], p1.x, eta1);
                    h0 = fma(v_t[l], p0.y, h0);
                    h1 = fma(v_t[l], p1.y, h1);
                    al0[l] = (STAGE == 1) ? p0.x : p0.y;
                    al1[l] = (STAGE == 1) ? p1.x : p1.y;
                }
                double r0, r1;
                trunc_exp2(fma(0.5, h0, eta0), fma(0.5, h1, eta1), r0, r1);
                const double c0 = (STAGE == 1) ? (double)yrow[n] - r0 : r0;
                const double c1 = (STAGE == 1) ? (double)yrow[n2] - r1 : r1;
#pragma unroll
                for (int l = 0; l < LT; ++l) acc[l] = fma(c1, al1[l], fma(c0, al0[l], acc[l]));
            }
            if (n < N) {
                double al0[LT];
                double eta0 = bb[n].x, h0 = 0.0;
#pragma unroll
                for (int l = 0; l < LT; ++l) {
                    const double2 p0 = load_a(aa, l * N + n);
                    eta0 = fma(mu_t[l], p0.x, eta0);
                    h0 = fma(v_t[l], p0.y, h0);
                    al0[l] = (STAGE == 1) ? p0.x : p0.y;
                }
                const double r0 = trunc_exp(fma(0.5, h0, eta0));
                const double c0 = (STAGE == 1) ? (double)yrow[n] - r0 : r0;
#pragma unroll
                for (int l = 0; l < LT; ++l) acc[l] = fma(c0, al0[l], acc[l]);
            }
        } else {
#pragma unroll 2
            for (int n = k; n < N; n += p.tpb) {
                const double2 bn = bb[n];                       // (bias, 1 / noise)
                double al[LT], eta = bn.x, h = 0.0;
#pragma unroll
                for (int l = 0; l < LT; ++l) {
                    const double2 p2 = load_a(aa, l * N + n);  // (a, a^2)
                    eta = fma(mu_t[l], p2.x, eta);
                    h = fma(v_t[l], p2.y, h);
                    al[l] = (STAGE == 1) ? p2.x : p2.y;
                }
                const bool pois = s.pois[n] != 0;
                const double rate = trunc_exp(fma(0.5, h, eta));       // computed for every channel: no branch
                double coef;
                if (STAGE == 1) {
                    const double yv = p.ydtype == VLGP_Y_U8 ? (double)s.ys[t * N + n]
                                                           : ((const double *)p.y)[(bin0 + t) * N + n];
                    coef = pois ? yv - rate : (yv - eta) * bn.y;
                } else {
                    coef = pois ? rate : bn.y;
                }
#pragma unroll
                for (int l = 0; l < LT; ++l) acc[l] = fma(coef, al[l], acc[l]);
            }
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

// ---- per-latent r x r work.  Gram, variance and the mean-step mat-vecs are flattened over ALL threads of the CTA as
// (latent, entry) items so no warp idles; only the short symmetric sweep runs one warp per latent. --------------------

// idx -> (latent, lower-triangle entry (i, j)) of the per-latent Gram matrices; pairoff[l] = first item of latent l
__device__ __forceinline__ void decode_pair(const SegArgs &p, int LT, int idx, int &l, int &i, int &j) {
    l = 0;
    while (l + 1 < LT && idx >= p.pairoff[l + 1]) ++l;
    tri_decode(idx - p.pairoff[l], i, j);
}

// M_l <- I + G_l' diag(w_l) G_l for every latent (both triangles written)
template <int LT>
__device__ __forceinline__ void gram_all(const SegArgs &p, const Smem<LT> &s) {
    const int W = p.W;
    for (int idx = threadIdx.x; idx < p.pair_total; idx += NT) {
        int l, i, j;
        decode_pair(p, LT, idx, l, i, j);
        const int ldg = p.ldg[l];
        const double *g = s.Gs + p.goff[l];
        const double *wl = s.w + l;
        double c0 = 0.0, c1 = 0.0;
        int t = 0;
        for (; t + 1 < W; t += 2) {
            c0 = fma(g[t * ldg + i] * wl[t * LT], g[t * ldg + j], c0);
            c1 = fma(g[(t + 1) * ldg + i] * wl[(t + 1) * LT], g[(t + 1) * ldg + j], c1);
        }
        if (t < W) c0 = fma(g[t * ldg + i] * wl[t * LT], g[t * ldg + j], c0);
        const double c = c0 + c1 + ((i == j) ? 1.0 : 0.0);
        double *M = s.Mi + p.moff[l];
        M[i * p.ldm[l] + j] = c;
        M[j * p.ldm[l] + i] = c;
    }
}

// In-place symmetric sweep of one latent's matrix by one warp: M <- -M^-1.  Returns false (warp-uniform) if a pivot is
// not positive (the matrix is not positive definite).
__device__ __forceinline__ bool warp_sweep(double *M, int ldm, int nc, double *colk) {
    const int lane = threadIdx.x & 31;
    const float inv_nc = 1.0f / (float)nc;
    const int nn = nc * nc;
    for (int k = 0; k < nc; ++k) {
        for (int i = lane; i < nc; i += 32) colk[i] = M[i * ldm + k];
        __syncwarp();
        const double d = colk[k];
        if (!(d > 0.0)) return false;
        const double pinv = fast_rcp(d);
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

// One latent on the FP64 tensor path, by one warp, without block barriers: Gram matrix I + G' diag(w) G accumulated by
// DMMA straight from the compact factor in SMEM (A operand = w-scaled column fragment, B operand = the same fragment
// unscaled), blocked symmetric sweep on the NB x NB tile matrix in registers (dmma.cuh, as in the H-step kernel),
// -Minv stored to SMEM (both triangles, ld = 8 NB + 4: conflict-free fragment reads), then the marginal variances
// v_t = G_t Minv G_t' as (G row tile) x Minv by DMMA and a quad reduction.  Index math validated by a 32-lane NumPy
// emulation (scripts/dmma_estep_emulation.py).
template <int LT, int NB>
__device__ __forceinline__ bool factor_variance_dmma_nb(const SegArgs &p, const Smem<LT> &s, int l, bool do_var) {
    constexpr int NTL = NB * (NB + 1) / 2;
    const int lane = threadIdx.x & 31, r = lane >> 2, c0 = 2 * (lane & 3);
    const int W = p.W, nc = p.nc[l], ldg = p.ldg[l];
    const double *G = s.Gs + p.goff[l];
    const double *wl = s.w + l;
    Tile A[NTL];
#pragma unroll
    for (int t = 0; t < NTL; ++t) A[t].x = A[t].y = 0.0;
    for (int k = 0; 4 * k < W; ++k) {
        const int t = 4 * k + (lane & 3);
        const bool tin = t < W;
        const double wt = tin ? wl[t * LT] : 0.0;
        double g[NB];
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            const int c = 8 * b + r;
            g[b] = (tin && c < nc) ? G[t * ldg + c] : 0.0;
        }
#pragma unroll
        for (int i = 0; i < NB; ++i) {
            const double gw = g[i] * wt;
#pragma unroll
            for (int j = 0; j <= i; ++j) dmma(A[tix(i, j)], gw, g[j]);
        }
    }
#pragma unroll
    for (int i = 0; i < NB; ++i) {           // + I (also on the padding, so the sweep stays well defined)
        A[tix(i, i)].x += (r == c0) ? 1.0 : 0.0;
        A[tix(i, i)].y += (r == c0 + 1) ? 1.0 : 0.0;
    }
    bool ok = true;
#pragma unroll
    for (int kb = 0; kb < NB; ++kb) {
        Tile P = A[tix(kb, kb)];
        ok = tile_spd_inverse(P, lane) && ok;
        const double Pt0 = tform(P, 0, lane), Pt1 = tform(P, 1, lane);
        const double Pn0 = nform(P, 0, lane), Pn1 = nform(P, 1, lane);
        double V0[NB], V1[NB];
#pragma unroll
        for (int m = 0; m < NB; ++m) {
            if (m == kb) continue;
            if (m > kb) {
                V0[m] = nform(A[tix(m, kb)], 0, lane);
                V1[m] = nform(A[tix(m, kb)], 1, lane);
            } else {
                V0[m] = tform(A[tix(kb, m)], 0, lane);
                V1[m] = tform(A[tix(kb, m)], 1, lane);
            }
        }
#pragma unroll
        for (int m = 0; m < NB; ++m) {
            if (m == kb) continue;
            Tile T{0.0, 0.0};
            if (m > kb) {
                dmma(T, V0[m], Pt0);
                dmma(T, V1[m], Pt1);
                A[tix(m, kb)] = T;
            } else {
                dmma(T, Pn0, V0[m]);
                dmma(T, Pn1, V1[m]);
                A[tix(kb, m)] = T;
            }
        }
#pragma unroll
        for (int i = 0; i < NB; ++i) {
            if (i == kb) continue;
            double T0, T1;
            if (i > kb) {
                T0 = -nform(A[tix(i, kb)], 0, lane);
                T1 = -nform(A[tix(i, kb)], 1, lane);
            } else {
                T0 = -tform(A[tix(kb, i)], 0, lane);
                T1 = -tform(A[tix(kb, i)], 1, lane);
            }
#pragma unroll
            for (int j = 0; j <= i; ++j) {
                if (j == kb) continue;
                dmma(A[tix(i, j)], T0, V0[j]);
                dmma(A[tix(i, j)], T1, V1[j]);
            }
        }
        A[tix(kb, kb)].x = -P.x;
        A[tix(kb, kb)].y = -P.y;
    }
    // -Minv -> SMEM, row-major, both triangles
    constexpr int LDM = 8 * NB + 4;
    double *M = s.Mi + p.moff[l];
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
    __syncwarp();
    if (!ok) return false;
    if (do_var) {
        double Bop[2 * NB][NB];                       // Minv as B operand: [c = 4 k + lane%4][j = 8 jt + lane/4]
#pragma unroll
        for (int k = 0; k < 2 * NB; ++k)
#pragma unroll
            for (int jt = 0; jt < NB; ++jt) Bop[k][jt] = M[(4 * k + (lane & 3)) * LDM + 8 * jt + r];
        for (int tt = 0; 8 * tt < W; ++tt) {
            const int trow = 8 * tt + r;
            const bool tin = trow < W;
            const double *grow = G + trow * ldg;
            double aop[2 * NB];
#pragma unroll
            for (int k = 0; k < 2 * NB; ++k) {
                const int c = 4 * k + (lane & 3);
                aop[k] = (tin && c < nc) ? grow[c] : 0.0;
            }
            double acc = 0.0;
#pragma unroll
            for (int jt = 0; jt < NB; ++jt) {
                Tile T{0.0, 0.0};
                // k4 steps whose four columns all lie beyond nc multiply the zero padding of the G rows: left out
                // (warp-uniform test; nc = 10 at NB = 2 saves one DMMA in four)
#pragma unroll
                for (int k = 0; k < 2 * NB; ++k)
                    if (4 * k < nc) dmma(T, aop[k], Bop[k][jt]);
                const int c = 8 * jt + c0;
                const double g0 = (tin && c < nc) ? grow[c] : 0.0;
                const double g1 = (tin && c + 1 < nc) ? grow[c + 1] : 0.0;
                acc = fma(T.x, g0, acc);
                acc = fma(T.y, g1, acc);
            }
            acc += __shfl_xor_sync(FULL, acc, 1);
            acc += __shfl_xor_sync(FULL, acc, 2);
            if ((lane & 3) == 0 && tin) s.v[trow * LT + l] = -acc;
        }
    }
    return true;
}

template <int LT, int NBMAX>
__device__ __forceinline__ bool factor_variance_dmma(const SegArgs &p, const Smem<LT> &s, int l, bool do_var) {
    // The kernel is instantiated twice: NBMAX = 2 serves nc <= 16 (the steady-state regime, nc = 6..12) at ~80
    // registers / 3 CTAs per SM; NBMAX = 4 serves nc <= 32 (omega near its upper bound: the first EM iterations) at
    // 128 registers / 2 CTAs per SM.  Keeping the 24- and 32-column code out of the first instantiation is what keeps
    // the rate passes at full occupancy.
    if (p.nc[l] <= 8) return factor_variance_dmma_nb<LT, 1>(p, s, l, do_var);
    if (NBMAX == 2 || p.nc[l] <= 16) return factor_variance_dmma_nb<LT, 2>(p, s, l, do_var);
    if (p.nc[l] <= 24) return factor_variance_dmma_nb<LT, (NBMAX >= 3 ? 3 : 2)>(p, s, l, do_var);
    return factor_variance_dmma_nb<LT, (NBMAX >= 4 ? 4 : 2)>(p, s, l, do_var);
}

// ---- fused pipeline (tensor-path rate passes + tensor-path factorisation) ---------------------------------------------
// An iteration of the unfused pipeline spends more time in the r x r phases than in the rate passes (ncu, config 2: 35 %
// of the warp samples in the two rate passes, 22 % in the five barrier-separated mat-vec stages of the mean step, 35-40 %
// in the one-warp-per-latent factor / variance phase during which three of the eight warps idle).  The fused pipeline
// gives the row-tile warps of the rate passes the r x r work that belongs to their own 8 bins and shortens the rest:
//   mean step     delta = u - G Minv G' W u with u = G G' ra - mu equals, exactly,  G Minv G' (ra + w o mu) - mu
//                 (Minv = (I + G'WG)^-1: G'Wu = (Minv^-1 - I) p - q with p = G' ra, q = G'(w o mu), so
//                 Minv G'Wu = p - Minv (p + q)): three products instead of five.  The first one, s = G' z, is summed per
//                 row tile by the warp that just produced ra (phase A), the small r x r product m = Minv s is spread
//                 over the CTA (phase B), and G m for a warp's own bins opens its second rate pass (phase C);
//   variance      v_t = G_t Minv G_t' of a warp's own bins opens its FIRST rate pass of the next iteration (phase A),
//                 (row tile) x Minv by DMMA as before, 7 warps x L latents instead of L warps x 7 row tiles;
//   factor        Gram + blocked sweep only, still one warp per latent (phase D).
// Four block barriers per iteration instead of eight.
template <int LT, int NB>
__device__ __forceinline__ void variance_rows_nb(const SegArgs &p, const Smem<LT> &s, int l) {
    constexpr int LDM = 8 * NB + 4;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, r = lane >> 2, q = lane & 3, c0 = 2 * q;
    const int nc = p.nc[l], ldg = p.ldg[l];
    const double *M = s.Mi + p.moff[l];
    const int trow = 8 * wid + r;
    const bool tin = trow < p.W;
    const double *grow = s.Gs + p.goff[l] + (tin ? trow : 0) * ldg;
    double aop[2 * NB];
#pragma unroll
    for (int k = 0; k < 2 * NB; ++k) {
        const int c = 4 * k + q;
        aop[k] = (tin && c < nc) ? grow[c] : 0.0;
    }
    // g' Minv g over the lower block triangle only (Minv is symmetric): per column block the diagonal block once and the
    // blocks below it twice -- NB (NB + 1) DMMA per row tile instead of 2 NB^2
    double acc = 0.0;
#pragma unroll
    for (int jt = 0; jt < NB; ++jt) {
        Tile Td{0.0, 0.0}, To{0.0, 0.0};
#pragma unroll
        for (int k = 2 * jt; k < 2 * jt + 2; ++k)
            if (4 * k < nc) dmma(Td, aop[k], M[(4 * k + q) * LDM + 8 * jt + r]);
#pragma unroll
        for (int k = 2 * jt + 2; k < 2 * NB; ++k)
            if (4 * k < nc) dmma(To, aop[k], M[(4 * k + q) * LDM + 8 * jt + r]);
        const int c = 8 * jt + c0;
        const double g0 = (tin && c < nc) ? grow[c] : 0.0;
        const double g1 = (tin && c + 1 < nc) ? grow[c + 1] : 0.0;
        acc = fma(fma(2.0, To.x, Td.x), g0, acc);
        acc = fma(fma(2.0, To.y, Td.y), g1, acc);
    }
    acc += __shfl_xor_sync(FULL, acc, 1);
    acc += __shfl_xor_sync(FULL, acc, 2);
    if (q == 0 && tin) s.v[trow * LT + l] = -acc;
}

template <int LT, int NBMAX>
__device__ __forceinline__ void variance_rows(const SegArgs &p, const Smem<LT> &s, int l) {
    if (p.nc[l] <= 8) variance_rows_nb<LT, 1>(p, s, l);
    else if (NBMAX == 2 || p.nc[l] <= 16) variance_rows_nb<LT, 2>(p, s, l);
    else if (p.nc[l] <= 24) variance_rows_nb<LT, (NBMAX >= 3 ? 3 : 2)>(p, s, l);
    else variance_rows_nb<LT, (NBMAX >= 4 ? 4 : 2)>(p, s, l);
}

__device__ __forceinline__ int col_latent(const SegArgs &p, int LT, int idx) {
    int l = 0;
    while (l + 1 < LT && idx >= p.coloff[l + 1]) ++l;
    return l;
}

// Phase A: [variance of the own bins from the previous factorisation] -> rate pass 1 -> z = ra + w o mu -> this row
// tile's part of s_l = G_l' z_l for every latent.  Ends with a block barrier.
template <int LT, int NBMAX, bool F32>
__device__ __forceinline__ void fused_a(const SegArgs &p, const Smem<LT> &s, const int *bad, bool do_var) {
    constexpr int NOT = (LT + 7) / 8;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, r = lane >> 2, q = lane & 3;
    const int W = p.W;
    if (8 * wid < W) {
        if (do_var) {
            for (int l = 0; l < LT; ++l)
                if (!bad[l]) variance_rows<LT, NBMAX>(p, s, l);      // a failed solve keeps v (vlgp/core.py:112)
            __syncwarp();
        }
        Tile acc[NOT];
        if (F32) rate_tiles_f32<LT, 1>(p, s, acc);
        else if (!(p.skip & 1)) rate_tiles<LT, 1, true>(p, s, acc);
        else
            for (int o = 0; o < NOT; ++o) acc[o].x = acc[o].y = 0.0;
        const int t = 8 * wid + r;
        if (t < W) {                                  // z = (y - rate) a_l' + w o mu
#pragma unroll
            for (int o = 0; o < NOT; ++o) {
                const int l0 = 8 * o + 2 * q;
                if (l0 < LT) s.ra[t * LT + l0] = fma(s.w[t * LT + l0], s.mu[t * LT + l0], s.ya[t * LT + l0] - acc[o].x);
                if (l0 + 1 < LT)
                    s.ra[t * LT + l0 + 1] = fma(s.w[t * LT + l0 + 1], s.mu[t * LT + l0 + 1], s.ya[t * LT + l0 + 1] - acc[o].y);
            }
        }
        __syncwarp();
        const int rows = min(8, W - 8 * wid);
        double *part = s.part + wid * p.col_total;
        for (int idx = lane; idx < p.col_total; idx += 32) {
            const int l = col_latent(p, LT, idx), ldg = p.ldg[l];
            const double *g = s.Gs + p.goff[l] + 8 * wid * ldg + (idx - p.coloff[l]);
            const double *z = s.ra + 8 * wid * LT + l;
            double a = 0.0;
            for (int tt = 0; tt < rows; ++tt) a = fma(g[tt * ldg], z[tt * LT], a);
            part[idx] = a;
        }
    }
    __syncthreads();
}

// Phase B: m_l = Minv_l s_l with s_l the sum of the row tiles' parts (fixed order); four lanes per entry.  M holds -Minv.
template <int LT>
__device__ __forceinline__ void fused_b(const SegArgs &p, const Smem<LT> &s) {
    const int tid = threadIdx.x, sub = tid & 3;
    const int nw = (p.W + 7) >> 3, ct = p.col_total;
    double *mvec = s.part + NWARP * ct;
    for (int base = 0; base < ct; base += NT / 4) {
        const int idx = base + (tid >> 2);
        double acc = 0.0;
        if (idx < ct) {
            const int l = col_latent(p, LT, idx), i = idx - p.coloff[l], nc = p.nc[l], ldm = p.ldm[l];
            const double *M = s.Mi + p.moff[l];
            const double *part0 = s.part + p.coloff[l];
            for (int j = sub; j < nc; j += 4) {
                double sj = 0.0;
                for (int w = 0; w < nw; ++w) sj += part0[w * ct + j];
                acc = fma(M[j * ldm + i], sj, acc);
            }
        }
        acc += __shfl_xor_sync(FULL, acc, 1);
        acc += __shfl_xor_sync(FULL, acc, 2);
        if (idx < ct && sub == 0) mvec[idx] = -acc;
    }
    __syncthreads();
}

// Phase C: delta = clip(G m - mu) on the own bins (a failed factorisation zeroes the step, vlgp/core.py:92-94), then
// rate pass 2 -> w.  Ends with a block barrier.
template <int LT, bool F32>
__device__ __forceinline__ void fused_c(const SegArgs &p, const Smem<LT> &s, const int *bad) {
    constexpr int NOT = (LT + 7) / 8;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, r = lane >> 2, q = lane & 3;
    const int W = p.W;
    if (8 * wid < W) {
        const double *mvec = s.part + NWARP * p.col_total;
        for (int idx = lane; idx < 8 * LT; idx += 32) {
            const int l = idx >> 3, t = 8 * wid + (idx & 7);
            if (t < W) {
                double d = 0.0;
                if (!bad[l]) {
                    const int nc = p.nc[l];
                    const double *g = s.Gs + p.goff[l] + t * p.ldg[l];
                    const double *mv = mvec + p.coloff[l];
                    double a = 0.0;
                    for (int j = 0; j < nc; ++j) a = fma(g[j], mv[j], a);
                    d = clipd(a - s.mu[t * LT + l], p.dmu_bound);
                }
                s.dmu[t * LT + l] = d;
                s.mu[t * LT + l] += d;
            }
        }
        __syncwarp();
        Tile acc[NOT];
        if (F32) rate_tiles_f32<LT, 2>(p, s, acc);
        else if (!(p.skip & 1)) rate_tiles<LT, 2, true>(p, s, acc);
        else
            for (int o = 0; o < NOT; ++o) acc[o].x = acc[o].y = 0.5;
        const int t = 8 * wid + r;
        if (t < W) {
#pragma unroll
            for (int o = 0; o < NOT; ++o) {
                const int l0 = 8 * o + 2 * q;
                if (l0 < LT) s.w[t * LT + l0] = 2.0 * acc[o].x;              // Bx holds a^2 / 2
                if (l0 + 1 < LT) s.w[t * LT + l0 + 1] = 2.0 * acc[o].y;
            }
        }
    }
    __syncthreads();
}

// Factorisation for every latent: on return M_l = -(I + G_l' W_l G_l)^-1, bad[l] says whether that failed (not positive
// definite), and -- if do_var -- v holds the new marginal variances of the latents that did not fail.  Ends with a
// block barrier.
template <int LT>
__device__ __forceinline__ void variance_all(const SegArgs &p, const Smem<LT> &s, const int *bad);

template <int LT, int NBMAX>
__device__ __forceinline__ void factor_all(const SegArgs &p, const Smem<LT> &s, int *bad, bool do_var) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (p.use_dmma) {
        for (int l = wid; l < LT; l += NWARP) {
            const bool ok = factor_variance_dmma<LT, NBMAX>(p, s, l, do_var);
            if (lane == 0) {
                bad[l] = ok ? 0 : 1;
                if (!ok) atomicAdd(p.flags, 1);
            }
        }
        __syncthreads();
        return;
    }
    gram_all<LT>(p, s);
    __syncthreads();
    for (int l = wid; l < LT; l += NWARP) {
        const bool ok = warp_sweep(s.Mi + p.moff[l], p.ldm[l], p.nc[l], s.vec + (size_t)l * 192);
        if (lane == 0) {
            bad[l] = ok ? 0 : 1;
            if (!ok) atomicAdd(p.flags, 1);
        }
    }
    __syncthreads();
    if (do_var) {
        variance_all<LT>(p, s, bad);
        __syncthreads();
    }
}

// v_t = G_t Minv G_t' for every (latent, bin)   (M holds -Minv)
template <int LT>
__device__ __forceinline__ void variance_all(const SegArgs &p, const Smem<LT> &s, const int *bad) {
    const int W = p.W;
    const float inv_w = 1.0f / (float)W;
    for (int idx = threadIdx.x; idx < LT * W; idx += NT) {
        const int l = (int)(((float)idx + 0.5f) * inv_w);
        const int t = idx - l * W;
        if (bad[l]) continue;                                      // failed solve: v keeps its value (core.py:112)
        const int nc = p.nc[l], ld = p.ldg[l], ldm = p.ldm[l];
        const double *g = s.Gs + p.goff[l] + t * ld;
        const double *M = s.Mi + p.moff[l];
        double acc = 0.0;
        for (int i = 0; i < nc; ++i) {
            double inner = 0.0;
            for (int j = 0; j < nc; ++j) inner = fma(M[i * ldm + j], g[j], inner);
            acc = fma(g[i], inner, acc);
        }
        s.v[t * LT + l] = -acc;
    }
}

// Newton step of the posterior mean of every latent (vlgp/core.py:81-97), Jacobi over latents.
// vec: per latent 3 x 64 doubles (pv / cv / uv).
template <int LT>
__device__ __forceinline__ void mean_step_all(const SegArgs &p, const Smem<LT> &s, const int *bad) {
    const int W = p.W, tid = threadIdx.x;
    const float inv_w = 1.0f / (float)W;
    // p = G' (resid a_l)
    for (int idx = tid; idx < p.col_total; idx += NT) {
        int l = 0;
        while (l + 1 < LT && idx >= p.coloff[l + 1]) ++l;
        const int j = idx - p.coloff[l], ld = p.ldg[l];
        const double *g = s.Gs + p.goff[l];
        double acc = 0.0;
        for (int t = 0; t < W; ++t) acc = fma(g[t * ld + j], s.ra[t * LT + l], acc);
        s.vec[l * 192 + j] = acc;
    }
    __syncthreads();
    // u = G p - mu_l
    for (int idx = tid; idx < LT * W; idx += NT) {
        const int l = (int)(((float)idx + 0.5f) * inv_w), t = idx - l * W;
        const int nc = p.nc[l], ld = p.ldg[l];
        const double *g = s.Gs + p.goff[l] + t * ld;
        const double *pv = s.vec + l * 192;
        double acc = 0.0;
        for (int j = 0; j < nc; ++j) acc = fma(g[j], pv[j], acc);
        s.vec[l * 192 + 128 + t] = acc - s.mu[t * LT + l];
    }
    __syncthreads();
    // c = G' (w_l o u)
    for (int idx = tid; idx < p.col_total; idx += NT) {
        int l = 0;
        while (l + 1 < LT && idx >= p.coloff[l + 1]) ++l;
        const int j = idx - p.coloff[l], ld = p.ldg[l];
        const double *g = s.Gs + p.goff[l];
        const double *uv = s.vec + l * 192 + 128;
        double acc = 0.0;
        for (int t = 0; t < W; ++t) acc = fma(g[t * ld + j], s.w[t * LT + l] * uv[t], acc);
        s.vec[l * 192 + 64 + j] = acc;
    }
    __syncthreads();
    // m = Minv c   (M = -Minv, symmetric)
    for (int idx = tid; idx < p.col_total; idx += NT) {
        int l = 0;
        while (l + 1 < LT && idx >= p.coloff[l + 1]) ++l;
        const int i = idx - p.coloff[l], nc = p.nc[l], ldm = p.ldm[l];
        const double *M = s.Mi + p.moff[l];
        const double *cv = s.vec + l * 192 + 64;
        double acc = 0.0;
        for (int j = 0; j < nc; ++j) acc = fma(M[j * ldm + i], cv[j], acc);
        s.vec[l * 192 + i] = -acc;
    }
    __syncthreads();
    // delta = clip(u - G m); a failed factorisation zeroes the step (vlgp/core.py:92-94)
    for (int idx = tid; idx < LT * W; idx += NT) {
        const int l = (int)(((float)idx + 0.5f) * inv_w), t = idx - l * W;
        double d = 0.0;
        if (!bad[l]) {
            const int nc = p.nc[l], ld = p.ldg[l];
            const double *g = s.Gs + p.goff[l] + t * ld;
            const double *mv = s.vec + l * 192;
            double acc = 0.0;
            for (int j = 0; j < nc; ++j) acc = fma(g[j], mv[j], acc);
            d = clipd(s.vec[l * 192 + 128 + t] - acc, p.dmu_bound);
        }
        s.dmu[t * LT + l] = d;
        s.mu[t * LT + l] += d;
    }
    __syncthreads();
}

// out[j] = sum_t G[t][j] f(t) for the nc columns of one latent's factor, by one warp: the bins are split over 32 / J
// groups of J = 8, 16 or 32 lanes (J >= nc when nc <= 32) whose partial sums are combined by shuffles.
template <class F>
__device__ __forceinline__ void warp_gt_product(const double *g, int ld, int nc, int W, F f, double *out) {
    const int lane = threadIdx.x & 31;
    if (nc <= 16) {
        const int J = nc <= 8 ? 8 : 16, j = lane & (J - 1), part = lane / J, nparts = 32 / J;
        double acc = 0.0;
        if (j < nc)
            for (int t = part; t < W; t += nparts) acc = fma(g[t * ld + j], f(t), acc);
        for (int o = 16; o >= J; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane < nc) out[lane] = acc;
    } else {
        for (int j = lane; j < nc; j += 32) {
            double acc = 0.0;
            for (int t = 0; t < W; ++t) acc = fma(g[t * ld + j], f(t), acc);
            out[j] = acc;
        }
    }
}

// Newton step of the posterior mean of ONE latent by ONE warp (vlgp/core.py:81-97): the five mat-vec stages are
// separated by __syncwarp only, so the latents proceed concurrently in their own warps and the whole mean step costs
// one block barrier instead of five.  vec: per latent 3 x 64 doubles (p, m | c | u).
template <int LT>
__device__ __forceinline__ void mean_step_warp(const SegArgs &p, const Smem<LT> &s, int l, bool bad) {
    const int lane = threadIdx.x & 31, W = p.W, nc = p.nc[l], ld = p.ldg[l], ldm = p.ldm[l];
    const double *g = s.Gs + p.goff[l], *M = s.Mi + p.moff[l];
    double *pv = s.vec + l * 192, *cv = pv + 64, *uv = pv + 128;
    warp_gt_product(g, ld, nc, W, [&](int t) { return s.ra[t * LT + l]; }, pv);          // p = G' (resid a_l)
    __syncwarp();
    for (int t = lane; t < W; t += 32) {                                                  // u = G p - mu_l
        double acc = 0.0;
        for (int j = 0; j < nc; ++j) acc = fma(g[t * ld + j], pv[j], acc);
        uv[t] = acc - s.mu[t * LT + l];
    }
    __syncwarp();
    warp_gt_product(g, ld, nc, W, [&](int t) { return s.w[t * LT + l] * uv[t]; }, cv);    // c = G' (w_l o u)
    __syncwarp();
    for (int i = lane; i < nc; i += 32) {                                                 // m = Minv c  (M = -Minv)
        double acc = 0.0;
        for (int j = 0; j < nc; ++j) acc = fma(M[j * ldm + i], cv[j], acc);
        pv[i] = -acc;
    }
    __syncwarp();
    for (int t = lane; t < W; t += 32) {        // delta = clip(u - G m); a failed factorisation zeroes the step (:92-94)
        double d = 0.0;
        if (!bad) {
            double acc = 0.0;
            for (int j = 0; j < nc; ++j) acc = fma(g[t * ld + j], pv[j], acc);
            d = clipd(uv[t] - acc, p.dmu_bound);
        }
        s.dmu[t * LT + l] = d;
        s.mu[t * LT + l] += d;
    }
}

template <int LT, int NBMAX, int FAST>
#ifdef VLGP_ESTEP_TWO_BINS
__global__ void __launch_bounds__(NT, 2) estep_seg_kernel(SegArgs p) {
#else
__global__ void __launch_bounds__(NT, (NBMAX <= 2 ? 3 : 2)) estep_seg_kernel(SegArgs p) {
#endif
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Smem<LT> s(smem_raw, p);
    __shared__ int bad[VLGP_MAX_L];
    const int tid = threadIdx.x;
    const int W = p.W, N = p.N;

    // The CTAs resident on one SM run the same phase sequence at the same pace: started together they stay in lockstep,
    // all of them in the pipe-bound rate passes at the same time and all of them in the latency-bound r x r phases at the
    // same time.  The k-th CTA to arrive on an SM therefore waits k * stagger cycles once, so that the phases interleave.
    if (p.stagger > 0) {
        __shared__ int slot_s;
        if (tid == 0) {
            unsigned smid;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            slot_s = atomicAdd(p.sm_slots + smid, 1);
        }
        __syncthreads();
        const long long wait = (long long)slot_s * p.stagger, t0 = clock64();
        while (clock64() - t0 < wait) __nanosleep(200);
    }

    // ---- once per CTA: parameters and the compact prior factors ------------------------------------------------------
#ifndef VLGP_ESTEP_SCALAR_RATE_PASS
    if (FAST == 2) {     // single-precision rate passes: (a, a^2 / 2) pairs and the bias as floats
        float2 *af = (float2 *)s.a;
        float *bf = (float *)(af + LT * N);
        for (int i = tid; i < LT * N; i += NT) af[i] = make_float2((float)p.pa[i].x, 0.5f * (float)p.pa[i].y);
        for (int n = tid; n < N; n += NT) bf[n] = (float)p.pb[n].x;
    } else if (FAST) {   // operand of the tensor-path rate passes: rows a_l | a_l^2 / 2 | b | 0, columns padded with 0
        for (int i = tid; i < p.kp * p.np; i += NT) {
            const int k = i / p.np, n = i - k * p.np;
            double val = 0.0;
            if (n < N) {
                if (k < LT) val = p.pa[k * N + n].x;
                else if (k < 2 * LT) val = 0.5 * p.pa[(k - LT) * N + n].y;
                else if (k == 2 * LT) val = p.pb[n].x;
            }
            s.a[i] = val;
        }
    } else
#endif
    {
        for (int i = tid; i < LT * N; i += NT) ((double2 *)s.a)[i] = p.pa[i];
        for (int n = tid; n < N; n += NT) ((double2 *)s.b)[n] = p.pb[n];
    }
    for (int n = tid; n < N; n += NT) s.pois[n] = p.poisson[n];
    if (tid < 32) s.etab[tid] = VLGP_EXP_T[tid];
    if (NBMAX == 4 && p.g_global) {
        s.Gs = p.G;                  // read in place (generic loads; only the NBMAX = 4 instantiations pay for that)
    } else {
        double *Gc = const_cast<double *>(s.Gs);
        for (int l = 0; l < LT; ++l) {
            const int nc = p.nc[l], ldg = p.ldg[l];
            const double *Gsrc = p.G + (size_t)l * W * p.rank;
            double *Gd = Gc + p.goff[l];
            for (int i = tid; i < W * nc; i += NT) {
                const int t = i / nc, c = i - t * nc;
                Gd[t * ldg + c] = Gsrc[(size_t)t * p.rank + c];
            }
        }
    }
    __syncthreads();

    for (int seg = blockIdx.x; seg < p.n_seg; seg += gridDim.x) {
        const int64_t bin0 = (int64_t)(p.subset ? p.subset[seg] : seg) * W;
        for (int i = tid; i < W * LT; i += NT) {
            s.mu[i] = p.mu[bin0 * LT + i];
            s.v[i] = p.v[bin0 * LT + i];
            s.w[i] = p.w[bin0 * LT + i];
            s.dmu[i] = 0.0;
        }
        if (FAST && p.fused) {
            // The count tile is read once per segment (y a_l' below): staged through the partial-sum region with
            // 128-bit loads (the buffer starts at the source's offset from a 16-byte boundary, so body chunks are
            // aligned on both sides).  Read in place by the row-tile warps it costs one HBM round trip per column tile
            // on a cold L2 -- with few segments per GPU (8-GPU shards) that was 0.8 ms of a 2.1 ms launch.
            const uint8_t *ysrc = (const uint8_t *)p.y + bin0 * N;
            const int mis = (int)((uintptr_t)ysrc & 15), nby = W * N;
            uint8_t *ybuf = (uint8_t *)(((uintptr_t)s.part + 15) & ~(uintptr_t)15) + mis;
            const int head = min(nby, (16 - mis) & 15);
            for (int i = tid; i < head; i += NT) ybuf[i] = ysrc[i];
            const int nvec = (nby - head) >> 4;
            const uint4 *src4 = (const uint4 *)(ysrc + head);
            uint4 *dst4 = (uint4 *)(ybuf + head);
            for (int i = tid; i < nvec; i += NT) dst4[i] = src4[i];
            for (int i = head + 16 * nvec + tid; i < nby; i += NT) ybuf[i] = ysrc[i];
            __syncthreads();
            if (8 * (tid >> 5) < W) ya_tiles<LT, FAST == 2>(p, s, ybuf);
        } else if (p.ydtype == VLGP_Y_U8) {
            const uint8_t *ysrc = (const uint8_t *)p.y + bin0 * N;
            for (int i = tid; i < W * N; i += NT) s.ys[i] = ysrc[i];
        }
        if (tid < LT) bad[tid] = 0;
        __syncthreads();

        if (FAST && p.fused && p.n_iter > 0) {
            factor_all<LT, NBMAX>(p, s, bad, false);                  // the first mean step uses the incoming w
            for (int it = 0; it < p.n_iter; ++it) {
                fused_a<LT, NBMAX, FAST == 2>(p, s, bad, it > 0 && p.method_vb && !(p.skip & 8));
                if (!(p.skip & 2)) fused_b<LT>(p, s);
                fused_c<LT, FAST == 2>(p, s, bad);
                if ((p.method_vb || it + 1 < p.n_iter) && !(p.skip & 4)) factor_all<LT, NBMAX>(p, s, bad, false);
            }
            if (p.method_vb && p.n_iter > 0) {                         // variances of the last factorisation
                if (8 * (tid >> 5) < W)
                    for (int l = 0; l < LT; ++l)
                        if (!bad[l]) variance_rows<LT, NBMAX>(p, s, l);
                __syncthreads();
            }
        } else
        for (int it = 0; it < p.n_iter; ++it) {
            if (!(p.skip & 1)) rate_pass<LT, 1, FAST>(p, s, bin0);    // ends with a barrier; part (aliases vec) is free again
            if (it == 0) factor_all<LT, NBMAX>(p, s, bad, false);      // the first mean step uses the incoming w
            if (!(p.skip & 2)) {
#ifdef VLGP_ESTEP_WARP_MEAN_STEP     // one warp per latent, one barrier: measured 0.2 ms SLOWER per launch than the flat form
                for (int l = tid >> 5; l < LT; l += NWARP) mean_step_warp<LT>(p, s, l, bad[l] != 0);
                __syncthreads();
#else
                mean_step_all<LT>(p, s, bad);
#endif
            }
            if (!(p.skip & 1)) rate_pass<LT, 2, FAST>(p, s, bin0);
            if ((p.method_vb || it + 1 < p.n_iter) && !(p.skip & 4)) factor_all<LT, NBMAX>(p, s, bad, p.method_vb != 0);
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

static __global__ void pack_params_kernel(int LN, int N, const double *__restrict__ a, const double *__restrict__ b,
                                   const double *__restrict__ noise, double2 *__restrict__ pa, double2 *__restrict__ pb) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < LN) {
        const double x = a[i];
        pa[i] = make_double2(x, x * x);
    }
    if (i < N) pb[i] = make_double2(b[i], 1.0 / noise[i]);
}

template <int LT, int NBMAX, int FAST>
int launch_seg_t(vlgp_ctx *ctx, TrialSet *ts, SegArgs &p, size_t smem, bool *handled) {
    const int LN = LT * p.N;
    if (!ctx->d_ppack) CK(cudaMalloc(&ctx->d_ppack, (size_t)(VLGP_MAX_L + 1) * p.N * sizeof(double2)));
    p.pa = (const double2 *)ctx->d_ppack;
    p.pb = p.pa + LN;
    pack_params_kernel<<<(LN + 255) / 256, 256, 0, ctx->stream>>>(LN, p.N, p.a, p.b, p.noise, (double2 *)p.pa,
                                                                  (double2 *)p.pb);
    CKL();
    CK(cudaFuncSetAttribute(estep_seg_kernel<LT, NBMAX, FAST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, estep_seg_kernel<LT, NBMAX, FAST>, NT, smem));
    if (per_sm < 1) return VLGP_OK;       // does not fit: let the general kernel handle it
    int grid = per_sm * ctx->prop.multiProcessorCount;
    if (grid > p.n_seg) grid = p.n_seg;
    if (!ctx->d_smslots) CK(cudaMalloc(&ctx->d_smslots, 1024 * sizeof(int)));
    p.sm_slots = ctx->d_smslots;
    {
        static const char *env = getenv("VLGP_ESTEP_STAGGER");
        p.stagger = env ? atoi(env) : 0;
    }
    if (p.stagger > 0) CK(cudaMemsetAsync(ctx->d_smslots, 0, 1024 * sizeof(int), ctx->stream));
    estep_seg_kernel<LT, NBMAX, FAST><<<grid, NT, smem, ctx->stream>>>(p);
    CKL();
    *handled = true;
    return VLGP_OK;
}


template <int NBMAX, int FAST>
int launch_seg_variant(vlgp_ctx *ctx, TrialSet *ts, SegArgs &p, size_t smem, bool *handled);

}   // namespace segk

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


// Each variant of the kernel is instantiated in its own translation unit (estep_seg_v*.cu) so that they compile in
// parallel and do not share a register budget.
#define VLGP_DEFINE_SEG_VARIANT(NBMAX, FAST)                                                                       \
    namespace segk {                                                                                               \
    template <>                                                                                                    \
    int launch_seg_variant<NBMAX, FAST>(vlgp_ctx * ctx, TrialSet * ts, SegArgs & p, size_t smem, bool *handled) { \
        int rc = VLGP_OK;                                                                                          \
        DISPATCH_L(ctx->L, (rc = launch_seg_t<LT, NBMAX, FAST>(ctx, ts, p, smem, handled)));                      \
        return rc;                                                                                                 \
    }                                                                                                              \
    }
