// K5/K6: M-step -- per-neuron Newton / least-squares updates of the loading a and bias b.
//
// Replaces core.mstep (vlgp/core.py:129-249).  One Newton iteration =
//   mstep_stats   : one streaming pass over all bins x neurons (HBM/L2 -> registers): rate, gradient, packed Hessian and
//                   noise moments per neuron, accumulated in registers, reduced per CTA, written as per-CTA partials
//   reduce_parts  : deterministic sum of the per-CTA partials -> nstat x N sufficient statistics
//   (NCCL sum-allreduce of that one fused buffer when trials are sharded over GPUs)
//   mstep_solve   : one thread per neuron: L x L Cholesky solve (gradient fallback when not PD), clip, update
// The rate is computed once per iteration and every neuron is updated from it (vlgp/core.py:174-176).
#include "common.cuh"
#include "p2p.cuh"
#include "tma.cuh"

// regress.cu: general regressors
int vlgp_launch_xb(vlgp_ctx *ctx, TrialSet *ts);
int vlgp_launch_bstats(vlgp_ctx *ctx, TrialSet *ts);
int vlgp_launch_bsolve(vlgp_ctx *ctx, int use_hessian, double eps, double lr, double db_bound);
int vlgp_bstat_muxb_offset(vlgp_ctx *ctx);

namespace {

// Per-iteration statistic slots per neuron (Poisson): [0,L) sum_t s_l r (s = mu + v o a_n) ; [L, L+L(L+1)/2) packed lower
// Hessian ; then sum_t r, sum e, sum e^2 (e = y - eta; only accumulated in the LAST iteration -- the reference keeps
// the noise of the last iteration, vlgp/core.py:177,242).  The y-moments mu'y (L) and sum y do not depend on (a, b):
// they are computed once per M-step (FIRST) into their own buffer instead of 25 times.
__host__ __device__ constexpr int nstat_of(int L) { return L + L * (L + 1) / 2 + 3; }
__host__ __device__ constexpr int nymom_of(int L) { return L + 1; }

struct MstatArgs {
    int64_t nbin;
    int N, NC, J;                // neurons, neurons per chunk (blockDim = J*NC rounded up to 32)
    const void *y;
    int ydtype;
    const double *mu, *v, *a, *b;
    const uint8_t *poisson;
    const double *xb;            // nbin x N offsets einsum(x, b) for general regressors (regress.cu), or null: b[n]
    double *part;                // gridDim.x x nstat x N
    double *ypart;               // gridDim.x x (L+1) x N   (FIRST only)
    int last;                    // accumulate the noise moments
    int rounds;                  // mstep_stats_tma_kernel: bin pairs per thread per stage
};

#ifndef VLGP_MS_U
#define VLGP_MS_U 2
#endif
constexpr int MS_U = VLGP_MS_U;          // bins per thread per tile: MS_U independent exp chains in flight
constexpr int MS_TB_MAX = 128;   // bins per SMEM tile (= MS_U * J <= 128)

template <int LT, bool FIRST, bool XB = false>
__global__ void __launch_bounds__((LT <= 5) ? 512 : 256) mstep_stats_kernel(MstatArgs p) {
    constexpr int NS = nstat_of(LT);
    extern __shared__ double sm[];
    double *muv = sm;                              // MS_TB_MAX x 2 LT : (mu, v) of the tile's bins
    double *red = sm + MS_TB_MAX * 2 * LT;          // 2 x blockDim
    const int tid = threadIdx.x;
    const int j = tid / p.NC;                      // bin lane
    const int nloc = tid - j * p.NC;
    const int n = blockIdx.y * p.NC + nloc;
    const bool active = j < p.J && n < p.N;
    const int64_t per = (p.nbin + gridDim.x - 1) / gridDim.x;
    const int64_t b0 = (int64_t)blockIdx.x * per;
    const int64_t b1 = b0 + per < p.nbin ? b0 + per : p.nbin;
    const int TB = MS_U * p.J;

    double acc[NS], yacc[FIRST ? LT + 1 : 1];
#pragma unroll
    for (int s = 0; s < NS; ++s) acc[s] = 0.0;
#pragma unroll
    for (int s = 0; s < (FIRST ? LT + 1 : 1); ++s) yacc[s] = 0.0;

    double al[LT], a2[LT];
    double bn = 0.0;
    bool pois = false;
    if (active) {
#pragma unroll
        for (int l = 0; l < LT; ++l) {
            al[l] = p.a[l * p.N + n];
            a2[l] = al[l] * al[l];
        }
        bn = p.b[n];
        pois = p.poisson[n] != 0;
    }
    // Software pipeline: the (mu, v) rows and this thread's own counts of tile k+1 are fetched into registers while
    // tile k is being processed, so the L2 latency of the loads is off the critical path.
    constexpr int PFMAX = 2;                       // host guarantees 2 * TB * LT <= PFMAX * blockDim
    double mvnext[PFMAX], ynext_d[MS_U];
    unsigned int ynext_b[MS_U];                    // raw count bytes: converted at use, so the prefetch does not wait
    auto prefetch = [&](int64_t t0) {
        const int nb = (int)((b1 - t0 < TB) ? (b1 - t0) : TB);
#pragma unroll
        for (int q = 0; q < PFMAX; ++q) {
            const int i = tid + q * blockDim.x;        // element i of the tile's [t][mu(LT) | v(LT)] block
            mvnext[q] = 0.0;
            if (t0 < b1 && i < nb * 2 * LT) {
                const int t = i / (2 * LT), c = i - t * 2 * LT;
                mvnext[q] = c < LT ? p.mu[(t0 + t) * LT + c] : p.v[(t0 + t) * LT + (c - LT)];
            }
        }
#pragma unroll
        for (int u = 0; u < MS_U; ++u) {
            const int t = j + u * p.J;
            ynext_b[u] = 0u;
            ynext_d[u] = 0.0;
            if (active && t0 < b1 && t < nb) {
                const int64_t idx = (t0 + t) * p.N + n;
                if (p.ydtype == VLGP_Y_U8) ynext_b[u] = ((const uint8_t *)p.y)[idx];
                else ynext_d[u] = ((const double *)p.y)[idx];
            }
        }
    };
    prefetch(b0);
    for (int64_t t0 = b0; t0 < b1; t0 += TB) {
        const int nb = (int)((b1 - t0 < TB) ? (b1 - t0) : TB);
        __syncthreads();
#pragma unroll
        for (int q = 0; q < PFMAX; ++q) {
            const int i = tid + q * blockDim.x;
            if (i < nb * 2 * LT) muv[i] = mvnext[q];
        }
        double ycur[MS_U];
#pragma unroll
        for (int u = 0; u < MS_U; ++u) ycur[u] = p.ydtype == VLGP_Y_U8 ? (double)ynext_b[u] : ynext_d[u];
        __syncthreads();
        prefetch(t0 + TB);
        if (!active) continue;
        // phase A: MS_U independent rate evaluations (arguments first, then the exponentials two at a time with
        // interleaved Horner chains)
        double r[MS_U], yv[MS_U], eta[MS_U], lin[MS_U];
#pragma unroll
        for (int u = 0; u < MS_U; ++u) {
            const int t = j + u * p.J;
            r[u] = 0.0;
            yv[u] = 0.0;
            eta[u] = 0.0;
            lin[u] = 0.0;
            if (t < nb) {
                const double *mv = muv + t * 2 * LT;
                double e = XB ? p.xb[(t0 + t) * p.N + n] : bn, h = 0.0;
#pragma unroll
                for (int l = 0; l < LT; ++l) {
                    e = fma(mv[l], al[l], e);
                    h = fma(mv[LT + l], a2[l], h);
                }
                eta[u] = e;
                yv[u] = ycur[u];
                lin[u] = e + 0.5 * h;
            }
        }
        if (pois) {
#pragma unroll
            for (int u = 0; u < MS_U; u += 2) trunc_exp2(lin[u], lin[u + 1], r[u], r[u + 1]);
        }
        // phase B: accumulate
#pragma unroll
        for (int u = 0; u < MS_U; ++u) {
            const int t = j + u * p.J;
            if (t >= nb) continue;
            const double *mv = muv + t * 2 * LT;
            if (FIRST) {
#pragma unroll
                for (int l = 0; l < LT; ++l) yacc[l] = fma(mv[l], yv[u], yacc[l]);
                yacc[LT] += yv[u];
            }
            if (p.last) {
                const double e = yv[u] - eta[u];
                acc[NS - 2] += e;
                acc[NS - 1] = fma(e, e, acc[NS - 1]);
            }
            if (pois) {
                const double rr = r[u];
                acc[NS - 3] += rr;
                double s[LT];
#pragma unroll
                for (int l = 0; l < LT; ++l) {
                    s[l] = fma(mv[LT + l], al[l], mv[l]);
                    acc[l] = fma(s[l], rr, acc[l]);
                }
                int q = LT;
#pragma unroll
                for (int l = 0; l < LT; ++l) {
                    const double rs = rr * s[l];
#pragma unroll
                    for (int k = 0; k <= l; ++k) {
                        acc[q] = fma(rs, s[k], acc[q]);
                        ++q;
                    }
                    acc[q - 1] = fma(rr, mv[LT + l], acc[q - 1]);     // + diag(r' v)   (vlgp/core.py:189)
                }
            }
        }
    }
    // reduce over the J bin lanes of this CTA, one statistic at a time (double-buffered: one barrier per statistic)
    __syncthreads();
#pragma unroll
    for (int s = 0; s < NS + (FIRST ? LT + 1 : 0); ++s) {
        double *buf = red + (s & 1) * blockDim.x;
        const double val = s < NS ? acc[s < NS ? s : 0] : yacc[FIRST ? (s >= NS ? s - NS : 0) : 0];
        buf[tid] = active ? val : 0.0;
        __syncthreads();
        if (j == 0 && n < p.N) {
            double x = 0.0;
            for (int jj = 0; jj < p.J; ++jj) x += buf[jj * p.NC + nloc];
            if (s < NS) p.part[((size_t)blockIdx.x * NS + s) * p.N + n] = x;
            else p.ypart[((size_t)blockIdx.x * (LT + 1) + (s - NS)) * p.N + n] = x;
        }
    }
}

// The same statistics for the iterations that need neither the y-moments (first) nor the noise moments (last): 23 of the
// 25 Newton iterations of an M-step read only (mu, v).  Their rows are staged CB bins at a time by TMA bulk copies into
// a ring of MT_STAGES shared-memory stages (one elected thread issues them, mbarrier completion), so the loop carries
// no global loads, no address arithmetic and no bounds checks; a thread = (bin-pair lane j, neuron) takes the adjacent
// bins (2 pr, 2 pr + 1), pr = j, j + J, ...: the two rows are 2 LT consecutive doubles, read with LT 128-bit broadcast
// loads per array.  Per entry ~60 FP64 instructions and ~9 others (the general kernel above: ~67 and ~100, which on
// this issue-bound pipe cost as much as the arithmetic).  Same per-entry arithmetic as mstep_stats_kernel; only the
// assignment of bins to accumulators (the summation order over bins) differs.
constexpr int MT_STAGES = 4;
constexpr int MT_ROUNDS = 16;            // bin pairs per thread per stage (measured on B200: 0.1428 / 0.1397 / 0.1438 ms at 8 / 16 / 32)

// MODE 0: middle iterations (no counts needed); 1: first iteration (adds the y-moments mu'y, sum y); 2: last iteration
// (adds the noise moments sum e, sum e^2 of e = y - eta).  Modes 1 and 2 also stage the uint8 count tile of the chunk
// (CB x N bytes, one more bulk copy per stage) and run for Gaussian channels too; they need all neurons in one CTA.
template <int LT, int MODE>
__global__ void __launch_bounds__((LT <= 5) ? 512 : 256) mstep_stats_tma_kernel(MstatArgs p, int64_t per) {
    constexpr int NS = nstat_of(LT);
    extern __shared__ __align__(16) unsigned char mt_raw[];
    const int J = p.J, CB = 2 * J * p.rounds;
    const size_t ybytes = MODE ? (((size_t)CB * p.N + 15) & ~(size_t)15) : 0;      // count tile of one stage
    const size_t stage_doubles = (size_t)2 * CB * LT + ybytes / sizeof(double);
    double *stage = (double *)mt_raw;                                  // MT_STAGES x [mu: CB x LT | v: CB x LT | y: CB x N u8]
    double *etab = stage + (size_t)MT_STAGES * stage_doubles;          // 32
    uint64_t *bar = (uint64_t *)(etab + 32);                           // MT_STAGES
    double *red = (double *)mt_raw;                                    // after the loop: 2 x blockDim
    const int tid = threadIdx.x;
    const int j = tid / p.NC;
    const int nloc = tid - j * p.NC;
    const int n = blockIdx.y * p.NC + nloc;
    const bool active = j < J && n < p.N;
    const int64_t b0 = (int64_t)blockIdx.x * per;
    const int64_t b1 = b0 + per < p.nbin ? b0 + per : p.nbin;
    const int nchunk = b1 > b0 ? (int)((b1 - b0 + CB - 1) / CB) : 0;

    auto issue = [&](int c) {                     // one thread: bulk copies of chunk c into its stage
        const int64_t t0 = b0 + (int64_t)c * CB;
        const int nb = (int)((b1 - t0 < CB) ? (b1 - t0) : CB);
        const unsigned bytes = (unsigned)(((size_t)nb * LT * sizeof(double) + 15) & ~(size_t)15);   // buffers are padded
        const unsigned yb = MODE ? (unsigned)(((size_t)nb * p.N + 15) & ~(size_t)15) : 0u;
        double *dst = stage + (size_t)(c % MT_STAGES) * stage_doubles;
        uint64_t *b = bar + (c % MT_STAGES);
        mbar_expect_tx(b, 2 * bytes + yb);
        tma_load_1d(dst, p.mu + t0 * LT, bytes, b);
        tma_load_1d(dst + (size_t)CB * LT, p.v + t0 * LT, bytes, b);
        if (MODE) tma_load_1d(dst + (size_t)2 * CB * LT, (const uint8_t *)p.y + t0 * p.N, yb, b);
    };
    if (tid == 0) {
        for (int s = 0; s < MT_STAGES; ++s) mbar_init(bar + s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < 32) etab[tid] = VLGP_EXP_T[tid];
    __syncthreads();
    if (tid == 0)
        for (int c = 0; c < MT_STAGES && c < nchunk; ++c) issue(c);

    double acc[NS], yacc[MODE == 1 ? LT + 1 : 1];
#pragma unroll
    for (int s = 0; s < NS; ++s) acc[s] = 0.0;
#pragma unroll
    for (int s = 0; s < (MODE == 1 ? LT + 1 : 1); ++s) yacc[s] = 0.0;
    double al[LT], a2h[LT];
    double bn = 0.0;
    bool pois = false;
    if (active) {
#pragma unroll
        for (int l = 0; l < LT; ++l) {
            al[l] = p.a[l * p.N + n];
            a2h[l] = 0.5 * (al[l] * al[l]);
        }
        bn = p.b[n];
        pois = p.poisson[n] != 0;
    }
    const bool work = active && (pois || MODE != 0);

    // exp(min(x, 10)) of two arguments, chains interleaved, table in shared memory (bitwise trunc_exp2)
    auto exp2 = [&](double x0, double x1, double &e0, double &e1) {
        x0 = x0 > 10.0 ? 10.0 : x0;
        x1 = x1 > 10.0 ? 10.0 : x1;
        x0 = x0 < -708.0 ? -708.0 : x0;
        x1 = x1 < -708.0 ? -708.0 : x1;
        const double shift = 6755399441055744.0;
        const double m0 = fma(x0, VLGP_EXP_INV, shift), m1 = fma(x1, VLGP_EXP_INV, shift);
        const int i0 = __double2loint(m0), i1 = __double2loint(m1);
        const double tj0 = etab[i0 & 31], tj1 = etab[i1 & 31];
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
        const int ex0 = max(i0 >> 5, -1022), ex1 = max(i1 >> 5, -1022);
        e0 = __hiloint2double(__double2hiint(q0) + (ex0 << 20), __double2loint(q0));
        e1 = __hiloint2double(__double2hiint(q1) + (ex1 << 20), __double2loint(q1));
    };
    auto accumulate = [&](const double *m, const double *vv, double rr) {
        acc[NS - 3] += rr;
        double s[LT];
#pragma unroll
        for (int l = 0; l < LT; ++l) {
            s[l] = fma(vv[l], al[l], m[l]);
            acc[l] = fma(s[l], rr, acc[l]);
        }
        int q = LT;
#pragma unroll
        for (int l = 0; l < LT; ++l) {
            const double rs = rr * s[l];
#pragma unroll
            for (int k = 0; k <= l; ++k) {
                acc[q] = fma(rs, s[k], acc[q]);
                ++q;
            }
            acc[q - 1] = fma(rr, vv[l], acc[q - 1]);
        }
    };
    auto lin_of = [&](const double *m, const double *vv, double &eta) {
        double e = bn, h = 0.0;
#pragma unroll
        for (int l = 0; l < LT; ++l) {
            e = fma(m[l], al[l], e);
            h = fma(vv[l], a2h[l], h);
        }
        eta = e;
        return e + h;
    };
    auto moments = [&](const double *m, double yv, double eta) {       // modes 1 / 2 only
        if (MODE == 1) {
#pragma unroll
            for (int l = 0; l < LT; ++l) yacc[l] = fma(m[l], yv, yacc[l]);
            yacc[MODE == 1 ? LT : 0] += yv;
        }
        if (MODE == 2) {
            const double e = yv - eta;
            acc[NS - 2] += e;
            acc[NS - 1] = fma(e, e, acc[NS - 1]);
        }
    };

    for (int c = 0; c < nchunk; ++c) {
        const int st = c % MT_STAGES;
        const int64_t t0 = b0 + (int64_t)c * CB;
        const int nb = (int)((b1 - t0 < CB) ? (b1 - t0) : CB);
        mbar_wait(bar + st, (unsigned)((c / MT_STAGES) & 1));
        if (work) {
            const double *smu = stage + (size_t)st * stage_doubles;
            const double *sv = smu + (size_t)CB * LT;
            const uint8_t *sy = (const uint8_t *)(sv + (size_t)CB * LT);
            const int npair = nb >> 1;
#pragma unroll 1
            for (int pr = j; pr < npair; pr += J) {
                double m[2 * LT], vv[2 * LT];
                const double2 *m2 = (const double2 *)(smu + (size_t)pr * 2 * LT);
                const double2 *v2 = (const double2 *)(sv + (size_t)pr * 2 * LT);
#pragma unroll
                for (int q = 0; q < LT; ++q) {
                    const double2 x = m2[q], y = v2[q];
                    m[2 * q] = x.x;
                    m[2 * q + 1] = x.y;
                    vv[2 * q] = y.x;
                    vv[2 * q + 1] = y.y;
                }
                double r0, r1, eta0, eta1;
                const double x0 = lin_of(m, vv, eta0), x1 = lin_of(m + LT, vv + LT, eta1);
                if (MODE) {
                    moments(m, (double)sy[(size_t)(2 * pr) * p.N + n], eta0);
                    moments(m + LT, (double)sy[(size_t)(2 * pr + 1) * p.N + n], eta1);
                }
                if (MODE == 0 || pois) {
                    exp2(x0, x1, r0, r1);
                    accumulate(m, vv, r0);
                    accumulate(m + LT, vv + LT, r1);
                }
            }
            if ((nb & 1) && (npair % J) == j) {       // odd bin at the very end of the bin range
                double m[LT], vv[LT];
#pragma unroll
                for (int l = 0; l < LT; ++l) {
                    m[l] = smu[(size_t)(nb - 1) * LT + l];
                    vv[l] = sv[(size_t)(nb - 1) * LT + l];
                }
                double r0, r1, eta0;
                const double x = lin_of(m, vv, eta0);
                if (MODE) moments(m, (double)sy[(size_t)(nb - 1) * p.N + n], eta0);
                if (MODE == 0 || pois) {
                    exp2(x, x, r0, r1);
                    accumulate(m, vv, r0);
                }
            }
        }
        __syncthreads();
        if (tid == 0 && c + MT_STAGES < nchunk) issue(c + MT_STAGES);
    }
    // reduce over the J bin lanes of this CTA (as in mstep_stats_kernel); the stages are free now
    __syncthreads();
#pragma unroll
    for (int s = 0; s < NS + (MODE == 1 ? LT + 1 : 0); ++s) {
        double *buf = red + (s & 1) * blockDim.x;
        const double val = s < NS ? acc[s < NS ? s : 0] : yacc[MODE == 1 ? (s >= NS ? s - NS : 0) : 0];
        buf[tid] = work ? val : 0.0;
        __syncthreads();
        if (j == 0 && n < p.N) {
            double x = 0.0;
            for (int jj = 0; jj < J; ++jj) x += buf[jj * p.NC + nloc];
            if (s < NS) p.part[((size_t)blockIdx.x * NS + s) * p.N + n] = x;
            else p.ypart[((size_t)blockIdx.x * (LT + 1) + (s - NS)) * p.N + n] = x;
        }
    }
}

// out[k] = sum_g part[g * K + k]    (deterministic; K = nstat x N)
__global__ void reduce_parts_kernel(const double *__restrict__ part, int G, int K, double *__restrict__ out) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= K) return;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    int g = 0;
    for (; g + 3 < G; g += 4) {
        s0 += part[(size_t)g * K + k];
        s1 += part[(size_t)(g + 1) * K + k];
        s2 += part[(size_t)(g + 2) * K + k];
        s3 += part[(size_t)(g + 3) * K + k];
    }
    for (; g < G; ++g) s0 += part[(size_t)g * K + k];
    out[k] = (s0 + s1) + (s2 + s3);
}

// Shared moments of the Gaussian channels: mu'mu (L x L), sum v (L), sum mu (L)  (vlgp/core.py:224-232)
template <int LT>
__global__ void __launch_bounds__(256) gauss_moments_kernel(int64_t nbin, const double *__restrict__ mu,
                                                            const double *__restrict__ v, double *part) {
    constexpr int K = LT * LT + 2 * LT;
    __shared__ double red[32];
    double acc[K];
#pragma unroll
    for (int s = 0; s < K; ++s) acc[s] = 0.0;
    for (int64_t bin = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; bin < nbin; bin += (int64_t)gridDim.x * blockDim.x) {
        double m[LT];
#pragma unroll
        for (int l = 0; l < LT; ++l) m[l] = mu[bin * LT + l];
#pragma unroll
        for (int l = 0; l < LT; ++l) {
#pragma unroll
            for (int k = 0; k < LT; ++k) acc[l * LT + k] = fma(m[l], m[k], acc[l * LT + k]);
            acc[LT * LT + l] += v[bin * LT + l];
            acc[LT * LT + LT + l] += m[l];
        }
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int s = 0; s < K; ++s) {
        const double x = warp_sum(acc[s]);
        __syncthreads();
        if (lane == 0) red[wid] = x;
        __syncthreads();
        if (threadIdx.x == 0) {
            double r = 0.0;
            for (int i = 0; i < (int)(blockDim.x >> 5); ++i) r += red[i];
            part[(size_t)blockIdx.x * K + s] = r;
        }
    }
}

struct MsolveArgs {
    int N;
    double count;                // total number of bins (all ranks)
    const double *stat;          // nstat x N
    const double *ymom;          // (L+1) x N : mu'y per latent, sum y
    const double *gshared;       // L*L + 2L (Gaussian channels) or null
    const uint8_t *poisson;
    double *a, *b, *noise, *da, *db;
    int use_hessian, last;
    double eps, lr, da_bound, db_bound;
    int *flags;                  // flags[1] += gradient fallbacks
    const double *gmuxb;         // general regressors: L x N sums mu'(x b) for the Gaussian channels; b is then updated
                                 // by mstep_bsolve_kernel (regress.cu), not here
};

// Solve H x = g for SPD H (L x L, full storage, destroyed).  Returns false if not positive definite
// (same criterion as LAPACK posv behind scipy.linalg.solve(sym_pos=True), vlgp/core.py:193).
template <int LT>
__device__ __forceinline__ bool chol_solve_small(double (&H)[LT][LT], double (&g)[LT]) {
#pragma unroll
    for (int k = 0; k < LT; ++k) {
        double d = H[k][k];
#pragma unroll
        for (int m = 0; m < k; ++m) d = fma(-H[k][m], H[k][m], d);
        if (!(d > 0.0)) return false;
        const double lkk = sqrt(d);
        H[k][k] = lkk;
#pragma unroll
        for (int i = k + 1; i < LT; ++i) {
            double s = H[i][k];
#pragma unroll
            for (int m = 0; m < k; ++m) s = fma(-H[i][m], H[k][m], s);
            H[i][k] = s / lkk;
        }
    }
#pragma unroll
    for (int i = 0; i < LT; ++i) {          // forward
        double s = g[i];
#pragma unroll
        for (int m = 0; m < i; ++m) s = fma(-H[i][m], g[m], s);
        g[i] = s / H[i][i];
    }
#pragma unroll
    for (int i = LT - 1; i >= 0; --i) {     // backward
        double s = g[i];
#pragma unroll
        for (int m = i + 1; m < LT; ++m) s = fma(-H[m][i], g[m], s);
        g[i] = s / H[i][i];
    }
    return true;
}

// Newton / least-squares update of neuron n from its sufficient statistics S(s) and y-moments Y(s).
template <int LT, class SF, class YF>
__device__ __forceinline__ void mstep_solve_neuron(const MsolveArgs &p, int n, SF S, YF Y) {
    constexpr int NS = nstat_of(LT);
    const int N = p.N;
    if (p.last) {
        const double me = S(NS - 2) / p.count;
        p.noise[n] = S(NS - 1) / p.count - me * me;   // np.var(y - eta, ddof=0), vlgp/core.py:177
    }
    if (p.poisson[n]) {
        double g[LT], H[LT][LT];
#pragma unroll
        for (int l = 0; l < LT; ++l) g[l] = Y(l) - S(l);      // mu'y - (mu + v o a)' r
        double step[LT];
        bool newton = p.use_hessian != 0;
        if (newton) {
            int q = LT;
#pragma unroll
            for (int l = 0; l < LT; ++l)
#pragma unroll
                for (int k = 0; k <= l; ++k) {
                    const double h = S(q++);
                    H[l][k] = h;
                    H[k][l] = h;
                }
#pragma unroll
            for (int l = 0; l < LT; ++l) {
                H[l][l] += p.eps;
                step[l] = g[l];
            }
            if (!chol_solve_small<LT>(H, step)) {
                newton = false;
                atomicAdd(p.flags + 1, 1);
            }
        }
        if (!newton) {
#pragma unroll
            for (int l = 0; l < LT; ++l) step[l] = p.lr * g[l];
        }
#pragma unroll
        for (int l = 0; l < LT; ++l) {
            const double d = clipd(step[l], p.da_bound);
            p.da[l * N + n] = d;
            p.a[l * N + n] += d;
        }
        if (p.gmuxb != nullptr) return;                        // b: regress.cu
        const double gb = Y(LT) - S(NS - 3);                   // sum (y - r)
        double sb;
        const double hb = S(NS - 3) + p.eps;
        if (p.use_hessian && hb > 0.0) sb = gb / hb;
        else {
            sb = p.lr * gb;
            if (p.use_hessian) atomicAdd(p.flags + 1, 1);
        }
        sb = clipd(sb, p.db_bound);
        p.db[n] = sb;
        p.b[n] += sb;
    } else if (p.gshared != nullptr) {
        // least squares for a Gaussian channel (vlgp/core.py:221-235); da/db are left untouched like the reference
        double H[LT][LT], rhs[LT], smu[LT];
        const double bn = p.b[n];
#pragma unroll
        for (int l = 0; l < LT; ++l) {
#pragma unroll
            for (int k = 0; k < LT; ++k) H[l][k] = p.gshared[l * LT + k];
            H[l][l] += p.gshared[LT * LT + l];
            smu[l] = p.gshared[LT * LT + LT + l];
            rhs[l] = p.gmuxb ? Y(l) - p.gmuxb[(size_t)l * N + n] : Y(l) - smu[l] * bn;
        }
        if (chol_solve_small<LT>(H, rhs)) {
            double dot = 0.0;
#pragma unroll
            for (int l = 0; l < LT; ++l) {
                p.a[l * N + n] = rhs[l];
                dot = fma(smu[l], rhs[l], dot);
            }
            if (p.gmuxb == nullptr) p.b[n] = (Y(LT) - dot) / p.count;
        } else {
            atomicAdd(p.flags + 1, 1);
        }
    }
}

template <int LT>
__global__ void mstep_solve_kernel(MsolveArgs p) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= p.N) return;
    const int N = p.N;
    mstep_solve_neuron<LT>(p, n, [&](int s) { return p.stat[(size_t)s * N + n]; },
                           [&](int s) { return p.ymom[(size_t)s * N + n]; });
}

// reduce_parts + (peer-memory allreduce over ranks) + solve in ONE launch: CTA c owns MR_NPB neurons.  Its warps sum the
// statistics kernel's per-CTA partials of those neurons (warp w takes partials w, w + 8, ...; lanes take the entries),
// the eight warp sums are added in warp order, the result is exchanged with the peers from inside the kernel
// (p2p.cuh; chunk = CTA index) and one thread per neuron runs the L x L solve from shared memory.  Replaces three
// launches and an NCCL call per Newton iteration (25 per M-step) with one; deterministic summation order.
constexpr int MR_NPB = 4;

template <int LT>
__global__ void __launch_bounds__(256) mstep_reduce_solve_kernel(MsolveArgs p, const double *__restrict__ part,
                                                                 const double *__restrict__ ypart, int G, int first,
                                                                 P2PDev pd, double *__restrict__ stat_out,
                                                                 double *__restrict__ ymom_out) {
    constexpr int NS = nstat_of(LT), NY = LT + 1, EMAX = (NS + NY) * MR_NPB;
    __shared__ double red[8][EMAX];
    __shared__ double tot[EMAX];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int N = p.N, n0 = blockIdx.x * MR_NPB;
    const int nn = N - n0 < MR_NPB ? N - n0 : MR_NPB;
    const int E = NS * MR_NPB, ET = E + (first ? NY * MR_NPB : 0);
    for (int e = lane; e < ET; e += 32) {
        const bool isy = e >= E;
        const int e2 = isy ? e - E : e;
        const int st = e2 / MR_NPB, nl = e2 - st * MR_NPB;
        const double *src = (isy ? ypart : part) + (size_t)st * N + n0 + nl;
        const size_t stride = (size_t)(isy ? NY : NS) * N;
        double a0 = 0.0, a1 = 0.0;
        if (nl < nn) {
            int g = wid;
            for (; g + 8 < G; g += 16) {
                a0 += src[(size_t)g * stride];
                a1 += src[(size_t)(g + 8) * stride];
            }
            if (g < G) a0 += src[(size_t)g * stride];
        }
        red[wid][e] = a0 + a1;
    }
    __syncthreads();
    for (int e = tid; e < ET; e += 256) {
        double x = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) x += red[w][e];
        tot[e] = x;
    }
    __syncthreads();
    p2p_allreduce_cta(pd, blockIdx.x, (size_t)blockIdx.x * EMAX, tot, ET);
    for (int e = tid; e < ET; e += 256) {
        const bool isy = e >= E;
        const int e2 = isy ? e - E : e;
        const int st = e2 / MR_NPB, nl = e2 - st * MR_NPB;
        if (nl < nn) (isy ? ymom_out : stat_out)[(size_t)st * N + n0 + nl] = tot[e];
    }
    if (tid < nn) {
        const int n = n0 + tid;
        const double *ym = p.ymom;
        if (first)
            mstep_solve_neuron<LT>(p, n, [&](int s) { return tot[s * MR_NPB + tid]; },
                                   [&](int s) { return tot[E + s * MR_NPB + tid]; });
        else
            mstep_solve_neuron<LT>(p, n, [&](int s) { return tot[s * MR_NPB + tid]; },
                                   [&](int s) { return ym[(size_t)s * N + n]; });
    }
}

// One M-step as a resumable job: set-up once, then Newton iterations enqueued a few at a time (vlgp_mstep_job_pump), so
// that an overlapped M-step's ~100 launches are issued while the host waits for H-step rounds instead of ahead of them.
struct MstepJob {
    MstatArgs sa{};
    MsolveArgs so{};
    dim3 grid;
    int nt = 0, K = 0, KY = 0, LT = 0;
    size_t smem = 0;
    int64_t gx = 0;
    double *ypart = nullptr;
    int n_iter = 0, next_it = 0;
    bool general_x = false;
    TrialSet *ts = nullptr;
    size_t smem_tma = 0;         // mstep_stats_tma_kernel: dynamic shared memory (0: not usable for this shape)
    size_t smem_tma_y = 0;       // the same for its first / last-iteration modes, which also stage the count tile
    int64_t per_tma = 0;         // its bins per CTA (a multiple of 16, so that every chunk starts on a 16-byte boundary)
};

template <int LT>
int mstep_setup_t(vlgp_ctx *ctx, TrialSet *ts, MstepJob &job, int n_iter, int use_hessian, double eps, double lr,
                  double da_bound, double db_bound) {
    constexpr int NS = nstat_of(LT);
    constexpr int MAXT = (LT <= 5) ? 512 : 256;
    const int N = ctx->N;
    const int NC = N < MAXT ? N : MAXT;
    const int nchunk = (N + NC - 1) / NC;
    int J = (MAXT / NC) < (MS_TB_MAX / MS_U) ? (MAXT / NC) : (MS_TB_MAX / MS_U);
    while (J > 1 && 2 * MS_U * J * LT > 2 * (((J * NC + 31) / 32) * 32)) --J;      // prefetch registers: PFMAX = 2
    int nt = ((J * NC + 31) / 32) * 32;
    const size_t smem = ((size_t)MS_TB_MAX * 2 * LT + 2 * (size_t)nt) * sizeof(double);
    int per_sm = 1;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, mstep_stats_kernel<LT, false>, nt, smem));
    if (per_sm < 1) per_sm = 1;
    int64_t gx = (int64_t)per_sm * ctx->prop.multiProcessorCount / nchunk;
    const int64_t min_bins = (int64_t)MS_U * J;    // at least one full tile per CTA
    if (gx > (ts->nbin + min_bins - 1) / min_bins) gx = (ts->nbin + min_bins - 1) / min_bins;
    if (gx < 1) gx = 1;
    const int K = NS * N;
    const int KY = (LT + 1) * N;
    if (ctx->mpart_grid < gx * (K + KY + 1)) {
        if (ctx->d_mpart) CK(cudaFree(ctx->d_mpart));
        ctx->d_mpart = nullptr;
        CK(cudaMalloc(&ctx->d_mpart, (size_t)gx * (K + KY + 64) * sizeof(double)));
        ctx->mpart_grid = (int)(gx * (K + KY + 1));
    }
    if (!ctx->d_ymom) CK(cudaMalloc(&ctx->d_ymom, (size_t)(VLGP_MAX_L + 1) * N * sizeof(double)));
    double *ypart = ctx->d_mpart + (size_t)gx * K;

    // total bin count over all ranks (the divisor of np.var / the Gaussian bias)
    double count = (double)ts->nbin;
    if (ctx->n_ranks > 1 && ts->nbin_all_ranks > 0.0) {
        count = ts->nbin_all_ranks;
    } else if (ctx->n_ranks > 1) {
        ctx->h_pin[0] = count;
        CK(cudaMemcpyAsync(ctx->d_small, ctx->h_pin, sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        int rc = vlgp_allreduce_dev(ctx, ctx->d_small, 1, 0);
        if (rc) return rc;
        CK(cudaMemcpyAsync(ctx->h_pin, ctx->d_small, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        count = ctx->h_pin[0];
        ts->nbin_all_ranks = count;
    }

    double *gshared = nullptr;
    if (ctx->any_gauss) {
        constexpr int KG = LT * LT + 2 * LT;
        const int gg = 2 * ctx->prop.multiProcessorCount;
        double *gpart = nullptr;
        CK(cudaMalloc(&gpart, (size_t)gg * KG * sizeof(double)));
        if (!ctx->d_gshared) CK(cudaMalloc(&ctx->d_gshared, (VLGP_MAX_L * VLGP_MAX_L + 2 * VLGP_MAX_L) * sizeof(double)));
        gauss_moments_kernel<LT><<<gg, 256, 0, ctx->stream>>>(ts->nbin, ts->d_mu, ts->d_v, gpart);
        CKL();
        reduce_parts_kernel<<<(KG + 127) / 128, 128, 0, ctx->stream>>>(gpart, gg, KG, ctx->d_gshared);
        CKL();
        int rc = vlgp_allreduce_dev(ctx, ctx->d_gshared, KG, 0);
        if (rc) return rc;
        CK(cudaStreamSynchronize(ctx->stream));
        CK(cudaFree(gpart));
        gshared = ctx->d_gshared;
    }

    MstatArgs sa{};
    sa.nbin = ts->nbin; sa.N = N; sa.NC = NC; sa.J = J;
    sa.y = ts->d_y; sa.ydtype = ts->ydtype;
    sa.mu = ts->d_mu; sa.v = ts->d_v; sa.a = ctx->d_a; sa.b = ctx->d_b; sa.poisson = ctx->d_poisson;
    sa.part = ctx->d_mpart;
    sa.ypart = ypart;
    MsolveArgs so{};
    so.N = N; so.count = count; so.stat = ctx->d_mstat; so.ymom = ctx->d_ymom; so.gshared = gshared; so.poisson = ctx->d_poisson;
    so.a = ctx->d_a; so.b = ctx->d_b; so.noise = ctx->d_noise; so.da = ctx->d_da; so.db = ctx->d_db;
    so.use_hessian = use_hessian; so.eps = eps; so.lr = lr; so.da_bound = da_bound; so.db_bound = db_bound;
    so.flags = ctx->d_flags;
    job.sa = sa; job.so = so;
    job.grid = dim3((unsigned)gx, nchunk);
    job.nt = nt; job.K = K; job.KY = KY; job.LT = LT; job.smem = smem; job.gx = gx; job.ypart = ypart;
    job.n_iter = n_iter; job.next_it = 0;
    job.general_x = ts->d_x != nullptr;
    job.ts = ts;
    {   // TMA-staged kernel
        static const int rounds_env = getenv("VLGP_MSTEP_ROUNDS") ? atoi(getenv("VLGP_MSTEP_ROUNDS")) : 0;
        job.sa.rounds = (rounds_env >= 8 && rounds_env <= 64 && rounds_env % 8 == 0) ? rounds_env : MT_ROUNDS;   // 16 J bins: count tiles stay 16-byte aligned
        const size_t cb = (size_t)2 * J * job.sa.rounds;
        const size_t redb = 2 * (size_t)nt * sizeof(double);
        auto need = [&](size_t ybytes) {
            const size_t ring = ((size_t)MT_STAGES * (2 * cb * LT + ybytes / sizeof(double)) + 32 + MT_STAGES) * sizeof(double);
            return ring > redb ? ring : redb;
        };
        job.smem_tma = need(0);
        job.smem_tma_y = need((cb * (size_t)N + 15) & ~(size_t)15);
        job.per_tma = (((ts->nbin + gx - 1) / gx) + 15) & ~(int64_t)15;
        const size_t cap = (size_t)200 << 10;
        if (job.smem_tma > cap || getenv("VLGP_MSTEP_NO_TMA")) job.smem_tma = 0;
        // the first / last-iteration modes read the counts: uint8 storage, all neurons in one CTA
        if (!job.smem_tma || job.smem_tma_y > cap || nchunk != 1 || ts->ydtype != VLGP_Y_U8 || n_iter < 2 ||
            getenv("VLGP_MSTEP_NO_TMA_Y"))
            job.smem_tma_y = 0;
        if (job.smem_tma > (size_t)48 << 10)
            CK(cudaFuncSetAttribute(mstep_stats_tma_kernel<LT, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)job.smem_tma));
        if (job.smem_tma_y > (size_t)48 << 10) {
            CK(cudaFuncSetAttribute(mstep_stats_tma_kernel<LT, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)job.smem_tma_y));
            CK(cudaFuncSetAttribute(mstep_stats_tma_kernel<LT, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)job.smem_tma_y));
        }
    }
    return VLGP_OK;
}

template <int LT>
int mstep_iter_t(vlgp_ctx *ctx, MstepJob &job, int it) {
    MstatArgs &sa = job.sa;
    MsolveArgs &so = job.so;
    const int K = job.K, KY = job.KY, N = so.N;
    const int64_t gx = job.gx;
    sa.last = so.last = (it == job.n_iter - 1);
    if (job.general_x) {
        int rcx = vlgp_launch_xb(ctx, job.ts);
        if (rcx) return rcx;
        sa.xb = job.ts->d_xb;
    }
    {
        ProfScope ps(ctx, 1);
        if (job.general_x) {
            if (it == 0)
                mstep_stats_kernel<LT, true, true><<<job.grid, job.nt, job.smem, ctx->stream>>>(sa);
            else
                mstep_stats_kernel<LT, false, true><<<job.grid, job.nt, job.smem, ctx->stream>>>(sa);
        } else if (it == 0 && job.smem_tma_y) {
            mstep_stats_tma_kernel<LT, 1><<<job.grid, job.nt, job.smem_tma_y, ctx->stream>>>(sa, job.per_tma);
        } else if (it == 0) {
            mstep_stats_kernel<LT, true><<<job.grid, job.nt, job.smem, ctx->stream>>>(sa);
        } else if (it < job.n_iter - 1 && job.smem_tma) {
            mstep_stats_tma_kernel<LT, 0><<<job.grid, job.nt, job.smem_tma, ctx->stream>>>(sa, job.per_tma);
        } else if (job.smem_tma_y) {
            mstep_stats_tma_kernel<LT, 2><<<job.grid, job.nt, job.smem_tma_y, ctx->stream>>>(sa, job.per_tma);
        } else {
            mstep_stats_kernel<LT, false><<<job.grid, job.nt, job.smem, ctx->stream>>>(sa);
        }
        CKL();
    }
    int rc;
    if (job.general_x) {
        // general regressors: statistics of b in their own pass (same parameters as the loading statistics above), then
        // the loading solve, then the b solve (Gaussian channels use the NEW loading, vlgp/core.py:226-234)
        rc = vlgp_launch_bstats(ctx, job.ts);
        if (rc) return rc;
        so.gmuxb = ctx->d_bstat + (size_t)vlgp_bstat_muxb_offset(ctx) * N;
        if (it == 0) {
            reduce_parts_kernel<<<(KY + 127) / 128, 128, 0, ctx->stream>>>(job.ypart, (int)gx, KY, ctx->d_ymom);
            CKL();
            rc = vlgp_allreduce_dev(ctx, ctx->d_ymom, KY, 0);
            if (rc) return rc;
        }
        reduce_parts_kernel<<<(K + 127) / 128, 128, 0, ctx->stream>>>(ctx->d_mpart, (int)gx, K, ctx->d_mstat);
        CKL();
        rc = vlgp_allreduce_dev(ctx, ctx->d_mstat, K, 0);
        if (rc) return rc;
        mstep_solve_kernel<LT><<<(N + 63) / 64, 64, 0, ctx->stream>>>(so);
        CKL();
        return vlgp_launch_bsolve(ctx, so.use_hessian, so.eps, so.lr, so.db_bound);
    }
    // one launch for reduce + exchange + solve when the exchange can happen inside the kernel (one rank, or peer memory)
    constexpr int EMAX = (nstat_of(LT) + LT + 1) * MR_NPB;
    const int nblk = (N + MR_NPB - 1) / MR_NPB;
    const bool fused = (ctx->n_ranks == 1 || vlgp_p2p_enabled(ctx)) && nblk <= VLGP_P2P_NCH &&
                       (size_t)nblk * EMAX <= (size_t)VLGP_P2P_PAY && !getenv("VLGP_MSTEP_UNFUSED");
    if (fused) {
        P2PDev pd = vlgp_p2p_next(ctx);
        mstep_reduce_solve_kernel<LT><<<nblk, 256, 0, ctx->stream>>>(so, ctx->d_mpart, job.ypart, (int)gx, it == 0 ? 1 : 0,
                                                                      pd, ctx->d_mstat, ctx->d_ymom);
        CKL();
        return VLGP_OK;
    }
    if (it == 0) {      // y-moments mu'y, sum y: once per M-step
        reduce_parts_kernel<<<(KY + 127) / 128, 128, 0, ctx->stream>>>(job.ypart, (int)gx, KY, ctx->d_ymom);
        CKL();
        rc = vlgp_allreduce_dev(ctx, ctx->d_ymom, KY, 0);
        if (rc) return rc;
    }
    reduce_parts_kernel<<<(K + 127) / 128, 128, 0, ctx->stream>>>(ctx->d_mpart, (int)gx, K, ctx->d_mstat);
    CKL();
    rc = vlgp_allreduce_dev(ctx, ctx->d_mstat, K, 0);
    if (rc) return rc;
    mstep_solve_kernel<LT><<<(N + 63) / 64, 64, 0, ctx->stream>>>(so);
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

// Job state lives behind vlgp_ctx::mstep_job (one M-step at a time per context).
static MstepJob *job_of(vlgp_ctx *ctx) {
    if (!ctx->mstep_job) ctx->mstep_job = new MstepJob();
    return (MstepJob *)ctx->mstep_job;
}

void vlgp_mstep_job_free(vlgp_ctx *ctx) {
    delete (MstepJob *)ctx->mstep_job;
    ctx->mstep_job = nullptr;
}

// Set-up on ctx->stream (grids, scratch, bin count over ranks, Gaussian-channel moments); enqueues no Newton iteration.
int vlgp_mstep_job_setup(vlgp_ctx *ctx, TrialSet *ts, int n_iter, int use_hessian, double eps, double lr,
                         double da_bound, double db_bound) {
    MstepJob &job = *job_of(ctx);
    int rc = VLGP_OK;
    DISPATCH_L(ctx->L, rc = mstep_setup_t<LT>(ctx, ts, job, n_iter, use_hessian, eps, lr, da_bound, db_bound));
    return rc;
}

int vlgp_mstep_job_remaining(vlgp_ctx *ctx) {
    const MstepJob *job = (const MstepJob *)ctx->mstep_job;
    return job ? job->n_iter - job->next_it : 0;
}

// Enqueue up to max_iters further Newton iterations on ctx->stream.
int vlgp_mstep_job_pump(vlgp_ctx *ctx, int max_iters) {
    MstepJob &job = *job_of(ctx);
    for (int k = 0; k < max_iters && job.next_it < job.n_iter; ++k) {
        int rc = VLGP_OK;
        DISPATCH_L(job.LT, rc = mstep_iter_t<LT>(ctx, job, job.next_it));
        if (rc) return rc;
        job.next_it++;
    }
    return VLGP_OK;
}

int vlgp_launch_mstep(vlgp_ctx *ctx, TrialSet *ts, int n_iter, int use_hessian, double eps, double lr,
                      double da_bound, double db_bound) {
    int rc = vlgp_mstep_job_setup(ctx, ts, n_iter, use_hessian, eps, lr, da_bound, db_bound);
    if (rc) return rc;
    return vlgp_mstep_job_pump(ctx, n_iter);
}
