// Shared declarations of libvlgp_b200: context, trial-set layout in HBM, error handling, small device helpers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <string>
#include <vector>

#include "../../include/vlgp_b200.h"

#define VLGP_MAX_L 16        // latents (register-array bound of the templated kernels)
#define VLGP_MAX_RANK 64     // rank of the prior factor (reference hard-codes 50, vlgp/preprocess.py:75)
#define VLGP_MAX_W 64        // window length handled by the SMEM-resident segment kernels (reference default 50)

// ---------------------------------------------------------------------------------------------------------------------
// HBM layout of one trial set (SURVEY.md section 7 "data model"): all bins concatenated, time-major.
// ---------------------------------------------------------------------------------------------------------------------
struct PriorFactor {           // one unique trial length
    int length = 0;
    double *d_G = nullptr;     // L x length x rank, natural row order, sigma already applied
    int *d_ncol = nullptr;     // L: number of leading non-zero columns (columns >= ncol are exactly zero)
    int *d_piv = nullptr;      // L x rank pivots (-1 padded)
    std::vector<int> h_ncol;
};

struct TrialSet {
    bool used = false;
    int n_trials = 0;
    int64_t nbin = 0;
    int max_len = 0, min_len = 0;
    std::vector<int> h_len;
    std::vector<int64_t> h_start;
    std::vector<int> h_fidx;           // trial -> index into factors
    int *d_len = nullptr;
    int64_t *d_start = nullptr;
    int *d_fidx = nullptr;
    std::vector<PriorFactor> factors;  // one per unique length
    double **d_Gptr = nullptr;         // device table: factor index -> d_G
    int **d_ncolptr = nullptr;         // device table: factor index -> d_ncol
    void *d_y = nullptr;
    int ydtype = VLGP_Y_F64;
    double *d_mu = nullptr, *d_v = nullptr, *d_w = nullptr, *d_dmu = nullptr;   // nbin x L
    double *d_ra = nullptr;            // nbin x L scratch: residual @ a^T
    double *d_u = nullptr;             // nbin scratch
    double *d_minv = nullptr;          // per-CTA scratch: grid x L x rank x rank
    int minv_grid = 0;
    // H-step
    double *d_M = nullptr;             // L x W x W second moments of mu
    double *d_K = nullptr;             // 2 x W x W: K and dK/dlog(omega) of the current evaluation
    double *d_hpart = nullptr;         // per-segment partials (2 x n_trials)
    double *d_hout = nullptr;          // MAX_L x 8 per-evaluation outputs + MAX_L x 2 reducible sums
    bool h_prepared = false;
    double h_nseg_total = 0.0;         // segments over all ranks
    double nbin_all_ranks = 0.0;       // bins over all ranks (M-step divisor); 0 until the first M-step asked for it
    int h_seg_grid = 1;
    bool h_geometry = false;
    int dmma_grid = 0;                 // cached launch geometry of the DMMA segment kernel
    double *d_mompart = nullptr;       // per-chunk partial second moments
};

// A batch of H-step objective evaluations (one per latent when the host optimisers run in lockstep), by value.
struct HEvalBatch {
    int n;
    int latent[VLGP_MAX_L];
    double sigmasq[VLGP_MAX_L], omega[VLGP_MAX_L], eps[VLGP_MAX_L];
};

struct NcclApi;   // dlopen'ed subset of NCCL (comm.cu)

struct vlgp_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t stream2 = nullptr;              // side stream for the latency-bound K^-1 kernel of the H-step
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    // The M-step reads only what the E-step left (mu, v, y, a, b) and the H-step only (mu, w): vlgp_mstep_begin puts
    // the M-step on its own stream (and, with several ranks, its own communicator) so that the two overlap.
    cudaStream_t stream_m = nullptr;
    cudaEvent_t ev_m_start = nullptr, ev_m_done = nullptr;
    bool mstep_pending = false;
    void *mstep_job = nullptr;                   // MstepJob (mstep.cu): the M-step being enqueued piecewise
    cudaDeviceProp prop{};
    std::string err;
    // model
    int N = 0, L = 0, rank = 0;
    double gp_noise = 1e-4, dt = 1.0;
    uint8_t *d_poisson = nullptr;
    bool any_gauss = false;
    double *d_a = nullptr, *d_b = nullptr, *d_noise = nullptr, *d_da = nullptr, *d_db = nullptr;
    std::vector<double> h_sigma, h_omega;
    std::vector<TrialSet> sets;
    // M-step scratch
    double *d_mpart = nullptr;   // grid x nstat x N
    double *d_mstat = nullptr;   // nstat x N (+ tail)
    double *d_ymom = nullptr;    // (L+1) x N : mu'y, sum y (constant during one M-step)
    void *d_ppack = nullptr;     // (L+1) x N double2: (a, a^2) and (b, 1/noise) pairs for the segment E-step
    int mpart_grid = 0;
    double *d_gshared = nullptr; // Gaussian-channel shared moments: L*L + 2L + 1
    int *d_flags = nullptr;      // device counters (failures etc.), 16 ints
    int *h_flags = nullptr;      // pinned
    double *h_pin = nullptr;     // pinned 4 KB staging for tiny D2H/H2D
    double *d_small = nullptr;   // 4 KB device staging
    void *h_stage[2] = {nullptr, nullptr};        // pinned double buffer of the y upload pipeline
    cudaEvent_t stage_ev[2] = {nullptr, nullptr};
    // comm
    NcclApi *nccl = nullptr;
    void *comm = nullptr;
    void *comm_m = nullptr;      // ncclCommSplit duplicate of comm for the overlapped M-step (null: no overlap when n_ranks > 1)
    void *shm = nullptr;         // host-side allreduce handle (shmcomm.cu) for the scalars the host consumes
    void *p2p = nullptr;         // P2PState (p2p.cu): peer-memory mailboxes for in-kernel allreduces, or null
    int p2p_chan = 0;            // channel of the current stream: 0 main, 1 overlapped M-step
    int rank_id = 0, n_ranks = 1;
    // measurement
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    int64_t counters[4] = {0, 0, 0, 0};
    int profile = 0;             // bit i set: time kernel class i with CUDA events (adds a sync per launch)
    double prof_ms[4] = {0, 0, 0, 0};
    int64_t prof_n[4] = {0, 0, 0, 0};
    cudaEvent_t pev0 = nullptr, pev1 = nullptr;
    void *d_flush = nullptr;
    size_t flush_bytes = 0;
};

int vlgp_fail(vlgp_ctx *ctx, int code, const char *fmt, ...);

// Trial-set buffers are allocated / freed in stream order from the device's default memory pool (release threshold
// raised in vlgp_create): a Session per vem() call then costs microseconds of allocator time instead of milliseconds.
template <typename T>
static inline cudaError_t vlgp_dalloc(vlgp_ctx *ctx, T **p, size_t bytes) {
    return cudaMallocAsync((void **)p, bytes, ctx->stream);
}
template <typename T>
static inline cudaError_t vlgp_dfree(vlgp_ctx *ctx, T *p) {
    return p ? cudaFreeAsync((void *)p, ctx->stream) : cudaSuccess;
}
int vlgp_allreduce_dev(vlgp_ctx *ctx, double *d_buf, size_t n, int op);   // comm.cu; no-op when n_ranks == 1
// capi.cu: enqueue up to max_iters further Newton iterations of a pending overlapped M-step (no-op otherwise); called
// by the H-step objective between its launches and its synchronisation, where the host would otherwise idle.
int vlgp_mstep_pump(vlgp_ctx *ctx, int max_iters);

#define CK(call)                                                                                              \
    do {                                                                                                      \
        cudaError_t e_ = (call);                                                                              \
        if (e_ != cudaSuccess)                                                                                \
            return vlgp_fail(ctx, VLGP_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #call,                 \
                             cudaGetErrorString(e_));                                                         \
    } while (0)

#define CKL()                                                                                                 \
    do {                                                                                                      \
        cudaError_t e_ = cudaGetLastError();                                                                  \
        if (e_ != cudaSuccess)                                                                                \
            return vlgp_fail(ctx, VLGP_ERR_CUDA, "%s:%d kernel launch -> %s", __FILE__, __LINE__,             \
                             cudaGetErrorString(e_));                                                         \
        ctx->counters[0]++;                                                                                   \
    } while (0)

#define REQUIRE(cond, ...)                                                                                    \
    do {                                                                                                      \
        if (!(cond)) return vlgp_fail(ctx, VLGP_ERR_ARG, __VA_ARGS__);                                        \
    } while (0)

static inline TrialSet *get_set(vlgp_ctx *ctx, int id) {
    if (!ctx || id < 0 || id >= (int)ctx->sets.size() || !ctx->sets[id].used) return nullptr;
    return &ctx->sets[id];
}

struct ProfScope {   // accumulates device time of one kernel class when profiling is enabled
    vlgp_ctx *ctx;
    int which;
    ProfScope(vlgp_ctx *c, int w) : ctx(c), which(w) {
        if (ctx->profile & (1 << which)) cudaEventRecord(ctx->pev0, ctx->stream);
    }
    ~ProfScope() {
        if (ctx->profile & (1 << which)) {
            cudaEventRecord(ctx->pev1, ctx->stream);
            cudaEventSynchronize(ctx->pev1);
            float ms = 0.f;
            cudaEventElapsedTime(&ms, ctx->pev0, ctx->pev1);
            ctx->prof_ms[which] += ms;
            ctx->prof_n[which] += 1;
        }
    }
};

// ---------------------------------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ double load_y(const void *y, int ydtype, int64_t idx) {
    return ydtype == VLGP_Y_U8 ? (double)((const uint8_t *)y)[idx] : ((const double *)y)[idx];
}

__device__ __forceinline__ double warp_sum(double x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    return x;
}

// exp(min(x, 10)) -- vlgp/math.py:24-38.  Branch-free (the library exp() carries a slow-path branch for out-of-range
// arguments that keeps the compiler from interleaving several evaluations; this link function is the single most
// executed operation of the E- and M-step): x = t ln2 + r with t = rint(x log2 e) via the 2^52+2^51 shift, |r| <= ln2/2,
// degree-13 Taylor polynomial in Horner form (truncation 4e-18), scaling by 2^t through the exponent field.
// Maximum error 1 ulp on [-708, 10] (checked against numpy.exp on 4e6 points, scripts/check_exp.py).  Arguments
// below -708 are clamped: the result is then 3e-308 instead of a denormal/0, an absolute difference of 3e-308.
//
// The Taylor coefficients 1/13! .. 1/2!, 1 live in constant memory: written as literals, ptxas re-materialised every one
// of them with two UMOV per loop iteration (24 of the 156 instructions of the E-step's two-neuron rate-pass body, as
// many in the M-step kernel); from the constant bank they are loaded into uniform registers once, outside the loops
// (136 instructions; same FP64 instructions in the same order, so results are unchanged bit for bit).
static __constant__ double VLGP_EXP_C[13] = {1.6059043836821613e-10, 2.08767569878681e-09, 2.505210838544172e-08,
                                             2.755731922398589e-07,  2.7557319223985893e-06, 2.48015873015873e-05,
                                             1.984126984126984e-04,  1.388888888888889e-03,  8.333333333333333e-03,
                                             4.1666666666666664e-02, 1.6666666666666666e-01, 0.5, 1.0};

__device__ __forceinline__ double trunc_exp(double x) {
    x = x > 10.0 ? 10.0 : x;                                       // plain compare-select: no NaN-propagating min/max
    x = x < -708.0 ? -708.0 : x;
    const double shift = 6755399441055744.0;                       // 2^52 + 2^51
    const double tmp = fma(x, 1.4426950408889634, shift);
    const int ti = __double2loint(tmp);                            // rint(x log2 e) in the low word
    const double t = tmp - shift;
    double r = fma(t, -6.93147180369123816490e-01, x);             // ln2 high part
    r = fma(t, -1.90821492927058770002e-10, r);                    // ln2 low part
    double p = VLGP_EXP_C[0];                                      // 1/13!
#pragma unroll
    for (int k = 1; k < 13; ++k) p = fma(p, r, VLGP_EXP_C[k]);     // 1/12! ... 1/2!, 1
    p = fma(p, r, 1.0);
    return p * __hiloint2double((ti + 1023) << 20, 0);             // 2^t, t in [-1022, 15]
}

// Two independent evaluations with their Horner chains interleaved statement by statement (the polynomial is a chain
// of 14 dependent DFMAs; two chains in flight double the FP64-pipe utilisation of a warp -- when ptxas keeps them
// interleaved, which it does not under an 80-register cap).  Bitwise identical to two calls of
// trunc_exp.
__device__ __forceinline__ void trunc_exp2(double x0, double x1, double &e0, double &e1) {
    x0 = x0 > 10.0 ? 10.0 : x0;
    x1 = x1 > 10.0 ? 10.0 : x1;
    x0 = x0 < -708.0 ? -708.0 : x0;
    x1 = x1 < -708.0 ? -708.0 : x1;
    const double shift = 6755399441055744.0;
    const double m0 = fma(x0, 1.4426950408889634, shift), m1 = fma(x1, 1.4426950408889634, shift);
    const int i0 = __double2loint(m0), i1 = __double2loint(m1);
    const double t0 = m0 - shift, t1 = m1 - shift;
    double r0 = fma(t0, -6.93147180369123816490e-01, x0), r1 = fma(t1, -6.93147180369123816490e-01, x1);
    r0 = fma(t0, -1.90821492927058770002e-10, r0);
    r1 = fma(t1, -1.90821492927058770002e-10, r1);
    double p0 = VLGP_EXP_C[0], p1 = VLGP_EXP_C[0];
#pragma unroll
    for (int k = 1; k < 13; ++k) {
        p0 = fma(p0, r0, VLGP_EXP_C[k]);
        p1 = fma(p1, r1, VLGP_EXP_C[k]);
    }
    p0 = fma(p0, r0, 1.0);
    p1 = fma(p1, r1, 1.0);
    e0 = p0 * __hiloint2double((i0 + 1023) << 20, 0);
    e1 = p1 * __hiloint2double((i1 + 1023) << 20, 0);
}

// 1 / d for a normal, finite d (sweep pivots: >= 1 for I + PSD matrices, > 0 otherwise) without the special-case
// handling of IEEE division: MUFU seed (about 20 bits) + two Newton steps, error <= 1 ulp, ~60 cycles of latency
// instead of ~300 on the critical path of every sweep step.
__device__ __forceinline__ double fast_rcp(double d) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
    double e = fma(-d, y, 1.0);
    y = fma(y, e, y);
    e = fma(-d, y, 1.0);
    y = fma(y, e, y);
    return y;
}

__device__ __forceinline__ double clipd(double x, double bound) { return fmin(fmax(x, -bound), bound); }
#endif
