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
#define VLGP_MAX_XDIM 8      // regressors per neuron (xdim = max(history, 1), vlgp/preprocess.py:59)
#define VLGP_MAX_W 64        // window length handled by the SMEM-resident segment kernels (reference default 50)
#define VLGP_MAX_W_H 160     // window length of the H-step's wide path (hstep_wide.cu: one W x W matrix in shared memory)

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
    // general regressors (regress.cu): null for the all-ones bias column
    double *d_x = nullptr;             // nbin x xdim x N
    double *d_xb = nullptr;            // nbin x N : einsum(x, b), the offset of the linear predictor
    // long-trial E-step (estep_long.cu): item tables, partial sums and inverses, allocated at the first call
    bool k3_ready = false;
    int k3_items = 0;
    int *d_k3_tab = nullptr, *d_k3_bad = nullptr;
    double *d_k3_buf = nullptr;
    // state prefetch (vlgp_trials_prefetch_state): mu | v | w | dmu copied to ctx->h_prefetch behind the E-step; valid
    // while no entry point has written the state since (state_version)
    uint64_t state_version = 0, prefetch_version = ~0ull;
    int prefetch_mask = 0;
    double *prefetch_ext[4] = {nullptr, nullptr, nullptr, nullptr};   // caller-owned pinned destinations (or the context's area)
    int gen = 0;                       // generation of the slot: stale handles of freed sets are refused
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
    int xdim = 1;                                // regressors per neuron; b, db are xdim x N
    int estep_f32 = 0;                           // vlgp_set_precision: single-precision rate passes in the segment E-step
    double *d_bpart = nullptr, *d_bstat = nullptr;   // regression statistics of the general-x M-step (regress.cu)
    size_t bpart_len = 0;
    int set_gen = 0;
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
    int *d_smslots = nullptr;    // one counter per SM: the resident CTAs of the segment E-step learn their slot on the SM
    int *h_flags = nullptr;      // pinned
    double *h_pin = nullptr;     // pinned 4 KB staging for tiny D2H/H2D
    double *d_small = nullptr;   // 4 KB device staging
    void *h_prefetch = nullptr;                   // pinned: one set's mu | v | w | dmu, filled by a copy stream
    size_t prefetch_cap = 0;
    int prefetch_set = -1;                        // handle of the set it belongs to
    cudaStream_t stream_copy = nullptr;
    cudaEvent_t ev_prefetch_go = nullptr, ev_prefetch_done = nullptr;
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
    // 256 bytes of slack: bulk copies (TMA) round the tail of a range up to 16 bytes
    return cudaMallocAsync((void **)p, bytes + 256, ctx->stream);
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

// A set handle is slot + (generation << 12): a handle that outlived its set (vlgp_trials_free, or vlgp_set_model, which
// drops every set) never resolves to whatever occupies the slot now.
static inline TrialSet *get_set(vlgp_ctx *ctx, int id) {
    if (!ctx || id < 0) return nullptr;
    const int slot = id & 0xfff, gen = id >> 12;
    if (slot >= (int)ctx->sets.size() || !ctx->sets[slot].used || ctx->sets[slot].gen != gen) return nullptr;
    return &ctx->sets[slot];
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

// exp(min(x, 10)) -- vlgp/math.py:24-38.  Branch-free and table-driven (the library exp() carries a slow-path branch
// that keeps the compiler from interleaving several evaluations, and this link function is the single most executed
// operation of the E- and M-step): x = m ln2/32 + r with m = rint(x 32/ln2) via the 2^52+2^51 shift, |r| <= ln2/64, so
// exp(x) = 2^(m >> 5) T[m & 31] exp(r) with T[j] = 2^(j/32) and a degree-6 Taylor polynomial for exp(r) (truncation
// 3e-18).  11 FP64-pipe instructions per value against 19 for the table-free degree-13 form used before; maximum error
// 0.9 ulp on [-700, 10] against a 50-digit reference (scripts/check_exp.py).  Arguments below -708 are clamped: the
// result is then ~3e-308 instead of a denormal / 0, an absolute difference of 3e-308.
static __constant__ double VLGP_EXP_C[6] = {1.0 / 720.0, 1.0 / 120.0, 1.0 / 24.0, 1.0 / 6.0, 0.5, 1.0};
static __device__ const double VLGP_EXP_T[32] = {
    1.00000000000000000e+00, 1.02189714865411663e+00, 1.04427378242741375e+00, 1.06714040067682370e+00,
    1.09050773266525769e+00, 1.11438674259589243e+00, 1.13878863475669156e+00, 1.16372485877757748e+00,
    1.18920711500272103e+00, 1.21524735998046896e+00, 1.24185781207348400e+00, 1.26905095719173322e+00,
    1.29683955465100964e+00, 1.32523664315974132e+00, 1.35425554693689265e+00, 1.38390988196383202e+00,
    1.41421356237309515e+00, 1.44518080697704665e+00, 1.47682614593949935e+00, 1.50916442759342284e+00,
    1.54221082540794074e+00, 1.57598084510788650e+00, 1.61049033194925428e+00, 1.64575547815396495e+00,
    1.68179283050742900e+00, 1.71861929812247793e+00, 1.75625216037329945e+00, 1.79470907500310717e+00,
    1.83400808640934243e+00, 1.87416763411029996e+00, 1.91520656139714740e+00, 1.95714412417540018e+00};
#define VLGP_EXP_INV 4.61662413084468283841e+01      // 32 / ln2
#define VLGP_EXP_HI 2.16608493865351192653e-02       // ln2 / 32, upper 32 bits
#define VLGP_EXP_LO 5.96317165397058656257e-12       // ln2 / 32 - HI

__device__ __forceinline__ double trunc_exp(double x) {
    x = x > 10.0 ? 10.0 : x;                                       // plain compare-select: no NaN-propagating min/max
    x = x < -708.0 ? -708.0 : x;
    const double shift = 6755399441055744.0;                       // 2^52 + 2^51
    const double tmp = fma(x, VLGP_EXP_INV, shift);
    const int mi = __double2loint(tmp);                            // rint(x 32 / ln2) in the low word
    const double t = tmp - shift;
    double r = fma(t, -VLGP_EXP_HI, x);
    r = fma(t, -VLGP_EXP_LO, r);
    double p = VLGP_EXP_C[0];
#pragma unroll
    for (int k = 1; k < 6; ++k) p = fma(p, r, VLGP_EXP_C[k]);      // 1/120 ... 1/2, 1
    p *= r;                                                        // exp(r) - 1
    const double tj = __ldg(&VLGP_EXP_T[mi & 31]);
    const double e = fma(tj, p, tj);
    const int ex = max(mi >> 5, -1022);
    return __hiloint2double(__double2hiint(e) + (ex << 20), __double2loint(e));      // e 2^ex through the exponent field
}

// Two independent evaluations with their chains interleaved statement by statement.  Bitwise identical to two calls
// of trunc_exp.
__device__ __forceinline__ void trunc_exp2(double x0, double x1, double &e0, double &e1) {
    x0 = x0 > 10.0 ? 10.0 : x0;
    x1 = x1 > 10.0 ? 10.0 : x1;
    x0 = x0 < -708.0 ? -708.0 : x0;
    x1 = x1 < -708.0 ? -708.0 : x1;
    const double shift = 6755399441055744.0;
    const double m0 = fma(x0, VLGP_EXP_INV, shift), m1 = fma(x1, VLGP_EXP_INV, shift);
    const int i0 = __double2loint(m0), i1 = __double2loint(m1);
    const double tj0 = __ldg(&VLGP_EXP_T[i0 & 31]), tj1 = __ldg(&VLGP_EXP_T[i1 & 31]);
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
}

// Four independent evaluations, statement-interleaved (two accumulator tiles of the tensor-path rate pass at a time).
__device__ __forceinline__ void trunc_exp4(double x0, double x1, double x2, double x3, double &e0, double &e1, double &e2,
                                           double &e3) {
    double x[4] = {x0, x1, x2, x3}, m[4], t[4], r[4], tj[4], p[4];
    int mi[4];
    const double shift = 6755399441055744.0;
#pragma unroll
    for (int i = 0; i < 4; ++i) x[i] = x[i] > 10.0 ? 10.0 : x[i];
#pragma unroll
    for (int i = 0; i < 4; ++i) x[i] = x[i] < -708.0 ? -708.0 : x[i];
#pragma unroll
    for (int i = 0; i < 4; ++i) m[i] = fma(x[i], VLGP_EXP_INV, shift);
#pragma unroll
    for (int i = 0; i < 4; ++i) mi[i] = __double2loint(m[i]);
#pragma unroll
    for (int i = 0; i < 4; ++i) tj[i] = __ldg(&VLGP_EXP_T[mi[i] & 31]);
#pragma unroll
    for (int i = 0; i < 4; ++i) t[i] = m[i] - shift;
#pragma unroll
    for (int i = 0; i < 4; ++i) r[i] = fma(t[i], -VLGP_EXP_HI, x[i]);
#pragma unroll
    for (int i = 0; i < 4; ++i) r[i] = fma(t[i], -VLGP_EXP_LO, r[i]);
#pragma unroll
    for (int i = 0; i < 4; ++i) p[i] = VLGP_EXP_C[0];
#pragma unroll
    for (int k = 1; k < 6; ++k) {
#pragma unroll
        for (int i = 0; i < 4; ++i) p[i] = fma(p[i], r[i], VLGP_EXP_C[k]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) p[i] *= r[i];
#pragma unroll
    for (int i = 0; i < 4; ++i) p[i] = fma(tj[i], p[i], tj[i]);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int ex = max(mi[i] >> 5, -1022);
        p[i] = __hiloint2double(__double2hiint(p[i]) + (ex << 20), __double2loint(p[i]));
    }
    e0 = p[0];
    e1 = p[1];
    e2 = p[2];
    e3 = p[3];
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
