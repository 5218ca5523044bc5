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
int vlgp_allreduce_dev(vlgp_ctx *ctx, double *d_buf, size_t n, int op);   // comm.cu; no-op when n_ranks == 1

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

// exp(min(x, 10)) -- vlgp/math.py:24-38
__device__ __forceinline__ double trunc_exp(double x) { return exp(fmin(x, 10.0)); }

__device__ __forceinline__ double clipd(double x, double bound) { return fmin(fmax(x, -bound), bound); }
#endif
