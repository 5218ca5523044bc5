/*
 * vlgp_b200.h -- C ABI of libvlgp_b200.so, the B200 (sm_100a) variational-EM engine behind vlgp.fit().
 *
 * The reference (catniplab/vlgp) has no FFI: its seam is a set of Python functions with the signature
 * (trials, params, config) that mutate three dicts in place (vlgp/api.py:7-11 binds them by name).  Each entry point
 * below replaces the arithmetic of one of those functions; the Python host (vlgp_b200/*.py) packs the dicts into the
 * flat buffers these calls take and unpacks the results into the same dict keys.  All entry points are batched over
 * trials x latents, take plain pointers and sizes, never throw, and return 0 on success or a negative vlgp_status
 * (message via vlgp_last_error).  Host pointers are borrowed for the duration of the call; device memory is owned by
 * the context.  One host thread drives one context; one context drives one GPU.
 *
 * Layout conventions (identical to the reference's NumPy arrays, C-contiguous, float64 on the host):
 *   bins of all trials of a "trial set" are concatenated in trial order: nbin = sum(lengths)
 *   y            nbin x N          (float64, or uint8 when every count is an integer in [0,255])
 *   mu,v,w,dmu   nbin x L
 *   a, da        L x N             (params["a"], params["da"])
 *   b, db        xdim x N          (params["b"]; xdim == 1 with an all-ones regressor by default, vlgp/preprocess.py:43-44)
 *   noise        N
 *   G            L x length x rank (params["cholesky"][length], vlgp/gp.py:150-162)
 */
#ifndef VLGP_B200_H
#define VLGP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define VLGP_API __attribute__((visibility("default")))
#else
#define VLGP_API
#endif

typedef struct vlgp_ctx vlgp_ctx;

enum vlgp_status {
    VLGP_OK = 0,
    VLGP_ERR_ARG = -1,      /* bad argument / call order */
    VLGP_ERR_CUDA = -2,     /* CUDA runtime error (message has the call site) */
    VLGP_ERR_NCCL = -3,     /* NCCL error or libnccl could not be loaded */
    VLGP_ERR_NOMEM = -4,
    VLGP_ERR_UNSUPPORTED = -5
};

enum vlgp_ydtype { VLGP_Y_F64 = 0, VLGP_Y_U8 = 1 };

/* ---- lifecycle ---------------------------------------------------------------------------------------------------- */
VLGP_API int vlgp_create(int device, vlgp_ctx **out);
VLGP_API int vlgp_destroy(vlgp_ctx *ctx);
/* Last error message of this context (or of the failed vlgp_create when ctx == NULL). */
VLGP_API const char *vlgp_last_error(const vlgp_ctx *ctx);
VLGP_API int vlgp_device_info(vlgp_ctx *ctx, int *sm_count, int *cc_major, int *cc_minor, uint64_t *total_mem, char name[128]);
VLGP_API int vlgp_sync(vlgp_ctx *ctx);
/* CUDA-event timer on the context's stream (the stream every kernel of this library is launched on). */
VLGP_API int vlgp_timer_start(vlgp_ctx *ctx);
VLGP_API int vlgp_timer_stop(vlgp_ctx *ctx, float *elapsed_ms);
/* counters[0] = kernels launched by this library since vlgp_create, [1] = r x r SPD solves (E-step),
 * [2] = W x W factorisations (H-step), [3] = NCCL collectives issued. */
VLGP_API int vlgp_counters(vlgp_ctx *ctx, int64_t counters[4]);

/* ---- model: replaces params["a","b","noise","sigma","omega","likelihood","rank","gp_noise","dt"] ---------------- */
/* get_params, vlgp/preprocess.py:49-81.  poisson_mask[n] != 0 -> Poisson channel, 0 -> Gaussian. */
VLGP_API int vlgp_set_model(vlgp_ctx *ctx, int n_neurons, int n_latents, int rank, const uint8_t *poisson_mask,
                   double gp_noise, double dt);
/* Any pointer may be NULL (left unchanged / not returned). */
VLGP_API int vlgp_set_params(vlgp_ctx *ctx, const double *a, const double *b, const double *noise, const double *sigma,
                    const double *omega);
VLGP_API int vlgp_get_params(vlgp_ctx *ctx, double *a, double *b, double *noise, double *da, double *db, double *sigma,
                    double *omega);

/* Number of regressors per neuron, xdim = max(history, 1) (vlgp/preprocess.py:59): b and db become xdim x N (row-major,
 * params["b"]).  Call right after vlgp_set_model (which resets xdim to 1) and before creating trial sets. */
VLGP_API int vlgp_set_regressors(vlgp_ctx *ctx, int xdim);

/* ---- trial sets: replaces the list of trial dicts ------------------------------------------------------------------ */
VLGP_API int vlgp_trials_create(vlgp_ctx *ctx, int n_trials, const int32_t *lengths, int *set_id);
VLGP_API int vlgp_trials_free(vlgp_ctx *ctx, int set_id);
VLGP_API int vlgp_trials_set_y(vlgp_ctx *ctx, int set_id, const void *y, int ydtype);
/* Same from per-trial blocks (parts[i]: rows[i] x N, C-contiguous, dtype src_dtype) without a host-side concatenation:
 * blocks are converted/copied into pinned staging by host threads and uploaded in a double-buffered pipeline.  float64
 * blocks holding only integer counts in [0,255] are stored as uint8; *stored_dtype returns the choice. */
VLGP_API int vlgp_trials_set_y_parts(vlgp_ctx *ctx, int set_id, int n_parts, const void *const *parts, const int64_t *rows,
                            int src_dtype, int *stored_dtype);
/* General regressors x (nbin x xdim x N float64, the trial dicts' "x" concatenated, vlgp/core.py:66): only needed when x
 * is not the all-ones bias column; the E-/M-step then use eta = mu a + einsum(x, b) and update b with the design x
 * (vlgp/core.py:205-220,229-235; csrc/regress.cu). */
VLGP_API int vlgp_trials_set_x(vlgp_ctx *ctx, int set_id, const double *x);
VLGP_API int vlgp_trials_set_state(vlgp_ctx *ctx, int set_id, const double *mu, const double *v, const double *w);
/* Initial posterior means on the device (vlgp/preprocess.py:36-41: mu = FactorAnalysis.transform(y) per trial):
 * mu[bin] = ((y[bin] - mean) P) Cz with mean (N), P = Wpsi' (N x L) and Cz = cov_z (L x L) of the factor model fitted
 * on the host (sklearn FactorAnalysis.transform evaluates the two products in this order). */
/* Start copying the listed state arrays (bit 0 mu, 1 v, 2 w, 3 dmu) to pinned host memory on a copy stream, behind what is
 * enqueued so far; a later vlgp_trials_get_state_parts of one of them is served from that copy unless the set's state has
 * been written in between.  vem() calls this after the E-step of its last iteration (vlgp/core.py:307-326 downloads
 * nothing: its trial dicts ARE the state), so the transfer runs under the M- and H-step. */
VLGP_API int vlgp_trials_prefetch_state(vlgp_ctx *ctx, int set_id, int which_mask);
/* The same with caller-owned page-locked destinations: dst[k] != NULL sends array k (nbin x L doubles) straight into that
 * block instead of the context's staging area.  vlgp_trials_prefetch_take waits for the copy and reports whether block k
 * holds the set's CURRENT array (nothing has written the state since): the host code then hands views of the block out
 * as trial["w"] / trial["dmu"] -- the reference rebinds those keys to new arrays in every E-step (vlgp/core.py:117-120)
 * -- with no further copy.  vlgp_host_alloc / vlgp_host_free: page-locked blocks for that purpose. */
VLGP_API int vlgp_trials_prefetch_state_into(vlgp_ctx *ctx, int set_id, int which_mask, double *const *dst);
VLGP_API int vlgp_trials_prefetch_take(vlgp_ctx *ctx, int set_id, int which, int *valid);
VLGP_API int vlgp_trials_prefetch_wait(vlgp_ctx *ctx, int set_id);   /* the prefetch has landed; its blocks may be reused */
VLGP_API int vlgp_host_alloc(void **p, size_t bytes);
VLGP_API int vlgp_host_free(void *p);
VLGP_API int vlgp_trials_project_y(vlgp_ctx *ctx, int set_id, const double *mean, const double *P, const double *Cz);
/* Per-trial-block variants (which: 0 mu, 1 v, 2 w, 3 dmu [get only]; parts[i]: rows[i] x L float64, C-contiguous):
 * gather / scatter through the pinned double-buffered pipeline, so segment views are read and written in place. */
VLGP_API int vlgp_trials_set_state_parts(vlgp_ctx *ctx, int set_id, int which, int n_parts, const double *const *parts,
                                const int64_t *rows);
VLGP_API int vlgp_trials_get_state_parts(vlgp_ctx *ctx, int set_id, int which, int n_parts, double *const *parts,
                                const int64_t *rows);
VLGP_API int vlgp_trials_get_state(vlgp_ctx *ctx, int set_id, double *mu, double *v, double *w, double *dmu);

/* ---- prior factor: gp.make_cholesky (vlgp/gp.py:150-162) over math.ichol_gauss (vlgp/math.py:76-126) ------------- */
VLGP_API int vlgp_make_cholesky(vlgp_ctx *ctx, int set_id);
/* G: L x length x rank; pivots: L x rank (-1 padded); ncol: L.  Any of the three may be NULL. */
VLGP_API int vlgp_get_cholesky(vlgp_ctx *ctx, int set_id, int length, double *G, int32_t *pivots, int32_t *ncol);
/* Inject a factor (tests: decouple E-step parity from pivot parity). */
VLGP_API int vlgp_set_cholesky(vlgp_ctx *ctx, int set_id, int length, const double *G);

/* ---- E-step: core.estep / infer_single_trial (vlgp/core.py:22-126), update_w (:419-442), update_v (:445-471) ------ */
/* n_failed: number of (trial, latent, iteration) r x r systems that were not positive definite (their update is
 * skipped exactly like the reference's except-branches, vlgp/core.py:92-94,112). */
VLGP_API int vlgp_estep(vlgp_ctx *ctx, int set_id, int n_iter, double dmu_bound, int method_vb, int *n_failed);
/* The same on n_segments listed members of the set only (each index once; the others are untouched).  With the two
 * row operations below it reproduces the reference's E-step on OVERLAPPING windows, whose segments are views of one
 * trial that are updated in place one after the other (vlgp/util.py:482-498, vlgp/core.py:96-97,112,123-126): the host
 * runs the segments level by level along each chain of overlapping windows (vlgp_b200/core.py::_Aliasing). */
VLGP_API int vlgp_estep_subset(vlgp_ctx *ctx, int set_id, int n_iter, double dmu_bound, int method_vb,
                               const int32_t *segments, int n_segments, int *n_failed);
/* For i < n: bin dst[i] <- bin src[i] in the per-bin arrays selected by which_mask (bit 0 mu, 1 v, 2 w, 3 dmu); bins
 * are indices into the set's concatenated bins.  No bin may be written twice or be both read and written. */
VLGP_API int vlgp_trials_copy_rows(vlgp_ctx *ctx, int set_id, int which_mask, const int64_t *src, const int64_t *dst,
                                   int64_t n);
VLGP_API int vlgp_update_w(vlgp_ctx *ctx, int set_id);
VLGP_API int vlgp_update_v(vlgp_ctx *ctx, int set_id, int *n_failed);

/* ---- M-step: core.mstep (vlgp/core.py:129-249) --------------------------------------------------------------------- */
/* n_fallback: number of per-neuron Newton systems that fell back to the gradient step (vlgp/core.py:194-198). */
VLGP_API int vlgp_mstep(vlgp_ctx *ctx, int set_id, int n_iter, int use_hessian, double eps, double learning_rate,
               double da_bound, double db_bound, int *n_fallback);
/* The same M-step split in two so that the host can drive the H-step while it runs: vem's M-step (vlgp/core.py:318-320)
 * reads mu, v, y and writes a, b, noise; its H-step (:322-325) reads mu, w and writes sigma, omega -- independent given
 * the E-step.  _begin enqueues the whole M-step on a second stream (and, with several ranks, a second communicator)
 * and returns; _end waits for it.  Calls that touch what the M-step reads or writes wait for it by themselves. */
VLGP_API int vlgp_mstep_begin(vlgp_ctx *ctx, int set_id, int n_iter, int use_hessian, double eps, double learning_rate,
                     double da_bound, double db_bound);
VLGP_API int vlgp_mstep_end(vlgp_ctx *ctx, int *n_fallback);

/* ---- H-step objective: gp.construct_posterior_cov + gp.elbo (vlgp/gp.py:12-62,126-147) --------------------------- */
/* Once per H-step (mu, w fixed during it): per-latent second moments of mu over the segments (all of length W).
 * W <= 160: windows of up to 56 bins run on the FP64 tensor path (one warp per segment, bordered variant when W is one or
 * two bins beyond a multiple of 8), 57..64 on a register sweep, 65..160 on the shared-memory kernels of
 * csrc/hstep_wide.cu; VLGP_ERR_ARG beyond (the reference itself takes any window, vlgp/gp.py:65-123). */
VLGP_API int vlgp_hstep_prepare(vlgp_ctx *ctx, int set_id);
/* hyper = (sigma^2, omega, eps) (already exponentiated).  Returns ll and d ll / d log(omega) (the only slot the
 * reference's mask [0,1,0] keeps, vlgp/gp.py:16,85).  info != 0: K is not positive definite (vlgp/gp.py:17-20,132). */
VLGP_API int vlgp_hstep_objective(vlgp_ctx *ctx, int set_id, int latent, const double hyper[3], double *ll, double *dll,
                         int *info);

/* n evaluations (latents[e], hypers[3e..3e+2]) in one pass over the segments: one set of launches, one allreduce and
 * one synchronisation for the whole batch (the host runs the per-latent L-BFGS-B optimisers in lockstep). n <= 16. */
VLGP_API int vlgp_hstep_objective_batch(vlgp_ctx *ctx, int set_id, int n, const int32_t *latents, const double *hypers,
                               double *ll, double *dll, int32_t *info);

/* The whole H-step optimisation of n_lat latents in one call: replaces the per-latent scipy L-BFGS-B runs of
 * gp.optimize / gp.optimze1d (vlgp/gp.py:82-92,100-123).  Runs vlgp_hstep_prepare, then the restated L-BFGS-B
 * (csrc/lbfgsb.cuh; scipy's defaults m = 10, factr = 1e7, pgtol = 1e-5, maxls = 20) of every latent in lockstep, one
 * batched objective pass per round.  All vectors are in the optimiser's variables x = log(sigma^2, omega, eps):
 * log_initial n_lat x 3, log_bounds 3 x 2 (lower, upper), mask 3 (gradient mask, the reference uses [0,1,0]),
 * log_result n_lat x 3.  collapse_tol > 0 ends a line search whose whole bracket lies within that distance of its
 * starting point (lbfgsb.cuh; 0 = the full scipy sequence).  Optional outputs: fval (minus ELBO at the result), nfev
 * (evaluations the optimiser asked for), task (lbfgsb::Task termination code), n_rounds (batched device passes). */
VLGP_API int vlgp_hstep_optimize(vlgp_ctx *ctx, int set_id, int n_lat, const int32_t *latents, const double *log_initial,
                                 const double *log_bounds, const int32_t *mask, double collapse_tol, double *log_result,
                                 double *fval, int32_t *nfev, int32_t *task, int32_t *n_rounds);
/* The restated L-BFGS-B in reverse communication, context-free (no device): tests drive it next to scipy's setulb on
 * identical objectives.  _advance: pass f, g at the x returned by the previous call (ignored on the first call); returns
 * 1 when f, g are needed at x, 0 when finished (*task = termination code), < 0 on bad arguments. */
VLGP_API int vlgp_lbfgsb_new(int n, const double *x0, const double *lower, const double *upper, double factr, double pgtol,
                             int maxls, double collapse_tol, void **handle);
VLGP_API int vlgp_lbfgsb_advance(void *handle, double f, const double *g, double *x, int *task);
VLGP_API int vlgp_lbfgsb_info(void *handle, double *f, int *nfev, int *nit, int *n_collapsed);
VLGP_API int vlgp_lbfgsb_free(void *handle);

/* ---- full posterior covariance: the T x T algebra of api.sample_posterior / util.posterior_cov ------------------------
 * cov (T x T, host) = inv(inv(G G' + reg I) + diag(w)) for one latent of one member of the set (vlgp/api.py:160-166;
 * reg = 0: K - K (1/W + K)^-1 K of vlgp/util.py:541-547), from the set's prior factor and its weights w, through the
 * rank-r identity of csrc/postcov.cu (one r x r inverse instead of two T x T ones). */
VLGP_API int vlgp_posterior_cov(vlgp_ctx *ctx, int set_id, int trial, int latent, double reg, double *cov);

/* ---- constraints and convergence bookkeeping (vlgp/core.py:300-305,350-354,366-416) ------------------------------ */
/* mu <- (mu - shift) @ M for every bin; shift (L) and M (L x L, row-major) may be NULL (0 / identity). */
VLGP_API int vlgp_latent_affine(vlgp_ctx *ctx, int set_id, const double *shift, const double *M);
/* The same map on the n_rows listed bins only (each once): the reference's constrain_loading / constrain_latent walk the
 * segment list and rescale every segment's mu in place (vlgp/core.py:384-389,414-416), so a bin shared by two
 * overlapping windows is mapped once per window holding it. */
VLGP_API int vlgp_latent_affine_rows(vlgp_ctx *ctx, int set_id, const double *shift, const double *M,
                                     const int64_t *rows, int64_t n_rows);
/* out[0] = sum mu^2, out[1] = sum dmu^2 over all bins and latents (summed over ranks when a communicator is set). */
VLGP_API int vlgp_norms(vlgp_ctx *ctx, int set_id, double out[2]);
/* Per-latent sum(mu), sum(mu^2) and the bin count (summed over ranks when a communicator is set). */
VLGP_API int vlgp_latent_moments(vlgp_ctx *ctx, int set_id, double *sum, double *sumsq, int64_t *count);

/* ---- host packing (context-free, no device needed) ------------------------------------------------------------------
 * The reference hands y over as float64 (it promotes whatever arrives, vlgp/core.py:60); spike counts are stored in HBM
 * as uint8 when every entry is an integer in [0, 255].  This is the conversion + exactness check that the host threads of
 * vlgp_trials_set_y_parts run on disjoint ranges: dst[k] = (uint8) src[k]; returns 1 when every entry was such a count
 * (dst is then exact), 0 otherwise (dst is then unspecified).  vlgp_host_pack_isa: 1 when the AVX2 body is in use. */
VLGP_API int vlgp_host_f64_to_u8(const double *src, unsigned char *dst, int64_t cnt);
VLGP_API int vlgp_host_pack_isa(void);
/* Self-test of the persistent host thread pool those threads come from (hostpack.cpp): `rounds` parallel calls of
 * n_tasks tasks; returns 0 when every task of every call ran exactly once before its call returned. */
VLGP_API int vlgp_host_pool_selftest(int n_tasks, int rounds, int *workers);

/* ---- multi-GPU: one process per GPU, NCCL sum-allreduce of the M-/H-step sufficient statistics ------------------- */
VLGP_API int vlgp_comm_unique_id(vlgp_ctx *ctx, const char *libnccl_path, char id[128]);
VLGP_API int vlgp_comm_init(vlgp_ctx *ctx, const char *libnccl_path, int rank, int n_ranks, const char id[128]);
/* In-place allreduce of a small host buffer through the device (op: 0 = sum, 1 = max). */
VLGP_API int vlgp_comm_allreduce(vlgp_ctx *ctx, double *buf, int n, int op);
/* In-place sum-allreduce of a large host array (staged through a temporary device buffer): used by the SPMD fit() to
 * give every rank the posterior of every trial. */
VLGP_API int vlgp_comm_allreduce_bulk(vlgp_ctx *ctx, double *buf, int64_t n);

/* ---- host-side allreduce between the processes of one node (POSIX shared memory; no GPU involved) ------------------
 * The scalars the HOST consumes every round -- (ll, dll) of each H-step objective evaluation for L-BFGS-B
 * (vlgp/gp.py:107-114), the norms of vem's convergence test (vlgp/core.py:300-305,350-354) -- are summed over ranks here
 * instead of through an NCCL launch + device round trip.  name: "/something", the same on every rank of the job; rank 0
 * creates the segment.  op: 0 sum, 1 max; n <= 256.  Results are bit-identical on every rank (fixed rank order).
 * vlgp_comm_attach_shm hands the handle to a context (which closes it on destroy): from then on vlgp_comm_allreduce,
 * vlgp_norms, vlgp_latent_moments and the H-step objective use it; every rank must attach, or none. */
VLGP_API int vlgp_shm_open(const char *name, int rank, int n_ranks, void **handle);
VLGP_API int vlgp_shm_allreduce(void *handle, double *buf, int n, int op);
VLGP_API int vlgp_shm_close(void *handle, int unlink_name);
VLGP_API int vlgp_comm_attach_shm(vlgp_ctx *ctx, void *handle);

/* ---- in-kernel allreduce through peer memory (NVLink / NVSwitch; csrc/p2p.cuh) --------------------------------------
 * Collective over the ranks (all on this node; needs the shared-memory handle attached): every rank allocates a mailbox,
 * exports it with cudaIpcGetMemHandle and maps the peers'.  From then on the small device-side reductions (M-step
 * statistics, H-step partial sums and moments) are exchanged by remote stores from inside the kernels that produce them
 * instead of an NCCL launch in between.  *enabled = 0 (and nothing changes) when some rank cannot map some peer. */
VLGP_API int vlgp_comm_enable_p2p(vlgp_ctx *ctx, int *enabled);

/* ---- GPFA branch (vlgp/gpfa.py:20-56; reached through gpfa.fit and api.fastfit) --------------------------------------
 * vlgp_gpfa_estep: mu <- A B^-1 (y - d) of every equal-length segment of the set (vlgp/gpfa.py:37-45, before the mean
 * subtraction), evaluated as P h with h = bigC' bigR^-1 (y - d) and P = (I + bigK S)^-1 bigK formed by the caller:
 * C is L x N, d N, rho W x N (1 / noise applied to (bin, neuron)), PT the transpose of P, (L W) x (L W), vector index
 * l W + t.  vlgp_gpfa_stats: Z1'Z1 ((L+1)^2), Z1'Y ((L+1) x N) and sum y^2 (N) with Z1 = [mu, 1] -- the normal
 * equations of the lstsq M-step (vlgp/gpfa.py:49-53,83-88). */
VLGP_API int vlgp_gpfa_estep(vlgp_ctx *ctx, int set_id, const double *C, const double *d, const double *rho, const double *PT);
VLGP_API int vlgp_gpfa_stats(vlgp_ctx *ctx, int set_id, double *ZtZ, double *ZtY, double *yy);

/* ---- arithmetic of the E-step rate passes (BASELINE.json configs[2]: "fp32") ----------------------------------------
 * bits = 64 (default): everything in double precision, results match the reference (vlgp/core.py:68-113) to 1e-10.
 * bits = 32: the rate passes of the segment E-step (linear predictor, exp link, sums over neurons -- the bulk of the
 * arithmetic) run in single precision; Gram matrices, inverses, variances, the mean step, the M- and H-step and all state
 * stay double.  Applies to window-length segments with Poisson channels and uint8 counts; other inputs keep 64.
 * Posterior means then agree with the reference to about 1e-5 relative (tests/test_gpu_parity.py). */
VLGP_API int vlgp_set_precision(vlgp_ctx *ctx, int bits);

/* ---- measurement helpers (used by bench.py only) ------------------------------------------------------------------ */
/* Measured FP64 FMA peak (TFLOP/s) of this GPU with a register-resident DFMA loop, and with mma.sync.m8n8k4.f64. */
VLGP_API int vlgp_peak_fp64(vlgp_ctx *ctx, double *dfma_tflops, double *dmma_tflops);
/* Measured device-to-device copy bandwidth (GB/s, read + write bytes) on a buffer of nbytes. */
VLGP_API int vlgp_peak_hbm(vlgp_ctx *ctx, uint64_t nbytes, double *gbs);
/* Write nbytes (> L2) to evict the L2 between timed iterations. */
VLGP_API int vlgp_flush_l2(vlgp_ctx *ctx);
/* Device time (ms) of the kernel classes, accumulated with CUDA events on the context's stream:
 * class 0 = E-step kernel, 1 = M-step statistics kernel, 2 = H-step per-segment kernel, 3 = ichol kernel.
 * mask bit i enables class i (each timed launch adds one event synchronisation); enabling resets the accumulators. */
VLGP_API int vlgp_profile_enable(vlgp_ctx *ctx, int mask);
VLGP_API int vlgp_profile_get(vlgp_ctx *ctx, int which, double *total_ms, int64_t *n);

#ifdef __cplusplus
}
#endif
#endif /* VLGP_B200_H */
