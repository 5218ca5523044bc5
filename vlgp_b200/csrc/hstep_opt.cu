// H-step optimiser without the Python interpreter in the round loop (vlgp/gp.py:65-123).
//
// Part 1 (this section, context-free, runs on the CPU): C entry points around lbfgsb.cuh so that tests can drive the
// restated L-BFGS-B next to scipy's setulb on identical objectives (tests/test_lbfgsb_port.py).
#include <new>

#include "common.cuh"
#include "lbfgsb.cuh"
#include "p2p.cuh"

extern "C" {

VLGP_API int vlgp_lbfgsb_new(int n, const double *x0, const double *lower, const double *upper, double factr, double pgtol,
                             int maxls, double collapse_tol, void **handle) {
    if (!handle || !x0 || !lower || !upper || n < 1 || n > lbfgsb::NMAX) return VLGP_ERR_ARG;
    lbfgsb::State *s = new (std::nothrow) lbfgsb::State();
    if (!s) return VLGP_ERR_NOMEM;
    lbfgsb::init(*s, n, x0, lower, upper, factr, pgtol, maxls, 15000, 15000, collapse_tol);
    *handle = s;
    return VLGP_OK;
}

VLGP_API int vlgp_lbfgsb_advance(void *handle, double f, const double *g, double *x, int *task) {
    if (!handle) return VLGP_ERR_ARG;
    lbfgsb::State &s = *(lbfgsb::State *)handle;
    if (s.phase != lbfgsb::PH_START && g) {
        s.f = f;
        for (int i = 0; i < s.n; ++i) s.g[i] = g[i];
    }
    const int need = lbfgsb::advance(s);
    if (x)
        for (int i = 0; i < s.n; ++i) x[i] = s.x[i];
    if (task) *task = s.task;
    return need;
}

VLGP_API int vlgp_lbfgsb_info(void *handle, double *f, int *nfev, int *nit, int *n_collapsed) {
    if (!handle) return VLGP_ERR_ARG;
    const lbfgsb::State &s = *(const lbfgsb::State *)handle;
    if (f) *f = s.f;
    if (nfev) *nfev = s.nfgv;
    if (nit) *nit = s.nit;
    if (n_collapsed) *n_collapsed = s.n_collapsed;
    return VLGP_OK;
}

VLGP_API int vlgp_lbfgsb_free(void *handle) {
    delete (lbfgsb::State *)handle;
    return VLGP_OK;
}

}   // extern "C"

// ---------------------------------------------------------------------------------------------------------------------
// Part 2: the whole H-step optimisation of all latents in ONE call (replaces the per-latent loop of gp.optimize,
// vlgp/gp.py:82-92, and the closure scipy's L-BFGS-B evaluates, :100-123).  The per-latent optimisers advance in
// lockstep: every round evaluates the points all still-active optimisers ask for in one batched device pass (the
// existing objective launcher: K^-1 kernel beside the per-segment DMMA kernel, final reduction, one allreduce over
// ranks).  Each optimiser sees exactly the values it would see alone, so the iterates are those of the reference's
// sequential runs; no Python, no scipy, no ctypes marshalling between rounds (about 40 us per round before).
//
// Evaluations are memoised per latent on the bit pattern of x: the line search of L-BFGS-B re-evaluates the point it
// finally accepts (its "XTOL" exit sets stp = stx, dcsrch) and collapsing searches return to their starting point;
// the objective is deterministic, so a repeated x gets the values it got before without a device round.
// ---------------------------------------------------------------------------------------------------------------------
int vlgp_launch_hstep_prepare(vlgp_ctx *ctx, TrialSet *ts);
int vlgp_launch_hstep_objective(vlgp_ctx *ctx, TrialSet *ts, const HEvalBatch &eb, double *ll, double *dll, int *info);

namespace {
struct HCacheEntry {
    double x[3];
    double ll, dll;
};
}   // namespace

extern "C" {

VLGP_API int vlgp_hstep_optimize(vlgp_ctx *ctx, int set_id, int n_lat, const int32_t *latents, const double *log_initial,
                                 const double *log_bounds, const int32_t *mask, double collapse_tol, double *log_result,
                                 double *fval, int32_t *nfev, int32_t *task, int32_t *n_rounds) {
    TrialSet *ts = get_set(ctx, set_id);
    REQUIRE(ts, "hstep_optimize: bad set %d", set_id);
    REQUIRE(n_lat >= 1 && n_lat <= VLGP_MAX_L && latents && log_initial && log_bounds && mask && log_result,
            "hstep_optimize: bad arguments");
    REQUIRE(ts->min_len == ts->max_len, "hstep_optimize: all segments must have the same length (vlgp/gp.py:77-80)");
    REQUIRE(ts->max_len <= VLGP_MAX_W_H, "hstep_optimize: window %d > %d", ts->max_len, VLGP_MAX_W_H);
    for (int k = 0; k < n_lat; ++k)
        REQUIRE(latents[k] >= 0 && latents[k] < ctx->L, "hstep_optimize: latent %d out of range", latents[k]);
    CK(cudaSetDevice(ctx->device));
    int rc = vlgp_launch_hstep_prepare(ctx, ts);
    if (rc) return rc;

    std::vector<lbfgsb::State> st(n_lat);
    std::vector<std::vector<HCacheEntry>> cache(n_lat);
    std::vector<int> done(n_lat, 0), asked(n_lat, 0);
    double lo[3], up[3];
    for (int i = 0; i < 3; ++i) {
        lo[i] = log_bounds[2 * i];
        up[i] = log_bounds[2 * i + 1];
    }
    for (int k = 0; k < n_lat; ++k)
        lbfgsb::init(st[k], 3, log_initial + 3 * k, lo, up, 1e7, 1e-5, 20, 15000, 15000, collapse_tol);
    int rounds = 0;
    for (;;) {
        int pend[VLGP_MAX_L], np = 0;
        for (int k = 0; k < n_lat; ++k) {
            if (done[k]) continue;
            for (;;) {
                if (!lbfgsb::advance(st[k])) {
                    done[k] = 1;
                    break;
                }
                asked[k]++;
                const HCacheEntry *hit = nullptr;
                for (const HCacheEntry &c : cache[k])
                    if (memcmp(c.x, st[k].x, sizeof(c.x)) == 0) {
                        hit = &c;
                        break;
                    }
                if (!hit) {
                    pend[np++] = k;
                    break;
                }
                st[k].f = -hit->ll;
                for (int i = 0; i < 3; ++i) st[k].g[i] = -((i == 1 ? hit->dll : 0.0) * (double)mask[i]);
            }
        }
        if (np == 0) break;
        double hyper[VLGP_MAX_L][3], ll[VLGP_MAX_L], dll[VLGP_MAX_L];
        for (int q = 0; q < np; ++q)
            for (int i = 0; i < 3; ++i) hyper[q][i] = exp(st[pend[q]].x[i]);
        int todo[VLGP_MAX_L], nt = np;
        for (int q = 0; q < np; ++q) todo[q] = q;
        while (nt > 0) {
            HEvalBatch eb{};
            eb.n = nt;
            for (int j = 0; j < nt; ++j) {
                const int q = todo[j];
                eb.latent[j] = latents[pend[q]];
                eb.sigmasq[j] = hyper[q][0];
                eb.omega[j] = hyper[q][1];
                eb.eps[j] = hyper[q][2];
            }
            double l_[VLGP_MAX_L], d_[VLGP_MAX_L];
            int inf[VLGP_MAX_L];
            rc = vlgp_launch_hstep_objective(ctx, ts, eb, l_, d_, inf);
            if (rc) return rc;
            rounds++;
            int again = 0;
            for (int j = 0; j < nt; ++j) {
                const int q = todo[j];
                if (inf[j] == 1) {
                    hyper[q][1] += 2.302585092994046;      // the reference's retry when K is not PD (vlgp/gp.py:133-135)
                    todo[again++] = q;
                } else {
                    ll[q] = l_[j];
                    dll[q] = d_[j];
                }
            }
            nt = again;
        }
        for (int q = 0; q < np; ++q) {
            const int k = pend[q];
            HCacheEntry c;
            memcpy(c.x, st[k].x, sizeof(c.x));
            c.ll = ll[q];
            c.dll = dll[q];
            cache[k].push_back(c);
            st[k].f = -ll[q];
            for (int i = 0; i < 3; ++i) st[k].g[i] = -((i == 1 ? dll[q] : 0.0) * (double)mask[i]);
        }
    }
    for (int k = 0; k < n_lat; ++k) {
        for (int i = 0; i < 3; ++i) log_result[3 * k + i] = st[k].x[i];
        if (fval) fval[k] = st[k].f;
        if (nfev) nfev[k] = asked[k];
        if (task) task[k] = st[k].task;
    }
    if (n_rounds) *n_rounds = rounds;
    return vlgp_p2p_check(ctx);
}

}   // extern "C"
