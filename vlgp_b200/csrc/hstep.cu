// K7: H-step objective -- ELBO of a latent's GP hyperparameters and its derivative w.r.t. log(omega), batched over
// the evaluations the host optimiser asks for at the same time (one per latent when the L-BFGS-B runs are in lockstep).
//
// Replaces gp.construct_posterior_cov (vlgp/gp.py:126-147), gp.elbo (:12-43) and gp.kernel (:46-62) for the closure
// that scipy's L-BFGS-B evaluates (vlgp/gp.py:107-111).  The reference materialises S_i = (K^-1 + diag(w_i))^-1 for
// every segment and contracts W x W x S tensors.  With d = sqrt(w_i) and B_i = I + diag(d) K diag(d) (SPD, eigenvalues
// >= 1, so no inverse of the ill-conditioned K per segment):
//     tr(K^-1 S_i)              = tr(B_i^-1)
//     K^-1 S_i K^-1 - K^-1      = -diag(d) B_i^-1 diag(d)
// so    ll  = -1/2 tr(K^-1 M) - 1/2 sum_i tr(B_i^-1) - S sum(log diag chol K),        M = sum_i mu_i mu_i'
//       dll =  1/2 [ (K^-1 M K^-1) : dK  -  sum_i (diag(d) B_i^-1 diag(d)) : dK ],    dK = -omega D^2 o (K - eps I)
// (checked against the reference to 1e-12 with a NumPy prototype and by tests/test_gpu_parity.py on the golden
// vectors).  M is built once per H-step; each evaluation costs one W x W inverse per segment: hstep_dmma.cu does it on
// the FP64 tensor path with one warp per segment, the CTA-wide register sweep below is the fallback for W > 56.
#include "common.cuh"
#include "linalg.cuh"
#include "dmma.cuh"
#include "p2p.cuh"

int vlgp_launch_hstep_segments_dmma(vlgp_ctx *ctx, TrialSet *ts, const HEvalBatch &eb, bool *handled);   // hstep_dmma.cu
int vlgp_launch_hstep_moments_wide(vlgp_ctx *ctx, TrialSet *ts, int chunks, double *part);                // hstep_wide.cu
int vlgp_launch_hstep_wide(vlgp_ctx *ctx, TrialSet *ts, const HEvalBatch &eb);                            // hstep_wide.cu

namespace {

constexpr int NT = 256;

// ---- second moments of mu over segments: Mpart[chunk][l][a][b] -------------------------------------------------------
__global__ void __launch_bounds__(NT) hstep_moment_kernel(int nseg, int W, int L, const double *__restrict__ mu,
                                                          double *__restrict__ part) {
    __shared__ double tile[8][VLGP_MAX_W];
    const int l = blockIdx.y;
    const int per = (nseg + gridDim.x - 1) / gridDim.x;
    const int s0 = blockIdx.x * per, s1 = min(nseg, s0 + per);
    const int tid = threadIdx.x;
    constexpr int MAXE = (VLGP_MAX_W * VLGP_MAX_W + NT - 1) / NT;   // 16 entries per thread at W = 64
    double acc[MAXE];
#pragma unroll
    for (int e = 0; e < MAXE; ++e) acc[e] = 0.0;
    const int WW = W * W;
    for (int s = s0; s < s1; s += 8) {
        const int ns = min(8, s1 - s);
        __syncthreads();
        for (int i = tid; i < ns * W; i += NT) {
            const int q = i / W, t = i - q * W;
            tile[q][t] = mu[((size_t)(s + q) * W + t) * L + l];
        }
        __syncthreads();
#pragma unroll
        for (int e = 0; e < MAXE; ++e) {
            const int idx = tid + e * NT;
            if (idx < WW) {
                const int a = idx / W, b = idx - a * W;
                double x = acc[e];
                for (int q = 0; q < ns; ++q) x = fma(tile[q][a], tile[q][b], x);
                acc[e] = x;
            }
        }
    }
#pragma unroll
    for (int e = 0; e < MAXE; ++e) {
        const int idx = tid + e * NT;
        if (idx < WW) part[((size_t)blockIdx.x * L + l) * WW + idx] = acc[e];
    }
}

__global__ void reduce_parts_kernel2(const double *__restrict__ part, int G, int K, double *__restrict__ out) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= K) return;
    double s = 0.0;
    for (int g = 0; g < G; ++g) s += part[(size_t)g * K + k];
    out[k] = s;
}

// ---- per evaluation, one CTA (blockIdx.x = evaluation): K, dK, K^-1 terms -------------------------------------------
// out[e][0] = tr(K^-1 M), [1] = sum log diag chol(K), [2] = (K^-1 M K^-1):dK, [5] = 1 if K is not PD.
// This launch is on the critical path of an optimiser round when a rank holds few segments (the per-segment kernel next
// to it is then shorter), so it is built for latency: K, dK and M staged in shared memory (padded to whole 8 x 8 tiles),
// the CTA-wide register sweep with the pivots' logarithms taken in parallel afterwards, and the two W^3 products on the
// FP64 tensor path: with T = K^-1 M and U = dK K^-1 (all three matrices symmetric), tr(K^-1 M K^-1 dK) = sum_ij T_ij U_ij,
// so every warp forms matching tiles of T and U by DMMA and multiplies them element by element in registers.
__global__ void __launch_bounds__(NT) hstep_global_kernel(HEvalBatch eb, int W, double dt, const double *__restrict__ Mall,
                                                          double *__restrict__ Kall, double *__restrict__ outall,
                                                          int write_k) {
    extern __shared__ double sm[];
    const int e = blockIdx.x;
    const double sigmasq = eb.sigmasq[e], omega = eb.omega[e], eps = eb.eps[e];
    const double *M = Mall + (size_t)eb.latent[e] * W * W;
    double *Kout = Kall + (size_t)e * 2 * W * W, *dKout = Kout + (size_t)W * W;
    double *out = outall + e * 8;
    const int NB = (W + 7) >> 3, WP = 8 * NB, ld = WP + 1;
    double *Aw = sm;                   // WP x ld : K -> -K^-1
    double *Mw = Aw + WP * ld;         // WP x ld : M
    double *Dw = Mw + WP * ld;         // WP x ld : dK / dlog(omega)
    double *ck = Dw + WP * ld;         // 2 x 64
    double *piv = ck + 128;            // 64
    double *red = piv + 64;            // 32
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    for (int i = ty; i < WP; i += 16)
        for (int j = tx; j < WP; j += 16) {
            double k = 0.0, dk = 0.0, m = 0.0;
            if (i < W && j < W) {
                const double dx = (double)(i - j) * dt;
                const double d2 = dx * dx;
                const double ks = sigmasq * exp(-omega * d2);
                k = ks + (i == j ? eps : 0.0);
                dk = -ks * d2 * omega;
                m = M[i * W + j];
                if (write_k) {
                    Kout[i * W + j] = k;
                    dKout[i * W + j] = dk;
                }
            }
            Aw[i * ld + j] = k;
            Dw[i * ld + j] = dk;
            Mw[i * ld + j] = m;
        }
    __syncthreads();
    const bool ok = block_sweep_spd(Aw, ld, W, ck, nullptr, piv);
    if (!ok) {
        if (tid == 0) {
            out[0] = out[1] = out[2] = 0.0;
            out[5] = 1.0;
        }
        return;
    }
    const double logdet = block_sum(tid < W ? log(piv[tid]) : 0.0, red);
    // tiles of T' = (-K^-1) M and U' = dK (-K^-1); T' o U' = T o U, tr(T) = -tr(T')
    const int lane = tid & 31, wid = tid >> 5, r = lane >> 2, q = lane & 3;
    double t1 = 0.0, gr = 0.0;
    for (int tile = wid; tile < NB * NB; tile += NT / 32) {
        const int ti = tile / NB, tj = tile - ti * NB;
        Tile T{0.0, 0.0}, U{0.0, 0.0};
        const double *arow = Aw + (8 * ti + r) * ld + q, *drow = Dw + (8 * ti + r) * ld + q;
        const double *mcol = Mw + q * ld + 8 * tj + r, *acol = Aw + q * ld + 8 * tj + r;
        for (int k = 0; 4 * k < WP; ++k) {
            dmma(T, arow[4 * k], mcol[4 * k * ld]);
            dmma(U, drow[4 * k], acol[4 * k * ld]);
        }
        gr = fma(T.x, U.x, gr);
        gr = fma(T.y, U.y, gr);
        if (ti == tj) {
            if (r == 2 * q) t1 -= T.x;
            if (r == 2 * q + 1) t1 -= T.y;
        }
    }
    t1 = block_sum(t1, red);
    gr = block_sum(gr, red);
    if (tid == 0) {
        out[0] = t1;
        out[1] = 0.5 * logdet;
        out[2] = gr;
        out[5] = 0.0;
    }
}

// ---- fallback (W > 56): one CTA per segment, B_i^-1 by the CTA-wide register sweep; blockIdx.y = evaluation --------
__global__ void __launch_bounds__(NT, 2) hstep_segment_kernel(HEvalBatch eb, int nseg, int W, int L,
                                                              const double *__restrict__ w,
                                                              const double *__restrict__ Kall,
                                                              double *__restrict__ partall) {
    __shared__ double dv[64];          // sqrt(w)
    __shared__ double ck[128];
    __shared__ double red[32];
    const int e = blockIdx.y, l = eb.latent[e];
    const double *K = Kall + (size_t)e * 2 * W * W, *dK = K + (size_t)W * W;
    double *part = partall + (size_t)e * 2 * nseg;
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    double kreg[4][4];
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int i = ty + 16 * p, j = tx + 16 * q;
            kreg[p][q] = (i < W && j < W) ? K[i * W + j] : 0.0;
        }
    for (int seg = blockIdx.x; seg < nseg; seg += gridDim.x) {
        __syncthreads();
        if (tid < 64) dv[tid] = tid < W ? sqrt(fmax(w[((size_t)seg * W + tid) * L + l], 0.0)) : 0.0;
        __syncthreads();
        double di[4], dj[4], a[4][4];
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            di[p] = dv[ty + 16 * p];
            dj[p] = dv[tx + 16 * p];
        }
#pragma unroll
        for (int p = 0; p < 4; ++p)
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int i = ty + 16 * p, j = tx + 16 * q;
                a[p][q] = di[p] * kreg[p][q] * dj[q] + ((i == j && i < W) ? 1.0 : 0.0);
            }
        const bool ok = block_sweep_regs(a, W, ck, nullptr);
        double tr = 0.0, pd = 0.0;
        if (ok) {
#pragma unroll
            for (int p = 0; p < 4; ++p)
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int i = ty + 16 * p, j = tx + 16 * q;
                    const double binv = -a[p][q];
                    if (i == j) tr += binv;
                    if (i < W && j < W) pd = fma(binv * di[p] * dj[q], __ldg(dK + i * W + j), pd);
                }
        }
        tr = block_sum(tr, red);
        pd = block_sum(pd, red);
        if (tid == 0) {
            const double nan = __longlong_as_double(0x7ff8000000000000LL);
            part[seg] = ok ? tr : nan;
            part[nseg + seg] = ok ? pd : nan;
        }
    }
}

// red[e][0] = sum_i tr(B_i^-1), red[e][1] = sum_i (d B_i^-1 d):dK   (deterministic, one CTA per evaluation); with peer
// memory the two sums are exchanged with the other ranks here (p2p.cuh, chunk = evaluation) instead of on the host
__global__ void __launch_bounds__(NT) hstep_final_kernel(int nseg, const double *__restrict__ partall,
                                                         double *__restrict__ redall, P2PDev pd) {
    __shared__ double red[32];
    __shared__ double pair[2];
    const double *part = partall + (size_t)blockIdx.x * 2 * nseg;
    double a = 0.0, b = 0.0;
    for (int i = threadIdx.x; i < nseg; i += NT) {
        a += part[i];
        b += part[nseg + i];
    }
    a = block_sum(a, red);
    b = block_sum(b, red);
    if (threadIdx.x == 0) {
        pair[0] = a;
        pair[1] = b;
    }
    __syncthreads();
    p2p_allreduce_cta(pd, blockIdx.x, (size_t)blockIdx.x * 2, pair, 2);
    if (threadIdx.x == 0) {
        redall[blockIdx.x * 2] = pair[0];
        redall[blockIdx.x * 2 + 1] = pair[1];
    }
}

}   // namespace

static size_t hstep_global_smem(int W) {
    const int WP = 8 * ((W + 7) / 8);
    return ((size_t)3 * WP * (WP + 1) + 128 + 64 + 32) * sizeof(double);
}

int vlgp_launch_hstep_prepare(vlgp_ctx *ctx, TrialSet *ts) {
    const int W = ts->max_len, L = ctx->L, S = ts->n_trials;
    const int WW = W * W;
    int chunks = (S + 31) / 32;
    const int maxc = 4 * ctx->prop.multiProcessorCount / (L > 0 ? L : 1) + 1;
    if (chunks > maxc) chunks = maxc;
    if (chunks < 1) chunks = 1;
    if (!ts->d_mompart) CK(vlgp_dalloc(ctx, &ts->d_mompart, (size_t)maxc * L * WW * sizeof(double)));
    double *part = ts->d_mompart;
    if (!ts->d_M) CK(vlgp_dalloc(ctx, &ts->d_M, (size_t)L * WW * sizeof(double)));
    if (!ts->d_K) CK(vlgp_dalloc(ctx, &ts->d_K, (size_t)VLGP_MAX_L * 2 * WW * sizeof(double)));
    if (!ts->d_hpart) CK(vlgp_dalloc(ctx, &ts->d_hpart, (size_t)VLGP_MAX_L * 2 * S * sizeof(double)));
    if (!ts->d_hout) CK(vlgp_dalloc(ctx, &ts->d_hout, (size_t)VLGP_MAX_L * 10 * sizeof(double)));
    if (W > VLGP_MAX_W) {              // wide window: shared-memory kernels of hstep_wide.cu
        int rcw = vlgp_launch_hstep_moments_wide(ctx, ts, chunks, part);
        if (rcw) return rcw;
    } else {
        hstep_moment_kernel<<<dim3(chunks, L), NT, 0, ctx->stream>>>(S, W, L, ts->d_mu, part);
        CKL();
    }
    reduce_parts_kernel2<<<(L * WW + 127) / 128, 128, 0, ctx->stream>>>(part, chunks, L * WW, ts->d_M);
    CKL();
    int rc = vlgp_allreduce_dev(ctx, ts->d_M, (size_t)L * WW, 0);
    if (rc) return rc;
    // number of segments over all ranks (the S of the log-determinant term)
    ctx->h_pin[0] = (double)S;
    if (ctx->n_ranks > 1 && !ctx->shm) {
        CK(cudaMemcpyAsync(ctx->d_small, ctx->h_pin, sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        rc = vlgp_allreduce_dev(ctx, ctx->d_small, 1, 0);
        if (rc) return rc;
        CK(cudaMemcpyAsync(ctx->h_pin, ctx->d_small, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    }
    CK(cudaStreamSynchronize(ctx->stream));
    double nseg_all = ctx->h_pin[0];
    if (ctx->n_ranks > 1 && ctx->shm) {
        rc = vlgp_comm_allreduce(ctx, &nseg_all, 1, 0);
        if (rc) return rc;
    }
    ts->h_nseg_total = nseg_all;
    if (!ts->h_geometry && W <= VLGP_MAX_W) {
        // launch geometry of the per-segment fallback kernel (the DMMA kernel sizes its own grid), once per set
        int per_sm = 1;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, hstep_segment_kernel, NT, 0));
        ts->h_seg_grid = (per_sm < 1 ? 1 : per_sm) * ctx->prop.multiProcessorCount;
        if (ts->h_seg_grid > S) ts->h_seg_grid = S;
        const size_t smem_g = hstep_global_smem(W);
        if (smem_g > 48 * 1024)
            CK(cudaFuncSetAttribute(hstep_global_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_g));
        ts->h_geometry = true;
    }
    ts->h_prepared = true;
    return VLGP_OK;
}

// Evaluates eb.n (latent, hyper) pairs in one pass: 3 launches, one allreduce, one D2H copy, one synchronisation.
int vlgp_launch_hstep_objective(vlgp_ctx *ctx, TrialSet *ts, const HEvalBatch &eb, double *ll, double *dll, int *info) {
    const int W = ts->max_len, S = ts->n_trials, n = eb.n;
    const size_t smem_g = hstep_global_smem(W);
    double *red = ts->d_hout + VLGP_MAX_L * 8;
    // The K^-1 terms (one CTA per evaluation, latency-bound) run on a second stream concurrently with the per-segment
    // kernel; the DMMA kernel builds K itself, only the W > 56 fallback reads the K written by the global kernel.
    const bool dmma_ok = ts->max_len <= 56 && !getenv("VLGP_FORCE_SWEEP_HSTEP");
    const bool wide = W > VLGP_MAX_W;
    if (wide) {
        int rcw = vlgp_launch_hstep_wide(ctx, ts, eb);
        if (rcw) return rcw;
    } else if (dmma_ok) {
        CK(cudaEventRecord(ctx->ev_fork, ctx->stream));
        CK(cudaStreamWaitEvent(ctx->stream2, ctx->ev_fork, 0));
        hstep_global_kernel<<<n, NT, smem_g, ctx->stream2>>>(eb, W, ctx->dt, ts->d_M, ts->d_K, ts->d_hout, 0);
        CKL();
        CK(cudaEventRecord(ctx->ev_join, ctx->stream2));
    } else {
        hstep_global_kernel<<<n, NT, smem_g, ctx->stream>>>(eb, W, ctx->dt, ts->d_M, ts->d_K, ts->d_hout, 1);
        CKL();
    }
    if (!wide) {
        ProfScope ps(ctx, 2);
        bool handled = false;
        int rcd = vlgp_launch_hstep_segments_dmma(ctx, ts, eb, &handled);
        if (rcd) return rcd;
        if (!handled) {
            hstep_segment_kernel<<<dim3(ts->h_seg_grid, n), NT, 0, ctx->stream>>>(eb, S, W, ctx->L, ts->d_w, ts->d_K,
                                                                                 ts->d_hpart);
            CKL();
        }
    }
    if (dmma_ok && !wide) CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0));
    // (tr, pd) sums over ranks: inside the final kernel through peer memory when enabled; else on the host through
    // shared memory when attached (they are consumed there), else NCCL
    const bool p2p = ctx->n_ranks > 1 && vlgp_p2p_enabled(ctx);
    P2PDev pd{};
    pd.n_ranks = 1;
    if (p2p) pd = vlgp_p2p_next(ctx);
    hstep_final_kernel<<<n, NT, 0, ctx->stream>>>(S, ts->d_hpart, red, pd);
    CKL();
    int rc = (ctx->shm || p2p) ? VLGP_OK : vlgp_allreduce_dev(ctx, red, 2 * n, 0);
    if (rc) return rc;
    CK(cudaMemcpyAsync(ctx->h_pin, ts->d_hout, (size_t)VLGP_MAX_L * 10 * sizeof(double), cudaMemcpyDeviceToHost,
                       ctx->stream));
    static const int pump_env = [] {      // Newton iterations of an overlapped M-step enqueued per H-step round
        const char *e = getenv("VLGP_MSTEP_PUMP");
        const int v = e ? atoi(e) : 0;
        return (v >= 1 && v <= 64) ? v : 0;
    }();
    // 2 per round keeps the M-step just ahead of the H-step on one GPU; on a shard of an 8-GPU run the H-step has fewer,
    // shorter rounds than the M-step has iterations left at its end, so 3 (measured on 8 x B200: 3.43 vs 3.54 ms per step)
    const int pump = pump_env ? pump_env : (ctx->n_ranks > 1 ? 3 : 2);
    rc = vlgp_mstep_pump(ctx, pump);   // an overlapped M-step gets its next launches while this round runs
    if (rc) return rc;
    CK(cudaStreamSynchronize(ctx->stream));
    if (ctx->shm && ctx->n_ranks > 1 && !p2p) {
        rc = vlgp_comm_allreduce(ctx, ctx->h_pin + VLGP_MAX_L * 8, 2 * n, 0);
        if (rc) return rc;
    }
    const double *o = ctx->h_pin, *r = ctx->h_pin + VLGP_MAX_L * 8;
    for (int e = 0; e < n; ++e) {
        const double *oe = o + e * 8;
        ll[e] = -0.5 * oe[0] - 0.5 * r[2 * e] - ts->h_nseg_total * oe[1];
        dll[e] = 0.5 * (oe[2] - r[2 * e + 1]);
        info[e] = oe[5] != 0.0 ? 1 : ((r[2 * e] != r[2 * e]) ? 2 : 0);      // 1: K not PD; 2: some B_i not PD (NaN)
    }
    ctx->counters[2] += (int64_t)S * n;
    return VLGP_OK;
}
