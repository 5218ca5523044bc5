// K7 (per-segment part) on the FP64 tensor path: one WARP per segment, the W x W matrix B_i = I + d K d (d = sqrt(w_i))
// held entirely in registers as 8 x 8 tiles in the accumulator layout of mma.sync.m8n8k4.f64, inverted in place by a
// BLOCKED symmetric sweep whose rank-8 updates are DMMA instructions.
//
// Block sweep on pivot block k (block generalisation of the scalar sweep in linalg.cuh; after all blocks the tiles hold
// -B^-1):   P = A_kk^-1 ;  A_ik <- A_ik P ;  A_ij <- A_ij - A_ik P A_kj  (i, j != k) ;  A_kk <- -P
// Only the lower block triangle is stored (NB (NB+1) / 2 tiles, 2 doubles per lane each).  Operand fragments are
// produced from accumulator-layout tiles with intra-warp shuffles ("N-form": element [lane/4][4h + lane%4], "T-form":
// element [4h + lane%4][lane/4]); the 8 x 8 pivot block is inverted by an 8-step scalar sweep done with shuffles.
// No shared-memory traffic and no block barrier inside a segment; K and dK/dlog(omega) are built once per CTA into
// shared memory in tile order and reused by every segment the CTA processes.  The lane-level algorithm was validated against
// numpy.linalg.inv with a 32-lane emulation before it was written in CUDA (scripts/dmma_block_sweep_emulation.py).
#include "common.cuh"
#include "dmma.cuh"

namespace {

constexpr int WARPS = 4;


template <int NB, bool HALF_LAST>
__global__ void __launch_bounds__(WARPS * 32) hstep_segment_dmma_kernel(HEvalBatch eb, int nseg, int W, int L, double dt,
                                                                        const double *__restrict__ w,
                                                                        double *__restrict__ partall) {
    const int ev = blockIdx.y, l = eb.latent[ev];          // blockIdx.y = evaluation of the batch
    const double sigmasq = eb.sigmasq[ev], omega = eb.omega[ev], eps = eb.eps[ev];
    double *part = partall + (size_t)ev * 2 * nseg;
    constexpr int NT = NB * (NB + 1) / 2;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double2 *Ks = (double2 *)smem_raw;                 // NT x 32 : K in tile / lane order
    double2 *dKs = Ks + NT * 32;                       // NT x 32 : dK/dlog(omega)
    double *dsm = (double *)(dKs + NT * 32);           // WARPS x 64 : sqrt(w) of the warp's segment (0 beyond W)
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int r = lane >> 2, c0 = 2 * (lane & 3);

    for (int t = wid; t < NT; t += WARPS) {
        int i = 0;
        while (tix(i + 1, 0) <= t) ++i;
        const int j = t - tix(i, 0);
        const int gi = 8 * i + r, gj = 8 * j + c0;
        // K = sigma^2 exp(-omega D^2) + eps I and dK/dlog(omega) = -omega D^2 o (K - eps I), built here with the same
        // expressions as hstep_global_kernel (vlgp/gp.py:46-62) so that the two kernels can run concurrently
        double2 kv = make_double2(0.0, 0.0), dv = make_double2(0.0, 0.0);
        if (gi < W && gj < W) {
            const double dx = (double)(gi - gj) * dt, d2 = dx * dx;
            const double ks = sigmasq * exp(-omega * d2);
            kv.x = ks + (gi == gj ? eps : 0.0);
            dv.x = -ks * d2 * omega;
        }
        if (gi < W && gj + 1 < W) {
            const double dx = (double)(gi - gj - 1) * dt, d2 = dx * dx;
            const double ks = sigmasq * exp(-omega * d2);
            kv.y = ks + (gi == gj + 1 ? eps : 0.0);
            dv.y = -ks * d2 * omega;
        }
        Ks[t * 32 + lane] = kv;
        dKs[t * 32 + lane] = dv;
    }
    __syncthreads();

    double *dw = dsm + wid * 64;
    const int stride = gridDim.x * WARPS;
    for (int seg = blockIdx.x * WARPS + wid; seg < nseg; seg += stride) {
        __syncwarp();
        for (int t = lane; t < 64; t += 32)
            dw[t] = t < W ? sqrt(fmax(w[((size_t)seg * W + t) * L + l], 0.0)) : 0.0;
        __syncwarp();
        // ---- B = I + d K d in tiles (identity on the padding) --------------------------------------------------
        Tile A[NT];
#pragma unroll
        for (int i = 0; i < NB; ++i) {
            const double di = dw[8 * i + r];
#pragma unroll
            for (int j = 0; j <= i; ++j) {
                const double2 kv = Ks[tix(i, j) * 32 + lane];
                A[tix(i, j)].x = fma(di * kv.x, dw[8 * j + c0], (i == j && r == c0) ? 1.0 : 0.0);
                A[tix(i, j)].y = fma(di * kv.y, dw[8 * j + c0 + 1], (i == j && r == c0 + 1) ? 1.0 : 0.0);
            }
        }
        // ---- blocked symmetric sweep -> tiles hold -B^-1 ----------------------------------------------------------
        // When at most four rows of the LAST block are real (W = 50: rows 48, 49), columns 4..7 of that pivot block are
        // identity padding: the off-diagonal tiles of its block column are exactly zero there, so the second k4 half of
        // every product of the last sweep step adds exact zeros and is left out (27 of 378 DMMA and their operand
        // shuffles at NB = 7; results unchanged bit for bit).  HALF_LAST = (W - 8 (NB - 1) <= 4), chosen by the launcher.
        bool ok = true;
#pragma unroll
        for (int kb = 0; kb < NB; ++kb) {
            const bool hi = !(kb == NB - 1 && HALF_LAST);       // compile-time: kb is unrolled
            Tile P = A[tix(kb, kb)];
            ok = tile_spd_inverse(P, lane) && ok;
            const double Pt0 = tform(P, 0, lane), Pn0 = nform(P, 0, lane);
            double Pt1 = 0.0, Pn1 = 0.0;
            if (hi) {
                Pt1 = tform(P, 1, lane);
                Pn1 = nform(P, 1, lane);
            }
            double V0[NB], V1[NB];            // operand form of the OLD block column A_{m,kb} (rows = block m)
#pragma unroll
            for (int m = 0; m < NB; ++m) {
                if (m == kb) continue;
                V1[m] = 0.0;
                if (m > kb) {
                    V0[m] = nform(A[tix(m, kb)], 0, lane);
                    if (hi) V1[m] = nform(A[tix(m, kb)], 1, lane);
                } else {
                    V0[m] = tform(A[tix(kb, m)], 0, lane);
                    if (hi) V1[m] = tform(A[tix(kb, m)], 1, lane);
                }
            }
#pragma unroll
            for (int m = 0; m < NB; ++m) {    // new block column: A_{m,kb} P  (stored transposed for m < kb)
                if (m == kb) continue;
                Tile T{0.0, 0.0};
                if (m > kb) {
                    dmma(T, V0[m], Pt0);
                    if (hi) dmma(T, V1[m], Pt1);
                    A[tix(m, kb)] = T;
                } else {
                    dmma(T, Pn0, V0[m]);
                    if (hi) dmma(T, Pn1, V1[m]);
                    A[tix(kb, m)] = T;
                }
            }
#pragma unroll
            for (int i = 0; i < NB; ++i) {    // A_ij -= (A_ik P) A_kj
                if (i == kb) continue;
                double T0, T1 = 0.0;
                if (i > kb) {
                    T0 = -nform(A[tix(i, kb)], 0, lane);
                    if (hi) T1 = -nform(A[tix(i, kb)], 1, lane);
                } else {
                    T0 = -tform(A[tix(kb, i)], 0, lane);
                    if (hi) T1 = -tform(A[tix(kb, i)], 1, lane);
                }
#pragma unroll
                for (int j = 0; j <= i; ++j) {
                    if (j == kb) continue;
                    dmma(A[tix(i, j)], T0, V0[j]);
                    if (hi) dmma(A[tix(i, j)], T1, V1[j]);
                }
            }
            A[tix(kb, kb)].x = -P.x;
            A[tix(kb, kb)].y = -P.y;
        }
        // ---- tr(B^-1) and (d B^-1 d) : dK ---------------------------------------------------------------------------
        double tr = 0.0, pd = 0.0;
#pragma unroll
        for (int i = 0; i < NB; ++i) {
            const double di = dw[8 * i + r];            // sqrt(w) is re-read from SMEM: not kept live through the sweep
#pragma unroll
            for (int j = 0; j <= i; ++j) {
                const double2 dv = dKs[tix(i, j) * 32 + lane];
                const double bx = -A[tix(i, j)].x, by = -A[tix(i, j)].y;
                const double wgt = (i == j) ? di : 2.0 * di;
                pd = fma(wgt * bx * dw[8 * j + c0], dv.x, pd);
                pd = fma(wgt * by * dw[8 * j + c0 + 1], dv.y, pd);
                if (i == j) {
                    const int gi = 8 * i + r;
                    if (r == c0 && gi < W) tr += bx;
                    if (r == c0 + 1 && gi < W) tr += by;
                }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            tr += __shfl_xor_sync(FULL, tr, o);
            pd += __shfl_xor_sync(FULL, pd, o);
        }
        if (lane == 0) {
            if (!ok) tr = pd = __longlong_as_double(0x7ff8000000000000LL);     // NaN marks a non-PD B_i for the host
            part[seg] = tr;
            part[nseg + seg] = pd;
        }
    }
}

template <int NB, bool HALF_LAST>
int launch_th(vlgp_ctx *ctx, TrialSet *ts, const HEvalBatch &eb) {
    constexpr int NT = NB * (NB + 1) / 2;
    const size_t smem = (size_t)2 * NT * 32 * sizeof(double2) + WARPS * 64 * sizeof(double);
    const int S = ts->n_trials;
    if (ts->dmma_grid == 0) {
        if (smem > 48 * 1024)
            CK(cudaFuncSetAttribute(hstep_segment_dmma_kernel<NB, HALF_LAST>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)smem));
        int per_sm = 1;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, hstep_segment_dmma_kernel<NB, HALF_LAST>, WARPS * 32, smem));
        if (per_sm < 1) per_sm = 1;
        ts->dmma_grid = per_sm * ctx->prop.multiProcessorCount;
        if (ts->dmma_grid > (S + WARPS - 1) / WARPS) ts->dmma_grid = (S + WARPS - 1) / WARPS;
    }
    const int grid = ts->dmma_grid;
    hstep_segment_dmma_kernel<NB, HALF_LAST><<<dim3(grid, eb.n), WARPS * 32, smem, ctx->stream>>>(eb, S, ts->max_len, ctx->L, ctx->dt,
                                                                                      ts->d_w, ts->d_hpart);
    CKL();
    return VLGP_OK;
}

template <int NB>
int launch_t(vlgp_ctx *ctx, TrialSet *ts, const HEvalBatch &eb) {
    // at most four real rows in the last 8-row block: the second half of the last sweep step is all padding
    return ts->max_len - 8 * (NB - 1) <= 4 ? launch_th<NB, true>(ctx, ts, eb) : launch_th<NB, false>(ctx, ts, eb);
}

}   // namespace

// Returns VLGP_OK and sets *handled when the window fits the register-resident tile layout (W <= 56).
int vlgp_launch_hstep_segments_dmma(vlgp_ctx *ctx, TrialSet *ts, const HEvalBatch &eb, bool *handled) {
    *handled = false;
    if (getenv("VLGP_FORCE_SWEEP_HSTEP")) return VLGP_OK;
    const int NB = (ts->max_len + 7) / 8;
    int rc = VLGP_OK;
    switch (NB) {
        case 1: rc = launch_t<1>(ctx, ts, eb); break;
        case 2: rc = launch_t<2>(ctx, ts, eb); break;
        case 3: rc = launch_t<3>(ctx, ts, eb); break;
        case 4: rc = launch_t<4>(ctx, ts, eb); break;
        case 5: rc = launch_t<5>(ctx, ts, eb); break;
        case 6: rc = launch_t<6>(ctx, ts, eb); break;
        case 7: rc = launch_t<7>(ctx, ts, eb); break;
        default: return VLGP_OK;          // W > 56: the CTA-wide sweep kernel in hstep.cu handles it
    }
    if (rc == VLGP_OK) *handled = true;
    return rc;
}
