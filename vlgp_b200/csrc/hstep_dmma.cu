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

// Bordered variant for windows with one or two bins beyond a multiple of 8 (W = 50 = 6 x 8 + 2, the reference's default):
// the q = W - 8 NBC border rows are eliminated FIRST, in scalar code, so that the blocked sweep runs on NBC = NB - 1 block
// rows (240 instead of 351 DMMA at W = 50) and none of its tiles is padding.  With B = [[B11, B12], [B21, B22]]
// (B22: 2 x 2, identity-padded when q = 1), X = B12 B22^-1 and the Schur complement S = B11 - X B21:
//     B^-1 = [[S^-1, -S^-1 X], [-X' S^-1, B22^-1 + X' S^-1 X]],
// and both outputs are ELEMENT-WISE sums over the tiles of S^-1 (no product with the border is formed):
//     tr(B^-1)      = <S^-1, I + X X'> + tr(B22^-1)
//     tr(B^-1 C)    = <S^-1, C11 - X C21 - C12 X' + X C22 X'> + tr(B22^-1 C22),      C = d dK d.
// Per segment the warp builds seven 8 NBC-vectors in shared memory (u, v = the border columns of B; x1, x2 = columns of
// X; and G, H = rows of C21 minus half of C22 X') and the tile loops read them by row / column index.
template <int NBC>
__global__ void __launch_bounds__(WARPS * 32) hstep_segment_schur_kernel(HEvalBatch eb, int nseg, int W, int L, double dt,
                                                                         const double *__restrict__ w,
                                                                         double *__restrict__ partall) {
    const int ev = blockIdx.y, l = eb.latent[ev];
    const double sigmasq = eb.sigmasq[ev], omega = eb.omega[ev], eps = eb.eps[ev];
    double *part = partall + (size_t)ev * 2 * nseg;
    constexpr int NT = NBC * (NBC + 1) / 2, WC = 8 * NBC;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double2 *Ks = (double2 *)smem_raw;                 // NT x 32 : K (core block) in tile / lane order
    double2 *dKs = Ks + NT * 32;                       // NT x 32 : dK/dlog(omega)
    double *bord = (double *)(dKs + NT * 32);          // 4 x WC : K(i, a), K(i, b), dK(a, i), dK(b, i)   (a = WC, b = WC + 1)
    double *bsc = bord + 4 * WC;                       // 8 : K(a,a), K(a,b), K(b,b), dK(a,b)
    double *vecs = bsc + 8;                            // WARPS x (7 WC + 8): d | u | v | x1 | x2 | G | H per warp
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int r = lane >> 2, c0 = 2 * (lane & 3);
    const int qb = W - WC;                             // 1 or 2 border rows

    auto kern = [&](int i, int j, double &k, double &dk) {
        const double dx = (double)(i - j) * dt, d2 = dx * dx;
        const double ks = sigmasq * exp(-omega * d2);
        k = ks + (i == j ? eps : 0.0);
        dk = -ks * d2 * omega;
    };
    for (int t = wid; t < NT; t += WARPS) {
        int i = 0;
        while (tix(i + 1, 0) <= t) ++i;
        const int j = t - tix(i, 0);
        const int gi = 8 * i + r, gj = 8 * j + c0;
        double2 kv, dv;
        kern(gi, gj, kv.x, dv.x);
        kern(gi, gj + 1, kv.y, dv.y);
        Ks[t * 32 + lane] = kv;
        dKs[t * 32 + lane] = dv;
    }
    for (int i = tid; i < WC; i += WARPS * 32) {
        double k, dk;
        kern(i, WC, k, dk);
        bord[i] = k;
        bord[2 * WC + i] = dk;
        k = dk = 0.0;
        if (qb > 1) kern(i, WC + 1, k, dk);
        bord[WC + i] = k;
        bord[3 * WC + i] = dk;
    }
    if (tid == 0) {
        double k, dk;
        kern(WC, WC, k, dk);
        bsc[0] = k;
        bsc[1] = bsc[3] = 0.0;
        bsc[2] = 0.0;
        if (qb > 1) {
            kern(WC, WC + 1, k, dk);
            bsc[1] = k;
            bsc[3] = dk;
            kern(WC + 1, WC + 1, k, dk);
            bsc[2] = k;
        }
    }
    __syncthreads();

    double *dw = vecs + wid * (7 * WC + 8);
    double *uu = dw + WC + 8, *vv = uu + WC, *x1 = vv + WC, *x2 = x1 + WC, *Gv = x2 + WC, *Hv = Gv + WC;
    const int stride = gridDim.x * WARPS;
    for (int seg = blockIdx.x * WARPS + wid; seg < nseg; seg += stride) {
        __syncwarp();
        for (int t = lane; t < WC + 2; t += 32)
            dw[t] = t < W ? sqrt(fmax(w[((size_t)seg * W + t) * L + l], 0.0)) : 0.0;
        __syncwarp();
        // ---- border: B22^-1 = [[p, q], [q, rr]], X = B12 B22^-1, rows of C21 --------------------------------------
        const double da = dw[WC], db = dw[WC + 1];                 // db = 0 when there is one border row
        const double b00 = fma(da * bsc[0], da, 1.0), b01 = da * bsc[1] * db, b11 = fma(db * bsc[2], db, 1.0);
        const double det = fma(b00, b11, -b01 * b01);
        const bool okb = b00 > 0.0 && det > 0.0;                   // the two pivots of the border block
        const double idet = 1.0 / det;
        const double p = b11 * idet, q = -b01 * idet, rr = b00 * idet;
        const double c01 = da * bsc[3] * db;
        for (int t = lane; t < WC; t += 32) {
            const double di = dw[t];
            const double u = di * bord[t] * da, v = di * bord[WC + t] * db;
            const double xa = fma(p, u, q * v), xb = fma(q, u, rr * v);
            const double g = da * bord[2 * WC + t] * di, h = db * bord[3 * WC + t] * di;
            uu[t] = u;
            vv[t] = v;
            x1[t] = xa;
            x2[t] = xb;
            Gv[t] = fma(-0.5 * c01, xb, g);
            Hv[t] = fma(-0.5 * c01, xa, h);
        }
        __syncwarp();
        // ---- S = B11 - u x1' - v x2' in tiles ----------------------------------------------------------------------
        Tile A[NT];
#pragma unroll
        for (int i = 0; i < NBC; ++i) {
            const double di = dw[8 * i + r], ui = uu[8 * i + r], vi = vv[8 * i + r];
#pragma unroll
            for (int j = 0; j <= i; ++j) {
                const double2 kv = Ks[tix(i, j) * 32 + lane];
                const int cj = 8 * j + c0;
                double ax = fma(di * kv.x, dw[cj], (i == j && r == c0) ? 1.0 : 0.0);
                double ay = fma(di * kv.y, dw[cj + 1], (i == j && r == c0 + 1) ? 1.0 : 0.0);
                ax = fma(-ui, x1[cj], ax);
                ay = fma(-ui, x1[cj + 1], ay);
                A[tix(i, j)].x = fma(-vi, x2[cj], ax);
                A[tix(i, j)].y = fma(-vi, x2[cj + 1], ay);
            }
        }
        bool ok = tile_sweep<NBC>(A, lane) && okb;            // tiles hold -S^-1
        // ---- tr(B^-1) and (d B^-1 d) : dK ---------------------------------------------------------------------------
        double tr = 0.0, pd = 0.0;
#pragma unroll
        for (int i = 0; i < NBC; ++i) {
            const int ri = 8 * i + r;
            const double di = dw[ri], x1i = x1[ri], x2i = x2[ri], Gi = Gv[ri], Hi = Hv[ri];
#pragma unroll
            for (int j = 0; j <= i; ++j) {
                const double2 dv = dKs[tix(i, j) * 32 + lane];
                const double wgt = (i == j) ? 1.0 : 2.0;
                const double sx = -wgt * A[tix(i, j)].x, sy = -wgt * A[tix(i, j)].y;      // S^-1 entries, weighted
                const int cj = 8 * j + c0;
                {
                    const double x1j = x1[cj], x2j = x2[cj];
                    double kap = di * dv.x * dw[cj];
                    kap = fma(-x1i, Gv[cj], kap);
                    kap = fma(-x2i, Hv[cj], kap);
                    kap = fma(-x1j, Gi, kap);
                    kap = fma(-x2j, Hi, kap);
                    pd = fma(sx, kap, pd);
                    tr = fma(sx, fma(x1i, x1j, x2i * x2j), tr);
                }
                {
                    const double x1j = x1[cj + 1], x2j = x2[cj + 1];
                    double kap = di * dv.y * dw[cj + 1];
                    kap = fma(-x1i, Gv[cj + 1], kap);
                    kap = fma(-x2i, Hv[cj + 1], kap);
                    kap = fma(-x1j, Gi, kap);
                    kap = fma(-x2j, Hi, kap);
                    pd = fma(sy, kap, pd);
                    tr = fma(sy, fma(x1i, x1j, x2i * x2j), tr);
                }
                if (i == j) {
                    if (r == c0) tr += sx;
                    if (r == c0 + 1) tr += sy;
                }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            tr += __shfl_xor_sync(FULL, tr, o);
            pd += __shfl_xor_sync(FULL, pd, o);
        }
        if (lane == 0) {
            tr += p + (qb > 1 ? rr : 0.0);
            pd = fma(2.0 * q, c01, pd);
            if (!ok) tr = pd = __longlong_as_double(0x7ff8000000000000LL);
            part[seg] = tr;
            part[nseg + seg] = pd;
        }
    }
}

template <int NBC>
int launch_schur(vlgp_ctx *ctx, TrialSet *ts, const HEvalBatch &eb) {
    constexpr int NT = NBC * (NBC + 1) / 2, WC = 8 * NBC;
    const size_t smem = (size_t)2 * NT * 32 * sizeof(double2) + (4 * WC + 8 + WARPS * (7 * WC + 8)) * sizeof(double);
    const int S = ts->n_trials;
    if (ts->dmma_grid == 0) {
        if (smem > 48 * 1024)
            CK(cudaFuncSetAttribute(hstep_segment_schur_kernel<NBC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int per_sm = 1;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, hstep_segment_schur_kernel<NBC>, WARPS * 32, smem));
        if (per_sm < 1) per_sm = 1;
        ts->dmma_grid = per_sm * ctx->prop.multiProcessorCount;
        if (ts->dmma_grid > (S + WARPS - 1) / WARPS) ts->dmma_grid = (S + WARPS - 1) / WARPS;
    }
    hstep_segment_schur_kernel<NBC><<<dim3(ts->dmma_grid, eb.n), WARPS * 32, smem, ctx->stream>>>(
        eb, S, ts->max_len, ctx->L, ctx->dt, ts->d_w, ts->d_hpart);
    CKL();
    return VLGP_OK;
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
    // one or two bins beyond a multiple of 8 (the default window of 50): bordered variant
    const int qb = ts->max_len - 8 * (NB - 1);
    if (qb <= 2 && NB >= 3 && NB <= 7 && !getenv("VLGP_HSTEP_NO_SCHUR")) {
        switch (NB - 1) {
            case 2: rc = launch_schur<2>(ctx, ts, eb); break;
            case 3: rc = launch_schur<3>(ctx, ts, eb); break;
            case 4: rc = launch_schur<4>(ctx, ts, eb); break;
            case 5: rc = launch_schur<5>(ctx, ts, eb); break;
            default: rc = launch_schur<6>(ctx, ts, eb); break;
        }
        if (rc == VLGP_OK) *handled = true;
        return rc;
    }
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
