// Warp-level FP64 tensor-core (mma.sync.m8n8k4.f64, "DMMA") building blocks shared by the H-step and E-step kernels:
// 8 x 8 tiles in the accumulator layout, operand fragments made from accumulator tiles by intra-warp shuffles, and the
// in-warp inverse of an SPD 8 x 8 tile.  tcgen05 has no f64 kind, so this is the FP64 tensor path of sm_100a
// (measured 37.0 TFLOP/s vs 34.0 for DFMA on this pool's B200).
#pragma once
#include "common.cuh"

constexpr unsigned FULL = 0xffffffffu;

struct Tile {
    double x, y;      // [lane/4][2 (lane%4)], [lane/4][2 (lane%4) + 1]
};

__device__ __forceinline__ void dmma(Tile &c, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c.x), "+d"(c.y)
                 : "d"(a), "d"(b));
}

// Tile[lane/4][4h + lane%4]
__device__ __forceinline__ double nform(const Tile &t, int h, int lane) {
    const int src = (lane & ~3) | (2 * h + ((lane & 3) >> 1));
    const double a = __shfl_sync(FULL, t.x, src), b = __shfl_sync(FULL, t.y, src);
    return (lane & 1) ? b : a;
}

// Tile[4h + lane%4][lane/4]
__device__ __forceinline__ double tform(const Tile &t, int h, int lane) {
    const int src = 4 * (4 * h + (lane & 3)) + (lane >> 3);
    const double a = __shfl_sync(FULL, t.x, src), b = __shfl_sync(FULL, t.y, src);
    return ((lane >> 2) & 1) ? b : a;
}

// In-place inverse of an SPD 8 x 8 tile (accumulator layout) by an 8-step scalar sweep.  Returns false (warp-uniform)
// if a pivot is not positive, i.e. the matrix is not positive definite.
__device__ __forceinline__ bool tile_spd_inverse(Tile &t, int lane) {
    const int r = lane >> 2, c0 = 2 * (lane & 3);
    bool ok = true;
#pragma unroll
    for (int p = 0; p < 8; ++p) {
        const double comp = (p & 1) ? t.y : t.x;
        const double d = __shfl_sync(FULL, comp, 4 * p + (p >> 1));            // a[p][p]
        const double cr = __shfl_sync(FULL, comp, (lane & ~3) | (p >> 1));     // a[r][p]
        const double pc0 = __shfl_sync(FULL, t.x, 4 * p + (lane & 3));         // a[p][c0]
        const double pc1 = __shfl_sync(FULL, t.y, 4 * p + (lane & 3));         // a[p][c0 + 1]
        ok = ok && (d > 0.0);
        const double pinv = fast_rcp(d);
        const double crp = cr * pinv;
        double nx = fma(-crp, pc0, t.x), ny = fma(-crp, pc1, t.y);
        if (r == p) {
            nx = pc0 * pinv;
            ny = pc1 * pinv;
        }
        if (c0 == p) nx = (r == p) ? -pinv : crp;
        if (c0 + 1 == p) ny = (r == p) ? -pinv : crp;
        t.x = nx;
        t.y = ny;
    }
    t.x = -t.x;
    t.y = -t.y;
    return ok;
}

__host__ __device__ constexpr int tix(int i, int j) { return i * (i + 1) / 2 + j; }

// Blocked symmetric sweep of an SPD matrix of order 8 NB held by ONE warp as the lower block triangle of 8 x 8 tiles in
// the accumulator layout (A[tix(i, j)], j <= i): on return the tiles hold -A^-1.  Block step kb: P = A_kk^-1 (scalar
// sweep by shuffles), A_ik <- A_ik P, A_ij <- A_ij - A_ik P A_kj (i, j != kb) by DMMA, A_kk <- -P.  Returns false
// (warp-uniform) if some pivot was not positive.  (Same algorithm as the H-step's per-segment kernel and the segment
// E-step's factorisation; used by the long-trial E-step, csrc/estep_long.cu.)
template <int NB>
__device__ __forceinline__ bool tile_sweep(Tile (&A)[NB * (NB + 1) / 2], int lane) {
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
    return ok;
}
