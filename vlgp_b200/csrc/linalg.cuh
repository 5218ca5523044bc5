// Block-cooperative small dense linear algebra in shared memory (n <= 64), used by the E-step and H-step kernels.
#pragma once
#include "common.cuh"

// Symmetric sweep (Gauss-Jordan without pivoting) of an SPD matrix of order n <= 64 held in REGISTERS: the 256 threads
// of the CTA form a 16 x 16 grid and thread (ty, tx) owns the elements (ty + 16 p, tx + 16 q), p, q < 4.  On return the
// registers hold -A^-1.  The k-th pivot equals the k-th Schur-complement pivot of the Cholesky factorisation
// (d_k = L_kk^2), so "pivot <= 0" is exactly LAPACK potrf's "not positive definite" and sum_k log d_k = log det A.
// Per step only the current column travels through shared memory (ck: 2 x 64 doubles, double-buffered -> one barrier
// per step).  Elements with an index >= n must be zero on entry and stay zero.
// Returns false (uniformly) on a non-positive pivot; *logdet (if non-null) receives sum_k log d_k.
// piv (optional, n doubles of SMEM): receives the pivots instead of their logarithms being summed inside the loop (the
// caller takes the n logarithms in parallel afterwards: log() on the critical path of every step is most of its cost).
__device__ __forceinline__ bool block_sweep_regs(double (&a)[4][4], int n, double *ck, double *logdet,
                                                 double *piv = nullptr) {
    const int tid = threadIdx.x;
    const int ty = tid >> 4, tx = tid & 15;
    double lsum = 0.0;
    bool ok = true;
    for (int k = 0; k < n; ++k) {
        double *buf = ck + (k & 1) * 64;
        const int kq = k >> 4, kr = k & 15;
        if (tx == kr) {
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (q == kq) {
#pragma unroll
                    for (int p = 0; p < 4; ++p) buf[ty + 16 * p] = a[p][q];
                }
        }
        __syncthreads();
        const double d = buf[k];
        if (!(d > 0.0)) { ok = false; break; }      // uniform: every thread reads the same value
        if (piv) {
            if (tid == 0) piv[k] = d;
        } else if (logdet) {
            lsum += log(d);
        }
        const double pinv = fast_rcp(d);
        double ci[4], cj[4];
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            ci[p] = buf[ty + 16 * p] * pinv;         // entries >= n were written as zeros by their owners
            cj[p] = buf[tx + 16 * p];
        }
#pragma unroll
        for (int p = 0; p < 4; ++p)
#pragma unroll
            for (int q = 0; q < 4; ++q) a[p][q] = fma(-ci[p], cj[q], a[p][q]);
        if (ty == kr) {                              // row k <- c_j / d
#pragma unroll
            for (int p = 0; p < 4; ++p)
                if (p == kq) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) a[p][q] = cj[q] * pinv;
                }
        }
        if (tx == kr) {                              // column k <- c_i / d, pivot <- -1/d
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (q == kq) {
#pragma unroll
                    for (int p = 0; p < 4; ++p) a[p][q] = (ty + 16 * p == k) ? -pinv : ci[p];
                }
        }
    }
    __syncthreads();
    if (logdet) *logdet = lsum;
    return ok;
}

// Same sweep for a matrix held in full (both triangles) in SMEM with leading dimension ld: loads it into the register
// layout above, sweeps, stores -A^-1 back.  Must be called by all 256 threads.
__device__ __forceinline__ bool block_sweep_spd(double *Aw, int ld, int n, double *ck, double *logdet,
                                                double *piv = nullptr) {
    const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
    double a[4][4];
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int i = ty + 16 * p, j = tx + 16 * q;
            a[p][q] = (i < n && j < n) ? Aw[i * ld + j] : 0.0;
        }
    const bool ok = block_sweep_regs(a, n, ck, logdet, piv);
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int i = ty + 16 * p, j = tx + 16 * q;
            if (i < n && j < n) Aw[i * ld + j] = a[p][q];
        }
    __syncthreads();
    return ok;
}

__device__ __forceinline__ void tri_decode(int idx, int &bi, int &bj) {
    int r = (int)((sqrtf(8.0f * (float)idx + 1.0f) - 1.0f) * 0.5f);
    while ((r + 1) * (r + 2) / 2 <= idx) ++r;
    while (r * (r + 1) / 2 > idx) --r;
    bi = r;
    bj = idx - r * (r + 1) / 2;
}

// Deterministic block-wide sum (256 threads); result valid in every thread.  red: 32 doubles of SMEM.
__device__ __forceinline__ double block_sum(double x, double *red) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    x = warp_sum(x);
    __syncthreads();
    if (lane == 0) red[wid] = x;
    __syncthreads();
    double r = 0.0;
    const int nw = blockDim.x >> 5;
    for (int i = 0; i < nw; ++i) r += red[i];
    return r;
}
