// Block-cooperative small dense linear algebra in shared memory (n <= 64), used by the E-step and H-step kernels.
#pragma once
#include "common.cuh"

// In-place symmetric sweep of an SPD matrix held in full (both triangles) in SMEM with leading dimension ld:
// on return Aw = -A^-1.  Gauss-Jordan without pivoting; the k-th pivot equals the k-th Schur-complement pivot of the
// Cholesky factorisation (d_k = L_kk^2), so "pivot <= 0" is exactly LAPACK potrf's "not positive definite", and
// sum_k log(d_k) = log det A.  Must be called by all 256 threads of the CTA; ck is a 64-double SMEM scratch.
// Returns false (uniformly) on a non-positive pivot; *logdet (if non-null) receives sum_k log(d_k).
__device__ __forceinline__ bool block_sweep_spd(double *Aw, int ld, int n, double *ck, double *logdet) {
    const int tid = threadIdx.x;
    const int ty = tid >> 4, tx = tid & 15;
    double lsum = 0.0;
    for (int k = 0; k < n; ++k) {
        const double d = Aw[k * ld + k];
        if (!(d > 0.0)) return false;               // uniform: every thread reads the same value
        if (logdet) lsum += log(d);
        const double pinv = 1.0 / d;
        if (tid < n) ck[tid] = Aw[tid * ld + k];
        __syncthreads();
        for (int i = ty; i < n; i += 16) {
            const double cip = ck[i] * pinv;
            for (int j = tx; j < n; j += 16) {
                double val;
                if (i == k) val = (j == k) ? -pinv : ck[j] * pinv;
                else if (j == k) val = cip;
                else val = fma(-cip, ck[j], Aw[i * ld + j]);
                Aw[i * ld + j] = val;
            }
        }
        __syncthreads();
    }
    if (logdet) *logdet = lsum;
    return true;
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
