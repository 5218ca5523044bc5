// Throughput of individual FP64 instruction forms on sm_100a: cycles per warp instruction per SM sub-partition.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 scripts/mb/mb_fp64_ops.cu -o scripts/mb/mb_fp64
#include <cstdio>
#include <cuda_runtime.h>

__constant__ double CC[8] = {1.0 / 720.0, 1.0 / 120.0, 1.0 / 24.0, 1.0 / 6.0, 0.5, 1.0, 0.3, 0.7};

template <int OP>
__global__ void __launch_bounds__(256) k(int iters, double *out, double seed, double seed2) {
    double a[8];
    for (int i = 0; i < 8; ++i) a[i] = seed + i * 1e-3 + threadIdx.x * 1e-6;
    double y = seed2, z = seed2 * 0.5;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (OP == 0) a[i] = fma(a[i], y, z);                       // registers
                if (OP == 1) a[i] = fma(a[i], y, CC[(u + i) & 7]);         // constant bank operand
                if (OP == 2) a[i] = fma(a[i], y, 0.333333333333);          // immediate-like literal
                if (OP == 3) a[i] = a[i] + y;                              // DADD
                if (OP == 4) a[i] = a[i] * y;                              // DMUL
                if (OP == 5) a[i] = a[i] > 10.0 ? z : a[i] + 0.0 * y;      // DSETP + select (+ one DFMA)
                if (OP == 6) a[i] = a[i] > z ? y : a[i];                   // DSETP + 2 FSEL only
                if (OP == 7) a[i] = fma(a[i], CC[(u + i) & 7], z);         // constant as multiplicand
                if (OP == 8) a[i] = a[i] - 6755399441055744.0;             // DADD with a literal
            }
        }
    }
    double s = 0.0;
    for (int i = 0; i < 8; ++i) s += a[i];
    if (s == 1.2345) out[0] = s;
}

template <int OP>
void run(const char *name, double *out, int per) {
    const int iters = 2000, grid = 148 * 8;
    k<OP><<<grid, 256>>>(10, out, 1.0, 0.999);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<OP><<<grid, 256>>>(iters, out, 1.0, 0.999);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double warp_instr_per_smsp = (double)grid * 8 * iters * 32.0 * per / (148.0 * 4.0);
    printf("%-40s %8.3f ms  %5.2f cycles per FP64-pipe warp instruction per SMSP (%d per statement)\n", name, ms,
           ms * 1e-3 * 1.965e9 / warp_instr_per_smsp, per);
}

int main() {
    double *out;
    cudaMalloc(&out, 64);
    run<0>("DFMA r,r,r", out, 1);
    run<1>("DFMA r,r,c[]", out, 1);
    run<2>("DFMA r,r,literal", out, 1);
    run<7>("DFMA r,c[],r", out, 1);
    run<3>("DADD r,r", out, 1);
    run<8>("DADD r,literal", out, 1);
    run<4>("DMUL r,r", out, 1);
    run<5>("DSETP + sel + DFMA", out, 2);
    run<6>("DSETP + 2 FSEL", out, 1);
    return 0;
}
