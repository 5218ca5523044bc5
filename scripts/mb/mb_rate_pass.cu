// Micro-benchmark of the tensor-path rate pass (estep_seg_impl.cuh: rate_tiles): what one column-tile step costs per
// SM sub-partition as a function of what is left in it and of the number of warps resident.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -I vlgp_b200/csrc scripts/mb/mb_rate_pass.cu -o /tmp/mb_rate
#include <cstdio>
#include <cstdlib>
#include "common.cuh"
#include "dmma.cuh"

int vlgp_fail(vlgp_ctx *, int code, const char *, ...) { return code; }

constexpr int LT = 5, KS = 3, NP = 104, N = 100, W = 50;

// VAR bits: 1 = x contraction by DMMA, 2 = exponentials, 4 = accumulation by DMMA, 8 = accumulation by DFMA instead,
// 16 = table lookups from shared memory instead of __ldg
template <int VAR>
__global__ void __launch_bounds__(256, 3) mb_kernel(int reps, int warps_active, const double *Bx_g, const uint8_t *y_g, double *out) {
    extern __shared__ __align__(16) unsigned char raw[];
    double *Bx = (double *)raw;                    // 12 x 104
    double *mu = Bx + 12 * NP, *v = mu + W * LT;
    double *tab = v + W * LT;
    uint8_t *ys = (uint8_t *)(tab + 32);
    for (int i = threadIdx.x; i < 12 * NP; i += 256) Bx[i] = Bx_g[i];
    for (int i = threadIdx.x; i < W * LT; i += 256) { mu[i] = 0.01 * (i % 7); v[i] = 0.001 * (i % 5); }
    for (int i = threadIdx.x; i < W * N; i += 256) ys[i] = y_g[i];
    if (threadIdx.x < 32) tab[threadIdx.x] = VLGP_EXP_T[threadIdx.x];
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, r = lane >> 2, q = lane & 3;
    if (wid >= warps_active) return;
    const int t = 8 * wid + r;
    const bool tin = t < W;
    double afr[KS];
    for (int kk = 0; kk < KS; ++kk) {
        const int k = 4 * kk + q;
        double val = 0.0;
        if (tin) {
            if (k < LT) val = mu[t * LT + k];
            else if (k < 2 * LT) val = v[t * LT + k - LT];
            else if (k == 2 * LT) val = 1.0;
        }
        afr[kk] = val;
    }
    Tile acc{0.0, 0.0};
    double accs[LT];
    for (int l = 0; l < LT; ++l) accs[l] = 0.0;
    const uint8_t *yrow = ys + (tin ? t : 0) * N;
    for (int rep = 0; rep < reps; ++rep) {
        for (int j = 0; j < NP / 8; ++j) {
            Tile x{0.0, 0.0};
            if (VAR & 1) {
#pragma unroll
                for (int kk = 0; kk < KS; ++kk) dmma(x, afr[kk], Bx[(4 * kk + q) * NP + 8 * j + r]);
            } else {
                x.x = afr[0] * Bx[q * NP + 8 * j + r];
                x.y = afr[1] * Bx[(4 + q) * NP + 8 * j + r];
            }
            double e0 = x.x, e1 = x.y;
            if (VAR & 2) {
                if (VAR & 16) {
                    // same arithmetic, table in shared memory
                    double xx0 = x.x > 10.0 ? 10.0 : x.x, xx1 = x.y > 10.0 ? 10.0 : x.y;
                    xx0 = xx0 < -708.0 ? -708.0 : xx0;
                    xx1 = xx1 < -708.0 ? -708.0 : xx1;
                    const double shift = 6755399441055744.0;
                    const double m0 = fma(xx0, VLGP_EXP_INV, shift), m1 = fma(xx1, VLGP_EXP_INV, shift);
                    const int i0 = __double2loint(m0), i1 = __double2loint(m1);
                    const double tj0 = tab[i0 & 31], tj1 = tab[i1 & 31];
                    const double t0 = m0 - shift, t1 = m1 - shift;
                    double r0 = fma(t0, -VLGP_EXP_HI, xx0), r1 = fma(t1, -VLGP_EXP_HI, xx1);
                    r0 = fma(t0, -VLGP_EXP_LO, r0);
                    r1 = fma(t1, -VLGP_EXP_LO, r1);
                    double p0 = VLGP_EXP_C[0], p1 = VLGP_EXP_C[0];
#pragma unroll
                    for (int k = 1; k < 6; ++k) {
                        p0 = fma(p0, r0, VLGP_EXP_C[k]);
                        p1 = fma(p1, r1, VLGP_EXP_C[k]);
                    }
                    p0 *= r0;
                    p1 *= r1;
                    const double q0 = fma(tj0, p0, tj0), q1 = fma(tj1, p1, tj1);
                    const int ex0 = max(i0 >> 5, -1022), ex1 = max(i1 >> 5, -1022);
                    e0 = __hiloint2double(__double2hiint(q0) + (ex0 << 20), __double2loint(q0));
                    e1 = __hiloint2double(__double2hiint(q1) + (ex1 << 20), __double2loint(q1));
                } else {
                    trunc_exp2(x.x, x.y, e0, e1);
                }
            }
            const int n0 = 8 * j + 2 * q;
            const double y0 = (tin && n0 < N) ? (double)yrow[n0] : 0.0;
            const double y1 = (tin && n0 + 1 < N) ? (double)yrow[n0 + 1] : 0.0;
            const double c0 = y0 - e0, c1 = y1 - e1;
            if (VAR & 4) {
                double2 b2 = make_double2(0.0, 0.0);
                if (r < LT) b2 = *reinterpret_cast<const double2 *>(Bx + r * NP + n0);
                dmma(acc, c0, b2.x);
                dmma(acc, c1, b2.y);
            } else if (VAR & 8) {
#pragma unroll
                for (int l = 0; l < LT; ++l) {
                    const double2 b2 = *reinterpret_cast<const double2 *>(Bx + l * NP + n0);
                    accs[l] = fma(c0, b2.x, accs[l]);
                    accs[l] = fma(c1, b2.y, accs[l]);
                }
            } else {
                acc.x += c0;
                acc.y += c1;
            }
        }
    }
    double sum = acc.x + acc.y;
    for (int l = 0; l < LT; ++l) sum += accs[l];
    if (sum == 1.2345) out[threadIdx.x] = sum;
}

template <int VAR>
void run(const char *name, int ctas_per_sm, int warps_active, const double *Bx, const uint8_t *y, double *out) {
    const int reps = 400;
    size_t smem = ctas_per_sm == 3 ? 70000 : (ctas_per_sm == 2 ? 100000 : (ctas_per_sm == 1 ? 200000 : 50000));
    cudaFuncSetAttribute(mb_kernel<VAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int grid = 148 * ctas_per_sm;
    mb_kernel<VAR><<<grid, 256, smem>>>(10, warps_active, Bx, y, out);
    cudaEventRecord(e0);
    mb_kernel<VAR><<<grid, 256, smem>>>(reps, warps_active, Bx, y, out);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaError_t err = cudaGetLastError();
    // tile steps per SM sub-partition
    const double steps = (double)ctas_per_sm * warps_active * reps * 13 / 4.0;
    const double cyc = ms * 1e-3 * 1.965e9 / steps;
    printf("%-44s ctas/SM %d warps/CTA %d : %7.3f ms  %6.1f cycles per tile step per SMSP  (%s)\n", name, ctas_per_sm, warps_active, ms, cyc,
           cudaGetErrorString(err));
}

int main() {
    double *Bx, *out;
    uint8_t *y;
    cudaMalloc(&Bx, 12 * NP * 8);
    cudaMalloc(&out, 4096);
    cudaMalloc(&y, W * N);
    double h[12 * NP];
    for (int i = 0; i < 12 * NP; ++i) h[i] = 0.01 * ((i * 7) % 13 - 6);
    cudaMemcpy(Bx, h, sizeof(h), cudaMemcpyHostToDevice);
    cudaMemset(y, 1, W * N);
    for (int c : {3}) {
        for (int w : {8, 7, 6, 4}) {
            run<7>("full (x DMMA, exp, acc DMMA)", c, w, Bx, y, out);
            run<5>("no exp", c, w, Bx, y, out);
            run<2>("exp only", c, w, Bx, y, out);
            run<18>("exp only, table in smem", c, w, Bx, y, out);
            run<3>("x DMMA + exp", c, w, Bx, y, out);
            run<11>("x DMMA + exp + acc by DFMA", c, w, Bx, y, out);
            run<23>("full, table in smem", c, w, Bx, y, out);
            run<27>("x DMMA + exp(smem table) + acc DFMA", c, w, Bx, y, out);
        }
    }
    return 0;
}
