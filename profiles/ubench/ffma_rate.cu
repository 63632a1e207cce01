// Micro-benchmark: issue rate of the FFMA pattern of custom::Correlation's inner loop (acc[j][k] += a[k] * b[m], 108
// accumulators, a and b in registers) as scalar FFMA and as packed FFMA2, for 1 / 2 / 3 warps per scheduler.
// Reports SM-cycles per warp-FMA-instruction-equivalent (one lane-FMA x 32) per scheduler: 1.0 = the FP32 peak.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma_rate ffma_rate.cu
#include <cstdio>
#include <cuda_runtime.h>

typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float lo, float hi) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }

template <int MODE>
__global__ void __launch_bounds__(384, 1) k(int iters, float* out, long long* cycles, float s0, float s1)
{
    float a[4] = {s0, s1, s0 + 1.0f, s1 + 1.0f};
    float b[12];
#pragma unroll
    for (int m = 0; m < 12; ++m)
        b[m] = s0 * (m + 1) + threadIdx.x * 1e-6f;
    float acc[3][9][4];
    f32x2 accp[3][4][4];
    float accs[3][4];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
            for (int j = 0; j < 9; ++j)
                acc[i][j][kk] = 0.0f;
#pragma unroll
            for (int p = 0; p < 4; ++p)
                accp[i][kk][p] = 0ull;
            accs[i][kk] = 0.0f;
        }
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int m = 0; m < 12; ++m)
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) {
                        const int jj = m - kk;
                        if (jj >= 0 && jj < 9)
                            acc[i][jj][kk] = __fmaf_rn(a[kk], b[m], acc[i][jj][kk]);
                    }
        } else {
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int m = 0; m < 12; m += 2) {
                    const f32x2 bp = pk2(b[m], b[m + 1]);
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) {
                        const int jj = m - kk;
                        if (jj >= 0 && jj + 1 < 9) {
                            const int pp = (kk & 1) ? (jj - 1) / 2 : jj / 2;
                            accp[i][kk][pp] = fma2(pk2(a[kk], a[kk]), bp, accp[i][kk][pp]);
                        } else if (jj == 8) {
                            accs[i][kk] = __fmaf_rn(a[kk], b[m], accs[i][kk]);
                        } else if (jj == -1) {
                            accs[i][kk] = __fmaf_rn(a[kk], b[m + 1], accs[i][kk]);
                        }
                    }
                }
        }
        // rotate the multiplicands so that nothing is loop-invariant (4 + 12 moves per 108 FMAs)
        const float t = a[0]; a[0] = a[1]; a[1] = a[2]; a[2] = a[3]; a[3] = t;
    }
    const long long t1 = clock64();
    float sum = 0.0f;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
            for (int j = 0; j < 9; ++j)
                sum += acc[i][j][kk];
#pragma unroll
            for (int p = 0; p < 4; ++p)
                sum += __uint_as_float(static_cast<unsigned>(accp[i][kk][p])) + __uint_as_float(static_cast<unsigned>(accp[i][kk][p] >> 32));
            sum += accs[i][kk];
        }
    out[blockIdx.x * blockDim.x + threadIdx.x] = sum;
    if (threadIdx.x == 0 && blockIdx.x == 0)
        *cycles = t1 - t0;
}

int main()
{
    float* out;
    long long* cyc;
    cudaMalloc(&out, 148 * 384 * sizeof(float));
    cudaMallocManaged(&cyc, sizeof(long long));
    const int iters = 20000;
    for (int mode = 0; mode < 2; ++mode)
        for (int warps = 4; warps <= 12; warps += 4) {
            for (int rep = 0; rep < 2; ++rep) {
                if (mode == 0)
                    k<0><<<148, warps * 32>>>(rep ? iters : 100, out, cyc, 1.0f, 0.5f);
                else
                    k<1><<<148, warps * 32>>>(rep ? iters : 100, out, cyc, 1.0f, 0.5f);
                cudaDeviceSynchronize();
            }
            // 108 lane-FMAs per thread and iteration; warps / 4 warps per scheduler
            const double per = static_cast<double>(*cyc) / (static_cast<double>(iters) * 108 * (warps / 4));
            printf("%s  %2d warps/SM  %.3f cycles per warp-FMA per scheduler (%s)\n", mode ? "FFMA2" : "FFMA ", warps, per,
                cudaGetErrorString(cudaGetLastError()));
        }
    return 0;
}
