// Micro-benchmark: how many shared-memory wavefronts does one LDS.128 cost when several lanes of a warp read the SAME
// 16 bytes?  (Design question for custom::Correlation: lanes of different displacement groups that share in2 rows.)
// Patterns 6-10: the lane layouts of correlation_md4_share_kernel.  Patterns (quad index read by lane l):  0: l (512 B unique)   1: l % 16 (two half-warps read the same 256 B)
//   2: l / 2 (adjacent lane pairs share)   3: l % 8 (128 B unique)   4: 0 (one quad)
//   5: (l % 8) + 18 * (l / 24)  (three groups of 8 lanes share one 128 B row segment, the fourth reads another row)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lds_merge lds_merge.cu ; run: ./lds_merge
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(384, 1) k(int pattern, int iters, float* out, long long* cycles)
{
    __shared__ __align__(16) float4 s[2048];
    for (int i = threadIdx.x; i < 2048; i += blockDim.x)
        s[i] = make_float4(i, 1.0f, 2.0f, 3.0f);
    __syncthreads();
    const int l = threadIdx.x & 31;
    int q;
    switch (pattern) {
        case 0: q = l; break;
        case 1: q = l % 16; break;
        case 2: q = l / 2; break;
        case 3: q = l % 8; break;
        case 4: q = 0; break;
        case 5: q = (l % 8) + 18 * (l / 24); break;
        case 6: q = l / 3; break;                          // items (quad, member) flattened: 3 lanes per quad
        case 7: q = (l + 16) / 3; break;                   // the same, other phase
        case 8: q = l / 3 + 17 * 3 * (l % 3); break;       // in1: member m reads row 3m of a 68-float-stride tile
        case 9: q = (l + 16) / 3 + 17 * 3 * ((l + 16) % 3); break;
        default: q = l / 3 + 16 * 3 * (l % 3); break;      // in1 with 64-float rows: bank conflicts expected
    }
    q += (threadIdx.x >> 5) * 32;   // every warp its own region
    float4 a0 = make_float4(0, 0, 0, 0), a1 = a0, a2 = a0, a3 = a0;
    const float4* p = s + q;
    const long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
        // 8 independent loads per iteration at rotating offsets (all within the warp's region + 1024 quads)
        float4 v0 = p[0], v1 = p[32 * 12], v2 = p[64 * 12 % 1024], v3 = p[96], v4 = p[128], v5 = p[160], v6 = p[192], v7 = p[224];
        a0.x += v0.x; a0.y += v0.y; a0.z += v0.z; a0.w += v0.w;
        a1.x += v1.x; a1.y += v1.y; a1.z += v1.z; a1.w += v1.w;
        a2.x += v2.x; a2.y += v2.y; a2.z += v2.z; a2.w += v2.w;
        a3.x += v3.x; a3.y += v3.y; a3.z += v3.z; a3.w += v3.w;
        a0.x += v4.x; a0.y += v4.y; a0.z += v4.z; a0.w += v4.w;
        a1.x += v5.x; a1.y += v5.y; a1.z += v5.z; a1.w += v5.w;
        a2.x += v6.x; a2.y += v6.y; a2.z += v6.z; a2.w += v6.w;
        a3.x += v7.x; a3.y += v7.y; a3.z += v7.z; a3.w += v7.w;
        p = s + ((q + (i & 1) * 256) & 2047);
    }
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0.x + a0.y + a0.z + a0.w + a1.x + a1.y + a1.z + a1.w + a2.x + a2.y + a2.z + a2.w + a3.x + a3.y + a3.z + a3.w;
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
    const char* names[11] = {"32 distinct quads (512 B)", "l%16: half-warps share (256 B)", "l/2: lane pairs share (256 B)",
        "l%8 (128 B)", "one quad (16 B)", "3 groups share + 1 other row (256 B)", "l/3: three lanes per quad", "(l+16)/3", "in1 rows 3 apart, stride 68", "same, other phase", "in1 rows 3 apart, stride 64"};
    for (int warps = 4; warps <= 12; warps += 8)
        for (int pat = 0; pat < 11; ++pat) {
            k<<<148, warps * 32>>>(pat, 100, out, cyc);
            cudaDeviceSynchronize();
            k<<<148, warps * 32>>>(pat, iters, out, cyc);
            cudaError_t e = cudaDeviceSynchronize();
            const double per_lds = static_cast<double>(*cyc) / (static_cast<double>(iters) * 8 * warps);
            printf("warps %2d  pattern %d  %-40s  %.2f SM-cycles per warp-LDS.128  (%s)\n", warps, pat, names[pat], per_lds,
                cudaGetErrorString(e));
        }
    return 0;
}
