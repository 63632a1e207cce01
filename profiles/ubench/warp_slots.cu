// Which hardware warp slots (%warpid; scheduler = slot % 4) do the warps of two co-resident 6-warp CTAs get?
// (custom::Correlation's 32x8 shared-row kernel runs two 192-thread CTAs per SM.)
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(192, 2) k(int* out)
{
    extern __shared__ float sm[];
    unsigned smid, wid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    asm volatile("mov.u32 %0, %%warpid;" : "=r"(wid));
    if ((threadIdx.x & 31) == 0) {
        int* o = out + (blockIdx.x * 6 + (threadIdx.x >> 5)) * 2;
        o[0] = smid;
        o[1] = wid;
    }
    // stay resident so that the second CTA of the SM is co-resident
    const long long t0 = clock64();
    while (clock64() - t0 < 2000000) { }
    sm[threadIdx.x] = 0.0f;
}
int main()
{
    int* out;
    cudaMallocManaged(&out, 296 * 6 * 2 * sizeof(int));
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 102912);
    cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    k<<<296, 192, 102912>>>(out);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    int hist[8] = {0};
    for (int sm = 0; sm < 3; ++sm) {
        printf("SM %d:", sm);
        for (int b = 0; b < 296; ++b)
            if (out[b * 12] == sm) {
                printf("  cta %d slots", b);
                for (int w = 0; w < 6; ++w)
                    printf(" %d", out[(b * 6 + w) * 2 + 1]);
            }
        printf("\n");
    }
    // per SM: warps per scheduler
    int worst[5] = {0};
    for (int sm = 0; sm < 148; ++sm) {
        int cnt[4] = {0};
        for (int i = 0; i < 296 * 6; ++i)
            if (out[i * 2] == sm)
                cnt[out[i * 2 + 1] % 4]++;
        int mx = 0;
        for (int j = 0; j < 4; ++j) mx = cnt[j] > mx ? cnt[j] : mx;
        worst[mx > 4 ? 4 : mx]++;
    }
    printf("SMs by busiest scheduler's warp count: 3 warps: %d, 4 warps: %d, other: %d %d %d\n", worst[3], worst[4], worst[0], worst[1], worst[2]);
    (void)hist;
    return 0;
}
