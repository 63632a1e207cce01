// custom::Correlation for sm_100a -- replaces CorrelationKernel::ComputeCUDA,
// blob_rearrange_kernel, correlation_cuda_forward_kernel and correlation_old_kernel
// (reference: src/ort_custom_ops/src/opticalflow/correlation_cuda.cc:29-138,
// correlation_cuda.cu:33-61,98-175,183-265,334-442; CPU definition correlation.cc:148-275).
//
//   out[n, ph, pw, h, w] = sum_c in1[n,c,h,w] * in2[n,c,h+ph-md,w+pw-md]      (outside the image: 0)
//
// The reference transposes both inputs to NHWC into two buffers it cudaMallocs and cudaFrees per
// call, then runs one 32-thread block per output pixel with 81 barriers.  Here the contraction
// reads NCHW directly, with no scratch memory and no allocation:
//   * a CTA owns a 32x8 tile of output pixels for all 81 displacements;
//   * per chunk of KC channels it stages the in1 tile and the in2 tile + 4-pixel halo in shared
//     memory (zero-filled outside the image, which implements the zero padding);
//   * a thread owns 4 consecutive pixels x 9 horizontal x 3 vertical displacements = 108 fp32
//     accumulators in registers and walks the channels in order with explicit FMAs, so every
//     shared-memory word it loads (128-bit LDS) feeds ~2.7 FMAs and each output value is the same
//     sequential-over-c fp32 sum the reference's CPU kernel forms;
//   * results leave as 128-bit stores straight into the [N,9,9,H,W] layout (the legacy
//     [N,81,H,W] layout is byte-identical; it only adds the 1/C scale).
// Any max_displacement other than 4 takes a plain one-thread-per-output kernel (no shipped model
// uses one; model_spec.py:161-162).
#include "vsc_common.cuh"

namespace vsc {

constexpr int kMD = 4;
constexpr int kP = 2 * kMD + 1;           // 9
constexpr int kTW = 32, kTH = 8;          // output tile
constexpr int kBW = kTW + 2 * kMD;        // 40: in2 tile row (floats), 160 B = 16B-aligned rows
constexpr int kBH = kTH + 2 * kMD;        // 16
constexpr int kKC = 8;                    // channels per shared-memory stage
constexpr int kCorrThreads = 192;         // 64 pixel-quads x 3 vertical-displacement groups
constexpr int kASize = kTH * kTW;         // 256 floats per channel
constexpr int kBSize = kBH * kBW;         // 640 floats per channel

__global__ void __launch_bounds__(kCorrThreads) correlation_md4_kernel(const float* __restrict__ in1,
    const float* __restrict__ in2, float* __restrict__ out, int C, int H, int W, float divisor, int legacy, int vec_store)
{
    __shared__ __align__(16) float sA[kKC * kASize];
    __shared__ __align__(16) float sB[kKC * kBSize];

    const int tid = threadIdx.x;
    const int q = tid & 63;        // pixel quad inside the tile
    const int g = tid >> 6;        // vertical displacement group: ph = 3g .. 3g+2 (warp-uniform)
    const int r = q >> 3;          // tile row 0..7
    const int qc = (q & 7) * 4;    // first tile column of the quad
    const int w0 = blockIdx.x * kTW;
    const int h0 = blockIdx.y * kTH;
    const int n = blockIdx.z;
    const size_t HW = static_cast<size_t>(H) * W;
    const float* a_img = in1 + static_cast<size_t>(n) * C * HW;
    const float* b_img = in2 + static_cast<size_t>(n) * C * HW;

    float acc[3][kP][4];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < kP; ++j)
#pragma unroll
            for (int k = 0; k < 4; ++k)
                acc[i][j][k] = 0.0f;

    for (int cbase = 0; cbase < C; cbase += kKC) {
        const int kc = min(kKC, C - cbase);
        __syncthreads();  // previous stage fully consumed
        // in1 tile: kc x 8 x 32
        for (int i = tid; i < kc * kASize; i += kCorrThreads) {
            const int c = i / kASize, rem = i - c * kASize;
            const int y = h0 + rem / kTW, x = w0 + (rem % kTW);
            sA[i] = (y < H && x < W) ? __ldg(a_img + (cbase + c) * HW + static_cast<size_t>(y) * W + x) : 0.0f;
        }
        // in2 tile + halo: kc x 16 x 40
        for (int i = tid; i < kc * kBSize; i += kCorrThreads) {
            const int c = i / kBSize, rem = i - c * kBSize;
            const int y = h0 - kMD + rem / kBW, x = w0 - kMD + (rem % kBW);
            sB[i] = (y >= 0 && y < H && x >= 0 && x < W)
                ? __ldg(b_img + (cbase + c) * HW + static_cast<size_t>(y) * W + x)
                : 0.0f;
        }
        __syncthreads();
        for (int c = 0; c < kc; ++c) {
            const float4 a4 = *reinterpret_cast<const float4*>(&sA[c * kASize + r * kTW + qc]);
            const float av[4] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                const float* brow = &sB[c * kBSize + (r + 3 * g + i) * kBW + qc];
                const float4 b0 = *reinterpret_cast<const float4*>(brow);
                const float4 b1 = *reinterpret_cast<const float4*>(brow + 4);
                const float4 b2 = *reinterpret_cast<const float4*>(brow + 8);
                const float bv[12] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w, b2.x, b2.y, b2.z, b2.w};
#pragma unroll
                for (int j = 0; j < kP; ++j)
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        acc[i][j][k] = __fmaf_rn(av[k], bv[k + j], acc[i][j][k]);
            }
        }
    }

    const int y = h0 + r;
    const int x = w0 + qc;
    if (y >= H || x >= W)
        return;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const int ph = 3 * g + i;
#pragma unroll
        for (int j = 0; j < kP; ++j) {
            float* o = out + ((static_cast<size_t>(n) * kP + ph) * kP + j) * HW + static_cast<size_t>(y) * W + x;
            float v[4];
#pragma unroll
            for (int k = 0; k < 4; ++k)  // legacy: total_sum / (float)C, a true division (correlation_cuda.cu:259-261)
                v[k] = legacy ? acc[i][j][k] / divisor : acc[i][j][k];
            if (vec_store) {  // W % 4 == 0 and out 16B-aligned: x+3 < W holds as well
                stg_stream4(o, make_float4(v[0], v[1], v[2], v[3]));
            } else {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (x + k < W)
                        o[k] = v[k];
            }
        }
    }
}

// any max_displacement: one thread per output value, sequential fp32 sum over c
__global__ void __launch_bounds__(256) correlation_generic_kernel(const float* __restrict__ in1,
    const float* __restrict__ in2, float* __restrict__ out, int C, int H, int W, int md, float scale, int use_div)
{
    const int P = 2 * md + 1;
    const size_t HW = static_cast<size_t>(H) * W;
    const size_t per_n = static_cast<size_t>(P) * P * HW;
    const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= per_n)
        return;
    const int n = blockIdx.y;
    const int w = static_cast<int>(i % W);
    const int h = static_cast<int>((i / W) % H);
    const int pw = static_cast<int>((i / HW) % P);
    const int ph = static_cast<int>(i / (HW * P));
    const int h2 = h + ph - md, w2 = w + pw - md;
    float acc = 0.0f;
    if (h2 >= 0 && h2 < H && w2 >= 0 && w2 < W) {
        const float* a = in1 + static_cast<size_t>(n) * C * HW + static_cast<size_t>(h) * W + w;
        const float* b = in2 + static_cast<size_t>(n) * C * HW + static_cast<size_t>(h2) * W + w2;
        for (int c = 0; c < C; ++c)
            acc = __fmaf_rn(__ldg(a + c * HW), __ldg(b + c * HW), acc);
    }
    out[static_cast<size_t>(n) * per_n + i] = use_div ? acc / scale : acc;
}

}  // namespace vsc

extern "C" int vsc_correlation_f32(const float* in1, const float* in2, float* out, int N, int C, int H, int W,
    int max_displacement, int legacy, vsc_stream_t stream)
{
    using namespace vsc;
    if (!in1 || !in2 || !out || N <= 0 || C <= 0 || H <= 0 || W <= 0 || max_displacement < 0 || N > 65535)
        return VSC_E_INVALID;
    if (static_cast<long long>(H) * W * C > 0x7fffffffLL)
        return VSC_E_INVALID;
    if (!aligned4(in1) || !aligned4(in2) || !aligned4(out))
        return VSC_E_ALIGN;
    if (max_displacement == kMD) {
        const int vec = (W % 4 == 0) && aligned16(out);
        const dim3 grid(cdiv(W, kTW), cdiv(H, kTH), N);
        if (grid.y > 65535)
            return VSC_E_INVALID;
        correlation_md4_kernel<<<grid, kCorrThreads, 0, as_stream(stream)>>>(in1, in2, out, C, H, W,
            static_cast<float>(C), legacy ? 1 : 0, vec);
        count_launch();
        return launch_status();
    }
    const int P = 2 * max_displacement + 1;
    const size_t per_n = static_cast<size_t>(P) * P * H * W;
    const dim3 grid(cdiv(static_cast<long long>(per_n), 256), N);
    correlation_generic_kernel<<<grid, 256, 0, as_stream(stream)>>>(in1, in2, out, C, H, W, max_displacement,
        static_cast<float>(C), legacy ? 1 : 0);
    count_launch();
    return launch_status();
}
