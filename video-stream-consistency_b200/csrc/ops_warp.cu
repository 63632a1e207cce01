// custom::Warp for sm_100a -- replaces WarpKernel::ComputeCUDA / warp_forward_kernel
// (reference: src/ort_custom_ops/src/opticalflow/warp_cuda.cc:28-77, warp_cuda.cu:29-98; CPU
// definition warp.cc:71-134).
//
// The reference launches one 32-thread block per 32 pixels PER CHANNEL, so the flow is re-read
// and the sampling geometry + validity mask recomputed C times, with fp64 multiplies on every
// tap.  Here a thread owns VEC consecutive pixels of one image row: it reads the flow once,
// derives the four tap offsets / weights once (the mask decision in the reference's own
// float/double sequence, so the `mask > 0.999` threshold falls identically), and then streams
// over a chunk of channels doing only gathers + 4 fp32 multiply-adds per value, with 128-bit
// flow loads and output stores when the row length allows.  HBM traffic = the algorithmic
// 4*N*H*W*(2C+2) bytes (gathers hit L1/L2: neighbouring pixels sample neighbouring taps).
#include "vsc_common.cuh"

namespace vsc {

struct WarpTap {
    int o00, o10, o01, o11;      // element offsets inside one H*W plane (0 when the corner is unused)
    float w00, w10, w01, w11;    // bilinear weights (0 when unused)
    unsigned valid;              // bit k set: corner k is read
};

// geometry of one output pixel, following warp.cc:85-129 / warp_cuda.cu:42-81 type by type
__device__ __forceinline__ WarpTap warp_setup(int x, int y, float fu, float fv, int W, int H)
{
    WarpTap t;
    const float xf = static_cast<float>(x) + fu;
    const float yf = static_cast<float>(y) + fv;
    const float xL = floorf(xf);
    const float yT = floorf(yf);
    const float alpha = xf - xL;
    const float beta = yf - yT;
    const float right_edge = static_cast<float>(W - 1);
    const float bottom_edge = static_cast<float>(H - 1);
    const float xR = xL + 1.0f;  // == float(double(xL) + 1.0): one rounding of the exact sum
    const float yB = yT + 1.0f;
    const bool mL = (0.0f <= xL && xL <= right_edge);
    const bool mR = (0.0f <= xR && xR <= right_edge);
    const bool mT = (0.0f <= yT && yT <= bottom_edge);
    const bool mB = (0.0f <= yB && yB <= bottom_edge);
    // products in double, each += rounded back to float (the reference's mixed types)
    const double a1 = 1.0 - static_cast<double>(alpha);
    const double b1 = 1.0 - static_cast<double>(beta);
    const double d00 = a1 * b1;
    const double d10 = static_cast<double>(alpha) * b1;
    const double d01 = a1 * static_cast<double>(beta);
    const double d11 = static_cast<double>(alpha * beta);  // float*float first, as `(alpha) * (beta)` evaluates
    float mask = 0.0f;
    mask = static_cast<float>(static_cast<double>(mask) + d00 * ((mT && mL) ? 1.0 : 0.0));
    mask = static_cast<float>(static_cast<double>(mask) + d10 * ((mT && mR) ? 1.0 : 0.0));
    mask = static_cast<float>(static_cast<double>(mask) + d01 * ((mB && mL) ? 1.0 : 0.0));
    mask = static_cast<float>(static_cast<double>(mask) + d11 * ((mB && mR) ? 1.0 : 0.0));
    const bool keep = static_cast<double>(mask) > 0.999;
    const bool v00 = keep && mT && mL, v10 = keep && mT && mR, v01 = keep && mB && mL, v11 = keep && mB && mR;
    // integer coordinates are only formed for corners that passed the range test
    const int ixL = mL ? static_cast<int>(xL) : 0, ixR = mR ? static_cast<int>(xR) : 0;
    const int iyT = mT ? static_cast<int>(yT) : 0, iyB = mB ? static_cast<int>(yB) : 0;
    t.o00 = v00 ? iyT * W + ixL : 0;
    t.o10 = v10 ? iyT * W + ixR : 0;
    t.o01 = v01 ? iyB * W + ixL : 0;
    t.o11 = v11 ? iyB * W + ixR : 0;
    t.w00 = v00 ? static_cast<float>(d00) : 0.0f;
    t.w10 = v10 ? static_cast<float>(d10) : 0.0f;
    t.w01 = v01 ? static_cast<float>(d01) : 0.0f;
    t.w11 = v11 ? static_cast<float>(d11) : 0.0f;
    t.valid = (v00 ? 1u : 0u) | (v10 ? 2u : 0u) | (v01 ? 4u : 0u) | (v11 ? 8u : 0u);
    return t;
}

__device__ __forceinline__ float warp_sample(const float* __restrict__ plane, const WarpTap& t)
{
    // unused corners are not read (their storage may hold non-finite values)
    const float a = (t.valid & 1u) ? __ldg(plane + t.o00) : 0.0f;
    const float b = (t.valid & 2u) ? __ldg(plane + t.o10) : 0.0f;
    const float c = (t.valid & 4u) ? __ldg(plane + t.o01) : 0.0f;
    const float d = (t.valid & 8u) ? __ldg(plane + t.o11) : 0.0f;
    float v = t.w00 * a;
    v = v + t.w10 * b;
    v = v + t.w01 * c;
    v = v + t.w11 * d;
    return v;
}

// grid: x = pixel groups of one image, y = channel chunk, z = n
template <int VEC>
__global__ void __launch_bounds__(256) warp_nchw_kernel(const float* __restrict__ in, const float* __restrict__ flow,
    float* __restrict__ out, int C, int H, int W, int chunk)
{
    const int HW = H * W;
    const int groups = HW / VEC;  // VEC==4 only when W % 4 == 0
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= groups)
        return;
    const int n = blockIdx.z;
    const int c0 = blockIdx.y * chunk;
    const int c1 = min(C, c0 + chunk);
    const int p = g * VEC;
    const int y = p / W;
    const int x = p - y * W;

    const float* fu = flow + (static_cast<size_t>(n) * 2 + 0) * HW + p;
    const float* fv = flow + (static_cast<size_t>(n) * 2 + 1) * HW + p;
    WarpTap t[VEC];
    if constexpr (VEC == 4) {
        const float4 u4 = ldg_stream4(fu);
        const float4 v4 = ldg_stream4(fv);
        t[0] = warp_setup(x + 0, y, u4.x, v4.x, W, H);
        t[1] = warp_setup(x + 1, y, u4.y, v4.y, W, H);
        t[2] = warp_setup(x + 2, y, u4.z, v4.z, W, H);
        t[3] = warp_setup(x + 3, y, u4.w, v4.w, W, H);
    } else {
        t[0] = warp_setup(x, y, ldg_stream(fu), ldg_stream(fv), W, H);
    }

    const float* ip = in + (static_cast<size_t>(n) * C + c0) * HW;
    float* op = out + (static_cast<size_t>(n) * C + c0) * HW + p;
#pragma unroll 2
    for (int c = c0; c < c1; ++c, ip += HW, op += HW) {
        if constexpr (VEC == 4) {
            float4 r;
            r.x = warp_sample(ip, t[0]);
            r.y = warp_sample(ip, t[1]);
            r.z = warp_sample(ip, t[2]);
            r.w = warp_sample(ip, t[3]);
            stg_stream4(op, r);
        } else {
            *op = warp_sample(ip, t[0]);
        }
    }
}

}  // namespace vsc

extern "C" int vsc_warp_nchw_f32(const float* in, const float* flow, float* out, int N, int C, int H, int W,
    vsc_stream_t stream)
{
    using namespace vsc;
    if (!in || !flow || !out || N <= 0 || C <= 0 || H <= 0 || W <= 0)
        return VSC_E_INVALID;
    if (static_cast<long long>(H) * W > 0x7fffffffLL / 2 || N > 65535)
        return VSC_E_INVALID;
    if (!aligned4(in) || !aligned4(flow) || !aligned4(out))
        return VSC_E_ALIGN;
    const bool vec = (W % 4 == 0) && aligned16(flow) && aligned16(out);
    const int VEC = vec ? 4 : 1;
    const int groups = H * W / VEC;
    const unsigned gx = cdiv(groups, 256);
    // enough blocks for >= 4 waves of 148 SMs x 8 resident CTAs when the tensor allows, chunks of >= 4 channels
    const long long want = 4LL * sm_count() * 8;
    int nchunk = static_cast<int>((want + static_cast<long long>(gx) * N - 1) / (static_cast<long long>(gx) * N));
    if (nchunk < 1) nchunk = 1;
    if (nchunk > (C + 3) / 4) nchunk = (C + 3) / 4;
    if (nchunk > 65535) nchunk = 65535;
    const int chunk = (C + nchunk - 1) / nchunk;
    nchunk = (C + chunk - 1) / chunk;
    const dim3 grid(gx, nchunk, N);
    if (vec)
        warp_nchw_kernel<4><<<grid, 256, 0, as_stream(stream)>>>(in, flow, out, C, H, W, chunk);
    else
        warp_nchw_kernel<1><<<grid, 256, 0, as_stream(stream)>>>(in, flow, out, C, H, W, chunk);
    count_launch();
    return launch_status();
}
