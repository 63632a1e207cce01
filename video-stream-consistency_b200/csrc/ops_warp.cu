// custom::Warp for sm_100a -- replaces WarpKernel::ComputeCUDA / warp_forward_kernel
// (reference: src/ort_custom_ops/src/opticalflow/warp_cuda.cc:28-77, warp_cuda.cu:29-98; CPU
// definition warp.cc:71-134).
//
// The reference launches one 32-thread block per 32 pixels PER CHANNEL, so the flow is re-read
// and the sampling geometry + validity mask recomputed C times, with fp64 multiplies on every
// tap.  Here a thread owns one pixel: it reads the flow once, derives the four tap offsets /
// weights once (the mask decision in the reference's own float/double sequence, so the
// `mask > 0.999` threshold falls identically), and then streams over a chunk of channels doing
// only gathers + 4 fp32 multiply-adds per value.  HBM traffic = the algorithmic
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

// grid: x = 256-pixel groups of one image, y = channel chunk, z = n.  One pixel per thread: the 32 lanes of a
// warp sample 32 neighbouring positions, so each of the four gathers touches a handful of 32-byte sectors and
// the store is one full 128-byte line (a 4-pixels-per-thread variant with 128-bit stores was measured 4x
// worse: its gathers stride 16 bytes between lanes, 29 sectors per request -- profiles/r1_notes.md).
// The channel loop is unrolled 4x: 16 independent gathers in flight per thread.
__global__ void __launch_bounds__(256) warp_nchw_kernel(const float* __restrict__ in, const float* __restrict__ flow,
    float* __restrict__ out, int C, int H, int W, int chunk)
{
    const int HW = H * W;
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= HW)
        return;
    const int n = blockIdx.z;
    const int c0 = blockIdx.y * chunk;
    const int c1 = min(C, c0 + chunk);
    const int y = p / W;
    const int x = p - y * W;
    const float* fl = flow + static_cast<size_t>(n) * 2 * HW + p;
    const WarpTap t = warp_setup(x, y, ldg_stream(fl), ldg_stream(fl + HW), W, H);

    const float* ip = in + (static_cast<size_t>(n) * C + c0) * HW;
    float* op = out + (static_cast<size_t>(n) * C + c0) * HW + p;
    int c = c0;
    for (; c + 4 <= c1; c += 4, ip += 4 * static_cast<size_t>(HW), op += 4 * static_cast<size_t>(HW)) {
        const float v0 = warp_sample(ip, t);
        const float v1 = warp_sample(ip + HW, t);
        const float v2 = warp_sample(ip + 2 * static_cast<size_t>(HW), t);
        const float v3 = warp_sample(ip + 3 * static_cast<size_t>(HW), t);
        __stcs(op, v0);
        __stcs(op + HW, v1);
        __stcs(op + 2 * static_cast<size_t>(HW), v2);
        __stcs(op + 3 * static_cast<size_t>(HW), v3);
    }
    for (; c < c1; ++c, ip += HW, op += HW)
        __stcs(op, warp_sample(ip, t));
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
    const unsigned gx = cdiv(static_cast<long long>(H) * W, 256);
    // enough blocks for >= 4 waves of 148 SMs x 8 resident CTAs when the tensor allows, chunks of >= 8 channels
    const long long want = 4LL * sm_count() * 8;
    int nchunk = static_cast<int>((want + static_cast<long long>(gx) * N - 1) / (static_cast<long long>(gx) * N));
    if (nchunk < 1) nchunk = 1;
    if (nchunk > (C + 7) / 8) nchunk = (C + 7) / 8;
    if (nchunk > 65535) nchunk = 65535;
    int chunk = (C + nchunk - 1) / nchunk;
    chunk = (chunk + 3) / 4 * 4;  // whole unrolled groups
    nchunk = (C + chunk - 1) / chunk;
    const dim3 grid(gx, nchunk, N);
    warp_nchw_kernel<<<grid, 256, 0, as_stream(stream)>>>(in, flow, out, C, H, W, chunk);
    count_launch();
    return launch_status();
}
