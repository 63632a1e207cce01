// Sampling geometry of custom::Warp shared by the gather kernels (ops_warp.cu) and the TMA-staged kernel
// (ops_warp_staged.cu): one definition of the corners, weights and the validity mask, so every kernel decides
// `mask > 0.999` identically and accumulates in the same order (reference: warp.cc:85-129, warp_cuda.cu:42-81).
#pragma once
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
    // (the reference adds each double-precision product to the float value with ONE rounding, warp.cc:119-126: a
    // fused multiply-add is the closer restatement, and 3 instructions per value cheaper than mul + add)
    float v = t.w00 * a;
    v = __fmaf_rn(t.w10, b, v);
    v = __fmaf_rn(t.w01, c, v);
    v = __fmaf_rn(t.w11, d, v);
    return v;
}

}  // namespace vsc
