// Per-pixel device functions of the stabilization path, shared by the stand-alone kernels
// (stab_kernels.cu) and the fused stage-A kernel (stab_fused.cu) so that both produce
// bit-identical values.  Expressions are written in the reference's evaluation order
// (flowconsistency.cu:77-191); the library is built with -fmad=false, so they are evaluated
// exactly as written.
#pragma once
#include "vsc_common.cuh"

namespace vsc {

struct WarpGeom {
    int ix, iy;    // top-left tap
    float fx, fy;  // fractions
};

// kernel_warp geometry, flowconsistency.cu:88-102: clamp to [0, W-3] x [0, H-3], no validity mask
__device__ __forceinline__ WarpGeom hwc_warp_geom(int ix, int iy, float flo_x, float flo_y, int W, int H)
{
    WarpGeom g;
    const float map_fx = fmaxf(0.0f, fminf(static_cast<float>(ix) + flo_x, static_cast<float>(W - 3)));
    const float map_fy = fmaxf(0.0f, fminf(static_cast<float>(iy) + flo_y, static_cast<float>(H - 3)));
    g.ix = static_cast<int>(floorf(map_fx));
    g.iy = static_cast<int>(floorf(map_fy));
    g.fx = map_fx - static_cast<float>(g.ix);
    g.fy = map_fy - static_cast<float>(g.iy);
    return g;
}

// bilinear sample of the 3 channels of an HWC float3 image, flowconsistency.cu:104-113
__device__ __forceinline__ void hwc_warp_sample3(const float* __restrict__ in, int W, const WarpGeom& g, float o[3])
{
    const float* p0 = in + (static_cast<size_t>(g.iy) * W + g.ix) * 3;  // (ix,iy) and (ix+1,iy): 6 contiguous floats
    const float* p1 = p0 + static_cast<size_t>(W) * 3;                  // next row
    const float ofx = 1.0f - g.fx;
    const float ofy = 1.0f - g.fy;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float tmp_1 = __ldg(p0 + c) * ofx + __ldg(p0 + 3 + c) * g.fx;
        const float tmp_2 = __ldg(p1 + c) * ofx + __ldg(p1 + 3 + c) * g.fx;
        o[c] = tmp_1 * ofy + tmp_2 * g.fy;
    }
}

// one channel of the same sample (value-per-thread kernels): identical expression, identical result
__device__ __forceinline__ float hwc_warp_sample1(const float* __restrict__ in, int W, const WarpGeom& g, int c)
{
    const float* p0 = in + (static_cast<size_t>(g.iy) * W + g.ix) * 3 + c;
    const float* p1 = p0 + static_cast<size_t>(W) * 3;
    const float ofx = 1.0f - g.fx;
    const float ofy = 1.0f - g.fy;
    const float tmp_1 = __ldg(p0) * ofx + __ldg(p0 + 3) * g.fx;
    const float tmp_2 = __ldg(p1) * ofx + __ldg(p1 + 3) * g.fx;
    return tmp_1 * ofy + tmp_2 * g.fy;
}

// the same sample split into its four loads and its arithmetic, so that a kernel can issue the loads of several
// images back to back (memory-level parallelism) before combining them: identical expression, identical result
// (32-bit element offsets -- the launchers guarantee images of fewer than 2^31 floats -- so that an address is one
// IMAD.WIDE plus immediates instead of a 64-bit multiply-add chain: the fused kernel is instruction-bound, and
// address arithmetic was as many instructions as its floating-point work)
__device__ __forceinline__ void hwc_warp_taps(const float* __restrict__ in, int W, const WarpGeom& g, int c, float t[4])
{
    const int o0 = (g.iy * W + g.ix) * 3 + c;
    const int o1 = o0 + 3 * W;
    t[0] = __ldg(in + o0);
    t[1] = __ldg(in + o0 + 3);
    t[2] = __ldg(in + o1);
    t[3] = __ldg(in + o1 + 3);
}
__device__ __forceinline__ float hwc_warp_combine(const float t[4], const WarpGeom& g)
{
    const float ofx = 1.0f - g.fx;
    const float ofy = 1.0f - g.fy;
    const float tmp_1 = t[0] * ofx + t[1] * g.fx;
    const float tmp_2 = t[2] * ofx + t[3] * g.fx;
    return tmp_1 * ofy + tmp_2 * g.fy;
}

// kernel_adap_comb for one value, flowconsistency.cu:131-164
__device__ __forceinline__ void adap_comb_value(float ci, float cp, float pi, float pp, float ni, float np, float ls,
    float alpha, float& adp_in, float& adp_pr)
{
    float wt_prv = expf(-alpha * (ci - pi) * (ci - pi));
    float wt_nxt = expf(-alpha * (ci - ni) * (ci - ni));
    if (wt_prv > 0.45f) wt_prv = 0.45f;
    if (wt_nxt > 0.3f) wt_nxt = 0.3f;
    if (wt_prv < 0.001f) wt_prv = 0.0f;
    if (wt_nxt < 0.001f) wt_nxt = 0.0f;
    adp_in = wt_prv * pi + wt_nxt * ni + (1.0f - (wt_prv + wt_nxt)) * ci;
    adp_pr = wt_prv * pp + wt_nxt * np + (1.0f - (wt_prv + wt_nxt)) * cp;
    adp_pr = wt_prv * ls + (1.0f - wt_prv) * adp_pr;
}

// kernel_consist_wt for one value, flowconsistency.cu:176-189
__device__ __forceinline__ float consist_wt_value(float crnt, float adp, float beta, float gamma)
{
    float wt = gamma * expf(-beta * (crnt - adp) * (crnt - adp));
    if (wt < 0.001f) wt = 0.0f;
    return wt;
}

}  // namespace vsc
