// Fused "stage A" of VideoStabilizer::doOneStep for sm_100a (reference:
// src/stabilization/videostabilizer.cpp:182-198 calling flowconsistency.cu:77-191).
//
// The reference runs 7 kernels here (5x kernel_warp, kernel_adap_comb, kernel_consist_wt), each
// followed by cudaDeviceSynchronize, and round-trips six full-resolution intermediates through
// HBM (prevWarpIn/Pr, nextWarpIn/Pr, lastStabWarp, adapCmbIn): 180 + 108 + 36 = 324 B/pixel.
// This kernel does the whole stage in one pass: a thread reads the two flow vectors of its
// pixel, gathers the five bilinear samples straight from the source frames, evaluates the
// adaptive combination and the consistency weight in registers and writes only adapCmbPr and
// consWt (adapCmbIn on request): 7 images + 2 flows in, 2 images out = 132 B/pixel.  The
// per-value arithmetic is the same device code as the stand-alone kernels (stab_device.cuh),
// so the results are bit-identical to calling those one by one.
#include "stab_device.cuh"

namespace vsc {

__global__ void __launch_bounds__(256) stage_a_kernel(const float* __restrict__ origPrev,
    const float* __restrict__ origCur, const float* __restrict__ origNext, const float* __restrict__ procPrev,
    const float* __restrict__ procCur, const float* __restrict__ procNext, const float* __restrict__ lastStab,
    const float* __restrict__ flowFwd, const float* __restrict__ flowBwd, int flowC, float alpha, float beta,
    float gamma, float* __restrict__ adapCmbIn, float* __restrict__ adapCmbPr, float* __restrict__ consWt, int W,
    int H)
{
    const int ix = blockIdx.x * blockDim.x + threadIdx.x;
    const int iy = blockIdx.y;
    if (ix >= W)
        return;
    const size_t p = static_cast<size_t>(iy) * W + ix;
    const WarpGeom gb = hwc_warp_geom(ix, iy, ldg_stream(flowBwd + p * flowC), ldg_stream(flowBwd + p * flowC + 1), W, H);
    const WarpGeom gf = hwc_warp_geom(ix, iy, ldg_stream(flowFwd + p * flowC), ldg_stream(flowFwd + p * flowC + 1), W, H);
    float pi[3], pp[3], ls[3], ni[3], np[3];
    hwc_warp_sample3(origPrev, W, gb, pi);   // prevWarpIn   (:182)
    hwc_warp_sample3(procPrev, W, gb, pp);   // prevWarpPr   (:183)
    hwc_warp_sample3(origNext, W, gf, ni);   // nextWarpIn   (:186)
    hwc_warp_sample3(procNext, W, gf, np);   // nextWarpPr   (:187)
    hwc_warp_sample3(lastStab, W, gb, ls);   // lastStabWarp (:190)
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float ci = ldg_stream(origCur + p * 3 + c);
        const float cp = ldg_stream(procCur + p * 3 + c);
        float ai, ap;
        adap_comb_value(ci, cp, pi[c], pp[c], ni[c], np[c], ls[c], alpha, ai, ap);
        if (adapCmbIn)
            adapCmbIn[p * 3 + c] = ai;
        adapCmbPr[p * 3 + c] = ap;
        consWt[p * 3 + c] = consist_wt_value(ci, ai, beta, gamma);
    }
}

}  // namespace vsc

extern "C" int vsc_stage_a_fused(const float* origPrev, const float* origCur, const float* origNext,
    const float* procPrev, const float* procCur, const float* procNext, const float* lastStab, const float* flowFwd,
    const float* flowBwd, int flow_channels, float alpha, float beta, float gamma, float* adapCmbIn, float* adapCmbPr,
    float* consWt, int W, int H, vsc_stream_t stream)
{
    using namespace vsc;
    if (!origPrev || !origCur || !origNext || !procPrev || !procCur || !procNext || !lastStab || !flowFwd || !flowBwd
        || !adapCmbPr || !consWt || W < 2 || H < 2 || H > 65535 || (flow_channels != 2 && flow_channels != 3))
        return VSC_E_INVALID;
    const dim3 grid(cdiv(W, 256), H);
    stage_a_kernel<<<grid, 256, 0, as_stream(stream)>>>(origPrev, origCur, origNext, procPrev, procCur, procNext,
        lastStab, flowFwd, flowBwd, flow_channels, alpha, beta, gamma, adapCmbIn, adapCmbPr, consWt, W, H);
    count_launch();
    return launch_status();
}
