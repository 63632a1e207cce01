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

// One thread per VALUE (pixel, channel) of a row, not per pixel: the image is interleaved HWC, so consecutive
// threads touch consecutive floats -- every gather tap, every input read and every output write of a warp is
// one contiguous 128-byte request when the flow is locally smooth.  (The first version used one thread per
// pixel: stride-3 scalar accesses, 25 sectors per request, LSU-bound at 2.6 TB/s -- profiles/r1_notes.md.)
// The three threads of a pixel recompute its warp geometry (a dozen ALU ops; the flow loads are broadcasts).
__global__ void __launch_bounds__(256) stage_a_kernel(const float* __restrict__ origPrev,
    const float* __restrict__ origCur, const float* __restrict__ origNext, const float* __restrict__ procPrev,
    const float* __restrict__ procCur, const float* __restrict__ procNext, const float* __restrict__ lastStab,
    const float* __restrict__ flowFwd, const float* __restrict__ flowBwd, int flowC, float alpha, float beta,
    float gamma, float* __restrict__ adapCmbIn, float* __restrict__ adapCmbPr, float* __restrict__ consWt, int W,
    int H)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;  // float index inside the row
    const int iy = blockIdx.y;
    if (i >= 3 * W)
        return;
    const int ix = i / 3;
    const int c = i - 3 * ix;
    const size_t p = static_cast<size_t>(iy) * W + ix;
    const size_t v = p * 3 + c;
    const WarpGeom gb = hwc_warp_geom(ix, iy, __ldg(flowBwd + p * flowC), __ldg(flowBwd + p * flowC + 1), W, H);
    const WarpGeom gf = hwc_warp_geom(ix, iy, __ldg(flowFwd + p * flowC), __ldg(flowFwd + p * flowC + 1), W, H);
    const float pi = hwc_warp_sample1(origPrev, W, gb, c);   // prevWarpIn   (:182)
    const float pp = hwc_warp_sample1(procPrev, W, gb, c);   // prevWarpPr   (:183)
    const float ni = hwc_warp_sample1(origNext, W, gf, c);   // nextWarpIn   (:186)
    const float np = hwc_warp_sample1(procNext, W, gf, c);   // nextWarpPr   (:187)
    const float ls = hwc_warp_sample1(lastStab, W, gb, c);   // lastStabWarp (:190)
    const float ci = ldg_stream(origCur + v);
    const float cp = ldg_stream(procCur + v);
    float ai, ap;
    adap_comb_value(ci, cp, pi, pp, ni, np, ls, alpha, ai, ap);
    if (adapCmbIn)
        adapCmbIn[v] = ai;
    adapCmbPr[v] = ap;
    consWt[v] = consist_wt_value(ci, ai, beta, gamma);
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
    const dim3 grid(cdiv(3LL * W, 256), H);
    stage_a_kernel<<<grid, 256, 0, as_stream(stream)>>>(origPrev, origCur, origNext, procPrev, procCur, procNext,
        lastStab, flowFwd, flowBwd, flow_channels, alpha, beta, gamma, adapCmbIn, adapCmbPr, consWt, W, H);
    count_launch();
    return launch_status();
}
