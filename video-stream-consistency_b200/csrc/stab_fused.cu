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

// Stage A fused with the solver set-up of the fine pyramid level (and the coarse level's inputs):
// instead of writing adapCmbPr / consWt (24 B/px), re-reading them with the processed frame in
// solver_prepare_kernel (36 B/px read, 36 B/px written) and three times more in the pyramid-down resizes,
// each thread folds its value straight into the level-0 solver coefficients
//     A = -step*(cnt + w)      B = step*(w*tgt - lap(pr))          (same expressions as solver_prepare_kernel)
// and, for pixels with even x and y, also stores (pr, tgt, w) into the half-resolution level-1 inputs: with even
// W and H the reference's get_bilinear down-scale (flowconsistency.cu:50-75, no half-pixel offset) samples
// exactly those pixels with weights (1,0,0,0).  Results are bit-identical to the unfused sequence.
__global__ void __launch_bounds__(256, 4) stage_a_prep_kernel(const float* __restrict__ origPrev,
    const float* __restrict__ origCur, const float* __restrict__ origNext, const float* __restrict__ procPrev,
    const float* __restrict__ procCur, const float* __restrict__ procNext, const float* __restrict__ lastStab,
    const float* __restrict__ flowFwd, const float* __restrict__ flowBwd, int flowC, float alpha, float beta,
    float gamma, float step, float* __restrict__ coefA, float* __restrict__ coefB, float* __restrict__ pr1,
    float* __restrict__ tg1, float* __restrict__ wt1, int W, int H)
{
    pdl_enter();
    const int L = 3 * W;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;  // float index inside the row
    const int iy = blockIdx.y;
    if (i >= L)
        return;
    const int ix = i / 3;
    const int c = i - 3 * ix;
    const int p = iy * W + ix;      // 32-bit element offsets: the launcher checks 3*W*H < 2^31
    const int v = p * 3 + c;
    const int pf = p * flowC;
    const WarpGeom gb = hwc_warp_geom(ix, iy, __ldg(flowBwd + pf), __ldg(flowBwd + pf + 1), W, H);
    const WarpGeom gf = hwc_warp_geom(ix, iy, __ldg(flowFwd + pf), __ldg(flowFwd + pf + 1), W, H);
    // all 26 loads of the value are issued before the first one is consumed (64 registers, 4 CTAs per SM): the
    // kernel is bound by gather latency, not by occupancy
    float t_pi[4], t_pp[4], t_ni[4], t_np[4], t_ls[4];
    hwc_warp_taps(origPrev, W, gb, c, t_pi);
    hwc_warp_taps(procPrev, W, gb, c, t_pp);
    hwc_warp_taps(origNext, W, gf, c, t_ni);
    hwc_warp_taps(procNext, W, gf, c, t_np);
    hwc_warp_taps(lastStab, W, gb, c, t_ls);
    const float ci = ldg_stream(origCur + v);
    const float cp = __ldg(procCur + v);
    const bool has_r = (ix + 1) < (W - 1), has_l = (ix - 1) >= 0, has_d = (iy + 1) < (H - 1), has_u = (iy - 1) >= 0;
    const float n_r = has_r ? __ldg(procCur + v + 3) : 0.0f;
    const float n_l = has_l ? __ldg(procCur + v - 3) : 0.0f;
    const float n_d = has_d ? __ldg(procCur + v + L) : 0.0f;
    const float n_u = has_u ? __ldg(procCur + v - L) : 0.0f;
    const float pi = hwc_warp_combine(t_pi, gb);
    const float pp = hwc_warp_combine(t_pp, gb);
    const float ni = hwc_warp_combine(t_ni, gf);
    const float np = hwc_warp_combine(t_np, gf);
    const float ls = hwc_warp_combine(t_ls, gb);
    float ai, tgt;
    adap_comb_value(ci, cp, pi, pp, ni, np, ls, alpha, ai, tgt);
    const float w = consist_wt_value(ci, ai, beta, gamma);
    // Laplacian of the processed frame with the reference's inclusion tests (flowconsistency.cu:215-238)
    int cnt = 0;
    float lap = 0.0f;
    if (has_r) { lap += n_r; cnt += 1; }
    if (has_l) { lap += n_l; cnt += 1; }
    if (has_d) { lap += n_d; cnt += 1; }
    if (has_u) { lap += n_u; cnt += 1; }
    lap -= static_cast<float>(cnt) * cp;
    coefA[v] = -step * (static_cast<float>(cnt) + w);
    coefB[v] = step * (w * tgt - lap);
    if (((ix | iy) & 1) == 0) {
        const int v1 = ((iy >> 1) * (W >> 1) + (ix >> 1)) * 3 + c;
        pr1[v1] = cp;
        tg1[v1] = tgt;
        wt1[v1] = w;
    }
}

int launch_stage_a_prep(const float* origPrev, const float* origCur, const float* origNext, const float* procPrev,
    const float* procCur, const float* procNext, const float* lastStab, const float* flowFwd, const float* flowBwd,
    int flowC, float alpha, float beta, float gamma, float step, float* coefA, float* coefB, float* pr1, float* tg1,
    float* wt1, int W, int H, cudaStream_t st)
{
    const dim3 grid(cdiv(3LL * W, 256), H);
    const int rc = launch_pdl(stage_a_prep_kernel, grid, dim3(256), 0, st, origPrev, origCur, origNext, procPrev, procCur,
        procNext, lastStab, flowFwd, flowBwd, flowC, alpha, beta, gamma, step, coefA, coefB, pr1, tg1, wt1, W, H);
    count_launch();
    return rc ? rc : launch_status();
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
