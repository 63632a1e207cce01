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

#ifndef VSC_STAGE_A_MINBLOCKS
#define VSC_STAGE_A_MINBLOCKS 4   // 64 registers: 8 CTAs of 128 threads per SM (3 -> 85 registers, 6 CTAs: measured)
#endif

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


// ---- row-walking variants ---------------------------------------------------------------------------------
// stage_a_prep_kernel exposes TWO memory latencies per value, back to back: the flow of the pixel must arrive
// before the tap addresses exist, and the taps before any arithmetic can start; with one value per thread and
// CTAs that live for a single row nothing else of that thread can be in flight meanwhile (ncu: issue 63 %,
// long_scoreboard 5.9 warps per issue cycle).  Here a thread keeps its column and walks down `rows` image rows:
//   * the flow of the row after the next is already loading while the current row is combined (PIPE >= 1);
//   * PIPE == 2 also issues the 24 loads of row r+1 before the arithmetic of row r (two register sets, A / B);
//   * the processed frame's vertical neighbours stay in registers (row r's value is row r+1's upper neighbour):
//     3 loads of procCur per value instead of 5;
//   * the bottom taps of row r are the top taps of row r+1 of the same thread a few hundred cycles later: L1 hits
//     instead of a second trip to L2.
// Per-value arithmetic: the same device functions in the same order -> bit-identical results.
// Measured (ncu, profiles/r1_stage_a_walk_ncu*.txt; 4K / 1080p): one row per CTA 351 / 90 us (issue 56 %);
// walk + flow prefetch, 8 rows, 128-thread CTAs 254 / 70 us (issue 76 %, L1 hit 51 -> 64 %) = 0.70 / 0.64 of the
// measured HBM peak in algorithmic bytes -- the default.  PIPE == 2 needs 122 registers (2 CTAs per SM): 363 / 101 us.
// Squeezing the prefetch variant into 47 / 40 registers (5 / 6 CTAs per SM) makes ptxas serialise the loads:
// 281 / 270 us at 4K, slower than 64 registers at lower occupancy.
struct StageATaps {
    float pi[4], pp[4], ni[4], np[4], ls[4];
    float ci, n_r, n_l, cp_dn;
    float fxb, fyb, fxf, fyf;
};
struct StageAFlow {
    float bx, by, fx, fy;
};
struct StageAPtrs {
    const float *origPrev, *origCur, *origNext, *procPrev, *procCur, *procNext, *lastStab, *flowFwd, *flowBwd;
};

__device__ __forceinline__ StageAFlow stage_a_flow(const StageAPtrs& P, int pf)
{
    StageAFlow f;
    f.bx = __ldg(P.flowBwd + pf);
    f.by = __ldg(P.flowBwd + pf + 1);
    f.fx = __ldg(P.flowFwd + pf);
    f.fy = __ldg(P.flowFwd + pf + 1);
    return f;
}

// every load of value v of row iy (the row's flow has arrived).  PREP: also the processed frame's neighbours for
// the Laplacian; otherwise cp_dn holds the value's own processed sample.
template <bool PREP = true>
__device__ __forceinline__ void stage_a_issue(StageATaps& t, const StageAPtrs& P, const StageAFlow& f, int ix, int iy,
    int c, int v, int W, int H, bool has_r, bool has_l)
{
    const WarpGeom gb = hwc_warp_geom(ix, iy, f.bx, f.by, W, H);
    const WarpGeom gf = hwc_warp_geom(ix, iy, f.fx, f.fy, W, H);
    hwc_warp_taps(P.origPrev, W, gb, c, t.pi);
    hwc_warp_taps(P.procPrev, W, gb, c, t.pp);
    hwc_warp_taps(P.origNext, W, gf, c, t.ni);
    hwc_warp_taps(P.procNext, W, gf, c, t.np);
    hwc_warp_taps(P.lastStab, W, gb, c, t.ls);
    t.ci = ldg_stream(P.origCur + v);
    if constexpr (PREP) {
        t.n_r = has_r ? __ldg(P.procCur + v + 3) : 0.0f;
        t.n_l = has_l ? __ldg(P.procCur + v - 3) : 0.0f;
        t.cp_dn = (iy + 1 < H) ? __ldg(P.procCur + v + 3 * W) : 0.0f;   // lower neighbour = the next row's own value
    } else {
        t.n_r = t.n_l = 0.0f;
        t.cp_dn = ldg_stream(P.procCur + v);
    }
    t.fxb = gb.fx;
    t.fyb = gb.fy;
    t.fxf = gf.fx;
    t.fyf = gf.fy;
}

// arithmetic + stores of value v of row iy; cp / n_u = procCur at (iy, iy-1), carried in registers
__device__ __forceinline__ void stage_a_finish(const StageATaps& t, float cp, float n_u, int ix, int iy, int c, int v,
    int W, int H, bool has_r, bool has_l, float alpha, float beta, float gamma, float step, float* __restrict__ coefA,
    float* __restrict__ coefB, float* __restrict__ pr1, float* __restrict__ tg1, float* __restrict__ wt1)
{
    WarpGeom gb, gf;
    gb.ix = gb.iy = gf.ix = gf.iy = 0;
    gb.fx = t.fxb;
    gb.fy = t.fyb;
    gf.fx = t.fxf;
    gf.fy = t.fyf;
    const float pi = hwc_warp_combine(t.pi, gb);
    const float pp = hwc_warp_combine(t.pp, gb);
    const float ni = hwc_warp_combine(t.ni, gf);
    const float np = hwc_warp_combine(t.np, gf);
    const float ls = hwc_warp_combine(t.ls, gb);
    float ai, tgt;
    adap_comb_value(t.ci, cp, pi, pp, ni, np, ls, alpha, ai, tgt);
    const float w = consist_wt_value(t.ci, ai, beta, gamma);
    const bool has_d = (iy + 1) < (H - 1), has_u = (iy - 1) >= 0;
    int cnt = 0;
    float lap = 0.0f;
    if (has_r) { lap += t.n_r; cnt += 1; }
    if (has_l) { lap += t.n_l; cnt += 1; }
    if (has_d) { lap += t.cp_dn; cnt += 1; }
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

template <int PIPE>
__global__ void __launch_bounds__(256, PIPE == 2 ? 2 : VSC_STAGE_A_MINBLOCKS) stage_a_prep_rows_kernel(StageAPtrs P, int flowC,
    float alpha, float beta, float gamma, float step, float* __restrict__ coefA, float* __restrict__ coefB,
    float* __restrict__ pr1, float* __restrict__ tg1, float* __restrict__ wt1, int W, int H, int rows)
{
    pdl_enter();
    const int L = 3 * W;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;  // float index inside the row
    if (i >= L)
        return;
    const int ix = i / 3;
    const int c = i - 3 * ix;
    const int y0 = blockIdx.y * rows;
    const int y1 = min(H, y0 + rows);
    const bool has_r = (ix + 1) < (W - 1), has_l = (ix - 1) >= 0;
    const int fstep = W * flowC;
    int v = (y0 * W + ix) * 3 + c;      // 32-bit element offsets: the launcher checks 3*W*H < 2^31
    int pf = (y0 * W + ix) * flowC;
    float cp = __ldg(P.procCur + v);
    float n_u = y0 > 0 ? __ldg(P.procCur + v - L) : 0.0f;
    if constexpr (PIPE < 2) {
        StageAFlow f = stage_a_flow(P, pf);
        for (int iy = y0; iy < y1; ++iy, v += L, pf += fstep) {
            StageATaps t;
            stage_a_issue(t, P, f, ix, iy, c, v, W, H, has_r, has_l);
            if (PIPE == 1 && iy + 1 < y1)
                f = stage_a_flow(P, pf + fstep);
            stage_a_finish(t, cp, n_u, ix, iy, c, v, W, H, has_r, has_l, alpha, beta, gamma, step, coefA, coefB, pr1,
                tg1, wt1);
            n_u = cp;
            cp = t.cp_dn;
            if (PIPE == 0 && iy + 1 < y1)
                f = stage_a_flow(P, pf + fstep);
        }
    } else {
        // two register sets: A holds even rows of the chunk, B odd rows; F0 / F1 are their flows, fetched two rows
        // ahead of use
        StageATaps A, B;
        StageAFlow F0 = stage_a_flow(P, pf), F1 = F0;
        if (y0 + 1 < y1)
            F1 = stage_a_flow(P, pf + fstep);
        stage_a_issue(A, P, F0, ix, y0, c, v, W, H, has_r, has_l);
        if (y0 + 2 < y1)
            F0 = stage_a_flow(P, pf + 2 * fstep);
        for (int iy = y0; iy < y1; iy += 2, v += 2 * L, pf += 2 * fstep) {
            const bool second = iy + 1 < y1;
            if (second) {
                stage_a_issue(B, P, F1, ix, iy + 1, c, v + L, W, H, has_r, has_l);
                if (iy + 3 < y1)
                    F1 = stage_a_flow(P, pf + 3 * fstep);
            }
            stage_a_finish(A, cp, n_u, ix, iy, c, v, W, H, has_r, has_l, alpha, beta, gamma, step, coefA, coefB, pr1, tg1,
                wt1);
            n_u = cp;
            cp = A.cp_dn;
            if (second) {
                if (iy + 2 < y1) {
                    stage_a_issue(A, P, F0, ix, iy + 2, c, v + 2 * L, W, H, has_r, has_l);
                    if (iy + 4 < y1)
                        F0 = stage_a_flow(P, pf + 4 * fstep);
                }
                stage_a_finish(B, cp, n_u, ix, iy + 1, c, v + L, W, H, has_r, has_l, alpha, beta, gamma, step, coefA,
                    coefB, pr1, tg1, wt1);
                n_u = cp;
                cp = B.cp_dn;
            }
        }
    }
}

// vsc_stage_a_fused (adapCmbPr / consWt written out, no solver set-up) as a row walk with the flow prefetched
__global__ void __launch_bounds__(256, VSC_STAGE_A_MINBLOCKS) stage_a_rows_kernel(StageAPtrs P, int flowC, float alpha, float beta,
    float gamma, float* __restrict__ adapCmbIn, float* __restrict__ adapCmbPr, float* __restrict__ consWt, int W, int H,
    int rows)
{
    const int L = 3 * W;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= L)
        return;
    const int ix = i / 3;
    const int c = i - 3 * ix;
    const int y0 = blockIdx.y * rows;
    const int y1 = min(H, y0 + rows);
    const int fstep = W * flowC;
    int v = (y0 * W + ix) * 3 + c;
    int pf = (y0 * W + ix) * flowC;
    StageAFlow f = stage_a_flow(P, pf);
    for (int iy = y0; iy < y1; ++iy, v += L, pf += fstep) {
        StageATaps t;
        stage_a_issue<false>(t, P, f, ix, iy, c, v, W, H, false, false);
        if (iy + 1 < y1)
            f = stage_a_flow(P, pf + fstep);
        WarpGeom gb, gf;
        gb.ix = gb.iy = gf.ix = gf.iy = 0;
        gb.fx = t.fxb;
        gb.fy = t.fyb;
        gf.fx = t.fxf;
        gf.fy = t.fyf;
        const float pi = hwc_warp_combine(t.pi, gb);   // prevWarpIn   (videostabilizer.cpp:182)
        const float pp = hwc_warp_combine(t.pp, gb);   // prevWarpPr   (:183)
        const float ni = hwc_warp_combine(t.ni, gf);   // nextWarpIn   (:186)
        const float np = hwc_warp_combine(t.np, gf);   // nextWarpPr   (:187)
        const float ls = hwc_warp_combine(t.ls, gb);   // lastStabWarp (:190)
        float ai, ap;
        adap_comb_value(t.ci, t.cp_dn, pi, pp, ni, np, ls, alpha, ai, ap);
        if (adapCmbIn)
            adapCmbIn[v] = ai;
        adapCmbPr[v] = ap;
        consWt[v] = consist_wt_value(t.ci, ai, beta, gamma);
    }
}

// vsc_set_stage_a_mode: low 4 bits = kernel (0 default = 3 with 128-thread CTAs, 1 one row per CTA, 2 row walk,
// 3 row walk + flow prefetch, 4 row walk + loads one row ahead); bits 4-7 = log2(rows per CTA), 0 = 8 (fewer on
// small frames), or bits 12-19 = rows per CTA as a number; bit 8 = 128-thread CTAs
std::atomic<int> g_stage_a_mode = 0;
// rows per CTA forced by the mode word: bits 12-19 = the number itself, else bits 4-7 = its log2, else 0 (automatic)
static int stage_a_rows_forced()
{
    const int n = (g_stage_a_mode >> 12) & 0xFF, lg = (g_stage_a_mode >> 4) & 0xF;
    return n ? n : lg ? (1 << lg) : 0;
}

int launch_stage_a_prep(const float* origPrev, const float* origCur, const float* origNext, const float* procPrev,
    const float* procCur, const float* procNext, const float* lastStab, const float* flowFwd, const float* flowBwd,
    int flowC, float alpha, float beta, float gamma, float step, float* coefA, float* coefB, float* pr1, float* tg1,
    float* wt1, int W, int H, cudaStream_t st)
{
    int kind = g_stage_a_mode & 0xF;
    int rc;
    if (kind == 0)
        kind = 3;
    if (kind >= 2) {
        const bool dflt = (g_stage_a_mode & 0xF) == 0;
        const int bs = (dflt || (g_stage_a_mode & 0x100)) ? 128 : 256;
        const int lg = stage_a_rows_forced();
        int rows = lg ? lg : 8;
        // default: 8 rows per CTA unless that leaves fewer than ~4 waves of CTAs (small frames)
        while (!lg && rows > 1 && static_cast<long long>(cdiv(3LL * W, bs)) * cdiv(H, rows) < 4LL * sm_count() * 8)
            rows >>= 1;
        const dim3 grid(cdiv(3LL * W, bs), cdiv(H, rows));
        const StageAPtrs P{origPrev, origCur, origNext, procPrev, procCur, procNext, lastStab, flowFwd, flowBwd};
        auto* k = kind == 2 ? stage_a_prep_rows_kernel<0> : kind == 3 ? stage_a_prep_rows_kernel<1>
                                                                      : stage_a_prep_rows_kernel<2>;
        rc = launch_pdl(k, grid, dim3(bs), 0, st, P, flowC, alpha, beta, gamma, step, coefA, coefB, pr1, tg1, wt1, W,
            H, rows);
    } else {
        const dim3 grid(cdiv(3LL * W, 256), H);
        rc = launch_pdl(stage_a_prep_kernel, grid, dim3(256), 0, st, origPrev, origCur, origNext, procPrev, procCur,
            procNext, lastStab, flowFwd, flowBwd, flowC, alpha, beta, gamma, step, coefA, coefB, pr1, tg1, wt1, W, H);
    }
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
    // the row-walking kernel uses 32-bit element offsets; one row per CTA otherwise or on request (mode 1)
    if ((g_stage_a_mode & 0xF) != 1 && 3LL * W * H < 0x7fffffffLL) {
        const int lg = stage_a_rows_forced();
        int rows = lg ? lg : 8;
        while (!lg && rows > 1 && static_cast<long long>(cdiv(3LL * W, 128)) * cdiv(H, rows) < 4LL * sm_count() * 8)
            rows >>= 1;
        const dim3 grid(cdiv(3LL * W, 128), cdiv(H, rows));
        const StageAPtrs P{origPrev, origCur, origNext, procPrev, procCur, procNext, lastStab, flowFwd, flowBwd};
        stage_a_rows_kernel<<<grid, 128, 0, as_stream(stream)>>>(P, flow_channels, alpha, beta, gamma, adapCmbIn,
            adapCmbPr, consWt, W, H, rows);
    } else {
        const dim3 grid(cdiv(3LL * W, 256), H);
        stage_a_kernel<<<grid, 256, 0, as_stream(stream)>>>(origPrev, origCur, origNext, procPrev, procCur, procNext,
            lastStab, flowFwd, flowBwd, flow_channels, alpha, beta, gamma, adapCmbIn, adapCmbPr, consWt, W, H);
    }
    count_launch();
    return launch_status();
}

extern "C" int vsc_set_stage_a_mode(int mode)
{
    const int kind = mode & 0xF;
    if (mode < 0 || mode > 0xFFFFF || (mode & 0xE00) || kind > 4)
        return VSC_E_INVALID;
    vsc::g_stage_a_mode = mode;
    return VSC_OK;
}
