// The screened-Poisson consistency solver for sm_100a -- replaces get_consist_out /
// kernel_consist_out and the pyramid loop of doOneStep (reference:
// src/stabilization/flowconsistency.cu:193-258,350-374; videostabilizer.cpp:200-228).
//
// Reference: per sweep and per value it reads crntPr (5 taps), consisOut (5 taps, IN PLACE -- the
// buffer being written, a data race, :370), consWt, prevStabWarp and prevUpdt and writes 2 arrays
// (84 B/pixel/sweep), 225 dependent launches per frame, two cudaMallocs per call.
//
// Here:
//   * deterministic Jacobi ordering: sweep k reads only sweep k-1 (two ping-pong buffers);
//   * everything that does not change between sweeps is folded once per solve into two
//     coefficient images (solver_prepare_kernel):
//         cnt  = number of neighbours the reference includes (asymmetric tests, :215-236)
//         Lp   = sum_nb pr - cnt*pr                      (Laplacian of the processed frame)
//         A    = -step * (cnt + w)
//         B    =  step * (w*tgt - Lp)
//     so that one sweep is, with S = sum of the included neighbours of `out`:
//         u'   = step*S + (A*out + B)                    ( = step * grad_val of :245 )
//         out' = (out + u') + mom*u                      ( :249 ; u = 0 before the first sweep == isMom 0 )
//     i.e. 2 state images (out, u) + 2 coefficient images = 72 B/pixel/sweep unblocked, 7 flops;
//   * 128-bit accesses over the flattened row of 3W floats (x-neighbours are +-3 floats away).
// The sweep arithmetic uses explicit FMAs; it is mathematically the reference's update and
// differs from it only in fp32 rounding (parity: tests/test_stab_gpu.py -- test_solver_vs_jacobi_oracle, test_sequence_vs_reference_gpu_fixtures; tolerances stated there).
#include "vsc_common.cuh"

#ifndef VSC_SOLVER_QG_DEFAULT
#define VSC_SOLVER_QG_DEFAULT 1
#endif
#ifndef VSC_SOLVER_PERMUTE_DEFAULT
#define VSC_SOLVER_PERMUTE_DEFAULT 0
#endif

namespace vsc {

constexpr size_t kAlign = 256;
inline size_t align_up(size_t b) { return (b + kAlign - 1) / kAlign * kAlign; }

// one thread per value; also zeroes u and (optionally) copies out -> alt
__global__ void __launch_bounds__(256) solver_prepare_kernel(const float* __restrict__ pr,
    const float* __restrict__ tgt, const float* __restrict__ wt, float* __restrict__ coefA, float* __restrict__ coefB,
    float* __restrict__ u, const float* __restrict__ copy_src, float* __restrict__ copy_dst, int W, int H, float step)
{
    pdl_enter();
    const int L = 3 * W;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    if (i >= L)
        return;
    const int x = i / 3;
    const size_t idx = static_cast<size_t>(y) * L + i;
    const float c = __ldg(pr + idx);
    int cnt = 0;
    float lap = 0.0f;
    if ((x + 1) < (W - 1)) { lap += __ldg(pr + idx + 3); cnt += 1; }
    if ((x - 1) >= 0)      { lap += __ldg(pr + idx - 3); cnt += 1; }
    if ((y + 1) < (H - 1)) { lap += __ldg(pr + idx + L); cnt += 1; }
    if ((y - 1) >= 0)      { lap += __ldg(pr + idx - L); cnt += 1; }
    lap -= static_cast<float>(cnt) * c;
    const float w = __ldg(wt + idx);
    coefA[idx] = -step * (static_cast<float>(cnt) + w);
    coefB[idx] = step * (w * __ldg(tgt + idx) - lap);
    u[idx] = 0.0f;
    if (copy_dst)
        copy_dst[idx] = copy_src[idx];
}

__device__ __forceinline__ void sweep_value(float o, float S, float a, float b, float uo, float step, float mom,
    float& onew, float& unew)
{
    unew = __fmaf_rn(step, S, __fmaf_rn(a, o, b));
    onew = __fmaf_rn(mom, uo, o + unew);
}

// one Jacobi sweep, 4 consecutive floats of a row per thread (requires 3W % 4 == 0, 16B-aligned images)
__global__ void __launch_bounds__(256) solver_sweep_vec_kernel(const float* __restrict__ coefA,
    const float* __restrict__ coefB, float* __restrict__ u, const float* __restrict__ src, float* __restrict__ dst,
    int W, int H, float step, float mom)
{
    pdl_enter();
    const int L = 3 * W;
    const int L4 = L >> 2;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    if (j >= L4)
        return;
    const size_t base = static_cast<size_t>(y) * L + 4 * j;
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 cur = *reinterpret_cast<const float4*>(src + base);
    const float4 prv = j > 0 ? *reinterpret_cast<const float4*>(src + base - 4) : z;
    const float4 nxt = j < L4 - 1 ? *reinterpret_cast<const float4*>(src + base + 4) : z;
    const float4 up = (y - 1) >= 0 ? *reinterpret_cast<const float4*>(src + base - L) : z;
    const float4 dn = (y + 1) < (H - 1) ? *reinterpret_cast<const float4*>(src + base + L) : z;
    const float4 a = ldg_stream4(coefA + base);
    const float4 b = ldg_stream4(coefB + base);
    const float4 uo = *reinterpret_cast<const float4*>(u + base);

    const int i0 = 4 * j;
    const int rlim = 3 * (W - 2);  // right neighbour included iff i < 3(W-2)  <=>  x+1 < W-1
    const float o[4] = {cur.x, cur.y, cur.z, cur.w};
    const float r[4] = {cur.w, nxt.x, nxt.y, nxt.z};
    const float l[4] = {prv.y, prv.z, prv.w, cur.x};
    const float d[4] = {dn.x, dn.y, dn.z, dn.w};
    const float t[4] = {up.x, up.y, up.z, up.w};
    const float av[4] = {a.x, a.y, a.z, a.w};
    const float bv[4] = {b.x, b.y, b.z, b.w};
    const float uv[4] = {uo.x, uo.y, uo.z, uo.w};
    float on[4], un[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int i = i0 + k;
        const float rr = i < rlim ? r[k] : 0.0f;
        const float ll = i >= 3 ? l[k] : 0.0f;
        const float S = ((rr + ll) + d[k]) + t[k];
        sweep_value(o[k], S, av[k], bv[k], uv[k], step, mom, on[k], un[k]);
    }
    *reinterpret_cast<float4*>(u + base) = make_float4(un[0], un[1], un[2], un[3]);
    *reinterpret_cast<float4*>(dst + base) = make_float4(on[0], on[1], on[2], on[3]);
}

// scalar variant for rows whose length is not a multiple of 4 floats
__global__ void __launch_bounds__(256) solver_sweep_scalar_kernel(const float* __restrict__ coefA,
    const float* __restrict__ coefB, float* __restrict__ u, const float* __restrict__ src, float* __restrict__ dst,
    int W, int H, float step, float mom)
{
    pdl_enter();
    const int L = 3 * W;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    if (i >= L)
        return;
    const size_t idx = static_cast<size_t>(y) * L + i;
    const float o = src[idx];
    const float rr = i < 3 * (W - 2) ? src[idx + 3] : 0.0f;
    const float ll = i >= 3 ? src[idx - 3] : 0.0f;
    const float dd = (y + 1) < (H - 1) ? src[idx + L] : 0.0f;
    const float tt = (y - 1) >= 0 ? src[idx - L] : 0.0f;
    const float S = ((rr + ll) + dd) + tt;
    float on, un;
    sweep_value(o, S, coefA[idx], coefB[idx], u[idx], step, mom, on, un);
    u[idx] = un;
    dst[idx] = on;
}

__global__ void __launch_bounds__(256) copy_kernel(const float* __restrict__ src, float* __restrict__ dst, size_t n)
{
    const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n)
        dst[i] = src[i];
}

struct SolveBuffers {
    float *coefA, *coefB, *u, *alt, *u2;  // u2: second momentum buffer for the temporally blocked passes
};
constexpr int kSolveArrays = 5;

static SolveBuffers carve(void* ws, size_t n)
{
    char* p = static_cast<char*>(ws);
    const size_t s = align_up(n * sizeof(float));
    SolveBuffers b;
    b.coefA = reinterpret_cast<float*>(p);
    b.coefB = reinterpret_cast<float*>(p + s);
    b.u = reinterpret_cast<float*>(p + 2 * s);
    b.alt = reinterpret_cast<float*>(p + 3 * s);
    b.u2 = reinterpret_cast<float*>(p + 4 * s);
    return b;
}

static int launch_prepare(const float* pr, const float* tgt, const float* wt, const SolveBuffers& b,
    const float* copy_src, float* copy_dst, int W, int H, float step, cudaStream_t st)
{
    const dim3 grid(cdiv(3LL * W, 256), H);
    const int rc = launch_pdl(solver_prepare_kernel, grid, dim3(256), 0, st, pr, tgt, wt, b.coefA, b.coefB, b.u, copy_src,
        copy_dst, W, H, step);
    count_launch();
    return rc ? rc : launch_status();
}

// implemented in stab_solver_stream.cu: T sweeps per launch, T in {8, 4}
int solver_stream_pass(int T, const float* coefA, const float* coefB, const float* u_src, float* u_dst,
    const float* o_src, float* o_dst, int W, int H, float step, float mom, cudaStream_t st);
// implemented in stab_solver_rolled.cu: the same passes with a 4-step loop (16-byte aligned rows); false = not applicable
bool solver_rolled_pass(int T, const float* coefA, const float* coefB, const float* u_src, float* u_dst,
    const float* o_src, float* o_dst, int W, int H, float step, float mom, cudaStream_t st, int* rc);
extern std::atomic<int> g_stream_rolled, g_stream_qg, g_stream_permute;

std::atomic<int> g_solver_mode = 0;  // 0 auto, 1 unblocked sweeps only, 2 temporally blocked passes whenever iters >= 4
extern std::atomic<bool> g_stream_pair, g_stream_coop;  // stab_solver_stream.cu: variants of the blocked kernel
extern std::atomic<int> g_stream_band, g_stream_edge_top, g_stream_edge_bot;
std::atomic<bool> g_frame_fused = true;                 // vsc_frame_stabilize: fused stage A + solver set-up when possible

std::atomic<int> g_stream_tmain = 0;                     // deepest blocked pass: 0 = chosen per solve, else 8 or 10

// How `iters` sweeps are executed: n_hi blocked passes of t_hi sweeps, then n_lo passes of t_lo, then `rest` single
// unblocked sweeps.  Every pass is exact Jacobi, so any partition gives the same result.
struct SweepPlan {
    int n_hi, t_hi, n_lo, t_lo, rest;
    int passes() const { return n_hi + n_lo; }
    int flips() const { return n_hi + n_lo + rest; }  // number of out-buffer ping-pongs
    int depth(int k) const { return k < n_hi ? t_hi : t_lo; }
};
std::atomic<bool> g_plan_balanced = false;   // vsc_set_solver_mode(| 0x0800): passes of nearly equal depth, odd depths included
bool solver_rolled_takes(int W, const void* a, const void* b, const void* c, const void* d);   // stab_solver_rolled.cu

// even_only: the passes will run in stab_solver_stream.cu (rows that are not 16-byte aligned, or by request), which
// is built for depths 2, 4, 6, 8, 10 only
static SweepPlan plan_sweeps(int W, int H, int iters, bool even_only)
{
    SweepPlan p{0, 0, 0, 0, iters};
    // tiny images: the 3T-step pipeline fill and the 6T-float band halo dominate -> plain sweeps
    const bool big = H >= 48 && 3 * W >= 384;
    // images beyond 2^31 floats: the blocked kernel addresses with 32-bit element offsets
    const bool fits32 = 3LL * W * (static_cast<long long>(H) + 64) < 0x7fffffffLL;
    if (g_solver_mode == 1 || !fits32 || (g_solver_mode == 0 && !big) || iters < 2)
        return p;  // plain sweeps only
    if (g_plan_balanced && !even_only) {
        // The fewest passes: ceil(iters / tmax) passes of nearly equal depth, odd depths included, no single sweeps
        // (75 sweeps = 3 x 10 + 5 x 9, 150 = 15 x 10).  Bit-identical like every partition, and measured SLOWER than
        // the default plan at 1080p (733 vs 741-746 frames/s sustained, profiles/r2_plan_balanced_ab.txt): passes of 9
        // and 10 sweeps need the 384-float bands (168 registers), and at 960x540 the wider bands of the 8-sweep passes
        // win back more than the two passes saved.  Kept as a tested option.
        const int forced_t = g_stream_tmain;
        const int tmax = forced_t ? forced_t : (static_cast<long long>(W) * H >= 500000 ? 10 : 8);
        const int npass = (iters + tmax - 1) / tmax;
        const int base = iters / npass;
        p.n_hi = iters % npass;
        p.t_hi = base + 1;
        p.n_lo = npass - p.n_hi;
        p.t_lo = base;
        p.rest = 0;
        return p;
    }
    // main passes of 8 or 10 sweeps, one even tail pass, an odd sweep on its own.  With the quad-gather exchange ring an
    // 8-sweep pass costs 0.75 of a 10-sweep pass at every size (1080p 45.5 vs 60.9 us, 4K 143 vs 202 us, 960x540 18.2 vs
    // 25.2 us: profiles/r2_solver_qg_sweep.txt), so 10-sweep passes are taken only where they save enough passes
    // (150 = 18 x 8 + 6 beats 15 x 10; 20 = 2 x 10 beats 2 x 8 + 4)
    int tmain = 8;
    auto passes = [&](int t) { return iters / t + (((iters % t) & ~1) ? 1 : 0); };
    const bool large = static_cast<long long>(W) * H >= 900000;
    const int cost10 = g_stream_qg ? 134 : 121;   // cost of a 10-sweep pass in per cent of an 8-sweep pass
    if (g_stream_tmain == 10 || (g_stream_tmain == 0 && large && passes(10) * cost10 < passes(8) * 100))
        tmain = 10;
    p.n_hi = iters / tmain;
    p.t_hi = tmain;
    p.t_lo = (iters % tmain) & ~1;
    p.n_lo = p.t_lo ? 1 : 0;
    p.rest = iters & 1;
    // A pass costs about 11 us + 1.3 us per sweep at 960x540 (pipeline fill, launch): a 2-sweep tail pass costs
    // two thirds of an 8-sweep one (17.8 vs 23.5 us, profiles/r2_launches_bench_default.txt).  8 + 2 = 10: the last
    // main pass takes the tail along as ONE 10-sweep pass (75 sweeps at level 1 = 8 x 8 + 10 + 1 instead of
    // 9 x 8 + 2 + 1).  Same sweeps, same results.
    if (tmain == 8 && p.t_lo == 2 && p.n_hi >= 1) {
        p.n_hi -= 1;
        p.t_lo = 10;
    }
    return p;
}

// `iters` sweeps starting from the state in x (momentum in b.u, zeroed by the prepare kernel); the result
// lands in x if plan.flips() is even, else in y
// u_zero: b.u has NOT been zeroed; the first blocked pass is told so (null momentum input: it stages zeros instead
// of reading 12 bytes per pixel of them), and only a solve without any blocked pass zeroes the image here
static bool even_depths_only(const SolveBuffers& b, const float* x, const float* y, int W)
{
    return !solver_rolled_takes(W, b.coefA, b.coefB, b.u, b.u2) || !aligned16(x) || !aligned16(y);
}

static int run_sweeps(const SolveBuffers& b, float* x, float* y, int W, int H, int iters, float step, float mom,
    cudaStream_t st, float** result, bool u_zero = false)
{
    const SweepPlan plan = plan_sweeps(W, H, iters, even_depths_only(b, x, y, W));
    float* src = x;
    float* dst = y;
    float* us = b.u;
    float* ud = b.u2;
    int rc = VSC_OK;
    const int npass = plan.passes();
    if (u_zero && npass == 0 && iters > 0) {
        const cudaError_t e = cudaMemsetAsync(b.u, 0, static_cast<size_t>(W) * H * 3 * sizeof(float), st);
        if (e != cudaSuccess)
            return static_cast<int>(e);
        u_zero = false;
    }
    for (int k = 0; k < npass && rc == VSC_OK; ++k) {
        const int T = plan.depth(k);
        const float* uin = (u_zero && k == 0) ? nullptr : us;
        if (!solver_rolled_pass(T, b.coefA, b.coefB, uin, ud, src, dst, W, H, step, mom, st, &rc))
            rc = solver_stream_pass(T, b.coefA, b.coefB, uin, ud, src, dst, W, H, step, mom, st);
        float* t = src; src = dst; dst = t;
        t = us; us = ud; ud = t;
    }
    if (rc != VSC_OK)
        return rc;
    const bool vec = (3LL * W) % 4 == 0 && aligned16(x) && aligned16(y) && aligned16(b.coefA) && aligned16(b.coefB)
        && aligned16(us);
    for (int k = 0; k < plan.rest; ++k) {
        if (vec) {
            const dim3 grid(cdiv(3LL * W / 4, 128), H);
            rc = launch_pdl(solver_sweep_vec_kernel, grid, dim3(128), 0, st, b.coefA, b.coefB, us, src, dst, W, H, step, mom);
        } else {
            const dim3 grid(cdiv(3LL * W, 256), H);
            rc = launch_pdl(solver_sweep_scalar_kernel, grid, dim3(256), 0, st, b.coefA, b.coefB, us, src, dst, W, H, step,
                mom);
        }
        if (rc)
            return rc;
        float* t = src;
        src = dst;
        dst = t;
    }
    count_launch(plan.rest);
    *result = src;
    return launch_status();
}

}  // namespace vsc

using namespace vsc;

extern "C" size_t vsc_consist_solve_workspace_bytes(int W, int H)
{
    if (W <= 0 || H <= 0)
        return 0;
    return kSolveArrays * align_up(static_cast<size_t>(W) * H * 3 * sizeof(float));
}

extern "C" int vsc_consist_solve(const float* crntPr, const float* prevStabWarp, const float* consWt, int numIter,
    float stepSize, float momFac, float* consisOut, int W, int H, void* workspace, size_t workspace_bytes,
    vsc_stream_t stream)
{
    if (!crntPr || !prevStabWarp || !consWt || !consisOut || W <= 0 || H <= 0 || H > 65535 || numIter < 0)
        return VSC_E_INVALID;
    if (numIter == 0)
        return VSC_OK;
    if (!workspace || workspace_bytes < vsc_consist_solve_workspace_bytes(W, H) || !aligned16(workspace))
        return VSC_E_WORKSPACE;
    cudaStream_t st = as_stream(stream);
    const size_t n = static_cast<size_t>(W) * H * 3;
    const SolveBuffers b = carve(workspace, n);
    // start in the buffer that makes the last sweep (or blocked pass) land in consisOut
    // (the plan depends on whether the 4-step-loop kernel can take the passes: 16-byte aligned rows and buffers)
    const bool odd = (plan_sweeps(W, H, numIter, even_depths_only(b, consisOut, b.alt, W)).flips() & 1) != 0;
    int rc = launch_prepare(crntPr, prevStabWarp, consWt, b, odd ? consisOut : nullptr, odd ? b.alt : nullptr, W, H,
        stepSize, st);
    if (rc)
        return rc;
    float* res = nullptr;
    rc = run_sweeps(b, odd ? b.alt : consisOut, odd ? consisOut : b.alt, W, H, numIter, stepSize, momFac, st, &res);
    return rc;
}

extern "C" int vsc_set_solver_mode(int mode)
{
    const int lo = mode & 0xFFFF;
    if (mode < 0 || (lo & 0xF) > 2 || (lo & 0xC000) == 0xC000 || ((lo >> 8) & 7) > 4 || ((lo >> 12) & 3) > 2 || ((mode >> 29) & 1))
        return VSC_E_INVALID;
    g_solver_mode = lo & 0xF;
    g_stream_pair = (lo & 0x10) == 0;
    g_stream_coop = (lo & 0x20) == 0;
    g_frame_fused = (lo & 0x40) == 0;
    g_pdl = (lo & 0x80) == 0;
    g_stream_tmain = ((lo >> 12) & 3) == 0 ? 0 : 6 + 2 * ((lo >> 12) & 3);
    g_stream_band = (lo >> 8) & 7;
    g_stream_rolled = (lo & 0x8000) ? 0 : (lo & 0x4000) ? 2 : 1;
    g_plan_balanced = (lo & 0x0800) != 0;
    g_stream_qg = ((mode >> 28) & 1) ? !VSC_SOLVER_QG_DEFAULT : VSC_SOLVER_QG_DEFAULT;
    g_stream_permute = ((mode >> 30) & 1) ? !VSC_SOLVER_PERMUTE_DEFAULT : VSC_SOLVER_PERMUTE_DEFAULT;
    g_stream_edge_top = ((mode >> 16) & 0x3F) - 1;   // 0 = default
    g_stream_edge_bot = ((mode >> 22) & 0x3F) - 1;
    return VSC_OK;
}

extern "C" void vsc_hyper_params_default(vsc_hyper_params* p)
{
    if (!p)
        return;
    p->alpha = 6800.0f;
    p->beta = 6800.0f;
    p->gamma = 2.0f;
    p->pyramidLevels = 2;
    p->numIter = 150;
    p->stepSize = 0.15f;
    p->momFac = 0.15f;
}

namespace {
constexpr int kMaxLevels = 8;
struct LevelDims {
    int w[kMaxLevels], h[kMaxLevels];
    size_t n[kMaxLevels];
};
LevelDims level_dims(int W, int H, int levels)
{
    LevelDims d{};
    d.w[0] = W;
    d.h[0] = H;
    for (int j = 1; j < levels; ++j) {
        d.w[j] = d.w[j - 1] / 2;  // videostabilizer.cpp:126-127
        d.h[j] = d.h[j - 1] / 2;
    }
    for (int j = 0; j < levels; ++j)
        d.n[j] = static_cast<size_t>(d.w[j]) * d.h[j] * 3;
    return d;
}
}  // namespace

extern "C" size_t vsc_frame_solve_workspace_bytes(int W, int H, int pyramidLevels)
{
    if (W <= 0 || H <= 0 || pyramidLevels < 1 || pyramidLevels > kMaxLevels)
        return 0;
    const LevelDims d = level_dims(W, H, pyramidLevels);
    size_t total = kSolveArrays * align_up(d.n[0] * sizeof(float));  // level 0: A, B, u, alt, u2
    for (int j = 1; j < pyramidLevels; ++j)
        total += (4 + kSolveArrays) * align_up((d.n[j] ? d.n[j] : 1) * sizeof(float));  // pr, tgt, wt, out + 5
    return total;
}

namespace vsc {
int launch_stage_a_prep(const float* origPrev, const float* origCur, const float* origNext, const float* procPrev,
    const float* procCur, const float* procNext, const float* lastStab, const float* flowFwd, const float* flowBwd,
    int flowC, float alpha, float beta, float gamma, float step, float* coefA, float* coefB, float* pr1, float* tg1,
    float* wt1, int W, int H, cudaStream_t st);
}

// stageA != nullptr: the fused path -- level-0 coefficients and the level-1 inputs are produced by
// stage_a_prep_kernel (2 levels, even W and H) instead of by solver_prepare_kernel + three resizes
struct StageAInputs {
    const float *origPrev, *origCur, *origNext, *procPrev, *procNext, *lastStab, *flowFwd, *flowBwd;
    int flowC;
};

static int frame_solve_impl(const float* procCur, const float* adapCmbPr, const float* consWt,
    const vsc_hyper_params* p, float* consisOut, int W, int H, void* workspace, vsc_stream_t stream,
    const StageAInputs* stageA)
{
    const int levels = p->pyramidLevels;
    const LevelDims d = level_dims(W, H, levels);
    if (d.w[levels - 1] < 1 || d.h[levels - 1] < 1)
        return VSC_E_INVALID;
    cudaStream_t st = as_stream(stream);

    // carve the workspace
    char* ws = static_cast<char*>(workspace);
    SolveBuffers sb[kMaxLevels];
    const float* pr[kMaxLevels];
    const float* tg[kMaxLevels];
    const float* wt[kMaxLevels];
    float* out[kMaxLevels];
    sb[0] = carve(ws, d.n[0]);
    ws += kSolveArrays * align_up(d.n[0] * sizeof(float));
    pr[0] = procCur;
    tg[0] = adapCmbPr;
    wt[0] = consWt;
    out[0] = nullptr;  // chosen below
    for (int j = 1; j < levels; ++j) {
        const size_t s = align_up(d.n[j] * sizeof(float));
        float* q = reinterpret_cast<float*>(ws);
        pr[j] = q;
        tg[j] = reinterpret_cast<float*>(ws + s);
        wt[j] = reinterpret_cast<float*>(ws + 2 * s);
        out[j] = reinterpret_cast<float*>(ws + 3 * s);
        sb[j] = carve(ws + 4 * s, d.n[j]);
        ws += (4 + kSolveArrays) * s;
    }

    int rc;
    if (stageA) {
        rc = launch_stage_a_prep(stageA->origPrev, stageA->origCur, stageA->origNext, stageA->procPrev, procCur,
            stageA->procNext, stageA->lastStab, stageA->flowFwd, stageA->flowBwd, stageA->flowC, p->alpha, p->beta,
            p->gamma, p->stepSize, sb[0].coefA, sb[0].coefB, const_cast<float*>(pr[1]), const_cast<float*>(tg[1]),
            const_cast<float*>(wt[1]), W, H, st);
        if (rc) return rc;
        // (momentum starts at 0: the level-0 solve below is told so instead of zeroing 12 bytes per pixel here)
    }
    // pyramid down (videostabilizer.cpp:210-216).  pyrConsisOut[0] is a copy of the processed frame (:209),
    // so its downsampled version equals pyrPr[j] -- computed once and copied.
    for (int j = 1; j < levels && !stageA; ++j) {
        rc = vsc_bilinear(pr[j - 1], d.w[j - 1], d.h[j - 1], 3, const_cast<float*>(pr[j]), d.w[j], d.h[j], 3, stream);
        if (rc) return rc;
        rc = vsc_bilinear(tg[j - 1], d.w[j - 1], d.h[j - 1], 3, const_cast<float*>(tg[j]), d.w[j], d.h[j], 3, stream);
        if (rc) return rc;
        rc = vsc_bilinear(wt[j - 1], d.w[j - 1], d.h[j - 1], 3, const_cast<float*>(wt[j]), d.w[j], d.h[j], 3, stream);
        if (rc) return rc;
    }

    // coarse to fine (:219-227)
    float* coarse_result = nullptr;
    for (int j = levels - 1; j >= 0; --j) {
        const int iters = p->numIter / (j + 1);
        float* final_probe = (j == 0) ? consisOut : out[j];
        const bool odd = (plan_sweeps(d.w[j], d.h[j], iters, even_depths_only(sb[j], final_probe, sb[j].alt, d.w[j])).flips() & 1) != 0;
        // x = buffer holding the initial state, y = the other one; the result lands in `final`
        float* final_buf = (j == 0) ? consisOut : out[j];
        float* x = odd ? sb[j].alt : final_buf;
        float* y = odd ? final_buf : sb[j].alt;
        // initial state: coarsest level = downsampled processed frame (== pr[j]); else upsampled coarser result
        const float* copy_src = nullptr;
        float* copy_dst = nullptr;
        if (j == levels - 1) {
            copy_src = pr[j];
            copy_dst = x;
        } else {
            rc = vsc_bilinear(coarse_result, d.w[j + 1], d.h[j + 1], 3, x, d.w[j], d.h[j], 3, stream);
            if (rc) return rc;
        }
        if (!(stageA && j == 0)) {
            rc = launch_prepare(pr[j], tg[j], wt[j], sb[j], copy_src, copy_dst, d.w[j], d.h[j], p->stepSize, st);
            if (rc) return rc;
        }
        float* res = nullptr;
        rc = run_sweeps(sb[j], x, y, d.w[j], d.h[j], iters, p->stepSize, p->momFac, st, &res, stageA && j == 0);
        if (rc) return rc;
        coarse_result = res;  // == final_buf (also when iters == 0: x == final_buf since 0 is even)
    }
    return VSC_OK;
}

extern "C" int vsc_frame_solve(const float* procCur, const float* adapCmbPr, const float* consWt,
    const vsc_hyper_params* p, float* consisOut, int W, int H, void* workspace, size_t workspace_bytes,
    vsc_stream_t stream)
{
    if (!procCur || !adapCmbPr || !consWt || !p || !consisOut || W <= 0 || H <= 0 || H > 65535)
        return VSC_E_INVALID;
    const int levels = p->pyramidLevels;
    if (levels < 1 || levels > kMaxLevels || p->numIter < 0)
        return VSC_E_INVALID;
    if (!workspace || workspace_bytes < vsc_frame_solve_workspace_bytes(W, H, levels) || !aligned16(workspace))
        return VSC_E_WORKSPACE;
    return frame_solve_impl(procCur, adapCmbPr, consWt, p, consisOut, W, H, workspace, stream, nullptr);
}

extern "C" size_t vsc_frame_stabilize_workspace_bytes(int W, int H, int pyramidLevels)
{
    const size_t base = vsc_frame_solve_workspace_bytes(W, H, pyramidLevels);
    if (base == 0)
        return 0;
    return base + 2 * align_up(static_cast<size_t>(W) * H * 3 * sizeof(float));  // adapCmbPr, consWt (generic path)
}

extern "C" int vsc_frame_stabilize(const float* origPrev, const float* origCur, const float* origNext,
    const float* procPrev, const float* procCur, const float* procNext, const float* lastStab, const float* flowFwd,
    const float* flowBwd, int flow_channels, const vsc_hyper_params* p, float* consisOut, int W, int H,
    void* workspace, size_t workspace_bytes, vsc_stream_t stream)
{
    if (!origPrev || !origCur || !origNext || !procPrev || !procCur || !procNext || !lastStab || !flowFwd || !flowBwd
        || !p || !consisOut || W < 2 || H < 2 || H > 65535 || (flow_channels != 2 && flow_channels != 3))
        return VSC_E_INVALID;
    const int levels = p->pyramidLevels;
    if (levels < 1 || levels > kMaxLevels || p->numIter < 0)
        return VSC_E_INVALID;
    if (!workspace || workspace_bytes < vsc_frame_stabilize_workspace_bytes(W, H, levels) || !aligned16(workspace))
        return VSC_E_WORKSPACE;
    if (level_dims(W, H, levels).w[levels - 1] < 1 || level_dims(W, H, levels).h[levels - 1] < 1)
        return VSC_E_INVALID;
    // (the fused kernel addresses with 32-bit element offsets)
    const bool fused = levels == 2 && (W % 2 == 0) && (H % 2 == 0) && g_frame_fused && 3LL * W * H < 0x7fffffffLL;
    if (fused) {
        const StageAInputs in{origPrev, origCur, origNext, procPrev, procNext, lastStab, flowFwd, flowBwd,
            flow_channels};
        return frame_solve_impl(procCur, nullptr, nullptr, p, consisOut, W, H, workspace, stream, &in);
    }
    char* extra = static_cast<char*>(workspace) + vsc_frame_solve_workspace_bytes(W, H, levels);
    float* tgt = reinterpret_cast<float*>(extra);
    float* wt = reinterpret_cast<float*>(extra + align_up(static_cast<size_t>(W) * H * 3 * sizeof(float)));
    const int rc = vsc_stage_a_fused(origPrev, origCur, origNext, procPrev, procCur, procNext, lastStab, flowFwd,
        flowBwd, flow_channels, p->alpha, p->beta, p->gamma, nullptr, tgt, wt, W, H, stream);
    if (rc)
        return rc;
    return frame_solve_impl(procCur, tgt, wt, p, consisOut, W, H, workspace, stream, nullptr);
}
