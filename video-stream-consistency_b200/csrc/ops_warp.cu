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
#include "ops_warp.cuh"

namespace vsc {

// grid: x = 256-pixel groups of one image, y = channel chunk, z = n.  One pixel per thread: the 32 lanes of a
// warp sample 32 neighbouring positions, so each of the four gathers touches a handful of 32-byte sectors and
// the store is one full 128-byte line (a 4-pixels-per-thread variant with 128-bit stores was measured 4x
// worse: its gathers stride 16 bytes between lanes, 29 sectors per request -- profiles/r1_notes.md).
// The channel loop is unrolled 4x: 16 independent gathers in flight per thread.
// Measured and removed again (profiles/r1_notes.md, r1_warp_variants_ncu.txt; all bit-identical): a CTA walking over
// 2-8 pixel groups with the next group's flow prefetched (8-16 % slower: fewer, longer CTAs), and right-hand taps
// taken from the neighbour lane by shuffle instead of a second gather (58 vs 37 us: the shuffle waits for the load,
// which serialises what were independent gathers).  ncu: 27.75 instructions per value in this loop, 46 per value
// overall -- 40 % of all instructions are the per-pixel set-up (the reference's double-precision mask arithmetic,
// p / W) amortised over only 8-16 channels per thread; whole-tensor chunks (vsc_set_warp_mode(1 | 1 << 4)) cut
// that but leave too few CTAs, and time the same.
__global__ void __launch_bounds__(256) warp_nchw_kernel(const float* __restrict__ in, const float* __restrict__ flow,
    float* __restrict__ out, int C, int H, int W, int chunk)
{
    pdl_enter();
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



// ---- tiled variant ------------------------------------------------------------------------------------------
// The four taps of a pixel are the 2x2 quad at (xb, yb) = (xL, yT) clamped into the image, so ONE 64-bit address
// per channel (plus a second one row below) serves all four gathers with constant offsets 0 / +1, instead of four
// separately formed addresses: the per-value instruction count roughly halves.  At the image border, where only
// some corners are valid, the valid corners land in other slots of the clamped quad; the slots keep the
// reference's accumulation order (00, 10, 01, 11) among the corners that are used, and unused slots are neither
// read nor weighted.
struct WarpQuad {
    int o;                  // element offset of slot A = (yb, xb) inside one H*W plane
    float wA, wB, wC, wD;   // weights of slots (yb,xb) (yb,xb+1) (yb+1,xb) (yb+1,xb+1); 0 when unused
    unsigned used;          // bit k set: slot k is read
};

__device__ __forceinline__ WarpQuad warp_quad(int x, int y, float fu, float fv, int W, int H)
{
    const WarpTap t = warp_setup(x, y, fu, fv, W, H);
    const float xL = floorf(static_cast<float>(x) + fu);
    const float yT = floorf(static_cast<float>(y) + fv);
    // fmaxf/fminf drop a NaN operand, so the conversions below are always in range
    const int xb = static_cast<int>(fminf(fmaxf(xL, 0.0f), static_cast<float>(max(W - 2, 0))));
    const int yb = static_cast<int>(fminf(fmaxf(yT, 0.0f), static_cast<float>(max(H - 2, 0))));
    const int ob = yb * W + xb;
    // slot of a used corner = (row - yb) * 2 + (col - xb), recovered from its plane offset
    auto slot = [&](int o) {
        const int d = o - ob;   // 0, 1, W or W + 1 for a used corner
        return (d >= W ? 2 : 0) + ((d == 1 || d == W + 1) ? 1 : 0);
    };
    // W == 1: a used corner in the next row has d == W == 1; rows win (there is no column xb + 1)
    auto slot1 = [&](int o) { return (o - ob) >= 1 ? 2 : 0; };
    const int s00 = W > 1 ? slot(t.o00) : slot1(t.o00), s10 = W > 1 ? slot(t.o10) : slot1(t.o10);
    const int s01 = W > 1 ? slot(t.o01) : slot1(t.o01), s11 = W > 1 ? slot(t.o11) : slot1(t.o11);
    const bool v00 = t.valid & 1u, v10 = t.valid & 2u, v01 = t.valid & 4u, v11 = t.valid & 8u;
    auto pick = [&](int k) {
        return (v00 && s00 == k) ? t.w00 : (v10 && s10 == k) ? t.w10 : (v01 && s01 == k) ? t.w01
            : (v11 && s11 == k) ? t.w11 : 0.0f;
    };
    auto hit = [&](int k) { return (v00 && s00 == k) || (v10 && s10 == k) || (v01 && s01 == k) || (v11 && s11 == k); };
    WarpQuad q;
    q.o = ob;
    q.wA = pick(0);
    q.wB = pick(1);
    q.wC = pick(2);
    q.wD = pick(3);
    q.used = (hit(0) ? 1u : 0u) | (hit(1) ? 2u : 0u) | (hit(2) ? 4u : 0u) | (hit(3) ? 8u : 0u);
    return q;
}

// grid: x = 32-pixel column tiles, y = 8-row tiles, z = n * channel chunks.  A CTA is a 32x8 pixel tile: each
// warp stores one 128-byte row segment per channel, and the bottom taps of a row are the top taps of the row
// below it in the same CTA, so they hit in L1 instead of going back to L2.
constexpr int kWarpTileW = 32, kWarpTileH = 8;
// TILED: grid x = 32-pixel column tiles, y = 8-row tiles; otherwise grid x = 256-pixel groups of the flattened
// image, y = 1 (the linear kernel's mapping with the quad's cheaper addressing); z = n * channel chunks.
template <bool TILED>
__global__ void __launch_bounds__(kWarpTileW * kWarpTileH) warp_nchw_quad_kernel(const float* __restrict__ in,
    const float* __restrict__ flow, float* __restrict__ out, int C, int H, int W, int chunk, int nchunk)
{
    int x, y;
    if constexpr (TILED) {
        x = blockIdx.x * kWarpTileW + (threadIdx.x & (kWarpTileW - 1));
        y = blockIdx.y * kWarpTileH + (threadIdx.x / kWarpTileW);
        if (x >= W || y >= H)
            return;
    } else {
        const int pp = blockIdx.x * blockDim.x + threadIdx.x;
        if (pp >= H * W)
            return;
        y = pp / W;
        x = pp - y * W;
    }
    const int HW = H * W;
    const int n = blockIdx.z / nchunk;
    const int c0 = (blockIdx.z - n * nchunk) * chunk;
    const int c1 = min(C, c0 + chunk);
    const int p = y * W + x;
    const float* fl = flow + static_cast<size_t>(n) * 2 * HW + p;
    const WarpQuad q = warp_quad(x, y, ldg_stream(fl), ldg_stream(fl + HW), W, H);
    const bool uA = q.used & 1u, uB = q.used & 2u, uC = q.used & 4u, uD = q.used & 8u;

    const float* pt = in + (static_cast<size_t>(n) * C + c0) * HW + q.o;   // slot A of channel c
    const float* pb = pt + W;                                             // slot C
    float* op = out + (static_cast<size_t>(n) * C + c0) * HW + p;
    auto sample = [&](const float* t, const float* b) {
        const float va = uA ? __ldg(t) : 0.0f;
        const float vb = uB ? __ldg(t + 1) : 0.0f;
        const float vc = uC ? __ldg(b) : 0.0f;
        const float vd = uD ? __ldg(b + 1) : 0.0f;
        float v = q.wA * va;
        v = __fmaf_rn(q.wB, vb, v);
        v = __fmaf_rn(q.wC, vc, v);
        v = __fmaf_rn(q.wD, vd, v);
        return v;
    };
    int c = c0;
    for (; c + 4 <= c1; c += 4) {
        const float* t1 = pt + HW;
        const float* b1 = pb + HW;
        const float* t2 = t1 + HW;
        const float* b2 = b1 + HW;
        const float* t3 = t2 + HW;
        const float* b3 = b2 + HW;
        const float v0 = sample(pt, pb);
        const float v1 = sample(t1, b1);
        const float v2 = sample(t2, b2);
        const float v3 = sample(t3, b3);
        float* o1 = op + HW;
        float* o2 = o1 + HW;
        float* o3 = o2 + HW;
        __stcs(op, v0);
        __stcs(o1, v1);
        __stcs(o2, v2);
        __stcs(o3, v3);
        pt = t3 + HW;
        pb = b3 + HW;
        op = o3 + HW;
    }
    for (; c < c1; ++c, pt += HW, pb += HW, op += HW)
        __stcs(op, sample(pt, pb));
}

std::atomic<int> g_warp_mode = 0;    // 0 default, 1 linear one-pixel-per-thread kernel, 2 tiled, 3 quad addressing, flattened,
                        // 4 TMA-staged tiles wherever applicable (ops_warp_staged.cu)
bool warp_staged_applicable(const float* in, int N, int C, int H, int W);
int launch_warp_staged(const float* in, const float* flow, float* out, int N, int C, int H, int W, int nchunk_req,
    cudaStream_t st, bool* launched);
// smallest map the staged kernel takes by default (measured: profiles/r2_warp_staged_events.txt)
constexpr long long kWarpStagedMinPixels = 120000;
std::atomic<int> g_warp_nchunk = 0;  // 0 = automatic channel split of the linear kernel, else the number of channel chunks

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
    // measured (profiles/r1_time_ops_v13.txt, r1_warp_linear_vs_tiled_ncu.txt): on smooth flow -- what an optical-flow
    // network produces -- the linear kernel is 4-16 % faster at every level shape; the tiled one wins only on
    // scattered flow (i.i.d. sigma = 2 px: 53 vs 59 us at 32x544x960).  Default: linear; tiled on request.
    // large maps with 16-byte aligned rows: tiles staged in shared memory by TMA (gathers for tiles with scattered flow)
    const bool big = static_cast<long long>(H) * W >= kWarpStagedMinPixels;
    if ((g_warp_mode == 4 || (g_warp_mode == 0 && big)) && warp_staged_applicable(in, N, C, H, W)) {
        bool launched = false;
        const int rc = launch_warp_staged(in, flow, out, N, C, H, W, g_warp_nchunk, as_stream(stream), &launched);
        if (rc || launched)
            return rc;
    }
    if (g_warp_mode == 2 || g_warp_mode == 3) {
        const bool tiled = g_warp_mode == 2;
        const unsigned tx = tiled ? cdiv(W, kWarpTileW) : cdiv(static_cast<long long>(H) * W, 256);
        const unsigned ty = tiled ? cdiv(H, kWarpTileH) : 1;
        const long long tiles = static_cast<long long>(tx) * ty * N;
        const long long want = 4LL * sm_count() * 8;
        int nchunk = static_cast<int>((want + tiles - 1) / tiles);
        if (nchunk < 1) nchunk = 1;
        if (nchunk > (C + 7) / 8) nchunk = (C + 7) / 8;
        int chunk = (C + nchunk - 1) / nchunk;
        chunk = (chunk + 3) / 4 * 4;  // whole unrolled groups
        nchunk = (C + chunk - 1) / chunk;
        if (ty > 65535 || static_cast<long long>(N) * nchunk > 65535)
            return VSC_E_INVALID;
        const dim3 grid(tx, ty, static_cast<unsigned>(N * nchunk));
        if (tiled)
            warp_nchw_quad_kernel<true><<<grid, kWarpTileW * kWarpTileH, 0, as_stream(stream)>>>(in, flow, out, C, H, W,
                chunk, nchunk);
        else
            warp_nchw_quad_kernel<false><<<grid, 256, 0, as_stream(stream)>>>(in, flow, out, C, H, W, chunk, nchunk);
        count_launch();
        return launch_status();
    }
    const unsigned gx = cdiv(static_cast<long long>(H) * W, 256);
    // enough blocks for >= 4 waves of 148 SMs x 8 resident CTAs when the tensor allows, chunks of >= 8 channels
    const long long want = 4LL * sm_count() * 8;
    int nchunk = static_cast<int>((want + static_cast<long long>(gx) * N - 1) / (static_cast<long long>(gx) * N));
    if (nchunk < 1) nchunk = 1;
    if (nchunk > (C + 7) / 8) nchunk = (C + 7) / 8;
    if (const int forced = g_warp_nchunk; forced > 0) nchunk = forced < C ? forced : C;
    if (nchunk > 65535) nchunk = 65535;
    int chunk = (C + nchunk - 1) / nchunk;
    chunk = (chunk + 3) / 4 * 4;  // whole unrolled groups
    nchunk = (C + chunk - 1) / chunk;
    const dim3 grid(gx, nchunk, N);
    const int rc = launch_pdl(warp_nchw_kernel, grid, dim3(256), 0, as_stream(stream), in, flow, out, C, H, W, chunk);
    count_launch();
    return rc ? rc : launch_status();
}

extern "C" int vsc_set_warp_mode(int mode)
{
    if (mode < 0 || (mode & 0xF) > 4 || mode > 0xFFF)
        return VSC_E_INVALID;
    vsc::g_warp_mode = mode & 0xF;
    vsc::g_warp_nchunk = mode >> 4;
    return VSC_OK;
}
