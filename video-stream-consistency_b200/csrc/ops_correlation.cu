// custom::Correlation for sm_100a -- replaces CorrelationKernel::ComputeCUDA,
// blob_rearrange_kernel, correlation_cuda_forward_kernel and correlation_old_kernel
// (reference: src/ort_custom_ops/src/opticalflow/correlation_cuda.cc:29-138,
// correlation_cuda.cu:33-61,98-175,183-265,334-442; CPU definition correlation.cc:148-275).
//
//   out[n, ph, pw, h, w] = sum_c in1[n,c,h,w] * in2[n,c,h+ph-md,w+pw-md]      (outside the image: 0)
//
// The reference transposes both inputs to NHWC into two buffers it cudaMallocs and cudaFrees per
// call, then runs one 32-thread block per output pixel with 81 barriers.  Here the contraction
// reads NCHW directly, with no scratch memory and no allocation:
//   * a CTA owns a 32x8 tile of output pixels for all 81 displacements;
//   * per chunk of KC=8 channels the in1 tile [8][8][32] and the in2 tile + 4-pixel halo [8][16][40]
//     are staged in shared memory by TMA (cp.async.bulk.tensor, 4-D tensor maps over [N][C][H][W]).
//     The out-of-bounds zero fill of TMA implements the op's zero padding (negative start
//     coordinates at the image border, channels beyond C).  A dedicated producer warp keeps a
//     3-stage full/empty mbarrier ring ahead of the six consumer warps;
//   * a consumer thread owns 4 consecutive pixels x 9 horizontal x 3 vertical displacements = 108 fp32
//     accumulators in registers and walks the channels in order with explicit FMAs: per channel 10
//     128-bit LDS feed 108 FMAs, and every output value is the same sequential-over-c fp32 sum the
//     reference's CPU kernel forms;
//   * results leave as 128-bit streaming stores straight into the [N,9,9,H,W] layout (the legacy
//     [N,81,H,W] layout is byte-identical; it only adds the division by C).
// The default kernel for maps at least 96 pixels wide is correlation_md4_share_kernel<32> further down (skewed tiles
// whose lane pairs read the same in2 rows, two CTAs of six warps per SM); this first kernel serves narrower maps.
// TMA needs 16-byte aligned row strides (W % 4 == 0) and base pointers; other shapes take the same
// compute loop behind a plain-load stager.  Any max_displacement other than 4 takes a one-thread-
// per-output kernel (no shipped model uses one; model_spec.py:161-162).
#include <cuda.h>

#include "vsc_common.cuh"

namespace vsc {

constexpr int kMD = 4;
constexpr int kP = 2 * kMD + 1;           // 9
constexpr int kTW = 32, kTH = 8;          // output tile
constexpr int kBW = kTW + 2 * kMD;        // 40: in2 tile row (floats), 160 B
constexpr int kBH = kTH + 2 * kMD;        // 16
constexpr int kKC = 8;                    // channels per stage
constexpr int kConsumers = 192;           // 64 pixel-quads x 3 vertical-displacement groups (6 warps)
constexpr int kASize = kTH * kTW;         // 256 floats per channel
constexpr int kBSize = kBH * kBW;         // 640 floats per channel
constexpr int kStages = 3;
constexpr int kStageFloats = kKC * (kASize + kBSize);                 // 7168 floats = 28 KB
constexpr unsigned kStageBytes = kStageFloats * sizeof(float);

// ---- mbarrier / TMA primitives (PTX ISA 8.x, sm_90+) ------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return static_cast<unsigned>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_4d(unsigned dst, const CUtensorMap* map, unsigned bar, int c0, int c1, int c2,
    int c3)
{
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<unsigned long long>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}

// ---- the contraction of one staged channel chunk (shared by both stagers) ---------------------------
__device__ __forceinline__ void correlate_chunk(const float* __restrict__ sA, const float* __restrict__ sB, int r, int qc,
    int g, float (&acc)[3][kP][4])
{
#pragma unroll 1
    for (int c = 0; c < kKC; ++c) {
        const float4 a4 = *reinterpret_cast<const float4*>(&sA[c * kASize + r * kTW + qc]);
        const float av[4] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const float* brow = &sB[c * kBSize + (r + 3 * g + i) * kBW + qc];
            // the 12 in2 values of the row are consumed one 128-bit load at a time (value m feeds the
            // accumulators (j = m - k, k)), so only 4 of them are live: keeps the kernel at <= 128 registers,
            // i.e. two CTAs per SM
#pragma unroll
            for (int h = 0; h < 3; ++h) {
                const float4 b4 = *reinterpret_cast<const float4*>(brow + 4 * h);
                const float bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
                for (int mm = 0; mm < 4; ++mm) {
                    const int m = 4 * h + mm;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int j = m - k;
                        if (j >= 0 && j < kP)
                            acc[i][j][k] = __fmaf_rn(av[k], bv[mm], acc[i][j][k]);
                    }
                }
            }
        }
    }
}

// The 108 values of a thread leave as 27 streaming 128-bit stores.  The epilogue is kept LEAN on purpose: ncu's
// per-instruction samples of the 64x8 kernel (profiles/r2_correlation_ncu.txt) put a third of all warp time into the
// old epilogue -- 38 instructions per store (a 64-bit address product, the legacy select and the alignment test
// inside the loops) run once per tile from a cold instruction cache while the FMA pipes idle.  Here: one base
// pointer, one 64-bit add per store, the legacy division and the unaligned case hoisted out of the loops.
// rot: slot i of the thread holds vertical displacement 3g + (i + rot) % 3 (the shared-row kernel's odd lanes).
__device__ __forceinline__ void correlation_store(float* __restrict__ out, float (&acc)[3][kP][4], int n, int g, int y,
    int x, int H, int W, float divisor, int legacy, int vec_store, int rot = 0)
{
    if (y < 0 || y >= H || x >= W)
        return;
    const size_t HW = static_cast<size_t>(H) * W;
    if (legacy) {  // total_sum / (float)C, a true division (correlation_cuda.cu:259-261)
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < kP; ++j)
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    acc[i][j][k] = acc[i][j][k] / divisor;
    }
    float* base = out + (static_cast<size_t>(n) * kP + 3 * g) * kP * HW + static_cast<size_t>(y) * W + x;
    const size_t plane = kP * HW;
    if (vec_store) {  // W % 4 == 0 and out 16B-aligned: x+3 < W holds as well
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const int sh = (i + rot >= 3) ? i + rot - 3 : i + rot;
            float* o = base + sh * plane;
#pragma unroll
            for (int j = 0; j < kP; ++j) {
                stg_stream4(o, make_float4(acc[i][j][0], acc[i][j][1], acc[i][j][2], acc[i][j][3]));
                o += HW;
            }
        }
    } else {
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const int sh = (i + rot >= 3) ? i + rot - 3 : i + rot;
            float* o = base + sh * plane;
#pragma unroll
            for (int j = 0; j < kP; ++j) {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (x + k < W)
                        o[k] = acc[i][j][k];
                o += HW;
            }
        }
    }
}

// ---- TMA-staged kernel: 6 consumer warps + 1 producer warp ------------------------------------------
__global__ void __maxnreg__(128) correlation_md4_tma_kernel(
    const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, float* __restrict__ out, int C,
    int H, int W, float divisor, int legacy, int vec_store)
{
    extern __shared__ __align__(128) unsigned char smem_bytes[];
    float* stage_mem = reinterpret_cast<float*>(smem_bytes);
    __shared__ __align__(8) unsigned long long bars[2 * kStages];  // full[0..2], empty[0..2]

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int w0 = blockIdx.x * kTW;
    const int h0 = blockIdx.y * kTH;
    const int n = blockIdx.z;
    const int nchunks = (C + kKC - 1) / kKC;
    const unsigned bar0 = smem_u32(bars);

    if (tid == 0) {
#pragma unroll
        for (int i = 0; i < kStages; ++i) {
            mbar_init(bar0 + 8 * i, 1);                       // full: the producer's arrive.expect_tx
            mbar_init(bar0 + 8 * (kStages + i), kConsumers / 32);  // empty: one arrival per consumer warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == kConsumers / 32) {
        // ===== producer warp: one elected lane drives TMA =====
        if ((tid & 31) == 0) {
            for (int j = 0; j < nchunks; ++j) {
                const int s = j % kStages;
                if (j >= kStages)
                    mbar_wait(bar0 + 8 * (kStages + s), ((j / kStages) - 1) & 1);
                const unsigned full = bar0 + 8 * s;
                mbar_expect_tx(full, kStageBytes);
                const unsigned dstA = smem_u32(stage_mem + s * kStageFloats);
                const unsigned dstB = dstA + kKC * kASize * sizeof(float);
                tma_load_4d(dstA, &mapA, full, w0, h0, j * kKC, n);
                tma_load_4d(dstB, &mapB, full, w0 - kMD, h0 - kMD, j * kKC, n);
            }
        }
        return;
    }

    // ===== consumer warps =====
    const int q = tid & 63;        // pixel quad inside the tile
    const int g = tid >> 6;        // vertical displacement group: ph = 3g .. 3g+2 (warp-uniform)
    const int r = q >> 3;          // tile row 0..7
    const int qc = (q & 7) * 4;    // first tile column of the quad
    float acc[3][kP][4];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < kP; ++j)
#pragma unroll
            for (int k = 0; k < 4; ++k)
                acc[i][j][k] = 0.0f;

    for (int j = 0; j < nchunks; ++j) {
        const int s = j % kStages;
        mbar_wait(bar0 + 8 * s, (j / kStages) & 1);
        const float* sA = stage_mem + s * kStageFloats;
        correlate_chunk(sA, sA + kKC * kASize, r, qc, g, acc);
        __syncwarp();
        if ((tid & 31) == 0)
            mbar_arrive(bar0 + 8 * (kStages + s));
    }
    correlation_store(out, acc, n, g, h0 + r, w0 + qc, H, W, divisor, legacy, vec_store);
}

// ---- wide TMA kernel: 64x8 tile, ONE CTA of 12 warps per SM, all of them consumers -------------------
// The 32x8 kernel above is register-starved: two CTAs x (6 consumer + 1 producer warps) leave 128 registers
// per thread, 108 of them accumulators, so every 128-bit LDS is consumed immediately and its latency is exposed
// (ncu: short_scoreboard the top stall, FMA pipe 35 % busy).  Here a single CTA of exactly 12 warps owns the SM
// (168 registers per thread): the in2 loads are software-pipelined one 128-bit load ahead of the FMAs, and the
// TMA producer is simply lane 0 of warp 0, which refills the stage freed two iterations ago before computing.
constexpr int kWTW = 64;                               // tile width
constexpr int kWBW = kWTW + 2 * kMD;                   // 72 floats = 288 B rows
constexpr int kWASize = kTH * kWTW;                    // 512
constexpr int kWBSize = kBH * kWBW;                    // 1152
constexpr int kWThreads = 384;                         // 128 pixel quads x 3 displacement groups
constexpr int kWStageFloats = kKC * (kWASize + kWBSize);  // 13312 floats = 52 KB
constexpr unsigned kWStageBytes = kWStageFloats * sizeof(float);

// Persistent: gridDim.x = min(tiles, SMs) CTAs walk the tiles (tile = blockIdx.x + i * gridDim.x) with ONE
// continuous chunk counter, so the loads of the next tile's first channel chunks are already in flight while
// the 108 accumulators of the current tile are being stored (with 32-64 channels a tile has only 4-8 chunks:
// without this the TMA fill and the store drain of every tile were exposed).
__global__ void __launch_bounds__(kWThreads, 1) correlation_md4_tma64_kernel(
    const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, float* __restrict__ out, int C,
    int H, int W, float divisor, int legacy, int vec_store, int tiles_x, int tiles_y, int ntiles)
{
    pdl_enter();
    extern __shared__ __align__(128) unsigned char smem_bytes[];
    float* stage_mem = reinterpret_cast<float*>(smem_bytes);
    __shared__ __align__(8) unsigned long long bars[2 * kStages];  // full[0..2], empty[0..2]

    const int tid = threadIdx.x;
    const int nchunks = (C + kKC - 1) / kKC;
    const unsigned bar0 = smem_u32(bars);
    // my tiles: blockIdx.x, blockIdx.x + gridDim.x, ...
    const int my_tiles = (ntiles - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
    const int total = my_tiles * nchunks;   // chunks this CTA consumes, numbered G = 0 .. total-1

    if (tid == 0) {
#pragma unroll
        for (int i = 0; i < kStages; ++i) {
            mbar_init(bar0 + 8 * i, 1);
            mbar_init(bar0 + 8 * (kStages + i), kWThreads / 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    auto tile_coords = [&](int it, int& w0, int& h0, int& n) {
        const int t = blockIdx.x + it * gridDim.x;
        const int bx = t % tiles_x;
        const int rest = t / tiles_x;
        w0 = bx * kWTW;
        h0 = (rest % tiles_y) * kTH;
        n = rest / tiles_y;
    };
    auto issue = [&](int G) {  // thread 0 only: TMA for global chunk G into stage G % kStages
        int w0, h0, n;
        tile_coords(G / nchunks, w0, h0, n);
        const int j = G % nchunks;
        const int s = G % kStages;
        const unsigned full = bar0 + 8 * s;
        mbar_expect_tx(full, kWStageBytes);
        const unsigned dstA = smem_u32(stage_mem + s * kWStageFloats);
        const unsigned dstB = dstA + kKC * kWASize * sizeof(float);
        tma_load_4d(dstA, &mapA, full, w0, h0, j * kKC, n);
        tma_load_4d(dstB, &mapB, full, w0 - kMD, h0 - kMD, j * kKC, n);
    };
    if (tid == 0) {
        issue(0);
        if (total > 1)
            issue(1);
    }

    const int q = tid & 127;        // pixel quad inside the tile
    const int g = tid >> 7;         // vertical displacement group (4 warps each)
    const int r = q >> 4;           // tile row 0..7
    const int qc = (q & 15) * 4;    // first tile column of the quad
    float acc[3][kP][4];

    int G = 0;
    for (int it = 0; it < my_tiles; ++it) {
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < kP; ++j)
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    acc[i][j][k] = 0.0f;

        for (int j = 0; j < nchunks; ++j, ++G) {
            const int s = G % kStages;
            if (tid == 0 && G + 2 < total) {  // refill the stage consumed at iteration G-1 (possibly for the next tile)
                if (G >= 1)
                    mbar_wait(bar0 + 8 * (kStages + (G + 2) % kStages), ((G + 2) / kStages - 1) & 1);
                issue(G + 2);
            }
            mbar_wait(bar0 + 8 * s, (G / kStages) & 1);
            const float* sA = stage_mem + s * kWStageFloats + r * kWTW + qc;
            const float* sB = stage_mem + s * kWStageFloats + kKC * kWASize + (r + 3 * g) * kWBW + qc;
            // software pipeline: the next 128-bit in2 load is in flight while the current one feeds 16 FMAs
            float4 bn = *reinterpret_cast<const float4*>(sB);
            float4 an = *reinterpret_cast<const float4*>(sA);
            // (rolled: unrolling the 8 channels of the chunk 2x / 4x / 8x turns every offset into an immediate and removes
            // the compares and branches -- 36 % of the issued instructions -- and changes nothing: 96-100 us at
            // 32x544x960 either way, profiles/r2_correlation_unroll_events.txt.  ncu, r2_correlation_ncu.txt: issue 64 %,
            // FMA pipe 45 %, shared-memory wavefronts 57 %, stalls short_scoreboard / wait / long_scoreboard)
#pragma unroll 1
            for (int c = 0; c < kKC; ++c) {
                const float av[4] = {an.x, an.y, an.z, an.w};
                if (c + 1 < kKC)
                    an = *reinterpret_cast<const float4*>(sA + (c + 1) * kWASize);
#pragma unroll
                for (int i = 0; i < 3; ++i) {
#pragma unroll
                    for (int h = 0; h < 3; ++h) {
                        const float bv[4] = {bn.x, bn.y, bn.z, bn.w};
                        // address of the load after (c, i, h)
                        const int hn = (h + 1) % 3;
                        const int in_ = (h == 2) ? (i + 1) % 3 : i;
                        const bool wrap = (h == 2 && i == 2);
                        if (!wrap || c + 1 < kKC)
                            bn = *reinterpret_cast<const float4*>(
                                sB + (c + (wrap ? 1 : 0)) * kWBSize + in_ * kWBW + 4 * hn);
#pragma unroll
                        for (int mm = 0; mm < 4; ++mm) {
                            const int m = 4 * h + mm;
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const int jj = m - k;
                                if (jj >= 0 && jj < kP)
                                    acc[i][jj][k] = __fmaf_rn(av[k], bv[mm], acc[i][jj][k]);
                            }
                        }
                    }
                }
            }
            __syncwarp();
            if ((tid & 31) == 0)
                mbar_arrive(bar0 + 8 * (kStages + s));
        }
        int w0, h0, n;
        tile_coords(it, w0, h0, n);
        correlation_store(out, acc, n, g, h0 + r, w0 + qc, H, W, divisor, legacy, vec_store);
    }
}

// ---- shared-row kernel: the 64x8 persistent kernel with lane PAIRS that read the same in2 addresses ---------
// The kernel above is bound by the shared-memory crossbar: per channel a warp issues 10 LDS.128 = 40 wavefronts for
// 108 FFMAs = 27 SM-cycles of the FMA pipes.  Measured on the B200 (profiles/ubench/lds_merge.cu): an LDS.128 whose
// lanes 2i and 2i+1 read the SAME 16 bytes costs 2.16 cycles instead of 4.00 -- the two quarter-warps of a half-warp
// are served by one 128-byte wavefront.  (Sharing between other lane groupings -- half-warps, lanes l/3, l%8 -- is
// NOT merged: 4.00.)  So the lanes of a warp are laid out as 16 pixel quads x 2 MEMBERS that need the same in2
// rows at every load:
//   * a thread still owns 4 px x 9 x 3 displacements.  With pixel row p and vertical displacements dy = 3g-4+i it
//     reads in2 rows p + 3g - 4 + i.  The tile is SKEWED: for the staged in2 rows R0 .. R0+9 and k = 0..7, member
//     (g, k) owns pixel row p = R0 + k + 4 - 3g, so that all three groups g of one k read in2 rows R0+k .. R0+k+2
//     (in1 rows R0-2 .. R0+11 are staged: 14 x 80 + 10 x 72 floats per channel instead of 8 x 64 + 16 x 72);
//   * warps 0..7: lanes (2q, 2q+1) = groups g = 0, 1 of k = warp: every in2 load is pair-shared (2 wavefronts);
//   * warps 8..11: lanes (2q, 2q+1) = group g = 2 of k and of k+1; the odd lane walks its three rows rotated
//     (k+3, k+1, k+2), so two of its three rows coincide with the even lane's (k, k+1, k+2): 6 of 9 loads shared;
//   * in1 rows are 80 floats apart (64 + 8 columns either side): the two members' rows are 3 (or 1) rows apart =
//     16 banks mod 32, so the in1 load of a quarter-warp (4 quads of each row) is conflict-free.
// Per warp and channel: 4 + 18 (warps 0-7) or 4 + 24 (warps 8-11) wavefronts against 40 before.  Every output is the
// same sequential-over-c fp32 FMA chain as in the other kernels (bit-identical).
constexpr int kSAH = kTH + 6;                          // 14 in1 rows
constexpr int kSBH = kTH + 2;                          // 10 in2 rows
#ifndef VSC_CORR_SKC
#define VSC_CORR_SKC 8
#endif
#ifndef VSC_CORR_SSTAGES
#define VSC_CORR_SSTAGES 3
#endif
constexpr int kSKC = VSC_CORR_SKC;                     // channels per stage of the shared-row kernel
constexpr int kSStages = VSC_CORR_SSTAGES;             // ring depth
#ifndef VSC_CORR_UNROLL
#define VSC_CORR_UNROLL 1
#endif
#ifndef VSC_CORR_DEPTH
#define VSC_CORR_DEPTH 1
#endif
constexpr int kSU = VSC_CORR_UNROLL;                   // channels per trip of the channel loop
constexpr int kSD = VSC_CORR_DEPTH;                    // in2 loads in flight per thread
static_assert(kSKC % kSU == 0 && (9 * kSU) % kSD == 0 && kSD <= 9, "pipeline geometry");
#ifndef VSC_CORR_FFMA2
#define VSC_CORR_FFMA2 0
#endif
#ifndef VSC_CORR_ROWQUADS_DEFAULT
#define VSC_CORR_ROWQUADS_DEFAULT 1
#endif
#ifndef VSC_CORR_NARROW_DEFAULT
#define VSC_CORR_NARROW_DEFAULT 1
#endif

// packed pair of floats in a 64-bit register pair (sm_100a fma.rn.f32x2 -> FFMA2: two IEEE FMAs, one issue slot)
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float lo, float hi)
{
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ float lo2(f32x2 v) { return __uint_as_float(static_cast<unsigned>(v)); }
__device__ __forceinline__ float hi2(f32x2 v) { return __uint_as_float(static_cast<unsigned>(v >> 32)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c)
{
    f32x2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}

// TW = 64: one CTA of 12 warps per SM.  TW = 32: CTAs of 6 warps, TWO per SM (168 registers x 192 threads x 2 fits
// the register file): while one CTA stores its 108 accumulators per thread or waits at a chunk barrier, the other
// keeps the FMA pipes busy (ncu on the 64-wide kernel: 17 % of all warp time in the per-tile code, 16 % in the
// per-chunk code, all 12 warps in the same phase).
template <int TW>
__global__ void __launch_bounds__(TW * 6, TW == 64 ? 1 : 2) correlation_md4_share_kernel(
    const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, float* __restrict__ out, int C,
    int H, int W, float divisor, int legacy, int vec_store, int tiles_x, int tiles_y, int ntiles)
{
    constexpr int kThreads = TW * 6;                      // (TW / 4) quads x 8 rows x 3 groups
    constexpr int kAW = TW + 16, kBWid = TW + 2 * kMD;    // staged row lengths: in1 80 / 48, in2 72 / 40 floats
    constexpr int kASz = kSAH * kAW, kBSz = kSBH * kBWid;
    constexpr int kStageF = kSKC * (kASz + kBSz);
    constexpr unsigned kStageB = kStageF * sizeof(float);
    pdl_enter();
    extern __shared__ __align__(128) unsigned char smem_bytes[];
    float* stage_mem = reinterpret_cast<float*>(smem_bytes);
    __shared__ __align__(8) unsigned long long bars[2 * kSStages];  // full[0..2], empty[0..2]

    const int tid = threadIdx.x;
    const int nchunks = (C + kSKC - 1) / kSKC;
    const unsigned bar0 = smem_u32(bars);
    const int my_tiles = (ntiles - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
    const int total = my_tiles * nchunks;

    if (tid == 0) {
#pragma unroll
        for (int i = 0; i < kSStages; ++i) {
            mbar_init(bar0 + 8 * i, 1);
            mbar_init(bar0 + 8 * (kSStages + i), kThreads / 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // tile -> first column, first staged in2 row R0 = 8 ty - 4, image
    auto tile_coords = [&](int it, int& w0, int& R0, int& n) {
        const int t = blockIdx.x + it * gridDim.x;
        const int bx = t % tiles_x;
        const int rest = t / tiles_x;
        w0 = bx * TW;
        R0 = (rest % tiles_y) * kTH - kMD;
        n = rest / tiles_y;
    };
    auto issue = [&](int G) {
        int w0, R0, n;
        tile_coords(G / nchunks, w0, R0, n);
        const int j = G % nchunks;
        const int s = G % kSStages;
        const unsigned full = bar0 + 8 * s;
        mbar_expect_tx(full, kStageB);
        const unsigned dstA = smem_u32(stage_mem + s * kStageF);
        const unsigned dstB = dstA + kSKC * kASz * sizeof(float);
        tma_load_4d(dstA, &mapA, full, w0 - 8, R0 - 2, j * kSKC, n);
        tma_load_4d(dstB, &mapB, full, w0 - kMD, R0, j * kSKC, n);
    };
    if (tid == 0) {
        for (int G0 = 0; G0 < kSStages - 1 && G0 < total; ++G0)
            issue(G0);
    }

    // a tile has 12 member PAIRS (lanes 2q, 2q+1 over its TW / 4 quads): pairs 0..7 = groups 0, 1 of k = pair, pairs
    // 8..11 = group 2 of k = 2 (pair - 8) and k + 1.  A warp holds one pair (TW = 64) or two (TW = 32).
    constexpr int kQuads = TW / 4;
    const int pidx = tid >> 1;              // pair slot in the CTA
    const int pair = pidx / kQuads;         // 0..11
    const int qc = (pidx % kQuads) * 4;     // first tile column of the quad
    const int mem = tid & 1;
    // member -> group g, row index k, in2 row of slot 0 and of slot 1 (slot 2 = slot 1 + 1), rotation of the slots
    int g, k, brow0, brow1, rot;
    if (pair < 8) {
        g = mem;
        k = pair;
        brow0 = k;
        brow1 = k + 1;
        rot = 0;
    } else {
        g = 2;
        k = 2 * (pair - 8) + mem;
        brow0 = mem ? k + 2 : k;
        brow1 = mem ? k : k + 1;
        rot = mem ? 2 : 0;   // slot s holds vertical displacement 3g + (s + rot) % 3
    }
    const int prow = k + 6 - 3 * g;           // in1 tile row of the member's pixel row
    const int offA = prow * kAW + 8 + qc;
    const int offB0 = kSKC * kASz + brow0 * kBWid + qc;
    const int offB1 = kSKC * kASz + brow1 * kBWid + qc;
#if VSC_CORR_FFMA2
    // The loop is bound by instruction issue (ncu: the schedulers issue 97 % of the cycles spent in it; 108 FFMA + 10
    // LDS + 19 others per channel).  The FMAs are therefore issued as packed pairs: the accumulators of pixel kk are
    // paired over ADJACENT horizontal displacements so that both halves use in2 values b[m], b[m+1] with m EVEN, i.e.
    // an aligned register pair of the 128-bit load: even kk pair (jj, jj+1) = (0,1) (2,3) (4,5) (6,7) and keep jj = 8
    // single; odd kk pair (1,2) (3,4) (5,6) (7,8) and keep jj = 0 single.  The in1 value is duplicated into both
    // halves once per channel.  Per row 4 x (4 FFMA2 + 1 FFMA) = 20 issue slots instead of 36; every lane of a packed
    // FMA is the same IEEE operation, so results stay bit-identical.
    f32x2 accp[3][4][4];
    float accs[3][4];
#else
    float acc[3][kP][4];
#endif

    int G = 0;
    for (int it = 0; it < my_tiles; ++it) {
#if VSC_CORR_FFMA2
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
                for (int pp = 0; pp < 4; ++pp)
                    accp[i][kk][pp] = 0ull;
                accs[i][kk] = 0.0f;
            }
#else
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < kP; ++j)
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
                    acc[i][j][kk] = 0.0f;
#endif

        for (int j = 0; j < nchunks; ++j, ++G) {
            const int s = G % kSStages;
            if (tid == 0 && G + kSStages - 1 < total) {   // refill the stage consumed at chunk G-1
                if (G >= 1)
                    mbar_wait(bar0 + 8 * (kSStages + (G + kSStages - 1) % kSStages), ((G + kSStages - 1) / kSStages - 1) & 1);
                issue(G + kSStages - 1);
            }
            mbar_wait(bar0 + 8 * s, (G / kSStages) & 1);
            const float* st = stage_mem + s * kStageF;
            const float* sA = st + offA;
            const float* sB0 = st + offB0;
            const float* sB1 = st + offB1;
            // software pipeline: the in2 loads run kSD loads ahead of the FMAs they feed (a ring of kSD float4), the
            // channel loop is unrolled kSU times so that the ring slots are compile-time registers
            auto rowptr = [&](int i) { return i == 0 ? sB0 : sB1 + (i - 1) * kBWid; };
#if VSC_CORR_FFMA2
            ulonglong2 bq[kSD];
#pragma unroll
            for (int t = 0; t < kSD; ++t)
                bq[t] = *reinterpret_cast<const ulonglong2*>(rowptr(t / 3) + 4 * (t % 3));
            float4 an = *reinterpret_cast<const float4*>(sA);
#pragma unroll 1
            for (int c0 = 0; c0 < kSKC; c0 += kSU) {
#pragma unroll
                for (int u = 0; u < kSU; ++u) {
                    const float av[4] = {an.x, an.y, an.z, an.w};
                    const f32x2 a2[4] = {pk2(an.x, an.x), pk2(an.y, an.y), pk2(an.z, an.z), pk2(an.w, an.w)};
                    if (u + 1 < kSU || c0 + kSU < kSKC)
                        an = *reinterpret_cast<const float4*>(sA + (c0 + u + 1) * kASz);
#pragma unroll
                    for (int t = 0; t < 9; ++t) {
                        const int i = t / 3, h = t % 3;
                        const int slot = (u * 9 + t) % kSD;
                        const f32x2 bp[2] = {bq[slot].x, bq[slot].y};
                        const int un = u + (t + kSD) / 9, tn = (t + kSD) % 9;
                        if (un < kSU || c0 + kSU < kSKC)
                            bq[slot] = *reinterpret_cast<const ulonglong2*>(rowptr(tn / 3) + (c0 + un) * kBSz + 4 * (tn % 3));
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const int m = 4 * h + 2 * e;   // bp[e] = (b[m], b[m+1])
#pragma unroll
                            for (int kk = 0; kk < 4; ++kk) {
                                const int jj = m - kk;     // displacement of the low half
                                if (jj >= 0 && jj + 1 < kP) {
                                    const int pp = (kk & 1) ? (jj - 1) / 2 : jj / 2;
                                    accp[i][kk][pp] = fma2(a2[kk], bp[e], accp[i][kk][pp]);
                                } else if (jj == kP - 1) {
                                    accs[i][kk] = __fmaf_rn(av[kk], lo2(bp[e]), accs[i][kk]);
                                } else if (jj == -1) {
                                    accs[i][kk] = __fmaf_rn(av[kk], hi2(bp[e]), accs[i][kk]);
                                }
                            }
                        }
                    }
                }
            }
#else
            float4 bq[kSD];
#pragma unroll
            for (int t = 0; t < kSD; ++t)
                bq[t] = *reinterpret_cast<const float4*>(rowptr(t / 3) + 4 * (t % 3));
            float4 an = *reinterpret_cast<const float4*>(sA);
#pragma unroll 1
            for (int c0 = 0; c0 < kSKC; c0 += kSU) {
#pragma unroll
                for (int u = 0; u < kSU; ++u) {
                    const float av[4] = {an.x, an.y, an.z, an.w};
                    if (u + 1 < kSU || c0 + kSU < kSKC)
                        an = *reinterpret_cast<const float4*>(sA + (c0 + u + 1) * kASz);
#pragma unroll
                    for (int t = 0; t < 9; ++t) {
                        const int i = t / 3, h = t % 3;
                        const int slot = (u * 9 + t) % kSD;
                        const float bv[4] = {bq[slot].x, bq[slot].y, bq[slot].z, bq[slot].w};
                        // the load kSD steps ahead: channel c0 + un, row i' = tn / 3, quad tn % 3
                        const int un = u + (t + kSD) / 9, tn = (t + kSD) % 9;
                        if (un < kSU || c0 + kSU < kSKC)
                            bq[slot] = *reinterpret_cast<const float4*>(rowptr(tn / 3) + (c0 + un) * kBSz + 4 * (tn % 3));
#pragma unroll
                        for (int mm = 0; mm < 4; ++mm) {
                            const int m = 4 * h + mm;
#pragma unroll
                            for (int kk = 0; kk < 4; ++kk) {
                                const int jj = m - kk;
                                if (jj >= 0 && jj < kP)
                                    acc[i][jj][kk] = __fmaf_rn(av[kk], bv[mm], acc[i][jj][kk]);
                            }
                        }
                    }
                }
            }
#endif
            __syncwarp();
            if ((tid & 31) == 0)
                mbar_arrive(bar0 + 8 * (kSStages + s));
        }
        int w0, R0, n;
        tile_coords(it, w0, R0, n);
#if VSC_CORR_FFMA2
        float acc[3][kP][4];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
#pragma unroll
                for (int jj = 0; jj < kP; ++jj) {
                    const int single = (kk & 1) ? 0 : kP - 1;
                    const int off = (kk & 1) ? jj - 1 : jj;   // position among the paired displacements
                    acc[i][jj][kk] = jj == single ? accs[i][kk]
                        : (off & 1) ? hi2(accp[i][kk][off / 2]) : lo2(accp[i][kk][off / 2]);
                }
#endif
        correlation_store(out, acc, n, g, R0 - 2 + prow, w0 + qc, H, W, divisor, legacy, vec_store, rot);
    }
}

// ---- plain-load stager for shapes TMA cannot address (W % 4 != 0 or unaligned bases) ----------------
__global__ void __launch_bounds__(kConsumers) correlation_md4_ld_kernel(const float* __restrict__ in1,
    const float* __restrict__ in2, float* __restrict__ out, int C, int H, int W, float divisor, int legacy,
    int vec_store)
{
    __shared__ __align__(16) float sA[kKC * kASize];
    __shared__ __align__(16) float sB[kKC * kBSize];
    const int tid = threadIdx.x;
    const int q = tid & 63, g = tid >> 6, r = q >> 3, qc = (q & 7) * 4;
    const int w0 = blockIdx.x * kTW, h0 = blockIdx.y * kTH, n = blockIdx.z;
    const size_t HW = static_cast<size_t>(H) * W;
    const float* a_img = in1 + static_cast<size_t>(n) * C * HW;
    const float* b_img = in2 + static_cast<size_t>(n) * C * HW;
    float acc[3][kP][4];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < kP; ++j)
#pragma unroll
            for (int k = 0; k < 4; ++k)
                acc[i][j][k] = 0.0f;

    constexpr int kNA = (kKC * kASize + kConsumers - 1) / kConsumers;  // 11 loads per thread
    constexpr int kNB = (kKC * kBSize + kConsumers - 1) / kConsumers;  // 27
    for (int cbase = 0; cbase < C; cbase += kKC) {
        // issue every load of the stage before the first store: one memory round trip per stage
        float va[kNA], vb[kNB];
#pragma unroll
        for (int j = 0; j < kNA; ++j) {
            const int i = tid + j * kConsumers;
            const int c = i / kASize, rem = i % kASize;
            const int y = h0 + rem / kTW, x = w0 + rem % kTW;
            const bool ok = i < kKC * kASize && cbase + c < C && y < H && x < W;
            va[j] = ok ? __ldg(a_img + (cbase + c) * HW + static_cast<size_t>(y) * W + x) : 0.0f;
        }
#pragma unroll
        for (int j = 0; j < kNB; ++j) {
            const int i = tid + j * kConsumers;
            const int c = i / kBSize, rem = i % kBSize;
            const int y = h0 - kMD + rem / kBW, x = w0 - kMD + rem % kBW;
            const bool ok = i < kKC * kBSize && cbase + c < C && y >= 0 && y < H && x >= 0 && x < W;
            vb[j] = ok ? __ldg(b_img + (cbase + c) * HW + static_cast<size_t>(y) * W + x) : 0.0f;
        }
        __syncthreads();  // previous stage fully consumed
#pragma unroll
        for (int j = 0; j < kNA; ++j)
            if (tid + j * kConsumers < kKC * kASize)
                sA[tid + j * kConsumers] = va[j];
#pragma unroll
        for (int j = 0; j < kNB; ++j)
            if (tid + j * kConsumers < kKC * kBSize)
                sB[tid + j * kConsumers] = vb[j];
        __syncthreads();
        correlate_chunk(sA, sB, r, qc, g, acc);
    }
    correlation_store(out, acc, n, g, h0 + r, w0 + qc, H, W, divisor, legacy, vec_store);
}

// any max_displacement: one thread per output value, sequential fp32 sum over c
__global__ void __launch_bounds__(256) correlation_generic_kernel(const float* __restrict__ in1,
    const float* __restrict__ in2, float* __restrict__ out, int C, int H, int W, int md, float scale, int use_div)
{
    const int P = 2 * md + 1;
    const size_t HW = static_cast<size_t>(H) * W;
    const size_t per_n = static_cast<size_t>(P) * P * HW;
    const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= per_n)
        return;
    const int n = blockIdx.y;
    const int w = static_cast<int>(i % W);
    const int h = static_cast<int>((i / W) % H);
    const int pw = static_cast<int>((i / HW) % P);
    const int ph = static_cast<int>(i / (HW * P));
    const int h2 = h + ph - md, w2 = w + pw - md;
    float acc = 0.0f;
    if (h2 >= 0 && h2 < H && w2 >= 0 && w2 < W) {
        const float* a = in1 + static_cast<size_t>(n) * C * HW + static_cast<size_t>(h) * W + w;
        const float* b = in2 + static_cast<size_t>(n) * C * HW + static_cast<size_t>(h2) * W + w2;
#pragma unroll 8
        for (int c = 0; c < C; ++c)
            acc = __fmaf_rn(__ldg(a + c * HW), __ldg(b + c * HW), acc);
    }
    out[static_cast<size_t>(n) * per_n + i] = use_div ? acc / scale : acc;
}

// Small maps (the coarse PWC-Net levels) are pure latency: a few thousand pixels, each output a serial walk over
// up to 196 channels of DRAM-cold operands, and the tiled kernels above get only a handful of CTAs.  Here a
// thread owns one (ph, h, w) task = the 9 horizontal displacements of one pixel and one vertical displacement:
// per channel it loads in1 once and 9 neighbouring in2 values (1.1 loads per FMA instead of 2; lanes are
// consecutive w, so every load is a contiguous row segment).  The 8 warps of a CTA share 32 tasks and each walks
// one eighth of the channels (2-4 load batches instead of 25); the 8 partial sums of a value are combined in
// warp order through shared memory (deterministic).
// (kSplitC = 16 for the very smallest maps, whose grids do not even fill the SMs at 8.)
template <int kSplitC>
__global__ void __launch_bounds__(32 * kSplitC) correlation_md4_rows_kernel(const float* __restrict__ in1,
    const float* __restrict__ in2, float* __restrict__ out, int C, int H, int W, float scale, int use_div)
{
    pdl_enter();
    __shared__ float part[kSplitC][kP][32];
    __shared__ int obase[32];
    const int lane = threadIdx.x & 31, slice = threadIdx.x >> 5;
    const unsigned HW = static_cast<unsigned>(H) * W;
    const unsigned tasks = kP * HW;                       // per n: (ph, h, w), w fastest
    const unsigned i = blockIdx.x * 32u + lane;
    const int n = blockIdx.y;
    float acc[kP];
#pragma unroll
    for (int j = 0; j < kP; ++j)
        acc[j] = 0.0f;
    int ob = -1;
    if (i < tasks) {
        const unsigned ph = i / HW;
        const unsigned hw = i - ph * HW;
        const int h = static_cast<int>(hw / W);
        const int w = static_cast<int>(hw - static_cast<unsigned>(h) * W);
        ob = static_cast<int>(ph * kP * HW + hw);        // + pw * HW
        const int h2 = h + static_cast<int>(ph) - kMD;
        if (h2 >= 0 && h2 < H) {
            unsigned colmask = 0;
#pragma unroll
            for (int j = 0; j < kP; ++j)
                colmask |= (w + j - kMD >= 0 && w + j - kMD < W) ? (1u << j) : 0u;
            const int cb = static_cast<int>(static_cast<long long>(C) * slice / kSplitC);
            const int ce = static_cast<int>(static_cast<long long>(C) * (slice + 1) / kSplitC);
            const float* a = in1 + (static_cast<size_t>(n) * C + cb) * HW + hw;
            // column w-4 of row h2; lanes whose left columns fall outside the row never dereference them
            const float* b = in2 + (static_cast<size_t>(n) * C + cb) * HW + static_cast<size_t>(h2) * W + w - kMD;
#pragma unroll 4
            for (int c = 0; c < ce - cb; ++c, a += HW, b += HW) {
                const float av = __ldg(a);
#pragma unroll
                for (int j = 0; j < kP; ++j) {
                    const float bv = (colmask >> j) & 1u ? __ldg(b + j) : 0.0f;
                    acc[j] = __fmaf_rn(av, bv, acc[j]);
                }
            }
        }
    }
#pragma unroll
    for (int j = 0; j < kP; ++j)
        part[slice][j][lane] = acc[j];
    if (slice == 0)
        obase[lane] = ob;
    __syncthreads();
    // 9 x 32 values per CTA: warp k sums displacement k (and k + kSplitC); 128-byte row stores
    for (int j = slice; j < kP; j += kSplitC) {
        const int o = obase[lane];
        if (o >= 0) {
            float sum = part[0][j][lane];
#pragma unroll
            for (int k = 1; k < kSplitC; ++k)
                sum += part[k][j][lane];
            out[static_cast<size_t>(n) * kP * tasks + static_cast<unsigned>(o) + static_cast<size_t>(j) * HW]
                = use_div ? sum / scale : sum;
        }
    }
}

// The same channel-split scheme with a register tile for maps whose rows are whole 16-byte quads (W % 4 == 0, aligned
// tensors): a thread owns one (ph, h, pixel QUAD) task = 4 px x 9 horizontal displacements = 36 accumulators and loads
// per channel one 128-bit quad of in1 and the three quads of in2 around it -- 4 load instructions per 36 FMAs instead
// of the 40 of the kernel above, which is bound by the number of load instructions it issues (1.1 per FMA).  Quads
// lie entirely inside or outside a row, so the zero padding is two predicates.  Partial sums of the channel slices
// are combined in slice order through shared memory; results leave as 128-bit stores.
template <int kSplitC>
__global__ void __launch_bounds__(32 * kSplitC) correlation_md4_rowquads_kernel(const float* __restrict__ in1,
    const float* __restrict__ in2, float* __restrict__ out, int C, int H, int W, float scale, int use_div)
{
    pdl_enter();
    __shared__ __align__(16) float part[kSplitC][kP][32][4];
    __shared__ int obase[32];
    const int lane = threadIdx.x & 31, slice = threadIdx.x >> 5;
    const unsigned HW = static_cast<unsigned>(H) * W;
    const unsigned Wq = static_cast<unsigned>(W) / 4, HWq = static_cast<unsigned>(H) * Wq;
    const unsigned tasks = kP * HWq;                      // per n: (ph, h, quad), quad fastest
    const unsigned i = blockIdx.x * 32u + lane;
    const int n = blockIdx.y;
    float acc[kP][4];
#pragma unroll
    for (int j = 0; j < kP; ++j)
#pragma unroll
        for (int k = 0; k < 4; ++k)
            acc[j][k] = 0.0f;
    int ob = -1;
    if (i < tasks) {
        const unsigned ph = i / HWq;
        const unsigned r = i - ph * HWq;
        const int h = static_cast<int>(r / Wq);
        const int w = 4 * static_cast<int>(r - static_cast<unsigned>(h) * Wq);
        ob = static_cast<int>(ph * kP * HW + static_cast<unsigned>(h) * W + w);   // + pw * HW
        const int h2 = h + static_cast<int>(ph) - kMD;
        if (h2 >= 0 && h2 < H) {
            const bool okL = w >= 4, okR = w + 4 < W;
            const int cb = static_cast<int>(static_cast<long long>(C) * slice / kSplitC);
            const int ce = static_cast<int>(static_cast<long long>(C) * (slice + 1) / kSplitC);
            const float* a = in1 + (static_cast<size_t>(n) * C + cb) * HW + static_cast<size_t>(h) * W + w;
            const float* b = in2 + (static_cast<size_t>(n) * C + cb) * HW + static_cast<size_t>(h2) * W + w;
            const float4 zero = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
#pragma unroll 2
            for (int c = 0; c < ce - cb; ++c, a += HW, b += HW) {
                const float4 a4 = __ldg(reinterpret_cast<const float4*>(a));
                const float4 bl = okL ? __ldg(reinterpret_cast<const float4*>(b - 4)) : zero;
                const float4 bm = __ldg(reinterpret_cast<const float4*>(b));
                const float4 br = okR ? __ldg(reinterpret_cast<const float4*>(b + 4)) : zero;
                const float av[4] = {a4.x, a4.y, a4.z, a4.w};
                const float bv[12] = {bl.x, bl.y, bl.z, bl.w, bm.x, bm.y, bm.z, bm.w, br.x, br.y, br.z, br.w};
#pragma unroll
                for (int m = 0; m < 12; ++m)
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int j = m - k;
                        if (j >= 0 && j < kP)
                            acc[j][k] = __fmaf_rn(av[k], bv[m], acc[j][k]);
                    }
            }
        }
    }
#pragma unroll
    for (int j = 0; j < kP; ++j)
        *reinterpret_cast<float4*>(part[slice][j][lane]) = make_float4(acc[j][0], acc[j][1], acc[j][2], acc[j][3]);
    if (slice == 0)
        obase[lane] = ob;
    __syncthreads();
    // 9 x 32 quads per CTA: warp k sums displacement k (and k + kSplitC); 512-byte row stores
    for (int j = slice; j < kP; j += kSplitC) {
        const int o = obase[lane];
        if (o >= 0) {
            float4 sum = *reinterpret_cast<const float4*>(part[0][j][lane]);
#pragma unroll
            for (int k = 1; k < kSplitC; ++k) {
                const float4 p = *reinterpret_cast<const float4*>(part[k][j][lane]);
                sum.x += p.x;
                sum.y += p.y;
                sum.z += p.z;
                sum.w += p.w;
            }
            if (use_div) {
                sum.x = sum.x / scale;
                sum.y = sum.y / scale;
                sum.z = sum.z / scale;
                sum.w = sum.w / scale;
            }
            *reinterpret_cast<float4*>(out + static_cast<size_t>(n) * kP * kP * HW + static_cast<unsigned>(o)
                + static_cast<size_t>(j) * HW) = sum;
        }
    }
}

// ---- host side: tensor maps ---------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled()
{
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess
            && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
        else
            (void)cudaGetLastError();
    }
    return fn;
}

// [N][C][H][W] fp32, box [1][KC][bh][bw]; out-of-bounds elements read as zero
static bool make_map(CUtensorMap* m, const float* base, int N, int C, int H, int W, int bw, int bh, int kc = kKC)
{
    EncodeTiledFn enc = encode_tiled();
    if (!enc)
        return false;
    const cuuint64_t dims[4] = {static_cast<cuuint64_t>(W), static_cast<cuuint64_t>(H), static_cast<cuuint64_t>(C),
        static_cast<cuuint64_t>(N)};
    const cuuint64_t strides[3] = {static_cast<cuuint64_t>(W) * 4, static_cast<cuuint64_t>(W) * H * 4,
        static_cast<cuuint64_t>(W) * H * C * 4};
    const cuuint32_t box[4] = {static_cast<cuuint32_t>(bw), static_cast<cuuint32_t>(bh), static_cast<cuuint32_t>(kc), 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(base), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)
        == CUDA_SUCCESS;
}

std::atomic<int> g_corr_mode = 0;  // 0 auto, 1 plain-load stager, 2 TMA 32x8 tiles, 3 TMA 64x8 tiles, 4 channel-split (tests),
                      // 5 / 6 TMA shared-row tiles, 64x8 (one CTA per SM) / 32x8 (two), 7 channel-split quads

}  // namespace vsc

extern "C" int vsc_set_correlation_mode(int mode)
{
    if (mode < 0 || mode > 7)
        return VSC_E_INVALID;
    vsc::g_corr_mode = mode;
    return VSC_OK;
}

extern "C" int vsc_correlation_f32(const float* in1, const float* in2, float* out, int N, int C, int H, int W,
    int max_displacement, int legacy, vsc_stream_t stream)
{
    using namespace vsc;
    if (!in1 || !in2 || !out || N <= 0 || C <= 0 || H <= 0 || W <= 0 || max_displacement < 0 || N > 65535)
        return VSC_E_INVALID;
    if (static_cast<long long>(H) * W * C > 0x7fffffffLL)
        return VSC_E_INVALID;
    if (!aligned4(in1) || !aligned4(in2) || !aligned4(out))
        return VSC_E_ALIGN;
    cudaStream_t st = as_stream(stream);
    // Small maps (the coarse PWC-Net levels: 9x15 ... 36x60 at 1080p/2) give the tiled kernels a handful of CTAs
    // that each walk all channels serially (76 us for 196x9x15, two CTAs): they go to the channel-split rows
    // kernel, which spreads the work over 9*H*W/32 CTAs x 8 channel slices.  Its sums associate differently
    // from the tiled kernels' (8 partial sums), well inside the op's 1e-4 tolerance.
    const bool small_map = max_displacement == kMD
        && (g_corr_mode == 4 || g_corr_mode == 7 || (g_corr_mode == 0 && static_cast<long long>(H) * W <= 12288));
    if (max_displacement == kMD && !small_map) {
        const int vec = (W % 4 == 0) && aligned16(out);
        const dim3 grid(cdiv(W, kTW), cdiv(H, kTH), N);
        if (grid.y > 65535)
            return VSC_E_INVALID;
        const float divisor = static_cast<float>(C);
        const bool tma_ok = g_corr_mode != 1 && (W % 4 == 0) && aligned16(in1) && aligned16(in2);
        // wide tile when it is not mostly padding: image at least 1.5 tiles wide (mode 2 / 3 force 32 / 64)
        const bool wide = g_corr_mode == 3 || g_corr_mode == 5 || g_corr_mode == 6 || (g_corr_mode == 0 && W >= 96);
        // shared-row lane layout (lane pairs read the same in2 rows): the default wide kernel; mode 3 keeps the
        // one-row-per-half-warp layout for A/B runs
        if (tma_ok && wide && g_corr_mode != 3) {
            const bool narrow = g_corr_mode == 6 || (g_corr_mode == 0 && VSC_CORR_NARROW_DEFAULT);
            const int TW = narrow ? 32 : 64;
            CUtensorMap mapA, mapB;
            if (make_map(&mapA, in1, N, C, H, W, TW + 16, kSAH, kSKC) && make_map(&mapB, in2, N, C, H, W, TW + 2 * kMD, kSBH, kSKC)) {
                const size_t smem = static_cast<size_t>(kSStages) * kSKC * (kSAH * (TW + 16) + kSBH * (TW + 2 * kMD)) * sizeof(float);
                static unsigned long long configured64 = 0, configured32 = 0;
                if (const int e = narrow ? ensure_dynamic_smem(correlation_md4_share_kernel<32>, smem, true, configured32)
                                         : ensure_dynamic_smem(correlation_md4_share_kernel<64>, smem, false, configured64))
                    return e;
                const int tiles_x = static_cast<int>(cdiv(W, TW));
                const int tiles_y = (H > 2 ? (H - 2 + kTH - 1) / kTH : 0) + 1;   // skewed tiles: one more row of tiles
                const long long ntiles = static_cast<long long>(tiles_x) * tiles_y * N;
                if (ntiles > 0x7fffffffLL)
                    return VSC_E_INVALID;
                const long long slots = static_cast<long long>(sm_count()) * (narrow ? 2 : 1);
                const unsigned ctas = static_cast<unsigned>(ntiles < slots ? ntiles : slots);
                const int rc = narrow
                    ? launch_pdl(correlation_md4_share_kernel<32>, dim3(ctas), dim3(192), smem, st, mapA, mapB, out, C, H, W,
                          divisor, legacy ? 1 : 0, vec, tiles_x, tiles_y, static_cast<int>(ntiles))
                    : launch_pdl(correlation_md4_share_kernel<64>, dim3(ctas), dim3(384), smem, st, mapA, mapB, out, C, H, W,
                          divisor, legacy ? 1 : 0, vec, tiles_x, tiles_y, static_cast<int>(ntiles));
                count_launch();
                return rc ? rc : launch_status();
            }
        }
        if (tma_ok && wide) {
            CUtensorMap mapA, mapB;
            if (make_map(&mapA, in1, N, C, H, W, kWTW, kTH) && make_map(&mapB, in2, N, C, H, W, kWBW, kBH)) {
                constexpr size_t smem = kStages * kWStageBytes;
                static unsigned long long configured = 0;
                if (const int e = ensure_dynamic_smem(correlation_md4_tma64_kernel, smem, false, configured))
                    return e;
                const int tiles_x = static_cast<int>(cdiv(W, kWTW)), tiles_y = static_cast<int>(cdiv(H, kTH));
                const long long ntiles = static_cast<long long>(tiles_x) * tiles_y * N;
                if (ntiles > 0x7fffffffLL)
                    return VSC_E_INVALID;
                const unsigned ctas = static_cast<unsigned>(ntiles < sm_count() ? ntiles : sm_count());
                const int rc = launch_pdl(correlation_md4_tma64_kernel, dim3(ctas), dim3(kWThreads), smem, st, mapA, mapB,
                    out, C, H, W, divisor, legacy ? 1 : 0, vec, tiles_x, tiles_y, static_cast<int>(ntiles));
                count_launch();
                return rc ? rc : launch_status();
            }
        }
        if (tma_ok) {
            CUtensorMap mapA, mapB;
            if (make_map(&mapA, in1, N, C, H, W, kTW, kTH) && make_map(&mapB, in2, N, C, H, W, kBW, kBH)) {
                constexpr size_t smem = kStages * kStageBytes;
                static unsigned long long configured = 0;  // two 84 KB CTAs per SM need the max-shared carve-out
                if (const int e = ensure_dynamic_smem(correlation_md4_tma_kernel, smem, true, configured))
                    return e;
                correlation_md4_tma_kernel<<<grid, kConsumers + 32, smem, st>>>(mapA, mapB, out, C, H, W, divisor,
                    legacy ? 1 : 0, vec);
                count_launch();
                return launch_status();
            }
        }
        correlation_md4_ld_kernel<<<grid, kConsumers, 0, st>>>(in1, in2, out, C, H, W, divisor, legacy ? 1 : 0, vec);
        count_launch();
        return launch_status();
    }
    const int P = 2 * max_displacement + 1;
    const size_t per_n = static_cast<size_t>(P) * P * H * W;
    // quad form of the rows kernel: rows of whole 16-byte quads (mode 7 forces it, mode 4 forces the scalar form)
    if (max_displacement == kMD && static_cast<long long>(H) * W * kP * kP < 0x7fffffffLL && W % 4 == 0 && aligned16(in1)
        && aligned16(in2) && aligned16(out) && (g_corr_mode == 7 || (g_corr_mode == 0 && VSC_CORR_ROWQUADS_DEFAULT))) {
        const dim3 grids(cdiv(static_cast<long long>(kP) * H * (W / 4), 32), N);
        const float fC = static_cast<float>(C);
        const int lg = legacy ? 1 : 0;
        // about 16 channels per warp, at most 8 slices (36.9 KB of partial sums)
        const int rc = C >= 96
            ? launch_pdl(correlation_md4_rowquads_kernel<8>, grids, dim3(32 * 8), 0, st, in1, in2, out, C, H, W, fC, lg)
            : launch_pdl(correlation_md4_rowquads_kernel<4>, grids, dim3(32 * 4), 0, st, in1, in2, out, C, H, W, fC, lg);
        if (rc)
            return rc;
        count_launch();
        return launch_status();
    }
    if (max_displacement == kMD && static_cast<long long>(H) * W * kP * kP < 0x7fffffffLL) {
        const dim3 grids(cdiv(static_cast<long long>(kP) * H * W, 32), N);
        // channel slices per CTA: about 16 channels per warp (shorter slices drown in the per-task set-up and the
        // shared-memory reduction, longer ones serialise the DRAM round trips), 16 slices when the grid alone
        // cannot fill the SMs
        const bool tiny = static_cast<long long>(grids.x) * N < 2LL * sm_count() && C >= 64;
        const float fC = static_cast<float>(C);
        const int lg = legacy ? 1 : 0;
        int rc;
        if (tiny || C >= 160)
            rc = launch_pdl(correlation_md4_rows_kernel<16>, grids, dim3(32 * 16), 0, st, in1, in2, out, C, H, W, fC, lg);
        else if (C >= 96)
            rc = launch_pdl(correlation_md4_rows_kernel<8>, grids, dim3(32 * 8), 0, st, in1, in2, out, C, H, W, fC, lg);
        else
            rc = launch_pdl(correlation_md4_rows_kernel<4>, grids, dim3(32 * 4), 0, st, in1, in2, out, C, H, W, fC, lg);
        if (rc)
            return rc;
        count_launch();
        return launch_status();
    }
    const dim3 grid(cdiv(static_cast<long long>(per_n), 256), N);
    correlation_generic_kernel<<<grid, 256, 0, st>>>(in1, in2, out, C, H, W, max_displacement, static_cast<float>(C),
        legacy ? 1 : 0);
    count_launch();
    return launch_status();
}
