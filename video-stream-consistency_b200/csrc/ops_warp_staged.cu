// custom::Warp, TMA-staged kernel for the large feature maps (reference: warp_cuda.cu:29-98, warp.cc:71-134).
//
// The gather kernels of ops_warp.cu make every thread wait for two dependent memory round trips (flow -> corner
// addresses -> corners) per channel batch and spend 8 of their 27.75 instructions per value on forming addresses:
// 0.47-0.57 of the HBM peak.  Here the loads do not depend on per-thread state.  A CTA owns a 64x8 pixel tile
// (256 threads, two pixels each, rows r and r+4).  Its threads compute their corners and weights once (ops_warp.cuh:
// the same corners, weights, mask decision and accumulation order as every other kernel, so results are
// bit-identical), reduce the bounding box of all corners the tile will read and, when that box fits kBX x kBY floats
// (flow varying by up to ~4 px across the tile: what an optical-flow decoder level produces), ONE thread streams the
// box, kG channels at a time, into a kStages-deep shared-memory ring with cp.async.bulk.tensor.4d over
// [N][C][H][W] (out-of-bounds zero fill, mbarrier complete_tx), while all threads take their corners of the previous
// group from shared memory at constant offsets: per value 4 LDS + 7 FP + 1 STG.  Tiles whose corners are spread
// wider (scattered flow) gather directly, exactly like warp_nchw_kernel.
//
// The box origin is rounded down to a multiple of 4 floats (16 bytes): every row the TMA engine fetches then starts
// on a 16-byte boundary.  (Round 1's first version used unrounded origins and its copies never completed once a box
// started off a 16-byte boundary; the wait below is bounded and traps instead of hanging the device.)
//
// Needs W % 4 == 0 (tensor-map strides are multiples of 16 bytes), a 16-byte aligned input, W >= kBX, H >= kBY.
#include <cuda.h>

#include <climits>
#include <type_traits>

#include "ops_warp.cuh"

namespace vsc {

namespace {
constexpr int kTileW = 64, kTileH = 8;           // output pixels per CTA
constexpr int kThreads = 256;                    // two pixels per thread: rows ty and ty + 4
constexpr int kBX = 72, kBY = 12;                // staged box (floats x rows): 64 + 1 (xR) + 3 (alignment) + 4 slack
#ifndef VSC_WARP_G
#define VSC_WARP_G 4
#endif
constexpr int kG = VSC_WARP_G;                   // channels per stage
#ifndef VSC_WARP_STAGES
#define VSC_WARP_STAGES 3
#endif
constexpr int kStages = VSC_WARP_STAGES;
constexpr int kPlane = kBY * kBX;                // floats of one channel of a staged box
constexpr int kStageFloats = kG * kPlane;        // 3456 floats = 13.5 KB (a multiple of 128 bytes)
constexpr unsigned kStageBytes = kStageFloats * sizeof(float);
constexpr size_t kStagedSmem = kStages * kStageBytes + 64 + 8 * 4 * sizeof(int);
static_assert(kStageBytes % 128 == 0, "TMA destinations are 128-byte aligned");

__device__ __forceinline__ unsigned w_smem_u32(const void* p) { return static_cast<unsigned>(__cvta_generic_to_shared(p)); }

struct StagedPix {
    WarpTap t;
    int s00, s10, s01, s11;   // offsets of the corners inside one channel of the staged box
};
}  // namespace

__global__ void __launch_bounds__(kThreads, 4) warp_nchw_staged_kernel(const __grid_constant__ CUtensorMap map,
    const float* __restrict__ in, const float* __restrict__ flow, float* __restrict__ out, int C, int H, int W,
    int chunk, int nchunk)
{
    pdl_enter();
    extern __shared__ __align__(128) unsigned char staged_smem[];
    float* stage = reinterpret_cast<float*>(staged_smem);
    const unsigned bar0 = w_smem_u32(staged_smem + kStages * kStageBytes);   // kStages mbarriers of 8 bytes
    int* red = reinterpret_cast<int*>(staged_smem + kStages * kStageBytes + 64);   // [8 warps][4]

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int x = blockIdx.x * kTileW + (tid & (kTileW - 1));
    const int ty = tid / kTileW;                                  // 0..3
    const int HW = H * W;
    const int n = blockIdx.z / nchunk;
    const int c0 = (blockIdx.z - n * nchunk) * chunk;
    const int c1 = min(C, c0 + chunk);

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < kStages; ++s)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar0 + 8u * s), "r"(1u) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }

    StagedPix px[2];
    bool live[2];
    int xl[2], yt[2], pofs[2];
    int minc = INT_MAX, maxc = INT_MIN, minr = INT_MAX, maxr = INT_MIN;
    bool all_valid = true;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int y = blockIdx.y * kTileH + ty + 4 * i;
        live[i] = x < W && y < H;
        pofs[i] = live[i] ? y * W + x : 0;
        const float* fl = flow + static_cast<size_t>(n) * 2 * HW + pofs[i];
        const float fu = ldg_stream(fl), fv = ldg_stream(fl + HW);
        px[i].t = warp_setup(live[i] ? x : 0, live[i] ? y : 0, fu, fv, W, H);
        if (!live[i])
            px[i].t.valid = 0u;
        // integer corner coordinates: a readable corner implies finite xL / yT in [-1, W-1] / [-1, H-1]
        xl[i] = 0;
        yt[i] = 0;
        if (px[i].t.valid) {
            xl[i] = static_cast<int>(floorf(static_cast<float>(x) + fu));
            yt[i] = static_cast<int>(floorf(static_cast<float>(y) + fv));
        }
        const unsigned v = px[i].t.valid;
        if (v & 5u) { minc = min(minc, xl[i]); maxc = max(maxc, xl[i]); }           // corners 00 / 01: column xL
        if (v & 10u) { minc = min(minc, xl[i] + 1); maxc = max(maxc, xl[i] + 1); }  // corners 10 / 11: column xR
        if (v & 3u) { minr = min(minr, yt[i]); maxr = max(maxr, yt[i]); }           // corners 00 / 10: row yT
        if (v & 12u) { minr = min(minr, yt[i] + 1); maxr = max(maxr, yt[i] + 1); }  // corners 01 / 11: row yB
        all_valid = all_valid && v == 15u;
    }
    // bounding box of every corner this tile reads
    minc = __reduce_min_sync(0xffffffffu, minc);
    maxc = __reduce_max_sync(0xffffffffu, maxc);
    minr = __reduce_min_sync(0xffffffffu, minr);
    maxr = __reduce_max_sync(0xffffffffu, maxr);
    if (lane == 0) {
        red[warp * 4 + 0] = minc;
        red[warp * 4 + 1] = maxc;
        red[warp * 4 + 2] = minr;
        red[warp * 4 + 3] = maxr;
    }
    const int tile_all_valid = __syncthreads_and(all_valid ? 1 : 0);   // also publishes `red` and the mbarrier init
    {
        const int w8 = lane & 7;
        minc = __reduce_min_sync(0xffffffffu, red[w8 * 4 + 0]);
        maxc = __reduce_max_sync(0xffffffffu, red[w8 * 4 + 1]);
        minr = __reduce_min_sync(0xffffffffu, red[w8 * 4 + 2]);
        maxr = __reduce_max_sync(0xffffffffu, red[w8 * 4 + 3]);
    }
    const bool any = maxc >= minc && maxr >= minr;
    if (any)
        minc &= ~3;   // (two's complement: rounds negative origins down as well)
    const bool staged = any && (maxc - minc) < kBX && (maxr - minr) < kBY;   // CTA-uniform

    float* op0 = out + (static_cast<size_t>(n) * C + c0) * HW;
    if (!staged) {
        // scattered flow (or a tile without a single readable corner): direct gathers, as warp_nchw_kernel
        // (both pixels of a thread per channel pair: 16 independent gathers in flight, as in warp_nchw_kernel)
        const float* ip = in + (static_cast<size_t>(n) * C + c0) * HW;
        float* opa = op0 + pofs[0];
        float* opb = op0 + pofs[1];
        int c = c0;
        for (; c + 2 <= c1; c += 2, ip += 2 * static_cast<size_t>(HW), opa += 2 * static_cast<size_t>(HW),
             opb += 2 * static_cast<size_t>(HW)) {
            const float a0 = warp_sample(ip, px[0].t);
            const float a1 = warp_sample(ip + HW, px[0].t);
            const float b0 = warp_sample(ip, px[1].t);
            const float b1 = warp_sample(ip + HW, px[1].t);
            if (live[0]) {
                __stcs(opa, a0);
                __stcs(opa + HW, a1);
            }
            if (live[1]) {
                __stcs(opb, b0);
                __stcs(opb + HW, b1);
            }
        }
        if (c < c1) {
            const float a0 = warp_sample(ip, px[0].t);
            const float b0 = warp_sample(ip, px[1].t);
            if (live[0])
                __stcs(opa, a0);
            if (live[1])
                __stcs(opb, b0);
        }
        return;
    }

#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const unsigned v = px[i].t.valid;
        const int r0 = (yt[i] - minr) * kBX, cx = xl[i] - minc;
        px[i].s00 = (v & 1u) ? r0 + cx : 0;
        px[i].s10 = (v & 2u) ? r0 + cx + 1 : 0;
        px[i].s01 = (v & 4u) ? r0 + kBX + cx : 0;
        px[i].s11 = (v & 8u) ? r0 + kBX + cx + 1 : 0;
    }
    const int ngroups = (c1 - c0 + kG - 1) / kG;
    auto request = [&](int g) {   // thread 0: box of channels c0 + g*kG .. +kG-1 into stage g % kStages
        const unsigned bar = bar0 + 8u * (g % kStages);
        const unsigned dst = w_smem_u32(stage + (g % kStages) * kStageFloats);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(kStageBytes) : "memory");
        asm volatile(
            "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
            ::"r"(dst), "l"(reinterpret_cast<unsigned long long>(&map)), "r"(bar), "r"(minc), "r"(minr),
            "r"(c0 + g * kG), "r"(n)
            : "memory");
    };
    if (tid == 0)
        for (int g = 0; g < kStages && g < ngroups; ++g)
            request(g);

    // store pointers of my two pixels, advanced by one channel plane per value (forming base + j*HW per store costs
    // 10 instructions per value: 64-bit multiplies -- profiles/r2_warp_staged_ncu.txt)
    float* opx[2] = {op0 + pofs[0], op0 + pofs[1]};
    // FULL: all kG channels of the group exist (always, except for the last group when C % kG != 0): no per-channel
    // test -- even a uniform branch per channel ends the basic block and serialises the LDS -> FMA -> STG chains
    auto consume = [&](auto allv_tag, auto full_tag, const float* sb, int g) {
        constexpr bool ALLV = decltype(allv_tag)::value;
        constexpr bool FULL = decltype(full_tag)::value;
        const int cg = c0 + g * kG;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const WarpTap& t = px[i].t;
            float* op = opx[i];
#pragma unroll
            for (int j = 0; j < kG; ++j) {
                const float* sc = sb + j * kPlane;
                float a, b, c, d;
                if constexpr (ALLV) {
                    a = sc[px[i].s00];
                    b = sc[px[i].s10];
                    c = sc[px[i].s01];
                    d = sc[px[i].s11];
                } else {
                    // unused corners are not read: the box may hold non-finite values there
                    a = (t.valid & 1u) ? sc[px[i].s00] : 0.0f;
                    b = (t.valid & 2u) ? sc[px[i].s10] : 0.0f;
                    c = (t.valid & 4u) ? sc[px[i].s01] : 0.0f;
                    d = (t.valid & 8u) ? sc[px[i].s11] : 0.0f;
                }
                float v = t.w00 * a;
                v = __fmaf_rn(t.w10, b, v);
                v = __fmaf_rn(t.w01, c, v);
                v = __fmaf_rn(t.w11, d, v);
                if ((ALLV || live[i]) && (FULL || cg + j < c1))
                    __stcs(op, v);
                op += HW;
            }
            opx[i] = op;
        }
    };

    for (int g = 0; g < ngroups; ++g) {
        const unsigned bar = bar0 + 8u * (g % kStages);
        const unsigned parity = (g / kStages) & 1u;
        // bounded wait: a mis-programmed transfer must fail the launch, not hang the device
        unsigned done = 0;
        for (unsigned tries = 0; !done; ++tries) {
            asm volatile(
                "{\n"
                ".reg .pred p;\n"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
                "selp.u32 %0, 1, 0, p;\n"
                "}\n"
                : "=r"(done)
                : "r"(bar), "r"(parity)
                : "memory");
            if (!done && tries > (1u << 22))
                __trap();
        }
        const float* sb = stage + (g % kStages) * kStageFloats;
        if (c0 + g * kG + kG > c1)
            consume(std::false_type{}, std::false_type{}, sb, g);
        else if (tile_all_valid)
            consume(std::true_type{}, std::true_type{}, sb, g);
        else
            consume(std::false_type{}, std::true_type{}, sb, g);
        __syncthreads();   // every thread has read stage g % kStages: it may be refilled
        if (tid == 0 && g + kStages < ngroups)
            request(g + kStages);
    }
}

// ---- host side: tensor map over [N][C][H][W], box [1][kG][kBY][kBX] -----------------------------------------
typedef CUresult (*WarpEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static bool warp_make_map(CUtensorMap* m, const float* base, int N, int C, int H, int W)
{
    static WarpEncodeTiledFn enc = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* fp = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q) == cudaSuccess
            && q == cudaDriverEntryPointSuccess)
            enc = reinterpret_cast<WarpEncodeTiledFn>(fp);
        else
            (void)cudaGetLastError();
    }
    if (!enc)
        return false;
    const cuuint64_t dims[4] = {static_cast<cuuint64_t>(W), static_cast<cuuint64_t>(H), static_cast<cuuint64_t>(C),
        static_cast<cuuint64_t>(N)};
    const cuuint64_t strides[3] = {static_cast<cuuint64_t>(W) * 4, static_cast<cuuint64_t>(W) * H * 4,
        static_cast<cuuint64_t>(W) * H * C * 4};
    const cuuint32_t box[4] = {kBX, kBY, kG, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(base), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)
        == CUDA_SUCCESS;
}

// can the staged kernel take this tensor at all?
bool warp_staged_applicable(const float* in, int N, int C, int H, int W)
{
    return (W % 4) == 0 && W >= kBX && H >= kBY && C >= kG && aligned16(in) && N <= 65535;
}

// returns VSC_OK / an error; *launched = false when the tensor map could not be built (the caller falls back)
int launch_warp_staged(const float* in, const float* flow, float* out, int N, int C, int H, int W, int nchunk_req,
    cudaStream_t st, bool* launched)
{
    *launched = false;
    CUtensorMap map;
    if (!warp_make_map(&map, in, N, C, H, W))
        return VSC_OK;
    const unsigned tx = cdiv(W, kTileW), ty = cdiv(H, kTileH);
    const long long tiles = static_cast<long long>(tx) * ty * N;
    // the per-tile set-up (flow, corners, bounding box: a fifth of all instructions with 32 channels per tile) is paid
    // once per channel chunk, so chunks are only used to get to one wave of 148 SMs x 4 resident CTAs
    // (profiles/r2_warp_staged_events.txt: 32x544x960 34.8 / 38.9 / 47.1 us with 1 / 2 / 4 chunks)
    const long long want = 1LL * sm_count() * 4;
    int nchunk = nchunk_req > 0 ? nchunk_req : static_cast<int>((want + tiles - 1) / tiles);
    if (nchunk < 1) nchunk = 1;
    if (nchunk > (C + 2 * kG - 1) / (2 * kG)) nchunk = (C + 2 * kG - 1) / (2 * kG);   // >= 2 groups per CTA
    if (nchunk < 1) nchunk = 1;
    int chunk = (C + nchunk - 1) / nchunk;
    chunk = (chunk + kG - 1) / kG * kG;
    nchunk = (C + chunk - 1) / chunk;
    if (ty > 65535 || static_cast<long long>(N) * nchunk > 65535)
        return VSC_OK;
    static unsigned long long configured = 0;
    if (const int e = ensure_dynamic_smem(warp_nchw_staged_kernel, kStagedSmem, false, configured))
        return e;
    const dim3 grid(tx, ty, static_cast<unsigned>(N * nchunk));
    const int rc = launch_pdl(warp_nchw_staged_kernel, grid, dim3(kThreads), kStagedSmem, st, map, in, flow, out, C, H,
        W, chunk, nchunk);
    count_launch();
    *launched = true;
    return rc ? rc : launch_status();
}

}  // namespace vsc
