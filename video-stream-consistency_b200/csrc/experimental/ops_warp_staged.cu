// EXPERIMENTAL -- not compiled into libvsc_b200.so (build.py only takes csrc/*.cu).  Round-2 candidate for
// custom::Warp: 64x4 pixel tiles whose source bounding box is staged in shared memory by TMA.
//
// Why: the gather kernels of ops_warp.cu serialise memory round trips inside every thread (flow -> taps of a channel
// batch -> next batch) and reach 0.57 of the HBM peak; 40 % of their instructions are per-pixel set-up and 8 of the
// 27.75 instructions per value in the channel loop form addresses.  Here the loads no longer depend on per-thread
// state: the tile's threads reduce the bounding box of all corners they will read; if it fits a kBX x kBY box (flow
// varying by up to ~12 px in x and 7 px in y across the tile), ONE thread streams that box, kG channels at a time,
// into a kStages-deep shared-memory ring (cp.async.bulk.tensor.4d over [N][C][H][W], out-of-bounds zero fill) while
// all threads take their four corners of the previous group from shared memory at constant offsets: per value 4 LDS
// + the same 7 FP operations + 1 store, no address arithmetic, no exposed latency after the first group.  Tiles
// whose corners are spread wider (scattered flow) gather directly, exactly like warp_nchw_kernel.  Same corners, same
// weights, same accumulation order: results are meant to be bit-identical.
//
// Status (end of round 1, GPU budget exhausted): the first version ran test_warp_identity_and_shift at 32x544x960
// correctly for ZERO flow (every box origin a multiple of 4 floats) and died in the bounded mbarrier wait below
// (__trap -> cudaErrorIllegalInstruction, by design instead of a hang) for the integer shift (+3, -2), i.e. as soon as
// a box started at a column that is not a multiple of 4: the bulk-tensor copy never completed.  The box origin is now
// rounded down to 16 bytes (kBX = 80 leaves 12 columns of slack for flow variation); that version compiles (40
// registers, UTMALDG in the SASS) but has not been run.  To try it: move this file's kernel and warp_make_map into
// ops_warp.cu, dispatch it for W % 4 == 0, W >= kBX, H >= kBY, 16-byte aligned input, and add the mode to the
// warp_mode fixture of tests/test_ops_gpu.py and to profiles/time_stage_a.py --warp.
//
//     nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -fmad=false -I include \
//          -I video-stream-consistency_b200/csrc -c video-stream-consistency_b200/csrc/experimental/ops_warp_staged.cu
#include <cuda.h>

#include <climits>

#include "../ops_warp.cu"   // WarpTap, warp_setup, warp_sample (this file is its own translation unit, never linked)

namespace vsc {

constexpr int kTileW = 64, kTileH = 4;           // output pixels per CTA (256 threads)
constexpr int kBX = 80, kBY = 12;                // staged box (floats x rows)
constexpr int kG = 4;                            // channels per stage
constexpr int kStages = 3;
constexpr int kStageFloats = kG * kBY * kBX;     // 3840 floats = 15 KB
constexpr unsigned kStageBytes = kStageFloats * sizeof(float);
constexpr size_t kStagedSmem = kStages * kStageBytes + 64 + 8 * 4 * sizeof(int);

__device__ __forceinline__ unsigned w_smem_u32(const void* p) { return static_cast<unsigned>(__cvta_generic_to_shared(p)); }

__global__ void __launch_bounds__(kTileW * kTileH) warp_nchw_staged_kernel(const __grid_constant__ CUtensorMap map,
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
    const int y = blockIdx.y * kTileH + tid / kTileW;
    const bool live = x < W && y < H;
    const int HW = H * W;
    const int n = blockIdx.z / nchunk;
    const int c0 = (blockIdx.z - n * nchunk) * chunk;
    const int c1 = min(C, c0 + chunk);
    const int p = live ? y * W + x : 0;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < kStages; ++s)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar0 + 8u * s), "r"(1u) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }

    const float* fl = flow + static_cast<size_t>(n) * 2 * HW + p;
    const float fu = ldg_stream(fl), fv = ldg_stream(fl + HW);
    WarpTap t = warp_setup(live ? x : 0, live ? y : 0, fu, fv, W, H);
    if (!live)
        t.valid = 0u;
    // integer corner coordinates: a readable corner implies finite xL / yT in [-1, W-1] / [-1, H-1]
    int xl = 0, yt = 0;
    if (t.valid) {
        xl = static_cast<int>(floorf(static_cast<float>(x) + fu));
        yt = static_cast<int>(floorf(static_cast<float>(y) + fv));
    }
    const int xr = xl + 1, yb = yt + 1;

    // bounding box of every corner this tile reads
    int minc = INT_MAX, maxc = INT_MIN, minr = INT_MAX, maxr = INT_MIN;
    if (t.valid & 5u) { minc = min(minc, xl); maxc = max(maxc, xl); }   // corners 00 / 01: column xL
    if (t.valid & 10u) { minc = min(minc, xr); maxc = max(maxc, xr); }  // corners 10 / 11: column xR
    if (t.valid & 3u) { minr = min(minr, yt); maxr = max(maxr, yt); }   // corners 00 / 10: row yT
    if (t.valid & 12u) { minr = min(minr, yb); maxr = max(maxr, yb); }  // corners 01 / 11: row yB
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
    __syncthreads();   // also publishes the mbarrier initialisation
    {
        const int w8 = lane & 7;
        minc = __reduce_min_sync(0xffffffffu, red[w8 * 4 + 0]);
        maxc = __reduce_max_sync(0xffffffffu, red[w8 * 4 + 1]);
        minr = __reduce_min_sync(0xffffffffu, red[w8 * 4 + 2]);
        maxr = __reduce_max_sync(0xffffffffu, red[w8 * 4 + 3]);
    }
    const bool any = maxc >= minc && maxr >= minr;
    // FIX NOT YET RUN ON A GPU (see the header): the box origin is rounded down to a multiple of 4 floats, so that
    // every row the TMA engine fetches starts on a 16-byte boundary
    if (any)
        minc &= ~3;
    const bool staged = any && (maxc - minc) < kBX && (maxr - minr) < kBY;   // CTA-uniform

    float* op = out + (static_cast<size_t>(n) * C + c0) * HW + p;
    if (!staged) {
        // scattered flow (or a tile without a single readable corner): direct gathers, as warp_nchw_kernel
        const float* ip = in + (static_cast<size_t>(n) * C + c0) * HW;
        int c = c0;
        for (; c + 4 <= c1; c += 4, ip += 4 * static_cast<size_t>(HW), op += 4 * static_cast<size_t>(HW)) {
            const float v0 = warp_sample(ip, t);
            const float v1 = warp_sample(ip + HW, t);
            const float v2 = warp_sample(ip + 2 * static_cast<size_t>(HW), t);
            const float v3 = warp_sample(ip + 3 * static_cast<size_t>(HW), t);
            if (live) {
                __stcs(op, v0);
                __stcs(op + HW, v1);
                __stcs(op + 2 * static_cast<size_t>(HW), v2);
                __stcs(op + 3 * static_cast<size_t>(HW), v3);
            }
        }
        for (; c < c1; ++c, ip += HW, op += HW) {
            const float v = warp_sample(ip, t);
            if (live)
                __stcs(op, v);
        }
        return;
    }

    // offsets of my corners inside one channel of a staged box (0 for corners that are not read)
    const int s00 = (t.valid & 1u) ? (yt - minr) * kBX + (xl - minc) : 0;
    const int s10 = (t.valid & 2u) ? (yt - minr) * kBX + (xr - minc) : 0;
    const int s01 = (t.valid & 4u) ? (yb - minr) * kBX + (xl - minc) : 0;
    const int s11 = (t.valid & 8u) ? (yb - minr) * kBX + (xr - minc) : 0;
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
        const int cg = c0 + g * kG;
#pragma unroll
        for (int j = 0; j < kG; ++j) {
            const float* sc = sb + j * (kBY * kBX);
            const float a = (t.valid & 1u) ? sc[s00] : 0.0f;
            const float b = (t.valid & 2u) ? sc[s10] : 0.0f;
            const float c = (t.valid & 4u) ? sc[s01] : 0.0f;
            const float d = (t.valid & 8u) ? sc[s11] : 0.0f;
            float v = t.w00 * a;
            v = v + t.w10 * b;
            v = v + t.w01 * c;
            v = v + t.w11 * d;
            if (live && cg + j < c1)
                __stcs(op + static_cast<size_t>(g * kG + j) * HW, v);
        }
        __syncthreads();   // every thread has read stage g % kStages: it may be refilled
        if (tid == 0 && g + kStages < ngroups)
            request(g + kStages);
    }
}

}  // namespace vsc

// ---- host side of the staged variant: tensor map over [N][C][H][W], box [1][kG][kBY][kBX] ----------------------
namespace vsc {
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
}  // namespace vsc

