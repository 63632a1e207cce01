// EXPERIMENT, NOT BUILT (build.py compiles csrc/*.cu only).  Persistent form of the TMA-staged custom::Warp kernel
// (csrc/ops_warp_staged.cu), kept as the record of a measured negative result of round 2
// (profiles/r2_warp_persistent_events.txt): bit-identical on every Warp test (159 GPU tests with it as mode 5), and
// slower than the per-tile kernel -- 32x544x960: 41.0 us isolated / 36.0 us back to back with 2 CTAs per SM (128
// registers) against 34.9 / 38.2 us; 49 us with 3 or 4 CTAs per SM (80 / 64 registers, spills).  Holding a second
// tile's geometry costs the registers that the per-tile kernel spends on two more resident CTAs, and four
// independent CTAs hide the per-tile prologue better than one CTA's software pipeline does.
// To try it again: paste the block below before the host side of ops_warp_staged.cu, give launch_warp_staged a
// `persistent` flag (grid = min(tiles, CTAs per SM x SMs), dynamic smem kPersistSmem) and dispatch it for a warp mode.
#if 0
// ---- persistent variant -------------------------------------------------------------------------------------
// One CTA per SM slot walks the tiles (tile = blockIdx.x + i * gridDim.x) with ONE continuous ring of channel groups.
// What the per-tile kernel above exposes in every CTA -- flow load -> corners -> bounding box -> first TMA round trip,
// about 2 us that only the other resident CTAs overlap -- moves under the previous tile's channel loop: the next
// tile's flow is requested at the top of a tile, its corners and bounding box are computed after the current tile's
// third group, and its first boxes are requested into the ring slots the current tile's last groups free.
namespace {
struct PPix {
    float w00, w10, w01, w11;
    int s00;          // offset of corner 00 inside one channel of the staged box (the others: +1, +kBX, +kBX+1)
    unsigned valid;   // bit k set: corner k is read
    int pofs;         // y * W + x, or -1 when the pixel is outside the image
};
struct PTile {        // CTA-uniform description of a tile
    int n, c0, c1, x0, y0;
    int minc, minr;
    bool staged, allv;
};
constexpr int kPThreads = 256;
constexpr size_t kPersistSmem = kStages * kStageBytes + 64 + 2 * 8 * 4 * sizeof(int);
}  // namespace

#ifndef VSC_WARP_PCTAS
#define VSC_WARP_PCTAS 3
#endif
__global__ void __launch_bounds__(kPThreads, VSC_WARP_PCTAS) warp_nchw_persistent_kernel(const __grid_constant__ CUtensorMap map,
    const float* __restrict__ in, const float* __restrict__ flow, float* __restrict__ out, int C, int H, int W,
    int chunk, int nchunk, int tiles_x, int tiles_y, int ntiles)
{
    pdl_enter();
    extern __shared__ __align__(128) unsigned char staged_smem[];
    float* stage = reinterpret_cast<float*>(staged_smem);
    const unsigned bar0 = w_smem_u32(staged_smem + kStages * kStageBytes);
    int* red_all = reinterpret_cast<int*>(staged_smem + kStages * kStageBytes + 64);   // [2][8 warps][4]

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int tx = tid & (kTileW - 1), ty = tid / kTileW;   // ty 0..3: rows ty and ty + 4 of the tile
    const int HW = H * W;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < kStages; ++s)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar0 + 8u * s), "r"(1u) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    auto tile_origin = [&](int t, PTile& T) {
        const int bx = t % tiles_x;
        const int rest = t / tiles_x;
        const int by = rest % tiles_y;
        const int z = rest / tiles_y;
        T.n = z / nchunk;
        T.c0 = (z - T.n * nchunk) * chunk;
        T.c1 = min(C, T.c0 + chunk);
        T.x0 = bx * kTileW;
        T.y0 = by * kTileH;
    };
    // the four flow values of my two pixels (u0, v0, u1, v1); pixels outside the image read the tile's pixel 0
    auto load_flows = [&](const PTile& T, float (&f)[4]) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int x = T.x0 + tx, y = T.y0 + ty + 4 * i;
            const bool live = x < W && y < H;
            const float* fl = flow + static_cast<size_t>(T.n) * 2 * HW + (live ? y * W + x : 0);
            f[2 * i] = ldg_stream(fl);
            f[2 * i + 1] = ldg_stream(fl + HW);
        }
    };
    // corners, weights, bounding box of a tile; `which` selects the reduction scratch (0 / 1 alternate per tile)
    auto geometry = [&](PTile& T, const float (&f)[4], PPix (&px)[2], int which) {
        int* red = red_all + which * 32;
        int xl[2], yt[2];
        int minc = INT_MAX, maxc = INT_MIN, minr = INT_MAX, maxr = INT_MIN;
        bool all_valid = true;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int x = T.x0 + tx, y = T.y0 + ty + 4 * i;
            const bool live = x < W && y < H;
            const WarpTap t = warp_setup(live ? x : 0, live ? y : 0, f[2 * i], f[2 * i + 1], W, H);
            px[i].w00 = t.w00;
            px[i].w10 = t.w10;
            px[i].w01 = t.w01;
            px[i].w11 = t.w11;
            px[i].valid = live ? t.valid : 0u;
            px[i].pofs = live ? y * W + x : -1;
            xl[i] = 0;
            yt[i] = 0;
            if (px[i].valid) {
                xl[i] = static_cast<int>(floorf(static_cast<float>(x) + f[2 * i]));
                yt[i] = static_cast<int>(floorf(static_cast<float>(y) + f[2 * i + 1]));
            }
            const unsigned v = px[i].valid;
            if (v & 5u) { minc = min(minc, xl[i]); maxc = max(maxc, xl[i]); }
            if (v & 10u) { minc = min(minc, xl[i] + 1); maxc = max(maxc, xl[i] + 1); }
            if (v & 3u) { minr = min(minr, yt[i]); maxr = max(maxr, yt[i]); }
            if (v & 12u) { minr = min(minr, yt[i] + 1); maxr = max(maxr, yt[i] + 1); }
            all_valid = all_valid && v == 15u;
        }
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
        T.allv = __syncthreads_and(all_valid ? 1 : 0) != 0;
        const int w8 = lane & 7;
        minc = __reduce_min_sync(0xffffffffu, red[w8 * 4 + 0]);
        maxc = __reduce_max_sync(0xffffffffu, red[w8 * 4 + 1]);
        minr = __reduce_min_sync(0xffffffffu, red[w8 * 4 + 2]);
        maxr = __reduce_max_sync(0xffffffffu, red[w8 * 4 + 3]);
        const bool any = maxc >= minc && maxr >= minr;
        if (any)
            minc &= ~3;   // box rows start on 16-byte boundaries
        T.staged = any && (maxc - minc) < kBX && (maxr - minr) < kBY;
        T.minc = minc;
        T.minr = minr;
#pragma unroll
        for (int i = 0; i < 2; ++i)
            px[i].s00 = (yt[i] - minr) * kBX + (xl[i] - minc);
    };
    auto request = [&](const PTile& T, int g, int G) {   // thread 0: channel group g of tile T into ring slot G % kStages
        const unsigned bar = bar0 + 8u * (G % kStages);
        const unsigned dst = w_smem_u32(stage + (G % kStages) * kStageFloats);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(kStageBytes) : "memory");
        asm volatile(
            "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
            ::"r"(dst), "l"(reinterpret_cast<unsigned long long>(&map)), "r"(bar), "r"(T.minc), "r"(T.minr),
            "r"(T.c0 + g * kG), "r"(T.n)
            : "memory");
    };
    auto ngroups_of = [&](const PTile& T) { return (T.c1 - T.c0 + kG - 1) / kG; };

    PTile cur, nxt;
    PPix px[2], pn[2];
    float fc[4], fn[4];
    int G0 = 0;    // ring index of the current tile's group 0 (counts staged groups only)
    int req = 0;   // ring index of the next request (thread 0 issues them in order)
    int t = blockIdx.x;
    if (t >= ntiles)
        return;
    tile_origin(t, cur);
    load_flows(cur, fc);
    geometry(cur, fc, px, 0);
    int which = 1;
    for (; t < ntiles; t += gridDim.x) {
        const int tn = t + gridDim.x;
        const bool have_next = tn < ntiles;
        bool next_ready = false;
        if (have_next) {
            tile_origin(tn, nxt);
            load_flows(nxt, fn);   // in flight during this tile's channel loop
        }
        const int ngroups = ngroups_of(cur);
        float* op0 = out + (static_cast<size_t>(cur.n) * C + cur.c0) * HW;
        if (!cur.staged) {
            // scattered flow (or no readable corner at all): direct gathers, as warp_nchw_kernel
            WarpTap ta, tb;
            {
                const int xa = cur.x0 + tx, ya = cur.y0 + ty, yb = ya + 4;
                const bool la = px[0].pofs >= 0, lb = px[1].pofs >= 0;
                ta = warp_setup(la ? xa : 0, la ? ya : 0, fc[0], fc[1], W, H);
                tb = warp_setup(lb ? xa : 0, lb ? yb : 0, fc[2], fc[3], W, H);
                if (!la) ta.valid = 0u;
                if (!lb) tb.valid = 0u;
            }
            const float* ip = in + (static_cast<size_t>(cur.n) * C + cur.c0) * HW;
            float* opa = op0 + max(px[0].pofs, 0);
            float* opb = op0 + max(px[1].pofs, 0);
            for (int c = cur.c0; c < cur.c1; ++c, ip += HW, opa += HW, opb += HW) {
                const float a0 = warp_sample(ip, ta);
                const float b0 = warp_sample(ip, tb);
                if (px[0].pofs >= 0)
                    __stcs(opa, a0);
                if (px[1].pofs >= 0)
                    __stcs(opb, b0);
            }
        } else {
            if (tid == 0)
                while (req < G0 + min(kStages, ngroups)) {   // whatever the previous tile could not request for us
                    request(cur, req - G0, req);
                    ++req;
                }
            float* opx[2] = {op0 + max(px[0].pofs, 0), op0 + max(px[1].pofs, 0)};
            for (int g = 0; g < ngroups; ++g) {
                const int G = G0 + g;
                const unsigned bar = bar0 + 8u * (G % kStages);
                const unsigned parity = (G / kStages) & 1u;
                unsigned done = 0;
                for (unsigned tries = 0; !done; ++tries) {   // bounded: a mis-programmed transfer traps
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
                const float* sb = stage + (G % kStages) * kStageFloats;
                const bool full = cur.c0 + g * kG + kG <= cur.c1;
                if (cur.allv && full) {
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        const float* sc = sb + px[i].s00;
                        float* op = opx[i];
#pragma unroll
                        for (int j = 0; j < kG; ++j) {
                            const float a = sc[j * kPlane], b = sc[j * kPlane + 1];
                            const float c = sc[j * kPlane + kBX], d = sc[j * kPlane + kBX + 1];
                            float v = px[i].w00 * a;
                            v = __fmaf_rn(px[i].w10, b, v);
                            v = __fmaf_rn(px[i].w01, c, v);
                            v = __fmaf_rn(px[i].w11, d, v);
                            __stcs(op, v);
                            op += HW;
                        }
                        opx[i] = op;
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        const float* sc = sb + px[i].s00;
                        float* op = opx[i];
                        const unsigned vm = px[i].valid;
#pragma unroll
                        for (int j = 0; j < kG; ++j) {
                            // unused corners are not read: the box may hold non-finite values there
                            const float a = (vm & 1u) ? sc[j * kPlane] : 0.0f;
                            const float b = (vm & 2u) ? sc[j * kPlane + 1] : 0.0f;
                            const float c = (vm & 4u) ? sc[j * kPlane + kBX] : 0.0f;
                            const float d = (vm & 8u) ? sc[j * kPlane + kBX + 1] : 0.0f;
                            float v = px[i].w00 * a;
                            v = __fmaf_rn(px[i].w10, b, v);
                            v = __fmaf_rn(px[i].w01, c, v);
                            v = __fmaf_rn(px[i].w11, d, v);
                            if (px[i].pofs >= 0 && cur.c0 + g * kG + j < cur.c1)
                                __stcs(op, v);
                            op += HW;
                        }
                        opx[i] = op;
                    }
                }
                // the next tile's corners and bounding box, once its flow has had three groups of time to arrive
                if (have_next && !next_ready && (g == 2 || g == ngroups - 1)) {
                    geometry(nxt, fn, pn, which);   // (contains the block-wide synchronisation of this group)
                    next_ready = true;
                } else {
                    __syncthreads();   // every thread has read ring slot G % kStages: it may be refilled
                }
                if (tid == 0 && req == G + kStages) {
                    const int gi = req - G0;   // group index counted from this tile's first group
                    if (gi < ngroups) {
                        request(cur, gi, req);
                        ++req;
                    } else if (next_ready && nxt.staged && gi - ngroups < ngroups_of(nxt)) {
                        request(nxt, gi - ngroups, req);
                        ++req;
                    }
                }
            }
            G0 += ngroups;
        }
        if (have_next) {
            if (!next_ready)
                geometry(nxt, fn, pn, which);
            which ^= 1;
            cur = nxt;
#pragma unroll
            for (int i = 0; i < 2; ++i)
                px[i] = pn[i];
#pragma unroll
            for (int i = 0; i < 4; ++i)
                fc[i] = fn[i];
        }
    }
}

#endif
