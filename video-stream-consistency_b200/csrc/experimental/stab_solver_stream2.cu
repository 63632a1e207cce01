// EXPERIMENT, NOT BUILT (build.py compiles csrc/*.cu only).  Kept as the record of a measured negative result:
// bit-identical to the shipped kernel on every test (188 GPU tests green with it as the default), 1.5x fewer
// instructions per update -- and no faster: 196.7 us vs 187.2 us per 8-sweep pass at 4K, 57.4 vs 57.2 us at
// 1080p (profiles/r1_sweep_bands_packed_vs_scalar.txt).  With 228 registers per thread only 6-8 warps fit on an
// SM; ncu shows issue 34 %, stall reason "wait" (fixed-latency dependencies) 1.7 warps per issue cycle
// (profiles/r1_stream2_packed_ncu.txt): the pass is bound by per-warp dependent-issue latency and the exchange
// ring's shared-memory pipe, not by instruction issue, so halving the FP instruction count buys nothing.
// To resurrect it: move it back to csrc/, declare solver_stream2_pass in stab_solver.cu and try it before
// solver_stream_pass in run_sweeps.
//
// Temporally blocked solver sweep, second generation: the scheme of stab_solver_stream.cu (T Jacobi sweeps per
// launch, rows streamed through registers, skew of two rows per time level, neighbour-exchange ring in shared
// memory, warp-local 16-byte cp.async staging, neighbour-pair named barriers) with TWO band columns per thread
// and every floating-point operation issued as a packed pair (FADD2 / FFMA2, sm_100a `add.rn.f32x2` /
// `fma.rn.f32x2`).
//
// Why: the one-column kernel is bound by instruction issue (about 132 instructions per thread and step for 8
// updates, 62 % issue utilisation, profiles/r1_notes.md).  A thread that owns two columns holds their state in
// 64-bit register pairs, so the 7 floating-point operations of an update are issued once for both columns, and
// the per-step overhead (staging, addressing, barriers, loop control) is paid once per two columns: the
// instruction count per update roughly halves and the exchange ring's shared-memory bandwidth becomes the
// bound.  Each lane of a packed operation is the same IEEE operation as the scalar instruction, so the results
// are bit-identical to the one-column kernel and to the unblocked sweeps.
//
// Column ownership: warp w covers the 64 band columns [64w, 64w+64); lane l owns columns 64w+l ("lo") and
// 64w+32+l ("hi").  Every shared-memory access is therefore 32 consecutive floats per warp (conflict-free), the
// x-neighbours (+-3 floats) of both columns live in the same or an adjacent warp exactly as in the one-column
// kernel, and every global store is a full 128-byte line.
//
// Requires 3W % 4 == 0 and 16-byte aligned images (16-byte staging); the launcher falls back to the one-column
// kernel otherwise.
#include <type_traits>

#include "vsc_common.cuh"

namespace vsc {

__host__ __device__ constexpr int stream2_halo(int T) { return (3 * T + 3) / 4 * 4; }

typedef unsigned long long f32x2;   // (lo, hi) packed pair of floats in a 64-bit register pair

__device__ __forceinline__ f32x2 pk(float lo, float hi)
{
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ float lo_of(f32x2 v) { return __uint_as_float(static_cast<unsigned>(v)); }
__device__ __forceinline__ float hi_of(f32x2 v) { return __uint_as_float(static_cast<unsigned>(v >> 32)); }
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b)
{
    f32x2 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c)
{
    f32x2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}

// BW = band width in floats; BW / 2 threads per CTA, one CTA per SM
template <int T, int BW, bool PAIR>
__global__ void __launch_bounds__(BW / 2, 1) solver_stream2_kernel(const float* __restrict__ coefA,
    const float* __restrict__ coefB, const float* __restrict__ u_src, float* __restrict__ u_dst,
    const float* __restrict__ o_src, float* __restrict__ o_dst, int W, int H, int chunk_rows, float step, float mom)
{
    constexpr int NT = BW / 2;
    constexpr int HALO = stream2_halo(T);
    constexpr int S = BW - 2 * HALO;   // columns stored per band
    constexpr int U = 2 * T;           // unroll = period of every ring index
    constexpr int PF = U;              // staging ring depth (rows in flight from HBM)
    static_assert(U % 4 == 0 && BW % 128 == 0, "ring period / warp geometry");
    extern __shared__ float smem_raw[];
    float* sm = smem_raw + 4;                       // exchange ring [T*4][BW], 4 floats of padding on each side
    float* stage = smem_raw + T * 4 * BW + 8;       // staging ring [PF][4 arrays][BW]

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int cl = warp * 64 + lane;                // my "lo" band column; "hi" = cl + 32
    const int L = 3 * W;
    const int g0 = blockIdx.x * S - HALO;
    const int r0 = blockIdx.y * chunk_rows;
    const int r1 = min(H, r0 + chunk_rows);
    const int nsteps = (r1 - r0) + 3 * T;
    const int gl = g0 + cl, gh = gl + 32;
    const bool ok_l = gl >= 0 && gl < L, ok_h = gh >= 0 && gh < L;
    const bool store_l = ok_l && cl >= HALO && cl < HALO + S;
    const bool store_h = ok_h && cl + 32 >= HALO && cl + 32 < HALO + S;
    // publishes to the exchange ring: image columns except the last pixel column (never a valid right neighbour:
    // x+1 < W-1, flowconsistency.cu:215); columns outside the image never publish, their slots stay zero
    const bool pub_l = ok_l && gl < 3 * (W - 1), pub_h = ok_h && gh < 3 * (W - 1);

    f32x2 win[T][4], uu[T][4], Ar[U], Br[U];
#pragma unroll
    for (int t = 0; t < T; ++t)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            win[t][j] = 0ull;
            uu[t][j] = 0ull;
        }
#pragma unroll
    for (int j = 0; j < U; ++j) {
        Ar[j] = 0ull;
        Br[j] = 0ull;
    }
    for (int i = tid; i < T * 4 * BW + 8; i += NT)
        smem_raw[i] = 0.0f;
    const f32x2 step2 = pk(step, step), mom2 = pk(mom, mom);

    // ---- staging: the 64 columns of a warp x 4 images are 64 chunks of 16 bytes, two per lane.  Lane l copies
    // chunk q = l and q = l + 32: image q / 16, columns [64w + 4 * (q % 16), +4).  Producer and consumers of a
    // chunk are lanes of the same warp: cp.async.wait_group + __syncwarp is all the ordering needed.
    // Addresses: one running 32-bit ELEMENT offset per chunk (arrays hold < 2^31 floats, checked by the
    // launcher), turned into a pointer with one IMAD.WIDE.
    const int q0 = lane, q1 = lane + 32;
    const int a0 = q0 >> 4, a1 = q1 >> 4;                      // 0/1 and 2/3
    const int wc0 = warp * 64 + 4 * (q0 & 15), wc1 = warp * 64 + 4 * (q1 & 15);
    const bool cok0 = g0 + wc0 >= 0 && g0 + wc0 < L, cok1 = g0 + wc1 >= 0 && g0 + wc1 < L;
    const float* const src0 = a0 == 0 ? o_src : u_src;
    const float* const src1 = a1 == 2 ? coefA : coefB;
    // element offset of (row of the NEXT request, first column of the chunk); rows above the image give negative
    // offsets that are never dereferenced (src-size 0)
    int eo0 = (r0 - T) * L + (cok0 ? g0 + wc0 : 0);
    int eo1 = (r0 - T) * L + (cok1 ? g0 + wc1 : 0);
    const unsigned dst0 = static_cast<unsigned>(__cvta_generic_to_shared(stage + a0 * BW + wc0));
    const unsigned dst1 = static_cast<unsigned>(__cvta_generic_to_shared(stage + a1 * BW + wc1));
    auto request = [&](int y, int slot) {   // row y of the four images -> staging slot
        const bool row_ok = y >= 0 && y < H;
        const unsigned n0 = (cok0 && row_ok) ? 16u : 0u, n1 = (cok1 && row_ok) ? 16u : 0u;
        const unsigned so = static_cast<unsigned>(slot * 4 * BW * sizeof(float));
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst0 + so), "l"(src0 + eo0), "r"(n0)
                     : "memory");
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst1 + so), "l"(src1 + eo1), "r"(n1)
                     : "memory");
        asm volatile("cp.async.commit_group;" ::: "memory");
        eo0 += L;
        eo1 += L;
    };
#pragma unroll
    for (int j = 0; j < PF - 1; ++j)   // rows of steps 0 .. PF-2
        request(r0 - T + j, j);
    __syncthreads();

    // element offset of (row y_in - 2T, column gl): where level T stores at this step.  gl may lie left of the
    // image while gl + 32 is inside it; the offset is only dereferenced under the store predicates
    int so = (r0 - 3 * T) * L + gl;

    // ---- exchange-ring synchronisation: neighbour-pair named barriers, as in the one-column kernel
    constexpr int NW = NT / 32;
    auto ring_sync = [&]() {
        if constexpr (PAIR && NW <= 16) {
            const int first = (warp & 1) ? warp + 1 : warp;   // boundary ids: left = warp, right = warp + 1
            const int second = (warp & 1) ? warp : warp + 1;
            if (first >= 1 && first <= NW - 1)
                asm volatile("bar.sync %0, 64;" ::"r"(first) : "memory");
            if (second >= 1 && second <= NW - 1)
                asm volatile("bar.sync %0, 64;" ::"r"(second) : "memory");
        } else {
            __syncthreads();
        }
    };

    auto step_body = [&](auto rowmask_tag, const int k, const int y_in) {
        constexpr bool ROWMASK = decltype(rowmask_tag)::value;
#pragma unroll
        for (int t = T; t >= 1; --t) {
            const int rho = y_in - 2 * t;
            const f32x2 c = win[t - 1][(k + 2) & 3];   // produced at step s-2
            f32x2 up = win[t - 1][(k + 1) & 3];        // s-3
            f32x2 dn = win[t - 1][(k + 3) & 3];        // s-1
            const float* row = sm + ((t - 1) * 4 + ((k + 2) & 3)) * BW + cl;
            const f32x2 lf = pk(row[-3], row[29]);
            const f32x2 rt = pk(row[3], row[35]);
            if constexpr (ROWMASK) {
                dn = (rho + 1) < (H - 1) ? dn : 0ull;  // (flowconsistency.cu:227)
                up = rho >= 1 ? up : 0ull;             // (:232)
            }
            const f32x2 Ssum = add2(add2(add2(rt, lf), dn), up);
            const f32x2 a = Ar[(k + U - 2 * t) % U];
            const f32x2 b = Br[(k + U - 2 * t) % U];
            const f32x2 uo = uu[t - 1][(k + 2) & 3];
            const f32x2 un = fma2(step2, Ssum, fma2(a, c, b));
            const f32x2 on = fma2(mom2, uo, add2(c, un));
            if (t < T) {
                win[t % T][k & 3] = on;
                uu[t % T][k & 3] = un;
                float* pub = sm + ((t % T) * 4 + (k & 3)) * BW + cl;
                if (pub_l)
                    pub[0] = lo_of(on);
                if (pub_h)
                    pub[32] = hi_of(on);
            } else if (rho >= r0 && rho < r1) {
                if (store_l) {
                    o_dst[so] = lo_of(on);
                    u_dst[so] = lo_of(un);
                }
                if (store_h) {
                    o_dst[so + 32] = hi_of(on);
                    u_dst[so + 32] = hi_of(un);
                }
            }
        }
        // level 0 arrives: the row of step s was requested at step s-PF+1
        asm volatile("cp.async.wait_group %0;" ::"n"(PF - 2) : "memory");
        __syncwarp();
        {
            const float* st = stage + (k % PF) * 4 * BW + cl;
            const float ol = st[0], oh = st[32];
            win[0][k & 3] = pk(ol, oh);
            uu[0][k & 3] = pk(st[BW], st[BW + 32]);
            float* pub = sm + (k & 3) * BW + cl;
            if (pub_l)
                pub[0] = ol;
            if (pub_h)
                pub[32] = oh;
            Ar[k % U] = pk(st[2 * BW], st[2 * BW + 32]);
            Br[k % U] = pk(st[3 * BW], st[3 * BW + 32]);
        }
        // request the row of step s+PF-1 into the slot this warp read at step s-1
        request(y_in + PF - 1, (k + PF - 1) % PF);
        so += L;
        if ((k & 1) == 1)
            ring_sync();
    };

    for (int base = 0; base < nsteps; base += U) {
#pragma unroll
        for (int k = 0; k < U; ++k) {
            const int s = base + k;
            if (s < nsteps) {  // uniform across the CTA
                const int y_in = r0 - T + s;
                if (y_in <= 2 * T || y_in >= H)
                    step_body(std::true_type{}, k, y_in);
                else
                    step_body(std::false_type{}, k, y_in);
            }
        }
    }
}

struct Stream2Geom {
    int nb, nc, chunk_rows;
    long long cost;
};

static Stream2Geom stream2_geom(int T, int BW, int L, int H, int sms)
{
    Stream2Geom g;
    const int S = BW - 2 * stream2_halo(T);
    g.nb = (L + S - 1) / S;
    int nc = sms / g.nb;
    if (nc < 1) nc = 1;
    const int min_rows = 4 * T;
    if (nc > (H + min_rows - 1) / min_rows) nc = (H + min_rows - 1) / min_rows;
    if (nc < 1) nc = 1;
    g.chunk_rows = (H + nc - 1) / nc;
    g.nc = (H + g.chunk_rows - 1) / g.chunk_rows;
    const long long waves = (static_cast<long long>(g.nb) * g.nc + sms - 1) / sms;
    g.cost = waves * (g.chunk_rows + 3 * T) * BW;
    return g;
}

template <int T, int BW, bool PAIR>
static int launch_stream2_impl(const Stream2Geom& g, const float* coefA, const float* coefB, const float* u_src,
    float* u_dst, const float* o_src, float* o_dst, int W, int H, float step, float mom, cudaStream_t st)
{
    const size_t smem = (static_cast<size_t>(T) * 4 * BW + 8 + static_cast<size_t>(2 * T) * 4 * BW) * sizeof(float);
    static unsigned long long configured = 0;
    if (const int e = ensure_dynamic_smem(solver_stream2_kernel<T, BW, PAIR>, smem, false, configured))
        return e;
    const dim3 grid(g.nb, g.nc);
    solver_stream2_kernel<T, BW, PAIR><<<grid, BW / 2, smem, st>>>(coefA, coefB, u_src, u_dst, o_src, o_dst, W, H,
        g.chunk_rows, step, mom);
    count_launch();
    return launch_status();
}

extern bool g_stream_pair;
extern int g_stream_band;

template <int T>
static int launch_stream2_best(const float* coefA, const float* coefB, const float* u_src, float* u_dst,
    const float* o_src, float* o_dst, int W, int H, float step, float mom, cudaStream_t st)
{
    const int L = 3 * W, sms = sm_count();
    constexpr int NCAND = 3;
    const int cands[NCAND] = {512, 384, 256};
    int best = 0;
    Stream2Geom bg = stream2_geom(T, cands[0], L, H, sms);
    if (g_stream_band >= 1 && g_stream_band <= 4) {
        best = g_stream_band == 1 ? 0 : (g_stream_band == 4 ? 2 : 1);   // 512 / 448,384 -> 384 / 256
        bg = stream2_geom(T, cands[best], L, H, sms);
    } else {
        for (int i = 1; i < NCAND; ++i) {
            const Stream2Geom g = stream2_geom(T, cands[i], L, H, sms);
            if (g.cost < bg.cost) {
                bg = g;
                best = i;
            }
        }
    }
#define VSC_S2(BWV)                                                                                               \
    (g_stream_pair ? launch_stream2_impl<T, BWV, true>(bg, coefA, coefB, u_src, u_dst, o_src, o_dst, W, H, step, \
                         mom, st)                                                                                 \
                   : launch_stream2_impl<T, BWV, false>(bg, coefA, coefB, u_src, u_dst, o_src, o_dst, W, H, step, \
                         mom, st))
    switch (best) {
        case 0: return VSC_S2(512);
        case 1: return VSC_S2(384);
        default: return VSC_S2(256);
    }
#undef VSC_S2
}

// two-columns-per-thread pass; VSC_E_INVALID when the images do not qualify (caller uses the one-column kernel)
int solver_stream2_pass(int T, const float* coefA, const float* coefB, const float* u_src, float* u_dst,
    const float* o_src, float* o_dst, int W, int H, float step, float mom, cudaStream_t st)
{
    const long long n = 3LL * W * H;
    if ((3LL * W) % 4 != 0 || !aligned16(coefA) || !aligned16(coefB) || !aligned16(u_src) || !aligned16(o_src)
        || n + 64LL * W >= 0x7fffffffLL)
        return VSC_E_INVALID;
    if (T == 8)
        return launch_stream2_best<8>(coefA, coefB, u_src, u_dst, o_src, o_dst, W, H, step, mom, st);
    if (T == 6)
        return launch_stream2_best<6>(coefA, coefB, u_src, u_dst, o_src, o_dst, W, H, step, mom, st);
    if (T == 4)
        return launch_stream2_best<4>(coefA, coefB, u_src, u_dst, o_src, o_dst, W, H, step, mom, st);
    if (T == 2)
        return launch_stream2_best<2>(coefA, coefB, u_src, u_dst, o_src, o_dst, W, H, step, mom, st);
    return VSC_E_INVALID;
}

}  // namespace vsc
