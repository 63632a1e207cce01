// EXPERIMENT, NOT BUILT (build.py compiles csrc/*.cu only).  Kept as the record of a measured negative result of
// round 2 (profiles/r2_solver_notes.md, r2_solver_sweep_pixelpair.txt, r2_solver_pixelpair_ncu.txt): bit-identical to
// the shipped kernels on every test, 1.4x fewer instructions per sweep, and no faster -- 54.8 us vs 53.7 us per
// 8-sweep pass at 1080p, 182.8 vs 181 us at 4K.  It needs 246-248 registers, so 8 warps per SM (9 or 10 warps are
// capped at 168 registers: the register file is split over the four sub-partitions); ncu: issue 33 %, stalls wait
// 23 %, no_instruction 16 %, mio 10 %.  To resurrect it: move it back to csrc/, declare solver_stream2_pass in
// stab_solver.cu and try it before the other passes in run_sweeps (T = 8 only).
//
// Temporally blocked solver sweep, pixel-pair variant: the scheme of stab_solver_stream.cu (T Jacobi sweeps per
// launch, rows streamed through registers, skew of two rows per time level, neighbour-exchange ring in shared
// memory, warp-cooperative 16-byte cp.async staging, neighbour-pair named barriers; reference loop
// flowconsistency.cu:367-372) with a different ownership of the band's columns:
//
//   a thread owns TWO HORIZONTALLY ADJACENT PIXELS OF ONE CHANNEL -- band columns cA and cB = cA + 3 of the
//   interleaved row [H][3W] -- and keeps both in 64-bit register pairs.
//
// What that buys, per pair of value updates (the one-column kernel: 4 LDS + 2 STS + 14 FP = 20+ instructions):
//   * the right neighbour of cA is cB and the left neighbour of cB is cA: own registers.  Only the left neighbour
//     of cA and the right neighbour of cB come from the exchange ring, and each column is published once:
//     2 LDS + 2 STS instead of 4 LDS + 2 STS -- the ring's shared-memory traffic per update drops by a third;
//   * the arithmetic is issued as packed pairs (sm_100a add.rn.f32x2 / fma.rn.f32x2 -> FADD2 / FFMA2): 2 scalar
//     + 6 packed instructions instead of 14 scalar ones.  Each lane of a packed operation is the same IEEE
//     operation as the scalar instruction, so results stay bit-identical to the one-column kernel and to the
//     unblocked sweeps (tests/test_stab_gpu.py::test_blocked_solver_is_bit_identical_to_unblocked);
//   * the reference's left/right inclusion tests (flowconsistency.cu:215,221) that now fall INSIDE a thread (cB
//     outside the image or in the last pixel column, cA outside the image) are folded into the first addition as
//     a multiply by a per-thread constant 1.0f / 0.0f: fma(own, 1, ring) == own + ring, fma(own, 0, ring) ==
//     ring, exactly -- no extra instruction.
// Measured sensitivity of the one-column kernel (profiles/r2_solver_sensitivity.txt): removing 8 % of its
// instructions (one exchange load per update) saves 9.6 % of its time, removing 17 % (two additions) saves 11 %:
// the pass is bound by the number of instructions issued, not by one pipe.  (Round 1's two-column experiment paired
// columns 32 floats apart: FP halved but 6 ring accesses per pair remained, 228 registers, 8 warps -- no gain.)
//
// Geometry: a warp owns 60 consecutive band columns = 10 pixel pairs x 3 channels on lanes 0..29 (lanes 30, 31 idle
// along: 60 is the largest multiple of 6 columns a warp of two-column threads can own, and it keeps every warp's
// columns contiguous, so the staging stays warp-local).  NW warps -> band of 60 NW columns, 2 HALO of them halo.
// Requires 3W % 4 == 0 and 16-byte aligned images (16-byte staging chunks); the launcher falls back to the
// one-column kernel otherwise.
#include <type_traits>

#include "vsc_common.cuh"

#ifndef VSC_STREAM2_PF
#define VSC_STREAM2_PF 16
#endif

namespace vsc {

__host__ __device__ constexpr int stream2_halo(int T) { return (3 * T + 3) / 4 * 4; }

typedef unsigned long long f32x2;   // (A, B) packed pair of floats in a 64-bit register pair

namespace {
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
}  // namespace

int g_stream2 = 1;   // 0: never use this kernel (vsc_set_solver_mode | 0x8000); 1: auto
int g_stream2_nw = 0;   // 0: cost model; else force the number of warps (only 8 is built)
extern bool g_stream_pair;
extern int g_stream_edge_top, g_stream_edge_bot;   // stab_solver_rolled.cu

// NW warps per CTA (one CTA per SM); SYNC: 1 = neighbour-pair named barriers, 0 = CTA-wide barrier
template <int T, int NW, int SYNC>
__global__ void __launch_bounds__(NW * 32, 1) solver_stream2_kernel(const float* __restrict__ coefA,
    const float* __restrict__ coefB, const float* __restrict__ u_src, float* __restrict__ u_dst,
    const float* __restrict__ o_src, float* __restrict__ o_dst, int W, int H, int chunk_rows, int first_rows, float step,
    float mom)
{
    constexpr int NT = NW * 32;
    constexpr int BW = NW * 60;        // band columns
    constexpr int HALO = stream2_halo(T);
    constexpr int S = BW - 2 * HALO;   // columns stored per band
    constexpr int U = 2 * T;           // unroll = period of every ring index
    constexpr int PF = VSC_STREAM2_PF;   // staging ring depth (rows in flight from HBM); divides U
    static_assert(U % 4 == 0 && U % PF == 0 && S > 0 && S % 4 == 0, "ring periods / band geometry");
    // exchange ring: per (level, slot) one row of RW floats for the A columns and one for the B columns, indexed by
    // THREAD (conflict-free for the publishes and for both reads); 8 floats of never-written zero padding on each
    // side serve the first / last warp's out-of-band neighbours
    constexpr int RW = NT + 16;                         // thread `tid` publishes at [8 + tid] of a half row
    extern __shared__ float smem_raw[];
    float* ring = smem_raw;                             // [T*4][2 halves: A, B][RW]
    float* stage = smem_raw + T * 4 * 2 * RW;           // [PF][4 arrays][BW]

    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const bool live = lane < 30;                        // lanes 30, 31 own no columns
    const int q = (live ? lane : 29) / 3, ch = (live ? lane : 29) % 3;
    const int cA = warp * 60 + 6 * q + ch;              // band column of my left pixel; cB = cA + 3
    const int L = 3 * W;
    const int g0 = blockIdx.x * S - HALO;
    const int gA = g0 + cA, gB = gA + 3;
    const int r0 = blockIdx.y == 0 ? 0 : first_rows + (blockIdx.y - 1) * chunk_rows;
    const int r1 = min(H, blockIdx.y == 0 ? first_rows : r0 + chunk_rows);
    const int nsteps = (r1 - r0) + 3 * T;
    const bool okA = live && gA >= 0 && gA < L, okB = live && gB >= 0 && gB < L;
    const bool storeA = okA && cA >= HALO && cA < HALO + S;
    const bool storeB = okB && cA + 3 >= HALO && cA + 3 < HALO + S;
    // publishes to the exchange ring: image columns except the last pixel column (never a valid right neighbour:
    // x+1 < W-1, flowconsistency.cu:215); columns outside the image never publish, their slots stay zero
    const bool pubA = okA && gA < 3 * (W - 1), pubB = okB && gB < 3 * (W - 1);
    // the same tests for the neighbours that are my own registers: cB as right neighbour of cA, cA as left of cB
    const float mB = pubB ? 1.0f : 0.0f, mA = okA ? 1.0f : 0.0f;
    // ring reads: left neighbour of cA = column cA - 3 = the B column of thread lane-3 (or of lane 27..29 of the
    // previous warp), right neighbour of cB = column cA + 6 = the A column of lane+3 (or lane 0..2 of the next warp)
    const int idxL = RW + 8 + tid + (!live ? 0 : lane < 3 ? -5 : -3);    // into the B half of a ring row
    const int idxR = 8 + tid + (!live ? 0 : lane >= 27 ? 5 : 3);         // into the A half

    f32x2 win[T][4], uu[T][2], Ar[U], Br[U];
#pragma unroll
    for (int t = 0; t < T; ++t) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
            win[t][j] = 0ull;
        uu[t][0] = uu[t][1] = 0ull;
    }
#pragma unroll
    for (int j = 0; j < U; ++j) {
        Ar[j] = 0ull;
        Br[j] = 0ull;
    }
    for (int i = tid; i < T * 4 * 2 * RW; i += NT)
        smem_raw[i] = 0.0f;
    const f32x2 step2 = pk(step, step), mom2 = pk(mom, mom);

    // ---- staging: the 60 columns of a warp x 4 images are 60 chunks of 16 bytes: lane l copies chunk l and chunk
    // l + 32 (lanes 0..27); chunk i = image i / 15, columns [60 w + 4 (i % 15), +4).  Chunks lie entirely inside or
    // outside the image (3W % 4 == 0, g0 % 4 == 0).  Rows are requested two at a time right after the exchange
    // barrier, every lane waits for its own copies of the next interval's rows right before the barrier, and the
    // barrier (which always includes the whole warp) makes the other lanes' copies visible.
    const int i0 = lane, i1 = lane + 32;
    const bool has1 = i1 < 60;
    const int wc0 = warp * 60 + 4 * (i0 % 15), wc1 = warp * 60 + 4 * ((has1 ? i1 : 0) % 15);
    const int a0 = i0 / 15, a1 = (has1 ? i1 : 0) / 15;
    const bool cok0 = g0 + wc0 >= 0 && g0 + wc0 < L, cok1 = has1 && g0 + wc1 >= 0 && g0 + wc1 < L;
    const float* const src0 = a0 == 0 ? o_src : a0 == 1 ? u_src : a0 == 2 ? coefA : coefB;
    const float* const src1 = a1 == 0 ? o_src : a1 == 1 ? u_src : a1 == 2 ? coefA : coefB;
    int eo0 = (r0 - T) * L + (cok0 ? g0 + wc0 : 0);   // element offset of (next requested row, chunk)
    int eo1 = (r0 - T) * L + (cok1 ? g0 + wc1 : 0);
    const unsigned dst0 = static_cast<unsigned>(__cvta_generic_to_shared(stage + a0 * BW + wc0));
    const unsigned dst1 = static_cast<unsigned>(__cvta_generic_to_shared(stage + a1 * BW + wc1));
    auto request = [&](int y, int slot, bool commit) {   // row y of the four images -> staging slot
        const bool row_ok = y >= 0 && y < H;
        const unsigned n0 = (cok0 && row_ok) ? 16u : 0u, n1 = (cok1 && row_ok) ? 16u : 0u;
        const unsigned so = static_cast<unsigned>(slot * 4 * BW * sizeof(float));
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst0 + so), "l"(src0 + eo0), "r"(n0)
                     : "memory");
        if (has1)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst1 + so), "l"(src1 + eo1), "r"(n1)
                         : "memory");
        if (commit)
            asm volatile("cp.async.commit_group;" ::: "memory");
        eo0 += L;
        eo1 += L;
    };

    asm volatile("griddepcontrol.wait;" ::: "memory");   // the previous kernel of the stream wrote our inputs
#pragma unroll
    for (int j = 0; j < PF - 2; ++j)   // rows of steps 0 .. PF-3 as (PF-2)/2 groups of two
        request(r0 - T + j, j, (j & 1) == 1);
    asm volatile("cp.async.wait_group %0;" ::"n"((PF - 4) / 2) : "memory");   // the first group has landed
    __syncthreads();

    // element offset of (row y_in - 2T, column gA): where level T stores at this step; only dereferenced under
    // the store predicates
    int so = (r0 - 3 * T) * L + gA;

    auto ring_sync = [&]() {
        if constexpr (SYNC == 1 && NW <= 16) {
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

    // One step: levels T..1 (level T first: it reads the coefficient slot the arrival overwrites), then the arrival
    // of the level-0 row.  ROWMASK applies the top/bottom inclusion tests (flowconsistency.cu:227,232).
    auto step_body = [&](auto rowmask_tag, const int k, const int y_in) {
        constexpr bool ROWMASK = decltype(rowmask_tag)::value;
        // all exchange loads of the step first, in SOURCE order: their addresses (thread index plus a per-thread
        // neighbour offset) are opaque to the compiler's alias analysis, so it will not hoist a load above the
        // previous level's publish by itself -- and a level-by-level order serialises the T dependency chains
        // (LDS -> 7 dependent FP operations -> STS), which leaves two warps per scheduler nothing to overlap with
        float lf[T], rt[T];
#pragma unroll
        for (int t = T; t >= 1; --t) {
            const float* row = ring + ((t - 1) * 4 + ((k + 2) & 3)) * 2 * RW;
            lf[t - 1] = row[idxL];                     // left neighbour of cA
            rt[t - 1] = row[idxR];                     // right neighbour of cB
        }
        f32x2 on_pub[T];
#pragma unroll
        for (int t = T; t >= 1; --t) {
            const int rho = y_in - 2 * t;
            const f32x2 c = win[t - 1][(k + 2) & 3];   // produced at step s-2
            f32x2 up = win[t - 1][(k + 1) & 3];        // s-3
            f32x2 dn = win[t - 1][(k + 3) & 3];        // s-1
            if constexpr (ROWMASK) {
                dn = (rho + 1) < (H - 1) ? dn : 0ull;  // (:227)
                up = rho >= 1 ? up : 0ull;             // (:232)
            }
            // (rt + lf) per column: cA: own cB (masked) + ring; cB: ring + own cA (masked)
            const f32x2 s1 = pk(__fmaf_rn(hi_of(c), mB, lf[t - 1]), __fmaf_rn(lo_of(c), mA, rt[t - 1]));
            const f32x2 Ssum = add2(add2(s1, dn), up);
            const f32x2 a = Ar[(k + U - 2 * t) % U];
            const f32x2 b = Br[(k + U - 2 * t) % U];
            const f32x2 uo = uu[t - 1][k & 1];         // produced at step s-2
            const f32x2 un = fma2(step2, Ssum, fma2(a, c, b));
            const f32x2 on = fma2(mom2, uo, add2(c, un));
            on_pub[t - 1] = on;
            if (t < T) {
                win[t % T][k & 3] = on;   // (t % T only silences the bounds warning for t == T)
                uu[t % T][k & 1] = un;
            } else {
                // predicated, not branched around: a branch here would end the basic block
                const int in_rows = rho >= r0 && rho < r1;
                const int pa = storeA && in_rows, pb = storeB && in_rows;
                asm volatile("{\n .reg .pred p, q;\n setp.ne.s32 p, %8, 0;\n setp.ne.s32 q, %9, 0;\n"
                             " @p st.global.f32 [%0], %4;\n @p st.global.f32 [%1], %5;\n"
                             " @q st.global.f32 [%2], %6;\n @q st.global.f32 [%3], %7;\n}" ::"l"(o_dst + so),
                             "l"(u_dst + so), "l"(o_dst + so + 3), "l"(u_dst + so + 3), "f"(lo_of(on)), "f"(lo_of(un)),
                             "f"(hi_of(on)), "f"(hi_of(un)), "r"(pa), "r"(pb));
            }
        }
        // level 0: the row of this step landed, and became visible, before the barrier that opened the interval
        const float* st = stage + (k % PF) * 4 * BW + cA;
        const float oa = st[0], ob = st[3];
        win[0][k & 3] = pk(oa, ob);
        uu[0][k & 1] = pk(st[BW], st[BW + 3]);
        Ar[k % U] = pk(st[2 * BW], st[2 * BW + 3]);
        Br[k % U] = pk(st[3 * BW], st[3 * BW + 3]);
        // publishes last (see above)
#pragma unroll
        for (int t = 1; t < T; ++t) {
            float* pub = ring + (t * 4 + (k & 3)) * 2 * RW + 8 + tid;
            if (pubA)
                pub[0] = lo_of(on_pub[t - 1]);
            if (pubB)
                pub[RW] = hi_of(on_pub[t - 1]);
        }
        {
            float* pub = ring + (k & 3) * 2 * RW + 8 + tid;
            if (pubA)
                pub[0] = oa;
            if (pubB)
                pub[RW] = ob;
        }
        so += L;
    };
    // one barrier interval = request two rows, two steps, wait for the next interval's rows, barrier
    auto pair_open = [&](const int k, const int y_in) {
        request(y_in + PF - 2, (k + PF - 2) % PF, false);
        request(y_in + PF - 1, (k + PF - 1) % PF, true);
    };
    auto pair_close = [&]() {
        asm volatile("cp.async.wait_group %0;" ::"n"((PF - 4) / 2) : "memory");
        ring_sync();
    };

    for (int base = 0; base < nsteps; base += U) {
        const int y_first = r0 - T + base;               // y_in of the group's first step
        const bool edge = y_first <= 2 * T || y_first + U - 1 >= H || base + U > nsteps;
        if (edge) {
#pragma unroll
            for (int k = 0; k < U; k += 2)
                if (base + k < nsteps) {   // uniform across the CTA
                    pair_open(k, y_first + k);
                    step_body(std::true_type{}, k, y_first + k);
                    if (base + k + 1 < nsteps)
                        step_body(std::true_type{}, k + 1, y_first + k + 1);
                    pair_close();
                }
        } else {
#pragma unroll
            for (int k = 0; k < U; k += 2) {
                pair_open(k, y_first + k);
                step_body(std::false_type{}, k, y_first + k);
                step_body(std::false_type{}, k + 1, y_first + k + 1);
                pair_close();
            }
        }
    }
}

struct Stream2Geom {
    int nw, nb, nc, chunk_rows, first_rows;
    long long cost;
};

static Stream2Geom stream2_geom(int T, int NW, int L, int H, int sms)
{
    Stream2Geom g;
    g.nw = NW;
    const int S = NW * 60 - 2 * stream2_halo(T);
    g.nb = (L + S - 1) / S;
    int nc = sms / g.nb;
    if (nc < 1) nc = 1;
    const int min_rows = 4 * T;  // below this the 3T-step pipeline fill dominates
    if (nc > (H + min_rows - 1) / min_rows) nc = (H + min_rows - 1) / min_rows;
    if (nc < 1) nc = 1;
    int top = g_stream_edge_top >= 0 ? g_stream_edge_top : 0, bot = g_stream_edge_bot >= 0 ? g_stream_edge_bot : 0;
    if (nc < 3 || H < nc * (2 * T + top + bot)) top = bot = 0;
    g.chunk_rows = (H + top + bot + nc - 1) / nc;
    g.first_rows = g.chunk_rows - top;
    g.nc = H <= g.first_rows ? 1 : 1 + (H - g.first_rows + g.chunk_rows - 1) / g.chunk_rows;
    const long long waves = (static_cast<long long>(g.nb) * g.nc + sms - 1) / sms;
    // time of one step ~ a + b * warps (issue-bound); calibrated by profiles/sweep_solver_r2.py
    const int step_ns = 60 + 30 * NW;
    g.cost = waves * (g.chunk_rows + 3 * T) * step_ns;
    return g;
}

template <int T, int NW, int SYNC>
static int launch_stream2_impl(const Stream2Geom& g, const float* coefA, const float* coefB, const float* u_src,
    float* u_dst, const float* o_src, float* o_dst, int W, int H, float step, float mom, cudaStream_t st)
{
    constexpr int PF = VSC_STREAM2_PF;
    constexpr int RW = NW * 32 + 16;
    const size_t smem = (static_cast<size_t>(T) * 4 * 2 * RW + static_cast<size_t>(PF) * 4 * NW * 60) * sizeof(float);
    static unsigned long long configured = 0;
    if (const int e = ensure_dynamic_smem(solver_stream2_kernel<T, NW, SYNC>, smem, false, configured))
        return e;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(g.nb, g.nc);
    cfg.blockDim = dim3(NW * 32);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = g_pdl ? 1 : 0;
    const cudaError_t e = cudaLaunchKernelEx(&cfg, solver_stream2_kernel<T, NW, SYNC>, coefA, coefB, u_src, u_dst, o_src,
        o_dst, W, H, g.chunk_rows, g.first_rows, step, mom);
    count_launch();
    return e == cudaSuccess ? launch_status() : static_cast<int>(e);
}

template <int T, int NW>
static int launch_stream2(const Stream2Geom& g, const float* coefA, const float* coefB, const float* u_src,
    float* u_dst, const float* o_src, float* o_dst, int W, int H, float step, float mom, cudaStream_t st)
{
    if (g_stream_pair)
        return launch_stream2_impl<T, NW, 1>(g, coefA, coefB, u_src, u_dst, o_src, o_dst, W, H, step, mom, st);
    return launch_stream2_impl<T, NW, 0>(g, coefA, coefB, u_src, u_dst, o_src, o_dst, W, H, step, mom, st);
}

// true if this kernel can run the pass (then *rc is its status); false: the caller uses the one-column kernel
bool solver_stream2_pass(int T, const float* coefA, const float* coefB, const float* u_src, float* u_dst,
    const float* o_src, float* o_dst, int W, int H, float step, float mom, cudaStream_t st, int* rc)
{
    if (!g_stream2 || T != 8)
        return false;
    if ((3LL * W) % 4 != 0 || !aligned16(coefA) || !aligned16(coefB) || !aligned16(u_src) || !aligned16(o_src))
        return false;
    if (3LL * W * (static_cast<long long>(H) + 64) >= 0x7fffffffLL)
        return false;
    const int L = 3 * W, sms = sm_count();
    // Warps per CTA: the register file is split over the SM's four sub-partitions (16384 registers each), so a
    // CTA whose warp count is not a multiple of 4 is capped by its fullest sub-partition: 9 or 10 warps get 168
    // registers per thread, like 12, and this kernel needs 246 (T = 8: 160 of them are the pairs' state).  8 warps
    // = 2 per sub-partition = 255 registers.
    const Stream2Geom bg = stream2_geom(T, 8, L, H, sms);
    (void)g_stream2_nw;
    *rc = launch_stream2<8, 8>(bg, coefA, coefB, u_src, u_dst, o_src, o_dst, W, H, step, mom, st);
    return true;
}

}  // namespace vsc
