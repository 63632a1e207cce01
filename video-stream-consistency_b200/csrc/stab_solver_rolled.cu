// Temporally blocked solver sweep, rolled form: the scheme of stab_solver_stream.cu (T Jacobi sweeps per launch, rows
// streamed through registers, skew of two rows per time level, neighbour-exchange ring in shared memory,
// warp-cooperative 16-byte cp.async staging, neighbour-pair named barriers; reference loop
// flowconsistency.cu:367-372) with a step loop that is unrolled FOUR times instead of 2T times.
//
// Why.  In stab_solver_stream.cu every ring index is a compile-time constant because the step loop is unrolled over
// the longest ring period, 2T steps (the per-row coefficients A, B live 2T steps in registers).  At T = 10 that is
// 2470 instructions = 39.5 KB of hot loop, plus a second copy with the top/bottom row masks: more than the 32 KB
// instruction cache.  ncu (profiles/r2_solver_notes.md, r2_solver_fillspec_ncu.txt): a quarter of the warp samples INSIDE the hot loop are
// `no_instruction`, and code that a CTA executes only once (the masked copy in the image's first and last row chunk;
// an experiment that compiled the pipeline-fill steps without their idle levels: 12 % fewer instructions, 45 %
// SLOWER) costs twice as much per instruction as resident code.  Here only the period-4 rings (row windows, exchange
// ring slots) index by the unrolled step; the coefficient window is a register array of 2T+4 rows that is shifted
// down by four rows after every group of four steps (2T register moves per array and group, +8 % instructions), and
// the staging slot is a run-time base advanced once per group.  Hot loop: 4 steps = about 560 instructions = 9 KB,
// masked copy the same -- the whole kernel stays resident.
//
// Also here (all measured in profiles/r2_solver_notes.md):
//   * the momentum ring is two rows deep instead of four (20 registers at T = 10, which pay for the coefficient
//     window's four extra rows);
//   * staging hand-off by the exchange barrier instead of a per-step cp.async wait + __syncwarp: rows are requested
//     two at a time right after the barrier (into the slots read before it), every lane waits for its own copies of
//     the NEXT interval's rows right before the barrier, and the barrier (which always includes the whole warp)
//     makes the other lanes' copies visible; the level-T stores are predicated, not branched around, so an interval
//     of two steps is one basic block;
//   * the first and the last row chunk may be shorter than the others (the CTAs that run the masked copy are the
//     slowest of the one-wave grid).
//   * quad-gather exchange ring (QG, default): see the comment at solver_rolled_kernel.  (An mbarrier arrive / wait
//     hand-off between neighbouring warps with a step of slack was measured 10-50 % slower than the named barriers and
//     removed again: profiles/r2_solver_split_sweep.txt.)
// Results are bit-identical to stab_solver_stream.cu and to the unblocked sweeps (same arithmetic per value, only
// the schedule differs): tests/test_stab_gpu.py::test_blocked_solver_is_bit_identical_to_unblocked.
//
// Requires 3W % 4 == 0 and 16-byte aligned images (16-byte staging chunks); the launcher falls back to
// stab_solver_stream.cu otherwise.
#include <type_traits>

#include "vsc_common.cuh"

#ifndef VSC_SOLVER_PREFETCH_NB
#define VSC_SOLVER_PREFETCH_NB 0
#endif
#ifndef VSC_SOLVER_PERMUTE_DEFAULT
#define VSC_SOLVER_PERMUTE_DEFAULT 0
#endif
#ifndef VSC_SOLVER_QG_DEFAULT
#define VSC_SOLVER_QG_DEFAULT 1
#endif

namespace vsc {

__host__ __device__ constexpr int rolled_halo(int T) { return (3 * T + 3) / 4 * 4; }
// staging ring depth (rows in flight from HBM): about 2T, a multiple of 4 (the slot base advances by 4 per group)
__host__ __device__ constexpr int rolled_pf(int T) { return (2 * T + 3) / 4 * 4 < 8 ? 8 : (2 * T + 3) / 4 * 4; }
extern std::atomic<bool> g_stream_pair;          // stab_solver_stream.cu: neighbour-pair named barriers (default) or CTA barrier
extern std::atomic<bool> g_stream_coop;          // false: per-thread 4-byte staging requested -> stab_solver_stream.cu
extern std::atomic<int> g_stream_band;           // 0 = cost model; 1..4 force a band width (512, 448, 384, 256)
std::atomic<int> g_stream_rolled = 1;            // 0: never use this kernel (vsc_set_solver_mode | 0x8000), 1: auto, 2: always (| 0x4000)
std::atomic<int> g_stream_edge_top = -1;         // rows by which the first / last row chunk is shorter than the others (-1: default)
std::atomic<int> g_stream_edge_bot = -1;
std::atomic<int> g_stream_permute = VSC_SOLVER_PERMUTE_DEFAULT;   // 1: warps of one scheduler own adjacent column blocks (| 1 << 30 flips it)
std::atomic<int> g_stream_qg = VSC_SOLVER_QG_DEFAULT;   // 1: exchange ring in the quad-gather layout (vsc_set_solver_mode | 1 << 28 flips it)

// QG ("quad gather"): the exchange ring is laid out [slot][column][level] (LS floats per column) instead of
// [level][slot][column], so that a thread fetches the left / right neighbours of FOUR time levels with one LDS.128 and
// publishes its T new values with T/4 STS.128: 2 x ceil(T/4) + ceil(T/4) shared-memory instructions per step instead
// of 3T.  All levels of a step read the same ring slot (written two steps earlier) and write the same slot, so the
// batching changes no dependency.  LS = 12 floats (T > 4): 8 consecutive columns x 16 bytes then fall into 8 distinct
// bank groups (12 i mod 32), i.e. every quarter-warp access is conflict-free.
__host__ __device__ constexpr int rolled_ls(int T) { return T <= 4 ? 4 : 12; }

template <int T, int BW, int SYNC, bool QG = false>
__global__ void __launch_bounds__(BW, 1) solver_rolled_kernel(const float* __restrict__ coefA,
    const float* __restrict__ coefB, const float* __restrict__ u_src, float* __restrict__ u_dst,
    const float* __restrict__ o_src, float* __restrict__ o_dst, int W, int H, int chunk_rows, int first_rows, float step,
    float mom, int permute)
{
    constexpr int HALO = rolled_halo(T);
    constexpr int S = BW - 2 * HALO;   // columns stored per band
    constexpr int PF = rolled_pf(T);   // staging ring depth (rows in flight from HBM); a multiple of 4
    constexpr int NX = 2 * T + 4;      // coefficient window: rows s0-2T .. s0+3 of a group starting at step s0
    static_assert(BW % 64 == 0 && PF % 4 == 0 && PF >= 8 && T >= 2, "geometry");
    extern __shared__ __align__(16) float smem_raw[];
    // exchange ring: T*4 rows of RW = BW + 8 floats; the 8 floats between two rows are never written, so the band's
    // first / last three threads read 0.0f for their out-of-band neighbours
    // (QG: 4 slots x RW columns x LS floats; column tid lives at physical column tid + 4, same zero margins)
    constexpr int RW = BW + 8;
    constexpr int LS = rolled_ls(T);
    constexpr int RING = QG ? 4 * RW * LS : T * 4 * RW + 8;   // floats
    float* sm = smem_raw + 4;
    float* stage = smem_raw + RING;   // [PF][4 arrays][BW]
    constexpr int SLOT = 4 * BW;                // floats per staging slot

    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    // Column block of a warp.  permute: the warps of one scheduler (hardware warp slot % 4; a CTA owns the SM, so slot =
    // warp index) take ADJACENT column blocks, so that a warp's exchange partners sit on its own scheduler except at
    // three block boundaries: when a warp waits at the pair barrier, the partner it waits for inherits its issue slots.
    int tid = threadIdx.x;
    if (permute) {
        const int pw = tid >> 5, sch = pw & 3;
        int before = 0;
#pragma unroll
        for (int j = 0; j < 3; ++j)
            before += j < sch ? (BW / 32 - j + 3) / 4 : 0;
        tid = (before + (pw >> 2)) * 32 + (tid & 31);
    }
    const int L = 3 * W;
    const int g0 = blockIdx.x * S - HALO;
    const int gi = g0 + tid;
    const int r0 = blockIdx.y == 0 ? 0 : first_rows + (blockIdx.y - 1) * chunk_rows;
    const int r1 = min(H, blockIdx.y == 0 ? first_rows : r0 + chunk_rows);
    const bool col_ok = gi >= 0 && gi < L;
    const bool store_col = col_ok && tid >= HALO && tid < HALO + S;
    // publishes to the neighbour-exchange ring: image columns except the last pixel column (never a valid right
    // neighbour: x+1 < W-1, flowconsistency.cu:215); columns outside the image never publish, their slots stay zero
    const bool pub_ok = col_ok && gi < 3 * (W - 1);
    const int nsteps = (r1 - r0) + 3 * T;

    float win[T][4], uu[T][2], XA[NX], XB[NX];
#pragma unroll
    for (int t = 0; t < T; ++t) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
            win[t][j] = 0.0f;
        uu[t][0] = uu[t][1] = 0.0f;
    }
#pragma unroll
    for (int j = 0; j < NX; ++j) {
        XA[j] = 0.0f;
        XB[j] = 0.0f;
    }
    for (int i = tid; i < RING; i += BW)
        smem_raw[i] = 0.0f;

    // ---- warp-cooperative staging: the 32 columns of a warp x 4 images are 32 chunks of 16 bytes, one per lane
    // (lane l copies columns [g0 + 32*warp + 4*(l%8), +4) of image l/8); chunks lie entirely inside or outside the
    // image (3W % 4 == 0, g0 % 4 == 0); rows / columns outside the image are zero-filled (src-size 0)
    const int lane = tid & 31;
    const int arr = lane >> 3;
    const int wcol = (tid & ~31) + 4 * (lane & 7);
    const int gcol = g0 + wcol;
    const bool chunk_ok = gcol >= 0 && gcol < L;
    const float* const my_src = arr == 0 ? o_src : arr == 1 ? u_src : arr == 2 ? coefA : coefB;
    int my_eoff = (r0 - T) * L + (chunk_ok ? gcol : 0);   // element offset of (next requested row, my chunk)
    const unsigned my_dst = static_cast<unsigned>(__cvta_generic_to_shared(stage + arr * BW + wcol));
    // (u_src == nullptr: the momentum image is all zeros -- the first pass of a solve -- and is not read at all)
    const unsigned chunk_bytes = (chunk_ok && my_src != nullptr) ? 16u : 0u;
    auto request = [&](int y, int slot, bool commit) {
        const unsigned n = (y >= 0 && y < H) ? chunk_bytes : 0u;
        const unsigned d = my_dst + static_cast<unsigned>(slot * SLOT * sizeof(float));
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(my_src + my_eoff), "r"(n) : "memory");
        if (commit)
            asm volatile("cp.async.commit_group;" ::: "memory");
        my_eoff += L;
    };

    asm volatile("griddepcontrol.wait;" ::: "memory");   // the previous kernel of the stream wrote our inputs
#pragma unroll
    for (int j = 0; j < PF - 2; ++j)   // rows of steps 0 .. PF-3 as (PF-2)/2 groups of two
        request(r0 - T + j, j, (j & 1) == 1);
    asm volatile("cp.async.wait_group %0;" ::"n"((PF - 4) / 2) : "memory");   // the first group has landed
    __syncthreads();

    // element offset of (row y_in - 2T, column gi): where level T stores at this step; dereferenced only under the
    // store predicate
    int so = (r0 - 3 * T) * L + (col_ok ? gi : 0);

    constexpr int NW = BW / 32;
    const int warp = tid >> 5;
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

    // One step: levels T..1 (level T first: it reads the window row the arrival of step k+... never touches), then
    // the arrival of the level-0 row.  k in 0..3 = step within the group; st = staging slot of this step.
    // neighbours of all levels for step k: slot (k+2)&3, columns tid-3 and tid+3, levels 0..T-1 (level t reads ring t-1)
    [[maybe_unused]] auto load_nb = [&](const int k, float* lfq, float* rtq) {
        constexpr int NQ = (T + 3) / 4;
        const float* nb = smem_raw + ((((k + 2) & 3) * RW + tid + 4) * LS);
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            const float4 l4 = *reinterpret_cast<const float4*>(nb - 3 * LS + 4 * q);
            const float4 r4 = *reinterpret_cast<const float4*>(nb + 3 * LS + 4 * q);
            lfq[4 * q] = l4.x; lfq[4 * q + 1] = l4.y; lfq[4 * q + 2] = l4.z; lfq[4 * q + 3] = l4.w;
            rtq[4 * q] = r4.x; rtq[4 * q + 1] = r4.y; rtq[4 * q + 2] = r4.z; rtq[4 * q + 3] = r4.w;
        }
    };
    auto step_body = [&](auto rowmask_tag, const int k, const int y_in, const float* st, const float* pre = nullptr) {
        constexpr bool ROWMASK = decltype(rowmask_tag)::value;
        constexpr int NQ = (T + 3) / 4;
        [[maybe_unused]] float lfq[QG ? NQ * 4 : 1], rtq[QG ? NQ * 4 : 1];
        if constexpr (QG) {
            if (pre != nullptr) {   // (compile-time: the interval loaded both steps' neighbours up front)
#pragma unroll
                for (int i = 0; i < NQ * 4; ++i) {
                    lfq[i] = pre[i];
                    rtq[i] = pre[NQ * 4 + i];
                }
            } else {
                load_nb(k, lfq, rtq);
            }
        }
#pragma unroll
        for (int t = T; t >= 1; --t) {
            const int rho = y_in - 2 * t;
            const float c = win[t - 1][(k + 2) & 3];   // produced at step s-2
            float up = win[t - 1][(k + 1) & 3];        // s-3
            float dn = win[t - 1][(k + 3) & 3];        // s-1
            float lf, rt;
            if constexpr (QG) {
                lf = lfq[t - 1];
                rt = rtq[t - 1];
            } else {
                const float* row = sm + ((t - 1) * 4 + ((k + 2) & 3)) * RW + tid;
                lf = row[-3];
                rt = row[3];
            }
            if constexpr (ROWMASK) {
                dn = (rho + 1) < (H - 1) ? dn : 0.0f;  // (flowconsistency.cu:227)
                up = rho >= 1 ? up : 0.0f;             // (:232)
            }
            const float Ssum = ((rt + lf) + dn) + up;
            const float a = XA[2 * T + k - 2 * t];     // row s-2t of the coefficient window
            const float b = XB[2 * T + k - 2 * t];
            const float uo = uu[t - 1][k & 1];         // produced at step s-2
            const float un = __fmaf_rn(step, Ssum, __fmaf_rn(a, c, b));
            const float on = __fmaf_rn(mom, uo, c + un);
            if (t < T) {
                win[t % T][k & 3] = on;   // (t % T only silences the bounds warning for t == T)
                uu[t % T][k & 1] = un;
                if constexpr (!QG) {
                    if (pub_ok)
                        sm[((t % T) * 4 + (k & 3)) * RW + tid] = on;
                }
            } else {
                // predicated, not branched around: a branch here would end the basic block
                const int ok = store_col && rho >= r0 && rho < r1;
                asm volatile("{\n .reg .pred p;\n setp.ne.s32 p, %4, 0;\n @p st.global.f32 [%0], %2;\n"
                             " @p st.global.f32 [%1], %3;\n}" ::"l"(o_dst + so), "l"(u_dst + so), "f"(on), "f"(un), "r"(ok));
            }
        }
        // level 0: the row of this step landed, and became visible, before the barrier that opened the interval
        const float n_o = st[0];
        win[0][k & 3] = n_o;
        uu[0][k & 1] = st[BW];
        if constexpr (QG) {
            // publish the T new values of this column (levels 0..T-1) into slot k&3
            if (pub_ok) {
                float* pb = smem_raw + (((k & 3) * RW + tid + 4) * LS);
#pragma unroll
                for (int q = 0; q < NQ; ++q) {
                    float4 v;
                    v.x = win[(4 * q) % T][k & 3];
                    v.y = 4 * q + 1 < T ? win[(4 * q + 1) % T][k & 3] : 0.0f;
                    v.z = 4 * q + 2 < T ? win[(4 * q + 2) % T][k & 3] : 0.0f;
                    v.w = 4 * q + 3 < T ? win[(4 * q + 3) % T][k & 3] : 0.0f;
                    *reinterpret_cast<float4*>(pb + 4 * q) = v;
                }
            }
        } else {
            if (pub_ok)
                sm[(k & 3) * RW + tid] = n_o;
        }
        XA[2 * T + k] = st[2 * BW];
        XB[2 * T + k] = st[3 * BW];
        so += L;
    };

    // One group of four steps = two barrier intervals.  slot = staging slot of the group's first step (a multiple
    // of 4; PF is one too).  EDGE: the masked copy, also taken by a group that nsteps cuts short.
    auto run_group = [&](auto rowmask_tag, const int base, const int y_first, const int slot) {
        constexpr bool EDGE = decltype(rowmask_tag)::value;
        const float* st = stage + slot * SLOT + tid;
#pragma unroll
        for (int k = 0; k < 4; k += 2) {
            if (!EDGE || base + k < nsteps) {   // uniform across the CTA
                // rows of steps s+PF-2, s+PF-1 into the slots read at steps s-2, s-1 (all lanes are past the barrier)
                const int rq = slot + k - 2 < 0 ? PF - 2 : slot + k - 2;
                request(y_first + k + PF - 2, rq, false);
                request(y_first + k + PF - 1, rq + 1, true);
                if constexpr (QG && VSC_SOLVER_PREFETCH_NB) {
                    // both steps' neighbour quads are visible since the barrier that opened the interval: the second
                    // step's loads are in flight while the first step computes
                    constexpr int NQ4 = (T + 3) / 4 * 4;
                    float nb1[2 * NQ4];
                    load_nb(k + 1, nb1, nb1 + NQ4);
                    step_body(rowmask_tag, k, y_first + k, st + k * SLOT);
                    if (!EDGE || base + k + 1 < nsteps)
                        step_body(rowmask_tag, k + 1, y_first + k + 1, st + (k + 1) * SLOT, nb1);
                } else {
                    step_body(rowmask_tag, k, y_first + k, st + k * SLOT);
                    if (!EDGE || base + k + 1 < nsteps)
                        step_body(rowmask_tag, k + 1, y_first + k + 1, st + (k + 1) * SLOT);
                }
                // my copies of the next interval's two rows: all but the (PF-4)/2 youngest groups
                asm volatile("cp.async.wait_group %0;" ::"n"((PF - 4) / 2) : "memory");
                ring_sync();
            }
        }
    };

    int slot = 0;
    for (int base = 0; base < nsteps; base += 4) {
        const int y_first = r0 - T + base;   // y_in of the group's first step
        // the masked copy is needed while some level works on rows <= 0 or >= H-2
        const bool edge = y_first <= 2 * T || y_first + 3 >= H || base + 4 > nsteps;
        if (edge)
            run_group(std::true_type{}, base, y_first, slot);
        else
            run_group(std::false_type{}, base, y_first, slot);
        // slide the coefficient window down by the four rows of this group
#pragma unroll
        for (int j = 0; j < 2 * T; ++j) {
            XA[j] = XA[j + 4];
            XB[j] = XB[j + 4];
        }
        slot = slot + 4 == PF ? 0 : slot + 4;
    }
}

struct RolledGeom {
    int bw, nb, nc, chunk_rows, first_rows;
    long long cost;
};

// grid of one band width: bands x row chunks, at most ONE wave (1 CTA per SM); cost ~ per-SM time
static RolledGeom rolled_geom(int T, int BW, int L, int H, int sms)
{
    RolledGeom g;
    g.bw = BW;
    const int S = BW - 2 * rolled_halo(T);
    g.nb = (L + S - 1) / S;
    int nc = sms / g.nb;
    if (nc < 1) nc = 1;
    const int min_rows = 4 * T;  // below this the 3T-step pipeline fill dominates
    if (nc > (H + min_rows - 1) / min_rows) nc = (H + min_rows - 1) / min_rows;
    if (nc < 1) nc = 1;
    // the first / last chunk run the masked copy for 3T / 2T steps: shorter by default (profiles/r2_solver_sweep_pairs_edges.txt)
    const int et = g_stream_edge_top, eb = g_stream_edge_bot;
    int top = et >= 0 ? et : 12, bot = eb >= 0 ? eb : 6;   // (12 / 6 since the quad-gather ring: profiles/r2_solver_qg_sweep.txt)
    if (nc < 3 || H < nc * (2 * T + top + bot)) top = bot = 0;
    // H = (mid - top) + (nc - 2) * mid + last,  last <= mid - bot
    g.chunk_rows = (H + top + bot + nc - 1) / nc;
    g.first_rows = g.chunk_rows - top;
    g.nc = H <= g.first_rows ? 1 : 1 + (H - g.first_rows + g.chunk_rows - 1) / g.chunk_rows;
    const long long waves = (static_cast<long long>(g.nb) * g.nc + sms - 1) / sms;
    // measured time of one step in ns per band width (not proportional: a narrow band has fewer warps to hide its
    // per-step latency), scaled by the number of levels
    const int step_ns = (BW >= 512 ? 405 : BW >= 448 ? 368 : BW >= 384 ? 315 : 228) * (T + 2) / 10;
    g.cost = waves * (g.chunk_rows + 3 * T) * step_ns;
    return g;
}

template <int T, int BW, int SYNC, bool QG = false>
static int launch_rolled_impl(const RolledGeom& g, const float* coefA, const float* coefB, const float* u_src,
    float* u_dst, const float* o_src, float* o_dst, int W, int H, float step, float mom, cudaStream_t st)
{
    constexpr int PF = rolled_pf(T);
    const size_t ring = QG ? static_cast<size_t>(4) * (BW + 8) * rolled_ls(T) : static_cast<size_t>(T) * 4 * (BW + 8) + 8;
    const size_t smem = (ring + static_cast<size_t>(PF) * 4 * BW) * sizeof(float);
    static unsigned long long configured = 0;
    if (const int e = ensure_dynamic_smem(solver_rolled_kernel<T, BW, SYNC, QG>, smem, false, configured))
        return e;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(g.nb, g.nc);
    cfg.blockDim = dim3(BW);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = g_pdl ? 1 : 0;
    const cudaError_t e = cudaLaunchKernelEx(&cfg, solver_rolled_kernel<T, BW, SYNC, QG>, coefA, coefB, u_src, u_dst, o_src,
        o_dst, W, H, g.chunk_rows, g.first_rows, step, mom, static_cast<int>(g_stream_permute));
    count_launch();
    return e == cudaSuccess ? launch_status() : static_cast<int>(e);
}

template <int T, int BW>
static int launch_rolled(const RolledGeom& g, const float* coefA, const float* coefB, const float* u_src, float* u_dst,
    const float* o_src, float* o_dst, int W, int H, float step, float mom, cudaStream_t st)
{
    if constexpr (T % 2 == 0) {   // the CTA-barrier form (a test hook) is built for the even depths only
        if (!g_stream_pair)
            return launch_rolled_impl<T, BW, 0>(g, coefA, coefB, u_src, u_dst, o_src, o_dst, W, H, step, mom, st);
    }
    // quad-gather ring where its shared memory fits (not T = 5..8 with 512-float bands)
    constexpr bool qg_fits = (static_cast<size_t>(4) * (BW + 8) * rolled_ls(T) + static_cast<size_t>(rolled_pf(T)) * 4 * BW)
            * sizeof(float) <= 227u * 1024u;
    if constexpr (qg_fits) {
        if (g_stream_qg) {
            return launch_rolled_impl<T, BW, 1, true>(g, coefA, coefB, u_src, u_dst, o_src, o_dst, W, H, step, mom, st);
        }
    }
    return launch_rolled_impl<T, BW, 1>(g, coefA, coefB, u_src, u_dst, o_src, o_dst, W, H, step, mom, st);
}

// band widths per depth: what the register file allows (T = 10: 40 + 20 + 48 state registers -> 168 per thread ->
// 384 threads; T <= 8 fits 128 -> 512 threads)
template <int T>
static int launch_rolled_best(const float* coefA, const float* coefB, const float* u_src, float* u_dst,
    const float* o_src, float* o_dst, int W, int H, float step, float mom, cudaStream_t st)
{
    const int L = 3 * W, sms = sm_count();
    constexpr int NCAND = 4;
    const int cands[NCAND] = {512, 448, 384, 256};
    constexpr int first = T >= 9 ? 2 : 0;   // T = 9, 10: 384 and 256 only
    int best = first;
    RolledGeom bg = rolled_geom(T, cands[first], L, H, sms);
    if (g_stream_band >= 1 && g_stream_band <= NCAND) {
        best = g_stream_band - 1 < first ? first : g_stream_band - 1;
        bg = rolled_geom(T, cands[best], L, H, sms);
    } else {
        for (int i = first + 1; i < NCAND; ++i) {
            const RolledGeom g = rolled_geom(T, cands[i], L, H, sms);
            if (g.cost < bg.cost) {
                bg = g;
                best = i;
            }
        }
    }
    if constexpr (T < 9) {
        if (best == 0)
            return launch_rolled<T, 512>(bg, coefA, coefB, u_src, u_dst, o_src, o_dst, W, H, step, mom, st);
        if (best == 1)
            return launch_rolled<T, 448>(bg, coefA, coefB, u_src, u_dst, o_src, o_dst, W, H, step, mom, st);
    }
    if (best == 2)
        return launch_rolled<T, 384>(bg, coefA, coefB, u_src, u_dst, o_src, o_dst, W, H, step, mom, st);
    return launch_rolled<T, 256>(bg, coefA, coefB, u_src, u_dst, o_src, o_dst, W, H, step, mom, st);
}

// can this kernel take the passes of a solve at all (so that the planner may use odd depths)?
bool solver_rolled_takes(int W, const void* a, const void* b, const void* c, const void* d)
{
    return g_stream_rolled && g_stream_coop && g_stream_pair && (3LL * W) % 4 == 0 && aligned16(a) && aligned16(b)
        && aligned16(c) && aligned16(d);
}

// true if this kernel can run the pass (then *rc is its status); false: the caller uses stab_solver_stream.cu
bool solver_rolled_pass(int T, const float* coefA, const float* coefB, const float* u_src, float* u_dst,
    const float* o_src, float* o_dst, int W, int H, float step, float mom, cudaStream_t st, int* rc)
{
    if (!g_stream_rolled || !g_stream_coop)
        return false;
    if ((3LL * W) % 4 != 0 || !aligned16(coefA) || !aligned16(coefB) || !aligned16(u_src) || !aligned16(o_src))
        return false;   // (a null u_src counts as aligned)
    if (3LL * W * (static_cast<long long>(H) + 64) >= 0x7fffffffLL)   // 32-bit element offsets in the kernel
        return false;
    // 4K-class images: a CTA runs 570 steps per pass, the fully unrolled loop stays resident and its 8 % fewer
    // instructions win (201.6 vs 205.1 us per pass, profiles/r2_solver_sweep_rolled.txt); g_stream_rolled == 2 forces
    // this kernel regardless (tests)
    if (g_stream_rolled == 1 && T == 10 && static_cast<long long>(W) * H >= 4000000LL)
        return false;
    switch (T) {
        case 10: *rc = launch_rolled_best<10>(coefA, coefB, u_src, u_dst, o_src, o_dst, W, H, step, mom, st); return true;
        case 8: *rc = launch_rolled_best<8>(coefA, coefB, u_src, u_dst, o_src, o_dst, W, H, step, mom, st); return true;
        case 6: *rc = launch_rolled_best<6>(coefA, coefB, u_src, u_dst, o_src, o_dst, W, H, step, mom, st); return true;
        case 4: *rc = launch_rolled_best<4>(coefA, coefB, u_src, u_dst, o_src, o_dst, W, H, step, mom, st); return true;
        case 2: *rc = launch_rolled_best<2>(coefA, coefB, u_src, u_dst, o_src, o_dst, W, H, step, mom, st); return true;
        case 9: *rc = launch_rolled_best<9>(coefA, coefB, u_src, u_dst, o_src, o_dst, W, H, step, mom, st); return true;
        case 7: *rc = launch_rolled_best<7>(coefA, coefB, u_src, u_dst, o_src, o_dst, W, H, step, mom, st); return true;
        case 5: *rc = launch_rolled_best<5>(coefA, coefB, u_src, u_dst, o_src, o_dst, W, H, step, mom, st); return true;
        case 3: *rc = launch_rolled_best<3>(coefA, coefB, u_src, u_dst, o_src, o_dst, W, H, step, mom, st); return true;
        default: return false;
    }
}

}  // namespace vsc
