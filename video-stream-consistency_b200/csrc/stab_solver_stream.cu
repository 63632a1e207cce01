// Temporally blocked solver sweep for sm_100a: T Jacobi sweeps of the screened-Poisson update per launch,
// with every intermediate sweep kept on chip (reference loop: flowconsistency.cu:367-372, one launch and one
// full HBM round trip of 7 images per sweep).
//
// The unblocked sweep (stab_solver.cu) already runs at the HBM roofline, so the only way to go faster is to
// not touch HBM between sweeps.  Scheme (validated bit-for-bit against plain Jacobi by the NumPy emulation in
// tests/emul_stream_solver.py, of which this kernel is a transliteration):
//
//   * the image is the 2-D float array [H][L], L = 3W (interleaved channels; x-neighbours are +-3 floats);
//   * a CTA owns a band of BW = blockDim.x consecutive floats and STREAMS down a chunk of rows; thread `tid`
//     owns column g0+tid for ALL T time levels, held in registers:
//         win[t][4]  level-t values of the last 4 rows      uu[t][4]  the matching momentum terms
//         Ar/Br[2T]  the folded coefficients of the last 2T rows
//   * at step s the level-0 row y_in = r0-T+s arrives from HBM; level t (1..T) computes row y_in-2t from the
//     level t-1 rows y_in-2t-1 .. y_in-2t+1, which level t-1 produced at steps s-3, s-2, s-1.  The skew of
//     TWO rows per level makes the T updates of a step mutually independent: ONE barrier per step, T
//     independent dependency chains per thread (ILP), no redundant work in y inside a chunk;
//   * level-0 rows (out, u, A, B) are fetched 8 steps ahead with cp.async into a thread-private staging ring
//     (HBM latency is several steps long; with a 1-step register prefetch the kernel was latency-bound and
//     slower than the unblocked sweeps -- profiles/r1_notes.md);
//   * up/down neighbours are the thread's own registers; left/right neighbours (+-3 floats, other threads)
//     come from a shared-memory ring written two steps earlier: 2 LDS.32 + 1 STS.32 per value and sweep;
//   * level-T rows inside the chunk and inside the band's valid cone (3T floats from each band edge, T rows
//     from each chunk edge are halo) are stored.  HBM traffic per sweep drops from 24 B/value to
//     ~24/(T * efficiency) B/value; efficiency = (1-6T/BW) * rows/(rows+3T) ~ 0.8 at T=8, BW=512.
//
// Rings are indexed by (step mod 4) / (step mod 2T); the step loop is unrolled 2T times so every ring index
// is a compile-time constant and the rings live in registers without moves.
#include <type_traits>

#include "vsc_common.cuh"

namespace vsc {

// private staging depth: divides the unroll factor 2T (T in {2, 4, 6, 8})
__host__ __device__ constexpr int stream_private_pf(int T) { return T == 6 ? 6 : (T == 2 ? 4 : (T == 10 ? 10 : 8)); }
__host__ __device__ constexpr int stream_halo(int T) { return (3 * T + 3) / 4 * 4; }
std::atomic<bool> g_stream_coop = true;          // warp-cooperative 16-byte staging when the images allow it
std::atomic<bool> g_stream_pair = true;          // neighbour-pair named barriers instead of a CTA-wide barrier
std::atomic<int> g_stream_band = 0;              // 0 = cost model; 1..4 force a band candidate (benchmarks, vsc_set_solver_mode)

// BW = band width in floats = threads per CTA (one CTA per SM; the launcher picks the BW that fills the SMs best.
// Two 256-wide CTAs per SM were measured slower than every single-CTA geometry at every size from 640x360 to
// 4K -- profiles/r1_sweep_bands.txt -- and are not built)
// COOP: level-0 rows staged warp-cooperatively with 16-byte cp.async (one chunk per lane and row; needs
// 3W % 4 == 0 and 16-byte aligned images); otherwise every thread stages its own four floats (4-byte cp.async).
// SYNC: how the exchange ring is synchronised: 0 = CTA-wide barrier, 1 = between neighbouring warps only (named
// barriers).  A third scheme -- per-warp progress counters polled with ld.acquire, so that nobody waits for a warp
// that is behind -- was measured 35 % SLOWER (253 vs 186 us at 4K, profiles/r1_sweep_bands_flags_vs_barriers.txt):
// the per-step poll sits on every warp's critical path, the named barriers are nearly free.
template <int T, int BW, bool COOP, int SYNC>
__global__ void __launch_bounds__(BW, 1) solver_stream_kernel(const float* __restrict__ coefA,
    const float* __restrict__ coefB, const float* __restrict__ u_src, float* __restrict__ u_dst,
    const float* __restrict__ o_src, float* __restrict__ o_dst, int W, int H, int chunk_rows, float step, float mom)
{
    // band halo: 3 floats per sweep, rounded up to a multiple of 4 so that band starts stay 16-byte aligned
    constexpr int HALO = stream_halo(T);
    constexpr int S = BW - 2 * HALO;   // columns stored per band
    constexpr int U = 2 * T;        // unroll: lcm(4, 2T) for T in {4, 8}
    static_assert(U % 4 == 0, "ring period");
    constexpr int PF = COOP ? U : stream_private_pf(T);  // staging ring depth (rows in flight from HBM)
    static_assert(BW % 64 == 0 && U % PF == 0 && PF >= 4, "staging geometry: slot index = step mod PF = k mod PF");
    extern __shared__ float smem_raw[];
    // exchange ring: T*4 rows of RW = BW + 8 floats.  The 8 floats between two rows (4 behind one row, 4 before the
    // next) are never written: the band's first / last three threads read their out-of-band neighbours tid-3 /
    // tid+3 there and get 0.0f (their columns are halo either way; without the padding they read another row's
    // live values -- harmless, but a nondeterministic read that compute-sanitizer rightly reports as a hazard)
    constexpr int RW = BW + 8;
    float* sm = smem_raw + 4;
    // thread-private staging ring for the level-0 rows: stage[slot][array][tid], filled by cp.async PF steps
    // ahead (each thread copies and later reads only its own 4 floats per row: no cross-thread ordering needed)
    float* stage = smem_raw + T * 4 * RW + 8;

    // Programmatic dependent launch: let the next pass of the stream start placing its CTAs as ours retire (its
    // prologue -- index set-up, zeroing of the exchange ring -- then overlaps our tail); it blocks in
    // griddepcontrol.wait below until this whole grid has completed and flushed, before it touches the images.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const int tid = threadIdx.x;
    const int L = 3 * W;
    const int g0 = blockIdx.x * S - HALO;
    const int gi = g0 + tid;
    const int r0 = blockIdx.y * chunk_rows;
    const int r1 = min(H, r0 + chunk_rows);
    const bool col_ok = gi >= 0 && gi < L;
    const bool store_col = col_ok && tid >= HALO && tid < HALO + S;
    // publishes to the neighbour-exchange ring: image columns except the last pixel column (see step_body)
    const bool pub_ok = col_ok && gi < 3 * (W - 1);
    const int nsteps = (r1 - r0) + 3 * T;
    const size_t colofs = static_cast<size_t>(col_ok ? gi : 0);

    float win[T][4], uu[T][4], Ar[U], Br[U];
#pragma unroll
    for (int t = 0; t < T; ++t)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            win[t][j] = 0.0f;
            uu[t][j] = 0.0f;
        }
#pragma unroll
    for (int j = 0; j < U; ++j) {
        Ar[j] = 0.0f;
        Br[j] = 0.0f;
    }
    // the ring slots read before they are first written belong to the invalid cone; zero them so that no
    // NaN/Inf garbage is ever combined (finite garbage is harmless: it never reaches a stored value)
    for (int i = tid; i < T * 4 * RW + 8; i += BW)
        smem_raw[i] = 0.0f;

    // HBM latency (~1-2 us under load) is several steps long: keep PF rows in flight per thread with cp.async
    // into the staging ring; rows (or columns) outside the image are zero-filled (src-size 0: the source is
    // not read, so its address may lie outside the arrays).
    // Addressing: ONE running per-thread 32-bit ELEMENT offset `eoff` = offset of (row y_in, column gi) in any of
    // the images, advanced by one row per step (the launcher guarantees that an image plus the pipeline's
    // look-ahead holds fewer than 2^31 floats); the prefetch (PF rows ahead) and store (2T rows behind) row shifts
    // are folded into CTA-uniform base pointers, so forming an address is one IMAD.WIDE.
    const unsigned stage_base = static_cast<unsigned>(__cvta_generic_to_shared(stage + tid));
    int eoff = (r0 - T) * L + static_cast<int>(colofs);
    // (row shifts are applied to the 32-bit offset, not to the base pointers: with pre-shifted pointers the
    // compiler re-associates the sum and forms the address with eight 64-bit instructions instead of three)
    const int store_shift = 2 * T * L;

    // ---- private staging (COOP == false): 4 x 4-byte cp.async per thread and row, no cross-thread ordering
    auto copy_row = [&](const float* bo, const float* bu, const float* ba, const float* bb, int off, int y, int slot) {
        const unsigned n = (col_ok && y >= 0 && y < H) ? 4u : 0u;
        const unsigned d = stage_base + static_cast<unsigned>(slot * 4 * BW * sizeof(float));
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d), "l"(bo + off), "r"(n) : "memory");
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d + BW * 4), "l"(bu + off), "r"(bu ? n : 0u)
                     : "memory");   // (u_src == nullptr: zero momentum, not read)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d + 2 * BW * 4), "l"(ba + off), "r"(n)
                     : "memory");
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d + 3 * BW * 4), "l"(bb + off), "r"(n)
                     : "memory");
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    const int pf_shift = PF * L;

    // ---- warp-cooperative staging (COOP == true): the 32 columns of a warp x 4 images are 32 chunks of 16 bytes,
    // exactly one per lane: ONE 16-byte cp.async per thread and step instead of four 4-byte ones.  Lane l copies
    // columns [g0 + 32*warp + 4*(l%8), +4) of image l/8; chunks lie entirely inside or outside the image
    // (3W % 4 == 0, g0 % 4 == 0).  Producer and consumers of a chunk are lanes of the SAME warp, so the only
    // ordering needed is cp.async.wait_group + __syncwarp: a row is requested PF-1 steps early, into the slot
    // the warp read one step earlier.
    const int lane = tid & 31;
    const int arr = lane >> 3;
    const int wcol = (tid & ~31) + 4 * (lane & 7);   // first column (inside the band) of my chunk
    const int gcol = g0 + wcol;
    const bool chunk_ok = gcol >= 0 && gcol < L;
    const float* const my_src = arr == 0 ? o_src : arr == 1 ? u_src : arr == 2 ? coefA : coefB;
    // row of the NEXT request (PF-1 rows ahead of step s), first column of my chunk
    int my_eoff = (r0 - T + PF - 1) * L + (chunk_ok ? gcol : 0);
    const unsigned my_dst = static_cast<unsigned>(__cvta_generic_to_shared(stage + arr * BW + wcol));
    const unsigned chunk_bytes = (chunk_ok && my_src != nullptr) ? 16u : 0u;   // (null u_src: zero momentum, not read)
    auto coop_copy = [&](const float* src, bool row_ok, int slot) {
        const unsigned n = row_ok ? chunk_bytes : 0u;   // src-size 0: zero fill, the source is not read
        const unsigned d = my_dst + static_cast<unsigned>(slot * 4 * BW * sizeof(float));
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(n) : "memory");
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    asm volatile("griddepcontrol.wait;" ::: "memory");   // the previous kernel of the stream wrote our inputs
    if constexpr (COOP) {
#pragma unroll
        for (int j = 0; j < PF - 1; ++j)   // rows of steps 0 .. PF-2
            coop_copy(my_src + (my_eoff - (PF - 1 - j) * L), r0 - T + j >= 0 && r0 - T + j < H, j);
    } else {
#pragma unroll
        for (int j = 0; j < PF; ++j)       // rows of steps 0 .. PF-1
            copy_row(o_src, u_src, coefA, coefB, eoff + j * L, r0 - T + j, j);
    }
    __syncthreads();

    // Synchronisation of the neighbour-exchange ring.  A thread only ever reads columns tid+-3, i.e. its own
    // warp or an adjacent one, so instead of a CTA-wide barrier each warp synchronises with its two neighbours
    // through named barriers (id i = the boundary between warps i-1 and i, 64 threads each).  Even warps take
    // left then right, odd warps right then left, so all boundaries complete in two rounds instead of rippling.
    // Warps that are not neighbours may drift apart, which spreads the LDS-heavy and FMA-heavy phases of the
    // step over time instead of having all warps hit the same pipe at once.
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
    // One step of the pipeline.  ROWMASK selects the variant that applies the reference's top/bottom inclusion
    // tests (flowconsistency.cu:227,232); it is needed only while some level works on rows <= 0 or >= H-2, a
    // CTA-uniform condition true for a handful of steps of the first and last row chunk.
    // The left/right tests (:215,:221) cost nothing: a column outside the image, and the last pixel column
    // (never a valid right neighbour: x+1 < W-1), simply never PUBLISH to the shared ring, whose slots were
    // zeroed above -- so their readers add 0.0f, which is exactly the excluded term.
    auto step_body = [&](auto rowmask_tag, const int k, const int y_in) {
        constexpr bool ROWMASK = decltype(rowmask_tag)::value;
        // levels T..1 (level T first: it reads the coefficient slot the arrival below overwrites)
#pragma unroll
        for (int t = T; t >= 1; --t) {
            const int rho = y_in - 2 * t;
            const float c = win[t - 1][(k + 2) & 3];   // produced at step s-2
            float up = win[t - 1][(k + 1) & 3];        // s-3
            float dn = win[t - 1][(k + 3) & 3];        // s-1
            const float* row = sm + ((t - 1) * 4 + ((k + 2) & 3)) * RW + tid;
            const float lf = row[-3];
            const float rt = row[3];
            if constexpr (ROWMASK) {
                dn = (rho + 1) < (H - 1) ? dn : 0.0f;  // (:227)
                up = rho >= 1 ? up : 0.0f;             // (:232)
            }
            const float Ssum = ((rt + lf) + dn) + up;
            const float a = Ar[(k + U - 2 * t) % U];
            const float b = Br[(k + U - 2 * t) % U];
            const float uo = uu[t - 1][(k + 2) & 3];
            const float un = __fmaf_rn(step, Ssum, __fmaf_rn(a, c, b));
            const float on = __fmaf_rn(mom, uo, c + un);
            if (t < T) {
                win[t % T][k & 3] = on;   // (t % T only silences the bounds warning for t == T)
                uu[t % T][k & 3] = un;
                if (pub_ok)
                    sm[((t % T) * 4 + (k & 3)) * RW + tid] = on;
            } else if (store_col && rho >= r0 && rho < r1) {
                const int so = eoff - store_shift;   // row rho = y_in - 2T
                o_dst[so] = on;
                u_dst[so] = un;
            }
        }
        // level 0 arrives: the row of step s was requested PF steps ago.  Waiting is done every second step
        // for two rows at once: at an even step at most PF-2 younger groups may still be in flight
        if constexpr (COOP) {
            // the row of step s was requested at step s-PF+1: all but the PF-2 youngest groups must be done;
            // __syncwarp makes the other lanes' chunks visible and closes the previous step's reads of the slot
            // that is refilled below
            asm volatile("cp.async.wait_group %0;" ::"n"(PF - 2) : "memory");
            __syncwarp();
        } else {
            if ((k & 1) == 0)
                asm volatile("cp.async.wait_group %0;" ::"n"(PF - 2) : "memory");
        }
        {
            const float* st = stage + (k % PF) * 4 * BW + tid;
            const float n_o = st[0];
            win[0][k & 3] = n_o;
            uu[0][k & 3] = st[BW];
            if (pub_ok)
                sm[(k & 3) * RW + tid] = n_o;
            Ar[k % U] = st[2 * BW];
            Br[k % U] = st[3 * BW];
        }
        if constexpr (COOP) {
            // request the row of step s+PF-1 into the slot this warp read at step s-1
            // (mask-free steps have y_in > 2T: the requested row can only leave the image at the bottom)
            const int y_req = y_in + PF - 1;
            coop_copy(my_src + my_eoff, ROWMASK ? (y_req >= 0 && y_req < H) : (y_req < H), (k + PF - 1) % PF);
            my_eoff += L;
        } else {
            // refill the slot just consumed (same thread: program order) with the row PF steps ahead
            copy_row(o_src, u_src, coefA, coefB, eoff + pf_shift, y_in + PF, k % PF);
        }
        eoff += L;
        // ONE barrier per TWO steps: a step reads ring slots (s-2)&3 (and its partner (s-1)&3) and writes
        // slot s&3 (partner (s+1)&3) -- disjoint, and what step s needs was published before the barrier
        // that closed step s-1 (worst-case visibility checked in tests/emul_stream_solver.py, sync_every=2)
        if ((k & 1) == 1)
            ring_sync();
    };

    // The ROWMASK body is the general one; whole groups of U steps take it when any of their steps touches the
    // image's top or bottom rows, so that the steady-state loop is ONE contiguous run of the mask-free bodies
    // (the unrolled loop is 34-50 KB of code, more than the 32 KB instruction cache: ncu shows no_instruction
    // stalls, and interleaving the two bodies step by step spread the hot path over twice the address range).
    for (int base = 0; base < nsteps; base += U) {
        const int y_first = r0 - T + base;               // y_in of the group's first step
        const bool edge = y_first <= 2 * T || y_first + U - 1 >= H || base + U > nsteps;
        if (edge) {
#pragma unroll
            for (int k = 0; k < U; ++k)
                if (base + k < nsteps)   // uniform across the CTA
                    step_body(std::true_type{}, k, y_first + k);
        } else {
#pragma unroll
            for (int k = 0; k < U; ++k)
                step_body(std::false_type{}, k, y_first + k);
        }
    }
}

struct StreamGeom {
    int nb, nc, chunk_rows;
    long long cost;
};

// grid of one band width: bands x row chunks, at most ONE wave (1 CTA per SM: a grid of sms+1 CTAs takes twice
// as long as a grid of sms CTAs); cost ~ per-SM time = steps x warps
static StreamGeom stream_geom(int T, int BW, int L, int H, int sms)
{
    StreamGeom g;
    const int S = BW - 2 * stream_halo(T);
    g.nb = (L + S - 1) / S;
    int nc = sms / g.nb;
    if (nc < 1) nc = 1;
    const int min_rows = 4 * T;  // below this the 3T-step pipeline fill dominates
    if (nc > (H + min_rows - 1) / min_rows) nc = (H + min_rows - 1) / min_rows;
    if (nc < 1) nc = 1;
    g.chunk_rows = (H + nc - 1) / nc;
    g.nc = (H + g.chunk_rows - 1) / g.chunk_rows;
    const long long waves = (static_cast<long long>(g.nb) * g.nc + sms - 1) / sms;
    // measured time of one step in ns (profiles/r1_sweep_bands_final.txt, T = 8, 640x360 ... 4K): not proportional to the
    // band width -- a narrow band has fewer warps to hide its per-step latency
    const int step_ns = BW >= 512 ? 405 : BW >= 448 ? 368 : BW >= 384 ? 315 : 228;
    g.cost = waves * (g.chunk_rows + 3 * T) * step_ns;
    return g;
}

template <int T, int BW, bool COOP, int SYNC>
static int launch_stream_impl(const StreamGeom& g, const float* coefA, const float* coefB, const float* u_src,
    float* u_dst, const float* o_src, float* o_dst, int W, int H, float step, float mom, cudaStream_t st)
{
    constexpr int PF = COOP ? 2 * T : stream_private_pf(T);
    const size_t smem = (static_cast<size_t>(T) * 4 * (BW + 8) + 8 + static_cast<size_t>(PF) * 4 * BW) * sizeof(float);
    static unsigned long long configured = 0;
    if (const int e = ensure_dynamic_smem(solver_stream_kernel<T, BW, COOP, SYNC>, smem, false, configured))
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
    const cudaError_t e = cudaLaunchKernelEx(&cfg, solver_stream_kernel<T, BW, COOP, SYNC>, coefA, coefB, u_src, u_dst,
        o_src, o_dst, W, H, g.chunk_rows, step, mom);
    count_launch();
    return e == cudaSuccess ? launch_status() : static_cast<int>(e);
}

template <int T, int BW>
static int launch_stream(const StreamGeom& g, const float* coefA, const float* coefB, const float* u_src,
    float* u_dst, const float* o_src, float* o_dst, int W, int H, float step, float mom, cudaStream_t st)
{
    const bool coop = (3LL * W) % 4 == 0 && aligned16(coefA) && aligned16(coefB) && aligned16(u_src) && aligned16(o_src)
        && g_stream_coop;
    if (coop && g_stream_pair)
        return launch_stream_impl<T, BW, true, 1>(g, coefA, coefB, u_src, u_dst, o_src, o_dst, W, H, step, mom, st);
    if (coop)
        return launch_stream_impl<T, BW, true, 0>(g, coefA, coefB, u_src, u_dst, o_src, o_dst, W, H, step, mom, st);
    return launch_stream_impl<T, BW, false, 0>(g, coefA, coefB, u_src, u_dst, o_src, o_dst, W, H, step, mom, st);
}

template <int T>
static int launch_stream_best(const float* coefA, const float* coefB, const float* u_src, float* u_dst,
    const float* o_src, float* o_dst, int W, int H, float step, float mom, cudaStream_t st)
{
    const int L = 3 * W, sms = sm_count();
    constexpr int NCAND = 4;
    const int cands[NCAND] = {512, 448, 384, 256};
    int best = 0;
    StreamGeom bg = stream_geom(T, cands[0], L, H, sms);
    if (g_stream_band >= 1 && g_stream_band <= NCAND) {
        best = g_stream_band - 1;
        bg = stream_geom(T, cands[best], L, H, sms);
    } else {
        for (int i = 1; i < NCAND; ++i) {
            const StreamGeom g = stream_geom(T, cands[i], L, H, sms);
            if (g.cost < bg.cost) {
                bg = g;
                best = i;
            }
        }
    }
    switch (best) {
        case 0: return launch_stream<T, 512>(bg, coefA, coefB, u_src, u_dst, o_src, o_dst, W, H, step, mom, st);
        case 1: return launch_stream<T, 448>(bg, coefA, coefB, u_src, u_dst, o_src, o_dst, W, H, step, mom, st);
        case 2: return launch_stream<T, 384>(bg, coefA, coefB, u_src, u_dst, o_src, o_dst, W, H, step, mom, st);
        default: return launch_stream<T, 256>(bg, coefA, coefB, u_src, u_dst, o_src, o_dst, W, H, step, mom, st);
    }
}

// Deep passes (T = 10).  Up to T = 8 the time of a step hardly depends on the number of time levels it advances
// (T = 6 -> 8 costs 2 %: the step is bound by its synchronisation / latency chain), so more levels per step are
// cheap throughput as long as the state fits the register file: 10 levels need 168 registers, i.e. bands of at
// most 384 floats.  Measured per SWEEP (profiles/r1_sweep_bands_T8_T10_T12.txt): T = 10 is 4 % faster than T = 8 at
// 4K, 2 % at 1080p, equal at 720p and slower below; T = 12 is slower everywhere (10.6 vs 6.95 us at 1080p: the
// 24-step unrolled body no longer fits the instruction cache) and is not built.
template <int T>
static int launch_stream_deep(const float* coefA, const float* coefB, const float* u_src, float* u_dst,
    const float* o_src, float* o_dst, int W, int H, float step, float mom, cudaStream_t st)
{
    const int L = 3 * W, sms = sm_count();
    constexpr int NCAND = 2;   // 448 floats: 146 registers per thread, T = 10 spills
    const int cands[NCAND] = {384, 256};
    int best = 0;
    StreamGeom bg = stream_geom(T, cands[0], L, H, sms);
    if (g_stream_band >= 1 && g_stream_band <= 4) {
        best = g_stream_band == 4 ? 1 : 0;
        bg = stream_geom(T, cands[best], L, H, sms);
    } else {
        for (int i = 1; i < NCAND; ++i) {
            const StreamGeom g = stream_geom(T, cands[i], L, H, sms);
            if (g.cost < bg.cost) {
                bg = g;
                best = i;
            }
        }
    }
    if (best == 0)
        return launch_stream<T, 384>(bg, coefA, coefB, u_src, u_dst, o_src, o_dst, W, H, step, mom, st);
    return launch_stream<T, 256>(bg, coefA, coefB, u_src, u_dst, o_src, o_dst, W, H, step, mom, st);
}

// T in {2, 4, 6, 8, 10}; returns VSC_E_INVALID for any other value
int solver_stream_pass(int T, const float* coefA, const float* coefB, const float* u_src, float* u_dst,
    const float* o_src, float* o_dst, int W, int H, float step, float mom, cudaStream_t st)
{
    if (3LL * W * (static_cast<long long>(H) + 64) >= 0x7fffffffLL)   // 32-bit element offsets in the kernel
        return VSC_E_INVALID;
    if (T == 10)
        return launch_stream_deep<10>(coefA, coefB, u_src, u_dst, o_src, o_dst, W, H, step, mom, st);
    if (T == 8)
        return launch_stream_best<8>(coefA, coefB, u_src, u_dst, o_src, o_dst, W, H, step, mom, st);
    if (T == 6)
        return launch_stream_best<6>(coefA, coefB, u_src, u_dst, o_src, o_dst, W, H, step, mom, st);
    if (T == 4)
        return launch_stream_best<4>(coefA, coefB, u_src, u_dst, o_src, o_dst, W, H, step, mom, st);
    if (T == 2)
        return launch_stream_best<2>(coefA, coefB, u_src, u_dst, o_src, o_dst, W, H, step, mom, st);
    return VSC_E_INVALID;
}

}  // namespace vsc
