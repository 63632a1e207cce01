"""NumPy emulation of the row-streaming, time-skewed solver kernel (csrc/stab_solver_stream.cu): same
indexing, same rings, same masks, vectorised over the threads of a CTA.  Used to validate the scheme on the CPU
(tests/test_stream_emulation.py) -- the CUDA kernel is a transliteration of `cta()` below.

Scheme: a CTA owns a band of BW consecutive floats of the flattened rows (row length L = 3W; x-neighbours are
+-3 floats away) and streams down the rows [r0-T, r1+T).  Thread `tid` owns column g0+tid for ALL T time levels.
At step s the level-0 row y_in = r0-T+s arrives, and level t (1..T) computes its row y_in-2t from level t-1
rows (y_in-2t-1, y_in-2t, y_in-2t+1), which level t-1 produced at steps s-3, s-2, s-1 (skew 2 => all T levels
of a step are independent).  Left/right neighbours come from a shared ring written at step s-2.
Level T rows inside [r0,r1) x [g0+3T, g0+3T+S) are stored.  The skew also leaves one step of slack in the
shared ring, so ONE barrier per TWO steps is enough (sync_every=2; with 3 the emulation diverges).
"""
import numpy as np


def jacobi(out, u, A, B, W, H, step, mom, sweeps):
    """plain Jacobi sweeps in the folded form (stab_solver.cu: sweep_value), float32, numpy"""
    L = 3 * W
    out = out.copy()
    u = u.copy()
    gi = np.arange(L)
    f = np.float32
    for _ in range(sweeps):
        r = np.zeros_like(out)
        l = np.zeros_like(out)
        d = np.zeros_like(out)
        t = np.zeros_like(out)
        r[:, :-3] = out[:, 3:]
        l[:, 3:] = out[:, :-3]
        d[:-1] = out[1:]
        t[1:] = out[:-1]
        r[:, gi >= 3 * (W - 2)] = 0
        l[:, gi < 3] = 0
        d[np.arange(H) + 1 >= H - 1] = 0
        S = ((r + l) + d) + t
        un = (f(step) * S + (A * out + B)).astype(np.float32)
        on = (f(mom) * u + (out + un)).astype(np.float32)
        out, u = on, un
    return out, u


def cta(out_src, u_src, A, B, out_dst, u_dst, W, H, step, mom, T, BW, g0, r0, r1, sync_every=1):
    L = 3 * W
    S = BW - 6 * T
    tid = np.arange(BW)
    gi = g0 + tid
    col_ok = (gi >= 0) & (gi < L)
    gic = np.clip(gi, 0, L - 1)
    f = np.float32
    win = np.zeros((T, 4, BW), np.float32)   # win[t][slot]: level t row produced at step == slot (mod 4)
    uu = np.zeros((T, 4, BW), np.float32)
    Ar = np.zeros((2 * T, BW), np.float32)   # arrival ring, 2T slots
    Br = np.zeros((2 * T, BW), np.float32)
    sm = np.zeros((T, 4, BW), np.float32)    # shared exchange ring as VISIBLE to other threads
    sm_w = np.zeros((T, 4, BW), np.float32)  # writes since the last barrier (worst case: invisible until then)
    pub_ok = col_ok & (gi < 3 * (W - 1))     # columns that publish (see the kernel: masks are free this way)
    nsteps = (r1 - r0) + 3 * T
    for s in range(nsteps):
        y_in = r0 - T + s
        for t in range(T, 0, -1):            # level T first: it reads the AB slot the arrival overwrites below
            rho = y_in - 2 * t
            c = win[t - 1][(s - 2) & 3]
            up = win[t - 1][(s - 3) & 3]
            dn = win[t - 1][(s - 1) & 3]
            row = sm[t - 1][(s - 2) & 3]
            lf = np.zeros(BW, np.float32)
            rt = np.zeros(BW, np.float32)
            lf[3:] = row[:-3]
            rt[:-3] = row[3:]
            dn = dn if (rho + 1) < (H - 1) else np.zeros(BW, np.float32)
            up = up if rho >= 1 else np.zeros(BW, np.float32)
            Ssum = ((rt + lf) + dn) + up
            a = Ar[(s - 2 * t) % (2 * T)]
            b = Br[(s - 2 * t) % (2 * T)]
            uo = uu[t - 1][(s - 2) & 3]
            un = (f(step) * Ssum + (a * c + b)).astype(np.float32)
            on = (f(mom) * uo + (c + un)).astype(np.float32)
            if t < T:
                win[t][s & 3] = on
                uu[t][s & 3] = un
                sm_w[t][s & 3] = np.where(pub_ok, on, f(0))
            else:
                if r0 <= rho < r1:
                    ok = (tid >= 3 * T) & (tid < 3 * T + S) & col_ok
                    out_dst[rho, gi[ok]] = on[ok]
                    u_dst[rho, gi[ok]] = un[ok]
        # level 0 arrives
        if 0 <= y_in < H:
            o0 = np.where(col_ok, out_src[y_in, gic], f(0))
            u0 = np.where(col_ok, u_src[y_in, gic], f(0))
            a0 = np.where(col_ok, A[y_in, gic], f(0))
            b0 = np.where(col_ok, B[y_in, gic], f(0))
        else:
            o0 = u0 = a0 = b0 = np.zeros(BW, np.float32)
        win[0][s & 3] = o0
        uu[0][s & 3] = u0
        sm_w[0][s & 3] = np.where(pub_ok, o0, f(0))
        Ar[s % (2 * T)] = a0
        Br[s % (2 * T)] = b0
        if (s % sync_every) == sync_every - 1:   # __syncthreads(): everything written so far becomes visible
            sm[:] = sm_w


def stream_pass(out, u, A, B, W, H, step, mom, T, BW, nchunks, sync_every=1):
    L = 3 * W
    S = BW - 6 * T
    assert S > 0
    out_dst = np.full_like(out, np.nan)
    u_dst = np.full_like(u, np.nan)
    nb = -(-L // S)
    CH = -(-H // nchunks)
    for b in range(nb):
        for c in range(nchunks):
            r0, r1 = c * CH, min(H, (c + 1) * CH)
            if r0 >= r1:
                continue
            cta(out, u, A, B, out_dst, u_dst, W, H, step, mom, T, BW, b * S - 3 * T, r0, r1, sync_every)
    return out_dst, u_dst
