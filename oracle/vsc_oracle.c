/* TEST INFRASTRUCTURE -- the oracle.  NOT product code.
 *
 * Plain-C CPU restatement of the reference's per-frame hot path
 * (MaxReimann/video-stream-consistency): the two onnxruntime custom ops and the
 * stabilization kernels.  Every function cites the reference file:line it follows
 * (paths relative to /root/reference).  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this library; the
 * product (video-stream-consistency_b200/) never does.
 *
 * Pinning (see oracle/README.md and DESIGN.md):
 *   - custom ops: bit-for-bit against the reference's own CPU kernels compiled
 *     unmodified into oracle/_ref/libvsc_ref_cpu.so, and against the committed
 *     fixtures tests/golden/ops_*.npz generated from that library;
 *   - stabilization: the reference has NO CPU implementation and no tests; the
 *     oracle is pinned against fixtures produced by the reference's CUDA kernels
 *     compiled unmodified for sm_100a and run on a B200
 *     (tests/golden/stab_*.npz, generator tests/golden/make_stab_golden.py).
 *
 * Build: gcc -O2 -ffp-contract=off -fopenmp (oracle/Makefile).  No FMA contraction, so
 * the arithmetic is exactly the C expression order written here.
 * All images: float, interleaved HWC, data[(y*W + x)*C + c]  (gpuimage.h:15-20).
 */
#include <math.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define IDX(x, y, c, W, C) ((((size_t)(y)) * (size_t)(W) + (size_t)(x)) * (size_t)(C) + (size_t)(c))

/* ------------------------------------------------------------------------------------------
 * C1  custom::Correlation, legacy=0 -- src/ort_custom_ops/src/opticalflow/correlation.cc:148-182
 * (correlate_patch) and :203-275 (correlation_forward), with the fixed parameters of
 * ComputeCPU :62-80 (kernel 1, stride 1, pad 0, dilation 1, patch = 2*md+1).
 *   out[n, ph, pw, h, w] = sum_c in1[n,c,h,w] * in2[n,c,h+ph-md,w+pw-md]   (terms outside the image = 0)
 * Accumulation is sequential over c in float, product rounded before the add (:176).
 * ------------------------------------------------------------------------------------------ */
void vsc_oracle_correlation(const float* in1, const float* in2, float* out, int N, int C, int H, int W, int md)
{
    const int P = 2 * md + 1;
    const size_t HW = (size_t)H * W;
#pragma omp parallel for collapse(2) schedule(static)
    for (int n = 0; n < N; ++n) {
        for (int ph = 0; ph < P; ++ph) {
            const float* a = in1 + (size_t)n * C * HW;
            const float* b = in2 + (size_t)n * C * HW;
            for (int pw = 0; pw < P; ++pw) {
                float* o = out + (((size_t)n * P + ph) * P + pw) * HW;
                for (int h = 0; h < H; ++h) {
                    const int h2 = h + ph - md;
                    for (int w = 0; w < W; ++w) {
                        const int w2 = w + pw - md;
                        float acc = 0.0f;
                        if (h2 >= 0 && h2 < H && w2 >= 0 && w2 < W) {
                            for (int c = 0; c < C; ++c) {
                                const float v1 = a[(size_t)c * HW + (size_t)h * W + w];
                                const float v2 = b[(size_t)c * HW + (size_t)h2 * W + w2];
                                acc += v1 * v2;
                            }
                        }
                        o[(size_t)h * W + w] = acc;
                    }
                }
            }
        }
    }
}

/* C3  custom::Correlation, legacy=1 (CUDA only in the reference) --
 * src/ort_custom_ops/src/opticalflow/correlation_cuda.cu:183-265 (correlation_old_kernel) and
 * :268-331 (CorrelateData_old): inputs zero-padded by md (the reference leaves the pad region
 * uninitialised, :51; the original PWC-Net code zero-fills it -- zero is the defined behaviour),
 *   out[n, (dy+md)*P + (dx+md), h, w] = (1/C) * sum_c in1[n,c,h,w] * in2[n,c,h+dy,w+dx]
 * top_channel % P is the x displacement, top_channel / P the y displacement (:233-234); the
 * division by sumelems = C is a float division (:261).  The reference reduces 32 lane-partials
 * (:238-258); the oracle sums sequentially over c -- same value up to fp32 reassociation. */
void vsc_oracle_correlation_legacy(const float* in1, const float* in2, float* out, int N, int C, int H, int W, int md)
{
    const int P = 2 * md + 1;
    const size_t HW = (size_t)H * W;
#pragma omp parallel for collapse(2) schedule(static)
    for (int n = 0; n < N; ++n) {
        for (int tc = 0; tc < P * P; ++tc) {
            const float* a = in1 + (size_t)n * C * HW;
            const float* b = in2 + (size_t)n * C * HW;
            const int dx = tc % P - md;
            const int dy = tc / P - md;
            float* o = out + ((size_t)n * P * P + tc) * HW;
            for (int h = 0; h < H; ++h) {
                const int h2 = h + dy;
                for (int w = 0; w < W; ++w) {
                    const int w2 = w + dx;
                    float acc = 0.0f;
                    if (h2 >= 0 && h2 < H && w2 >= 0 && w2 < W) {
                        for (int c = 0; c < C; ++c)
                            acc += a[(size_t)c * HW + (size_t)h * W + w] * b[(size_t)c * HW + (size_t)h2 * W + w2];
                    }
                    o[(size_t)h * W + w] = acc / (float)C;
                }
            }
        }
    }
}

/* C4  custom::Warp -- src/ort_custom_ops/src/opticalflow/warp.cc:71-134.
 * Masked bilinear backward warp, NCHW, flow [N,2,H,W] in pixels.  The reference mixes float
 * state with double literals: every product is evaluated in double and each `+=` rounds the
 * double sum back to float (:103-126).  Restated with the same types. */
void vsc_oracle_warp_nchw(const float* input, const float* flow, float* dst, int N, int C, int H, int W)
{
    const size_t HW = (size_t)H * W;
#pragma omp parallel for collapse(2) schedule(static)
    for (int n = 0; n < N; ++n) {
        for (int c = 0; c < C; ++c) {
            const float* in = input + ((size_t)n * C + c) * HW;
            const float* fx = flow + ((size_t)n * 2 + 0) * HW;
            const float* fy = flow + ((size_t)n * 2 + 1) * HW;
            float* o = dst + ((size_t)n * C + c) * HW;
            for (int y = 0; y < H; ++y) {
                for (int x = 0; x < W; ++x) {
                    const float xf = (float)x + fx[(size_t)y * W + x];
                    const float yf = (float)y + fy[(size_t)y * W + x];
                    const float alpha = xf - floorf(xf);
                    const float beta = yf - floorf(yf);
                    const float right_edge = (float)(W - 1);
                    const float bottom_edge = (float)(H - 1);
                    const float xL = floorf(xf);
                    const float xR = (float)((double)floorf(xf) + 1.0);
                    const float yT = floorf(yf);
                    const float yB = (float)((double)floorf(yf) + 1.0);
                    const int maskL = (0 <= xL && xL <= right_edge) ? 1 : 0;
                    const int maskR = (0 <= xR && xR <= right_edge) ? 1 : 0;
                    const int maskT = (0 <= yT && yT <= bottom_edge) ? 1 : 0;
                    const int maskB = (0 <= yB && yB <= bottom_edge) ? 1 : 0;
                    float val = 0.0f;
                    float mask = 0.0f;
                    mask = (float)(mask + (1.0 - alpha) * (1.0 - beta) * (maskT + maskL == 2 ? 1.0 : 0.0));
                    mask = (float)(mask + (double)alpha * (1.0 - beta) * (maskT + maskR == 2 ? 1.0 : 0.0));
                    mask = (float)(mask + (1.0 - alpha) * (double)beta * (maskB + maskL == 2 ? 1.0 : 0.0));
                    mask = (float)(mask + (double)(alpha * beta) * (maskB + maskR == 2 ? 1.0 : 0.0)); /* float*float first */
                    if (mask > 0.999) {
                        val = (float)(val + (1.0 - alpha) * (1.0 - beta)
                                               * (maskT + maskL == 2 ? in[(size_t)((int)yT) * W + (int)xL] : 0.0));
                        val = (float)(val + (double)alpha * (1.0 - beta)
                                               * (maskT + maskR == 2 ? in[(size_t)((int)yT) * W + (int)xR] : 0.0));
                        val = (float)(val + (1.0 - alpha) * (double)beta
                                               * (maskB + maskL == 2 ? in[(size_t)((int)yB) * W + (int)xL] : 0.0));
                        val = (float)(val + (double)(alpha * beta)
                                               * (maskB + maskR == 2 ? in[(size_t)((int)yB) * W + (int)xR] : 0.0));
                    }
                    o[(size_t)y * W + x] = val;
                }
            }
        }
    }
}

/* ------------------------------------------------------------------------------------------
 * S1  get_warp_result / kernel_warp -- src/stabilization/flowconsistency.cu:77-114.
 * Clamped bilinear backward warp of a float3 HWC image by an HWC flow (flowC = 3 or 2
 * channels, only 0 and 1 are read).  No validity mask; clamp to [0, W-3] x [0, H-3].
 * ------------------------------------------------------------------------------------------ */
void vsc_oracle_warp_hwc3(const float* in, const float* flow, float* out, int W, int H, int flowC)
{
#pragma omp parallel for schedule(static)
    for (int iy = 0; iy < H; ++iy) {
        for (int ix = 0; ix < W; ++ix) {
            const float flo_x = flow[IDX(ix, iy, 0, W, flowC)];
            const float flo_y = flow[IDX(ix, iy, 1, W, flowC)];
            const float map_fx = fmaxf(0.0f, fminf((float)ix + flo_x, (float)(W - 3)));
            const float map_fy = fmaxf(0.0f, fminf((float)iy + flo_y, (float)(H - 3)));
            const int map_ix = (int)floorf(map_fx);
            const int map_iy = (int)floorf(map_fy);
            const float flo_fx = map_fx - (float)map_ix;
            const float flo_fy = map_fy - (float)map_iy;
            for (int c = 0; c < 3; ++c) {
                const float tmp_1 = in[IDX(map_ix, map_iy, c, W, 3)] * (1.0f - flo_fx)
                    + in[IDX(map_ix + 1, map_iy, c, W, 3)] * flo_fx;
                const float tmp_2 = in[IDX(map_ix, map_iy + 1, c, W, 3)] * (1.0f - flo_fx)
                    + in[IDX(map_ix + 1, map_iy + 1, c, W, 3)] * flo_fx;
                out[IDX(ix, iy, c, W, 3)] = tmp_1 * (1.0f - flo_fy) + tmp_2 * flo_fy;
            }
        }
    }
}

/* S2  get_adap_comb / kernel_adap_comb -- src/stabilization/flowconsistency.cu:116-166.
 * n = W*H*3 values; purely element-wise so the images are treated as flat arrays. */
void vsc_oracle_adap_comb(const float* crntIn, const float* crntPr, const float* prevWarpIn, const float* prevWarpPr,
    const float* nextWarpIn, const float* nextWarpPr, float* adapCmbIn, float* adapCmbPr, const float* lastStabWarp,
    float alpha, size_t n)
{
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; ++i) {
        const float ci = crntIn[i], pi = prevWarpIn[i], ni = nextWarpIn[i];
        const float cp = crntPr[i], pp = prevWarpPr[i], np = nextWarpPr[i];
        const float ls = lastStabWarp[i];
        float wt_prv = expf(-alpha * (ci - pi) * (ci - pi));
        float wt_nxt = expf(-alpha * (ci - ni) * (ci - ni));
        if (wt_prv > 0.45f) wt_prv = 0.45f;
        if (wt_nxt > 0.3f) wt_nxt = 0.3f;
        if (wt_prv < 0.001f) wt_prv = 0.0f;
        if (wt_nxt < 0.001f) wt_nxt = 0.0f;
        const float adp_in = wt_prv * pi + wt_nxt * ni + (1.0f - (wt_prv + wt_nxt)) * ci;
        float adp_pr = wt_prv * pp + wt_nxt * np + (1.0f - (wt_prv + wt_nxt)) * cp;
        adp_pr = wt_prv * ls + (1.0f - wt_prv) * adp_pr;
        if (adapCmbIn) adapCmbIn[i] = adp_in;
        adapCmbPr[i] = adp_pr;
    }
}

/* S3  get_consist_wt / kernel_consist_wt -- src/stabilization/flowconsistency.cu:168-191. */
void vsc_oracle_consist_wt(const float* adapCmbIn, const float* crntIn, float* consWt, float beta, float gamma, size_t n)
{
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; ++i) {
        const float d = crntIn[i] - adapCmbIn[i];
        float wt = gamma * expf(-beta * d * d);
        if (wt < 0.001f) wt = 0.0f;
        consWt[i] = wt;
    }
}

/* S4  get_bilinear / kernel_bilinear -- src/stabilization/flowconsistency.cu:50-75.
 * Resize without half-pixel offset; loops c < output channels, reads with the input's stride. */
void vsc_oracle_bilinear(const float* in, int Wi, int Hi, int Ci, float* out, int Wo, int Ho, int Co)
{
#pragma omp parallel for schedule(static)
    for (int oy = 0; oy < Ho; ++oy) {
        for (int ox = 0; ox < Wo; ++ox) {
            const float xx = ((float)ox * (float)Wi) / (float)Wo;
            const float yy = ((float)oy * (float)Hi) / (float)Ho;
            const int ix = (int)floorf(xx);
            const int iy = (int)floorf(yy);
            const float fx = xx - (float)ix;
            const float fy = yy - (float)iy;
            const int ix1 = ix + 1 < Wi - 1 ? ix + 1 : Wi - 1;
            const int iy1 = iy + 1 < Hi - 1 ? iy + 1 : Hi - 1;
            for (int c = 0; c < Co; ++c) {
                const float v00 = in[IDX(ix, iy, c, Wi, Ci)];
                const float v10 = in[IDX(ix1, iy, c, Wi, Ci)];
                const float v01 = in[IDX(ix, iy1, c, Wi, Ci)];
                const float v11 = in[IDX(ix1, iy1, c, Wi, Ci)];
                const float v = v00 * (1.0f - fx) * (1.0f - fy) + v10 * fx * (1.0f - fy) + v01 * (1.0f - fx) * fy
                    + v11 * fx * fy;
                out[IDX(ox, oy, c, Wo, Co)] = v;
            }
        }
    }
}

/* S5  get_consist_out / kernel_consist_out -- src/stabilization/flowconsistency.cu:193-258, :350-374.
 * numIter gradient-descent sweeps of the screened-Poisson energy with momentum on the previous
 * step (:247-255).  The reference passes consisOut as both the buffer read (prevConsis) and the
 * buffer written (:370), so its result depends on GPU scheduling.  mode selects the ordering:
 *   mode 0  Jacobi: every sweep reads the previous sweep's image only (double buffered).
 *           This is the deterministic definition the product implements.
 *   mode 1  in-place, raster order, channel-inner: what a sequential execution of :199-257
 *           produces (Gauss-Seidel); brackets the reference's possible outputs.
 * Neighbour inclusion tests are the reference's asymmetric ones (:215-236).
 * scratch-free API: allocates its own temporaries. */
static inline float solve_value(const float* prev, const float* pr, const float* tgt, const float* wt, int ix, int iy,
    int c, int W, int H)
{
    int cnt = 0;
    float lap_pr = 0.0f, lap_out = 0.0f;
    if ((ix + 1) < (W - 1)) {
        lap_pr += pr[IDX(ix + 1, iy, c, W, 3)];
        lap_out += prev[IDX(ix + 1, iy, c, W, 3)];
        cnt += 1;
    }
    if ((ix - 1) >= 0) {
        lap_pr += pr[IDX(ix - 1, iy, c, W, 3)];
        lap_out += prev[IDX(ix - 1, iy, c, W, 3)];
        cnt += 1;
    }
    if ((iy + 1) < (H - 1)) {
        lap_pr += pr[IDX(ix, iy + 1, c, W, 3)];
        lap_out += prev[IDX(ix, iy + 1, c, W, 3)];
        cnt += 1;
    }
    if ((iy - 1) >= 0) {
        lap_pr += pr[IDX(ix, iy - 1, c, W, 3)];
        lap_out += prev[IDX(ix, iy - 1, c, W, 3)];
        cnt += 1;
    }
    lap_pr -= (float)cnt * pr[IDX(ix, iy, c, W, 3)];
    lap_out -= (float)cnt * prev[IDX(ix, iy, c, W, 3)];
    const float wt_val = wt[IDX(ix, iy, c, W, 3)];
    const float tmp_1 = wt_val * (prev[IDX(ix, iy, c, W, 3)] - tgt[IDX(ix, iy, c, W, 3)]);
    const float tmp_2 = lap_out - lap_pr;
    return tmp_2 - tmp_1; /* grad_val */
}

int vsc_oracle_consist_out(const float* crntPr, const float* prevStabWarp, const float* consWt, int numIter,
    float stepSize, float momFac, float* consisOut, int W, int H, int mode)
{
    const size_t n = (size_t)W * H * 3;
    float* upd = (float*)calloc(n, sizeof(float));
    float* alt = mode == 0 ? (float*)malloc(n * sizeof(float)) : NULL;
    if (!upd || (mode == 0 && !alt)) {
        free(upd);
        free(alt);
        return 1;
    }
    float* cur = consisOut;
    float* nxt = alt;
    for (int k = 0; k < numIter; ++k) {
        if (mode == 0) {
#pragma omp parallel for schedule(static)
            for (int iy = 0; iy < H; ++iy)
                for (int ix = 0; ix < W; ++ix)
                    for (int c = 0; c < 3; ++c) {
                        const float g = solve_value(cur, crntPr, prevStabWarp, consWt, ix, iy, c, W, H);
                        const size_t i = IDX(ix, iy, c, W, 3);
                        float o;
                        if (k)
                            o = cur[i] + stepSize * g + momFac * upd[i];
                        else
                            o = cur[i] + stepSize * g;
                        upd[i] = stepSize * g;
                        nxt[i] = o;
                    }
            float* t = cur;
            cur = nxt;
            nxt = t;
        } else {
            for (int iy = 0; iy < H; ++iy)
                for (int ix = 0; ix < W; ++ix)
                    for (int c = 0; c < 3; ++c) {
                        const float g = solve_value(cur, crntPr, prevStabWarp, consWt, ix, iy, c, W, H);
                        const size_t i = IDX(ix, iy, c, W, 3);
                        float o;
                        if (k)
                            o = cur[i] + stepSize * g + momFac * upd[i];
                        else
                            o = cur[i] + stepSize * g;
                        upd[i] = stepSize * g;
                        cur[i] = o;
                    }
        }
    }
    if (mode == 0 && cur != consisOut)
        memcpy(consisOut, cur, n * sizeof(float));
    free(upd);
    free(alt);
    return 0;
}

/* S7  kernel_to_float_image -- src/stabilization/gpuimage.cu:39-51: float(double(u8) / 255.0), alpha ignored. */
void vsc_oracle_rgba8_to_f32x3(const uint8_t* rgba, float* out, int W, int H)
{
    const size_t P = (size_t)W * H;
#pragma omp parallel for schedule(static)
    for (size_t p = 0; p < P; ++p)
        for (int c = 0; c < 3; ++c)
            out[p * 3 + c] = (float)((double)(float)rgba[p * 4 + c] / 255.0);
}

/* S8  kernel_to_char_image -- src/stabilization/gpuimage.cu:54-67:
 * (unsigned char)__float2uint_rd(fabs(v) * 255): float multiply, round toward -inf to uint32
 * (saturating, NaN -> 0), then truncation to 8 bits (wraps mod 256, no clamp); alpha byte = 1. */
void vsc_oracle_f32x3_to_rgba8(const float* in, uint8_t* rgba, int W, int H)
{
    const size_t P = (size_t)W * H;
#pragma omp parallel for schedule(static)
    for (size_t p = 0; p < P; ++p) {
        for (int c = 0; c < 3; ++c) {
            const float s = fabsf(in[p * 3 + c]) * 255.0f;
            uint32_t u;
            if (!(s == s))
                u = 0u;
            else if (s >= 4294967296.0f)
                u = 0xFFFFFFFFu;
            else
                u = (uint32_t)floorf(s);
            rgba[p * 4 + c] = (uint8_t)(u & 0xFFu);
        }
        rgba[p * 4 + 3] = 1;
    }
}

/* One frame of VideoStabilizer::doOneStep -- src/stabilization/videostabilizer.cpp:177-247
 * (SURVEY Appendix B).  Pyramid of `levels` levels (2 in the reference, :109) with integer-halved
 * sizes (:126-127); numIter/(j+1) sweeps at level j (:226).  lastStab is updated in place with
 * the fp32 result (:247).  consisOut/rgba may be NULL.  Returns 0, or 1 on allocation failure. */
int vsc_oracle_do_one_step(const float* origPrev, const float* origCur, const float* origNext, const float* procPrev,
    const float* procCur, const float* procNext, float* lastStab, const float* flowFwd, const float* flowBwd, int W,
    int H, int flowC, float alpha, float beta, float gamma, int levels, int numIter, float stepSize, float momFac,
    int mode, float* consisOut, uint8_t* rgba)
{
    const size_t n = (size_t)W * H * 3;
    if (levels < 1 || levels > 8)
        return 1;
    float* buf = (float*)malloc(9 * n * sizeof(float));
    if (!buf)
        return 1;
    float *prevWarpIn = buf, *prevWarpPr = buf + n, *nextWarpIn = buf + 2 * n, *nextWarpPr = buf + 3 * n,
          *lastStabWarp = buf + 4 * n, *adapCmbIn = buf + 5 * n, *adapCmbPr = buf + 6 * n, *consWt = buf + 7 * n,
          *out0 = buf + 8 * n;

    vsc_oracle_warp_hwc3(origPrev, flowBwd, prevWarpIn, W, H, flowC);   /* :182 */
    vsc_oracle_warp_hwc3(procPrev, flowBwd, prevWarpPr, W, H, flowC);   /* :183 */
    vsc_oracle_warp_hwc3(origNext, flowFwd, nextWarpIn, W, H, flowC);   /* :186 */
    vsc_oracle_warp_hwc3(procNext, flowFwd, nextWarpPr, W, H, flowC);   /* :187 */
    vsc_oracle_warp_hwc3(lastStab, flowBwd, lastStabWarp, W, H, flowC); /* :190 */
    vsc_oracle_adap_comb(origCur, procCur, prevWarpIn, prevWarpPr, nextWarpIn, nextWarpPr, adapCmbIn, adapCmbPr,
        lastStabWarp, alpha, n);                                        /* :194 */
    vsc_oracle_consist_wt(adapCmbIn, origCur, consWt, beta, gamma, n);  /* :198 */

    /* pyramid (:203-217): level 0 aliases the full-res images, out0 starts as the processed frame */
    const float* pyrPr[8];
    const float* pyrTg[8];
    const float* pyrWt[8];
    float* pyrOut[8];
    float* own[8] = {0};
    int pw[8], ph[8];
    memcpy(out0, procCur, n * sizeof(float));
    pyrPr[0] = procCur;
    pyrTg[0] = adapCmbPr;
    pyrWt[0] = consWt;
    pyrOut[0] = out0;
    pw[0] = W;
    ph[0] = H;
    int rc = 0;
    for (int j = 1; j < levels && !rc; ++j) {
        pw[j] = pw[j - 1] / 2;
        ph[j] = ph[j - 1] / 2;
        const size_t m = (size_t)pw[j] * ph[j] * 3;
        own[j] = (float*)malloc(4 * (m ? m : 1) * sizeof(float));
        if (!own[j]) {
            rc = 1;
            break;
        }
        float *a = own[j], *b = own[j] + m, *c = own[j] + 2 * m, *d = own[j] + 3 * m;
        vsc_oracle_bilinear(pyrPr[j - 1], pw[j - 1], ph[j - 1], 3, a, pw[j], ph[j], 3);
        vsc_oracle_bilinear(pyrTg[j - 1], pw[j - 1], ph[j - 1], 3, b, pw[j], ph[j], 3);
        vsc_oracle_bilinear(pyrWt[j - 1], pw[j - 1], ph[j - 1], 3, c, pw[j], ph[j], 3);
        vsc_oracle_bilinear(pyrOut[j - 1], pw[j - 1], ph[j - 1], 3, d, pw[j], ph[j], 3);
        pyrPr[j] = a;
        pyrTg[j] = b;
        pyrWt[j] = c;
        pyrOut[j] = d;
    }
    for (int j = levels - 1; j >= 0 && !rc; --j) { /* :219-227 */
        if (j != levels - 1)
            vsc_oracle_bilinear(pyrOut[j + 1], pw[j + 1], ph[j + 1], 3, pyrOut[j], pw[j], ph[j], 3);
        rc = vsc_oracle_consist_out(pyrPr[j], pyrTg[j], pyrWt[j], numIter / (j + 1), stepSize, momFac, pyrOut[j],
            pw[j], ph[j], mode);
    }
    if (!rc) {
        if (rgba)
            vsc_oracle_f32x3_to_rgba8(out0, rgba, W, H); /* :237-238 */
        memcpy(lastStab, out0, n * sizeof(float));       /* :247 */
        if (consisOut)
            memcpy(consisOut, out0, n * sizeof(float));
    }
    for (int j = 1; j < levels; ++j)
        free(own[j]);
    free(buf);
    return rc;
}

/* torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU baseline sets the team size explicitly */
void vsc_oracle_set_num_threads(int n)
{
#ifdef _OPENMP
    if (n > 0)
        omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int vsc_oracle_num_threads(void)
{
    int n = 1;
#ifdef _OPENMP
#pragma omp parallel
    {
#pragma omp master
        n = omp_get_num_threads();
    }
#endif
    return n;
}
