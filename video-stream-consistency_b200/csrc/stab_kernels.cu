// Stand-alone stabilization kernels for sm_100a: the one-to-one replacements of
// get_warp_result, get_adap_comb, get_consist_wt, get_bilinear and of the RGBA8<->float3
// conversions (reference: src/stabilization/flowconsistency.cu:50-191,288-348 and
// gpuimage.cu:39-67).  These keep the reference's granularity so the flowconsistency.cuh shim
// can forward call by call; the per-frame pipeline uses the fused kernels instead
// (stab_fused.cu, stab_solver.cu).
//
// All are HBM-bound streaming kernels: one thread per pixel (or per 4 consecutive floats for
// the element-wise ones, 128-bit accesses), grids sized by the data, no shared memory needed.
// Unlike the reference they take a stream, do not synchronise the device and do not over-launch
// an extra block row/column (getGrid, flowconsistency.cu:262-266).
#include "stab_device.cuh"

namespace vsc {

// one thread per value (pixel, channel): consecutive threads touch consecutive floats of the interleaved image
__global__ void __launch_bounds__(256) warp_hwc3_kernel(const float* __restrict__ in, const float* __restrict__ flow,
    float* __restrict__ out, int W, int H, int flowC)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int iy = blockIdx.y;
    if (i >= 3 * W)
        return;
    const int ix = i / 3;
    const int c = i - 3 * ix;
    const size_t p = static_cast<size_t>(iy) * W + ix;
    const WarpGeom g = hwc_warp_geom(ix, iy, __ldg(flow + p * flowC), __ldg(flow + p * flowC + 1), W, H);
    out[p * 3 + c] = hwc_warp_sample1(in, W, g, c);
}

template <bool VEC>
__global__ void __launch_bounds__(256) adap_comb_kernel(const float* __restrict__ crntIn,
    const float* __restrict__ crntPr, const float* __restrict__ prevWarpIn, const float* __restrict__ prevWarpPr,
    const float* __restrict__ nextWarpIn, const float* __restrict__ nextWarpPr, float* __restrict__ adapCmbIn,
    float* __restrict__ adapCmbPr, const float* __restrict__ lastStabWarp, float alpha, size_t n)
{
    const size_t i = (static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x) * (VEC ? 4 : 1);
    if (i >= n)
        return;
    if constexpr (VEC) {
        const float4 ci = ldg_stream4(crntIn + i), cp = ldg_stream4(crntPr + i), pi = ldg_stream4(prevWarpIn + i),
                     pp = ldg_stream4(prevWarpPr + i), ni = ldg_stream4(nextWarpIn + i),
                     np = ldg_stream4(nextWarpPr + i), ls = ldg_stream4(lastStabWarp + i);
        float4 ai, ap;
        adap_comb_value(ci.x, cp.x, pi.x, pp.x, ni.x, np.x, ls.x, alpha, ai.x, ap.x);
        adap_comb_value(ci.y, cp.y, pi.y, pp.y, ni.y, np.y, ls.y, alpha, ai.y, ap.y);
        adap_comb_value(ci.z, cp.z, pi.z, pp.z, ni.z, np.z, ls.z, alpha, ai.z, ap.z);
        adap_comb_value(ci.w, cp.w, pi.w, pp.w, ni.w, np.w, ls.w, alpha, ai.w, ap.w);
        if (adapCmbIn)
            *reinterpret_cast<float4*>(adapCmbIn + i) = ai;
        *reinterpret_cast<float4*>(adapCmbPr + i) = ap;
    } else {
        float ai, ap;
        adap_comb_value(crntIn[i], crntPr[i], prevWarpIn[i], prevWarpPr[i], nextWarpIn[i], nextWarpPr[i],
            lastStabWarp[i], alpha, ai, ap);
        if (adapCmbIn)
            adapCmbIn[i] = ai;
        adapCmbPr[i] = ap;
    }
}

template <bool VEC>
__global__ void __launch_bounds__(256) consist_wt_kernel(const float* __restrict__ adapCmbIn,
    const float* __restrict__ crntIn, float* __restrict__ consWt, float beta, float gamma, size_t n)
{
    const size_t i = (static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x) * (VEC ? 4 : 1);
    if (i >= n)
        return;
    if constexpr (VEC) {
        const float4 a = ldg_stream4(adapCmbIn + i), c = ldg_stream4(crntIn + i);
        float4 w;
        w.x = consist_wt_value(c.x, a.x, beta, gamma);
        w.y = consist_wt_value(c.y, a.y, beta, gamma);
        w.z = consist_wt_value(c.z, a.z, beta, gamma);
        w.w = consist_wt_value(c.w, a.w, beta, gamma);
        *reinterpret_cast<float4*>(consWt + i) = w;
    } else {
        consWt[i] = consist_wt_value(crntIn[i], adapCmbIn[i], beta, gamma);
    }
}

// kernel_bilinear, flowconsistency.cu:50-75.  One thread per output PIXEL: the coordinate arithmetic (two IEEE
// divisions) is shared by the channels; a value-per-thread variant was measured 1.6-1.9x slower (r1_notes.md).
__global__ void __launch_bounds__(256) bilinear_kernel(const float* __restrict__ in, int Wi, int Hi, int Ci,
    float* __restrict__ out, int Wo, int Ho, int Co)
{
    pdl_enter();
    const int ox = blockIdx.x * blockDim.x + threadIdx.x;
    const int oy = blockIdx.y;
    if (ox >= Wo)
        return;
    const float xx = (static_cast<float>(ox) * static_cast<float>(Wi)) / static_cast<float>(Wo);
    const float yy = (static_cast<float>(oy) * static_cast<float>(Hi)) / static_cast<float>(Ho);
    const int ix = static_cast<int>(floorf(xx));
    const int iy = static_cast<int>(floorf(yy));
    const float fx = xx - static_cast<float>(ix);
    const float fy = yy - static_cast<float>(iy);
    const int ix1 = min(ix + 1, Wi - 1);
    const int iy1 = min(iy + 1, Hi - 1);
    const float* r0 = in + static_cast<size_t>(iy) * Wi * Ci;
    const float* r1 = in + static_cast<size_t>(iy1) * Wi * Ci;
    float* o = out + (static_cast<size_t>(oy) * Wo + ox) * Co;
    const float ofx = 1.0f - fx, ofy = 1.0f - fy;
    for (int c = 0; c < Co; ++c) {
        const float v00 = __ldg(r0 + static_cast<size_t>(ix) * Ci + c);
        const float v10 = __ldg(r0 + static_cast<size_t>(ix1) * Ci + c);
        const float v01 = __ldg(r1 + static_cast<size_t>(ix) * Ci + c);
        const float v11 = __ldg(r1 + static_cast<size_t>(ix1) * Ci + c);
        o[c] = v00 * ofx * ofy + v10 * fx * ofy + v01 * ofx * fy + v11 * fx * fy;
    }
}

// The exact x2 up-scale (flow 960x540 -> 1080p, pyramid level 1 -> 0: the only up-scales of the pipeline), one
// thread per INPUT value, which owns the 2x2 output values that sample around it.  With Wo = 2 Wi and
// ox * Wi < 2^24 the reference's coordinate (float(ox) * float(Wi)) / float(Wo) is exactly ox / 2 (exact product,
// exactly representable quotient), so ix = ox >> 1 and fx is 0 or 0.5 without a division; the four taps are loaded
// once for four outputs, and the interpolation expression is the reference's, evaluated in its order for each
// output: bit-identical to bilinear_kernel at a quarter of its loads and none of its divisions.
template <int C>
__global__ void __launch_bounds__(256) bilinear_up2_kernel(const float* __restrict__ in, int Wi, int Hi,
    float* __restrict__ out)
{
    pdl_enter();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;   // float index inside the input row
    const int iy = blockIdx.y;
    if (i >= Wi * C)
        return;
    const int ix = i / C;
    const int c = i - ix * C;
    const int dx = (ix + 1 < Wi) ? C : 0;                  // min(ix + 1, Wi - 1)
    const float* r0 = in + static_cast<size_t>(iy) * Wi * C + i;
    const float* r1 = (iy + 1 < Hi) ? r0 + static_cast<size_t>(Wi) * C : r0;
    const float v00 = __ldg(r0), v10 = __ldg(r0 + dx), v01 = __ldg(r1), v11 = __ldg(r1 + dx);
    auto interp = [&](float fx, float fy) {
        const float ofx = 1.0f - fx, ofy = 1.0f - fy;
        return v00 * ofx * ofy + v10 * fx * ofy + v01 * ofx * fy + v11 * fx * fy;
    };
    const size_t Lo = static_cast<size_t>(Wi) * 2 * C;
    float* o = out + static_cast<size_t>(2 * iy) * Lo + static_cast<size_t>(2 * ix) * C + c;
    __stcs(o, interp(0.0f, 0.0f));
    __stcs(o + C, interp(0.5f, 0.0f));
    __stcs(o + Lo, interp(0.0f, 0.5f));
    __stcs(o + Lo + C, interp(0.5f, 0.5f));
}

// kernel_to_float_image, gpuimage.cu:39-51.  float(double(u8)/255.0) == float(u8)/255.0f for all 256 inputs
// (checked exhaustively in tests/test_oracle.py), so the IEEE float division is used.
// A CTA converts 256 pixels: one coalesced uchar4 load per thread, the 768 floats are transposed through shared
// memory (stride-3 writes: conflict-free) and leave as three fully coalesced 128-byte-per-warp stores.  (A thread
// per pixel writes three floats at a 12-byte stride -- 13.8 us at 1080p against 5 us of compulsory traffic; a
// thread per value with byte loads was worse still, 19.7 us.)
__global__ void __launch_bounds__(256) rgba8_to_f32x3_kernel(const uchar4* __restrict__ in, float* __restrict__ out,
    size_t P)
{
    pdl_enter();
    __shared__ float tile[768];
    const size_t p0 = static_cast<size_t>(blockIdx.x) * 256;
    const size_t p = p0 + threadIdx.x;
    if (p < P) {
        const uchar4 v = __ldg(in + p);
        tile[threadIdx.x * 3 + 0] = static_cast<float>(v.x) / 255.0f;
        tile[threadIdx.x * 3 + 1] = static_cast<float>(v.y) / 255.0f;
        tile[threadIdx.x * 3 + 2] = static_cast<float>(v.z) / 255.0f;
    }
    __syncthreads();
    const size_t n = (P - p0 < 256 ? P - p0 : 256) * 3;   // floats of this CTA
    float* o = out + p0 * 3;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const unsigned i = threadIdx.x + k * 256;
        if (i < n)
            __stcs(o + i, tile[i]);
    }
}

// kernel_to_char_image, gpuimage.cu:54-67: floor(|v|*255) to uint32 (saturating, NaN->0), low 8 bits, alpha = 1
__device__ __forceinline__ unsigned char f32_to_u8(float v)
{
    return static_cast<unsigned char>(__float2uint_rd(fabsf(v) * 255.0f));
}

// (staging the 768 floats of a CTA through shared memory like rgba8_to_f32x3_kernel was measured slower here:
// 9.3 vs 8.6 us at 1080p -- the strided float loads are absorbed by L1)
__global__ void __launch_bounds__(256) f32x3_to_rgba8_kernel(const float* __restrict__ in, uchar4* __restrict__ out,
    size_t P)
{
    pdl_enter();
    const size_t p = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (p >= P)
        return;
    uchar4 o;
    o.x = f32_to_u8(in[p * 3 + 0]);
    o.y = f32_to_u8(in[p * 3 + 1]);
    o.z = f32_to_u8(in[p * 3 + 2]);
    o.w = 1;
    out[p] = o;
}

}  // namespace vsc

using namespace vsc;

extern "C" int vsc_warp_hwc3(const float* in, const float* flow, float* out, int W, int H, int flow_channels,
    vsc_stream_t stream)
{
    if (!in || !flow || !out || W < 2 || H < 2 || (flow_channels != 2 && flow_channels != 3) || H > 65535)
        return VSC_E_INVALID;
    const dim3 grid(cdiv(3LL * W, 256), H);
    warp_hwc3_kernel<<<grid, 256, 0, as_stream(stream)>>>(in, flow, out, W, H, flow_channels);
    count_launch();
    return launch_status();
}

extern "C" int vsc_adap_comb(const float* crntIn, const float* crntPr, const float* prevWarpIn,
    const float* prevWarpPr, const float* nextWarpIn, const float* nextWarpPr, float* adapCmbIn, float* adapCmbPr,
    const float* lastStabWarp, float alpha, int W, int H, vsc_stream_t stream)
{
    if (!crntIn || !crntPr || !prevWarpIn || !prevWarpPr || !nextWarpIn || !nextWarpPr || !adapCmbPr || !lastStabWarp
        || W <= 0 || H <= 0)
        return VSC_E_INVALID;
    const size_t n = static_cast<size_t>(W) * H * 3;
    const bool vec = n % 4 == 0 && aligned16(crntIn) && aligned16(crntPr) && aligned16(prevWarpIn)
        && aligned16(prevWarpPr) && aligned16(nextWarpIn) && aligned16(nextWarpPr) && aligned16(adapCmbPr)
        && aligned16(lastStabWarp) && (!adapCmbIn || aligned16(adapCmbIn));
    if (vec)
        adap_comb_kernel<true><<<cdiv(n / 4, 256), 256, 0, as_stream(stream)>>>(crntIn, crntPr, prevWarpIn,
            prevWarpPr, nextWarpIn, nextWarpPr, adapCmbIn, adapCmbPr, lastStabWarp, alpha, n);
    else
        adap_comb_kernel<false><<<cdiv(n, 256), 256, 0, as_stream(stream)>>>(crntIn, crntPr, prevWarpIn, prevWarpPr,
            nextWarpIn, nextWarpPr, adapCmbIn, adapCmbPr, lastStabWarp, alpha, n);
    count_launch();
    return launch_status();
}

extern "C" int vsc_consist_wt(const float* adapCmbIn, const float* crntIn, float* consWt, float beta, float gamma,
    int W, int H, vsc_stream_t stream)
{
    if (!adapCmbIn || !crntIn || !consWt || W <= 0 || H <= 0)
        return VSC_E_INVALID;
    const size_t n = static_cast<size_t>(W) * H * 3;
    const bool vec = n % 4 == 0 && aligned16(adapCmbIn) && aligned16(crntIn) && aligned16(consWt);
    if (vec)
        consist_wt_kernel<true><<<cdiv(n / 4, 256), 256, 0, as_stream(stream)>>>(adapCmbIn, crntIn, consWt, beta,
            gamma, n);
    else
        consist_wt_kernel<false><<<cdiv(n, 256), 256, 0, as_stream(stream)>>>(adapCmbIn, crntIn, consWt, beta, gamma,
            n);
    count_launch();
    return launch_status();
}

extern "C" int vsc_bilinear(const float* in, int Wi, int Hi, int Ci, float* out, int Wo, int Ho, int Co,
    vsc_stream_t stream)
{
    if (!in || !out || Wi <= 0 || Hi <= 0 || Ci <= 0 || Wo <= 0 || Ho <= 0 || Co <= 0 || Co > Ci || Ho > 65535)
        return VSC_E_INVALID;
    if (Wo == 2 * Wi && Ho == 2 * Hi && Ci == Co && (Co == 3 || Co == 2) && static_cast<long long>(Wo) * Wi <= (1 << 24)
        && static_cast<long long>(Ho) * Hi <= (1 << 24) && static_cast<long long>(Wi) * Ci < (1 << 29)) {
        const dim3 grid2(cdiv(static_cast<long long>(Wi) * Co, 256), Hi);
        const int rc = Co == 3 ? launch_pdl(bilinear_up2_kernel<3>, grid2, dim3(256), 0, as_stream(stream), in, Wi, Hi, out)
                               : launch_pdl(bilinear_up2_kernel<2>, grid2, dim3(256), 0, as_stream(stream), in, Wi, Hi, out);
        count_launch();
        return rc ? rc : launch_status();
    }
    const dim3 grid(cdiv(Wo, 256), Ho);
    const int rc = launch_pdl(bilinear_kernel, grid, dim3(256), 0, as_stream(stream), in, Wi, Hi, Ci, out, Wo, Ho, Co);
    count_launch();
    return rc ? rc : launch_status();
}

// Device-side input path for the flow network.  FlowModel::run scales the three window frames on
// the CPU with QImage::scaled(w, h, Qt::IgnoreAspectRatio, Qt::FastTransformation) and copies them through two
// host buffers and a pageable H2D per direction (flowmodel.cpp:121-150, imagehelpers.cpp:22-40).  Here the RGBA8
// frame already uploaded for the stabilization is scaled on the device: nearest neighbour, sampled at pixel
// centres in 16.16 fixed point (ix = 65536 * sw / dw truncated, source x = (ix / 2 + x * ix) >> 16) -- the scheme
// of Qt 5's raster scaler as far as it is documented; Qt is not in this image, so byte-equality with
// QImage::scaled is NOT claimed or tested (the definition above is, tests/test_stab_gpu.py).
namespace vsc {
__global__ void __launch_bounds__(256) rgba8_scale_nearest_kernel(const uchar4* __restrict__ src, int sw, int sh,
    uchar4* __restrict__ dst, int dw, int dh, unsigned ix, unsigned iy)
{
    pdl_enter();
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    if (x >= dw)
        return;
    const unsigned sx = min((ix / 2 + static_cast<unsigned>(x) * ix) >> 16, static_cast<unsigned>(sw - 1));
    const unsigned sy = min((iy / 2 + static_cast<unsigned>(y) * iy) >> 16, static_cast<unsigned>(sh - 1));
    dst[static_cast<size_t>(y) * dw + x] = __ldg(src + static_cast<size_t>(sy) * sw + sx);
}
}  // namespace vsc

extern "C" int vsc_rgba8_scale_nearest(const uint8_t* src_dev, int srcW, int srcH, uint8_t* dst_dev, int dstW, int dstH,
    vsc_stream_t stream)
{
    using namespace vsc;
    if (!src_dev || !dst_dev || srcW <= 0 || srcH <= 0 || dstW <= 0 || dstH <= 0 || dstH > 65535 || srcW > 32767
        || srcH > 32767 || dstW > 32767)
        return VSC_E_INVALID;
    if (!aligned4(src_dev) || !aligned4(dst_dev))
        return VSC_E_ALIGN;
    const unsigned ix = static_cast<unsigned>(65536.0 * static_cast<double>(srcW) / static_cast<double>(dstW));
    const unsigned iy = static_cast<unsigned>(65536.0 * static_cast<double>(srcH) / static_cast<double>(dstH));
    // (ix / 2 + x * ix) must fit 32 bits: x * ix < dstW * 65536 * srcW / dstW = 65536 * srcW <= 2^31
    const dim3 grid(cdiv(dstW, 256), dstH);
    const int rc = launch_pdl(rgba8_scale_nearest_kernel, grid, dim3(256), 0, as_stream(stream),
        reinterpret_cast<const uchar4*>(src_dev), srcW, srcH, reinterpret_cast<uchar4*>(dst_dev), dstW, dstH, ix, iy);
    count_launch();
    return rc ? rc : launch_status();
}

extern "C" int vsc_rgba8_to_f32x3(const uint8_t* rgba_dev, float* out, int W, int H, vsc_stream_t stream)
{
    if (!rgba_dev || !out || W <= 0 || H <= 0)
        return VSC_E_INVALID;
    if (!aligned4(rgba_dev))
        return VSC_E_ALIGN;
    const size_t P = static_cast<size_t>(W) * H;
    const int rc = launch_pdl(rgba8_to_f32x3_kernel, dim3(cdiv(P, 256)), dim3(256), 0, as_stream(stream),
        reinterpret_cast<const uchar4*>(rgba_dev), out, P);
    count_launch();
    return rc ? rc : launch_status();
}

extern "C" int vsc_f32x3_to_rgba8(const float* in, uint8_t* rgba_dev, int W, int H, vsc_stream_t stream)
{
    if (!in || !rgba_dev || W <= 0 || H <= 0)
        return VSC_E_INVALID;
    if (!aligned4(rgba_dev))
        return VSC_E_ALIGN;
    const size_t P = static_cast<size_t>(W) * H;
    const int rc = launch_pdl(f32x3_to_rgba8_kernel, dim3(cdiv(P, 256)), dim3(256), 0, as_stream(stream), in,
        reinterpret_cast<uchar4*>(rgba_dev), P);
    count_launch();
    return rc ? rc : launch_status();
}
