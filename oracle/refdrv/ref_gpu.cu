// TEST INFRASTRUCTURE -- not product code.
//
// Driver that runs the reference's OWN CUDA code on a B200: the custom ops
// (correlation_cuda.{cc,cu}, warp_cuda.{cc,cu}) through
// CorrelationKernel::Compute / WarpKernel::Compute with the CUDA provider, and the
// stabilization kernels (flowconsistency.cu, gpuimage.{cu,cpp}) through the six
// free functions of flowconsistency.cuh.  All reference files are compiled
// UNMODIFIED from /root/reference for sm_100a (oracle/Makefile) against the
// stand-in headers of standins/.  The per-frame call sequence of
// VideoStabilizer::doOneStep (videostabilizer.cpp:167-265) cannot be compiled
// (Qt containers) and is restated in vsc_ref_gpu_do_one_step below, calling the
// reference's functions in the reference's order.
//
// Used only by tests/ (golden generation on the GPU box, parity), and by
// bench.py's "reference kernels on the same box" timing.  Never by the product.
#include <ort_custom_ops/opticalflow/correlation.h>
#include <ort_custom_ops/opticalflow/warp.h>

#include "flowconsistency.cuh"
#include "gpuimage.h"

#include <cuda_runtime.h>

#include <cstdio>
#include <cstring>
#include <exception>
#include <memory>
#include <vector>

namespace {

struct OutSlot {
    void* ptr;
    size_t bytes;
};

void* take_output(void* user, size_t, size_t bytes)
{
    OutSlot* s = static_cast<OutSlot*>(user);
    if (bytes > s->bytes)
        throw std::runtime_error("ref driver: output buffer too small");
    return s->ptr;
}

const OrtApi g_api{};

// A GPUImage that views caller-owned device memory: the reference type always
// cudaMallocs in its ctor (gpuimage.cpp:30-39), so allocate 1 element and swap the
// public fields for the lifetime of the view.
struct View {
    GPUImage img;
    float* own;
    View(const float* p, int w, int h, int c) : img(1, 1, 1), own(img.data)
    {
        img.data = const_cast<float*>(p);
        img.width = w;
        img.height = h;
        img.channels = c;
    }
    ~View()
    {
        img.data = own;
        img.width = img.height = img.channels = 1;
    }
    operator GPUImage&() { return img; }
};

}  // namespace

extern "C" {

int vsc_ref_gpu_correlation(const float* in1, const float* in2, float* out, size_t out_bytes, int64_t N, int64_t C,
    int64_t H, int64_t W, int64_t max_displacement, int64_t legacy, void* stream)
{
    try {
        OrtKernelInfo info;
        info.legacy = legacy;
        info.max_displacement = max_displacement;
        CorrelationKernel k(g_api, &info, "CUDAExecutionProvider");
        OrtKernelContext ctx;
        OutSlot slot{out, out_bytes};
        ctx.alloc_output = take_output;
        ctx.alloc_user = &slot;
        ctx.gpu_stream = stream;
        ctx.inputs.resize(2);
        ctx.inputs[0].shape = {N, C, H, W};
        ctx.inputs[0].data = const_cast<float*>(in1);
        ctx.inputs[1].shape = {N, C, H, W};
        ctx.inputs[1].data = const_cast<float*>(in2);
        k.Compute(&ctx);
        return 0;
    } catch (const std::exception& e) {
        std::fprintf(stderr, "vsc_ref_gpu_correlation: %s\n", e.what());
        return 1;
    }
}

int vsc_ref_gpu_warp(const float* in, const float* flow, float* out, size_t out_bytes, int64_t N, int64_t C, int64_t H,
    int64_t W, void* stream)
{
    try {
        OrtKernelInfo info;
        WarpKernel k(g_api, &info, "CUDAExecutionProvider");
        OrtKernelContext ctx;
        OutSlot slot{out, out_bytes};
        ctx.alloc_output = take_output;
        ctx.alloc_user = &slot;
        ctx.gpu_stream = stream;
        ctx.inputs.resize(2);
        ctx.inputs[0].shape = {N, C, H, W};
        ctx.inputs[0].data = const_cast<float*>(in);
        ctx.inputs[1].shape = {N, 2, H, W};
        ctx.inputs[1].data = const_cast<float*>(flow);
        k.Compute(&ctx);
        return 0;
    } catch (const std::exception& e) {
        std::fprintf(stderr, "vsc_ref_gpu_warp: %s\n", e.what());
        return 1;
    }
}

// ---- stabilization: flowconsistency.cuh on caller-owned device buffers -------------------------

int vsc_ref_gpu_warp_result(const float* in, const float* flow, float* out, int W, int H, int flowC)
{
    try {
        View vi(in, W, H, 3), vf(flow, W, H, flowC), vo(out, W, H, 3);
        get_warp_result(vi, vf, vo);
        return 0;
    } catch (const std::exception& e) {
        std::fprintf(stderr, "vsc_ref_gpu_warp_result: %s\n", e.what());
        return 1;
    }
}

int vsc_ref_gpu_adap_comb(const float* crntIn, const float* crntPr, const float* prevWarpIn, const float* prevWarpPr,
    const float* nextWarpIn, const float* nextWarpPr, float* adapCmbIn, float* adapCmbPr, const float* lastStabWarp,
    float alpha, int W, int H)
{
    try {
        View a(crntIn, W, H, 3), b(crntPr, W, H, 3), c(prevWarpIn, W, H, 3), d(prevWarpPr, W, H, 3),
            e(nextWarpIn, W, H, 3), f(nextWarpPr, W, H, 3), g(adapCmbIn, W, H, 3), h(adapCmbPr, W, H, 3),
            i(lastStabWarp, W, H, 3);
        get_adap_comb(a, b, c, d, e, f, g, h, i, alpha);
        return 0;
    } catch (const std::exception& ex) {
        std::fprintf(stderr, "vsc_ref_gpu_adap_comb: %s\n", ex.what());
        return 1;
    }
}

int vsc_ref_gpu_consist_wt(const float* adapCmbIn, const float* crntIn, float* consWt, float beta, float gamma, int W,
    int H)
{
    try {
        View a(adapCmbIn, W, H, 3), b(crntIn, W, H, 3), c(consWt, W, H, 3);
        get_consist_wt(a, b, c, beta, gamma);
        return 0;
    } catch (const std::exception& ex) {
        std::fprintf(stderr, "vsc_ref_gpu_consist_wt: %s\n", ex.what());
        return 1;
    }
}

int vsc_ref_gpu_bilinear(const float* in, int Wi, int Hi, int Ci, float* out, int Wo, int Ho, int Co)
{
    try {
        View a(in, Wi, Hi, Ci), b(out, Wo, Ho, Co);
        get_bilinear(a, b);
        return 0;
    } catch (const std::exception& ex) {
        std::fprintf(stderr, "vsc_ref_gpu_bilinear: %s\n", ex.what());
        return 1;
    }
}

int vsc_ref_gpu_consist_out(const float* crntPr, const float* prevStabWarp, const float* consWt, int numIter,
    float stepSize, float momFac, float* consisOut, int W, int H)
{
    try {
        View a(crntPr, W, H, 3), b(prevStabWarp, W, H, 3), c(consWt, W, H, 3), d(consisOut, W, H, 3);
        get_consist_out(a, b, c, numIter, stepSize, momFac, d);
        return 0;
    } catch (const std::exception& ex) {
        std::fprintf(stderr, "vsc_ref_gpu_consist_out: %s\n", ex.what());
        return 1;
    }
}

// RGBA8888 host bytes -> float3 device image, via GPUImage::copyFromQImage (gpuimage.cpp:103-124)
int vsc_ref_gpu_to_float(const unsigned char* rgba_host, float* out, int W, int H)
{
    try {
        QImage q(W, H, QImage::Format_RGBA8888);
        std::memcpy(q.bits(), rgba_host, static_cast<size_t>(W) * H * 4);
        View o(out, W, H, 3);
        o.img.copyFromQImage(q);
        return 0;
    } catch (const std::exception& ex) {
        std::fprintf(stderr, "vsc_ref_gpu_to_float: %s\n", ex.what());
        return 1;
    }
}

// float3 device image -> RGBA8888 host bytes, via GPUImage::copyToQImage (gpuimage.cpp:127-135)
int vsc_ref_gpu_to_char(const float* in, unsigned char* rgba_host, int W, int H)
{
    try {
        QImage q(W, H, QImage::Format_RGBA8888);
        View i(in, W, H, 3);
        i.img.copyToQImage(q);
        std::memcpy(rgba_host, q.bits(), static_cast<size_t>(W) * H * 4);
        return 0;
    } catch (const std::exception& ex) {
        std::fprintf(stderr, "vsc_ref_gpu_to_char: %s\n", ex.what());
        return 1;
    }
}

// One frame of VideoStabilizer::doOneStep, steps 2-8 of SURVEY 3.2, in the reference's call
// order with the reference's functions and its own GPUImage temporaries (videostabilizer.cpp:177-247).
// All pointers are device float3 HWC images except flows (flowC channels) and rgba_host.
// lastStab is updated in place (the recurrence, :247).  pyramidLevels is 2 as in :109.
struct RefStep {
    int W, H, flowC;
    std::unique_ptr<GPUImage> prevWarpIn, prevWarpPr, nextWarpIn, nextWarpPr, lastStabWarp, consisOut, consWt,
        adapCmbIn, adapCmbPr;
    std::vector<std::unique_ptr<GPUImage>> pyrPr, pyrAdapCmbPr, pyrConsWt, pyrConsisOut;
};

void* vsc_ref_gpu_step_create(int W, int H, int flowC, int pyramidLevels)
{
    try {
        auto* s = new RefStep;
        s->W = W;
        s->H = H;
        s->flowC = flowC;
        auto mk = [&](int w, int h) { return std::unique_ptr<GPUImage>(new GPUImage(w, h, 3)); };
        s->prevWarpIn = mk(W, H);
        s->prevWarpPr = mk(W, H);
        s->nextWarpIn = mk(W, H);
        s->nextWarpPr = mk(W, H);
        s->lastStabWarp = mk(W, H);
        s->consisOut = mk(W, H);
        s->consWt = mk(W, H);
        s->adapCmbIn = mk(W, H);
        s->adapCmbPr = mk(W, H);
        int pw = W, ph = H;
        for (int i = 0; i < pyramidLevels; ++i) {  // videostabilizer.cpp:118-128
            s->pyrPr.push_back(mk(pw, ph));
            s->pyrAdapCmbPr.push_back(mk(pw, ph));
            s->pyrConsWt.push_back(mk(pw, ph));
            s->pyrConsisOut.push_back(mk(pw, ph));
            pw /= 2;
            ph /= 2;
        }
        return s;
    } catch (const std::exception& ex) {
        std::fprintf(stderr, "vsc_ref_gpu_step_create: %s\n", ex.what());
        return nullptr;
    }
}

void vsc_ref_gpu_step_destroy(void* p) { delete static_cast<RefStep*>(p); }

int vsc_ref_gpu_step(void* p, const float* origPrev, const float* origCur, const float* origNext,
    const float* procPrev, const float* procCur, const float* procNext, float* lastStab, const float* flowFwd,
    const float* flowBwd, float alpha, float beta, float gamma, int numIter, float stepSize, float momFac,
    float* consisOutCopy, unsigned char* rgba_host)
{
    try {
        RefStep& s = *static_cast<RefStep*>(p);
        const int W = s.W, H = s.H;
        View o0(origPrev, W, H, 3), o1(origCur, W, H, 3), o2(origNext, W, H, 3);
        View p0(procPrev, W, H, 3), p1(procCur, W, H, 3), p2(procNext, W, H, 3);
        View last(lastStab, W, H, 3), ff(flowFwd, W, H, s.flowC), fb(flowBwd, W, H, s.flowC);

        get_warp_result(o0, fb, *s.prevWarpIn);
        get_warp_result(p0, fb, *s.prevWarpPr);
        get_warp_result(o2, ff, *s.nextWarpIn);
        get_warp_result(p2, ff, *s.nextWarpPr);
        get_warp_result(last, fb, *s.lastStabWarp);

        get_adap_comb(o1, p1, *s.prevWarpIn, *s.prevWarpPr, *s.nextWarpIn, *s.nextWarpPr, *s.adapCmbIn, *s.adapCmbPr,
            *s.lastStabWarp, alpha);
        get_consist_wt(*s.adapCmbIn, o1, *s.consWt, beta, gamma);

        const int levels = static_cast<int>(s.pyrPr.size());
        for (int j = 0; j < levels; ++j) {
            if (j == 0) {
                s.pyrPr[0]->copyFrom(p1.img);
                s.pyrAdapCmbPr[0]->copyFrom(*s.adapCmbPr);
                s.pyrConsWt[0]->copyFrom(*s.consWt);
                s.pyrConsisOut[0]->copyFrom(p1.img);
            } else {
                get_bilinear(*s.pyrPr[j - 1], *s.pyrPr[j]);
                get_bilinear(*s.pyrAdapCmbPr[j - 1], *s.pyrAdapCmbPr[j]);
                get_bilinear(*s.pyrConsWt[j - 1], *s.pyrConsWt[j]);
                get_bilinear(*s.pyrConsisOut[j - 1], *s.pyrConsisOut[j]);
            }
        }
        for (int j = levels - 1; j >= 0; --j) {
            if (j != levels - 1)
                get_bilinear(*s.pyrConsisOut[j + 1], *s.pyrConsisOut[j]);
            get_consist_out(*s.pyrPr[j], *s.pyrAdapCmbPr[j], *s.pyrConsWt[j], numIter / (j + 1), stepSize, momFac,
                *s.pyrConsisOut[j]);
        }
        s.consisOut->copyFrom(*s.pyrConsisOut[0]);

        if (rgba_host) {
            QImage q(W, H, QImage::Format_RGBA8888);
            s.consisOut->copyToQImage(q);
            std::memcpy(rgba_host, q.bits(), static_cast<size_t>(W) * H * 4);
        }
        last.img.copyFrom(*s.consisOut);
        if (consisOutCopy)
            cudaMemcpy(consisOutCopy, s.consisOut->data, sizeof(float) * 3 * W * H, cudaMemcpyDeviceToDevice);
        return cudaDeviceSynchronize() == cudaSuccess ? 0 : 2;
    } catch (const std::exception& ex) {
        std::fprintf(stderr, "vsc_ref_gpu_step: %s\n", ex.what());
        return 1;
    }
}
}
