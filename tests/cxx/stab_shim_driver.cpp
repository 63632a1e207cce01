// TEST DRIVER for the product's stabilization shim (video-stream-consistency_b200/host/stabilization):
// the same role VideoStabilizer plays in the reference -- it owns GPUImage objects, fills them from RGBA host
// frames through GPUImage::copyFromQImage, and runs the doOneStep call sequence (videostabilizer.cpp:177-247)
// through the six flowconsistency.cuh functions, which here resolve to the shim (i.e. to libvsc_b200.so).
// Compiled against the reference's unmodified gpuimage.h / flowconsistency.cuh and the QImage stand-in.
#include "flowIO.h"
#include "flowconsistency.cuh"
#include "gpuimage.h"

#include <cuda_runtime.h>

#include <malloc.h>

#include <chrono>
#include <cstdio>
#include <cstring>
#include <exception>
#include <memory>
#include <vector>

namespace {
struct Stab {
    int W = 0, H = 0, flowC = 3, k = 0;
    std::vector<std::unique_ptr<GPUImage>> orig, proc;  // sliding window, front = prev
    std::unique_ptr<GPUImage> last, flowFwd, flowBwd, prevWarpIn, prevWarpPr, nextWarpIn, nextWarpPr, lastStabWarp,
        consisOut, consWt, adapCmbIn, adapCmbPr;
    std::vector<std::unique_ptr<GPUImage>> pyrPr, pyrAdapCmbPr, pyrConsWt, pyrConsisOut;
};
std::unique_ptr<GPUImage> mk(int w, int h, int c) { return std::unique_ptr<GPUImage>(new GPUImage(w, h, c)); }
}  // namespace

namespace {
// 4K frames are 33 MB: above glibc's largest dynamic mmap threshold (32 MB), so every QImage of the application
// would be a fresh mmap whose pages fault in on first touch (8 ms per frame and image) and are unmapped on free.
// The timing runs keep such blocks on the heap, as a long-running player's allocator does; it changes no result.
struct HeapForFrames {
    HeapForFrames()
    {
        mallopt(M_MMAP_THRESHOLD, 1 << 30);
        mallopt(M_TRIM_THRESHOLD, 1 << 30);
    }
} g_heap_for_frames;
}  // namespace

extern "C" {

// ReadFlowFile of the shim (flowIO.h signature): 0 = ok, 1 = it threw (message copied to msg), 2 = buffer too small
int vsc_shim_read_flo(const char* path, float* buf, size_t cap_floats, int* w, int* h, char* msg, size_t msg_cap)
{
    try {
        std::vector<float> flow;
        ReadFlowFile(flow, *w, *h, path);
        if (flow.size() > cap_floats)
            return 2;
        std::memcpy(buf, flow.data(), flow.size() * sizeof(float));
        return 0;
    } catch (const std::exception& e) {
        if (msg && msg_cap)
            std::snprintf(msg, msg_cap, "%s", e.what());
        return 1;
    }
}

void* vsc_shim_create(int W, int H, int flowC, int levels)
{
    try {
        Stab* s = new Stab;
        s->W = W;
        s->H = H;
        s->flowC = flowC;
        s->last = mk(W, H, 3);
        s->flowFwd = mk(W, H, flowC);
        s->flowBwd = mk(W, H, flowC);
        s->prevWarpIn = mk(W, H, 3);
        s->prevWarpPr = mk(W, H, 3);
        s->nextWarpIn = mk(W, H, 3);
        s->nextWarpPr = mk(W, H, 3);
        s->lastStabWarp = mk(W, H, 3);
        s->consisOut = mk(W, H, 3);
        s->consWt = mk(W, H, 3);
        s->adapCmbIn = mk(W, H, 3);
        s->adapCmbPr = mk(W, H, 3);
        int pw = W, ph = H;
        for (int i = 0; i < levels; ++i) {
            s->pyrPr.push_back(mk(pw, ph, 3));
            s->pyrAdapCmbPr.push_back(mk(pw, ph, 3));
            s->pyrConsWt.push_back(mk(pw, ph, 3));
            s->pyrConsisOut.push_back(mk(pw, ph, 3));
            pw /= 2;
            ph /= 2;
        }
        return s;
    } catch (const std::exception& e) {
        std::fprintf(stderr, "vsc_shim_create: %s\n", e.what());
        return nullptr;
    }
}

void vsc_shim_destroy(void* p) { delete static_cast<Stab*>(p); }

// loadFrame: imageToGPU x2 (imagehelpers.cpp:16-20); the third frame also seeds lastStabilizedFrame (:152)
int vsc_shim_push(void* p, const unsigned char* orig_rgba, const unsigned char* proc_rgba)
{
    try {
        Stab& s = *static_cast<Stab*>(p);
        QImage qo(s.W, s.H, QImage::Format_RGBA8888), qp(s.W, s.H, QImage::Format_RGBA8888);
        std::memcpy(qo.bits(), orig_rgba, static_cast<size_t>(s.W) * s.H * 4);
        std::memcpy(qp.bits(), proc_rgba, static_cast<size_t>(s.W) * s.H * 4);
        auto o = mk(s.W, s.H, 3), q = mk(s.W, s.H, 3);
        o->copyFromQImage(qo);
        q->copyFromQImage(qp);
        s.orig.push_back(std::move(o));
        s.proc.push_back(std::move(q));
        if (s.proc.size() == 3 && s.k == 0) {
            s.last->copyFrom(*s.proc.back());
            s.k = 1;
        }
        return 0;
    } catch (const std::exception& e) {
        std::fprintf(stderr, "vsc_shim_push: %s\n", e.what());
        return 1;
    }
}

// TIMING: `reps` times the doOneStep call sequence from the flow copies to copyToQImage (videostabilizer.cpp:177-238)
// on the window as it stands (nothing is popped); flows are uploaded once before.  Returns host-clock ms per step in
// *ms_per_step.  Every shim call is synchronous, like the reference's.
int vsc_shim_time_steps(void* p, const float* flowFwd_host, const float* flowBwd_host, int reps, int numIter,
    double* ms_per_step)
{
    try {
        Stab& s = *static_cast<Stab*>(p);
        if (s.orig.size() != 3)
            return 3;
        const int W = s.W, H = s.H;
        GPUImage resF(W, H, s.flowC), resB(W, H, s.flowC);   // flowResultsFwd/Bwd[batchIdx]
        resF.copyFrom(std::vector<float>(flowFwd_host, flowFwd_host + static_cast<size_t>(W) * H * s.flowC));
        resB.copyFrom(std::vector<float>(flowBwd_host, flowBwd_host + static_cast<size_t>(W) * H * s.flowC));
        const int levels = static_cast<int>(s.pyrPr.size());
        cudaDeviceSynchronize();
        const auto t0 = std::chrono::steady_clock::now();
        for (int r = 0; r < reps; ++r) {
            s.flowFwd->copyFrom(resF);
            s.flowBwd->copyFrom(resB);
            get_warp_result(*s.orig[0], *s.flowBwd, *s.prevWarpIn);
            get_warp_result(*s.proc[0], *s.flowBwd, *s.prevWarpPr);
            get_warp_result(*s.orig[2], *s.flowFwd, *s.nextWarpIn);
            get_warp_result(*s.proc[2], *s.flowFwd, *s.nextWarpPr);
            get_warp_result(*s.last, *s.flowBwd, *s.lastStabWarp);
            get_adap_comb(*s.orig[1], *s.proc[1], *s.prevWarpIn, *s.prevWarpPr, *s.nextWarpIn, *s.nextWarpPr,
                *s.adapCmbIn, *s.adapCmbPr, *s.lastStabWarp, 6800.0f);
            get_consist_wt(*s.adapCmbIn, *s.orig[1], *s.consWt, 6800.0f, 2.0f);
            for (int j = 0; j < levels; ++j) {
                if (j == 0) {
                    s.pyrPr[0]->copyFrom(*s.proc[1]);
                    s.pyrAdapCmbPr[0]->copyFrom(*s.adapCmbPr);
                    s.pyrConsWt[0]->copyFrom(*s.consWt);
                    s.pyrConsisOut[0]->copyFrom(*s.proc[1]);
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
                get_consist_out(*s.pyrPr[j], *s.pyrAdapCmbPr[j], *s.pyrConsWt[j], numIter / (j + 1), 0.15f, 0.15f,
                    *s.pyrConsisOut[j]);
            }
            s.consisOut->copyFrom(*s.pyrConsisOut[0]);
            QImage q(W, H, QImage::Format_RGBA8888);   // a fresh image per frame, as videostabilizer.cpp:237
            s.consisOut->copyToQImage(q);
            s.last->copyFrom(*s.consisOut);
        }
        *ms_per_step = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count() / reps;
        return 0;
    } catch (const std::exception& e) {
        std::fprintf(stderr, "vsc_shim_time_steps: %s\n", e.what());
        return 1;
    }
}

int vsc_shim_step(void* p, const float* flowFwd_host, const float* flowBwd_host, float alpha, float beta, float gamma,
    int numIter, float stepSize, float momFac, unsigned char* rgba_out, float* consis_out_host)
{
    try {
        Stab& s = *static_cast<Stab*>(p);
        if (s.orig.size() != 3)
            return 3;
        const int W = s.W, H = s.H;
        std::vector<float> ff(flowFwd_host, flowFwd_host + static_cast<size_t>(W) * H * s.flowC);
        std::vector<float> fb(flowBwd_host, flowBwd_host + static_cast<size_t>(W) * H * s.flowC);
        s.flowFwd->copyFrom(ff);
        s.flowBwd->copyFrom(fb);
        get_warp_result(*s.orig[0], *s.flowBwd, *s.prevWarpIn);
        get_warp_result(*s.proc[0], *s.flowBwd, *s.prevWarpPr);
        get_warp_result(*s.orig[2], *s.flowFwd, *s.nextWarpIn);
        get_warp_result(*s.proc[2], *s.flowFwd, *s.nextWarpPr);
        get_warp_result(*s.last, *s.flowBwd, *s.lastStabWarp);
        get_adap_comb(*s.orig[1], *s.proc[1], *s.prevWarpIn, *s.prevWarpPr, *s.nextWarpIn, *s.nextWarpPr, *s.adapCmbIn,
            *s.adapCmbPr, *s.lastStabWarp, alpha);
        get_consist_wt(*s.adapCmbIn, *s.orig[1], *s.consWt, beta, gamma);
        const int levels = static_cast<int>(s.pyrPr.size());
        for (int j = 0; j < levels; ++j) {
            if (j == 0) {
                s.pyrPr[0]->copyFrom(*s.proc[1]);
                s.pyrAdapCmbPr[0]->copyFrom(*s.adapCmbPr);
                s.pyrConsWt[0]->copyFrom(*s.consWt);
                s.pyrConsisOut[0]->copyFrom(*s.proc[1]);
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
        if (rgba_out) {
            QImage q(W, H, QImage::Format_RGBA8888);
            s.consisOut->copyToQImage(q);
            std::memcpy(rgba_out, q.bits(), static_cast<size_t>(W) * H * 4);
        }
        if (consis_out_host)
            cudaMemcpy(consis_out_host, s.consisOut->data, sizeof(float) * 3 * W * H, cudaMemcpyDeviceToHost);
        s.last->copyFrom(*s.consisOut);
        s.orig.erase(s.orig.begin());
        s.proc.erase(s.proc.begin());
        return 0;
    } catch (const std::exception& e) {
        std::fprintf(stderr, "vsc_shim_step: %s\n", e.what());
        return 1;
    }
}
}
