// TEST DRIVER for the product's VideoStabilizer drop-in (video-stream-consistency_b200/host/stabilization/
// vsc_videostabilizer.cpp), compiled against the reference's UNMODIFIED videostabilizer.h / flowmodel.h /
// imagehelpers.h / gpuimage.h / inference headers and the Qt stand-ins.  It plays the parts of the application that
// are not on the hot path:
//   * a StreamStabilizer-like subclass: loadFrame() appends frames from memory exactly like
//     stabilizestream.cpp:72-73 (originalFramesQt << image; originalFrames << imageToGPU(...)), outputFrame() keeps
//     the emitted QImage;
//   * FlowModel / OrtContext / InferenceModelVariant stand-ins: FlowModel::run keeps the reference's data flow
//     (flowmodel.cpp:121-168: batchSize pairs from window indices indexFirst + b / indexSecond + b, nearest-neighbour
//     resize to the network size, flow up-scaled with get_bilinear into results[b]) around a stand-in "network":
//         flow[h,w,3] = bilinear sample of float3(frame1) displaced by float3(frame2).rg      (vsc_warp_hwc3)
//     so that a test can reproduce every flow exactly through the Python binding.
#include "videostabilizer.h"

#include <cuda_runtime_api.h>

#include <malloc.h>

#include <chrono>
#include <cstring>
#include <exception>
#include <string>
#include <vector>

#include "flowconsistency.cuh"
#include "vsc/vsc.h"

// ---- stand-ins for src/inference and flowmodel.cpp (ORT is not in this image) --------------------------------
namespace Ort {
struct Env { };
struct Session { };
struct MemoryInfo { };
struct RunOptions { };
}  // namespace Ort

OrtContext::OrtContext() : mEnvironment(new Ort::Env), mEnvironmentRef(*mEnvironment) { }
OrtContext::~OrtContext() { }
InferenceModelVariant::~InferenceModelVariant() { }

namespace {
std::string g_error;
long g_flow_runs = 0;
double g_flow_ms = 0, g_load_ms = 0, g_step_ms = 0, g_out_ms = 0;   // host clock: FlowModel::run / loadFrame / doOneStep / outputFrame

struct Clock {
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    double ms() const { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(); }
};

void cuda_ok(cudaError_t e, const char* what)
{
    if (e != cudaSuccess)
        throw std::runtime_error(std::string(what) + ": " + cudaGetErrorString(e));
}
void vsc_ok(int rc, const char* what)
{
    if (rc)
        throw std::runtime_error(std::string(what) + ": " + vsc_error_string(rc));
}
}  // namespace

FlowModel::FlowModel(QString modelChoice, OrtContext* ort_context, int batchSize, int width, int height)
    : _ort_context(ort_context), batchSize(batchSize), width(width), height(height)
{
    if (modelChoice != QString("pwcnet-light") && modelChoice != QString("pwcnet"))
        throw std::runtime_error("Unknown model choice");   // flowmodel.cpp:57-59
}

void FlowModel::run(QList<QSharedPointer<QImage>>& originalFramesQt, QList<QSharedPointer<GPUImage>>& results,
    int indexFirst, int indexSecond, flowTiming* timing)
{
    ++g_flow_runs;
    const Clock clk;
    const int fw = originalFramesQt[0]->width(), fh = originalFramesQt[0]->height();
    const size_t fpx = static_cast<size_t>(fw) * fh, npx = static_cast<size_t>(width) * height;
    uint8_t *full = nullptr, *net1 = nullptr, *net2 = nullptr;
    float *a = nullptr, *b = nullptr;
    cuda_ok(cudaMalloc(reinterpret_cast<void**>(&full), fpx * 4), "cudaMalloc");
    cuda_ok(cudaMalloc(reinterpret_cast<void**>(&net1), npx * 4), "cudaMalloc");
    cuda_ok(cudaMalloc(reinterpret_cast<void**>(&net2), npx * 4), "cudaMalloc");
    cuda_ok(cudaMalloc(reinterpret_cast<void**>(&a), npx * 12), "cudaMalloc");
    cuda_ok(cudaMalloc(reinterpret_cast<void**>(&b), npx * 12), "cudaMalloc");
    auto to_net = [&](const QImage& img, uint8_t* dst) {   // QImage::scaled(FastTransformation) + CudaIO::setData
        cuda_ok(cudaMemcpy(full, img.bits(), fpx * 4, cudaMemcpyHostToDevice), "cudaMemcpy");
        if (fw == width && fh == height)
            cuda_ok(cudaMemcpy(dst, full, fpx * 4, cudaMemcpyDeviceToDevice), "cudaMemcpy");
        else
            vsc_ok(vsc_rgba8_scale_nearest(full, fw, fh, dst, width, height, nullptr), "vsc_rgba8_scale_nearest");
    };
    for (int n = 0; n < batchSize; ++n) {
        to_net(*originalFramesQt[indexFirst + n], net1);
        to_net(*originalFramesQt[indexSecond + n], net2);
        vsc_ok(vsc_rgba8_to_f32x3(net1, a, width, height, nullptr), "vsc_rgba8_to_f32x3");
        vsc_ok(vsc_rgba8_to_f32x3(net2, b, width, height, nullptr), "vsc_rgba8_to_f32x3");
        GPUImage flowLowRes(width, height, 3);
        vsc_ok(vsc_warp_hwc3(a, b, flowLowRes.data, width, height, 3, nullptr), "vsc_warp_hwc3");
        cuda_ok(cudaDeviceSynchronize(), "cudaDeviceSynchronize");
        if (fw != width || fh != height)
            get_bilinear(flowLowRes, *results[n].get());   // flowmodel.cpp:158-161
        else
            results[n]->copyFromCudaBuffer(flowLowRes.data, npx * 12);
    }
    cudaFree(full);
    cudaFree(net1);
    cudaFree(net2);
    cudaFree(a);
    cudaFree(b);
    timing->runTime += 1;
    g_flow_ms += clk.ms();
}

void FlowModel::runFlowVis(QList<QSharedPointer<GPUImage>>&, flowTiming*) { }

// ---- the application side -------------------------------------------------------------------------------------
namespace {

class MemoryStabilizer : public VideoStabilizer {
public:
    MemoryStabilizer(int W, int H, int batch, int T, const uint8_t* orig, const uint8_t* proc, uint8_t* outs, int* have)
        : VideoStabilizer(W, H, batch, std::optional<QString>(QString("pwcnet-light")), true), W_(W), H_(H), T_(T),
          orig_(orig), proc_(proc), outs_(outs), have_(have)
    {
    }
    int stabilizeAll()   // StreamStabilizer::stabilizeAll (stabilizestream.cpp:141-154) without the decoder thread
    {
        preloadProcessedFrames();
        g_flow_ms = g_load_ms = g_step_ms = g_out_ms = 0;   // the accumulators cover the doOneStep loop only
        timer.start();
        int steps = 0;
        for (int i = k;; i++) {
            ++steps;
            const Clock clk;
            const bool more = doOneStep(i);
            g_step_ms += clk.ms();
            if (!more)
                break;
        }
        return steps;
    }

protected:
    bool loadFrame(int) override   // like StreamStabilizer::loadFrame: the next frame of the stream, whatever the index
    {
        const Clock clk;
        const int i = next_++;
        if (i >= T_)
            return false;
        const size_t bytes = static_cast<size_t>(W_) * H_ * 4;
        QSharedPointer<QImage> o(new QImage(W_, H_, QImage::Format_RGBA8888));
        QImage p(W_, H_, QImage::Format_RGBA8888);
        std::memcpy(o->bits(), orig_ + i * bytes, bytes);
        std::memcpy(p.bits(), proc_ + i * bytes, bytes);
        originalFramesQt << o;                       // stabilizestream.cpp:71-73
        originalFrames << imageToGPU(*o.get());
        processedFrames << imageToGPU(p);
        g_load_ms += clk.ms();
        return true;
    }
    void outputFrame(int i, QSharedPointer<QImage> q) override
    {
        if (i < 0 || i >= T_)
            return;
        const Clock clk;
        std::memcpy(outs_ + static_cast<size_t>(i) * W_ * H_ * 4, q->bits(), static_cast<size_t>(W_) * H_ * 4);
        have_[i] += 1;
        g_out_ms += clk.ms();
    }

private:
    int W_, H_, T_, next_ = 0;
    const uint8_t *orig_, *proc_;
    uint8_t* outs_;
    int* have_;
};

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

const char* vsc_vs_test_last_error() { return g_error.c_str(); }
long vsc_vs_test_flow_runs() { return g_flow_runs; }
// accumulated host-clock milliseconds since the last call: {doOneStep total, of which FlowModel::run, of which
// loadFrame, of which outputFrame}
void vsc_vs_test_timing(double* out4)
{
    out4[0] = g_step_ms;
    out4[1] = g_flow_ms;
    out4[2] = g_load_ms;
    out4[3] = g_out_ms;
    g_step_ms = g_flow_ms = g_load_ms = g_out_ms = 0;
}

// Runs a whole clip of T frames through VideoStabilizer (preload, doOneStep until loadFrame fails, final frames).
// outs[T][H][W][4] receives every emitted frame, have[T] how often frame i was emitted.  numIter / gamma <= 0: defaults.
// Returns the number of doOneStep calls, or -1 (see vsc_vs_test_last_error).
int vsc_vs_test_run(int W, int H, int T, int batch, const uint8_t* orig, const uint8_t* proc, uint8_t* outs, int* have,
    int numIter, float gamma)
{
    try {
        MemoryStabilizer s(W, H, batch, T, orig, proc, outs, have);
        if (numIter > 0)
            s.getHyperParams()->numIter = numIter;   // what the GUI sliders do (hyperparameterwidget.cpp:112-128)
        if (gamma > 0)
            s.getHyperParams()->gamma = gamma;
        return s.stabilizeAll();
    } catch (const std::exception& e) {
        g_error = e.what();
        return -1;
    }
}
}
