// Drop-in implementation of the reference's stabilization interface on top of the C ABI
// (include/vsc/vsc.h).  It REPLACES three reference translation units in the host's build --
// src/stabilization/flowconsistency.cu, gpuimage.cu and gpuimage.cpp -- and is compiled against
// the reference's own, unmodified headers (flowconsistency.cuh:3-50, gpuimage.h:8-31), which it
// finds on the include path; VideoStabilizer, FlowModel, StreamStabilizer and imagehelpers.cpp
// then link against it unchanged.
//
// Interface contract kept: the six free functions take GPUImage&, return void and
// are synchronous -- each call returns with its result visible to a following blocking cudaMemcpy
// (here: enqueue on one private stream, then cudaStreamSynchronize); CUDA errors print and
// exit(1) like checkError (flowconsistency.cu:25-31); GPUImage keeps its public fields, deep-copy
// constructor and the six copy* methods with the reference's messages and exceptions.
// Ordering: the reference ran every kernel and every cudaMemcpy on the legacy default stream, so a
// D2D copyFrom followed by a get_* call was ordered implicitly.  Here the kernels run on a private
// non-blocking stream, which never orders itself against the legacy stream, and a D2D cudaMemcpy does not
// block the host -- so every GPUImage copy* method is issued on the SAME private stream and then
// synchronised: copyFrom -> get_* -> copyFrom sequences of the unmodified host code stay ordered.
// What changes underneath: no per-call cudaMalloc/cudaFree (get_consist_out's scratch and the
// RGBA staging buffers are cached and only ever grow), no cudaDeviceSynchronize, sm_100a kernels.
#include "flowconsistency.cuh"
#include "gpuimage.h"

#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <mutex>
#include <stdexcept>
#include <unordered_map>
#include <vector>

#include "vsc/vsc.h"

namespace {

struct Shim {
    cudaStream_t stream = nullptr;
    void* ws = nullptr;
    size_t ws_bytes = 0;
    unsigned char* rgba = nullptr;
    size_t rgba_bytes = 0;
    std::mutex mu;  // the reference is single-threaded here (Qt main thread); the lock makes misuse safe

    cudaStream_t s()
    {
        if (!stream && cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking) != cudaSuccess)
            stream = nullptr;  // falls back to the legacy default stream
        return stream;
    }
    void* workspace(size_t bytes)
    {
        if (bytes > ws_bytes) {
            if (ws)
                cudaFree(ws);
            ws = nullptr;
            ws_bytes = 0;
            if (cudaMalloc(&ws, bytes) != cudaSuccess)
                throw std::runtime_error("Unable to allocate CUDA memory.");
            ws_bytes = bytes;
        }
        return ws;
    }
    unsigned char* staging(size_t bytes)
    {
        if (bytes > rgba_bytes) {
            if (rgba)
                cudaFree(rgba);
            rgba = nullptr;
            rgba_bytes = 0;
            if (cudaMalloc(reinterpret_cast<void**>(&rgba), bytes) != cudaSuccess)
                throw std::runtime_error("Unable to allocate CUDA memory.");
            rgba_bytes = bytes;
        }
        return rgba;
    }
};

Shim& shim()
{
    static Shim s;
    return s;
}

// checkError + cudaDeviceSynchronize of the reference launchers, for one stream
void finish(int rc, const char* stage)
{
    cudaError_t e = rc > 0 ? static_cast<cudaError_t>(rc) : cudaSuccess;
    if (rc == 0)
        e = cudaStreamSynchronize(shim().s());
    if (rc < 0 || e != cudaSuccess) {
        std::cerr << "CUDA error at " << stage << ": " << (rc < 0 ? rc : static_cast<int>(e)) << ", "
                  << (rc < 0 ? vsc_error_string(rc) : cudaGetErrorString(e)) << std::endl;
        std::exit(1);
    }
}

// a cudaMemcpy of the reference, on the shim's stream (see "Ordering" above); false on failure
bool copy_sync(void* dst, const void* src, size_t bytes, cudaMemcpyKind kind)
{
    std::lock_guard<std::mutex> lock(shim().mu);
    cudaStream_t st = shim().s();
    if (cudaMemcpyAsync(dst, src, bytes, kind, st) != cudaSuccess)
        return false;
    return cudaStreamSynchronize(st) == cudaSuccess;
}

}  // namespace

// ------------------------------------------------------------------------------ flowconsistency.cuh
void get_bilinear(GPUImage& input, GPUImage& output)
{
    std::lock_guard<std::mutex> lock(shim().mu);
    finish(vsc_bilinear(input.data, input.width, input.height, input.channels, output.data, output.width,
               output.height, output.channels, shim().s()),
        "kernel_bilinear");
}

void get_warp_result(GPUImage& input, GPUImage& flow, GPUImage& inputWarp)
{
    std::lock_guard<std::mutex> lock(shim().mu);
    finish(vsc_warp_hwc3(input.data, flow.data, inputWarp.data, inputWarp.width, inputWarp.height, flow.channels,
               shim().s()),
        "kernel_warp");
}

void get_adap_comb(GPUImage& crntIn, GPUImage& crntPr, GPUImage& prevWarpIn, GPUImage& prevWarpPr,
    GPUImage& nextWarpIn, GPUImage& nextWarpPr, GPUImage& adapCmbIn, GPUImage& adapCmbPr, GPUImage& lastStabWarp,
    float alpha)
{
    std::lock_guard<std::mutex> lock(shim().mu);
    finish(vsc_adap_comb(crntIn.data, crntPr.data, prevWarpIn.data, prevWarpPr.data, nextWarpIn.data,
               nextWarpPr.data, adapCmbIn.data, adapCmbPr.data, lastStabWarp.data, alpha, crntIn.width,
               crntIn.height, shim().s()),
        "kernel_adap_comb");
}

void get_consist_wt(GPUImage& adapCmbIn, GPUImage& crntIn, GPUImage& consistWt, float beta, float gamma)
{
    std::lock_guard<std::mutex> lock(shim().mu);
    finish(vsc_consist_wt(adapCmbIn.data, crntIn.data, consistWt.data, beta, gamma, crntIn.width, crntIn.height,
               shim().s()),
        "kernel_consist_wt");
}

void get_consist_out(GPUImage& crntPr, GPUImage& prevStabWarp, GPUImage& consWt, int numIter, float stepSize,
    float momFac, GPUImage& consisOut)
{
    std::lock_guard<std::mutex> lock(shim().mu);
    const size_t need = vsc_consist_solve_workspace_bytes(consisOut.width, consisOut.height);
    void* ws = shim().workspace(need);
    finish(vsc_consist_solve(crntPr.data, prevStabWarp.data, consWt.data, numIter, stepSize, momFac, consisOut.data,
               consisOut.width, consisOut.height, ws, need, shim().s()),
        "kernel_consist_out");
}

// dead in the reference (no caller): the average of the neighbouring processed frames
void perform_consistency(GPUImage&, GPUImage& processedPrev, GPUImage&, GPUImage& processedNext, GPUImage&,
    GPUImage&, GPUImage&, GPUImage& stabilizedOut)
{
    (void)processedPrev;
    (void)processedNext;
    (void)stabilizedOut;
    std::cerr << "perform_consistency: not implemented (dead code in the reference, flowconsistency.cu:269-286)"
              << std::endl;
    std::exit(1);
}

// ------------------------------------------------------------------------------ gpuimage.h
// The application creates and destroys two GPUImages per frame (imageToGPU in loadFrame, pop_front in doOneStep,
// stabilizestream.cpp:72-73, videostabilizer.cpp:248-250) plus temporaries in FlowModel::run: with cudaMalloc /
// cudaFree underneath that is four driver calls with an implicit device synchronisation per frame.  Freed blocks are
// kept in a small per-size cache instead (every GPUImage operation here is synchronous, so a block is idle when its
// image dies); cudaMalloc is only reached when a size is seen for the first time.
namespace {
struct BlockCache {
    static constexpr size_t kPerSize = 8;                 // blocks kept per byte size
    static constexpr size_t kMaxBytes = size_t(3) << 30;   // and in total
    std::mutex mu;
    std::unordered_map<size_t, std::vector<void*>> free_blocks;
    size_t cached = 0;

    void* get(size_t bytes)
    {
        {
            std::lock_guard<std::mutex> lock(mu);
            auto it = free_blocks.find(bytes);
            if (it != free_blocks.end() && !it->second.empty()) {
                void* p = it->second.back();
                it->second.pop_back();
                cached -= bytes;
                return p;
            }
        }
        void* p = nullptr;
        if (cudaMalloc(&p, bytes ? bytes : 1) == cudaSuccess)
            return p;
        (void)cudaGetLastError();
        drop_all();   // out of memory: give the cached blocks back and try once more
        if (cudaMalloc(&p, bytes ? bytes : 1) != cudaSuccess)
            throw std::runtime_error("Unable to allocate CUDA memory.");
        return p;
    }
    void put(void* p, size_t bytes)
    {
        if (!p)
            return;
        {
            std::lock_guard<std::mutex> lock(mu);
            std::vector<void*>& v = free_blocks[bytes];
            if (v.size() < kPerSize && cached + bytes <= kMaxBytes) {
                v.push_back(p);
                cached += bytes;
                return;
            }
        }
        cudaFree(p);
    }
    void drop_all()
    {
        std::lock_guard<std::mutex> lock(mu);
        for (auto& kv : free_blocks)
            for (void* p : kv.second)
                cudaFree(p);
        free_blocks.clear();
        cached = 0;
    }
};
BlockCache& blocks()
{
    static BlockCache* c = new BlockCache;   // never destroyed: images with static storage may die after it would
    return *c;
}
size_t image_bytes(int w, int h, int c) { return static_cast<size_t>(w) * h * c * sizeof(float); }
}  // namespace

GPUImage::GPUImage(int width, int height, int channels) : width(width), height(height), channels(channels)
{
    data = static_cast<float*>(blocks().get(image_bytes(width, height, channels)));
}

GPUImage::GPUImage(const GPUImage& other) : GPUImage(other.width, other.height, other.channels) { copyFrom(other); }

GPUImage::~GPUImage() { blocks().put(data, image_bytes(width, height, channels)); }

void GPUImage::copyFrom(const GPUImage& other)
{
    if (other.width == width && other.height == height && other.channels == channels) {
        const size_t nelems = static_cast<size_t>(width) * height * channels;
        if (!copy_sync(data, other.data, nelems * sizeof(float), cudaMemcpyDeviceToDevice))
            throw std::runtime_error("Unable to copy data from device to device.");
    } else {
        std::cerr << "Invalid dimensions" << std::endl;
    }
}

void GPUImage::copyFrom(const std::vector<float>& vec)
{
    const size_t nelems = static_cast<size_t>(width) * height * channels;
    if (nelems == vec.size()) {
        if (!copy_sync(data, vec.data(), nelems * sizeof(float), cudaMemcpyHostToDevice))
            throw std::runtime_error("Unable to copy data from device to host.");
    } else {
        std::cerr << "Invalid dimensions" << std::endl;
    }
}

void GPUImage::copyFrom(const std::vector<std::byte>& vec)
{
    const size_t nelems = static_cast<size_t>(width) * height * channels;
    if (nelems * sizeof(float) == vec.size())
        (void)copy_sync(data, vec.data(), nelems * sizeof(float), cudaMemcpyHostToDevice);
    else
        throw std::runtime_error("Invalid dimensions. Length of byte-array differs from expected number of floats");
}

void GPUImage::copyFromCudaBuffer(const void* resourcePointer, size_t byteSize)
{
    const size_t nelems = static_cast<size_t>(width) * height * channels;
    if (nelems * sizeof(float) == byteSize) {
        if (!copy_sync(data, resourcePointer, byteSize, cudaMemcpyDeviceToDevice))
            throw std::runtime_error("Unable to copy data from device to host.");
    } else {
        throw std::runtime_error("Invalid dimensions. Length of byte-array differs from expected number of floats");
    }
}

void GPUImage::copyFromQImage(const QImage& image)
{
    if (image.width() != width || image.height() != height) {
        std::cerr << "Invalid dimensions" << std::endl;
        return;
    }
    if (channels != 3) {
        std::cerr << "copyFromQImage requires a 3 channel GPU image" << std::endl;
        return;
    }
    const size_t bytes = static_cast<size_t>(width) * height * 4;
    const QImage rgba = image.format() != QImage::Format_RGBA8888 ? image.convertToFormat(QImage::Format_RGBA8888) : image;
    std::lock_guard<std::mutex> lock(shim().mu);
    unsigned char* dev = shim().staging(bytes);
    if (cudaMemcpyAsync(dev, rgba.bits(), bytes, cudaMemcpyHostToDevice, shim().s()) != cudaSuccess)
        throw std::runtime_error("Unable to copy data from host to device.");
    finish(vsc_rgba8_to_f32x3(dev, data, width, height, shim().s()), "kernel_to_float_image");
}

void GPUImage::copyToQImage(QImage& image) const
{
    if (image.format() != QImage::Format_RGBA8888) {
        std::cerr << "copyToQImage needs to be in format RGBA8888" << std::endl;
        return;
    }
    const size_t bytes = static_cast<size_t>(width) * height * 4;
    std::lock_guard<std::mutex> lock(shim().mu);
    unsigned char* dev = shim().staging(bytes);
    const int rc = vsc_f32x3_to_rgba8(data, dev, width, height, shim().s());
    if (rc == 0 && cudaMemcpyAsync(image.bits(), dev, bytes, cudaMemcpyDeviceToHost, shim().s()) != cudaSuccess)
        throw std::runtime_error("Unable to copy data from device to host.");
    finish(rc, "kernel_to_char_image");
}
