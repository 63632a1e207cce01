// onnxruntime custom-op library on top of the C ABI (include/vsc/vsc.h): the drop-in for the
// reference's `-CustomOps` shared library (reference: src/ort_custom_ops/src/custom_ops.cpp:17-97,
// include/ort_custom_ops/opticalflow/correlation.h:10-84, warp.h:11-60, basekernel.h:13-27,
// src/opticalflow/correlation_cuda.cc:29-138, warp_cuda.cc:28-77).
//
// Same exported symbol (RegisterCustomOps), same domain ("custom"), same op names ("Correlation",
// "Warp"), same attributes (required int64 `legacy`, `max_displacement`), same tensor layouts and
// output shapes ({N,P,P,H,W} / {N,P*P,H,W} / {N,C,H,W}), same error behaviour (std::runtime_error
// out of the kernel constructor / Compute).  Differences, all deliberate:
//   * only the CUDAExecutionProvider kernels are registered: this library has no CPU path
//     (a CPU-EP session keeps using the reference's own CPU kernels);
//   * Compute enqueues on ORT's compute stream and returns: no cudaMalloc / cudaFree / implicit
//     device synchronisation inside an op (the reference does both per call,
//     correlation_cuda.cu:369-370,440-441);
//   * Ort::InitApi is called explicitly (the reference defines ORT_API_MANUAL_INIT but relies on
//     symbol interposition with the host's copy of the API pointer; SURVEY 8b pitfall);
//   * op objects are function-local statics instead of leaked `new`s.
// Builds against the real onnxruntime 1.20.1 headers, or against standins/ort for the
// compile check and the registration tests in this repository.
#ifndef ORT_API_MANUAL_INIT
#define ORT_API_MANUAL_INIT
#include <onnxruntime_cxx_api.h>
#undef ORT_API_MANUAL_INIT
#else
#include <onnxruntime_cxx_api.h>
#endif

#include <cstdio>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "vsc/vsc.h"

#if defined(_MSC_VER)
#define VSC_ORT_EXPORT __declspec(dllexport)
#else
#define VSC_ORT_EXPORT __attribute__((visibility("default")))
#endif

namespace {

const char* const kCudaProvider = "CUDAExecutionProvider";
const char* const kDomain = "custom";

void throw_on(int rc, const char* what)
{
    if (rc != VSC_OK)
        throw std::runtime_error(std::string(what) + ": " + vsc_error_string(rc));
}

std::vector<int64_t> nchw_or_throw(const Ort::ConstValue& v, const char* what)
{
    const auto info = v.GetTensorTypeAndShapeInfo();
    if (info.GetElementType() != ONNX_TENSOR_ELEMENT_DATA_TYPE_FLOAT)
        throw std::runtime_error("Unsupported input type. Must be float.");  // GetInputType only admits FLOAT
    std::vector<int64_t> d = info.GetShape();
    if (d.size() != 4)
        throw std::runtime_error(std::string(what) + ": expected a 4-D NCHW tensor");
    for (int64_t x : d)
        if (x <= 0 || x > 0x7fffffff)
            throw std::runtime_error(std::string(what) + ": bad dimension");
    return d;
}

struct VscCorrelationKernel {
    VscCorrelationKernel(const OrtApi& api, const OrtKernelInfo* info)
    {
        // required attributes, same messages as correlation.h:19-31
        int64_t legacy = 0;
        {
            Ort::Status st(api.KernelInfoGetAttribute_int64(info, "legacy", &legacy));
            if (!st.IsOK())
                throw std::runtime_error("Error reading attribute 'legacy', with error: " + st.GetErrorMessage());
        }
        legacy_ = legacy != 0;
        Ort::Status st2(api.KernelInfoGetAttribute_int64(info, "max_displacement", &max_displacement_));
        if (!st2.IsOK())
            throw std::runtime_error(
                "Error reading attribute 'max_displacement', with error: " + st2.GetErrorMessage());
        if (max_displacement_ < 0 || max_displacement_ > 1024)
            throw std::runtime_error("attribute 'max_displacement' out of range");
    }

    void Compute(OrtKernelContext* context)
    {
        Ort::KernelContext ctx{context};
        Ort::ConstValue in1 = ctx.GetInput(0);
        Ort::ConstValue in2 = ctx.GetInput(1);
        const std::vector<int64_t> d = nchw_or_throw(in1, "Correlation input 0");
        if (nchw_or_throw(in2, "Correlation input 1") != d)
            throw std::runtime_error("Correlation: the two inputs must have the same shape");
        const int64_t N = d[0], C = d[1], H = d[2], W = d[3];
        const int64_t P = 2 * max_displacement_ + 1;
        // correlation_cuda.cc:69-76
        const std::vector<int64_t> odims
            = legacy_ ? std::vector<int64_t>{N, P * P, H, W} : std::vector<int64_t>{N, P, P, H, W};
        auto out = ctx.GetOutput(0, odims.data(), odims.size());
        // the op must run on ORT's compute stream (correlation_cuda.cc:79-80)
        vsc_stream_t stream = reinterpret_cast<vsc_stream_t>(ctx.GetGPUComputeStream());
        throw_on(vsc_correlation_f32(in1.GetTensorData<float>(), in2.GetTensorData<float>(),
                     out.GetTensorMutableData<float>(), static_cast<int>(N), static_cast<int>(C),
                     static_cast<int>(H), static_cast<int>(W), static_cast<int>(max_displacement_), legacy_ ? 1 : 0,
                     stream),
            "custom::Correlation");
    }

private:
    bool legacy_ = false;
    int64_t max_displacement_ = 4;
};

struct VscWarpKernel {
    VscWarpKernel(const OrtApi&, const OrtKernelInfo*) { }

    void Compute(OrtKernelContext* context)
    {
        Ort::KernelContext ctx{context};
        Ort::ConstValue x = ctx.GetInput(0);
        Ort::ConstValue flow = ctx.GetInput(1);
        const std::vector<int64_t> d = nchw_or_throw(x, "Warp input");
        const std::vector<int64_t> f = nchw_or_throw(flow, "Warp flow");
        if (f[0] != d[0] || f[1] != 2 || f[2] != d[2] || f[3] != d[3])
            throw std::runtime_error("Warp: flow must be [N,2,H,W]");
        auto out = ctx.GetOutput(0, d.data(), d.size());  // warp_cuda.cc:43-45
        vsc_stream_t stream = reinterpret_cast<vsc_stream_t>(ctx.GetGPUComputeStream());
        throw_on(vsc_warp_nchw_f32(x.GetTensorData<float>(), flow.GetTensorData<float>(),
                     out.GetTensorMutableData<float>(), static_cast<int>(d[0]), static_cast<int>(d[1]),
                     static_cast<int>(d[2]), static_cast<int>(d[3]), stream),
            "custom::Warp");
    }
};

// correlation.h:60-84
struct VscCorrelationOp : Ort::CustomOpBase<VscCorrelationOp, VscCorrelationKernel> {
    void* CreateKernel(const OrtApi& api, const OrtKernelInfo* info) const { return new VscCorrelationKernel(api, info); }
    const char* GetName() const { return "Correlation"; }
    const char* GetExecutionProviderType() const { return kCudaProvider; }
    size_t GetInputTypeCount() const { return 2; }
    ONNXTensorElementDataType GetInputType(size_t) const { return ONNX_TENSOR_ELEMENT_DATA_TYPE_FLOAT; }
    size_t GetOutputTypeCount() const { return 1; }
    ONNXTensorElementDataType GetOutputType(size_t) const { return ONNX_TENSOR_ELEMENT_DATA_TYPE_FLOAT; }
};

// warp.h:38-60
struct VscWarpOp : Ort::CustomOpBase<VscWarpOp, VscWarpKernel> {
    void* CreateKernel(const OrtApi& api, const OrtKernelInfo* info) const { return new VscWarpKernel(api, info); }
    const char* GetName() const { return "Warp"; }
    const char* GetExecutionProviderType() const { return kCudaProvider; }
    size_t GetInputTypeCount() const { return 2; }
    ONNXTensorElementDataType GetInputType(size_t) const { return ONNX_TENSOR_ELEMENT_DATA_TYPE_FLOAT; }
    size_t GetOutputTypeCount() const { return 1; }
    ONNXTensorElementDataType GetOutputType(size_t) const { return ONNX_TENSOR_ELEMENT_DATA_TYPE_FLOAT; }
};

// domains must outlive the sessions that reference them (custom_ops.cpp:55-71)
struct DomainDeleter {
    const OrtApi* api;
    void operator()(OrtCustomOpDomain* d) const { api->ReleaseCustomOpDomain(d); }
};
std::vector<std::unique_ptr<OrtCustomOpDomain, DomainDeleter>>& domains()
{
    static std::vector<std::unique_ptr<OrtCustomOpDomain, DomainDeleter>> v;
    return v;
}
std::mutex& domains_mutex()
{
    static std::mutex m;
    return m;
}

}  // namespace

extern "C" VSC_ORT_EXPORT OrtStatus* ORT_API_CALL RegisterCustomOps(OrtSessionOptions* options, const OrtApiBase* api_base)
{
    const OrtApi* api = api_base->GetApi(ORT_API_VERSION);
    if (api == nullptr) {  // custom_ops.cpp:77-83: report and return "no status"
        std::fprintf(stderr, "vsc custom ops: onnxruntime %s does not provide API version %d\n",
            api_base->GetVersionString(), static_cast<int>(ORT_API_VERSION));
        return nullptr;
    }
    Ort::InitApi(api);

    static const VscCorrelationOp correlation_op;
    static const VscWarpOp warp_op;

    OrtCustomOpDomain* domain = nullptr;
    if (OrtStatus* st = api->CreateCustomOpDomain(kDomain, &domain))
        return st;
    {
        std::lock_guard<std::mutex> lock(domains_mutex());
        domains().emplace_back(domain, DomainDeleter{api});
    }
    if (OrtStatus* st = api->CustomOpDomain_Add(domain, &correlation_op))
        return st;
    if (OrtStatus* st = api->CustomOpDomain_Add(domain, &warp_op))
        return st;
    return api->AddCustomOpDomain(options, domain);
}
