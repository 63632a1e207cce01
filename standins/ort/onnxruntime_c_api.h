// STAND-IN for <onnxruntime_c_api.h> (onnxruntime 1.20.1 is not in this image; the reference
// downloads it at configure time, CMakeLists.txt:95-122).  Ours, not a copy of ORT: it declares
// only what the custom-op sources touch, with an in-memory "session" behind it, so that
//   * the reference's custom-op sources (basekernel.h, correlation.h, warp.h, correlation.cc,
//     warp.cc, correlation_cuda.cc, warp_cuda.cc, custom_ops.cpp) compile UNMODIFIED into
//     oracle/_ref/ and can be driven on host or device buffers (oracle/refdrv/*), and
//   * the product's ORT shim (video-stream-consistency_b200/host/ort_custom_ops) can be
//     compiled and exercised end to end -- RegisterCustomOps -> domain -> op -> CreateKernel ->
//     KernelCompute -- without onnxruntime (tests/test_host_shims*.py).
// With the real onnxruntime headers on the include path instead, the same sources build the
// real plug-in.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

#define ORT_API_VERSION 20
#define ORT_API_CALL
#define ORT_STANDIN 1

enum ONNXTensorElementDataType {
    ONNX_TENSOR_ELEMENT_DATA_TYPE_UNDEFINED = 0,
    ONNX_TENSOR_ELEMENT_DATA_TYPE_FLOAT = 1,
    ONNX_TENSOR_ELEMENT_DATA_TYPE_UINT8 = 2,
    ONNX_TENSOR_ELEMENT_DATA_TYPE_INT64 = 7,
    ONNX_TENSOR_ELEMENT_DATA_TYPE_DOUBLE = 11
};

struct OrtStatus {
    std::string msg;
};
typedef OrtStatus* OrtStatusPtr;

// attributes of one graph node, as the stand-in session holds them
struct OrtKernelInfo {
    bool has_legacy = true;
    bool has_max_displacement = true;
    int64_t legacy = 0;
    int64_t max_displacement = 4;
};

// a tensor the stand-in session owns: host or device pointer, the kernels do not care
struct OrtMockTensor {
    ONNXTensorElementDataType type = ONNX_TENSOR_ELEMENT_DATA_TYPE_FLOAT;
    std::vector<int64_t> shape;
    void* data = nullptr;
};

struct OrtKernelContext {
    std::vector<OrtMockTensor> inputs;
    std::vector<OrtMockTensor> outputs;
    // output allocator supplied by the driver: returns a buffer of `bytes` bytes
    void* (*alloc_output)(void* user, size_t index, size_t bytes) = nullptr;
    void* alloc_user = nullptr;
    void* gpu_stream = nullptr;
};

struct OrtApi;

// the C vtable ORT calls a custom op through (subset)
struct OrtCustomOp {
    uint32_t version = ORT_API_VERSION;
    void* (*CreateKernel)(const OrtCustomOp* op, const OrtApi* api, const OrtKernelInfo* info) = nullptr;
    const char* (*GetName)(const OrtCustomOp* op) = nullptr;
    const char* (*GetExecutionProviderType)(const OrtCustomOp* op) = nullptr;
    ONNXTensorElementDataType (*GetInputType)(const OrtCustomOp* op, size_t index) = nullptr;
    size_t (*GetInputTypeCount)(const OrtCustomOp* op) = nullptr;
    ONNXTensorElementDataType (*GetOutputType)(const OrtCustomOp* op, size_t index) = nullptr;
    size_t (*GetOutputTypeCount)(const OrtCustomOp* op) = nullptr;
    void (*KernelCompute)(void* op_kernel, OrtKernelContext* context) = nullptr;
    void (*KernelDestroy)(void* op_kernel) = nullptr;
};

struct OrtCustomOpDomain {
    std::string name;
    std::vector<const OrtCustomOp*> ops;
};

struct OrtSessionOptions {
    std::vector<OrtCustomOpDomain*> domains;
};

struct OrtApi {
    OrtStatusPtr KernelInfoGetAttribute_int64(const OrtKernelInfo* info, const char* name, int64_t* out) const
    {
        const std::string n(name);
        if (n == "legacy" && info->has_legacy) {
            *out = info->legacy;
            return nullptr;
        }
        if (n == "max_displacement" && info->has_max_displacement) {
            *out = info->max_displacement;
            return nullptr;
        }
        return new OrtStatus{"attribute '" + n + "' not found"};
    }
    OrtStatusPtr CreateCustomOpDomain(const char* domain, OrtCustomOpDomain** out) const
    {
        *out = new OrtCustomOpDomain{domain, {}};
        return nullptr;
    }
    OrtStatusPtr CustomOpDomain_Add(OrtCustomOpDomain* domain, const OrtCustomOp* op) const
    {
        for (const OrtCustomOp* o : domain->ops)
            if (std::string(o->GetName(o)) == op->GetName(op)
                && std::string(o->GetExecutionProviderType(o)) == op->GetExecutionProviderType(op))
                return new OrtStatus{"custom op registered twice for the same provider"};
        domain->ops.push_back(op);
        return nullptr;
    }
    OrtStatusPtr AddCustomOpDomain(OrtSessionOptions* options, OrtCustomOpDomain* domain) const
    {
        options->domains.push_back(domain);
        return nullptr;
    }
    void ReleaseCustomOpDomain(OrtCustomOpDomain* domain) const { delete domain; }
    void ReleaseStatus(OrtStatus* s) const { delete s; }
    const char* GetErrorMessage(const OrtStatus* s) const { return s ? s->msg.c_str() : ""; }
    OrtStatusPtr CreateStatus(int /*code*/, const char* msg) const { return new OrtStatus{msg}; }
};

struct OrtApiBase {
    const OrtApi* (*GetApi)(uint32_t version);
    const char* (*GetVersionString)();
};

// what the stand-in "runtime" reports as available execution providers (driver-settable)
inline std::vector<std::string>& OrtStandinProviders()
{
    static std::vector<std::string> p = {"CUDAExecutionProvider", "CPUExecutionProvider"};
    return p;
}

inline const OrtApiBase* OrtGetApiBase()
{
    static const OrtApi api{};
    static const OrtApiBase base{[](uint32_t v) -> const OrtApi* { return v <= ORT_API_VERSION ? &api : nullptr; },
        []() -> const char* { return "1.20.1-standin"; }};
    return &base;
}
