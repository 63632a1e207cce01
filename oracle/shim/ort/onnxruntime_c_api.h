// TEST INFRASTRUCTURE -- not product code.
//
// Minimal stand-in for <onnxruntime_c_api.h> (onnxruntime 1.20.1 is not in this
// image; the reference downloads it at configure time, CMakeLists.txt:95-122).
// It declares only what the reference's custom-op sources touch
// (basekernel.h, correlation.h, warp.h, correlation.cc, warp.cc,
// correlation_cuda.cc, warp_cuda.cc), so that those files compile UNMODIFIED
// from /root/reference into oracle/_ref/ and can be driven on host (or device)
// buffers by oracle/refdrv/*.  The same header also lets the ORT
// registration shim of the product (host/ort_custom_ops.cpp) be compile-checked.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

#define ORT_API_VERSION 20
#define ORT_API_CALL
#define ORT_MOCK_STANDIN 1

enum ONNXTensorElementDataType {
    ONNX_TENSOR_ELEMENT_DATA_TYPE_UNDEFINED = 0,
    ONNX_TENSOR_ELEMENT_DATA_TYPE_FLOAT = 1,
    ONNX_TENSOR_ELEMENT_DATA_TYPE_INT64 = 7,
    ONNX_TENSOR_ELEMENT_DATA_TYPE_DOUBLE = 11
};

struct OrtStatus {
    std::string msg;
};
typedef OrtStatus* OrtStatusPtr;

// attributes of one graph node, as the mock session would hold them
struct OrtKernelInfo {
    bool has_legacy = true;
    bool has_max_displacement = true;
    int64_t legacy = 0;
    int64_t max_displacement = 4;
};

// a tensor the mock "session" owns: host or device pointer, the kernels do not care
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

struct OrtSessionOptions;
struct OrtCustomOpDomain;
struct OrtCustomOp;

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
};

struct OrtApiBase {
    const OrtApi* (*GetApi)(uint32_t version);
    const char* (*GetVersionString)();
};
