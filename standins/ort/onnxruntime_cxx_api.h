// STAND-IN for <onnxruntime_cxx_api.h>: the subset of namespace Ort the custom-op sources use.
// See onnxruntime_c_api.h in this directory.
#pragma once
#include "onnxruntime_c_api.h"

#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace Ort {

// ORT_API_MANUAL_INIT semantics: the global api pointer must be set with InitApi before use
inline const OrtApi*& GlobalApi()
{
    static const OrtApi* api = nullptr;
    return api;
}
inline void InitApi(const OrtApi* api) { GlobalApi() = api; }
inline const OrtApi& GetApi()
{
    if (!GlobalApi())
        throw std::runtime_error("Ort::GetApi(): Ort::InitApi was never called (ORT_API_MANUAL_INIT)");
    return *GlobalApi();
}

inline std::vector<std::string> GetAvailableProviders() { return OrtStandinProviders(); }

struct Status {
    explicit Status(OrtStatusPtr p) : p_(p) { }
    ~Status() { delete p_; }
    Status(const Status&) = delete;
    Status& operator=(const Status&) = delete;
    bool IsOK() const { return p_ == nullptr; }
    std::string GetErrorMessage() const { return p_ ? p_->msg : std::string(); }

private:
    OrtStatusPtr p_;
};

struct TensorTypeAndShapeInfo {
    ONNXTensorElementDataType type;
    std::vector<int64_t> shape;
    ONNXTensorElementDataType GetElementType() const { return type; }
    std::vector<int64_t> GetShape() const { return shape; }
};

struct ConstValue {
    const OrtMockTensor* t;
    TensorTypeAndShapeInfo GetTensorTypeAndShapeInfo() const { return {t->type, t->shape}; }
    template <typename T>
    const T* GetTensorData() const
    {
        return static_cast<const T*>(t->data);
    }
};

struct UnownedValue {
    OrtMockTensor* t;
    TensorTypeAndShapeInfo GetTensorTypeAndShapeInfo() const { return {t->type, t->shape}; }
    template <typename T>
    T* GetTensorMutableData()
    {
        return static_cast<T*>(t->data);
    }
};

struct KernelContext {
    explicit KernelContext(OrtKernelContext* c) : c_(c) { }
    size_t GetInputCount() const { return c_->inputs.size(); }
    ConstValue GetInput(size_t i) const { return ConstValue{&c_->inputs.at(i)}; }
    UnownedValue GetOutput(size_t i, const int64_t* dims, size_t rank) const
    {
        if (c_->outputs.size() <= i)
            c_->outputs.resize(i + 1);
        OrtMockTensor& o = c_->outputs[i];
        o.type = c_->inputs.empty() ? ONNX_TENSOR_ELEMENT_DATA_TYPE_FLOAT : c_->inputs[0].type;
        o.shape.assign(dims, dims + rank);
        size_t n = 1;
        for (size_t k = 0; k < rank; ++k)
            n *= static_cast<size_t>(dims[k]);
        const size_t esz = (o.type == ONNX_TENSOR_ELEMENT_DATA_TYPE_DOUBLE) ? 8 : 4;
        if (!c_->alloc_output)
            throw std::runtime_error("stand-in ORT: no output allocator");
        o.data = c_->alloc_output(c_->alloc_user, i, n * esz);
        return UnownedValue{&o};
    }
    void* GetGPUComputeStream() const { return c_->gpu_stream; }

private:
    OrtKernelContext* c_;
};

// owning wrapper the reference's CustomOpRegistry holds (custom_ops.h:33-40)
struct CustomOpDomain {
    explicit CustomOpDomain(const char* name) : p_(new OrtCustomOpDomain{name, {}}) { }
    ~CustomOpDomain() { delete p_; }
    CustomOpDomain(const CustomOpDomain&) = delete;
    CustomOpDomain(CustomOpDomain&& o) : p_(o.p_) { o.p_ = nullptr; }
    operator OrtCustomOpDomain*() const { return p_; }

private:
    OrtCustomOpDomain* p_;
};

// CRTP base: fills the C vtable from TOp's member functions; TKernel needs Compute(OrtKernelContext*)
template <typename TOp, typename TKernel>
struct CustomOpBase : OrtCustomOp {
    CustomOpBase()
    {
        OrtCustomOp::version = ORT_API_VERSION;
        OrtCustomOp::CreateKernel = [](const OrtCustomOp* op, const OrtApi* api, const OrtKernelInfo* info) -> void* {
            return static_cast<const TOp*>(op)->CreateKernel(*api, info);
        };
        OrtCustomOp::GetName = [](const OrtCustomOp* op) { return static_cast<const TOp*>(op)->GetName(); };
        OrtCustomOp::GetExecutionProviderType
            = [](const OrtCustomOp* op) { return static_cast<const TOp*>(op)->GetExecutionProviderType(); };
        OrtCustomOp::GetInputTypeCount
            = [](const OrtCustomOp* op) { return static_cast<const TOp*>(op)->GetInputTypeCount(); };
        OrtCustomOp::GetInputType
            = [](const OrtCustomOp* op, size_t i) { return static_cast<const TOp*>(op)->GetInputType(i); };
        OrtCustomOp::GetOutputTypeCount
            = [](const OrtCustomOp* op) { return static_cast<const TOp*>(op)->GetOutputTypeCount(); };
        OrtCustomOp::GetOutputType
            = [](const OrtCustomOp* op, size_t i) { return static_cast<const TOp*>(op)->GetOutputType(i); };
        OrtCustomOp::KernelCompute
            = [](void* k, OrtKernelContext* ctx) { static_cast<TKernel*>(k)->Compute(ctx); };
        OrtCustomOp::KernelDestroy = [](void* k) { delete static_cast<TKernel*>(k); };
    }
};

}  // namespace Ort

#include "onnxruntime_session_standin.h"   // Env / SessionOptions / IoBinding / Session (flow session)
