// TEST INFRASTRUCTURE -- not product code.
//
// Minimal stand-in for <onnxruntime_cxx_api.h>: the subset of namespace Ort the
// reference custom-op sources use.  See onnxruntime_c_api.h in this directory.
#pragma once
#include "onnxruntime_c_api.h"

#include <stdexcept>
#include <string>
#include <vector>

namespace Ort {

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
            throw std::runtime_error("mock ORT: no output allocator");
        o.data = c_->alloc_output(c_->alloc_user, i, n * esz);
        return UnownedValue{&o};
    }
    void* GetGPUComputeStream() const { return c_->gpu_stream; }

private:
    OrtKernelContext* c_;
};

// the reference only inherits from it (CRTP); ORT's real one fills the OrtCustomOp vtable
template <typename TOp, typename TKernel>
struct CustomOpBase { };

}  // namespace Ort
