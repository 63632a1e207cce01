// See vsc_flow_session.h.  Builds against onnxruntime 1.20.1 (CUDA provider) or against standins/ort for the
// compile check and tests/test_host_shims.py::test_flow_session_*.
#include "vsc_flow_session.h"

#include <cuda_runtime_api.h>

#include <stdexcept>
#include <vector>

// the custom-op library's entry point (host/ort_custom_ops/vsc_custom_ops.cpp; custom_ops.h:49)
extern "C" OrtStatus* ORT_API_CALL RegisterCustomOps(OrtSessionOptions* options, const OrtApiBase* api);

namespace {

void cuda_or_throw(cudaError_t e, const char* what)
{
    if (e != cudaSuccess)
        throw std::runtime_error(std::string(what) + ": " + cudaGetErrorString(e));   // CudaIO.cpp:44-51 texts
}

void vsc_or_throw(int rc, const char* what)
{
    if (rc != VSC_OK)
        throw std::runtime_error(std::string(what) + ": " + vsc_error_string(rc));
}

}  // namespace

struct VscFlowSession::Binding {
    explicit Binding(Ort::Session& s) : io(s) { }
    Ort::IoBinding io;
    std::vector<Ort::Value> values;   // keep the bound tensors alive as long as the binding
};

VscFlowSession::VscFlowSession(Ort::Env& env, const std::string& model_path, int netW, int netH,
    vsc_stabilizer* stabilizer, int device_id, bool batch_directions, int batch_size)
    : netW_(netW), netH_(netH), batched_(batch_directions), batch_size_(batch_size), st_(stabilizer)
{
    if (!stabilizer || netW <= 0 || netH <= 0 || batch_size < 1 || (batch_directions && batch_size != 1))
        throw std::runtime_error("VscFlowSession: invalid arguments");
    if (vsc_stabilizer_batch_size(stabilizer) != batch_size)
        throw std::runtime_error("VscFlowSession: the stabilizer was created for another flow batch size");

    // InferenceModelVariant::createSession (:146-184), with the graph placed on the stabilizer's compute stream
    Ort::SessionOptions options;
    options.SetIntraOpNumThreads(1);
    options.SetGraphOptimizationLevel(ORT_ENABLE_ALL);
    if (OrtStatus* status = RegisterCustomOps(options, OrtGetApiBase())) {
        const std::string msg = OrtGetApiBase()->GetApi(ORT_API_VERSION)->GetErrorMessage(status);
        OrtGetApiBase()->GetApi(ORT_API_VERSION)->ReleaseStatus(status);
        throw std::runtime_error("RegisterCustomOps: " + msg);
    }
    OrtCUDAProviderOptions cuda{};
    cuda.device_id = device_id;
    cuda.has_user_compute_stream = 1;
    cuda.user_compute_stream = vsc_stabilizer_compute_stream(st_);
    options.AppendExecutionProvider_CUDA(cuda);
    session_ = std::make_unique<Ort::Session>(env, model_path.c_str(), options);

    // CudaIO's allocations (CudaIO.cpp:37-52), once; zeroed like the reference's
    const int64_t batch = batched_ ? 2 : batch_size_;
    const size_t frame_bytes = static_cast<size_t>(netW) * netH * 4;
    const size_t flow_bytes = static_cast<size_t>(netW) * netH * 3 * sizeof(float);
    for (int i = 0; i < 2; ++i) {
        cuda_or_throw(cudaMalloc(reinterpret_cast<void**>(&frame_[i]), batch * frame_bytes),
            "Unable to allocate CUDA memory.");
        cuda_or_throw(cudaMemset(frame_[i], 0, batch * frame_bytes), "Unable to zero out CUDA memory.");
    }
    if (batched_) {   // one [2,H,W,3] output: slot 1 is its second sample
        cuda_or_throw(cudaMalloc(reinterpret_cast<void**>(&flow_[0]), 2 * flow_bytes), "Unable to allocate CUDA memory.");
        cuda_or_throw(cudaMemset(flow_[0], 0, 2 * flow_bytes), "Unable to zero out CUDA memory.");
        flow_[1] = flow_[0] + static_cast<size_t>(netW) * netH * 3;
    } else {
        for (int i = 0; i < 2; ++i) {
            cuda_or_throw(cudaMalloc(reinterpret_cast<void**>(&flow_[i]), batch * flow_bytes),
                "Unable to allocate CUDA memory.");
            cuda_or_throw(cudaMemset(flow_[i], 0, batch * flow_bytes), "Unable to zero out CUDA memory.");
        }
    }

    // runStatic's tensors and binding (:270-313) -- built once instead of per run
    const Ort::MemoryInfo device_memory("Cuda", OrtArenaAllocator, device_id, OrtMemTypeDefault);
    const int64_t frame_shape[4] = {batch, netH, netW, 4};
    const int64_t flow_shape[4] = {batch, netH, netW, 3};
    for (int slot = 0; slot < (batched_ ? 1 : 2); ++slot) {
        bind_[slot] = std::make_unique<Binding>(*session_);
        Binding& b = *bind_[slot];
        b.values.reserve(3);
        // slot 1 binds the two frame buffers the other way round: the backward run of a frame re-uses the
        // forward run's inputs (flow 2 -> 1 after flow 1 -> 2) without writing them again
        b.values.push_back(Ort::Value::CreateTensor(device_memory, frame_[slot], batch * frame_bytes, frame_shape, 4,
            ONNX_TENSOR_ELEMENT_DATA_TYPE_UINT8));
        b.values.push_back(Ort::Value::CreateTensor(device_memory, frame_[1 - slot], batch * frame_bytes, frame_shape, 4,
            ONNX_TENSOR_ELEMENT_DATA_TYPE_UINT8));
        b.values.push_back(Ort::Value::CreateTensor(device_memory, flow_[slot], batch * flow_bytes, flow_shape, 4,
            ONNX_TENSOR_ELEMENT_DATA_TYPE_FLOAT));
        b.io.BindInput("frame1", b.values[0]);
        b.io.BindInput("frame2", b.values[1]);
        b.io.BindOutput("output", b.values[2]);
    }
    // the compute stream orders everything downstream; no provider synchronisation at the end of Run
    run_options_.AddConfigEntry("disable_synchronize_execution_providers", "1");
}

VscFlowSession::~VscFlowSession()
{
    if (st_)
        vsc_stabilizer_sync(st_);   // pending runs read / write the buffers below
    bind_[0].reset();
    bind_[1].reset();
    session_.reset();
    for (int i = 0; i < 2; ++i)
        cudaFree(frame_[i]);
    cudaFree(flow_[0]);
    if (!batched_)
        cudaFree(flow_[1]);
}

const float* VscFlowSession::run(int indexFirst, int indexSecond, int slot)
{
    if (batched_)
        throw std::runtime_error("VscFlowSession::run: the session runs both directions as one batch");
    if (slot < 0 || slot > 1)
        throw std::runtime_error("VscFlowSession::run: slot must be 0 or 1");
    // cpyNImagesToBuffer + QImage::scaled + CudaIO::setData (flowmodel.cpp:126-144) on the device
    const size_t frame_bytes = static_cast<size_t>(netW_) * netH_ * 4;
    for (int b = 0; b < batch_size_; ++b) {
        vsc_or_throw(vsc_stabilizer_flow_input(st_, indexFirst + b, frame_[slot] + b * frame_bytes, netW_, netH_),
            "flow input frame1");
        vsc_or_throw(vsc_stabilizer_flow_input(st_, indexSecond + b, frame_[1 - slot] + b * frame_bytes, netW_, netH_),
            "flow input frame2");
    }
    session_->Run(run_options_, bind_[slot]->io);
    return flow_[slot];
}

void VscFlowSession::stabilizeCurrentFrame(uint8_t* out_rgba_host)
{
    if (batched_) {
        // frame1 = [cur, next], frame2 = [next, cur]: sample 0 is the forward pair (:271), sample 1 the backward (:272)
        const size_t frame_bytes = static_cast<size_t>(netW_) * netH_ * 4;
        vsc_or_throw(vsc_stabilizer_flow_input(st_, 1, frame_[0], netW_, netH_), "flow input cur");
        vsc_or_throw(vsc_stabilizer_flow_input(st_, 2, frame_[0] + frame_bytes, netW_, netH_), "flow input next");
        cudaStream_t stream = static_cast<cudaStream_t>(vsc_stabilizer_compute_stream(st_));
        cuda_or_throw(cudaMemcpyAsync(frame_[1], frame_[0] + frame_bytes, frame_bytes, cudaMemcpyDeviceToDevice, stream),
            "Unable to copy data from device to device.");
        cuda_or_throw(cudaMemcpyAsync(frame_[1] + frame_bytes, frame_[0], frame_bytes, cudaMemcpyDeviceToDevice, stream),
            "Unable to copy data from device to device.");
        session_->Run(run_options_, bind_[0]->io);
        vsc_or_throw(vsc_stabilizer_step_lowres_flow(st_, flow_[0], flow_[1], netW_, netH_, out_rgba_host),
            "stabilizer step");
        return;
    }
    // every batch_size frames (videostabilizer.cpp:269): both directions for the whole batch
    if (step_in_batch_ == 0) {
        run(1, 2, 0);                                // :271
        session_->Run(run_options_, bind_[1]->io);   // :272, run(2, 1): the same frames, bound swapped
    }
    const size_t flow_elems = static_cast<size_t>(netW_) * netH_ * 3;
    const float* fwd = flow_[0] + step_in_batch_ * flow_elems;   // flowResultsFwd[batchIdx] (:176-179)
    const float* bwd = flow_[1] + step_in_batch_ * flow_elems;
    step_in_batch_ = (step_in_batch_ + 1) % batch_size_;
    vsc_or_throw(vsc_stabilizer_step_lowres_flow(st_, fwd, bwd, netW_, netH_, out_rgba_host), "stabilizer step");
}
