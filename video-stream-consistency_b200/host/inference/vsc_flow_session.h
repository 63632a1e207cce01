// Flow-network session on device buffers: the B200 data path of FlowModel::run + InferenceModelVariant::runStatic
// (reference: src/stabilization/flowmodel.cpp:43-98,121-168, src/inference/InferenceModelVariant.cpp:146-184,
// 270-326, src/inference/CudaIO.cpp:112-121, src/stabilization/imagehelpers.cpp:22-40).
//
// Per direction and frame the reference copies batch+2 frames host->host (cpyNImagesToBuffer), optionally resizes
// them on the CPU (QImage::scaled), uploads both network inputs with blocking pageable cudaMemcpy
// (CudaIO::setData), rebuilds Ort::Value / name vectors, runs the session with a provider synchronisation at the
// end, and device-copies the flow into a GPUImage (plus a temporary GPUImage + get_bilinear when FLOWDOWNSCALE > 1).
//
// Here nothing of that touches the host: the frames are already on the device (vsc_stabilizer_push_frame uploaded
// them from pinned memory on the copy stream), vsc_stabilizer_flow_input writes window frames straight into the
// tensors bound to "frame1"/"frame2" (nearest-neighbour scale on the GPU, = Qt::FastTransformation), the session is
// created on the stabilizer's compute stream (OrtCUDAProviderOptions::user_compute_stream) so graph, custom ops and
// stabilization kernels are ordered by ONE stream without any event or synchronisation, the Ort::IoBinding is built
// once, and Run is enqueue-only (disable_synchronize_execution_providers).  The flow stays in the bound output
// buffer and is consumed in place by vsc_stabilizer_step_lowres_flow (which up-scales when the net runs at a lower
// resolution).
#pragma once
#include <onnxruntime_cxx_api.h>

#include <memory>
#include <string>

#include "vsc/vsc.h"

class VscFlowSession {
public:
    // model_path: PWCNet-*-wpreproc.onnx (inputs "frame1","frame2" uint8 [1,netH,netW,4]; output "output" float
    // [1,netH,netW,3], flowmodel.cpp:62-88).  `stabilizer` owns the frames and the compute stream; it must outlive
    // the session.  Throws std::runtime_error / Ort::Exception like InferenceModelVariant::createSession.
    // batch_directions: stabilizeCurrentFrame() runs the graph ONCE on a batch of two frame pairs -- (cur, next) and
    // (next, cur) -- instead of twice on one pair: every custom op then sees N = 2 and half the launches disappear
    // (profiles/time_ops_batched.py: 88 -> 72 us of custom-op GPU time per frame at 1080p/2, 587 -> 515 us for the
    // dense model at 4K).  The model must accept a batch dimension of 2 (the reference feeds [batchSize, H, W, 4],
    // flowmodel.cpp:62-70).
    // batch_size (the CLI's `-b`, main.cpp:48-63; must equal the stabilizer's, vsc_stabilizer_create_batched): the
    // session's tensors are [batch_size, netH, netW, 4] / [batch_size, netH, netW, 3] and the graph runs once per
    // direction every batch_size frames -- on window frames 1..batch_size against 2..batch_size+1
    // (videostabilizer.cpp:269-273, flowmodel.cpp:137-143) -- while the frames in between only index into the two
    // flow batches (:176-179).  Not combinable with batch_directions.
    VscFlowSession(Ort::Env& env, const std::string& model_path, int netW, int netH, vsc_stabilizer* stabilizer,
        int device_id = 0, bool batch_directions = false, int batch_size = 1);
    ~VscFlowSession();
    VscFlowSession(const VscFlowSession&) = delete;
    VscFlowSession& operator=(const VscFlowSession&) = delete;

    // FlowModel::run(frames, results, indexFirst, indexSecond): flow from window frame indexFirst to indexSecond
    // (0 = previous, 1 = current, 2 = next), enqueued on the compute stream; the result lands in output slot
    // `slot` (0 or 1: forward / backward, so that both directions of a frame can be pending).  Returns the device
    // pointer of the [netH, netW, 3] float flow; valid until the next run into the same slot.  Not available on a
    // session created with batch_directions (throws).  With batch_size > 1 the call fills the whole batch (frames
    // indexFirst + b / indexSecond + b) and returns the [batch_size, netH, netW, 3] result.
    const float* run(int indexFirst, int indexSecond, int slot);

    // retrieveOpticalFlow + doOneStep (videostabilizer.cpp:167-279): both directions, then the stabilization step
    // consuming the two flows in place.  out_rgba_host as in vsc_stabilizer_step.
    void stabilizeCurrentFrame(uint8_t* out_rgba_host);

    int netWidth() const { return netW_; }
    int netHeight() const { return netH_; }

private:
    struct Binding;
    int netW_, netH_;
    bool batched_;
    int batch_size_ = 1, step_in_batch_ = 0;
    vsc_stabilizer* st_;
    uint8_t* frame_[2] = {nullptr, nullptr};   // bound inputs ([1,H,W,4] each; batched: [2,H,W,4] each)
    float* flow_[2] = {nullptr, nullptr};      // bound outputs, one per slot (batched: flow_[1] = flow_[0] + H*W*3)
    std::unique_ptr<Ort::Session> session_;
    std::unique_ptr<Binding> bind_[2];         // one persistent IoBinding per output slot
    Ort::RunOptions run_options_;
};
