// vsc_stabilizer: the per-stream pipeline object (include/vsc/vsc.h), the native counterpart of
// VideoStabilizer's per-frame recurrence (reference: src/stabilization/videostabilizer.cpp:46-153,
// 167-265; GPUImage::copyFromQImage / copyToQImage, gpuimage.cpp:103-135).
//
// What the reference does per frame and what this object does instead:
//   * 13 member GPUImages + 8 pyramid images, plus per-call cudaMalloc/cudaFree in imageToGPU,
//     copyToQImage and get_consist_out          -> every buffer allocated once at creation;
//   * blocking pageable cudaMemcpy H2D per frame -> pinned staging + cudaMemcpyAsync on a copy
//     stream; the upload and u8->f32 conversion of frame t+2 overlap the solve of frame t
//     (4 window slots: prev, cur, next + the one being filled), ordered by events only;
//   * 8 full-image D2D copies (flowFwd/Bwd, pyramid level 0, consisOut, lastStabilizedFrame)
//                                                -> pointer swaps, no copies;
//   * 240 launches each followed by cudaDeviceSynchronize -> fused stage A + solver sweeps
//     enqueued on one compute stream, no host synchronisation inside a step;
//   * blocking D2H of the 8-bit result           -> async D2H on a third stream into pinned memory
//     (double buffered), overlapping the next frame.
// One object serves one video stream on the device that was current at creation; the recurrence
// makes frames of a stream strictly sequential, so multi-GPU = one object per GPU (no collective).
#include "vsc_common.cuh"

#include <cstring>
#include <new>
#include <string>
#include <thread>

namespace {

constexpr int kMaxBatch = 16;                 // flow batch sizes 1..16 (main.cpp:48-63: -b)
constexpr int kMaxSlots = 2 + kMaxBatch + 1;  // window of 2k + batchSize frames (k = 1) + the slot being filled

struct Pending {
    const uint8_t* src = nullptr;  // pinned result
    uint8_t* dst = nullptr;        // caller's pageable buffer
    size_t bytes = 0;
};

}  // namespace

struct vsc_stabilizer {
    int W = 0, H = 0, flowC = 3, device = 0;
    int batch = 1, wcap = 3, nslots = 4;   // flow batch size, window capacity 2k + batch, ring slots wcap + 1
    size_t P = 0, n = 0;
    vsc_hyper_params hp{};
    cudaStream_t compute = nullptr, copy = nullptr, d2h = nullptr;

    // window ring
    float* orig[kMaxSlots] = {};
    float* proc[kMaxSlots] = {};
    uint8_t* stage_dev[kMaxSlots][2] = {};
    uint8_t* stage_pin[kMaxSlots][2] = {};
    cudaEvent_t slot_ready[kMaxSlots] = {};   // copy stream: upload + conversion of the slot finished
    cudaEvent_t slot_h2d[kMaxSlots] = {};     // copy stream: H2D from the slot's pinned staging finished
    cudaEvent_t slot_released[kMaxSlots] = {};  // compute stream: last step reading the slot finished
    bool slot_used[kMaxSlots] = {};
    int head = 0, count = 0;
    long long pushed = 0;

    float *lastStab = nullptr, *consisOut = nullptr;
    float* flowUp[2] = {};
    float* flowIn[2] = {};      // device landing buffers for host flows (frame-sized)
    float* flowPin[2] = {};     // pinned staging for pageable host flows
    cudaEvent_t flow_ready = nullptr, flow_free = nullptr;
    bool flow_used = false;
    void* ws = nullptr;
    size_t ws_bytes = 0;
    int ws_levels = 0;

    // file mode (-f <flowdir>): two sets of pinned landing buffers (allocated on first use), so that a worker
    // thread can read the .flo pair of the next frame while this frame's pair is uploaded and consumed
    float* filePin[2][2] = {};
    cudaEvent_t file_h2d[2] = {};   // copy stream: H2D out of the set finished
    bool file_used[2] = {};
    int file_set = 0;               // set the next read goes to
    std::thread pf_thread;
    bool pf_active = false;
    int pf_set = 0, pf_frame = 0, pf_rc = 0;
    std::string pf_dir;

    uint8_t* out_dev[2] = {};
    uint8_t* out_pin[2] = {};
    cudaEvent_t out_conv[2] = {};  // compute stream: f32->u8 written
    cudaEvent_t out_done[2] = {};  // d2h stream: D2H finished
    bool out_used[2] = {};
    Pending pending[2];
    int out_idx = 0;
};

namespace {

using namespace vsc;

int cu(cudaError_t e) { return e == cudaSuccess ? VSC_OK : static_cast<int>(e); }

bool is_pinned(const void* p)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        (void)cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeHost;
}

void flush_pending(vsc_stabilizer* s, int k)
{
    Pending& pd = s->pending[k];
    if (pd.dst) {
        cudaEventSynchronize(s->out_done[k]);
        std::memcpy(pd.dst, pd.src, pd.bytes);
        pd = Pending{};
    }
}

int ensure_workspace(vsc_stabilizer* s, int levels)
{
    if (levels == s->ws_levels)
        return VSC_OK;
    const size_t need = vsc_frame_stabilize_workspace_bytes(s->W, s->H, levels);
    if (need == 0)
        return VSC_E_INVALID;
    if (need > s->ws_bytes) {
        // pyramidLevels is fixed at 2 in the reference (videostabilizer.cpp:109); a change of the level count is
        // the one case where the object re-allocates (outside any timed loop)
        cudaStreamSynchronize(s->compute);
        if (s->ws)
            cudaFree(s->ws);
        s->ws = nullptr;
        s->ws_bytes = 0;
        const int rc = cu(cudaMalloc(&s->ws, need));
        if (rc)
            return rc;
        s->ws_bytes = need;
    }
    s->ws_levels = levels;
    return VSC_OK;
}

int do_step(vsc_stabilizer* s, const float* flowFwd, const float* flowBwd, uint8_t* out_host)
{
    if (s->count < 3)
        return VSC_E_STATE;
    const vsc_hyper_params c = s->hp;  // snapshot once per frame (videostabilizer.cpp:192)
    int rc = ensure_workspace(s, c.pyramidLevels);
    if (rc)
        return rc;
    const int s0 = s->head, s1 = (s->head + 1) % s->nslots, s2 = (s->head + 2) % s->nslots;
    for (int k : {s0, s1, s2})
        if ((rc = cu(cudaStreamWaitEvent(s->compute, s->slot_ready[k], 0))))
            return rc;

    rc = vsc_frame_stabilize(s->orig[s0], s->orig[s1], s->orig[s2], s->proc[s0], s->proc[s1], s->proc[s2],
        s->lastStab, flowFwd, flowBwd, s->flowC, &c, s->consisOut, s->W, s->H, s->ws, s->ws_bytes, s->compute);
    if (rc)
        return rc;

    if (out_host) {
        const int k = s->out_idx;
        s->out_idx ^= 1;
        flush_pending(s, k);
        if (s->out_used[k] && (rc = cu(cudaStreamWaitEvent(s->compute, s->out_done[k], 0))))
            return rc;
        rc = vsc_f32x3_to_rgba8(s->consisOut, s->out_dev[k], s->W, s->H, s->compute);
        if (rc)
            return rc;
        cudaEventRecord(s->out_conv[k], s->compute);
        cudaStreamWaitEvent(s->d2h, s->out_conv[k], 0);
        const size_t bytes = s->P * 4;
        if (is_pinned(out_host)) {
            rc = cu(cudaMemcpyAsync(out_host, s->out_dev[k], bytes, cudaMemcpyDeviceToHost, s->d2h));
        } else {
            rc = cu(cudaMemcpyAsync(s->out_pin[k], s->out_dev[k], bytes, cudaMemcpyDeviceToHost, s->d2h));
            s->pending[k] = Pending{s->out_pin[k], out_host, bytes};
        }
        if (rc)
            return rc;
        cudaEventRecord(s->out_done[k], s->d2h);
        s->out_used[k] = true;
    }

    // recurrence (videostabilizer.cpp:247): lastStabilizedFrame <- consisOut, as a pointer swap
    float* t = s->lastStab;
    s->lastStab = s->consisOut;
    s->consisOut = t;
    // the three window slots were read by this step; slot s0 leaves the window (:248-250)
    for (int k : {s0, s1, s2})
        cudaEventRecord(s->slot_released[k], s->compute);
    s->head = s1;
    s->count -= 1;
    return cu(cudaGetLastError());
}

}  // namespace

extern "C" int vsc_stabilizer_create(vsc_stabilizer** out, int W, int H, int flow_channels)
{
    return vsc_stabilizer_create_batched(out, W, H, flow_channels, 1);
}

extern "C" int vsc_stabilizer_batch_size(const vsc_stabilizer* s) { return s ? s->batch : 0; }
extern "C" int vsc_stabilizer_window_count(const vsc_stabilizer* s) { return s ? s->count : 0; }

extern "C" int vsc_stabilizer_create_batched(vsc_stabilizer** out, int W, int H, int flow_channels, int batchSize)
{
    if (!out || W < 2 || H < 2 || H > 65535 || (flow_channels != 2 && flow_channels != 3) || batchSize < 1
        || batchSize > kMaxBatch)
        return VSC_E_INVALID;
    *out = nullptr;
    vsc_stabilizer* s = new (std::nothrow) vsc_stabilizer;
    if (!s)
        return VSC_E_INVALID;
    s->W = W;
    s->H = H;
    s->flowC = flow_channels;
    s->batch = batchSize;
    s->wcap = 2 + batchSize;
    s->nslots = s->wcap + 1;
    s->P = static_cast<size_t>(W) * H;
    s->n = s->P * 3;
    vsc_hyper_params_default(&s->hp);
    int rc = cu(cudaGetDevice(&s->device));
    const size_t fb = s->n * sizeof(float);
    auto dmalloc = [&](void** p, size_t b) {
        if (!rc)
            rc = cu(cudaMalloc(p, b));
    };
    auto hmalloc = [&](void** p, size_t b) {
        if (!rc)
            rc = cu(cudaHostAlloc(p, b, cudaHostAllocDefault));
    };
    auto mkevent = [&](cudaEvent_t* e) {
        if (!rc)
            rc = cu(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
    };
    // The compute stream gets the highest priority, the copy streams the lowest: a solver pass fills every SM
    // (one CTA each, all registers), so the conversion kernels of the frame being uploaded can only start when a
    // pass retires -- with equal priorities they then delay the next pass; with low priority they wait for the
    // level-1 passes, whose narrower CTAs leave half of each SM free.
    int prio_lo = 0, prio_hi = 0;
    if (!rc) rc = cu(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));   // lo = numerically greatest
    if (!rc) rc = cu(cudaStreamCreateWithPriority(&s->compute, cudaStreamNonBlocking, prio_hi));
    if (!rc) rc = cu(cudaStreamCreateWithPriority(&s->copy, cudaStreamNonBlocking, prio_lo));
    if (!rc) rc = cu(cudaStreamCreateWithPriority(&s->d2h, cudaStreamNonBlocking, prio_lo));
    for (int k = 0; k < s->nslots; ++k) {
        dmalloc(reinterpret_cast<void**>(&s->orig[k]), fb);
        dmalloc(reinterpret_cast<void**>(&s->proc[k]), fb);
        for (int i = 0; i < 2; ++i) {
            dmalloc(reinterpret_cast<void**>(&s->stage_dev[k][i]), s->P * 4);
            hmalloc(reinterpret_cast<void**>(&s->stage_pin[k][i]), s->P * 4);
        }
        mkevent(&s->slot_ready[k]);
        mkevent(&s->slot_h2d[k]);
        mkevent(&s->slot_released[k]);
    }
    dmalloc(reinterpret_cast<void**>(&s->lastStab), fb);
    dmalloc(reinterpret_cast<void**>(&s->consisOut), fb);
    for (int i = 0; i < 2; ++i) {
        dmalloc(reinterpret_cast<void**>(&s->flowUp[i]), s->P * flow_channels * sizeof(float));
        dmalloc(reinterpret_cast<void**>(&s->flowIn[i]), s->P * flow_channels * sizeof(float));
        hmalloc(reinterpret_cast<void**>(&s->flowPin[i]), s->P * flow_channels * sizeof(float));
        dmalloc(reinterpret_cast<void**>(&s->out_dev[i]), s->P * 4);
        hmalloc(reinterpret_cast<void**>(&s->out_pin[i]), s->P * 4);
        mkevent(&s->out_conv[i]);
        mkevent(&s->out_done[i]);
    }
    mkevent(&s->flow_ready);
    mkevent(&s->flow_free);
    if (!rc)
        rc = ensure_workspace(s, s->hp.pyramidLevels);
    if (rc) {
        vsc_stabilizer_destroy(s);
        return rc;
    }
    *out = s;
    return VSC_OK;
}

extern "C" void vsc_stabilizer_destroy(vsc_stabilizer* s)
{
    if (!s)
        return;
    if (s->compute) cudaStreamSynchronize(s->compute);
    if (s->copy) cudaStreamSynchronize(s->copy);
    if (s->d2h) cudaStreamSynchronize(s->d2h);
    for (int k = 0; k < kMaxSlots; ++k) {
        cudaFree(s->orig[k]);
        cudaFree(s->proc[k]);
        for (int i = 0; i < 2; ++i) {
            cudaFree(s->stage_dev[k][i]);
            cudaFreeHost(s->stage_pin[k][i]);
        }
        if (s->slot_ready[k]) cudaEventDestroy(s->slot_ready[k]);
        if (s->slot_h2d[k]) cudaEventDestroy(s->slot_h2d[k]);
        if (s->slot_released[k]) cudaEventDestroy(s->slot_released[k]);
    }
    cudaFree(s->lastStab);
    cudaFree(s->consisOut);
    for (int i = 0; i < 2; ++i) {
        cudaFree(s->flowUp[i]);
        cudaFree(s->flowIn[i]);
        cudaFreeHost(s->flowPin[i]);
        cudaFree(s->out_dev[i]);
        cudaFreeHost(s->out_pin[i]);
        if (s->out_conv[i]) cudaEventDestroy(s->out_conv[i]);
        if (s->out_done[i]) cudaEventDestroy(s->out_done[i]);
    }
    if (s->flow_ready) cudaEventDestroy(s->flow_ready);
    if (s->flow_free) cudaEventDestroy(s->flow_free);
    cudaFree(s->ws);
    if (s->pf_thread.joinable())
        s->pf_thread.join();
    for (int k = 0; k < 2; ++k) {
        cudaFreeHost(s->filePin[k][0]);
        cudaFreeHost(s->filePin[k][1]);
        if (s->file_h2d[k]) cudaEventDestroy(s->file_h2d[k]);
    }
    if (s->compute) cudaStreamDestroy(s->compute);
    if (s->copy) cudaStreamDestroy(s->copy);
    if (s->d2h) cudaStreamDestroy(s->d2h);
    (void)cudaGetLastError();
    delete s;
}

extern "C" vsc_hyper_params* vsc_stabilizer_hyper_params(vsc_stabilizer* s) { return s ? &s->hp : nullptr; }

extern "C" int vsc_stabilizer_push_frame(vsc_stabilizer* s, const uint8_t* orig_rgba_host,
    const uint8_t* proc_rgba_host)
{
    if (!s || !orig_rgba_host || !proc_rgba_host)
        return VSC_E_INVALID;
    if (s->count >= s->wcap)
        return VSC_E_STATE;  // the window is full: call vsc_stabilizer_step first
    const int k = (s->head + s->count) % s->nslots;
    const size_t bytes = s->P * 4;
    int rc;
    // the slot may still be read by an in-flight step (it was window[0] two steps ago)
    if (s->slot_used[k] && (rc = cu(cudaStreamWaitEvent(s->copy, s->slot_released[k], 0))))
        return rc;
    const uint8_t* src[2] = {orig_rgba_host, proc_rgba_host};
    float* dst[2] = {s->orig[k], s->proc[k]};
    for (int i = 0; i < 2; ++i) {
        const uint8_t* from = src[i];
        if (!is_pinned(from)) {
            if (s->slot_used[k])
                cudaEventSynchronize(s->slot_h2d[k]);  // previous H2D out of this staging buffer (4 pushes ago)
            std::memcpy(s->stage_pin[k][i], from, bytes);
            from = s->stage_pin[k][i];
        }
        if ((rc = cu(cudaMemcpyAsync(s->stage_dev[k][i], from, bytes, cudaMemcpyHostToDevice, s->copy))))
            return rc;
    }
    cudaEventRecord(s->slot_h2d[k], s->copy);
    for (int i = 0; i < 2; ++i)
        if ((rc = vsc_rgba8_to_f32x3(s->stage_dev[k][i], dst[i], s->W, s->H, s->copy)))
            return rc;
    s->pushed += 1;
    if (s->pushed == s->wcap) {
        // preloadProcessedFrames: lastStabilizedFrame <- processedFrames.back() (videostabilizer.cpp:152), the LAST of
        // the 2k + batchSize preloaded frames.  The
        // compute stream may still own lastStab only after a step, and none has run since create/reset.
        if ((rc = cu(cudaMemcpyAsync(s->lastStab, s->proc[k], s->n * sizeof(float), cudaMemcpyDeviceToDevice,
                 s->copy))))
            return rc;
    }
    cudaEventRecord(s->slot_ready[k], s->copy);
    s->slot_used[k] = true;
    s->count += 1;
    return cu(cudaGetLastError());
}

extern "C" int vsc_stabilizer_step(vsc_stabilizer* s, const float* flowFwd_dev, const float* flowBwd_dev,
    uint8_t* out_rgba_host)
{
    if (!s || !flowFwd_dev || !flowBwd_dev)
        return VSC_E_INVALID;
    return do_step(s, flowFwd_dev, flowBwd_dev, out_rgba_host);
}

extern "C" int vsc_stabilizer_step_lowres_flow(vsc_stabilizer* s, const float* flowFwd_dev, const float* flowBwd_dev,
    int flowW, int flowH, uint8_t* out_rgba_host)
{
    if (!s || !flowFwd_dev || !flowBwd_dev || flowW <= 0 || flowH <= 0)
        return VSC_E_INVALID;
    if (s->count < 3)
        return VSC_E_STATE;
    if (flowW == s->W && flowH == s->H)
        return do_step(s, flowFwd_dev, flowBwd_dev, out_rgba_host);
    // flowmodel.cpp:156-165: get_bilinear from the model's resolution, flow values not rescaled
    int rc = vsc_bilinear(flowFwd_dev, flowW, flowH, s->flowC, s->flowUp[0], s->W, s->H, s->flowC, s->compute);
    if (rc)
        return rc;
    rc = vsc_bilinear(flowBwd_dev, flowW, flowH, s->flowC, s->flowUp[1], s->W, s->H, s->flowC, s->compute);
    if (rc)
        return rc;
    return do_step(s, s->flowUp[0], s->flowUp[1], out_rgba_host);
}

extern "C" int vsc_stabilizer_step_host_flow(vsc_stabilizer* s, const float* flowFwd_host, const float* flowBwd_host,
    int flowW, int flowH, uint8_t* out_rgba_host)
{
    if (!s || !flowFwd_host || !flowBwd_host || flowW <= 0 || flowH <= 0 || flowW > s->W || flowH > s->H)
        return VSC_E_INVALID;
    if (s->count < 3)
        return VSC_E_STATE;
    const size_t bytes = static_cast<size_t>(flowW) * flowH * s->flowC * sizeof(float);
    int rc;
    // the landing buffers were read by the previous host-flow step on the compute stream
    if (s->flow_used && (rc = cu(cudaStreamWaitEvent(s->copy, s->flow_free, 0))))
        return rc;
    const float* src[2] = {flowFwd_host, flowBwd_host};
    for (int i = 0; i < 2; ++i) {
        const float* from = src[i];
        if (!is_pinned(from)) {
            if (s->flow_used)
                cudaEventSynchronize(s->flow_ready);  // previous H2D out of the pinned staging finished
            std::memcpy(s->flowPin[i], from, bytes);
            from = s->flowPin[i];
        }
        if ((rc = cu(cudaMemcpyAsync(s->flowIn[i], from, bytes, cudaMemcpyHostToDevice, s->copy))))
            return rc;
    }
    cudaEventRecord(s->flow_ready, s->copy);
    s->flow_used = true;
    if ((rc = cu(cudaStreamWaitEvent(s->compute, s->flow_ready, 0))))
        return rc;
    if (flowW == s->W && flowH == s->H) {
        rc = do_step(s, s->flowIn[0], s->flowIn[1], out_rgba_host);
    } else {
        rc = vsc_bilinear(s->flowIn[0], flowW, flowH, s->flowC, s->flowUp[0], s->W, s->H, s->flowC, s->compute);
        if (!rc)
            rc = vsc_bilinear(s->flowIn[1], flowW, flowH, s->flowC, s->flowUp[1], s->W, s->H, s->flowC, s->compute);
        if (!rc)
            rc = do_step(s, s->flowUp[0], s->flowUp[1], out_rgba_host);
    }
    cudaEventRecord(s->flow_free, s->compute);
    return rc;
}

namespace vsc {
int flo_read_into(const char* path, float* dst, size_t cap_floats, int* width, int* height);  // flo_io.cu
}

namespace {

// FileStabilizer::retrieveOpticalFlow (stabilizefiles.cpp:135-149) for frame `currentFrame`, into landing set k
int read_flow_pair(vsc_stabilizer* s, const std::string& dir, int currentFrame, int k)
{
    char path[2][4096];
    int rc;
    // "flow file i describes i-1 -> i; bwd flow file describes i -> i-1" (stabilizefiles.cpp:139-144)
    if ((rc = vsc_flo_frame_path(dir.c_str(), currentFrame + 1, 0, path[0], sizeof(path[0]))))
        return rc;
    if ((rc = vsc_flo_frame_path(dir.c_str(), currentFrame, 1, path[1], sizeof(path[1]))))
        return rc;
    int res[2] = {VSC_OK, VSC_OK};
    auto one = [&](int i) {
        int w = 0, h = 0;
        // a file of another size must not be read into the frame-sized buffer: check the header first
        if ((res[i] = vsc_flo_read_header(path[i], &w, &h)))
            return;
        if (w != s->W || h != s->H) {
            res[i] = VSC_E_FLO_DIMS;          // initializeFlowImage, imagehelpers.cpp:45-49
            return;
        }
        res[i] = vsc::flo_read_into(path[i], s->filePin[k][i], s->P * 2, &w, &h);
    };
    std::thread other(one, 1);
    one(0);
    other.join();
    return res[0] ? res[0] : res[1];
}

int ensure_file_buffers(vsc_stabilizer* s)
{
    if (s->filePin[0][0])
        return VSC_OK;
    int rc = VSC_OK;
    for (int k = 0; k < 2 && !rc; ++k) {
        for (int i = 0; i < 2 && !rc; ++i)
            rc = cu(cudaHostAlloc(reinterpret_cast<void**>(&s->filePin[k][i]), s->P * 2 * sizeof(float),
                cudaHostAllocDefault));
        if (!rc)
            rc = cu(cudaEventCreateWithFlags(&s->file_h2d[k], cudaEventDisableTiming));
    }
    return rc;
}

// set k is about to be overwritten by the host: its previous upload must have left it
void wait_file_set(vsc_stabilizer* s, int k)
{
    if (s->file_used[k])
        cudaEventSynchronize(s->file_h2d[k]);
}

void drop_prefetch(vsc_stabilizer* s)
{
    if (s->pf_thread.joinable())
        s->pf_thread.join();
    s->pf_active = false;
}

}  // namespace

extern "C" int vsc_stabilizer_prefetch_flow_files(vsc_stabilizer* s, const char* flow_dir, int currentFrame)
{
    if (!s || !flow_dir || s->flowC != 2)
        return VSC_E_INVALID;
    int rc = ensure_file_buffers(s);
    if (rc)
        return rc;
    if (s->pf_active && s->pf_frame == currentFrame && s->pf_dir == flow_dir)
        return VSC_OK;   // already on its way
    drop_prefetch(s);
    const int k = s->file_set;
    wait_file_set(s, k);
    s->pf_set = k;
    s->pf_frame = currentFrame;
    s->pf_dir = flow_dir;
    s->pf_rc = VSC_OK;
    s->pf_active = true;
    s->pf_thread = std::thread([s, k, currentFrame]() { s->pf_rc = read_flow_pair(s, s->pf_dir, currentFrame, k); });
    return VSC_OK;
}

// FileStabilizer::retrieveOpticalFlow + doOneStep: the two .flo files of frame `currentFrame` are read straight
// into pinned landing buffers (by the prefetch worker if vsc_stabilizer_prefetch_flow_files announced the
// frame, else here) and uploaded on the copy stream.
extern "C" int vsc_stabilizer_step_flow_files(vsc_stabilizer* s, const char* flow_dir, int currentFrame,
    uint8_t* out_rgba_host)
{
    if (!s || !flow_dir || s->flowC != 2)
        return VSC_E_INVALID;
    if (s->count < 3)
        return VSC_E_STATE;
    int rc = ensure_file_buffers(s);
    if (rc)
        return rc;
    int k;
    if (s->pf_active && s->pf_frame == currentFrame && s->pf_dir == flow_dir) {
        drop_prefetch(s);   // joins the worker
        k = s->pf_set;
        rc = s->pf_rc;
    } else {
        drop_prefetch(s);
        k = s->file_set;
        wait_file_set(s, k);
        rc = read_flow_pair(s, flow_dir, currentFrame, k);
    }
    if (rc)
        return rc;
    rc = vsc_stabilizer_step_host_flow(s, s->filePin[k][0], s->filePin[k][1], s->W, s->H, out_rgba_host);
    cudaEventRecord(s->file_h2d[k], s->copy);
    s->file_used[k] = true;
    s->file_set = k ^ 1;
    return rc;
}

extern "C" int vsc_stabilizer_flow_input(vsc_stabilizer* s, int window_index, uint8_t* dst_dev, int netW, int netH)
{
    if (!s || !dst_dev || window_index < 0 || window_index >= s->wcap || netW <= 0 || netH <= 0)
        return VSC_E_INVALID;
    if (window_index >= s->count)
        return VSC_E_STATE;
    const int k = (s->head + window_index) % s->nslots;
    int rc = cu(cudaStreamWaitEvent(s->compute, s->slot_ready[k], 0));
    if (rc)
        return rc;
    // stage_dev[k][0] keeps the slot's original RGBA8 frame until the slot is pushed again, which waits for the
    // step that releases it -- and that step is enqueued after this call on the same stream
    if (netW == s->W && netH == s->H)
        return cu(cudaMemcpyAsync(dst_dev, s->stage_dev[k][0], s->P * 4, cudaMemcpyDeviceToDevice, s->compute));
    return vsc_rgba8_scale_nearest(s->stage_dev[k][0], s->W, s->H, dst_dev, netW, netH, s->compute);
}

extern "C" int vsc_stabilizer_sync(vsc_stabilizer* s)
{
    if (!s)
        return VSC_E_INVALID;
    int rc = cu(cudaStreamSynchronize(s->copy));
    if (!rc) rc = cu(cudaStreamSynchronize(s->compute));
    if (!rc) rc = cu(cudaStreamSynchronize(s->d2h));
    flush_pending(s, 0);
    flush_pending(s, 1);
    return rc;
}

extern "C" int vsc_stabilizer_wait_uploads(vsc_stabilizer* s)
{
    if (!s)
        return VSC_E_INVALID;
    // every H2D goes through the copy stream; the events below are recorded right behind the copies
    int rc = VSC_OK;
    for (int k = 0; k < s->nslots && !rc; ++k)
        if (s->slot_used[k])
            rc = cu(cudaEventSynchronize(s->slot_h2d[k]));
    if (!rc && s->flow_used)
        rc = cu(cudaEventSynchronize(s->flow_ready));
    return rc;
}

extern "C" const float* vsc_stabilizer_last_output_dev(vsc_stabilizer* s) { return s ? s->lastStab : nullptr; }

extern "C" int vsc_stabilizer_copy_last_output(vsc_stabilizer* s, float* dst_dev)
{
    if (!s || !dst_dev)
        return VSC_E_INVALID;
    return cu(cudaMemcpyAsync(dst_dev, s->lastStab, s->n * sizeof(float), cudaMemcpyDeviceToDevice, s->compute));
}

extern "C" vsc_stream_t vsc_stabilizer_compute_stream(vsc_stabilizer* s) { return s ? s->compute : nullptr; }

extern "C" int vsc_stabilizer_reset(vsc_stabilizer* s)
{
    if (!s)
        return VSC_E_INVALID;
    drop_prefetch(s);
    const int rc = vsc_stabilizer_sync(s);
    s->head = 0;
    s->count = 0;
    s->pushed = 0;
    for (int k = 0; k < s->nslots; ++k)
        s->slot_used[k] = false;
    return rc;
}
