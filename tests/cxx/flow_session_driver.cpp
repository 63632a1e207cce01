// TEST DRIVER for the product's flow session (video-stream-consistency_b200/host/inference/vsc_flow_session.cpp):
// plays onnxruntime + the PWC-Net graph with the stand-in API (standins/ort).  The "model" is a graph function
// registered under a path: it checks that the custom-op domain was registered on the session options, and computes
// a flow-shaped output from BOTH bound inputs with kernels of the C ABI on the provider's compute stream:
//     output[1,H,W,3] = bilinear sample of float3(frame1) displaced by float3(frame2).rg      (vsc_warp_hwc3)
// so a test can reproduce it exactly and see stale inputs, swapped bindings or missing stream ordering.
#include <cuda_runtime_api.h>
#include <onnxruntime_cxx_api.h>

#include <cstring>
#include <exception>
#include <memory>
#include <string>

#include "../../video-stream-consistency_b200/host/inference/vsc_flow_session.h"

namespace {

std::string g_error;
const char* const kModel = "models/standin-flow-wpreproc.onnx";

struct Graph {
    float *a = nullptr, *b = nullptr;   // scratch: float3 images of the two inputs
    size_t px = 0;
    bool saw_custom_domain = false;
    long calls = 0;
};
Graph g_graph;

void graph_fn(OrtStandinRun& r)
{
    Graph& g = g_graph;
    ++g.calls;
    for (OrtCustomOpDomain* d : r.domains)
        if (d->name == "custom" && d->ops.size() >= 2)
            g.saw_custom_domain = true;
    const OrtMockTensor& f1 = r.inputs.at("frame1");
    const OrtMockTensor& f2 = r.inputs.at("frame2");
    OrtMockTensor& out = r.outputs.at("output");
    if (f1.type != ONNX_TENSOR_ELEMENT_DATA_TYPE_UINT8 || f1.shape.size() != 4 || f1.shape[3] != 4
        || out.type != ONNX_TENSOR_ELEMENT_DATA_TYPE_FLOAT || out.shape.size() != 4 || out.shape[3] != 3
        || f1.shape != f2.shape || out.shape[1] != f1.shape[1] || out.shape[2] != f1.shape[2])
        throw Ort::Exception("stand-in graph: unexpected tensor types / shapes");
    const int B = static_cast<int>(f1.shape[0]), H = static_cast<int>(f1.shape[1]), W = static_cast<int>(f1.shape[2]);
    if (static_cast<size_t>(W) * H > g.px || out.shape[0] != f1.shape[0])
        throw Ort::Exception("stand-in graph: scratch too small / batch mismatch");
    int rc = 0;
    for (int n = 0; n < B && !rc; ++n) {   // per sample, like the real graph's batch dimension
        const size_t px = static_cast<size_t>(W) * H;
        rc = vsc_rgba8_to_f32x3(static_cast<const uint8_t*>(f1.data) + n * px * 4, g.a, W, H, r.stream);
        if (!rc) rc = vsc_rgba8_to_f32x3(static_cast<const uint8_t*>(f2.data) + n * px * 4, g.b, W, H, r.stream);
        if (!rc) rc = vsc_warp_hwc3(g.a, g.b, static_cast<float*>(out.data) + n * px * 3, W, H, 3, r.stream);
    }
    if (rc)
        throw Ort::Exception(std::string("stand-in graph: ") + vsc_error_string(rc));
}

struct Rig {
    vsc_stabilizer* st = nullptr;
    Ort::Env env;
    std::unique_ptr<VscFlowSession> fs;
    int W = 0, H = 0, netW = 0, netH = 0;
};

}  // namespace

extern "C" {

const char* vsc_fs_test_last_error() { return g_error.c_str(); }

void* vsc_fs_test_create_b(int W, int H, int netW, int netH, const char* model_path, int batch_directions, int batch_size);
void* vsc_fs_test_create(int W, int H, int netW, int netH, const char* model_path, int batch_directions)
{
    return vsc_fs_test_create_b(W, H, netW, netH, model_path, batch_directions, 1);
}

// batch_size = the CLI's -b: window of 2 + batch_size frames, [batch_size, H, W, 4] session tensors
void* vsc_fs_test_create_b(int W, int H, int netW, int netH, const char* model_path, int batch_directions, int batch_size)
{
    try {
        OrtStandinGraphs()[kModel] = graph_fn;
        OrtStandinState().synchronize = [](void* s) { cudaStreamSynchronize(static_cast<cudaStream_t>(s)); };
        auto rig = std::make_unique<Rig>();
        rig->W = W; rig->H = H; rig->netW = netW; rig->netH = netH;
        if (static_cast<size_t>(netW) * netH > g_graph.px) {
            cudaFree(g_graph.a);
            cudaFree(g_graph.b);
            g_graph.px = static_cast<size_t>(netW) * netH;
            if (cudaMalloc(reinterpret_cast<void**>(&g_graph.a), g_graph.px * 12) != cudaSuccess
                || cudaMalloc(reinterpret_cast<void**>(&g_graph.b), g_graph.px * 12) != cudaSuccess)
                throw std::runtime_error("driver: cudaMalloc failed");
        }
        const int rc = vsc_stabilizer_create_batched(&rig->st, W, H, 3, batch_size);
        if (rc)
            throw std::runtime_error(vsc_error_string(rc));
        rig->fs = std::make_unique<VscFlowSession>(rig->env, model_path ? model_path : kModel, netW, netH, rig->st, 0,
            batch_directions != 0, batch_size);
        return rig.release();
    } catch (const std::exception& e) {
        g_error = e.what();
        return nullptr;
    }
}

void vsc_fs_test_destroy(void* h)
{
    Rig* rig = static_cast<Rig*>(h);
    if (!rig)
        return;
    rig->fs.reset();
    vsc_stabilizer_destroy(rig->st);
    delete rig;
}

vsc_stabilizer* vsc_fs_test_stabilizer(void* h) { return static_cast<Rig*>(h)->st; }

// retrieveOpticalFlow + doOneStep for the window's current frame; out_rgba_host valid after vsc_stabilizer_sync
int vsc_fs_test_step(void* h, uint8_t* out_rgba_host)
{
    try {
        static_cast<Rig*>(h)->fs->stabilizeCurrentFrame(out_rgba_host);
        return 0;
    } catch (const std::exception& e) {
        g_error = e.what();
        return 1;
    }
}

// one direction, result copied to the host (synchronises)
int vsc_fs_test_flow(void* h, int indexFirst, int indexSecond, int slot, float* dst_host)
{
    Rig* rig = static_cast<Rig*>(h);
    try {
        const float* d = rig->fs->run(indexFirst, indexSecond, slot);
        vsc_stabilizer_sync(rig->st);
        return static_cast<int>(cudaMemcpy(dst_host, d, static_cast<size_t>(rig->netW) * rig->netH * 12,
            cudaMemcpyDeviceToHost));
    } catch (const std::exception& e) {
        g_error = e.what();
        return 1;
    }
}

// runs, provider_syncs, sessions_created, tensors_bound, graph calls, saw the custom-op domain
void vsc_fs_test_counters(long* out6)
{
    const OrtStandinCounters& c = OrtStandinState();
    out6[0] = c.runs;
    out6[1] = c.provider_syncs;
    out6[2] = c.sessions_created;
    out6[3] = c.tensors_bound;
    out6[4] = g_graph.calls;
    out6[5] = g_graph.saw_custom_domain ? 1 : 0;
}
}
