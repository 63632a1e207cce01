// TEST DRIVER for the product's ORT custom-op shim (video-stream-consistency_b200/host/ort_custom_ops):
// plays the part of an onnxruntime session with the stand-in API (standins/ort): calls the exported
// RegisterCustomOps, looks the op up in the registered domain by name and execution provider, creates the
// kernel from node attributes and runs KernelCompute on caller-supplied (device) buffers.
#include <onnxruntime_cxx_api.h>

#include <cstdio>
#include <cstring>
#include <exception>
#include <string>

extern "C" OrtStatus* RegisterCustomOps(OrtSessionOptions* options, const OrtApiBase* api);

namespace {

std::string g_last_error;

struct OutSlot {
    void* ptr;
    size_t bytes;
    size_t asked;
};
void* take_output(void* user, size_t, size_t bytes)
{
    OutSlot* s = static_cast<OutSlot*>(user);
    s->asked = bytes;
    if (bytes > s->bytes)
        throw std::runtime_error("test driver: output buffer too small");
    return s->ptr;
}

OrtSessionOptions& session()
{
    static OrtSessionOptions opts;
    static bool done = false;
    if (!done) {
        done = true;
        OrtStatus* st = RegisterCustomOps(&opts, OrtGetApiBase());
        if (st) {
            g_last_error = st->msg;
            delete st;
        }
    }
    return opts;
}

const OrtCustomOp* find_op(const char* name, const char* provider)
{
    for (OrtCustomOpDomain* d : session().domains)
        if (d->name == "custom")
            for (const OrtCustomOp* op : d->ops)
                if (!std::strcmp(op->GetName(op), name) && !std::strcmp(op->GetExecutionProviderType(op), provider))
                    return op;
    return nullptr;
}

int run(const OrtCustomOp* op, const OrtKernelInfo& info, OrtKernelContext& ctx)
{
    if (!op) {
        g_last_error = "op not registered";
        return 2;
    }
    void* k = nullptr;
    try {
        k = op->CreateKernel(op, OrtGetApiBase()->GetApi(ORT_API_VERSION), &info);
        op->KernelCompute(k, &ctx);
        op->KernelDestroy(k);
        return 0;
    } catch (const std::exception& e) {
        g_last_error = e.what();
        if (k)
            op->KernelDestroy(k);
        return 1;
    }
}

}  // namespace

extern "C" {

const char* vsc_ort_test_last_error() { return g_last_error.c_str(); }

// "domain|op|provider|nin|nout|intype0|outtype0;" per registered op; returns the number of ops
int vsc_ort_test_registry(char* buf, size_t n)
{
    std::string s;
    int count = 0;
    for (OrtCustomOpDomain* d : session().domains)
        for (const OrtCustomOp* op : d->ops) {
            char line[256];
            std::snprintf(line, sizeof line, "%s|%s|%s|%zu|%zu|%d|%d;", d->name.c_str(), op->GetName(op),
                op->GetExecutionProviderType(op), op->GetInputTypeCount(op), op->GetOutputTypeCount(op),
                static_cast<int>(op->GetInputType(op, 0)), static_cast<int>(op->GetOutputType(op, 0)));
            s += line;
            ++count;
        }
    std::snprintf(buf, n, "%s", s.c_str());
    return count;
}

int vsc_ort_test_correlation(const float* in1, const float* in2, float* out, size_t out_bytes, int64_t N, int64_t C,
    int64_t H, int64_t W, int64_t md, int64_t legacy, int has_legacy, int has_md, void* stream, int64_t* out_dims,
    int* out_rank)
{
    OrtKernelInfo info;
    info.legacy = legacy;
    info.max_displacement = md;
    info.has_legacy = has_legacy != 0;
    info.has_max_displacement = has_md != 0;
    OrtKernelContext ctx;
    OutSlot slot{out, out_bytes, 0};
    ctx.alloc_output = take_output;
    ctx.alloc_user = &slot;
    ctx.gpu_stream = stream;
    ctx.inputs.resize(2);
    ctx.inputs[0].shape = {N, C, H, W};
    ctx.inputs[0].data = const_cast<float*>(in1);
    ctx.inputs[1].shape = {N, C, H, W};
    ctx.inputs[1].data = const_cast<float*>(in2);
    const int rc = run(find_op("Correlation", "CUDAExecutionProvider"), info, ctx);
    if (rc == 0 && out_rank) {
        *out_rank = static_cast<int>(ctx.outputs[0].shape.size());
        for (size_t i = 0; i < ctx.outputs[0].shape.size(); ++i)
            out_dims[i] = ctx.outputs[0].shape[i];
    }
    return rc;
}

int vsc_ort_test_warp(const float* in, const float* flow, float* out, size_t out_bytes, int64_t N, int64_t C, int64_t H,
    int64_t W, int64_t flow_channels, void* stream)
{
    OrtKernelInfo info;
    OrtKernelContext ctx;
    OutSlot slot{out, out_bytes, 0};
    ctx.alloc_output = take_output;
    ctx.alloc_user = &slot;
    ctx.gpu_stream = stream;
    ctx.inputs.resize(2);
    ctx.inputs[0].shape = {N, C, H, W};
    ctx.inputs[0].data = const_cast<float*>(in);
    ctx.inputs[1].shape = {N, flow_channels, H, W};
    ctx.inputs[1].data = const_cast<float*>(flow);
    return run(find_op("Warp", "CUDAExecutionProvider"), info, ctx);
}
}
