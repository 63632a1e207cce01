// TEST INFRASTRUCTURE -- not product code.
//
// Driver that runs the reference's OWN CPU custom-op kernels
// (/root/reference/src/ort_custom_ops/src/opticalflow/{correlation,warp}.cc, compiled
// unmodified against the stand-in ORT headers in standins/ort) on plain host
// buffers.  It goes through CorrelationKernel::Compute / WarpKernel::Compute
// (correlation.h:33-49, warp.h:18-31), i.e. through the same shape logic and
// attribute handling ORT would exercise.  Built by oracle/Makefile into
// oracle/_ref/libvsc_ref_cpu.so; used by tests/ (golden-vector generation, oracle
// pinning) and by bench.py's cpu_baseline / --impl reference legs only.
#include <ort_custom_ops/opticalflow/correlation.h>
#include <ort_custom_ops/opticalflow/warp.h>
#include <ort_custom_ops/custom_ops.h>

#include "flowIO.h"  // the reference's Middlebury .flo reader/writer (src/stabilization/flowIO.cpp), unmodified

#include <cstdio>
#include <cstring>
#include <exception>
#include <string>

namespace {

struct OutSlot {
    void* ptr;
    size_t bytes;
};

void* take_output(void* user, size_t /*index*/, size_t bytes)
{
    OutSlot* s = static_cast<OutSlot*>(user);
    if (bytes > s->bytes)
        throw std::runtime_error("ref driver: output buffer too small");
    s->bytes = bytes;
    return s->ptr;
}

const OrtApi g_api{};

}  // namespace

extern "C" {

// returns 0 on success; 1 = exception (message printed); out_rank/out_dims report what the op asked ORT for
int vsc_ref_cpu_correlation(const float* in1, const float* in2, float* out, size_t out_bytes, int64_t N, int64_t C,
    int64_t H, int64_t W, int64_t max_displacement, int64_t legacy, int64_t* out_dims, int* out_rank)
{
    try {
        OrtKernelInfo info;
        info.legacy = legacy;
        info.max_displacement = max_displacement;
        CorrelationKernel k(g_api, &info, "CPUExecutionProvider");
        OrtKernelContext ctx;
        OutSlot slot{out, out_bytes};
        ctx.alloc_output = take_output;
        ctx.alloc_user = &slot;
        ctx.inputs.resize(2);
        ctx.inputs[0].shape = {N, C, H, W};
        ctx.inputs[0].data = const_cast<float*>(in1);
        ctx.inputs[1].shape = {N, C, H, W};
        ctx.inputs[1].data = const_cast<float*>(in2);
        k.Compute(&ctx);
        if (out_rank)
            *out_rank = static_cast<int>(ctx.outputs[0].shape.size());
        if (out_dims)
            for (size_t i = 0; i < ctx.outputs[0].shape.size(); ++i)
                out_dims[i] = ctx.outputs[0].shape[i];
        return 0;
    } catch (const std::exception& e) {
        std::fprintf(stderr, "vsc_ref_cpu_correlation: %s\n", e.what());
        return 1;
    }
}

int vsc_ref_cpu_warp(const float* in, const float* flow, float* out, size_t out_bytes, int64_t N, int64_t C, int64_t H,
    int64_t W)
{
    try {
        OrtKernelInfo info;
        WarpKernel k(g_api, &info, "CPUExecutionProvider");
        OrtKernelContext ctx;
        OutSlot slot{out, out_bytes};
        ctx.alloc_output = take_output;
        ctx.alloc_user = &slot;
        ctx.inputs.resize(2);
        ctx.inputs[0].shape = {N, C, H, W};
        ctx.inputs[0].data = const_cast<float*>(in);
        ctx.inputs[1].shape = {N, 2, H, W};
        ctx.inputs[1].data = const_cast<float*>(flow);
        k.Compute(&ctx);
        return 0;
    } catch (const std::exception& e) {
        std::fprintf(stderr, "vsc_ref_cpu_warp: %s\n", e.what());
        return 1;
    }
}

// the reference's own registration (custom_ops.cpp:73-97) against the stand-in session:
// "domain|op|provider|nin|nout|intype0|outtype0;" per registered op; returns the number of ops, -1 on error
int vsc_ref_cpu_registry(char* buf, size_t n)
{
    static OrtSessionOptions opts;
    static bool done = false;
    if (!done) {
        done = true;
        Ort::InitApi(OrtGetApiBase()->GetApi(ORT_API_VERSION));  // the host's auto-initialised copy, SURVEY 8b pitfall
        if (OrtStatus* st = RegisterCustomOps(&opts, OrtGetApiBase())) {
            std::fprintf(stderr, "reference RegisterCustomOps: %s\n", st->msg.c_str());
            delete st;
            return -1;
        }
    }
    std::string s;
    int count = 0;
    for (OrtCustomOpDomain* d : opts.domains)
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

// ReadFlowFile / WriteFlowFile of the reference (flowIO.cpp:31-107) on caller buffers; 0 ok, 1 = it threw
int vsc_ref_read_flo(const char* path, float* buf, size_t cap_floats, int* w, int* h)
{
    try {
        std::vector<float> flow;
        ReadFlowFile(flow, *w, *h, path);
        if (flow.size() > cap_floats)
            return 2;
        std::memcpy(buf, flow.data(), flow.size() * sizeof(float));
        return 0;
    } catch (const std::exception& e) {
        std::fprintf(stderr, "vsc_ref_read_flo: %s\n", e.what());
        return 1;
    }
}

// same, returning the reference's exception text
int vsc_ref_read_flo_msg(const char* path, float* buf, size_t cap_floats, int* w, int* h, char* msg, size_t msg_cap)
{
    try {
        std::vector<float> flow;
        ReadFlowFile(flow, *w, *h, path);
        if (flow.size() > cap_floats)
            return 2;
        std::memcpy(buf, flow.data(), flow.size() * sizeof(float));
        return 0;
    } catch (const std::exception& e) {
        if (msg && msg_cap)
            std::snprintf(msg, msg_cap, "%s", e.what());
        return 1;
    }
}

int vsc_ref_write_flo(const char* path, const float* buf, int w, int h)
{
    try {
        std::vector<float> flow(buf, buf + static_cast<size_t>(w) * h * 2);
        WriteFlowFile(flow, w, h, path);
        return 0;
    } catch (const std::exception& e) {
        std::fprintf(stderr, "vsc_ref_write_flo: %s\n", e.what());
        return 1;
    }
}

// missing-attribute behaviour of the reference ctor (correlation.h:19-31): returns 1 if it threw
int vsc_ref_cpu_correlation_ctor_throws(int has_legacy, int has_max_displacement)
{
    try {
        OrtKernelInfo info;
        info.has_legacy = has_legacy != 0;
        info.has_max_displacement = has_max_displacement != 0;
        CorrelationKernel k(g_api, &info, "CPUExecutionProvider");
        return 0;
    } catch (const std::exception&) {
        return 1;
    }
}
}
