// STAND-IN for the session half of <onnxruntime_cxx_api.h> (Env, SessionOptions, CUDA provider options, MemoryInfo,
// Value, IoBinding, RunOptions, Session): the subset the product's flow session
// (video-stream-consistency_b200/host/inference/vsc_flow_session.cpp) uses, spelled like onnxruntime 1.20.1 so that
// the same source builds against the real headers.  Ours, not a copy of ORT.
//
// Behind it is an in-memory "runtime": a model path names a graph function the test driver registered
// (OrtStandinGraphs()); Session::Run(RunOptions, IoBinding) hands that function the bound tensors, the session's
// custom-op domains and the CUDA provider's compute stream, and -- like the real CUDA provider -- synchronises
// that stream afterwards unless the run was configured with disable_synchronize_execution_providers = 1.  The
// driver supplies the synchronise function (the stand-in has no CUDA dependency) and can read the counters to
// check that a run really was enqueue-only.
#pragma once
#include "onnxruntime_c_api.h"

#include <functional>
#include <map>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

enum OrtAllocatorType { OrtInvalidAllocator = -1, OrtDeviceAllocator = 0, OrtArenaAllocator = 1 };
enum OrtMemType { OrtMemTypeCPUInput = -2, OrtMemTypeCPUOutput = -1, OrtMemTypeCPU = -1, OrtMemTypeDefault = 0 };
enum GraphOptimizationLevel { ORT_DISABLE_ALL = 0, ORT_ENABLE_BASIC = 1, ORT_ENABLE_EXTENDED = 2, ORT_ENABLE_ALL = 99 };
enum OrtLoggingLevel { ORT_LOGGING_LEVEL_VERBOSE, ORT_LOGGING_LEVEL_INFO, ORT_LOGGING_LEVEL_WARNING,
    ORT_LOGGING_LEVEL_ERROR, ORT_LOGGING_LEVEL_FATAL };

// the legacy CUDA provider options struct (onnxruntime_c_api.h); only the fields the product sets
struct OrtCUDAProviderOptions {
    int device_id = 0;
    int has_user_compute_stream = 0;
    void* user_compute_stream = nullptr;
    int do_copy_in_default_stream = 1;
};

// what a registered graph function sees of one Run
struct OrtStandinRun {
    std::map<std::string, OrtMockTensor> inputs, outputs;   // bound tensors by graph name
    std::vector<OrtCustomOpDomain*> domains;                // custom-op domains registered on the session options
    void* stream = nullptr;                                 // the CUDA provider's compute stream
};
using OrtStandinGraph = std::function<void(OrtStandinRun&)>;

inline std::map<std::string, OrtStandinGraph>& OrtStandinGraphs()
{
    static std::map<std::string, OrtStandinGraph> g;
    return g;
}
struct OrtStandinCounters {
    long runs = 0, provider_syncs = 0, sessions_created = 0, tensors_bound = 0;
    void (*synchronize)(void* stream) = nullptr;   // driver-supplied cudaStreamSynchronize
};
inline OrtStandinCounters& OrtStandinState()
{
    static OrtStandinCounters c;
    return c;
}

namespace Ort {

struct Exception : std::runtime_error {
    using std::runtime_error::runtime_error;
};

struct Env {
    explicit Env(OrtLoggingLevel = ORT_LOGGING_LEVEL_WARNING, const char* = "") { }
};

struct SessionOptions {
    SessionOptions& SetIntraOpNumThreads(int) { return *this; }
    SessionOptions& SetGraphOptimizationLevel(GraphOptimizationLevel) { return *this; }
    SessionOptions& AppendExecutionProvider_CUDA(const OrtCUDAProviderOptions& o)
    {
        cuda = o;
        has_cuda = true;
        return *this;
    }
    operator OrtSessionOptions*() { return &c; }
    OrtSessionOptions c;
    OrtCUDAProviderOptions cuda;
    bool has_cuda = false;
};

struct MemoryInfo {
    MemoryInfo(const char* name_, OrtAllocatorType, int device_, OrtMemType) : name(name_), device(device_) { }
    std::string name;
    int device;
};

// a tensor over caller-owned memory (Ort::Value::CreateTensor(info, data, bytes, shape, rank, type))
struct Value {
    static Value CreateTensor(const MemoryInfo& info, void* data, size_t bytes, const int64_t* shape, size_t rank,
        ONNXTensorElementDataType type)
    {
        Value v;
        v.t.type = type;
        v.t.shape.assign(shape, shape + rank);
        v.t.data = data;
        v.bytes = bytes;
        v.on_device = info.name == "Cuda";
        return v;
    }
    OrtMockTensor t;
    size_t bytes = 0;
    bool on_device = false;
};

struct Session;

struct IoBinding {
    explicit IoBinding(Session&) { }
    void BindInput(const char* name, const Value& v)
    {
        run.inputs[name] = v.t;
        ++OrtStandinState().tensors_bound;
    }
    void BindOutput(const char* name, const Value& v)
    {
        run.outputs[name] = v.t;
        ++OrtStandinState().tensors_bound;
    }
    void ClearBoundInputs() { run.inputs.clear(); }
    void ClearBoundOutputs() { run.outputs.clear(); }
    OrtStandinRun run;
};

struct RunOptions {
    RunOptions& AddConfigEntry(const char* key, const char* value)
    {
        config[key] = value;
        return *this;
    }
    std::map<std::string, std::string> config;
};

struct Session {
    Session(Env&, const char* model_path, SessionOptions& options) : path(model_path)
    {
        const auto it = OrtStandinGraphs().find(path);
        if (it == OrtStandinGraphs().end())
            throw Exception("Load model from " + path + " failed: File doesn't exist");
        graph = it->second;
        domains = options.c.domains;
        stream = options.has_cuda && options.cuda.has_user_compute_stream ? options.cuda.user_compute_stream : nullptr;
        ++OrtStandinState().sessions_created;
    }
    void Run(const RunOptions& ro, IoBinding& binding)
    {
        OrtStandinRun& r = binding.run;
        r.domains = domains;
        r.stream = stream;
        graph(r);
        ++OrtStandinState().runs;
        const auto it = ro.config.find("disable_synchronize_execution_providers");
        if (it == ro.config.end() || it->second != "1") {
            ++OrtStandinState().provider_syncs;
            if (OrtStandinState().synchronize)
                OrtStandinState().synchronize(stream);
        }
    }
    std::string path;
    OrtStandinGraph graph;
    std::vector<OrtCustomOpDomain*> domains;
    void* stream = nullptr;
};

}  // namespace Ort
