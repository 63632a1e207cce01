// Library-level entry points of the C ABI (include/vsc/vsc.h): version, error strings, launch
// counter, pinned host memory.
#include "vsc_common.cuh"

namespace vsc {

unsigned long long g_launches = 0;
std::atomic<bool> g_pdl = true;

int sm_count()
{
    static int cached[64] = {};
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) {
        (void)cudaGetLastError();
        return 148;  // B200; not cached so that a later call with a device present re-queries
    }
    int& slot = cached[dev & 63];
    if (slot == 0) {
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
            slot = n;
        else {
            (void)cudaGetLastError();
            return 148;
        }
    }
    return slot;
}

}  // namespace vsc

extern "C" int vsc_version(void) { return VSC_VERSION; }

extern "C" const char* vsc_error_string(int code)
{
    switch (code) {
        case VSC_OK: return "ok";
        case VSC_E_INVALID: return "invalid argument (null pointer, non-positive size or unsupported channel count)";
        case VSC_E_WORKSPACE: return "workspace missing, misaligned or too small";
        case VSC_E_STATE: return "stabilizer called out of order";
        case VSC_E_ALIGN: return "pointer not sufficiently aligned";
        // the reference's exception texts (flowIO.cpp:35-74, imagehelpers.cpp:48); callers append the file name
        case VSC_E_FLO_OPEN: return "ReadFlowFile: could not open";
        case VSC_E_FLO_HEADER: return "ReadFlowFile: problem reading file";
        case VSC_E_FLO_TAG: return "ReadFlowFile: wrong tag (possibly due to big-endian machine?)";
        case VSC_E_FLO_WIDTH: return "ReadFlowFile: illegal width";
        case VSC_E_FLO_HEIGHT: return "ReadFlowFile: illegal height";
        case VSC_E_FLO_SHORT: return "ReadFlowFile: file is too short";
        case VSC_E_FLO_LONG: return "ReadFlowFile: file is too long";
        case VSC_E_FLO_DIMS: return "Flow image size does not match image size";
        default: break;
    }
    if (code > 0)
        return cudaGetErrorString(static_cast<cudaError_t>(code));
    return "unknown vsc error";
}

extern "C" uint64_t vsc_launch_count(void)
{
    return static_cast<uint64_t>(__atomic_load_n(&vsc::g_launches, __ATOMIC_RELAXED));
}

extern "C" int vsc_host_alloc(void** p, size_t bytes)
{
    if (!p || bytes == 0)
        return VSC_E_INVALID;
    const cudaError_t e = cudaHostAlloc(p, bytes, cudaHostAllocDefault);
    return e == cudaSuccess ? VSC_OK : static_cast<int>(e);
}

extern "C" int vsc_host_free(void* p)
{
    if (!p)
        return VSC_OK;
    const cudaError_t e = cudaFreeHost(p);
    return e == cudaSuccess ? VSC_OK : static_cast<int>(e);
}
