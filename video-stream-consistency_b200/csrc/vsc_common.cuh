// Shared helpers of the sm_100a kernels (internal; the public surface is include/vsc/vsc.h).
//
// Floating-point policy: the library is compiled with -fmad=false, so every expression is
// evaluated exactly as written (IEEE fp32, round-to-nearest), which makes the small kernels
// bit-identical to the plain-C restatement of the reference.  Fused multiply-adds appear only
// where they are written explicitly (__fmaf_rn) in the two compute-heavy loops: the Correlation
// contraction and the solver sweep.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

#include "vsc/vsc.h"

namespace vsc {

extern unsigned long long g_launches;  // host-side counter, see vsc_launch_count()

inline void count_launch(unsigned n = 1) { __atomic_fetch_add(&g_launches, (unsigned long long)n, __ATOMIC_RELAXED); }

inline int launch_status()
{
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? VSC_OK : static_cast<int>(e);
}

inline cudaStream_t as_stream(vsc_stream_t s) { return static_cast<cudaStream_t>(s); }

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
inline bool aligned4(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 3u) == 0; }

inline unsigned cdiv(long long a, long long b) { return static_cast<unsigned>((a + b - 1) / b); }

// number of SMs of the current device (cached); B200 = 148
int sm_count();

// Opt a kernel into `smem` bytes of dynamic shared memory (and optionally the max-shared carve-out).  Function
// attributes are per device, so the "already done" flag is a bit per device ordinal, not one per process.
template <class Kernel>
inline int ensure_dynamic_smem(Kernel kernel, size_t smem, bool max_carveout, unsigned long long& done_mask)
{
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess)
        return static_cast<int>(cudaGetLastError());
    const unsigned long long bit = 1ull << (dev & 63);
    if (__atomic_load_n(&done_mask, __ATOMIC_ACQUIRE) & bit)
        return VSC_OK;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e == cudaSuccess && max_carveout)
        e = cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess)
        return static_cast<int>(e);
    __atomic_fetch_or(&done_mask, bit, __ATOMIC_RELEASE);
    return VSC_OK;
}

// Programmatic dependent launch (sm_90+): a kernel launched with launch_pdl() may start while its predecessor in
// the stream is still draining; it must call pdl_wait() before it touches anything the predecessor wrote (the
// wait returns once the predecessor grid has completed and flushed), and it calls pdl_trigger() first so that ITS
// successor can be placed early as well.  Without the launch attribute both are no-ops, so the kernels can also be
// launched the ordinary way.  The per-frame chain is ~45 dependent launches: this hides their launch latencies.
extern std::atomic<bool> g_pdl;   // vsc_set_solver_mode(| 0x80) turns it off (A/B runs)
__device__ __forceinline__ void pdl_enter()
{
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
}
template <class... KArgs, class... Args>
inline int launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = g_pdl ? 1 : 0;
    const cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
    return e == cudaSuccess ? VSC_OK : static_cast<int>(e);
}

// streaming (read-once) loads/stores: do not allocate in L1
__device__ __forceinline__ float4 ldg_stream4(const float* p)
{
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p));
    return v;
}
__device__ __forceinline__ float ldg_stream(const float* p)
{
    float v;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ void stg_stream4(float* p, float4 v)
{
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
                 "f"(v.w)
                 : "memory");
}

}  // namespace vsc
