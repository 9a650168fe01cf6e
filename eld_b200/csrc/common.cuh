// common.cuh - shared host-side plumbing of libeld_b200.so (error channel, context).
#pragma once
#include <cstdlib>
#include <cuda_runtime.h>
#include <cuda.h>
#include <cstdint>
#include <cstdio>
#include <cstdarg>
#include <atomic>

#include "../../include/eld_b200.h"

namespace eld {

void set_error(const char* fmt, ...);

#define ELD_CHECK_CUDA(expr)                                                                   \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) {                                                               \
            ::eld::set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #expr,                \
                             cudaGetErrorString(_e));                                          \
            return ELD_E_CUDA;                                                                 \
        }                                                                                      \
    } while (0)

#define ELD_REQUIRE(cond, ...)                                                                 \
    do {                                                                                       \
        if (!(cond)) {                                                                         \
            ::eld::set_error(__VA_ARGS__);                                                     \
            return ELD_E_ARG;                                                                  \
        }                                                                                      \
    } while (0)

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                    const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

}  // namespace eld

struct eld_ctx {
    int device;
    int num_sms;
    int smem_optin;
    eld::PFN_encodeTiled encode_tiled;
    std::atomic<int64_t> launches;
};

namespace eld {
inline void count_launch(eld_ctx* ctx, int n = 1) { ctx->launches.fetch_add(n, std::memory_order_relaxed); }

// Launch with programmatic stream serialization (PDL): the kernel may start while its predecessor drains; it blocks in
// griddepcontrol.wait before touching anything the predecessor writes.  ELD_NO_PDL=1 restores plain launches.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), int grid, int block, size_t smem, cudaStream_t st, Args... args)
{
    static const bool no_pdl = getenv("ELD_NO_PDL") != nullptr;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3((unsigned)block); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = no_pdl ? 0 : 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}


// device side of PDL (see umma.cuh for the tcgen05 tiles): wait for the predecessor, then release the successor
__device__ __forceinline__ void pdl_wait_and_release()
{
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
}
