// common.cuh - shared host-side plumbing of libeld_b200.so (error channel, context).
#pragma once
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
}
