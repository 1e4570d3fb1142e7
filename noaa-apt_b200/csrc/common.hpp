// Shared host-side helpers: status/last-error plumbing and CUDA call checking.
#pragma once

#include <cstdarg>
#include <cstdio>
#include <string>

#include <cuda_runtime.h>

#include "aptb200.h"

namespace aptb200 {

std::string &last_error_slot();                 // thread-local message buffer
int fail(int status, const char *fmt, ...);     // records the message, returns status

inline int cuda_fail(cudaError_t e, const char *what, const char *file, int line) {
    return fail(APT_ERR_CUDA, "%s failed at %s:%d: %s", what, file, line, cudaGetErrorString(e));
}

#define APT_CUDA(call)                                                                    \
    do {                                                                                  \
        cudaError_t e__ = (call);                                                         \
        if (e__ != cudaSuccess) {                                                         \
            int st__ = e__ == cudaErrorMemoryAllocation ? APT_ERR_NOMEM : APT_ERR_CUDA;   \
            ::aptb200::cuda_fail(e__, #call, __FILE__, __LINE__);                         \
            return st__;                                                                  \
        }                                                                                 \
    } while (0)

#define APT_TRY(expr)                      \
    do {                                   \
        int st__ = (expr);                 \
        if (st__ != APT_OK) return st__;   \
    } while (0)

}  // namespace aptb200
