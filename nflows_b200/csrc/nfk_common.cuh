// Shared host-side helpers of libnfk_sm100.so: error reporting, launch accounting.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

#include "../../include/nfk.h"

namespace nfk {

extern thread_local char g_last_error[512];
extern std::atomic<int64_t> g_launch_count;

int fail(int code, const char* fmt, ...);

inline int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(NFK_E_CUDA, "%s: %s", what, cudaGetErrorString(e));
    g_launch_count.fetch_add(1, std::memory_order_relaxed);
    return NFK_OK;
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace nfk

#define NFK_REQUIRE(cond, ...)                                   \
    do {                                                         \
        if (!(cond)) return nfk::fail(NFK_E_INVALID, __VA_ARGS__); \
    } while (0)
