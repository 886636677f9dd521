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

// Latch for one-time PER-DEVICE setup (cudaFuncSetAttribute and the like): a process may run flows on several GPUs.
struct DeviceOnce {
    std::atomic<uint64_t> done{0};
    bool pending(int* dev) {
        int d = 0;
        cudaGetDevice(&d);
        *dev = d;
        return ((done.load(std::memory_order_acquire) >> (d & 63)) & 1ull) == 0;
    }
    void mark(int dev) { done.fetch_or(1ull << (dev & 63), std::memory_order_release); }
};

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace nfk

#define NFK_REQUIRE(cond, ...)                                   \
    do {                                                         \
        if (!(cond)) return nfk::fail(NFK_E_INVALID, __VA_ARGS__); \
    } while (0)
