// common.cuh -- shared host-side plumbing of libpanib200.so (error text, launch accounting).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

#include "../../include/panib200.h"

namespace panib {

constexpr uint64_t kEmpty = 0xFFFFFFFFFFFFFFFFull;  // empty slot of a sketch-table bucket
constexpr int kBucketSlots = PANIB_BUCKET_SLOTS;

void set_error(const char *fmt, ...);
extern std::atomic<uint64_t> g_launches;

inline int check_launch(const char *what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return PANIB_E_CUDA;
    }
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return PANIB_OK;
}

#define PANIB_CUDA(call)                                                          \
    do {                                                                          \
        cudaError_t e__ = (call);                                                 \
        if (e__ != cudaSuccess) {                                                 \
            panib::set_error("%s failed: %s", #call, cudaGetErrorString(e__));   \
            return PANIB_E_CUDA;                                                  \
        }                                                                         \
    } while (0)

}  // namespace panib
