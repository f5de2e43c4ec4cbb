// Shared host/device helpers for the pdm_b200 kernels.
#pragma once

#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/pdm_b200.h"

namespace pdm {

// ---- error reporting (thread-local message behind pdm_last_error) ----
void set_error(const char* fmt, ...);
int  cuda_fail(cudaError_t e, const char* what);

#define PDM_CUDA_CHECK(expr)                                            \
    do {                                                                \
        cudaError_t _e = (expr);                                        \
        if (_e != cudaSuccess) return ::pdm::cuda_fail(_e, #expr);      \
    } while (0)

#define PDM_REQUIRE(cond, ...)                                          \
    do {                                                                \
        if (!(cond)) { ::pdm::set_error(__VA_ARGS__); return PDM_ERR_INVALID_ARG; } \
    } while (0)

inline cudaStream_t as_stream(pdm_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }
__host__ __device__ inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
__host__ __device__ inline int64_t round_up(int64_t a, int64_t b) { return ceil_div(a, b) * b; }

// Cached per-device facts (sm count, cc).  Returns PDM_OK or an error.
struct DeviceInfo { int sm_count; int cc_major; int cc_minor; };
int current_device_info(DeviceInfo* out);
int device_info(int device, DeviceInfo* out);

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kBigE  = 1.0e38f;     // energy of a masked (out-of-range) dataset column
constexpr float kMaxE  = 1.0e30f;     // clamp on e = (E - m)/T so that w*e stays finite when w == 0

}  // namespace pdm
