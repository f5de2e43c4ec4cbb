// Error plumbing and device queries behind the C ABI (include/pdm_b200.h).
#include "pdm_common.cuh"

#include <mutex>

namespace pdm {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
    set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
    return PDM_ERR_CUDA;
}

int current_device_info(DeviceInfo* out) {
    static std::mutex mu;
    static DeviceInfo cache[64];
    static bool have[64] = {false};
    int dev = 0;
    PDM_CUDA_CHECK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) { set_error("device ordinal %d out of range", dev); return PDM_ERR_INVALID_ARG; }
    std::lock_guard<std::mutex> lock(mu);
    if (!have[dev]) {
        cudaDeviceProp prop;
        PDM_CUDA_CHECK(cudaGetDeviceProperties(&prop, dev));
        cache[dev] = DeviceInfo{prop.multiProcessorCount, prop.major, prop.minor};
        have[dev] = true;
    }
    *out = cache[dev];
    return PDM_OK;
}

}  // namespace pdm

extern "C" const char* pdm_last_error(void) { return pdm::g_err; }

extern "C" int pdm_abi_version(void) { return PDM_ABI_VERSION; }

extern "C" int pdm_device_info(int device, int* sm_count, int* cc_major, int* cc_minor) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) return pdm::cuda_fail(e, "cudaGetDeviceCount");
    if (device < 0 || device >= n) { pdm::set_error("no CUDA device %d (count %d)", device, n); return PDM_ERR_CUDA; }
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) return pdm::cuda_fail(e, "cudaGetDeviceProperties");
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (cc_major) *cc_major = prop.major;
    if (cc_minor) *cc_minor = prop.minor;
    return PDM_OK;
}
