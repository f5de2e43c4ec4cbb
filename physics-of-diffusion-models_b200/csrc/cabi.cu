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

// cudaGetDeviceProperties takes milliseconds (it walks the whole driver property table, and contends with
// anything else that talks to the resource manager, e.g. a polling nvidia-smi); the three facts the
// launchers need are read once per device with cudaDeviceGetAttribute and cached.
int device_info(int dev, DeviceInfo* out) {
    static std::mutex mu;
    static DeviceInfo cache[64];
    static bool have[64] = {false};
    if (dev < 0 || dev >= 64) { set_error("device ordinal %d out of range", dev); return PDM_ERR_INVALID_ARG; }
    std::lock_guard<std::mutex> lock(mu);
    if (!have[dev]) {
        int n = 0;
        PDM_CUDA_CHECK(cudaGetDeviceCount(&n));
        if (dev >= n) { set_error("no CUDA device %d (count %d)", dev, n); return PDM_ERR_CUDA; }
        DeviceInfo d;
        PDM_CUDA_CHECK(cudaDeviceGetAttribute(&d.sm_count, cudaDevAttrMultiProcessorCount, dev));
        PDM_CUDA_CHECK(cudaDeviceGetAttribute(&d.cc_major, cudaDevAttrComputeCapabilityMajor, dev));
        PDM_CUDA_CHECK(cudaDeviceGetAttribute(&d.cc_minor, cudaDevAttrComputeCapabilityMinor, dev));
        cache[dev] = d;
        have[dev] = true;
    }
    *out = cache[dev];
    return PDM_OK;
}

int current_device_info(DeviceInfo* out) {
    int dev = 0;
    PDM_CUDA_CHECK(cudaGetDevice(&dev));
    return device_info(dev, out);
}

}  // namespace pdm

extern "C" const char* pdm_last_error(void) { return pdm::g_err; }

extern "C" int pdm_abi_version(void) { return PDM_ABI_VERSION; }

extern "C" int pdm_device_info(int device, int* sm_count, int* cc_major, int* cc_minor) {
    pdm::DeviceInfo d;
    int rc = pdm::device_info(device, &d);
    if (rc != PDM_OK) return rc;
    if (sm_count) *sm_count = d.sm_count;
    if (cc_major) *cc_major = d.cc_major;
    if (cc_minor) *cc_minor = d.cc_minor;
    return PDM_OK;
}
