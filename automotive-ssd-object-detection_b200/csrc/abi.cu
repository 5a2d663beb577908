// abi.cu -- library-level entry points of libssdhot.so (include/ssdhot.h).
#include <map>
#include <mutex>
#include <utility>

#include "common.cuh"

namespace ssdhot {
unsigned long long g_launches = 0ull;
unsigned long long* g_timeline = nullptr;

// cudaFuncAttributeMaxDynamicSharedMemorySize is a property of (device, kernel): the opt-in is tracked per current device
// under a mutex, raised once per pair (never lowered), and therefore never re-issued inside a graph capture after the first
// (warm-up) call on that device.
int ensure_dyn_smem(const void* kern, size_t bytes) {
    static std::mutex mu;
    static std::map<std::pair<int, const void*>, size_t> configured;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return (int)e;
    std::lock_guard<std::mutex> lock(mu);
    size_t& have = configured[std::make_pair(dev, kern)];
    if (bytes > have) {
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        if (e != cudaSuccess) return (int)e;
        have = bytes;
    }
    return SSDHOT_OK;
}
}

// Debug hook: device buffer of (units x 16) uint64 that the per-image kernels (train_image_kernel,
// nms_image_kernel) fill with %globaltimer stamps of their phases (null = off, the default).
extern "C" int ssdhot_debug_timeline(void* dev_buffer) {
    ssdhot::g_timeline = reinterpret_cast<unsigned long long*>(dev_buffer);
    return SSDHOT_OK;
}

extern "C" int ssdhot_abi_version(void) { return SSDHOT_ABI_VERSION; }

extern "C" unsigned long long ssdhot_launch_count(void) { return ssdhot::g_launches; }

extern "C" const char* ssdhot_status_string(int status) {
    switch (status) {
        case SSDHOT_OK: return "ok";
        case SSDHOT_ERR_NULL: return "a required pointer is NULL";
        case SSDHOT_ERR_SHAPE: return "a size is outside the supported range";
        case SSDHOT_ERR_VALUE: return "a scalar argument is invalid";
        case SSDHOT_ERR_DEVICE: return "not an sm_100 device";
        case SSDHOT_ERR_ALIGN: return "a pointer is not aligned as documented";
        default: break;
    }
    if (status > 0) return cudaGetErrorString((cudaError_t)status);
    return "unknown ssdhot status";
}
