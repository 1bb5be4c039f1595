// softdp_host.h -- host-side helpers shared by the translation units of libb200dp.so:
// the thread-local error message behind b200dp_last_error(), cached device attributes and
// the per-kernel dynamic shared-memory grant.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <map>
#include <mutex>
#include <string>
#include <utility>

namespace b200dp_host {

inline thread_local std::string g_err;

inline int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}
inline int cuda_fail(cudaError_t e, const char* what) {
    g_err = std::string(what) + ": " + cudaGetErrorString(e);
    return (int)e;
}

inline bool aligned(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }

struct DevInfo {
    int sms = 0;
    int smem_optin = 0;
    int smem_per_sm = 0;
};

inline bool dev_info(DevInfo& out) {
    static DevInfo cache[64];
    static bool have[64] = {false};
    static std::mutex mu;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return false;
    std::lock_guard<std::mutex> lk(mu);
    if (!have[dev]) {
        DevInfo d;
        if (cudaDeviceGetAttribute(&d.sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return false;
        cudaDeviceGetAttribute(&d.smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
        cudaDeviceGetAttribute(&d.smem_per_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev);
        cache[dev] = d;
        have[dev] = true;
    }
    out = cache[dev];
    return true;
}

// cudaFuncSetAttribute is not free (and may serialise with work in flight): remember,
// per device and kernel, the largest dynamic shared-memory size already granted.
template <class Kern>
int set_smem(Kern k, size_t smem, const char* fn) {
    static std::mutex mu;
    static std::map<std::pair<int, const void*>, size_t> granted;
    int dev = 0;
    cudaGetDevice(&dev);
    const std::pair<int, const void*> key(dev, reinterpret_cast<const void*>(k));
    {
        std::lock_guard<std::mutex> lk(mu);
        auto it = granted.find(key);
        if (it != granted.end() && it->second >= smem) return 0;
    }
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_fail(e, fn);
    std::lock_guard<std::mutex> lk(mu);
    size_t& g = granted[key];
    if (g < smem) g = smem;
    return 0;
}

}  // namespace b200dp_host
