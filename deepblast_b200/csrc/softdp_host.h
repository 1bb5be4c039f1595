// softdp_host.h -- host-side helpers shared by the translation units of libb200dp.so:
// the thread-local error message behind b200dp_last_error(), cached device attributes and
// the per-kernel dynamic shared-memory grant.
#pragma once
#include <cuda.h>
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

// ---- TMA tensor maps (cuTensorMapEncodeTiled through the runtime's driver entry point) ----------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    return fn;
}

// rank-3 map over a contiguous [B, N, M] fp32 tensor, box boxdim cols x boxrows rows x 1 (default: 32 x 32)
inline bool encode_row_map(CUtensorMap* map, const float* ptr, int B, int N, int M, int boxdim = 32, int boxrows = 0) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return false;
    cuuint64_t dims[3] = {(cuuint64_t)M, (cuuint64_t)N, (cuuint64_t)B};
    cuuint64_t strides[2] = {(cuuint64_t)M * 4, (cuuint64_t)N * M * 4};
    cuuint32_t box[3] = {(cuuint32_t)boxdim, (cuuint32_t)(boxrows ? boxrows : boxdim), 1};
    cuuint32_t estr[3] = {1, 1, 1};
    // L2 promotion 256 B: a box row is only 64-128 B, but the neighbouring columns of the row
    // are consumed a few blocks later, and DRAM serves 256-byte pieces far better than 64-byte
    // ones (measured on B200, C2 forward: 0.272 ms with 128 B promotion, 0.246 ms with 256 B)
    const CUtensorMapL2promotion promo = CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, promo,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

}  // namespace b200dp_host
