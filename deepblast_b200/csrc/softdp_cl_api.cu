// softdp_cl_api.cu -- C ABI of the cluster kernels (softdp_cl.cuh): small batches of long, equal-size
// lattices, one thread-block cluster per pair, boundary rows handed over through distributed shared memory.
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstring>
#include <string>

#include "../../include/b200dp.h"
#include "softdp_host.h"
#include "softdp_cl.cuh"

using namespace b200dp;
using namespace b200dp_host;

namespace {

// CTAs per cluster: one per strip up to 8, and no more than the pipeline can keep busy (a CTA's
// next strip is C strips further down: with a hop of about 42 steps per strip the first CTA is free
// again after M + 31 steps, so more than (M + 31) / 42 + 1 CTAs would only wait)
int cl_pick_size(int N, int M, int want) {
    const int K = (N + kTile - 1) / kTile;
    int c = std::min(8, K);
    c = std::min(c, (M + 31) / 42 + 1);
    if (want > 0) c = std::min(want, 8);
    int p = 1;
    while (p * 2 <= c) p *= 2;                   // 1, 2, 4, 8: sizes that tile the 148 SMs well
    return p;
}

// warps per CTA: the ring of a cluster has csize x W warps; more than the pair has strips is useless, and a
// CTA must leave room for every cluster of the batch to be resident (a pair waits for a free cluster otherwise)
int cl_pick_warps(int B, int N, int csize, size_t warp_smem, int want) {
    DevInfo di;
    if (!dev_info(di)) return 1;
    const int K = (N + kTile - 1) / kTile;
    int W = 1;
    while (W < 4 && csize * W < K) {
        const int c = W * 2;
        if ((size_t)c * warp_smem + 1024 > (size_t)di.smem_optin) break;
        const long long per_sm = (long long)di.smem_per_sm / ((long long)c * (long long)warp_smem + 1024);
        if (per_sm * di.sms < (long long)B * csize) break;                     // not all clusters resident any more
        W = c;
    }
    if (want > 0) W = std::min(want, 4);
    return W;
}

template <class Kern, class... Args>
int cl_launch(Kern k, size_t warp_smem, int csize, int W, int B, cudaStream_t st, const char* fn, Args... args) {
    DevInfo di;
    if (!dev_info(di)) return fail(-2, std::string(fn) + ": cannot query the CUDA device");
    const size_t smem = warp_smem * (size_t)W;
    if (smem > (size_t)di.smem_optin) return fail(-3, std::string(fn) + ": lattice too wide for the cluster kernels");
    if (int rc = set_smem(k, smem, fn)) return rc;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)csize;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.blockDim = dim3(32u * (unsigned)W, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    // resident clusters: every cluster loops over pairs c, c + #clusters, ...
    cfg.gridDim = dim3((unsigned)csize, 1, 1);
    int maxcl = 0;
    if (cudaOccupancyMaxActiveClusters(&maxcl, k, &cfg) != cudaSuccess || maxcl < 1) {
        cudaGetLastError();
        return fail(-3, std::string(fn) + ": no cluster of this size fits on the device");
    }
    const int ncl = std::max(1, std::min(B, maxcl));
    cfg.gridDim = dim3((unsigned)(ncl * csize), 1, 1);
    cudaError_t e = cudaLaunchKernelEx(&cfg, k, args...);
    if (e != cudaSuccess) return cuda_fail(e, fn);
    return 0;
}

int cl_check(const char* fn, int B, int N, int M, int mode) {
    if (B < 0 || N < 1 || M < 1) return fail(-1, std::string(fn) + ": need B >= 0, N >= 1, M >= 1");
    if (mode != B200DP_MODE_NW && mode != B200DP_MODE_SW) return fail(-1, std::string(fn) + ": bad mode");
    if (M % 4 != 0) return fail(-5, std::string(fn) + ": the cluster kernels need M % 4 == 0 (16-byte rows)");
    if (N <= kTile) return fail(-5, std::string(fn) + ": a single strip has nothing to hand over (use b200dp_sq_*)");
    return 0;
}

}  // namespace

extern "C" {

int b200dp_cl_applicable(int B, int N, int M) {
    DevInfo di;
    if (B < 1 || N <= kTile || M < 64 || M % 4 != 0 || !dev_info(di) || !get_encode()) return 0;
    if (cl_bwd_smem_bytes<2>(M) > (size_t)di.smem_optin) return 0;
    // Measured on B200 against the strip-queue kernels (scripts/gpu_x7.py): the cluster FORWARD wins 5-13 % while
    // the ring of a cluster covers the pair (at most 32 strips) and the batch is small (2 x 96 x 128: 0.028 against
    // 0.030 ms, 64 x 256^2: 0.061 / 0.066, 32 x 512^2: 0.120 / 0.129, 1 x 1024^2: 0.230 / 0.245); from 32 x 1024^2
    // on the two are level, with more strips per pair the queue wins.  (The cluster BACKWARD loses 5-20 %
    // everywhere: its steps are too short for the hand-off to matter; the Python side does not dispatch it.)
    const long long K = (N + kTile - 1) / kTile;
    return (K <= 32 && (long long)B * K <= 512) ? cl_pick_size(N, M, 0) : 0;
}

int b200dp_cl_fwd(const float* theta, const float* A, float* Q, float* Vt, int B, int N, int M, int mode,
                  int flags, void* stream) {
    if (int rc = cl_check("b200dp_cl_fwd", B, N, M, mode)) return rc;
    if (B == 0) return 0;
    if (!theta || !A || !Vt) return fail(-1, "b200dp_cl_fwd: null pointer");
    if (!aligned(theta, 16) || !aligned(A, 16) || (Q && !aligned(Q, 16)))
        return fail(-1, "b200dp_cl_fwd: theta, A and Q must be 16-byte aligned");
    CUtensorMap tmT, tmA;
    if (!encode_row_map(&tmT, theta, B, N, M, kG) || !encode_row_map(&tmA, A, B, N, M, kG))
        return fail(-4, "b200dp_cl_fwd: cuTensorMapEncodeTiled failed");
    ClParams p;
    memset(&p, 0, sizeof(p));
    p.theta = theta;
    p.A = A;
    p.Q = Q;
    p.Vt = Vt;
    p.B = B;
    p.N = N;
    p.M = M;
    const int csize = cl_pick_size(N, M, (flags >> B200DP_WARPS_SHIFT) & 0xF);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const bool sw = mode == B200DP_MODE_SW;
    const size_t smem = cl_fwd_smem_bytes<4>(M);
    const int W = cl_pick_warps(B, N, csize, smem, (flags >> B200DP_CTAS_SHIFT) & 0xF);
#define B200DP_CLF(SW_, STORE_) \
    return cl_launch(softdp_cl_fwd_kernel<SW_, STORE_, 4>, smem, csize, W, B, st, "b200dp_cl_fwd", p, tmT, tmA)
    if (!Q) {
        if (sw) B200DP_CLF(true, false);
        B200DP_CLF(false, false);
    }
    if (sw) B200DP_CLF(true, true);
    B200DP_CLF(false, true);
#undef B200DP_CLF
}

int b200dp_cl_bwd(const float* Et, long long et_stride, const float* Q, float* E, int B, int N, int M, int mode,
                  int flags, void* stream) {
    if (int rc = cl_check("b200dp_cl_bwd", B, N, M, mode)) return rc;
    if (B == 0) return 0;
    if (!Et || !Q || !E) return fail(-1, "b200dp_cl_bwd: null pointer");
    if (!aligned(Q, 16)) return fail(-1, "b200dp_cl_bwd: Q storage must be 16-byte aligned");
    ClParams p;
    memset(&p, 0, sizeof(p));
    p.Et = Et;
    p.et_stride = et_stride;
    p.Qin = Q;
    p.Eout = E;
    p.B = B;
    p.N = N;
    p.M = M;
    const int csize = cl_pick_size(N, M, (flags >> B200DP_WARPS_SHIFT) & 0xF);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const size_t smem = cl_bwd_smem_bytes<2>(M);
    const int W = cl_pick_warps(B, N, csize, smem, (flags >> B200DP_CTAS_SHIFT) & 0xF);
    if (mode == B200DP_MODE_SW)
        return cl_launch(softdp_cl_bwd_kernel<true, 2>, smem, csize, W, B, st, "b200dp_cl_bwd", p);
    return cl_launch(softdp_cl_bwd_kernel<false, 2>, smem, csize, W, B, st, "b200dp_cl_bwd", p);
}

}  // extern "C"
