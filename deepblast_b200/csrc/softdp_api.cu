// softdp_api.cu -- C ABI of libb200dp.so (declared in include/b200dp.h).
// Host side: argument checks, TMA tensor-map encoding, launch geometry, launches.
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <utility>
#include <string>

#include "../../include/b200dp.h"
#include "softdp_host.h"
#include "softdp_adjoint.cuh"
#include "softdp_bwd.cuh"
#include "softdp_bwd2.cuh"
#include "softdp_bwd3.cuh"
#include "softdp_fwd.cuh"
#include "softdp_fwd2.cuh"
#include "softdp_fwd3.cuh"
#include "softdp_loss.cuh"
#include "softdp_traceback.cuh"

using namespace b200dp;
using namespace b200dp_host;

namespace {

// The chained kernels run one warp per pair: their time is one pair's latency, flat up to ~7
// pairs per SM, while the hand-off kernels (W warps per pair) scale with the batch.  Measured
// on B200 the two meet at about 4 pairs per SM (256 x 256: forward 0.177 ms chained against
// 0.085 ms hand-off at 2 pairs per SM).
constexpr int kChainedMinPairsPerSM = 4;


QLayout q_layout(int N, int M) {
    QLayout ql;
    ql.K = (N + kTile - 1) / kTile;
    // chained-dense: strip k+1 starts M steps after strip k, so the 31 ramp steps of
    // consecutive strips interleave lane-wise (strip k's tail uses lanes t > sigma - M of a
    // step, strip k+1's head lanes t <= sigma - M) and no step is padding
    ql.strip_stride = (long long)M * kStepFloats;
    ql.pair_stride = ql.K * ql.strip_stride + 31 * kStepFloats;
    return ql;
}

int check_common(const char* fn, int B, int N, int M) {
    if (B < 0 || N < 1 || M < 1) return fail(-1, std::string(fn) + ": need B >= 0, N >= 1, M >= 1");
    if ((long long)N + M > (1 << 24)) return fail(-1, std::string(fn) + ": lattice too large");
    return 0;
}

// Launch geometry.  A pair is cut into K = ceil(N/32) strips; W warps of one CTA work
// on one pair at a time.  Pick W (a power of two) and the grid so that the estimated
// sweep time max(latency, throughput) is smallest (measured on B200: about 7 resident
// warps per SM saturate the forward kernel); ties go to the smaller W (no hand-offs).
struct Geometry {
    int W;
    int grid;
    size_t smem;
};

template <class SmemFn>
int pick_geometry(const char* fn, int B, int N, int M, int flags, SmemFn smem_of, Geometry& g,
                  bool varlen = false) {
    DevInfo di;
    if (!dev_info(di)) return fail(-2, std::string(fn) + ": cannot query the CUDA device");
    const int K = (N + kTile - 1) / kTile;
    int forceW = (flags >> B200DP_WARPS_SHIFT) & 0xF;
    int forceG = (flags >> B200DP_CTAS_SHIFT) & 0xFFFF;
    double best = 1e300;
    g.W = 0;
    for (int W = 1; W <= 8; W *= 2) {
        if (forceW ? (W != forceW) : (W > 1 && W > K)) continue;
        // ragged batches: the host does not know the lengths, but a CTA's strips form one
        // sequence across pairs, so short pairs do not idle the warps of a wide CTA while
        // long pairs need the width: keep the widest CTA that fits
        // (skip W only when 2 W is itself admissible: 2 W <= K and it fits)
        if (!forceW && varlen && W < 8 && 2 * W <= K && smem_of(2 * W, M) <= (size_t)di.smem_optin) continue;
        size_t smem = smem_of(W, M);
        if (smem > (size_t)di.smem_optin) continue;
        int per_sm = (int)((size_t)di.smem_per_sm / (smem + 1024));   // 1 KB reserved per CTA
        per_sm = per_sm < 1 ? 1 : per_sm;
        int by_threads = 2048 / (32 * W);
        if (per_sm > by_threads) per_sm = by_threads;
        if (per_sm > 32) per_sm = 32;
        long long resident = (long long)di.sms * per_sm;
        int grid = (int)(B < resident ? B : resident);
        if (grid < 1) grid = 1;
        // latency term: a pair's strips run back to back on its W warps; throughput term:
        // the SM's issue/XU capacity is shared by all resident warps and does not depend
        // on W, except for a few percent of hand-off cost per doubling
        double rounds = (double)((B + grid - 1) / grid);
        double lat = rounds * (double)((K + W - 1) / W) * (double)(M + 33);
        double thr = (double)B * K * (double)(M + 33) / ((double)di.sms * 7.0);
        double pen = 1.0 + (W >= 2 ? 0.12 : 0.0) + (W >= 4 ? 0.05 : 0.0) + (W >= 8 ? 0.05 : 0.0);
        double t = (lat > thr ? lat : thr) * pen;
        if (t < best * 0.999) {
            best = t;
            g.W = W;
            g.grid = grid;
            g.smem = smem;
        }
    }
    if (!g.W && forceW) {
        // the requested warps-per-pair does not fit in shared memory: take the largest that does
        for (int W = forceW / 2; W >= 1 && !g.W; W /= 2) {
            size_t smem = smem_of(W, M);
            if (smem > (size_t)di.smem_optin) continue;
            g.W = W;
            g.smem = smem;
            long long resident = (long long)di.sms * (long long)((size_t)di.smem_per_sm / (smem + 1024) ? (size_t)di.smem_per_sm / (smem + 1024) : 1);
            g.grid = (int)(B < resident ? B : resident);
        }
    }
    if (!g.W) return fail(-3, std::string(fn) + ": M too large for shared memory boundary rows");
    if (forceG > 0) g.grid = forceG;
    return 0;
}

// Dispatch overrides travel in `flags` (include/b200dp.h); the library reads no environment
// variables.
bool no_chained(int flags) { return (flags & B200DP_NO_CHAINED) != 0; }
int ring_flag(int flags) { return (flags >> B200DP_SQ_RING_SHIFT) & 0xF; }

// Chained backward kernel (softdp_bwd3.cuh).  Same return convention as launch_fwd3.
template <int RING>
int launch_bwd3_t(const BwdParams& p, int grid, size_t smem, bool sw, cudaStream_t st) {
    int rc = 0;
    auto run = [&](auto kern) {
        rc = set_smem(kern, smem, "b200dp_bwd");
        if (!rc) kern<<<grid, 32, smem, st>>>(p);
    };
    if (sw) run(softdp_bwd3_kernel<true, RING>);
    else run(softdp_bwd3_kernel<false, RING>);
    return rc;
}

int launch_bwd3(const BwdParams& p, int B, int N, int M, int mode, int flags, cudaStream_t st) {
    if (no_chained(flags)) return 1;
    DevInfo di;
    if (!dev_info(di)) return 1;
    const int minB = (flags & B200DP_FORCE_CHAINED) ? 1 : kChainedMinPairsPerSM * di.sms;
    if (B < minB || N < kTile || M < 64 || M % 32 != 0) return 1;
    const int want_ring = ring_flag(flags);
    const int forceG = (flags >> B200DP_CTAS_SHIFT) & 0xFFFF;
    const int rings[4] = {3, 4, 6, 2};     // measured on B200: 3 slots win
    int ring = 0, grid = 0;
    size_t smem = 0;
    long long best_rounds = 0;
    for (int ri = 0; ri < 4; ++ri) {
        const int r = rings[ri];
        if (want_ring ? (r != want_ring) : (r == 2)) continue;
        const size_t sm = (size_t)r * kDiagElems * 4 + kB3StageBytes + (size_t)r * 8 + 16 + (size_t)M * 4 + 128;
        if (sm > (size_t)di.smem_optin) continue;
        int per_sm = (int)((size_t)di.smem_per_sm / (sm + 1024));
        if (per_sm > 32) per_sm = 32;
        if (per_sm < 1) continue;
        const long long resident = (long long)di.sms * per_sm;
        const long long rounds = (B + resident - 1) / resident;
        if (!ring || rounds < best_rounds) {
            ring = r;
            grid = (int)((B + rounds - 1) / rounds);
            smem = sm;
            best_rounds = rounds;
        }
    }
    if (!ring) return 1;
    if (forceG > 0) grid = forceG;
    const bool sw = mode == B200DP_MODE_SW;
    int rc = 0;
    if (ring == 6) rc = launch_bwd3_t<6>(p, grid, smem, sw, st);
    else if (ring == 4) rc = launch_bwd3_t<4>(p, grid, smem, sw, st);
    else if (ring == 3) rc = launch_bwd3_t<3>(p, grid, smem, sw, st);
    else rc = launch_bwd3_t<2>(p, grid, smem, sw, st);
    if (rc) return -(1000 + rc);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return -(1000 + cuda_fail(e, "b200dp_bwd launch"));
    return 0;
}

// Chained forward kernel: one warp per CTA, NCH strips side by side, RING tile slots.
// Returns 0 when launched, > 0 when the shape / batch is not for this kernel (the caller
// falls through to softdp_fwd2), < 0 (-(1000 + code)) on error.
template <int NCH, int RING>
int launch_fwd3_t(const CUtensorMap& tmT, const CUtensorMap& tmA, const CUtensorMap& pfT, const CUtensorMap& pfA,
                  const FwdParams& p, int grid, size_t smem, bool sw, cudaStream_t st) {
    int rc = 0;
    auto run = [&](auto kern) {
        rc = set_smem(kern, smem, "b200dp_fwd");
        if (!rc) kern<<<grid, 32, smem, st>>>(tmT, tmA, pfT, pfA, p);
    };
#ifndef B200DP_FWD3_DBG
#define B200DP_FWD3_DBG 0        // diagnostic builds (scripts/gpu_x16.sh): fwd2_step's DBG bits; results are wrong
#endif
    if (sw) run(softdp_fwd3_kernel<true, NCH, RING, B200DP_FWD3_DBG>);
    else run(softdp_fwd3_kernel<false, NCH, RING, B200DP_FWD3_DBG>);
    return rc;
}

int launch_fwd3(const CUtensorMap& tmT, const CUtensorMap& tmA, FwdParams p, int B, int N, int M, int mode,
                int flags, cudaStream_t st) {
    if (no_chained(flags)) return 1;
    DevInfo di;
    if (!dev_info(di)) return 1;
    const int minB = (flags & B200DP_FORCE_CHAINED) ? 1 : kChainedMinPairsPerSM * di.sms;
    if (B < minB || N < kTile || M < 64) return 1;
    // (two chains per warp, NCH = 2, lose to one: the kernel is bound by the memory system, not by
    // issue latency; only NCH = 1 is instantiated)
    const int want_ring = ring_flag(flags);
    const int forceG = (flags >> B200DP_CTAS_SHIFT) & 0xFFFF;
    // measured on B200 with the two-state Q: 4 slots (two events of lead) beat 3 by 2-4 %
    const int rings[4] = {4, 3, 6, 8};
    int ring = 0, grid = 0;
    size_t smem = 0;
    long long best_rounds = 0;
    for (int ri = 0; ri < 4; ++ri) {
        const int r = rings[ri];
        if (want_ring && r != want_ring) continue;
        const size_t sm = (size_t)r * 4096 + (size_t)r * 8 + 16 + (size_t)M * 4 + 128;
        if (sm > (size_t)di.smem_optin) continue;
        int per_sm = (int)((size_t)di.smem_per_sm / (sm + 1024));
        if (per_sm > 32) per_sm = 32;
        if (per_sm < 1) continue;
        const long long resident = (long long)di.sms * per_sm;
        // equal number of pairs per CTA: rounds = ceil(B / resident), grid = ceil(B / rounds);
        // fewer rounds = more chains resident wins, among equals the ring tried first
        const long long rounds = (B + resident - 1) / resident;
        if (!ring || rounds < best_rounds) {
            ring = r;
            grid = (int)((B + rounds - 1) / rounds);
            smem = sm;
            best_rounds = rounds;
        }
    }
    if (!ring) return 1;
    if (forceG > 0) grid = forceG;
    p.pf_tiles = 0;
    p.pf_dist = 0;
    const bool sw = mode == B200DP_MODE_SW;
    int rc = 0;
#define B200DP_F3(RING_) rc = launch_fwd3_t<1, RING_>(tmT, tmA, tmT, tmA, p, grid, smem, sw, st)
    if (ring == 8) B200DP_F3(8); else if (ring == 6) B200DP_F3(6); else if (ring == 4) B200DP_F3(4); else B200DP_F3(3);
#undef B200DP_F3
    if (rc) return -(1000 + rc);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return -(1000 + cuda_fail(e, "b200dp_fwd launch"));
    return 0;
}

}  // namespace

extern "C" {

int b200dp_version(void) { return 100; }

const char* b200dp_last_error(void) { return g_err.c_str(); }

int b200dp_q_layout(int N, int M, int* K, long long* strip_stride, long long* pair_stride, long long* pad) {
    if (N < 1 || M < 1) return fail(-1, "b200dp_q_layout: need N >= 1, M >= 1");
    QLayout ql = q_layout(N, M);
    if (K) *K = ql.K;
    if (strip_stride) *strip_stride = ql.strip_stride;
    if (pair_stride) *pair_stride = ql.pair_stride;
    if (pad) *pad = (long long)kDiagRows * kStepFloats;   // tile reads may run past the last strip
    return 0;
}

int b200dp_fwd(const float* theta, const float* A, float* Q, float* Vt, const int32_t* xlen, const int32_t* ylen,
               int B, int N, int M, int mode, int flags, void* stream) {
    if (int rc = check_common("b200dp_fwd", B, N, M)) return rc;
    if (mode != B200DP_MODE_NW && mode != B200DP_MODE_SW) return fail(-1, "b200dp_fwd: bad mode");
    if (B == 0) return 0;
    if (!theta || !A || !Q || !Vt) return fail(-1, "b200dp_fwd: null pointer");
    if (!aligned(Q, 16)) return fail(-1, "b200dp_fwd: Q storage must be 16-byte aligned");
    Geometry g;
    if (int rc = pick_geometry("b200dp_fwd", B, N, M, flags, fwd_smem_bytes, g, xlen || ylen)) return rc;
    FwdParams p;
    p.theta = theta;
    p.A = A;
    p.Q = Q;
    p.Qin = nullptr;
    p.has_za = 0;
    p.has_e = 0;
    p.Vt = Vt;
    p.d = PairDims{xlen, ylen, B, N, M};
    p.ql = q_layout(N, M);
    p.i0 = mode == B200DP_MODE_SW ? 2 : 1;
    p.flags = flags;
    p.pf_tiles = 0;
    p.pf_dist = 0;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    CUtensorMap tmT, tmA;
    memset(&tmT, 0, sizeof(tmT));
    memset(&tmA, 0, sizeof(tmA));
    bool tma = !(flags & B200DP_NO_TMA) && (M % 4 == 0) && M >= kTile && N >= kTile &&
               aligned(theta, 16) && aligned(A, 16);
    const bool fast = tma && !(flags & B200DP_V1_KERNELS) && N >= kG && M >= 2 * kG;
    if (fast && encode_row_map(&tmT, theta, B, N, M, kG) && encode_row_map(&tmA, A, B, N, M, kG)) {
        // large batches of equal-size lattices: chained single-warp kernel (softdp_fwd3.cuh)
        if (!xlen && !ylen && M % 16 == 0 && !((flags >> B200DP_WARPS_SHIFT) & 0xF)) {
            int rc3 = launch_fwd3(tmT, tmA, p, B, N, M, mode, flags, st);
            if (rc3 <= 0) return rc3 < 0 ? -rc3 - 1000 : 0;    // 0 = launched, < 0 = error, > 0 = not applicable
        }
        // a deeper tile ring lengthens the TMA lead (RING - 2 events of 16 steps); take the
        // deepest ring that costs no residency (read latency under the kernel's own write
        // traffic is several microseconds)
        Geometry g2;
        if (int rc = pick_geometry("b200dp_fwd", B, N, M, flags, fwd2_smem_bytes<3>, g2, xlen || ylen)) return rc;
        int ring = 3;
        {
            const int want = ring_flag(flags) ? ring_flag(flags) : 8;
            Geometry gt;
            if (want >= 4 && pick_geometry("b200dp_fwd", B, N, M, flags, fwd2_smem_bytes<4>, gt, xlen || ylen) == 0 &&
                gt.W == g2.W && gt.grid >= g2.grid) {
                ring = 4;
                g2 = gt;
                if (want >= 6 && pick_geometry("b200dp_fwd", B, N, M, flags, fwd2_smem_bytes<6>, gt, xlen || ylen) == 0 &&
                    gt.W == g2.W && gt.grid >= g2.grid) {
                    ring = 6;
                    g2 = gt;
                    if (want >= 8 &&
                        pick_geometry("b200dp_fwd", B, N, M, flags, fwd2_smem_bytes<8>, gt, xlen || ylen) == 0 &&
                        gt.W == g2.W && gt.grid >= g2.grid) {
                        ring = 8;
                        g2 = gt;
                    }
                }
            }
        }
        int rc2 = 0;
        auto run = [&](auto kern) {
            rc2 = set_smem(kern, g2.smem, "b200dp_fwd");
            if (!rc2) kern<<<g2.grid, 32 * g2.W, g2.smem, st>>>(tmT, tmA, p);
        };
        const bool sw = mode == B200DP_MODE_SW;
        if (ring == 8) {
            sw ? run(softdp_fwd2_kernel<true, 8>) : run(softdp_fwd2_kernel<false, 8>);
        } else if (ring == 6) {
            sw ? run(softdp_fwd2_kernel<true, 6>) : run(softdp_fwd2_kernel<false, 6>);
        } else if (ring == 4) {
            sw ? run(softdp_fwd2_kernel<true, 4>) : run(softdp_fwd2_kernel<false, 4>);
        } else {
            sw ? run(softdp_fwd2_kernel<true, 3>) : run(softdp_fwd2_kernel<false, 3>);
        }
        if (rc2) return rc2;
        cudaError_t e2 = cudaGetLastError();
        if (e2 != cudaSuccess) return cuda_fail(e2, "b200dp_fwd launch");
        return 0;
    }
    if (tma) tma = encode_row_map(&tmT, theta, B, N, M) && encode_row_map(&tmA, A, B, N, M);
    if (tma) {
        if (int rc = set_smem(softdp_fwd_kernel<true>, g.smem, "b200dp_fwd")) return rc;
        softdp_fwd_kernel<true><<<g.grid, 32 * g.W, g.smem, st>>>(tmT, tmA, p);
    } else {
        if (int rc = set_smem(softdp_fwd_kernel<false>, g.smem, "b200dp_fwd")) return rc;
        softdp_fwd_kernel<false><<<g.grid, 32 * g.W, g.smem, st>>>(tmT, tmA, p);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "b200dp_fwd launch");
    return 0;
}

static int bwd_impl(const float* Et, long long et_stride, const float* Q, float* E, float* Ei, int* wrote_ei,
                    const int32_t* xlen, const int32_t* ylen, int B, int N, int M, int mode, int flags, void* stream);

int b200dp_bwd(const float* Et, long long et_stride, const float* Q, float* E, const int32_t* xlen,
               const int32_t* ylen, int B, int N, int M, int mode, int flags, void* stream) {
    return bwd_impl(Et, et_stride, Q, E, nullptr, nullptr, xlen, ylen, B, N, M, mode, flags, stream);
}

int b200dp_bwd_keep_interior(const float* Et, long long et_stride, const float* Q, float* E, float* Ei,
                             int* wrote_ei, int B, int N, int M, int mode, int flags, void* stream) {
    return bwd_impl(Et, et_stride, Q, E, Ei, wrote_ei, nullptr, nullptr, B, N, M, mode, flags, stream);
}

static int bwd_impl(const float* Et, long long et_stride, const float* Q, float* E, float* Ei, int* wrote_ei,
                    const int32_t* xlen, const int32_t* ylen, int B, int N, int M, int mode, int flags, void* stream) {
    if (wrote_ei) *wrote_ei = 0;
    if (int rc = check_common("b200dp_bwd", B, N, M)) return rc;
    if (mode != B200DP_MODE_NW && mode != B200DP_MODE_SW) return fail(-1, "b200dp_bwd: bad mode");
    if (B == 0) return 0;
    if (!Et || !Q || !E) return fail(-1, "b200dp_bwd: null pointer");
    if (!aligned(Q, 16)) return fail(-1, "b200dp_bwd: Q storage must be 16-byte aligned");
    BwdParams p;
    p.Et = Et;
    p.et_stride = et_stride;
    p.Q = Q;
    p.QdE = nullptr;
    p.E = E;
    p.Ei = Ei;
    p.d = PairDims{xlen, ylen, B, N, M};
    p.ql = q_layout(N, M);
    p.i0 = mode == B200DP_MODE_SW ? 2 : 1;
    p.flags = flags;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const bool tma = !(flags & B200DP_NO_TMA);
    Geometry g;
    if (tma && !(flags & B200DP_V1_KERNELS) && !xlen && !ylen && !((flags >> B200DP_WARPS_SHIFT) & 0xF)) {
        // large batches of equal-size lattices: chained single-warp kernel (softdp_bwd3.cuh)
        int rc3 = launch_bwd3(p, B, N, M, mode, flags, st);
        if (rc3 == 0 && wrote_ei && Ei) *wrote_ei = 1;      // only the chained kernel writes the second copy
        if (rc3 <= 0) return rc3 < 0 ? -rc3 - 1000 : 0;
    }
    if (tma && !(flags & B200DP_V1_KERNELS) && M >= 2 * kG) {
        if (int rc = pick_geometry("b200dp_bwd", B, N, M, flags, bwd2_smem_bytes, g, xlen || ylen)) return rc;
        if (mode == B200DP_MODE_SW) {
            if (int rc = set_smem(softdp_bwd2_kernel<true>, g.smem, "b200dp_bwd")) return rc;
            softdp_bwd2_kernel<true><<<g.grid, 32 * g.W, g.smem, st>>>(p);
        } else {
            if (int rc = set_smem(softdp_bwd2_kernel<false>, g.smem, "b200dp_bwd")) return rc;
            softdp_bwd2_kernel<false><<<g.grid, 32 * g.W, g.smem, st>>>(p);
        }
    } else {
        if (int rc = pick_geometry("b200dp_bwd", B, N, M, flags, bwd_smem_bytes, g, xlen || ylen)) return rc;
        if (tma) {
            if (int rc = set_smem(softdp_bwd_kernel<true>, g.smem, "b200dp_bwd")) return rc;
            softdp_bwd_kernel<true><<<g.grid, 32 * g.W, g.smem, st>>>(p);
        } else {
            if (int rc = set_smem(softdp_bwd_kernel<false>, g.smem, "b200dp_bwd")) return rc;
            softdp_bwd_kernel<false><<<g.grid, 32 * g.W, g.smem, st>>>(p);
        }
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "b200dp_bwd launch");
    return 0;
}

int b200dp_adj_fwd(const float* Q, const float* Ztheta, const float* ZA, float* Vtd, float* Qd,
                   const int32_t* xlen, const int32_t* ylen, int B, int N, int M, int flags, void* stream) {
    if (int rc = check_common("b200dp_adj_fwd", B, N, M)) return rc;
    if (B == 0) return 0;
    if (!Q || !Ztheta || !ZA || !Vtd || !Qd) return fail(-1, "b200dp_adj_fwd: null pointer");
    if (!aligned(Q, 16) || !aligned(Qd, 16))
        return fail(-1, "b200dp_adj_fwd: Q/Qd storage must be 16-byte aligned");
    Geometry g;
    if (int rc = pick_geometry("b200dp_adj_fwd", B, N, M, flags, adj_fwd_smem_bytes, g, xlen || ylen)) return rc;
    AdjFwdParams p;
    p.Q = Q;
    p.Ztheta = Ztheta;
    p.ZA = ZA;
    p.Vtd = Vtd;
    p.Qd = Qd;
    p.d = PairDims{xlen, ylen, B, N, M};
    p.ql = q_layout(N, M);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (!(flags & B200DP_NO_TMA)) {
        if (int rc = set_smem(softdp_adj_fwd_kernel<true>, g.smem, "b200dp_adj_fwd")) return rc;
        softdp_adj_fwd_kernel<true><<<g.grid, 32 * g.W, g.smem, st>>>(p);
    } else {
        if (int rc = set_smem(softdp_adj_fwd_kernel<false>, g.smem, "b200dp_adj_fwd")) return rc;
        softdp_adj_fwd_kernel<false><<<g.grid, 32 * g.W, g.smem, st>>>(p);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "b200dp_adj_fwd launch");
    return 0;
}

int b200dp_adj_bwd(const float* E, const float* Q, const float* Qd, float* Ed, const int32_t* xlen,
                   const int32_t* ylen, int B, int N, int M, int flags, void* stream) {
    if (int rc = check_common("b200dp_adj_bwd", B, N, M)) return rc;
    if (B == 0) return 0;
    if (!E || !Q || !Qd || !Ed) return fail(-1, "b200dp_adj_bwd: null pointer");
    if (!aligned(Q, 16) || !aligned(Qd, 16))
        return fail(-1, "b200dp_adj_bwd: Q/Qd storage must be 16-byte aligned");
    Geometry g;
    if (int rc = pick_geometry("b200dp_adj_bwd", B, N, M, flags, adj_bwd_smem_bytes, g, xlen || ylen)) return rc;
    AdjBwdParams p;
    p.E = E;
    p.Q = Q;
    p.Qd = Qd;
    p.Ed = Ed;
    p.d = PairDims{xlen, ylen, B, N, M};
    p.ql = q_layout(N, M);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (!(flags & B200DP_NO_TMA)) {
        if (int rc = set_smem(softdp_adj_bwd_kernel<true>, g.smem, "b200dp_adj_bwd")) return rc;
        softdp_adj_bwd_kernel<true><<<g.grid, 32 * g.W, g.smem, st>>>(p);
    } else {
        if (int rc = set_smem(softdp_adj_bwd_kernel<false>, g.smem, "b200dp_adj_bwd")) return rc;
        softdp_adj_bwd_kernel<false><<<g.grid, 32 * g.W, g.smem, st>>>(p);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "b200dp_adj_bwd launch");
    return 0;
}

int b200dp_traceback(const float* grad, long long sb, long long si, long long sj, const int32_t* xlen,
                     const int32_t* ylen, int B, int N, int M, int variant, int32_t* out, int cap, int32_t* len,
                     void* stream) {
    if (int rc = check_common("b200dp_traceback", B, N, M)) return rc;
    if (variant != 0 && variant != 1) return fail(-1, "b200dp_traceback: bad variant");
    if (B == 0) return 0;
    if (!grad || !out || !len || cap < 1) return fail(-1, "b200dp_traceback: null pointer / cap < 1");
    TracebackParams p;
    p.grad = grad;
    p.sb = sb;
    p.si = si;
    p.sj = sj;
    p.xlen = xlen;
    p.ylen = ylen;
    p.B = B;
    p.N = N;
    p.M = M;
    p.variant = variant;
    p.out = out;
    p.cap = cap;
    p.len = len;
    softdp_traceback_kernel<<<(B + kTbWarps - 1) / kTbWarps, 32 * kTbWarps, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "b200dp_traceback launch");
    return 0;
}

// ---- chained adjoint pair (double backward on large batches of equal-size lattices) ----------
static bool adj3_shape_ok(int B, int N, int M) {
    DevInfo di;
    if (!dev_info(di)) return false;
    // (measured on B200: the chained adjoint pair beats the general kernels at every batch size)
    return B >= 1 && N >= kTile && M >= 64 && M % 32 == 0 && get_encode() != nullptr;
}

static int chained_grid(int B, size_t smem, int& grid, int flags) {
    DevInfo di;
    if (!dev_info(di)) return -1;
    if (smem > (size_t)di.smem_optin) return -1;
    int per_sm = (int)((size_t)di.smem_per_sm / (smem + 1024));
    if (per_sm > 32) per_sm = 32;
    if (per_sm < 1) return -1;
    const long long resident = (long long)di.sms * per_sm;
    const long long rounds = (B + resident - 1) / resident;
    grid = (int)((B + rounds - 1) / rounds);
    if ((flags >> B200DP_CTAS_SHIFT) & 0xFFFF) grid = (flags >> B200DP_CTAS_SHIFT) & 0xFFFF;
    return (int)rounds;
}

int b200dp_adj3_applicable(int B, int N, int M) { return (B > 0 && N > 0 && M > 0 && adj3_shape_ok(B, N, M)) ? 1 : 0; }

int b200dp_adj_fwd3(const float* Q, const float* Zt, const float* ZA, const float* E, float* Vtd, float* QdE, int B,
                    int N, int M, int flags, void* stream) {
    if (int rc = check_common("b200dp_adj_fwd3", B, N, M)) return rc;
    if (B == 0) return 0;
    if (!Q || !Zt || !Vtd || !QdE) return fail(-1, "b200dp_adj_fwd3: null pointer");
    if (!adj3_shape_ok(B, N, M)) return fail(-4, "b200dp_adj_fwd3: shape not taken by the chained kernels (b200dp_adj3_applicable)");
    if (!aligned(Q, 16) || !aligned(QdE, 16) || !aligned(Zt, 16) || (ZA && !aligned(ZA, 16)) || (E && !aligned(E, 16)))
        return fail(-1, "b200dp_adj_fwd3: pointers must be 16-byte aligned");
    CUtensorMap tmZ, tmA, tmE;
    if (!encode_row_map(&tmZ, Zt, B, N, M, kG) || !encode_row_map(&tmA, ZA ? ZA : Zt, B, N, M, kG) ||
        !encode_row_map(&tmE, E ? E : Zt, B, N, M, kG))
        return fail(-2, "b200dp_adj_fwd3: cuTensorMapEncodeTiled failed");
    FwdParams p;
    p.theta = Zt;
    p.A = ZA;
    p.Q = QdE;
    p.Qin = Q;
    p.Vt = Vtd;
    p.d = PairDims{nullptr, nullptr, B, N, M};
    p.ql = q_layout(N, M);
    p.i0 = 1;
    p.flags = flags;
    p.has_za = ZA ? 1 : 0;
    p.has_e = E ? 1 : 0;
    p.pf_tiles = 0;
    p.pf_dist = 0;
    // three Q tile slots unless two save a whole round of CTAs (M >= 512 at 1024 pairs)
    int g3 = 0, g2 = 0;
    const size_t s3 = fwd3_smem_bytes<1, 3, true, 3>(M), s2 = fwd3_smem_bytes<1, 3, true, 2>(M);
    const int r3 = chained_grid(B, s3, g3, flags), r2 = chained_grid(B, s2, g2, flags);
    if (r2 < 0) return fail(-3, "b200dp_adj_fwd3: M too large for shared memory");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (r3 > 0 && r3 <= r2) {
        auto kern = softdp_fwd3_kernel<false, 1, 3, 0, true, 3>;
        if (int rc = set_smem(kern, s3, "b200dp_adj_fwd3")) return rc;
        kern<<<g3, 32, s3, st>>>(tmZ, tmA, tmE, tmE, p);
    } else {
        auto kern = softdp_fwd3_kernel<false, 1, 3, 0, true, 2>;
        if (int rc = set_smem(kern, s2, "b200dp_adj_fwd3")) return rc;
        kern<<<g2, 32, s2, st>>>(tmZ, tmA, tmE, tmE, p);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "b200dp_adj_fwd3 launch");
    return 0;
}

int b200dp_adj_bwd3(const float* Q, const float* QdE, float* Ed, float* Ed_interior, int B, int N, int M, int flags,
                    void* stream) {
    if (int rc = check_common("b200dp_adj_bwd3", B, N, M)) return rc;
    if (B == 0) return 0;
    if (!Q || !QdE || (!Ed && !Ed_interior)) return fail(-1, "b200dp_adj_bwd3: null pointer");
    if (!adj3_shape_ok(B, N, M)) return fail(-4, "b200dp_adj_bwd3: shape not taken by the chained kernels (b200dp_adj3_applicable)");
    if (!aligned(Q, 16) || !aligned(QdE, 16)) return fail(-1, "b200dp_adj_bwd3: Q / QdE storage must be 16-byte aligned");
    BwdParams p;
    p.Et = nullptr;
    p.et_stride = 0;
    p.Q = Q;
    p.QdE = QdE;
    p.E = Ed;
    p.Ei = Ed_interior;
    p.d = PairDims{nullptr, nullptr, B, N, M};
    p.ql = q_layout(N, M);
    p.i0 = 1;
    p.flags = flags;
    // three Q + QdE slots unless the smaller ring saves a whole round of CTAs
    int g3 = 0, g2 = 0;
    const size_t s3 = bwd3_smem_bytes<3, true>(M), s2 = bwd3_smem_bytes<2, true>(M);
    const int r3 = chained_grid(B, s3, g3, flags), r2 = chained_grid(B, s2, g2, flags);
    if (r2 < 0) return fail(-3, "b200dp_adj_bwd3: M too large for shared memory");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (r3 > 0 && r3 <= r2) {
        auto kern = softdp_bwd3_kernel<false, 3, true>;
        if (int rc = set_smem(kern, s3, "b200dp_adj_bwd3")) return rc;
        kern<<<g3, 32, s3, st>>>(p);
    } else {
        auto kern = softdp_bwd3_kernel<false, 2, true>;
        if (int rc = set_smem(kern, s2, "b200dp_adj_bwd3")) return rc;
        kern<<<g2, 32, s2, st>>>(p);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "b200dp_adj_bwd3 launch");
    return 0;
}

int b200dp_mxent_fwd(const float* Ytrue, const float* Ypred, long long pb, long long pi, const float* G,
                     const int32_t* xlen, const int32_t* ylen, int B, int N, int M, float* pair_loss,
                     float* pair_count, void* stream) {
    if (int rc = check_common("b200dp_mxent_fwd", B, N, M)) return rc;
    if (B == 0) return 0;
    if (!Ytrue || !Ypred || !pair_loss || !pair_count) return fail(-1, "b200dp_mxent_fwd: null pointer");
    LossParams p{};
    p.Ytrue = Ytrue; p.Ypred = Ypred; p.G = G; p.xlen = xlen; p.ylen = ylen;
    p.pb = pb; p.pi = pi; p.B = B; p.N = N; p.M = M;
    p.pair_loss = pair_loss; p.pair_count = pair_count;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    // the slabs accumulate into pair_loss / pair_count: start from zero
    cudaError_t e0 = cudaMemsetAsync(pair_loss, 0, (size_t)B * 4, st);
    if (e0 == cudaSuccess) e0 = cudaMemsetAsync(pair_count, 0, (size_t)B * 4, st);
    if (e0 != cudaSuccess) return cuda_fail(e0, "b200dp_mxent_fwd memset");
    // 16-byte loads where every row of the three tensors starts on a 16-byte boundary
    const bool vec = M % 4 == 0 && pb % 4 == 0 && pi % 4 == 0 && aligned(Ytrue, 16) && aligned(Ypred, 16) &&
                     (!G || aligned(G, 16));
    if (vec) softdp_mxent_fwd_kernel<true><<<dim3((N + kLossRows - 1) / kLossRows, B), 256, 0, st>>>(p);
    else softdp_mxent_fwd_kernel<false><<<dim3((N + kLossRows - 1) / kLossRows, B), 256, 0, st>>>(p);
    softdp_mxent_fin_kernel<<<(B + 255) / 256 < 64 ? (B + 255) / 256 : 64, 256, 0, st>>>(p);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "b200dp_mxent_fwd launch");
    return 0;
}

int b200dp_mxent_bwd(const float* Ytrue, const float* Ypred, long long pb, long long pi, const float* G,
                     const int32_t* xlen, const int32_t* ylen, int B, int N, int M, const float* pair_count,
                     const float* gout, float* grad, void* stream) {
    if (int rc = check_common("b200dp_mxent_bwd", B, N, M)) return rc;
    if (B == 0) return 0;
    if (!Ytrue || !Ypred || !pair_count || !gout || !grad) return fail(-1, "b200dp_mxent_bwd: null pointer");
    LossParams p{};
    p.Ytrue = Ytrue; p.Ypred = Ypred; p.G = G; p.xlen = xlen; p.ylen = ylen;
    p.pb = pb; p.pi = pi; p.B = B; p.N = N; p.M = M;
    p.pair_count = const_cast<float*>(pair_count); p.gout = gout; p.grad = grad;
    const bool vec = M % 4 == 0 && pb % 4 == 0 && pi % 4 == 0 && aligned(Ytrue, 16) && aligned(Ypred, 16) &&
                     (!G || aligned(G, 16)) && aligned(grad, 16);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (vec) softdp_mxent_bwd_kernel<true><<<dim3((N + kLossRows - 1) / kLossRows, B), 256, 0, st>>>(p);
    else softdp_mxent_bwd_kernel<false><<<dim3((N + kLossRows - 1) / kLossRows, B), 256, 0, st>>>(p);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "b200dp_mxent_bwd launch");
    return 0;
}

}  // extern "C"

// ---- host-buffer entry point: chunked H2D -> fwd -> bwd -> D2H pipeline ---------------------
//
// What a caller holding theta / A in HOST memory pays for NeedlemanWunschDecoder.decode
// (deepblast/nw_cuda.py:319-325: forward, then autograd.grad of sum(Vt)) is dominated by
// the PCIe copies (8 B/cell in, 4 B/cell out), not by the kernels.  The batch is cut into
// chunks that flow through three workspace slots on three internal streams, so that the
// upload of chunk c+1, the two sweeps of chunk c and the download of chunk c-1 overlap
// (PCIe is full duplex): the call costs about max(H2D, D2H) instead of their sum.
namespace {

struct HostPipe {
    bool ready = false;
    cudaStream_t s_in = nullptr, s_cmp = nullptr, s_out = nullptr, s_walk = nullptr;
    cudaEvent_t fork = nullptr, join = nullptr;
    cudaEvent_t walk_done[3] = {nullptr, nullptr, nullptr};
    cudaEvent_t in_done[3] = {nullptr, nullptr, nullptr};
    cudaEvent_t cmp_done[3] = {nullptr, nullptr, nullptr};
    cudaEvent_t out_done[3] = {nullptr, nullptr, nullptr};
};
std::mutex g_pipe_mu;
HostPipe g_pipes[64];

__global__ void fill_one_kernel(float* p) { p[0] = 1.0f; }

size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

struct HostSlotLayout {
    size_t theta, A, Q, E, Vt, Et, paths, len, bytes;
};
inline int host_path_cap(int N, int M) { return 2 * (N + M) + 8; }
HostSlotLayout host_slot_layout(int N, int M, int cp, bool with_paths = false) {
    const QLayout ql = q_layout(N, M);
    HostSlotLayout L;
    size_t off = 0;
    L.theta = off; off = align256(off + (size_t)cp * N * M * 4);
    L.A = off;     off = align256(off + (size_t)cp * N * M * 4);
    L.Q = off;     off = align256(off + ((size_t)cp * ql.pair_stride + (size_t)kDiagRows * kStepFloats) * 4);
    L.E = off;     off = align256(off + (size_t)cp * (N + 2) * (M + 2) * 4);
    L.Vt = off;    off = align256(off + (size_t)cp * 4);
    L.Et = off;    off = align256(off + (size_t)cp * 4);
    L.paths = off;
    if (with_paths) off = align256(off + (size_t)cp * host_path_cap(N, M) * 3 * 4);
    L.len = off;
    if (with_paths) off = align256(off + (size_t)cp * 4);
    L.bytes = off;
    return L;
}

}  // namespace

extern "C" {

size_t b200dp_decode_host_workspace(int N, int M, int chunk_pairs) {
    if (N < 1 || M < 1 || chunk_pairs < 1) return 0;
    return 3 * host_slot_layout(N, M, chunk_pairs).bytes + 256;
}

size_t b200dp_align_host_workspace(int N, int M, int chunk_pairs) {
    if (N < 1 || M < 1 || chunk_pairs < 1) return 0;
    return 3 * host_slot_layout(N, M, chunk_pairs, true).bytes + 256;
}

int b200dp_align_host_path_cap(int N, int M) { return host_path_cap(N, M); }

// Shared pipeline of b200dp_decode_host (E_h: the padded expected-alignment matrices come back) and
// b200dp_align_host (paths_h / len_h: the walk runs on the device, only the paths come back).
static int host_pipeline(const float* theta_h, const float* A_h, const float* Et_h, float* Vt_h, float* E_h,
                         int32_t* paths_h, int32_t* len_h, int variant, int B, int N, int M, int mode, int chunk_pairs,
                         void* workspace, size_t workspace_bytes, int flags, void* stream) {
    const bool walk = paths_h != nullptr;
    if (int rc = check_common("b200dp_decode_host", B, N, M)) return rc;
    if (mode != B200DP_MODE_NW && mode != B200DP_MODE_SW) return fail(-1, "b200dp_decode_host: bad mode");
    if (B == 0) return 0;
    if (!theta_h || !A_h || !Vt_h || (!E_h && !walk) || (walk && !len_h) || !workspace)
        return fail(-1, "b200dp_decode_host: null pointer");
    if (walk && variant != 0 && variant != 1) return fail(-1, "b200dp_align_host: bad variant");
    if (chunk_pairs < 1) return fail(-1, "b200dp_decode_host: chunk_pairs < 1");
    if (!aligned(workspace, 256)) return fail(-1, "b200dp_decode_host: workspace must be 256-byte aligned");
    if (workspace_bytes < (walk ? b200dp_align_host_workspace(N, M, chunk_pairs) : b200dp_decode_host_workspace(N, M, chunk_pairs)))
        return fail(-1, "b200dp_decode_host: workspace too small (see b200dp_decode_host_workspace / b200dp_align_host_workspace)");
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return fail(-2, "b200dp_decode_host: no device");
    std::lock_guard<std::mutex> lk(g_pipe_mu);      // one pipeline per device, enqueued by one thread at a time
    HostPipe& hp = g_pipes[dev];
    if (!hp.ready) {
        cudaError_t e = cudaSuccess;
        auto mk = [&](cudaEvent_t* ev) { if (e == cudaSuccess) e = cudaEventCreateWithFlags(ev, cudaEventDisableTiming); };
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&hp.s_in, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&hp.s_cmp, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&hp.s_out, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&hp.s_walk, cudaStreamNonBlocking);
        mk(&hp.fork);
        mk(&hp.join);
        for (int i = 0; i < 3; ++i) {
            mk(&hp.in_done[i]);
            mk(&hp.cmp_done[i]);
            mk(&hp.out_done[i]);
            mk(&hp.walk_done[i]);
        }
        if (e != cudaSuccess) return cuda_fail(e, "b200dp_decode_host: stream/event creation");
        hp.ready = true;
    }
    cudaStream_t user = reinterpret_cast<cudaStream_t>(stream);
    const HostSlotLayout L = host_slot_layout(N, M, chunk_pairs, walk);
    const int cap = host_path_cap(N, M);
    unsigned char* ws = static_cast<unsigned char*>(workspace);
    float* one = reinterpret_cast<float*>(ws + 3 * L.bytes);
    const size_t pairTA = (size_t)N * M, pairE = (size_t)(N + 2) * (M + 2);
#define B200DP_CK(call, what)                                      \
    do {                                                           \
        cudaError_t e_ = (call);                                   \
        if (e_ != cudaSuccess) return cuda_fail(e_, what);         \
    } while (0)
    // The enqueue sequence runs inside a lambda so that an error in the middle still reaches the join
    // below: work already queued on the internal streams keeps using the workspace and the host
    // buffers, and the caller's stream must not run past it (the Python side may drop its
    // references as soon as this function returns an error).
    auto enqueue = [&]() -> int {
    // fork: everything below is ordered after the work already queued on the caller's stream
    B200DP_CK(cudaEventRecord(hp.fork, user), "b200dp_decode_host: fork");
    B200DP_CK(cudaStreamWaitEvent(hp.s_in, hp.fork, 0), "b200dp_decode_host: fork");
    B200DP_CK(cudaStreamWaitEvent(hp.s_cmp, hp.fork, 0), "b200dp_decode_host: fork");
    B200DP_CK(cudaStreamWaitEvent(hp.s_out, hp.fork, 0), "b200dp_decode_host: fork");
    B200DP_CK(cudaStreamWaitEvent(hp.s_walk, hp.fork, 0), "b200dp_decode_host: fork");
    if (!Et_h) {
        fill_one_kernel<<<1, 1, 0, hp.s_cmp>>>(one);
        B200DP_CK(cudaGetLastError(), "b200dp_decode_host: fill launch");
    }
    int c = 0;
    for (int b0 = 0; b0 < B; b0 += chunk_pairs, ++c) {
        const int nb = (B - b0 < chunk_pairs) ? (B - b0) : chunk_pairs;
        const int sl = c % 3;
        unsigned char* sb = ws + (size_t)sl * L.bytes;
        float* d_theta = reinterpret_cast<float*>(sb + L.theta);
        float* d_A = reinterpret_cast<float*>(sb + L.A);
        float* d_Q = reinterpret_cast<float*>(sb + L.Q);
        float* d_E = reinterpret_cast<float*>(sb + L.E);
        float* d_Vt = reinterpret_cast<float*>(sb + L.Vt);
        float* d_Et = reinterpret_cast<float*>(sb + L.Et);
        // upload: the slot's theta / A are free once the sweeps of chunk c-3 have run
        B200DP_CK(cudaStreamWaitEvent(hp.s_in, hp.cmp_done[sl], 0), "b200dp_decode_host: wait");
        B200DP_CK(cudaMemcpyAsync(d_theta, theta_h + (size_t)b0 * pairTA, (size_t)nb * pairTA * 4,
                                  cudaMemcpyHostToDevice, hp.s_in), "b200dp_decode_host: H2D theta");
        B200DP_CK(cudaMemcpyAsync(d_A, A_h + (size_t)b0 * pairTA, (size_t)nb * pairTA * 4, cudaMemcpyHostToDevice,
                                  hp.s_in), "b200dp_decode_host: H2D A");
        if (Et_h)
            B200DP_CK(cudaMemcpyAsync(d_Et, Et_h + b0, (size_t)nb * 4, cudaMemcpyHostToDevice, hp.s_in),
                      "b200dp_decode_host: H2D Et");
        B200DP_CK(cudaEventRecord(hp.in_done[sl], hp.s_in), "b200dp_decode_host: record");
        // sweeps: E / Vt of the slot are free once chunk c-3 has been downloaded
        B200DP_CK(cudaStreamWaitEvent(hp.s_cmp, hp.in_done[sl], 0), "b200dp_decode_host: wait");
        B200DP_CK(cudaStreamWaitEvent(hp.s_cmp, hp.out_done[sl], 0), "b200dp_decode_host: wait");
        if (int rc = b200dp_fwd(d_theta, d_A, d_Q, d_Vt, nullptr, nullptr, nb, N, M, mode, flags, hp.s_cmp)) return rc;
        if (int rc = b200dp_bwd(Et_h ? d_Et : one, Et_h ? 1 : 0, d_Q, d_E, nullptr, nullptr, nb, N, M, mode, flags,
                                hp.s_cmp))
            return rc;
        int32_t* d_paths = reinterpret_cast<int32_t*>(sb + L.paths);
        int32_t* d_len = reinterpret_cast<int32_t*>(sb + L.len);
        B200DP_CK(cudaEventRecord(hp.cmp_done[sl], hp.s_cmp), "b200dp_decode_host: record");
        if (walk) {
            // the greedy walk of nw_cuda.py:273-317 over the interior of the padded E, all pairs of the chunk.
            // A walk is a chain of dependent loads (about a microsecond per step whatever the chunk size): it
            // runs on a stream of its own, beside the sweeps of the following chunks
            B200DP_CK(cudaStreamWaitEvent(hp.s_walk, hp.cmp_done[sl], 0), "b200dp_align_host: wait");
            if (int rc = b200dp_traceback(d_E + (M + 2) + 1, (long long)pairE, (long long)(M + 2), 1, nullptr, nullptr, nb, N, M,
                                          variant, d_paths, cap, d_len, hp.s_walk))
                return rc;
            B200DP_CK(cudaEventRecord(hp.walk_done[sl], hp.s_walk), "b200dp_align_host: record");
        }
        // download
        B200DP_CK(cudaStreamWaitEvent(hp.s_out, walk ? hp.walk_done[sl] : hp.cmp_done[sl], 0), "b200dp_decode_host: wait");
        if (walk) {
            B200DP_CK(cudaMemcpyAsync(paths_h + (size_t)b0 * cap * 3, d_paths, (size_t)nb * cap * 3 * 4, cudaMemcpyDeviceToHost,
                                      hp.s_out), "b200dp_align_host: D2H paths");
            B200DP_CK(cudaMemcpyAsync(len_h + b0, d_len, (size_t)nb * 4, cudaMemcpyDeviceToHost, hp.s_out),
                      "b200dp_align_host: D2H lengths");
        }
        if (E_h)
            B200DP_CK(cudaMemcpyAsync(E_h + (size_t)b0 * pairE, d_E, (size_t)nb * pairE * 4, cudaMemcpyDeviceToHost,
                                      hp.s_out), "b200dp_decode_host: D2H E");
        B200DP_CK(cudaMemcpyAsync(Vt_h + b0, d_Vt, (size_t)nb * 4, cudaMemcpyDeviceToHost, hp.s_out),
                  "b200dp_decode_host: D2H Vt");
        B200DP_CK(cudaEventRecord(hp.out_done[sl], hp.s_out), "b200dp_decode_host: record");
    }
    return 0;
    };
    const int rc = enqueue();
    // join: the caller's stream continues after the last download (s_out runs in order, and every
    // download waited for its sweeps, which waited for their uploads); after an error all three
    // internal streams are joined, whatever they got to
    cudaStream_t tojoin[4] = {hp.s_out, hp.s_cmp, hp.s_in, hp.s_walk};
    for (int i = 0; i < (rc ? 4 : 1); ++i) {
        cudaError_t e1 = cudaEventRecord(hp.join, tojoin[i]);
        if (e1 == cudaSuccess) e1 = cudaStreamWaitEvent(user, hp.join, 0);
        if (e1 != cudaSuccess && !rc) return cuda_fail(e1, "b200dp_decode_host: join");
    }
    if (rc) return rc;
#undef B200DP_CK
    return 0;
}

int b200dp_decode_host(const float* theta_h, const float* A_h, const float* Et_h, float* Vt_h, float* E_h, int B,
                       int N, int M, int mode, int chunk_pairs, void* workspace, size_t workspace_bytes, int flags,
                       void* stream) {
    return host_pipeline(theta_h, A_h, Et_h, Vt_h, E_h, nullptr, nullptr, 0, B, N, M, mode, chunk_pairs, workspace,
                         workspace_bytes, flags, stream);
}

int b200dp_align_host(const float* theta_h, const float* A_h, float* Vt_h, int32_t* paths_h, int32_t* len_h, float* E_h,
                      int B, int N, int M, int mode, int variant, int chunk_pairs, void* workspace,
                      size_t workspace_bytes, int flags, void* stream) {
    if (mode != B200DP_MODE_NW && mode != B200DP_MODE_SW) return fail(-1, "b200dp_align_host: bad mode");
    if (!paths_h || !len_h) return fail(-1, "b200dp_align_host: null pointer");
    return host_pipeline(theta_h, A_h, nullptr, Vt_h, E_h, paths_h, len_h, variant, B, N, M, mode, chunk_pairs, workspace,
                         workspace_bytes, flags, stream);
}

}  // extern "C"
