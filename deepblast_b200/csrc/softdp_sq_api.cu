// softdp_sq_api.cu -- C ABI of the strip-queue kernels (softdp_sq.cuh): the host-side plan
// builder (strip records in dependency order, packed offsets) and the four launchers.
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <functional>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/b200dp.h"
#include "softdp_host.h"
#include "softdp_sq.cuh"

using namespace b200dp;
using namespace b200dp_host;

namespace {

constexpr size_t kSqCtlBytes = 256;       // control word (+ padding) in front of the boundary scratch

// resident CTAs (= warps) per SM of one kernel instantiation at its shared-memory size, cached
template <class Kern>
int sq_occupancy(Kern k, size_t smem) {
    static std::mutex mu;
    static std::map<std::pair<int, const void*>, int> cache;
    int dev = 0;
    cudaGetDevice(&dev);
    const std::pair<int, const void*> key(dev, reinterpret_cast<const void*>(k));
    {
        std::lock_guard<std::mutex> lk(mu);
        auto it = cache.find(key);
        if (it != cache.end()) return it->second;
    }
    if (set_smem(k, smem, "b200dp_sq")) return 0;
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, 32, smem) != cudaSuccess) occ = 0;
    std::lock_guard<std::mutex> lk(mu);
    cache[key] = occ;
    return occ;
}

template <class Kern, class... Maps>
int sq_launch(Kern k, size_t smem, const SqParams& p, int flags, cudaStream_t st, const char* fn, const Maps&... maps) {
    DevInfo di;
    if (!dev_info(di)) return fail(-2, std::string(fn) + ": cannot query the CUDA device");
    if (int rc = set_smem(k, smem, fn)) return rc;
    const int occ = sq_occupancy(k, smem);
    if (occ < 1) return fail(-3, std::string(fn) + ": kernel does not fit on an SM");
    long long grid = (long long)di.sms * occ;
    const int forceG = (flags >> B200DP_CTAS_SHIFT) & 0xFFFF;
    if (forceG > 0) grid = forceG;
    if (grid > p.nstrips) grid = p.nstrips;
    if (grid < 1) grid = 1;
    k<<<(int)grid, 32, smem, st>>>(p, maps...);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, fn);
    return 0;
}

int sq_check(const char* fn, const void* tab, int nstrips, const void* ws) {
    if (nstrips < 0) return fail(-1, std::string(fn) + ": nstrips < 0");
    if (nstrips == 0) return 0;
    if (!tab || !ws) return fail(-1, std::string(fn) + ": null plan table / workspace");
    if (!aligned(tab, 16) || !aligned(ws, 256)) return fail(-1, std::string(fn) + ": table must be 16-byte, workspace 256-byte aligned");
    return 0;
}

SqParams sq_params(const void* tab, int nstrips, void* ws) {
    SqParams p;
    memset(&p, 0, sizeof(p));
    p.tab = static_cast<const StripRec*>(tab);
    p.nstrips = nstrips;
    p.ctl = static_cast<unsigned long long*>(ws);
    p.bnd = reinterpret_cast<unsigned long long*>(static_cast<unsigned char*>(ws) + kSqCtlBytes);
    return p;
}

int ring_of(int flags) { return (flags >> B200DP_SQ_RING_SHIFT) & 0xF; }

thread_local void* g_trace = nullptr;      // b200dp_sq_set_trace (diagnostics)

}  // namespace

extern "C" {

int b200dp_plan_build(const int32_t* xlen, const int32_t* ylen, int B, int N, int M, int packed, int warps_fwd,
                      int warps_bwd, b200dp_plan_info* info, long long* pair_off, long long* q_off, void* fwd_tab, void* bwd_tab,
                      int tab_capacity) {
    if (B < 0 || N < 1 || M < 1) return fail(-1, "b200dp_plan_build: need B >= 0, N >= 1, M >= 1");
    if (!info) return fail(-1, "b200dp_plan_build: info is null");
    if (!packed && (M % 4) != 0)
        return fail(-5, "b200dp_plan_build: the dense layout needs M % 4 == 0 (16-byte rows); use the packed layout");
    constexpr long long kHopSteps = 56;          // steps by which a strip trails the strip it depends on
    std::vector<int> n(B), m(B), K(B);
    std::vector<long long> toff(B), qoff(B), boff(B);
    long long tcur = 0, qcur = 0, bcur = 0, cells = 0;
    int nstrips = 0, maxm = 0;
    for (int b = 0; b < B; ++b) {
        // lengths are clamped to the tensor like the reference's slices theta[b, :n, :m]
        // (deepblast/alignment.py:166-169); an empty pair has no strips
        int nb = xlen ? xlen[b] : N, mb = ylen ? ylen[b] : M;
        nb = nb < 0 ? 0 : (nb > N ? N : nb);
        mb = mb < 0 ? 0 : (mb > M ? M : mb);
        if (nb == 0 || mb == 0) nb = mb = 0;
        n[b] = nb;
        m[b] = mb;
        K[b] = (nb + kTile - 1) / kTile;
        nstrips += K[b];
        maxm = std::max(maxm, mb);
        cells += (long long)nb * mb;
        if (packed) {
            const int pitch = (mb + 3) & ~3;
            tcur = (tcur + 31) & ~31ll;                    // every pair starts on a 128-byte line
            toff[b] = tcur;
            tcur += (long long)nb * pitch;
        } else {
            toff[b] = (long long)b * N * M;
        }
        qoff[b] = qcur;
        if (K[b] > 0) qcur += (long long)K[b] * mb * kStepFloats + 31ll * kStepFloats;
        boff[b] = bcur;
        if (K[b] > 1) bcur += (long long)(K[b] - 1) * mb;
        if (pair_off) pair_off[b] = toff[b];
        if (q_off) q_off[b] = qoff[b];
    }
    // Grid size of each sweep.  work = warp-steps of all strips, chain = the longest dependency
    // chain (strip k+1 trails strip k by kHopSteps).  With every warp productive a step slows
    // down roughly linearly in the resident warps per SM, so when the batch is bound by its
    // longest chain (ragged batches with a few long pairs) more resident warps only slow that
    // chain: keep just enough warps that the throughput time matches the chain
    // (measured on B200, BASELINE configs[4]: 8 warps per SM 0.345 ms forward, 13 per SM 0.426 ms).
    long long work = 0, chain = 1;
    for (int b = 0; b < B; ++b) {
        if (K[b] == 0) continue;
        work += (long long)K[b] * (m[b] + 31);
        chain = std::max(chain, (long long)(K[b] - 1) * kHopSteps + m[b] + 31);
    }
    auto pick_grid = [&](int resident) {
        if (resident < 1) resident = 148 * 12;
        if (nstrips <= resident) return std::max(nstrips, 1);
        long long want = (work + chain / 2) / chain;                    // warps at which throughput time = chain time
        const long long floor_ = (long long)resident * 45 / 100;        // (below that the step time no longer improves)
        if (want < floor_) want = floor_;
        return (int)std::min<long long>(want, resident);
    };
    info->grid_fwd = pick_grid(warps_fwd);
    info->grid_bwd = pick_grid(warps_bwd);
    info->nstrips = nstrips;
    info->max_m = maxm;
    info->q_floats = qcur + (long long)kDiagRows * kStepFloats;     // tile reads may run past the last strip
    info->bnd_words = bcur;
    info->packed_floats = packed ? ((tcur + 31) & ~31ll) : (long long)B * N * M;
    info->cells = cells;
    if (!fwd_tab && !bwd_tab) return 0;
    if (tab_capacity < nstrips) return fail(-1, "b200dp_plan_build: table capacity too small");

    // Ticket order = a LIST SCHEDULE computed here on the host: `warps` virtual warps, every
    // strip takes m + 31 steps, strip k+1 of a pair may start kHopSteps after strip k started
    // (it trails its predecessor by the 31 steps of skew inside a strip plus the hand-off);
    // whenever a warp is free it takes, among the strips that may start, the one with the longest
    // remaining dependency chain.  Tickets are numbered in the order of these virtual starts,
    // so every strip comes after the strip it depends on, long pairs are pipelined from the
    // first moment with other pairs' strips filling the gaps, and a warp seldom takes a strip
    // whose predecessor is not yet far enough ahead.
    StripRec* ft = static_cast<StripRec*>(fwd_tab);
    StripRec* bt = static_cast<StripRec*>(bwd_tab);
    auto fill = [&](StripRec& s, int b, int k, bool fwd) {
        const int pitch = packed ? ((m[b] + 3) & ~3) : M;
        memset(&s, 0, sizeof(s));
        s.t_off = toff[b] + (long long)k * kTile * pitch;
        s.q_off = qoff[b] + (long long)k * m[b] * kStepFloats;
        // boundary j sits between strips j and j+1
        const long long above = k > 0 ? boff[b] + (long long)(k - 1) * m[b] : -1;
        const long long below = k + 1 < K[b] ? boff[b] + (long long)k * m[b] : -1;
        s.b_in = fwd ? above : below;
        s.b_out = fwd ? below : above;
        s.rows = std::min(kTile, n[b] - k * kTile);
        s.m = m[b];
        s.pitch = pitch;
        s.pair = b;
        s.flags = (k == 0 ? kSqFirst : 0) | (k + 1 == K[b] ? kSqLast : 0);
        s.k = k;
    };
    auto schedule = [&](int warps, StripRec* tab, bool fwd) {
        if (!tab) return;
        if (warps < 1) warps = 148 * 12;
        struct Item {
            long long key;      // ready heap: remaining chain (max first); pending heap: ready time (min first)
            int b, r;
        };
        auto by_prio = [](const Item& x, const Item& y) { return x.key < y.key || (x.key == y.key && x.b > y.b); };
        auto by_time = [](const Item& x, const Item& y) { return x.key > y.key || (x.key == y.key && x.b > y.b); };
        std::vector<Item> ready, pending;
        std::vector<long long> freeat;                     // min-heap of warp free times
        auto chain = [&](int b, int r) { return (long long)(K[b] - 1 - r) * kHopSteps + m[b] + 31; };
        for (int b = 0; b < B; ++b)
            if (K[b] > 0) ready.push_back(Item{chain(b, 0), b, 0});
        std::make_heap(ready.begin(), ready.end(), by_prio);
        const int nw = std::min<long long>(warps, std::max(1, nstrips));
        freeat.assign(nw, 0);
        int tk = 0;
        while (tk < nstrips) {
            std::pop_heap(freeat.begin(), freeat.end(), std::greater<long long>());
            long long t = freeat.back();
            if (ready.empty() && !pending.empty() && pending.front().key > t) t = pending.front().key;
            while (!pending.empty() && pending.front().key <= t) {
                std::pop_heap(pending.begin(), pending.end(), by_time);
                Item it = pending.back();
                pending.pop_back();
                it.key = chain(it.b, it.r);
                ready.push_back(it);
                std::push_heap(ready.begin(), ready.end(), by_prio);
            }
            std::pop_heap(ready.begin(), ready.end(), by_prio);
            const Item it = ready.back();
            ready.pop_back();
            fill(tab[tk++], it.b, fwd ? it.r : K[it.b] - 1 - it.r, fwd);
            freeat.back() = t + m[it.b] + 31;
            std::push_heap(freeat.begin(), freeat.end(), std::greater<long long>());
            if (it.r + 1 < K[it.b]) {
                pending.push_back(Item{t + kHopSteps, it.b, it.r + 1});
                std::push_heap(pending.begin(), pending.end(), by_time);
            }
        }
    };
    schedule(info->grid_fwd, ft, true);
    schedule(info->grid_bwd, bt, false);
    return 0;
}

void b200dp_sq_set_trace(void* trace) { g_trace = trace; }

size_t b200dp_sq_workspace_bytes(long long bnd_words) {
    return kSqCtlBytes + (size_t)(bnd_words > 0 ? bnd_words : 0) * 8 + 256;
}

int b200dp_sq_resident_warps(int kind) {
    DevInfo di;
    if (!dev_info(di)) return 0;
    int occ = 0;
    switch (kind) {
        case 0: occ = sq_occupancy(softdp_sq_fwd_kernel<false, false, true, 4>, sq_fwd_smem_bytes<false, 4>()); break;
        case 1: occ = sq_occupancy(softdp_sq_bwd_kernel<false, false, 2>, sq_bwd_smem_bytes<2, false>()); break;
        case 2: occ = sq_occupancy(softdp_sq_fwd_kernel<false, true, true, 3>, sq_fwd_smem_bytes<true, 3>()); break;
        default: occ = sq_occupancy(softdp_sq_bwd_kernel<false, true, 2>, sq_bwd_smem_bytes<2, true>()); break;
    }
    return di.sms * occ;
}

int b200dp_sq_fwd(const void* tab, int nstrips, void* workspace, const float* theta, const float* A,
                  float* Q, float* Vt, int mode, int flags, void* stream) {
    if (int rc = sq_check("b200dp_sq_fwd", tab, nstrips, workspace)) return rc;
    if (mode != B200DP_MODE_NW && mode != B200DP_MODE_SW) return fail(-1, "b200dp_sq_fwd: bad mode");
    if (nstrips == 0) return 0;
    if (!theta || !A || !Vt) return fail(-1, "b200dp_sq_fwd: null pointer");
    if (!aligned(theta, 16) || !aligned(A, 16) || (Q && !aligned(Q, 16)))
        return fail(-1, "b200dp_sq_fwd: theta, A and Q must be 16-byte aligned");
    SqParams p = sq_params(tab, nstrips, workspace);
    p.dbg = (flags >> B200DP_SQ_DBG_SHIFT) & 0xF;
    p.trace = static_cast<unsigned long long*>(g_trace);
    p.theta = theta;
    p.A = A;
    p.Q = Q;
    p.Vt = Vt;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const bool sw = mode == B200DP_MODE_SW;
    const int ring = ring_of(flags);
#define B200DP_SQF(SW_, STORE_, RING_) \
    return sq_launch(softdp_sq_fwd_kernel<SW_, false, STORE_, RING_>, sq_fwd_smem_bytes<false, RING_>(), p, flags, st, "b200dp_sq_fwd")
    if (!Q) {
        // score only (deepblast/alignment.py:127-137: ddp(theta, A) under no_grad): Vt alone, 8 B/cell
        if (sw) B200DP_SQF(true, false, 4);
        B200DP_SQF(false, false, 4);
    }
    if (ring == 3) {
        if (sw) B200DP_SQF(true, true, 3);
        B200DP_SQF(false, true, 3);
    }
    if (ring == 6) {
        if (sw) B200DP_SQF(true, true, 6);
        B200DP_SQF(false, true, 6);
    }
    if (ring == 8) {
        if (sw) B200DP_SQF(true, true, 8);
        B200DP_SQF(false, true, 8);
    }
    if (sw) B200DP_SQF(true, true, 4);
    B200DP_SQF(false, true, 4);
#undef B200DP_SQF
}

int b200dp_sq_fwd_dense(const void* tab, int nstrips, void* workspace, const float* theta, const float* A,
                        float* Q, float* Vt, int B, int N, int M, int mode, int flags, void* stream) {
    if (int rc = sq_check("b200dp_sq_fwd_dense", tab, nstrips, workspace)) return rc;
    if (mode != B200DP_MODE_NW && mode != B200DP_MODE_SW) return fail(-1, "b200dp_sq_fwd_dense: bad mode");
    if (nstrips == 0) return 0;
    if (!theta || !A || !Vt) return fail(-1, "b200dp_sq_fwd_dense: null pointer");
    if (B < 1 || N < 1 || M < 4 || (M % 4) != 0) return fail(-1, "b200dp_sq_fwd_dense: need B, N >= 1 and M % 4 == 0");
    if (!aligned(theta, 16) || !aligned(A, 16) || (Q && !aligned(Q, 16)))
        return fail(-1, "b200dp_sq_fwd_dense: theta, A and Q must be 16-byte aligned");
    CUtensorMap tmT, tmA;
    if (!encode_row_map(&tmT, theta, B, N, M, kG) || !encode_row_map(&tmA, A, B, N, M, kG))
        return b200dp_sq_fwd(tab, nstrips, workspace, theta, A, Q, Vt, mode, flags, stream);      // no TMA maps: LDGSTS staging
    SqParams p = sq_params(tab, nstrips, workspace);
    p.dbg = (flags >> B200DP_SQ_DBG_SHIFT) & 0xF;
    p.trace = static_cast<unsigned long long*>(g_trace);
    p.theta = theta;
    p.A = A;
    p.Q = Q;
    p.Vt = Vt;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const bool sw = mode == B200DP_MODE_SW;
    const int ring = ring_of(flags);
#define B200DP_SQFT(SW_, STORE_, RING_)                                                                                  \
    return sq_launch(softdp_sq_fwd_tma_kernel<SW_, false, STORE_, RING_>, sq_fwd_smem_bytes<false, RING_>(), p, flags, st, \
                     "b200dp_sq_fwd_dense", tmT, tmA, tmT)
    if (!Q) {
        if (sw) B200DP_SQFT(true, false, 4);
        B200DP_SQFT(false, false, 4);
    }
    if (ring == 3) {
        if (sw) B200DP_SQFT(true, true, 3);
        B200DP_SQFT(false, true, 3);
    }
    if (ring == 6) {
        if (sw) B200DP_SQFT(true, true, 6);
        B200DP_SQFT(false, true, 6);
    }
    if (sw) B200DP_SQFT(true, true, 4);
    B200DP_SQFT(false, true, 4);
#undef B200DP_SQFT
}

int b200dp_sq_bwd(const void* tab, int nstrips, void* workspace, const float* Et, long long et_stride,
                  const float* Q, float* E, int mode, int flags, void* stream) {
    if (int rc = sq_check("b200dp_sq_bwd", tab, nstrips, workspace)) return rc;
    if (mode != B200DP_MODE_NW && mode != B200DP_MODE_SW) return fail(-1, "b200dp_sq_bwd: bad mode");
    if (nstrips == 0) return 0;
    if (!Et || !Q || !E) return fail(-1, "b200dp_sq_bwd: null pointer");
    if (!aligned(Q, 16)) return fail(-1, "b200dp_sq_bwd: Q storage must be 16-byte aligned");
    SqParams p = sq_params(tab, nstrips, workspace);
    p.dbg = (flags >> B200DP_SQ_DBG_SHIFT) & 0xF;
    p.trace = static_cast<unsigned long long*>(g_trace);
    p.Et = Et;
    p.et_stride = et_stride;
    p.Qin = Q;
    p.Eout = E;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const bool sw = mode == B200DP_MODE_SW;
    const int ring = ring_of(flags);
#define B200DP_SQB(SW_, RING_) \
    return sq_launch(softdp_sq_bwd_kernel<SW_, false, RING_>, sq_bwd_smem_bytes<RING_, false>(), p, flags, st, "b200dp_sq_bwd")
    // default: two Q tiles per warp (18.6 KB of shared memory, 12 warps per SM).  Measured on B200: a third
    // tile buys nothing once the ring covers one bulk-TMA round trip, the extra resident warps do
    // (1024 x 256^2: 0.169 ms against 0.177 ms with three tiles and 9 warps per SM; 1024 x 512^2: 0.639 / 0.674)
    if (ring == 3) {
        if (sw) B200DP_SQB(true, 3);
        B200DP_SQB(false, 3);
    }
    if (ring == 4) {
        if (sw) B200DP_SQB(true, 4);
        B200DP_SQB(false, 4);
    }
    if (ring == 6) {
        if (sw) B200DP_SQB(true, 6);
        B200DP_SQB(false, 6);
    }
    if (sw) B200DP_SQB(true, 2);
    B200DP_SQB(false, 2);
#undef B200DP_SQB
}

int b200dp_sq_adj_fwd(const void* tab, int nstrips, void* workspace, const float* Q, const float* Zt,
                      const float* ZA, const float* E, float* Vtd, float* QdE, int flags, void* stream) {
    if (int rc = sq_check("b200dp_sq_adj_fwd", tab, nstrips, workspace)) return rc;
    if (nstrips == 0) return 0;
    if (!Q || !Zt || !Vtd || !QdE) return fail(-1, "b200dp_sq_adj_fwd: null pointer");
    if (!aligned(Q, 16) || !aligned(QdE, 16) || !aligned(Zt, 16) || (ZA && !aligned(ZA, 16)) || (E && !aligned(E, 16)))
        return fail(-1, "b200dp_sq_adj_fwd: pointers must be 16-byte aligned");
    SqParams p = sq_params(tab, nstrips, workspace);
    p.dbg = (flags >> B200DP_SQ_DBG_SHIFT) & 0xF;
    p.trace = static_cast<unsigned long long*>(g_trace);
    p.theta = Zt;
    p.A = ZA;
    p.E = E;
    p.Qin = Q;
    p.Q = QdE;
    p.Vt = Vtd;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    return sq_launch(softdp_sq_fwd_kernel<false, true, true, 3>, sq_fwd_smem_bytes<true, 3>(), p, flags, st,
                     "b200dp_sq_adj_fwd");
}

int b200dp_sq_adj_bwd(const void* tab, int nstrips, void* workspace, const float* Q, const float* QdE,
                      float* Ed, int flags, void* stream) {
    if (int rc = sq_check("b200dp_sq_adj_bwd", tab, nstrips, workspace)) return rc;
    if (nstrips == 0) return 0;
    if (!Q || !QdE || !Ed) return fail(-1, "b200dp_sq_adj_bwd: null pointer");
    if (!aligned(Q, 16) || !aligned(QdE, 16)) return fail(-1, "b200dp_sq_adj_bwd: Q / QdE storage must be 16-byte aligned");
    SqParams p = sq_params(tab, nstrips, workspace);
    p.dbg = (flags >> B200DP_SQ_DBG_SHIFT) & 0xF;
    p.trace = static_cast<unsigned long long*>(g_trace);
    p.Qin = Q;
    p.QdE = QdE;
    p.Eout = Ed;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (ring_of(flags) == 3)
        return sq_launch(softdp_sq_bwd_kernel<false, true, 3>, sq_bwd_smem_bytes<3, true>(), p, flags, st, "b200dp_sq_adj_bwd");
    return sq_launch(softdp_sq_bwd_kernel<false, true, 2>, sq_bwd_smem_bytes<2, true>(), p, flags, st, "b200dp_sq_adj_bwd");
}

}  // extern "C"
