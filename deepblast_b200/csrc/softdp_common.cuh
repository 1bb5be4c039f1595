// softdp_common.cuh -- shared device helpers for the sm_100a soft-DP kernels.
//
// Data layout in HBM:
//   theta, A, ZA : [B, N, M] fp32 row-major (reference layout, deepblast/nw.py:65-117)
//   E, Ed, Ztheta: [B, N+2, M+2] fp32 row-major (reference layout, nw.py:347)
//   Q, Qd        : the reference's [B, N+2, M+2, 3] (nw.py:105) stored STRIP-MAJOR in
//                  the order the wavefront produces it, TWO of the three states per cell:
//                  the x and y components; the m component is implied (Q sums to 1 over
//                  the states, Qd to 0: q_m = 1 - q_x - q_y, qd_m = -(qd_x + qd_y)), which
//                  takes a third off the Q traffic of every sweep.  A cell whose Q is all
//                  zero (first row / column of the sw.py lattice, sw.py:54-55) carries the
//                  mark q_x = kQZeroMark (< 0, not a probability).
//                  Lattice cell (i, j), 1-based, belongs to strip k = (i-1)/32, lane
//                  t = (i-1)%32 and is touched at wavefront step sigma = (j-1) + t of that
//                  strip:  elem(b,i,j,c) = b*pair_stride + k*strip_stride + sigma*64 + c*32 + t
//                    (c = 0: x, c = 1: y)
//                    strip_stride = M*64, pair_stride = ceil(N/32)*strip_stride + 31*64
//                  (chained-dense: the 31 ramp steps of consecutive strips interleave
//                  lane-wise, no step of a pair's storage is padding).
//                  One step of a strip is 256 contiguous bytes (2 states x 32 lanes) and
//                  consecutive steps are contiguous, so the forward streams Q out
//                  sequentially and the backward streams it back in with 1-D bulk TMA.
//                  Border cells (zeros, Q[N+1,M+1,:]=1) are implicit, never stored.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#ifdef B200DP_DEBUG_WAIT
#include <cstdio>
#endif

namespace b200dp {

constexpr unsigned kFull = 0xffffffffu;
constexpr int kTile = 32;            // rows per strip == lanes; row-major tile is 32 x 32
constexpr int kTileElems = kTile * kTile;
constexpr int kRowRing = 3;          // theta/A (row-major, skewed read): 2 live tiles + 1 in flight
constexpr int kStepFloats = 64;      // Q floats per wavefront step of a strip: 2 stored states x 32 lanes
constexpr int kQY = 32;              // offset of the y component inside a step (x at 0)
constexpr float kQZeroMark = -1.0f;  // q_x of a cell whose Q is identically zero (sw.py first row / column)
constexpr int kDiagRows = 16;        // wavefront steps per Q tile
constexpr int kDiagElems = kDiagRows * kStepFloats;
constexpr int kDiagRing = 4;         // Q tiles: 1 live + 3 in flight
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---- mbarrier ------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// A wait that cannot complete is a protocol bug.  Release builds wait on wall-clock time, not
// on a poll count (a legitimate wait under preemption, MPS time-slicing, a debugger or ncu
// replay may take many polls): the launch is only failed (trap: the host sees an error instead
// of a hung device) after kWaitLimitNs of real time; -DB200DP_NO_WAIT_LIMIT compiles the bound out.
constexpr unsigned long long kWaitLimitNs = 8000000000ull;      // 8 s
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long v;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(v));
    return v;
}
struct WaitGuard {
    unsigned polls = 0;
    unsigned long long t0 = 0;
    // call once per failed poll; cheap until the wait has lasted a few thousand polls
    __device__ __forceinline__ void tick() {
#ifndef B200DP_NO_WAIT_LIMIT
        if ((++polls & 0x3ffu) == 0) {
            const unsigned long long now = global_ns();
            if (t0 == 0) t0 = now;
            else if (now - t0 > kWaitLimitNs) __trap();
        }
#endif
    }
};
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    WaitGuard g;
    while (!mbar_try_wait(bar, parity)) g.tick();
}

// One lane of a converged warp, known to the compiler to be exactly one (elect.sync):
// TMA issue under this predicate needs no per-lane serialisation loop.
__device__ __forceinline__ bool elect_one() {
    uint32_t p;
    asm volatile(
        "{\n\t.reg .pred P;\n\t"
        "elect.sync _|P, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t}"
        : "=r"(p));
    return p != 0;
}

// ---- TMA (cp.async.bulk.tensor) -------------------------------------------
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1,
                                            int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_u32(dst)),
        "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1,
                                            int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(smem_u32(dst)),
        "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
// 1-D bulk copy global -> shared (size multiple of 16 B, both addresses 16-B aligned)
__device__ __forceinline__ void tma_bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
// L2 prefetch of a tensor-map box (no shared-memory destination, no completion)
__device__ __forceinline__ void tma_prefetch_l2_3d(const CUtensorMap* map, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];" ::"l"(map), "r"(c0), "r"(c1),
                 "r"(c2)
                 : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

// ---- cp.async (generic fallback when a tensor is not TMA-addressable) -------
__device__ __forceinline__ void cp_async4_zfill(void* dst, const void* src, bool valid) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(smem_u32(dst)), "l"(src),
                 "r"(valid ? 4 : 0)
                 : "memory");
}
__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
// 16-byte copy with a 256-byte L2 prefetch hint; !valid zero-fills the destination (nothing is read)
__device__ __forceinline__ void cp_async16_zfill_l2(void* dst, const void* src, bool valid) {
    asm volatile("cp.async.cg.shared.global.L2::256B [%0], [%1], 16, %2;" ::"r"(smem_u32(dst)), "l"(src),
                 "r"(valid ? 16 : 0)
                 : "memory");
}
__device__ __forceinline__ void cp_async16_zfill(void* dst, const void* src, bool valid) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst)), "l"(src), "r"(valid ? 16 : 0)
                 : "memory");
}
__device__ __forceinline__ void cp_async_mbar_arrive_noinc(uint64_t* bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---- cross-warp progress words (release/acquire at CTA scope) ------------------
__device__ __forceinline__ void st_release_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.cta.shared.u64 [%0], %1;" ::"r"(smem_u32(p)), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.cta.shared.u64 %0, [%1];" : "=l"(v) : "r"(smem_u32(p)) : "memory");
    return v;
}

// ---- cross-CTA hand-off words (coherent at GPU scope, i.e. served by L2; no fences: the
// payload and its validity tag travel in ONE 8-byte word, see softdp_sq.cuh) -------------------
__device__ __forceinline__ void st_relaxed_gpu_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// the same hand-off as a reduction performed AT the L2 (never parked in an SM-side write buffer):
// tags grow from launch to launch, so max() replaces whatever an earlier launch left behind
__device__ __forceinline__ void red_max_gpu_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("red.relaxed.gpu.global.max.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_gpu_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// ---- fast math -----------------------------------------------------------------
__device__ __forceinline__ float fast_ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float fast_lg2(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float fast_rcp(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// 1/S for S in [1, 3] on the FMA pipe: cubic minimax seed in (S - 2) (relative error
// 1.0e-2) and two Newton steps (1.6e-7, i.e. the accuracy of rcp.approx).  The XU pipe
// (ex2 / lg2 / rcp) is the forward kernel's bottleneck on B200, the FMA pipe is not.
__device__ __forceinline__ float rcp_1to3(float S) {
    const float x = S - 2.0f;
    float r = fmaf(-0.08242285f, x, 0.16481542f);
    r = fmaf(r, x, -0.24747011f);
    r = fmaf(r, x, 0.4949553f);
    r = r * fmaf(-S, r, 2.0f);
    r = r * fmaf(-S, r, 2.0f);
    return r;
}

// ---- persistent strip iterator ---------------------------------------------------
// A CTA owns pairs blockIdx.x, blockIdx.x + gridDim.x, ...; every pair is cut into
// strips of 32 rows.  All strips of all the CTA's pairs form ONE linear sequence
// q = 0, 1, 2, ...; warp w executes the strips with q % W == w in increasing q.
// Strip q hands its boundary row to strip q+1 (when both belong to the same pair)
// through shared memory slot q % (W+1).  Because strip q+W+1 is executed by the
// same warp as strip q+1 and after it, a slot is never overwritten while read.
struct Strip {
    int pair;     // batch index
    int k;        // strip ordinal inside the pair in PROCESSING order (0 = first processed)
    int K;        // strips in the pair
    int n, m;     // lattice of the pair
    int round;    // how many pairs this CTA has been dealt
    unsigned q;   // linear sequence number inside the CTA
    bool valid;
};

struct PairDims {
    const int* xlen;
    const int* ylen;
    int B, N, M;
};

__device__ __forceinline__ void strip_load_pair(Strip& s, const PairDims& d) {
    // lengths are clamped to the tensor like the reference's slices theta[b, :n, :m]
    // (deepblast/alignment.py:166-169): an over-long length must not run past the pair's storage
    s.n = d.xlen ? min(max(d.xlen[s.pair], 0), d.N) : d.N;
    s.m = d.ylen ? min(max(d.ylen[s.pair], 0), d.M) : d.M;
    s.K = (s.n > 0 && s.m > 0) ? ((s.n + kTile - 1) / kTile) : 0;
}
__device__ __forceinline__ void strip_seek(Strip& s, const PairDims& d, int w, int W) {
    for (;;) {
        if (s.pair >= d.B) {
            s.valid = false;
            return;
        }
        if (s.k >= s.K) {
            // next round, boustrophedon: rounds alternate direction so that a batch sorted
            // by descending work is dealt evenly to the CTAs
            s.round++;
            const int lane = (s.round & 1) ? (int)(gridDim.x - 1 - blockIdx.x) : (int)blockIdx.x;
            s.pair = s.round * (int)gridDim.x + lane;
            s.k = 0;
            if (s.pair < d.B) strip_load_pair(s, d);
            continue;
        }
        if ((int)(s.q & (unsigned)(W - 1)) == w) {
            s.valid = true;
            return;
        }
        s.k++;
        s.q++;
    }
}
__device__ __forceinline__ void strip_first(Strip& s, const PairDims& d, int w, int W) {
    s.pair = blockIdx.x;
    s.round = 0;
    s.k = 0;
    s.q = 0;
    s.K = 0;
    s.n = s.m = 0;
    s.valid = false;
    if (s.pair < d.B) strip_load_pair(s, d);
    strip_seek(s, d, w, W);
}
__device__ __forceinline__ void strip_next(Strip& s, const PairDims& d, int w, int W) {
    s.k++;
    s.q++;
    strip_seek(s, d, w, W);
}

// wait until boundary `q` has published at least `need` entries
__device__ __forceinline__ int progress_wait(const unsigned long long* word, unsigned q, int need) {
#ifdef B200DP_DEBUG_WAIT
    unsigned spins = 0;
#endif
    WaitGuard g;
    for (;;) {
        unsigned long long v = ld_acquire_u64(word);
        if ((unsigned)(v >> 32) == q && (int)(unsigned)v >= need) return (int)(unsigned)v;
        __nanosleep(32);
#ifdef B200DP_DEBUG_WAIT
        if (++spins > (1u << 18)) {
            if ((threadIdx.x & 31) == 0)
                printf("progress_wait stuck: cta %d warp %d wants q %u need %d, word q %u count %d\n", (int)blockIdx.x,
                       (int)(threadIdx.x >> 5), q, need, (unsigned)(v >> 32), (int)(unsigned)v);
            return need;
        }
#else
        g.tick();
#endif
    }
}

// ---- run-ahead gate ------------------------------------------------------------------
// Boundary slot q % (W+1) is reused by strip q + (W+1).  Inside one pair the hand-off chain
// itself keeps a writer behind the previous reader of its slot, but the first strip of a
// pair waits for nobody, so across pair boundaries (ragged batches, many short pairs per
// CTA) a warp could run several strips ahead of its left neighbour and overwrite a row --
// and its progress word -- that a slow warp is still reading.  Gate: strip q starts only
// after strip q - (W+1), the previous strip of warp w-1, has finished; by induction every
// earlier user of the slot has finished too.  fin[w] = 1 + the last strip warp w finished.
__device__ __forceinline__ void strip_gate(const unsigned long long* fin, unsigned q, int w, int W) {
    if (W > 1 && q >= (unsigned)(W + 1)) {
        const unsigned long long* f = fin + ((w + W - 1) % W);
        const unsigned long long want = (unsigned long long)q - (unsigned)W;      // 1 + (q - (W+1))
        WaitGuard g;
        while (ld_acquire_u64(f) < want) {
            __nanosleep(32);
            g.tick();
        }
    }
}
__device__ __forceinline__ void strip_done(unsigned long long* fin, unsigned q, int w, int W) {
    if (W > 1) {
        __syncwarp();
        if ((threadIdx.x & 31) == 0) st_release_u64(fin + w, (unsigned long long)q + 1ull);
    }
}

struct QLayout {
    long long pair_stride;    // floats per pair  = K * strip_stride
    long long strip_stride;   // floats per strip = M * 64 (consecutive strips overlap by 31 ramp steps)
    int K;                    // strips per pair  = ceil(N / 32)
};

}  // namespace b200dp
