// softdp_cl.cuh -- CLUSTER kernels: forward fill and backward sweep for SMALL batches of LONG,
// equal-size lattices (reference: deepblast/nw.py:46-62,120-135 and the sw.py twins; replaces
// deepblast/nw_cuda.py:46-102 where the reference actually trains and infers: a few dozen pairs of up
// to 1024 x 1024, deepblast/trainer.py:375, or one pair at a time, alignment.py:165-169).
//
// Such a batch is bound by ONE pair's dependency chain: strip k+1 (32 rows) can only follow strip k
// by the 31 steps of lane skew plus the hand-off of the boundary row.  The strip-queue kernels
// (softdp_sq.cuh) hand that row over through L2 -- a store, its visibility, one or two polling round
// trips: about 2 us, 20-35 wavefront steps per strip, 60 % of the whole sweep at 32 x 1024^2.  Here the
// strips of one pair are dealt round-robin to the CTAs of one thread-block CLUSTER (one warp per CTA,
// up to 8 CTAs on 8 SMs) and the row travels through DISTRIBUTED SHARED MEMORY:
//   * the producer lane stores each boundary value straight into the consumer CTA's shared memory as ONE
//     8-byte word {tag = use count of the row buffer, value} (st.relaxed.cluster.shared::cluster.b64):
//     payload and flag cannot be seen apart, so no fence and no mbarrier sits on the producer's path
//     (a release at cluster scope every 8 steps was measured: it waits for the warp's outstanding Q
//     stores and makes the sweep 2.5 times slower);
//   * the consumer polls its OWN shared memory (30 cycles a try) 8 entries at a time, twice per 16-step
//     block: the hop costs a few hundred cycles instead of the few thousand of the L2 round trips;
//   * two boundary rows per CTA (strips alternate) and a consumed-strips counter written back into the
//     producer's shared memory give flow control when a producer could otherwise lap its consumer
//     (pairs with fewer strips than the cluster has CTAs).
// Same cell arithmetic (fwd2_step, difference form, log2 units), same strip-major two-state Q, same
// operand staging (16 x 16 TMA boxes, bulk-TMA Q tiles), same E staging and row-major drain as the
// strip-queue kernels; a static schedule instead of tickets (cluster c takes pairs c, c + #clusters, ...).
#pragma once
#include "softdp_sq.cuh"

namespace b200dp {

struct ClParams {
    // forward: theta, A -> Q (or null: score only), Vt          backward: Et, Qin -> Eout (interior [B, N, M])
    const float* theta;
    const float* A;
    float* Q;
    float* Vt;
    const float* Et;
    long long et_stride;
    const float* Qin;
    float* Eout;
    int B, N, M;
};

// ---- cluster primitives ------------------------------------------------------------------------
__device__ __forceinline__ unsigned cl_rank() {
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ unsigned cl_size() {
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cl_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `local` (a shared::cta address of this CTA) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t cl_map(const void* local, unsigned rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(local)), "r"(rank));
    return r;
}
__device__ __forceinline__ void cl_st_relaxed_u64(uint32_t addr, unsigned long long v) {
    asm volatile("st.relaxed.cluster.shared::cluster.b64 [%0], %1;" ::"r"(addr), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long cl_ld_relaxed_u64(const unsigned long long* local) {
    unsigned long long v;
    asm volatile("ld.relaxed.cluster.shared::cta.b64 %0, [%1];" : "=l"(v) : "r"(smem_u32(local)) : "memory");
    return v;
}
__device__ __forceinline__ void cl_st_release_u32(uint32_t addr, unsigned v) {
    asm volatile("st.release.cluster.shared::cluster.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned cl_ld_acquire_u32(const unsigned* local) {
    unsigned v;
    asm volatile("ld.acquire.cluster.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(local)) : "memory");
    return v;
}

constexpr int kClGroup = 8;          // boundary values taken at a time (two takes per 16-step block)

// shared memory of one CTA: [operand / Q ring | E staging (bwd)] [ring mbarriers] [consumed counter]
// [2 boundary rows of Mp tagged 8-byte words] [bv: 16 floats] [16 zeros]
__host__ __device__ inline size_t cl_tail_bytes(int M) {
    const size_t Mp = ((size_t)M + 15) & ~(size_t)15;
    return 16 + 2 * Mp * 8 + 64 + 64;
}
// (per WARP; a CTA of W warps takes W times as much, every warp's region 1 KB aligned)
template <int RING>
__host__ __device__ inline size_t cl_fwd_smem_bytes(int M) {
    size_t b = (size_t)RING * 4096 + (size_t)RING * 8;
    b = (b + 15) & ~(size_t)15;
    return (b + cl_tail_bytes(M) + 1023) & ~(size_t)1023;
}
template <int RING>
__host__ __device__ inline size_t cl_bwd_smem_bytes(int M) {
    size_t b = (size_t)RING * kDiagElems * 4 + kB3StageBytes + (size_t)RING * 8;
    b = (b + 15) & ~(size_t)15;
    return (b + cl_tail_bytes(M) + 1023) & ~(size_t)1023;
}

// Everything the two kernels share about the hand-off: the CTA's own (consumer side) rows and counter,
// and the mapped addresses of the next CTA's rows (I produce for it) and the previous CTA's counter (I
// consume from it).
struct ClLink {
    unsigned* consumed;          // local: strips of MINE that my consumer has finished (written remotely by it)
    unsigned long long* rows;    // [2][Mp] local boundary rows of tagged words
    float* bv;                   // [16] the current block's values
    float* zero_row;             // [16]
    uint32_t r_rows;             // the consumer CTA's rows
    uint32_t r_consumed;         // the producer CTA's counter
    int Mp;
    unsigned n_in, n_out;        // strips with an input row consumed / with an output row produced so far

    // `base`: the tail of THIS warp's region; `warp_bytes`: size of one warp's region (all CTAs and warps
    // have the same layout, so a neighbour's tail is at the same offset of its region)
    __device__ __forceinline__ void init(unsigned char* base, int M, int t, int w, int W, size_t warp_bytes) {
        Mp = (M + 15) & ~15;
        consumed = reinterpret_cast<unsigned*>(base);
        rows = reinterpret_cast<unsigned long long*>(base + 16);
        bv = reinterpret_cast<float*>(base + 16 + (size_t)2 * Mp * 8);
        zero_row = bv + 16;
        for (int e = t; e < 2 * Mp; e += 32) rows[e] = 0ull;          // tag 0: never a valid use count
        if (t == 0) *consumed = 0u;
        if (t < 16) {
            bv[t] = 0.f;
            zero_row[t] = 0.f;
        }
        // the warps of a cluster form ONE ring: warp w of CTA r is position r W + w; its consumer is the next
        // position (the next warp of the CTA, or warp 0 of the next CTA), its producer the previous one
        const int C = (int)cl_size(), rank = (int)cl_rank();
        const int nw = (w + 1 == W) ? 0 : w + 1, nr = (w + 1 == W) ? ((rank + 1 == C) ? 0 : rank + 1) : rank;
        const int pw = (w == 0) ? W - 1 : w - 1, pr = (w == 0) ? ((rank == 0) ? C - 1 : rank - 1) : rank;
        r_rows = cl_map(reinterpret_cast<unsigned char*>(rows) + (ptrdiff_t)(nw - w) * (ptrdiff_t)warp_bytes, (unsigned)nr);
        r_consumed = cl_map(reinterpret_cast<unsigned char*>(consumed) + (ptrdiff_t)(pw - w) * (ptrdiff_t)warp_bytes, (unsigned)pr);
        n_in = n_out = 0;
    }
    // consumer: entries [s0 + lo, s0 + hi) of the current input row -> bv[lo..hi) (lanes lo..hi-1 poll their
    // own word in LOCAL shared memory until it carries this use's tag)
    __device__ __forceinline__ void take(int s0, int lo, int hi, int m, int t) const {
        const bool mine = t >= lo && t < hi && s0 + t < m;
        const unsigned long long* src = rows + (n_in & 1) * Mp + s0 + (mine ? t : 0);
        const unsigned want = n_in + 1u;
        unsigned long long w = 0;
        bool ok = !mine;
        if (mine) {
            w = cl_ld_relaxed_u64(src);
            ok = (unsigned)(w >> 32) == want;
        }
        if (!__all_sync(kFull, ok)) {
            WaitGuard g;
            for (;;) {
                if (!ok) {
                    w = cl_ld_relaxed_u64(src);
                    ok = (unsigned)(w >> 32) == want;
                }
                if (__all_sync(kFull, ok)) break;
                g.tick();
            }
        }
        if (t >= lo && t < hi) bv[t] = mine ? __uint_as_float((unsigned)w) : 0.f;
        __syncwarp();
    }
    // consumer: the whole input row has been read (once per strip with an input)
    __device__ __forceinline__ void release_in(int t) {
        n_in++;
        if (t == 0) cl_st_release_u32(r_consumed, n_in);
    }
    // producer: before the first value of output row n_out is stored, the row that used the same buffer
    // (n_out - 2) must have been consumed
    __device__ __forceinline__ void acquire_out() const {
        if (n_out >= 2) {
            WaitGuard g;
            while (cl_ld_acquire_u32(consumed) + 2u <= n_out) g.tick();
        }
    }
    // producer lane: value e of the current output row
    __device__ __forceinline__ void publish(int e, float v) const {
        cl_st_relaxed_u64(r_rows + (uint32_t)((n_out & 1) * Mp + e) * 8u,
                          ((unsigned long long)(n_out + 1u) << 32) | (unsigned long long)__float_as_uint(v));
    }
};

// ---------------------------------------------------------------------------------------------
template <bool SWM, bool STOREQ, int RING>
__global__ void __launch_bounds__(128) softdp_cl_fwd_kernel(const ClParams p, const __grid_constant__ CUtensorMap tm_theta,
                                                            const __grid_constant__ CUtensorMap tm_A) {
    constexpr int kGroupBytes = 2048, kSlot = 4096;
    constexpr int FDBG = STOREQ ? 0 : 4;
    extern __shared__ __align__(1024) unsigned char smem_all[];
    const int t = threadIdx.x & 31, w = threadIdx.x >> 5, W = blockDim.x >> 5, g = t >> 4, tp = t & 15;
    const int N = p.N, M = p.M, K = (N + 31) >> 5;
    const int C = (int)cl_size(), rank = (int)cl_rank();
    const int cid = (int)blockIdx.x / C, ncl = (int)gridDim.x / C;
    const int VW = C * W, vpos = rank * W + w;           // ring of warps: size, this warp's position
    const size_t warp_bytes = cl_fwd_smem_bytes<RING>(M);
    unsigned char* smem_raw = smem_all + (size_t)w * warp_bytes;

    unsigned char* ring = smem_raw;
    size_t off = (size_t)RING * kSlot;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + off);
    off = (off + (size_t)RING * 8 + 15) & ~(size_t)15;
    ClLink L;
    L.init(smem_raw + off, M, t, w, W, warp_bytes);
    const float* zero_row = L.zero_row;
    if (t == 0) {
        for (int s = 0; s < RING; ++s) mbar_init(&bars[s], 1);
        tma_prefetch_desc(&tm_theta);
        tma_prefetch_desc(&tm_A);
    }
    fence_mbar_init();
    __syncwarp();
    cl_sync();                                     // every CTA's rows are cleared before anybody stores into them

    const int T16 = (M + 15) >> 4, NE = T16 + 1, NBk = (M + 31 + 15) >> 4;
    const long long SS = (long long)M * kStepFloats, PS = (long long)K * SS + 31ll * kStepFloats;
    const int lanebase = g * kGroupBytes + tp * 60;

    // strip sequence of this warp: (pair, i) with i = vpos, vpos + VW, ... < K, pair = cid, cid + ncl, ... < B
    auto valid = [&](int pair, int i) { return pair < p.B && i < K; };
    auto advance = [&](int& pair, int& i) {
        i += VW;
        if (i >= K) {
            i = vpos;
            pair += ncl;
        }
    };
    auto issue = [&](int pair, int i, int e, unsigned slot) {
        unsigned bytes = 0;
#pragma unroll
        for (int gg = 0; gg < 2; ++gg)
            if (e - gg >= 0 && e - gg < T16) bytes += 2048u;
        if (elect_one()) {
            mbar_expect_tx(&bars[slot], bytes);
#pragma unroll
            for (int gg = 0; gg < 2; ++gg) {
                const int tile = e - gg;
                if (tile >= 0 && tile < T16) {
                    unsigned char* d = ring + slot * kSlot + gg * kGroupBytes;
                    const int row0 = i * kTile + gg * kG;
                    tma_load_3d(d, &tm_theta, &bars[slot], tile * kG, row0, pair);
                    tma_load_3d(d + 1024, &tm_A, &bars[slot], tile * kG, row0, pair);
                }
            }
        }
    };

    int pair = cid, i = vpos;
    int npair = pair, ni = i;
    advance(npair, ni);
    int issued = 0;
    unsigned islot = 0, wslot = 0, phases = 0;

    while (valid(pair, i)) {
        const int rows = min(kTile, N - i * kTile);
        const bool first = i == 0, last = i == K - 1;
        const bool has_up = i > 0, feeds_down = i + 1 < K;
        const bool row_ok = t < rows;
        const bool rowcomp = row_ok && !(SWM && first && t == 0);      // sw.py: i >= 2
        const bool full_rows = rows == kTile;
        const bool nxt_ok = valid(npair, ni);
        float* qp = p.Q + (long long)pair * PS + (long long)i * SS + t;
        if (feeds_down) L.acquire_out();

        float v = 0.f, h = 0.f;                       // differences, log2 units
        float acc_hi = 0.f, acc_lo = 0.f;             // sum_j h[i, j] of the lane's row
        unsigned slotA = 0, slotB = 0;

        for (int b = 0; b < NBk; ++b) {
            __syncwarp();
            while (issued <= b + RING - 2) {
                if (issued < NE) issue(pair, i, issued, islot);
                else if (nxt_ok && issued - NE <= T16) issue(npair, ni, issued - NE, islot);
                else break;
                issued++;
                islot = (islot + 1 == RING) ? 0u : islot + 1;
            }
            slotA = slotB;
            if (b < NE) {
                mbar_wait(&bars[wslot], (phases >> wslot) & 1u);
                phases ^= 1u << wslot;
                slotB = wslot;
                wslot = (wslot + 1 == RING) ? 0u : wslot + 1;
            }
            const int s0 = b * 16;
            // ---- the row above: values stored into this CTA's shared memory by the producer CTA ----
            const bool take = has_up && s0 < M;
            const float* br = zero_row;
            if (take) {
                L.take(s0, 0, kClGroup, M, t);
                br = L.bv;
            }
            const unsigned char* sA = ring + slotA * kSlot + lanebase + 64;
            const unsigned char* sB = ring + slotB * kSlot + lanebase;
            const bool steady = full_rows && b >= 2 && s0 + 16 <= M;
            float part = 0.f;
            // one unrolled, operand-hoisted body for steady (EDGE = false) and ramp blocks (see softdp_sq.cuh)
            auto block = [&](auto edge_tag) {
                constexpr bool EDGE = decltype(edge_tag)::value;
                float th_[16], a_[16], bv_[16];
#pragma unroll
                for (int ss = 0; ss < 16; ++ss) {
                    const float* tb = reinterpret_cast<const float*>((tp <= ss) ? sB : sA);
                    th_[ss] = tb[ss];
                    a_[ss] = tb[ss + 256];
                }
                auto load_bv = [&](int half) {
#pragma unroll
                    for (int q4 = 2 * half; q4 < 2 * half + 2; ++q4) {
                        const float4 b4 = reinterpret_cast<const float4*>(br)[q4];
                        bv_[4 * q4] = b4.x;
                        bv_[4 * q4 + 1] = b4.y;
                        bv_[4 * q4 + 2] = b4.z;
                        bv_[4 * q4 + 3] = b4.w;
                    }
                };
                load_bv(0);
                if (!take) load_bv(1);
                const bool live = !(SWM && first && t == 0);      // sw.py: row 1 is below the origin
                const bool cap = EDGE && last && (((M - 1 + (rows - 1)) >> 4) == b);
                const int c0 = s0 - t;
#pragma unroll
                for (int ss = 0; ss < 16; ++ss) {
                    if (ss == 8 && take) {
                        L.take(s0, kClGroup, 16, M, t);
                        load_bv(1);
                    }
                    float hup = __shfl_up_sync(kFull, h, 1);
                    hup = (t == 0) ? bv_[ss] : hup;
                    if (!EDGE) {
                        h = fwd2_step<false, SWM, FDBG>(th_[ss], a_[ss], hup, v, qp + ss * kStepFloats, true, live);
                        part += h;
                        if (t == 31 && feeds_down) L.publish(c0 + ss, h);
                    } else {
                        const int c = c0 + ss;
                        const bool in = row_ok && (unsigned)c < (unsigned)M;
                        const bool comp = SWM ? (in && rowcomp && c >= 1) : in;      // sw.py: j >= 2
                        h = fwd2_step<true, SWM, FDBG>(th_[ss], a_[ss], hup, v, qp + ss * kStepFloats, in, comp);
                        part += h;
                        if (t == 31 && feeds_down && in) L.publish(c, h);
                        // Vt = V[n, m] = ln 2 * sum_j h[n, j]
                        if (cap && in && t == rows - 1 && c == M - 1) p.Vt[pair] = (acc_hi + (acc_lo + part)) * kLn2;
                    }
                }
                qp += 16 * kStepFloats;
            };
            if (steady) block(std::false_type{});
            else block(std::true_type{});
            {
                const float t1 = part + acc_lo;       // Fast2Sum fold of the block's partial row sum
                const float nh = acc_hi + t1;
                acc_lo = t1 - (nh - acc_hi);
                acc_hi = nh;
            }
        }
        __syncwarp();
        if (has_up) L.release_in(t);
        if (feeds_down) L.n_out++;
        issued -= NE;
        pair = npair;
        i = ni;
        advance(npair, ni);
    }
    cl_sync();                                     // nobody leaves while a neighbour may still store into it
}

// ---------------------------------------------------------------------------------------------
// Backward sweep: strips bottom-up (sequence index i = K-1-k), push form, right to left, lane 31 leads.
template <bool SWM, int RING>
__global__ void __launch_bounds__(128) softdp_cl_bwd_kernel(const ClParams p) {
    extern __shared__ __align__(1024) unsigned char smem_all[];
    const int t = threadIdx.x & 31, w = threadIdx.x >> 5, W = blockDim.x >> 5, u = 31 - t;
    const int N = p.N, M = p.M, K = (N + 31) >> 5;
    const int C = (int)cl_size(), rank = (int)cl_rank();
    const int cid = (int)blockIdx.x / C, ncl = (int)gridDim.x / C;
    const int VW = C * W, vpos = rank * W + w;
    const size_t warp_bytes = cl_bwd_smem_bytes<RING>(M);
    unsigned char* smem_raw = smem_all + (size_t)w * warp_bytes;

    float* qring = reinterpret_cast<float*>(smem_raw);
    float* stage = qring + RING * kDiagElems;
    size_t off = (size_t)RING * kDiagElems * 4 + kB3StageBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + off);
    off = (off + (size_t)RING * 8 + 15) & ~(size_t)15;
    ClLink L;
    L.init(smem_raw + off, M, t, w, W, warp_bytes);
    const float* zero_row = L.zero_row;
    if (t == 0)
        for (int s = 0; s < RING; ++s) mbar_init(&bars[s], 1);
    fence_mbar_init();
    __syncwarp();
    cl_sync();

    const int NBk = (M + 31 + 15) >> 4;           // blocks == Q tiles: sweep steps 0 .. M+30
    const long long SS = (long long)M * kStepFloats, PS = (long long)K * SS + 31ll * kStepFloats;
    auto valid = [&](int pair, int i) { return pair < p.B && i < K; };
    auto advance = [&](int& pair, int& i) {
        i += VW;
        if (i >= K) {
            i = vpos;
            pair += ncl;
        }
    };
    // tile a of a strip covers sweep steps [16a, 16a+16) = wavefront steps [M+15-16a, M+30-16a]
    auto issue = [&](int pair, int i, int a, unsigned slot) {
        const float* strip = p.Qin + (long long)pair * PS + (long long)(K - 1 - i) * SS;
        q_tile_load<true>(qring + slot * kDiagElems, &bars[slot], strip, M + 15 - kDiagRows * a, t, 1);
    };

    int pair = cid, i = vpos;
    int npair = pair, ni = i;
    advance(npair, ni);
    int issued = 0;
    unsigned islot = 0, wslot = 0, phases = 0;

    while (valid(pair, i)) {
        const int k = K - 1 - i;
        const int rows = min(kTile, N - k * kTile);
        const bool top = k == 0, bottom = k == K - 1;
        const bool has_below = i > 0, feeds_up = i + 1 < K;
        const bool row_ok = t < rows;
        const bool rowcomp = row_ok && !(SWM && top && t == 0);       // sw.py: i >= 2
        const bool full_rows = rows == kTile;
        const bool sw_dead = SWM && top && t == 0;
        const bool nxt_ok = valid(npair, ni);
        float* Erow0 = p.Eout + ((long long)pair * N + (long long)k * kTile) * M;
        const float et = p.Et[(long long)pair * p.et_stride];
        if (feeds_up) L.acquire_out();

        float zout = 0.f, dprev = 0.f, yprev = 0.f;
        int next_drain = (M - 1) >> 5;                // highest column tile not yet drained
        int srow = 0;                                 // (16 b) mod kB2StageSteps

        for (int b = 0; b < NBk; ++b) {
            __syncwarp();
            while (issued <= b + RING - 1) {
                if (issued < NBk) issue(pair, i, issued, islot);
                else if (nxt_ok && issued - NBk < NBk) issue(npair, ni, issued - NBk, islot);
                else break;
                issued++;
                islot = (islot + 1 == RING) ? 0u : islot + 1;
            }
            mbar_wait(&bars[wslot], (phases >> wslot) & 1u);
            phases ^= 1u << wslot;
            const float* qt = qring + wslot * kDiagElems + t;
            wslot = (wslot + 1 == RING) ? 0u : wslot + 1;
            const int s0 = b * 16;
            while (next_drain >= 0 && (M + 30 - 32 * next_drain) < s0) {
                if (full_rows && next_drain * kTile + kTile <= M) sq_drain_tile<true>(stage, Erow0, next_drain, M, rows, M, t);
                else sq_drain_tile<false>(stage, Erow0, next_drain, M, rows, M, t);
                next_drain--;
            }
            __syncwarp();
            // ---- the row below: entry e = sweep step at which lane 31 needs it (column M-1-e) ----
            const bool take = has_below && s0 < M;
            const float* br = zero_row;
            if (take) {
                L.take(s0, 0, kClGroup, M, t);
                br = L.bv;
            }
            float* st = stage + srow * kB2StagePitch + t;
            const bool steady = full_rows && s0 >= 32 && s0 + 15 <= M - 1 - (SWM ? 1 : 0);
            if (steady) {
                float bv_[16];
                auto load_bv = [&](int half) {
#pragma unroll
                    for (int q4 = 2 * half; q4 < 2 * half + 2; ++q4) {
                        const float4 b4 = reinterpret_cast<const float4*>(br)[q4];
                        bv_[4 * q4] = b4.x;
                        bv_[4 * q4 + 1] = b4.y;
                        bv_[4 * q4 + 2] = b4.z;
                        bv_[4 * q4 + 3] = b4.w;
                    }
                };
                load_bv(0);
                if (!take) load_bv(1);
                float qx_[16], qy_[16], qm_[16];
#pragma unroll
                for (int ss = 0; ss < 16; ++ss) {
                    qx_[ss] = qt[(15 - ss) * kStepFloats];
                    qy_[ss] = qt[(15 - ss) * kStepFloats + kQY];
                }
#pragma unroll
                for (int ss = 0; ss < 16; ++ss) qm_[ss] = (1.f - qx_[ss]) - qy_[ss];      // >= 0 by the forward's clamp
#pragma unroll
                for (int ss = 0; ss < 16; ++ss) {
                    if (ss == 8 && take) {
                        L.take(s0, kClGroup, 16, M, t);
                        load_bv(1);
                    }
                    float zin = __shfl_down_sync(kFull, zout, 1);
                    if (t == 31) zin = bv_[ss];
                    float e = zin + yprev;
                    if (SWM) e = sw_dead ? 0.f : e;           // row 1: E = 0, nothing pushed
                    const float X = qx_[ss] * e, Y = qy_[ss] * e, D = qm_[ss] * e;
                    st[ss * kB2StagePitch] = e;
                    zout = X + dprev;
                    dprev = D;
                    yprev = Y;
                    if (t == 0 && feeds_up) L.publish(s0 - 31 + ss, zout);
                }
            } else {
                // ramp blocks.  Q in the ramps was never written by the forward (arbitrary bits):
                // products are selected, not multiplied by zero.
                const bool seed_blk = bottom && s0 < 32;      // E[n, m] = Et lives here
#pragma unroll 4
                for (int ss = 0; ss < 16; ++ss) {
                    if (ss == 8 && take) L.take(s0, kClGroup, 16, M, t);
                    const int c = M - 1 - (s0 + ss - u);
                    float zin = __shfl_down_sync(kFull, zout, 1);
                    if (t == 31) zin = br[ss];
                    const bool in = row_ok && (unsigned)c < (unsigned)M;
                    const bool comp = SWM ? (in && rowcomp && c >= 1) : in;
                    float e = zin + yprev;
                    if (seed_blk && t == rows - 1 && c == M - 1) e = et;   // nw.py:125-127
                    e = comp ? e : 0.f;
                    const float qx = qt[(15 - ss) * kStepFloats], qy = qt[(15 - ss) * kStepFloats + kQY];
                    const float X = comp ? qx * e : 0.f;
                    const float Y = comp ? qy * e : 0.f;
                    const float D = comp ? ((1.f - qx) - qy) * e : 0.f;
                    st[ss * kB2StagePitch] = e;
                    zout = X + dprev;
                    dprev = D;
                    yprev = Y;
                    // lane 0 publishes the value the strip above needs at ITS sweep step M-1-c
                    if (t == 0 && feeds_up && in) L.publish(M - 1 - c, zout);
                }
            }
            srow += 16;
            if (srow == kB2StageSteps) srow = 0;
        }
        __syncwarp();
        while (next_drain >= 0) {
            if (full_rows && next_drain * kTile + kTile <= M) sq_drain_tile<true>(stage, Erow0, next_drain, M, rows, M, t);
            else sq_drain_tile<false>(stage, Erow0, next_drain, M, rows, M, t);
            next_drain--;
        }
        __syncwarp();
        if (has_below) L.release_in(t);
        if (feeds_up) L.n_out++;
        issued -= NBk;
        pair = npair;
        i = ni;
        advance(npair, ni);
    }
    cl_sync();
}

}  // namespace b200dp
