// softdp_sq.cuh -- STRIP-QUEUE kernels: forward fill, backward sweep and the adjoint pair for
// batches of ANY shape -- ragged (per-pair lengths), packed, small batches of long pairs,
// large batches of equal pairs (reference: deepblast/nw.py:46-62,120-135,178-199,251-267 and
// the sw.py twins; replaces deepblast/nw_cuda.py:46-165; the ragged semantics are those of the
// per-pair loop of deepblast/alignment.py:165-169).
//
// Same cell arithmetic (softdp_fwd2.cuh / softdp_fwd3.cuh / softdp_bwd3.cuh: difference
// form, log2 units, two stored Q states) and the same strip-major Q as the other fast
// kernels; what is new is how strips find warps:
//   * WORK QUEUE.  The host cuts every pair into strips of 32 rows and lists them in ONE
//     table of 64-byte records (StripRec) in an order in which every strip comes after the
//     strip it depends on (the strip above it in the forward direction, below it in the
//     backward direction).  Each CTA is ONE warp; warps take tickets from an atomic counter
//     and execute the strip of their ticket, so a warp only ever waits for a strip whose
//     ticket is smaller, i.e. one that a running warp already holds: no dead-lock whatever
//     the grid size, any number of warps on one long pair (a 1024 x 1024 lattice is 32
//     strips pipelined 48 steps apart on 32 warps of 32 different SMs), and the load balances
//     itself on ragged batches.  Resident warps per SM are bounded by shared memory
//     (12-23 KB per warp), not by the batch size.
//   * HAND-OFF THROUGH L2, WITHOUT FENCES.  Strip k hands its bottom boundary row (one float
//     per column) to strip k+1 through a scratch row in global memory.  Every entry is ONE
//     8-byte word {tag = launch epoch, value}: the producer lane stores it with a single
//     relaxed 64-bit store at GPU scope as soon as the value exists, the consumer loads the
//     16 entries of its next block one block ahead (relaxed, GPU scope: L2) and re-polls only
//     entries whose tag is not the epoch yet.  Payload and flag cannot be seen apart, so no
//     release/acquire fence sits on the critical path of either warp, and stale entries of
//     earlier launches (older epochs) are never mistaken for data.
//   * OPERANDS OF ANY PITCH.  theta / A (/ Ztheta / E) tiles are staged with 16-byte
//     cp.async (LDGSTS, 256-byte L2 prefetch hint) completing on the slot's mbarrier: each
//     quad of lanes fetches the 64-byte piece of one row, so a warp instruction covers 8
//     rows x 64 B and the only requirement is a row pitch that is a multiple of 4 floats --
//     dense [B, N, M] tensors and PACKED buffers (pair b at an offset of its own, pitch
//     roundup(m_b, 4)) run through the same code.  Q streams stay 1-D bulk TMA (4 KB tiles).
//   * E is written straight into the caller's [.., n, m] interior layout (dense or packed):
//     no padded border, no zero fill.
#pragma once
#include "softdp_bwd3.cuh"
#include "softdp_fwd3.cuh"

namespace b200dp {

#ifndef B200DP_SQ_REARM
#define B200DP_SQ_REARM 3
#endif
constexpr int kSqRearmStep = B200DP_SQ_REARM;   // step of a backward block at which the block's own Q slot is re-armed (its Q is in registers by then)
constexpr int kSqPfStep = 11;    // step of a 16-step block at which the next block's boundary entries are fetched
constexpr int kSqFirst = 1;      // StripRec.flags: strip 0 of its pair (holds lattice row 1)
constexpr int kSqLast = 2;       //                 last strip of its pair (holds lattice row n)

// One ticket of the work queue (built on the host by b200dp_plan_build, softdp_api.cu).
struct alignas(16) StripRec {
    long long t_off;     // element offset of (first row of the strip, column 0) in theta / A / E / ...
    long long q_off;     // float offset of wavefront step 0 of the strip in the strip-major Q storage
    long long b_in;      // boundary scratch (8-byte words) this strip consumes, -1 = none (zero border)
    long long b_out;     // boundary scratch this strip publishes, -1 = none
    int rows;            // rows of the strip inside the lattice, 1..32
    int m;               // columns of the pair
    int pitch;           // row pitch of theta / A / E in floats (multiple of 4)
    int pair;            // batch index (Vt, Et)
    int flags;           // kSqFirst | kSqLast
    int k;               // strip ordinal inside the pair (0 = rows 1..32)
    int pad0, pad1;
};
static_assert(sizeof(StripRec) == 64, "StripRec is 64 bytes");

struct SqParams {
    const StripRec* tab;
    int nstrips;
    unsigned long long* ctl;     // ctl[0]: ticket counter (low 32 bits) | exited warps (high 32 bits), self-resetting;
                                 // ctl[1]: launches done on this workspace -- the tag of a launch's boundary words
                                 // is ctl[1] + 1 (never 0, so a zero-filled workspace holds no valid word); the last
                                 // warp to leave bumps it, so the tag lives on the DEVICE and a launch captured in a
                                 // CUDA graph gets a fresh tag at every replay
    unsigned long long* bnd;     // boundary scratch
    // forward:          theta, A -> Q, Vt
    // adjoint forward:  theta = Ztheta (interior layout), A = ZA or null, E or null, Qin -> Q = Qd * E, Vt = Vtd
    const float* theta;
    const float* A;
    const float* E;
    const float* Qin;
    float* Q;
    float* Vt;
    // backward:         Et, Qin -> Eout            adjoint backward:  Qin, QdE -> Eout (= Ed)
    const float* Et;
    long long et_stride;
    const float* QdE;
    float* Eout;
    unsigned long long* trace;   // diagnostics: per ticket {start, end} in globaltimer ns, or null
    int dbg;                     // diagnostics (B200DP_SQ_DBG_SHIFT; results are wrong when set): 1 = no L2 prefetch
                                 // hint on the operand copies, 2 = no operand staging, 4 = no boundary exchange
};

// (measured: a reduction performed at the L2, red.max on the tagged word, is no faster than the
// plain relaxed store -- the hop latency is not in the SM's store path)
__device__ __forceinline__ void sq_publish(unsigned long long* p, unsigned long long v, int) {
    st_relaxed_gpu_u64(p, v);
}
__device__ __forceinline__ unsigned long long sq_pack(unsigned epoch, float v) {
    return ((unsigned long long)epoch << 32) | (unsigned long long)__float_as_uint(v);
}

__device__ __forceinline__ StripRec sq_load_rec(const StripRec* tab, int tk, int nstrips) {
    StripRec r;
    if (tk < nstrips) {
        const int4* s = reinterpret_cast<const int4*>(tab + tk);
        int4* d = reinterpret_cast<int4*>(&r);
        d[0] = __ldg(s);
        d[1] = __ldg(s + 1);
        d[2] = __ldg(s + 2);
        d[3] = __ldg(s + 3);
    } else {
        r.rows = 0;
        r.m = 0;
        r.t_off = r.q_off = 0;
        r.b_in = r.b_out = -1;
        r.pitch = 4;
        r.pair = 0;
        r.flags = 0;
        r.k = 0;
    }
    return r;
}

// A warp leaves: count it, and the last one to leave resets the control word for the next launch
// (every pull of every warp precedes its own exit on the same address, so nothing is pending).
__device__ __forceinline__ void sq_exit(unsigned long long* ctl, unsigned epoch) {
    if ((threadIdx.x & 31) == 0) {
        const unsigned long long old = atomicAdd(ctl, 1ull << 32);
        if ((unsigned)(old >> 32) == gridDim.x - 1) {
            ctl[1] = epoch;                      // every warp of this launch has read it (they all started)
            atomicExch(ctl, 0ull);
        }
    }
}
__device__ __forceinline__ unsigned sq_epoch(const unsigned long long* ctl) {
    return (unsigned)ld_relaxed_gpu_u64(ctl + 1) + 1u;
}

// Boundary entries of one block (16 columns): `pfv` was loaded one block ahead by lanes 0..15;
// re-poll the entries whose tag is not this launch's yet, then park the 16 values in `bv`.
__device__ __forceinline__ void sq_bnd_take(unsigned long long pfv, bool need, const unsigned long long* src,
                                            unsigned epoch, float* bv, int t) {
    bool ok = !need || (unsigned)(pfv >> 32) == epoch;
    if (!__all_sync(kFull, ok)) {
        // trailing the producer closely: poll back to back (one L2 round trip per try, only the
        // lanes whose entry is still missing), the hop latency is on the pair's critical path
        WaitGuard g;
        for (;;) {
            if (!ok) {
                pfv = ld_relaxed_gpu_u64(src);
                ok = (unsigned)(pfv >> 32) == epoch;
            }
            if (__all_sync(kFull, ok)) break;
            g.tick();
        }
    }
    if (t < 16) bv[t] = need ? __uint_as_float((unsigned)pfv) : 0.f;
    __syncwarp();
}

template <bool ADJ, int RING>
__host__ __device__ inline size_t sq_fwd_smem_bytes() {
    size_t b = (size_t)RING * (ADJ ? 6144 : 4096);               // [RING][2 groups][theta, A (, E)][16][16]
    if (ADJ) b += (size_t)3 * kDiagElems * 4 + 3 * 8;            // Q tiles
    b += (size_t)RING * 8;
    b = (b + 15) & ~(size_t)15;
    b += 128;                                                    // bv[16], zero[16]
    return b;
}

// ---------------------------------------------------------------------------------------------
// Forward fill (ADJ = false; STOREQ = false: score only, Vt alone) and adjoint forward (ADJ).
// TMAOP: the operands are DENSE [B, N, M] tensors described by tensor maps -- one elected lane issues
// 16 x 16 TMA boxes (same shared-memory image as the LDGSTS path) instead of every lane issuing four
// 16-byte copies per tensor and block with their address arithmetic.  Measured on B200 at 1024 x 256^2:
// 49 instead of 67 instructions per wavefront step and 64 instead of 128 registers; the time is the
// same (0.228 ms) -- the sweep is bound by the latency of its dependent steps under the mixed
// read / write traffic, not by issue slots.
template <bool SWM, bool ADJ, bool STOREQ, int RING, bool TMAOP>
__device__ __forceinline__ void sq_fwd_body(const SqParams& p, const CUtensorMap* tm_theta, const CUtensorMap* tm_A,
                                            const CUtensorMap* tm_E) {
    static_assert(!(ADJ && SWM), "the adjoint sweeps cover the full range (sw.py:150-151)");
    constexpr int kGroupBytes = ADJ ? 3072 : 2048;
    constexpr int kSlot = 2 * kGroupBytes;
    constexpr int QR = 3;
    constexpr int FDBG = STOREQ ? 0 : 4;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const int t = threadIdx.x, g = t >> 4, tp = t & 15, r8 = t >> 2, ch = t & 3;
    const unsigned epoch = sq_epoch(p.ctl);

    unsigned char* ring = smem_raw;
    float* qring = reinterpret_cast<float*>(smem_raw + (size_t)RING * kSlot);
    size_t off = (size_t)RING * kSlot + (ADJ ? (size_t)QR * kDiagElems * 4 : 0);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + off);
    uint64_t* qbars = bars + RING;
    off += (size_t)(RING + (ADJ ? QR : 0)) * 8;
    off = (off + 15) & ~(size_t)15;
    float* bv = reinterpret_cast<float*>(smem_raw + off);
    float* zero_row = bv + 16;

    if (t == 0) {
        for (int s = 0; s < RING; ++s) mbar_init(&bars[s], TMAOP ? 1 : 32);
        if (ADJ)
            for (int s = 0; s < QR; ++s) mbar_init(&qbars[s], 1);
        if (TMAOP) {
            tma_prefetch_desc(tm_theta);
            if (tm_A) tma_prefetch_desc(tm_A);
            if (tm_E) tma_prefetch_desc(tm_E);
        }
    }
    if (t < 16) {
        bv[t] = 0.f;
        zero_row[t] = 0.f;
    }
    fence_mbar_init();
    __syncthreads();

    const int lanebase = g * kGroupBytes + tp * 60;      // group, row, -4*tp column skew (bytes)
    const bool has_a = !ADJ || p.A != nullptr;
    const bool has_e = ADJ && p.E != nullptr;

    // event e of a strip = {group 0 (rows 0..15): tile e, group 1 (rows 16..31): tile e-1} x tensors
    auto issue = [&](const StripRec& st, int e, unsigned slot) {
        if (p.dbg & 2) return;
        const int T16 = (st.m + 15) >> 4;
        if (TMAOP) {
            const unsigned per = 1024u * (1u + (has_a ? 1u : 0u) + (has_e ? 1u : 0u));
            unsigned bytes = 0;
#pragma unroll
            for (int gg = 0; gg < 2; ++gg)
                if (e - gg >= 0 && e - gg < T16) bytes += per;
            if (elect_one()) {
                mbar_expect_tx(&bars[slot], bytes);
#pragma unroll
                for (int gg = 0; gg < 2; ++gg) {
                    const int tile = e - gg;
                    if (tile >= 0 && tile < T16) {
                        unsigned char* d = ring + slot * kSlot + gg * kGroupBytes;
                        const int row0 = st.k * kTile + gg * kG;
                        tma_load_3d(d, tm_theta, &bars[slot], tile * kG, row0, st.pair);
                        if (has_a) tma_load_3d(d + 1024, tm_A, &bars[slot], tile * kG, row0, st.pair);
                        if (has_e) tma_load_3d(d + 2048, tm_E, &bars[slot], tile * kG, row0, st.pair);
                    }
                }
            }
            return;
        }
        unsigned char* dst = ring + slot * kSlot + r8 * 64 + ch * 16;
#pragma unroll
        for (int gg = 0; gg < 2; ++gg) {
            const int tile = e - gg;
            if (tile >= 0 && tile < T16) {
                const int col = tile * 16 + ch * 4;
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const int row = gg * 16 + i * 8 + r8;
                    const bool ok = row < st.rows && col < st.pitch;
                    const long long o = st.t_off + (ok ? (long long)row * st.pitch + col : 0ll);
                    unsigned char* d = dst + gg * kGroupBytes + i * 512;
                    if (p.dbg & 1) {
                        cp_async16_zfill(d, p.theta + o, ok);
                        if (has_a) cp_async16_zfill(d + 1024, p.A + o, ok);
                        if (has_e) cp_async16_zfill(d + 2048, p.E + o, ok);
                    } else {
                        cp_async16_zfill_l2(d, p.theta + o, ok);
                        if (has_a) cp_async16_zfill_l2(d + 1024, p.A + o, ok);
                        if (has_e) cp_async16_zfill_l2(d + 2048, p.E + o, ok);
                    }
                }
            }
        }
        cp_async_mbar_arrive_noinc(&bars[slot]);
    };

    // ---- tickets: all from the counter, so that a warp only ever waits for strips that RUNNING
    // warps hold (no assumption that the whole grid is resident) -----------------------------
    int cur_tk = 0, nxt_tk = 0;
    {
        unsigned long long first = 0;
        if (t == 0) first = atomicAdd(p.ctl, 1ull);
        cur_tk = (int)(unsigned)__shfl_sync(kFull, (unsigned)first, 0);
    }
    StripRec cur = sq_load_rec(p.tab, cur_tk, p.nstrips);
    StripRec nxt;
    nxt.rows = 0;
    nxt.m = 0;
    bool nxt_ready = false;
    int issued = 0, qissued = 0;
    unsigned islot = 0, wslot = 0, phases = 0;
    unsigned qislot = 0, qwslot = 0, qphases = 0;

    while (cur.rows > 0) {
        const int m = cur.m, rows = cur.rows;
        const int T16 = (m + 15) >> 4, NE = T16 + 1, NBk = (m + 31 + 15) >> 4;
        const bool first = (cur.flags & kSqFirst) != 0, last = (cur.flags & kSqLast) != 0;
        const bool has_up = cur.b_in >= 0 && !(p.dbg & 4), feeds_down = cur.b_out >= 0 && !(p.dbg & 4);
        const bool row_ok = t < rows;
        const bool rowcomp = row_ok && !(SWM && first && t == 0);      // sw.py: i >= 2
        const bool full_rows = rows == kTile;
        const unsigned long long* bin = p.bnd + (has_up ? cur.b_in : 0);
        unsigned long long* bout = p.bnd + (feeds_down ? cur.b_out : 0);
        float* qp = p.Q + cur.q_off + t;
        const float* qstrip = ADJ ? p.Qin + cur.q_off : nullptr;
        if (p.trace && t == 0) p.trace[2 * cur_tk] = global_ns();

        // The next ticket is taken late -- three blocks before this strip ends, its record read
        // one block later: enough to hide the atomic, the record load and the first operand
        // tiles, while a claimed strip never sits unstarted for more than ~50 steps (strips differ
        // in length by 10x on ragged batches; a strip claimed at the start of a long one would
        // hold up its whole dependency chain)
        unsigned long long pulled = 0;
        const int b_pull = NBk > 3 ? NBk - 3 : 0, b_load = b_pull + 1;
        // the first 16 entries of the row above
        unsigned long long pfv = 0;
        if (has_up && t < 16 && t < m) pfv = ld_relaxed_gpu_u64(bin + t);

        float v = 0.f, h = 0.f;                       // differences, log2 units
        float acc_hi = 0.f, acc_lo = 0.f;             // sum_j h[i, j] of the lane's row
        unsigned slotA = 0, slotB = 0;

        for (int b = 0; b < NBk; ++b) {
            __syncwarp();
            if (b == b_pull && t == 0) pulled = atomicAdd(p.ctl, 1ull);
            if (b == b_load) {
                nxt_tk = (int)(unsigned)__shfl_sync(kFull, (unsigned)pulled, 0);
                nxt = sq_load_rec(p.tab, nxt_tk, p.nstrips);
                nxt_ready = true;
            }
            // ---- producer: operand events up to b + RING - 2 (running into the next strip) ----
            while (issued <= b + RING - 2) {
                if (issued < NE) issue(cur, issued, islot);
                else if (nxt_ready && nxt.rows > 0 && issued - NE <= ((nxt.m + 15) >> 4)) issue(nxt, issued - NE, islot);
                else break;
                issued++;
                islot = (islot + 1 == RING) ? 0u : islot + 1;
            }
            const float* qt = nullptr;
            if (ADJ) {
                while (qissued <= b + QR - 1) {
                    if (qissued < NBk)
                        q_tile_load<true>(qring + qislot * kDiagElems, &qbars[qislot], qstrip, kDiagRows * qissued, t);
                    else if (nxt_ready && nxt.rows > 0 && qissued - NBk < ((nxt.m + 31 + 15) >> 4))
                        q_tile_load<true>(qring + qislot * kDiagElems, &qbars[qislot], p.Qin + nxt.q_off,
                                          kDiagRows * (qissued - NBk), t);
                    else break;
                    qissued++;
                    qislot = (qislot + 1 == QR) ? 0u : qislot + 1;
                }
            }
            slotA = slotB;
            if (b < NE) {
                if (!(p.dbg & 2)) mbar_wait(&bars[wslot], (phases >> wslot) & 1u);
                phases ^= 1u << wslot;
                slotB = wslot;
                wslot = (wslot + 1 == RING) ? 0u : wslot + 1;
            }
            if (ADJ) {
                mbar_wait(&qbars[qwslot], (qphases >> qwslot) & 1u);
                qphases ^= 1u << qwslot;
                qt = qring + qwslot * kDiagElems + t;
                qwslot = (qwslot + 1 == QR) ? 0u : qwslot + 1;
            }
            const int s0 = b * 16;
            // ---- the row above: 16 tagged words from L2, loaded one block ahead ----------------
            const float* br = zero_row;
            if (has_up && s0 < m) {
                sq_bnd_take(pfv, t < 16 && s0 + t < m, bin + s0 + t, epoch, bv, t);
                br = bv;
            }
            // the next block's entries are fetched kSqPfStep steps into this block: early enough to
            // hide an L2 round trip, late enough that a strip trailing its predecessor closely finds
            // them there (fetched at the start of the block the stable distance would be 16 steps more)
            const int nx = s0 + 16 + t;
            const bool pf_next = has_up && t < 16 && nx < m;
            const unsigned char* sA = ring + slotA * kSlot + lanebase + 64;
            const unsigned char* sB = ring + slotB * kSlot + lanebase;
            const bool steady = full_rows && b >= 2 && s0 + 16 <= m;
            float part = 0.f;
            // One unrolled body for both kinds of block, operands hoisted into registers: EDGE = false when every
            // lane is inside the lattice for the 16 steps (no predicates), EDGE = true for the ramp blocks at
            // both ends of a strip and for partial strips (the same step with the lattice-membership selects;
            // the 31 + 8 steps by which a strip trails its predecessor are ramp steps, so their speed sets
            // the dependency stagger of a long pair).
            auto block = [&](auto edge_tag) {
                constexpr bool EDGE = decltype(edge_tag)::value;
                float th_[16], a_[16], bv_[16];
                float e_[ADJ ? 16 : 1], qx_[ADJ ? 16 : 1], qy_[ADJ ? 16 : 1];
#pragma unroll
                for (int ss = 0; ss < 16; ++ss) {
                    const float* tb = reinterpret_cast<const float*>((tp <= ss) ? sB : sA);
                    th_[ss] = tb[ss];
                    a_[ss] = has_a ? tb[ss + 256] : 0.f;
                    if (ADJ) {
                        e_[ss] = has_e ? tb[ss + 512] : 1.f;
                        qx_[ss] = qt[ss * kStepFloats];
                        qy_[ss] = qt[ss * kStepFloats + kQY];
                    }
                }
#pragma unroll
                for (int q4 = 0; q4 < 4; ++q4) {
                    const float4 b4 = reinterpret_cast<const float4*>(br)[q4];
                    bv_[4 * q4] = b4.x;
                    bv_[4 * q4 + 1] = b4.y;
                    bv_[4 * q4 + 2] = b4.z;
                    bv_[4 * q4 + 3] = b4.w;
                }
                const bool live = !(SWM && first && t == 0);      // sw.py: row 1 is below the origin
                const bool cap = EDGE && last && (((m - 1 + (rows - 1)) >> 4) == b);
                const int c0 = s0 - t;                            // the lane's column at step 0 of the block
#pragma unroll
                for (int ss = 0; ss < 16; ++ss) {
                    float hup = __shfl_up_sync(kFull, h, 1);
                    hup = (t == 0) ? bv_[ss] : hup;
                    if (!EDGE) {
                        if (ADJ)
                            h = adj3_step<false>(th_[ss], a_[ss], e_[ss], qx_[ss], qy_[ss], hup, v, qp + ss * kStepFloats,
                                                 true, true);
                        else
                            h = fwd2_step<false, SWM, FDBG>(th_[ss], a_[ss], hup, v, qp + ss * kStepFloats, true, live);
                        part += h;
                        if (t == 31 && feeds_down) sq_publish(bout + (c0 + ss), sq_pack(epoch, h), p.dbg);
                    } else {
                        const int c = c0 + ss;
                        const bool in = row_ok && (unsigned)c < (unsigned)m;
                        const bool comp = SWM ? (in && rowcomp && c >= 1) : in;      // sw.py: j >= 2
                        if (ADJ)
                            h = adj3_step<true>(th_[ss], a_[ss], e_[ss], in ? qx_[ss] : 0.f, in ? qy_[ss] : 0.f, hup, v,
                                                qp + ss * kStepFloats, in, comp);
                        else
                            h = fwd2_step<true, SWM, FDBG>(th_[ss], a_[ss], hup, v, qp + ss * kStepFloats, in, comp);
                        part += h;
                        if (t == 31 && feeds_down && in) sq_publish(bout + c, sq_pack(epoch, h), p.dbg);
                        // Vt = V[n, m] = ln 2 * sum_j h[n, j]   (adjoint: Vtd = sum_j hd[n, j])
                        if (cap && in && t == rows - 1 && c == m - 1)
                            p.Vt[cur.pair] = (acc_hi + (acc_lo + part)) * (ADJ ? 1.f : kLn2);
                    }
                    if (ss == kSqPfStep && pf_next) pfv = ld_relaxed_gpu_u64(bin + nx);
                }
                qp += 16 * kStepFloats;
            };
            if (steady) block(std::false_type{});
            else block(std::true_type{});
            {
                // fold the block's partial row sum into the two-float accumulator (Fast2Sum)
                const float t1 = part + acc_lo;
                const float nh = acc_hi + t1;
                acc_lo = t1 - (nh - acc_hi);
                acc_hi = nh;
            }
        }
        if (p.trace && t == 0) p.trace[2 * cur_tk + 1] = global_ns();
        issued -= NE;
        if (ADJ) qissued -= NBk;
        cur = nxt;
        cur_tk = nxt_tk;
        nxt_ready = false;
        nxt.rows = 0;
    }
    sq_exit(p.ctl, epoch);
}

template <bool SWM, bool ADJ, bool STOREQ, int RING>
__global__ void __launch_bounds__(32) softdp_sq_fwd_kernel(const SqParams p) {
    sq_fwd_body<SWM, ADJ, STOREQ, RING, false>(p, nullptr, nullptr, nullptr);
}
template <bool SWM, bool ADJ, bool STOREQ, int RING>
__global__ void __launch_bounds__(32) softdp_sq_fwd_tma_kernel(const SqParams p, const __grid_constant__ CUtensorMap tm_theta,
                                                               const __grid_constant__ CUtensorMap tm_A,
                                                               const __grid_constant__ CUtensorMap tm_E) {
    sq_fwd_body<SWM, ADJ, STOREQ, RING, true>(p, &tm_theta, p.A ? &tm_A : nullptr, p.E ? &tm_E : nullptr);
}

// ---------------------------------------------------------------------------------------------
// Backward sweep (ADJ = false) and adjoint backward (ADJ): push form, right to left, lane 31 leads.
template <int RING, bool ADJ>
__host__ __device__ inline size_t sq_bwd_smem_bytes() {
    size_t b = (size_t)RING * kDiagElems * (ADJ ? 2 : 1) * 4 + kB3StageBytes;
    b += (size_t)RING * 8;
    b = (b + 15) & ~(size_t)15;
    b += 128;                                      // bv[16], zero[16]
    return b;
}

// Drain column tile tc (32 columns) of the current strip from the step-major staging ring to
// row-major E: element (r, col) was produced at sweep step (m-1-col) + (31-r), i.e. staging row
// (m + 30 - 32 tc - t - r) mod 80 for lane t = column 32 tc + t (the strip's step 0 sits in
// staging row 0).  FULL: all 32 rows and all 32 columns exist.
template <bool FULL>
__device__ __forceinline__ void sq_drain_tile(const float* __restrict__ stage, float* __restrict__ Erow0, int tc, int m,
                                              int rows, int pitch, int t) {
    const int col = tc * kTile + t;
    int sr0 = (m + 30 - tc * kTile) % kB2StageSteps - t;
    sr0 += (sr0 < 0) ? kB2StageSteps : 0;
    const float* p0 = stage + sr0 * kB2StagePitch;
    const float* p1 = p0 + kB2StageFloats;
    float* dstp = Erow0 + col;
    const bool colok = FULL || col < m;
#pragma unroll
    for (int r0 = 0; r0 < kTile; r0 += 8) {
        float v[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int r = r0 + q;
            const float* ps = (r > sr0) ? p1 : p0;
            v[q] = ps[-(kB2StagePitch - 1) * r];
        }
#pragma unroll
        for (int q = 0; q < 8; ++q)
            if (FULL || (colok && r0 + q < rows)) dstp[(long long)(r0 + q) * pitch] = v[q];
    }
}

template <bool SWM, bool ADJ, int RING>
__global__ void __launch_bounds__(32) softdp_sq_bwd_kernel(const SqParams p) {
    static_assert(!(ADJ && SWM), "the adjoint sweeps cover the full range (sw.py:199-201)");
    constexpr int kSlotElems = kDiagElems * (ADJ ? 2 : 1);      // [Q tile | QdE tile]
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const int t = threadIdx.x, u = 31 - t;
    const unsigned epoch = sq_epoch(p.ctl);

    float* qring = reinterpret_cast<float*>(smem_raw);
    float* stage = qring + RING * kSlotElems;
    size_t off = (size_t)RING * kSlotElems * 4 + kB3StageBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + off);
    off = (off + (size_t)RING * 8 + 15) & ~(size_t)15;
    float* bv = reinterpret_cast<float*>(smem_raw + off);
    float* zero_row = bv + 16;

    if (t == 0)
        for (int s = 0; s < RING; ++s) mbar_init(&bars[s], 1);
    if (t < 16) {
        bv[t] = 0.f;
        zero_row[t] = 0.f;
    }
    fence_mbar_init();
    __syncthreads();

    // tile a of a strip covers sweep steps [16a, 16a+16) = wavefront steps [m+15-16a, m+30-16a]
    auto issue = [&](const StripRec& st, int a, unsigned slot) {
        q_tile_load<true>(qring + slot * kSlotElems, &bars[slot], p.Qin + st.q_off, st.m + 15 - kDiagRows * a, t,
                          ADJ ? 2 : 1);
        if (ADJ)
            q_tile_load<true>(qring + slot * kSlotElems + kDiagElems, &bars[slot], p.QdE + st.q_off,
                              st.m + 15 - kDiagRows * a, t, 0);
    };

    int cur_tk = 0, nxt_tk = 0;
    {
        unsigned long long first = 0;
        if (t == 0) first = atomicAdd(p.ctl, 1ull);
        cur_tk = (int)(unsigned)__shfl_sync(kFull, (unsigned)first, 0);
    }
    StripRec cur = sq_load_rec(p.tab, cur_tk, p.nstrips);
    StripRec nxt;
    nxt.rows = 0;
    nxt.m = 0;
    bool nxt_ready = false;
    int issued = 0;
    unsigned islot = 0, wslot = 0, phases = 0;

    while (cur.rows > 0) {
        const int m = cur.m, rows = cur.rows;
        const int NBk = (m + 31 + 15) >> 4;           // blocks == Q tiles: sweep steps 0 .. m+30
        const bool top = (cur.flags & kSqFirst) != 0, bottom = (cur.flags & kSqLast) != 0;
        const bool has_below = cur.b_in >= 0 && !(p.dbg & 4), feeds_up = cur.b_out >= 0 && !(p.dbg & 4);
        const bool row_ok = t < rows;
        const bool rowcomp = row_ok && !(SWM && top && t == 0);       // sw.py: i >= 2
        const bool full_rows = rows == kTile;
        const bool sw_dead = SWM && top && t == 0;
        const unsigned long long* bin = p.bnd + (has_below ? cur.b_in : 0);
        unsigned long long* bout = p.bnd + (feeds_up ? cur.b_out : 0);
        float* Erow0 = p.Eout + cur.t_off;
        const float et = ADJ ? 0.f : p.Et[(long long)cur.pair * p.et_stride];
        if (p.trace && t == 0) p.trace[2 * cur_tk] = global_ns();

        unsigned long long pulled = 0;                // next ticket: taken late, see the forward kernel
        const int b_pull = NBk > 3 ? NBk - 3 : 0, b_load = b_pull + 1;
        // lane l < 16 fetches the entry lane 31 needs at step l: column m-1-l
        unsigned long long pfv = 0;
        if (has_below && t < 16 && m - 1 - t >= 0) pfv = ld_relaxed_gpu_u64(bin + (m - 1 - t));

        float zout = 0.f, dprev = 0.f, yprev = 0.f;
        int next_drain = (m - 1) >> 5;                // highest column tile not yet drained
        int srow = 0;                                 // (16 b) mod kB2StageSteps

        for (int b = 0; b < NBk; ++b) {
            __syncwarp();
            if (b == b_pull && t == 0) pulled = atomicAdd(p.ctl, 1ull);
            if (b == b_load) {
                nxt_tk = (int)(unsigned)__shfl_sync(kFull, (unsigned)pulled, 0);
                nxt = sq_load_rec(p.tab, nxt_tk, p.nstrips);
                nxt_ready = true;
            }
            // tiles up to `limit` go out (this strip's, then the next strip's first ones)
            auto pump = [&](int limit) {
                while (issued <= limit) {
                    if (issued < NBk) issue(cur, issued, islot);
                    else if (nxt_ready && nxt.rows > 0 && issued - NBk < ((nxt.m + 31 + 15) >> 4)) issue(nxt, issued - NBk, islot);
                    else break;
                    issued++;
                    islot = (islot + 1 == RING) ? 0u : islot + 1;
                }
            };
            pump(b + RING - 1);
            mbar_wait(&bars[wslot], (phases >> wslot) & 1u);
            phases ^= 1u << wslot;
            const float* qt = qring + wslot * kSlotElems + t;
            wslot = (wslot + 1 == RING) ? 0u : wslot + 1;
            const int s0 = b * 16;
            // drain every column tile completed before this block: tile tc is complete once
            // lane 0 has passed column 32 tc, i.e. after step m + 30 - 32 tc
            while (next_drain >= 0 && (m + 30 - 32 * next_drain) < s0) {
                if (full_rows && next_drain * kTile + kTile <= m) sq_drain_tile<true>(stage, Erow0, next_drain, m, rows, cur.pitch, t);
                else sq_drain_tile<false>(stage, Erow0, next_drain, m, rows, cur.pitch, t);
                next_drain--;
            }
            __syncwarp();
            // ---- the row below: lane 31 needs column m-1-s0-ss at step ss -----------------------
            const float* br = zero_row;
            if (has_below && s0 < m) {
                const int c = m - 1 - s0 - t;
                sq_bnd_take(pfv, t < 16 && c >= 0, bin + (c >= 0 ? c : 0), epoch, bv, t);
                br = bv;
            }
            const int cn = m - 1 - s0 - t - 16;                  // the next block's entry of this lane
            const bool pf_next = has_below && t < 16 && cn >= 0;
            float* st = stage + srow * kB2StagePitch + t;
            // lane 31's column at step s is m-1-s; lane 0's is m+30-s
            const bool steady = full_rows && s0 >= 32 && s0 + 15 <= m - 1 - (SWM ? 1 : 0);
            // One unrolled body for both kinds of block (see the forward kernel): EDGE = true for the ramp blocks.
            // Q in the ramps was never written by the forward (arbitrary bits): products are selected there,
            // not multiplied by zero.  (The empty asm statements pin the hoisted loads in front of the steps:
            // without them ptxas sinks the shared-memory loads into the dependent chain.)
            auto block = [&](auto edge_tag) {
                constexpr bool EDGE = decltype(edge_tag)::value;
                float bv_[16];
#pragma unroll
                for (int q4 = 0; q4 < 4; ++q4) {
                    const float4 b4 = reinterpret_cast<const float4*>(br)[q4];
                    bv_[4 * q4] = b4.x;
                    bv_[4 * q4 + 1] = b4.y;
                    bv_[4 * q4 + 2] = b4.z;
                    bv_[4 * q4 + 3] = b4.w;
                }
                float qx_[16], qy_[16], qm_[16];
                float px_[ADJ ? 16 : 1], py_[ADJ ? 16 : 1], pm_[ADJ ? 16 : 1];
#pragma unroll
                for (int ss = 0; ss < 16; ++ss) {
                    qx_[ss] = qt[(15 - ss) * kStepFloats];
                    qy_[ss] = qt[(15 - ss) * kStepFloats + kQY];
                    if (ADJ) {
                        px_[ss] = qt[kDiagElems + (15 - ss) * kStepFloats];
                        py_[ss] = qt[kDiagElems + (15 - ss) * kStepFloats + kQY];
                    }
                }
#pragma unroll
                for (int ss = 0; ss < 16; ++ss) {
                    asm volatile("" : "+f"(qx_[ss]), "+f"(qy_[ss]));
                    if (ADJ) asm volatile("" : "+f"(px_[ss]), "+f"(py_[ss]));
                }
#pragma unroll
                for (int ss = 0; ss < 16; ++ss) {
                    qm_[ss] = (1.f - qx_[ss]) - qy_[ss];      // >= 0 by the forward's clamp
                    if (ADJ) pm_[ss] = -(px_[ss] + py_[ss]);  // Qd sums to 0 over the states
                    if (ADJ && !EDGE) {
                        const bool live = qx_[ss] >= 0.f;     // a marked cell (Q == 0) pushes nothing
                        qx_[ss] = live ? qx_[ss] : 0.f;
                        qy_[ss] = live ? qy_[ss] : 0.f;
                        qm_[ss] = live ? qm_[ss] : 0.f;
                    }
                }
                const bool seed_blk = EDGE && !ADJ && bottom && s0 < 32;      // E[n, m] = Et lives here
                const int c0 = m - 1 - s0 + u;                // the lane's column at step 0 of the block
                unsigned long long* bw = bout + (m + 30 - s0);
#pragma unroll
                for (int ss = 0; ss < 16; ++ss) {
                    float zin = __shfl_down_sync(kFull, zout, 1);
                    if (t == 31) zin = bv_[ss];
                    float e = zin + yprev;
                    float X, Y, D;
                    if (!EDGE) {
                        if (SWM) e = sw_dead ? 0.f : e;       // row 1: E = 0, nothing pushed (0 * mark = 0)
                        X = ADJ ? fmaf(qx_[ss], e, px_[ss]) : qx_[ss] * e;
                        Y = ADJ ? fmaf(qy_[ss], e, py_[ss]) : qy_[ss] * e;
                        D = ADJ ? fmaf(qm_[ss], e, pm_[ss]) : qm_[ss] * e;
                    } else {
                        const int c = c0 - ss;
                        const bool in = row_ok && (unsigned)c < (unsigned)m;
                        const bool comp = SWM ? (in && rowcomp && c >= 1) : in;
                        if (seed_blk && t == rows - 1 && c == m - 1) e = et;   // nw.py:125-127
                        e = comp ? e : 0.f;
                        if (ADJ) {
                            const bool live = comp && qx_[ss] >= 0.f;
                            X = live ? fmaf(qx_[ss], e, px_[ss]) : 0.f;
                            Y = live ? fmaf(qy_[ss], e, py_[ss]) : 0.f;
                            D = live ? fmaf(qm_[ss], e, pm_[ss]) : 0.f;
                        } else {
                            X = comp ? qx_[ss] * e : 0.f;
                            Y = comp ? qy_[ss] * e : 0.f;
                            D = comp ? qm_[ss] * e : 0.f;
                        }
                    }
                    st[ss * kB2StagePitch] = e;
                    zout = X + dprev;
                    dprev = D;
                    yprev = Y;
                    // (lane 0's column at this step is m + 30 - s0 - ss: a warp-uniform address, immediate offset)
                    if (t == 0 && feeds_up && (!EDGE || (row_ok && (unsigned)(m + 30 - s0 - ss) < (unsigned)m)))
                        sq_publish(bw - ss, sq_pack(epoch, zout), p.dbg);
                    if (ss == kSqPfStep && pf_next) pfv = ld_relaxed_gpu_u64(bin + cn);
                    if (ss == kSqRearmStep) {
                        // the block's Q is in registers (shared-memory loads of a warp return in order: when the
                        // gate is back, so are they), so its slot can take the tile RING blocks ahead now instead
                        // of at the top of the next block: RING tiles in flight instead of RING - 1
                        const float gate = *reinterpret_cast<const volatile float*>(zero_row);
                        if (gate == 0.f) pump(b + RING);
                    }
                }
            };
            if (steady) block(std::false_type{});
            else block(std::true_type{});
            srow += 16;
            if (srow == kB2StageSteps) srow = 0;
        }
        __syncwarp();
        while (next_drain >= 0) {
            if (full_rows && next_drain * kTile + kTile <= m) sq_drain_tile<true>(stage, Erow0, next_drain, m, rows, cur.pitch, t);
            else sq_drain_tile<false>(stage, Erow0, next_drain, m, rows, cur.pitch, t);
            next_drain--;
        }
        if (p.trace && t == 0) p.trace[2 * cur_tk + 1] = global_ns();
        issued -= NBk;
        cur = nxt;
        cur_tk = nxt_tk;
        nxt_ready = false;
        nxt.rows = 0;
    }
    sq_exit(p.ctl, epoch);
}

}  // namespace b200dp
