// softdp_bwd3.cuh -- backward sweep, CHAINED fast path for large batches of equal-size
// lattices (reference: deepblast/nw.py:120-135, sw.py:100-115; replaces
// deepblast/nw_cuda.py:82-102).
//
// Same push-form step as softdp_bwd2.cuh, organised like softdp_fwd3.cuh: one warp per
// CTA, no hand-offs, and the wavefront never drains.  The strips of all the CTA's pairs
// (each pair bottom-up) form one linear sequence of segments of M columns swept right to
// left; at step S lane t (u = 31 - t, lane 31 leads) sits at linear position S - u, and a
// lane that finishes its row starts the same row of the strip above on the next step.
//   * Q: in the chained strip-major layout (strip_stride = M steps) the line of wavefront
//     step sigma of strip kb+1 IS the line of step sigma + M of strip kb, so every lane of
//     the warp -- whichever of two strips it is in -- reads the same 256-byte line, and the
//     sweep walks a pair's Q storage sequentially from its end to its beginning: one 4 KB
//     1-D bulk-TMA tile per 16 steps.  Only at a pair boundary two tiles are live (the
//     first 31 lines of the old pair, the last lines of the new one).
//   * E is staged step-major (pitch 33) and complete 32-column tiles are drained row-major
//     exactly as in softdp_bwd2.cuh; in linear time one tile completes every 32 steps.
//   * the row below a strip (what lane 0 of the previous segment pushed up) lives in one
//     boundary row of M floats per warp.
// Requirements (checked by the host): no per-pair lengths, M % 32 == 0, M >= 64.
#pragma once
#include <type_traits>

#include "softdp_bwd2.cuh"

namespace b200dp {

constexpr int kB3StageBytes = ((kB2StageFloats * 4 + 127) / 128) * 128;

// ADJ: the same sweep as the ADJOINT backward pass (nw.py:251-267, replaces nw_cuda.py:142-165):
//   Ed[i,j] = sum over the three successors of ( Qd E + Q Ed ), push form: the cell (i,j),
//   once ed = Ed[i,j] is known, hands X = Qd_x E + Q_x ed (and likewise D, Y) to its
//   predecessors.  The products Qd[i,j,:] * E[i,j] are formed by the adjoint FORWARD sweep
//   (softdp_adj3.cuh), which has E at hand in stream order, and arrive here as a second
//   strip-major stream (two states, the m state is -(x + y)) next to Q: one more 4 KB bulk
//   copy per tile and three FFMAs instead of FMULs per step.  No seed (Ed[N,M] = 0 + pushes),
//   all borders zero; cells whose Q carries the zero mark push nothing.
template <int RING, bool ADJ = false>
__host__ __device__ inline size_t bwd3_smem_bytes(int M) {
    size_t b = (size_t)RING * kDiagElems * (ADJ ? 2 : 1) * 4 + kB3StageBytes;
    b += (size_t)RING * 8;
    b = (b + 15) & ~(size_t)15;
    b += (size_t)M * 4;                            // boundary row
    b += 128;                                      // 16 zeros + slack
    return b;
}

// Drain column tile tc of the segment whose first step had staging row `segrow`
// (= (segment * M) mod 80): element (r, col) was produced at segment step
// (M-1-col) + (31-r), i.e. it sits in staging row (ub - t - r) mod 80 with the warp-uniform
// ub = segrow + M + 30 - 32 tc (lane t = column 32 tc + t).  One wrap at most, so row r is
// read from p0 - 32 r (r <= sr0) or p1 - 32 r (r > sr0): compare, select, LDS with an
// immediate offset, one IMAD.WIDE for the row address, STG -- 5 instructions per row.
// FULL = all 32 rows of the strip are inside the lattice (no store predicates).
template <bool FULL>
__device__ __forceinline__ void bwd3_drain_tile(const float* __restrict__ stage, float* __restrict__ Erow0, int tc,
                                                int M, int rmax, int pitch, int t, int segrow,
                                                float* __restrict__ Ei_row0 = nullptr) {
    int sr0 = (segrow + M + 30 - tc * kTile) % kB2StageSteps - t;
    sr0 += (sr0 < 0) ? kB2StageSteps : 0;
    const float* p0 = stage + sr0 * kB2StagePitch;
    const float* p1 = p0 + kB2StageFloats;
    float* dstp = Erow0 + tc * kTile + t + 1;
#pragma unroll
    for (int r0 = 0; r0 < kTile; r0 += 8) {
        float v[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int r = r0 + q;
            const float* ps = (r > sr0) ? p1 : p0;
            v[q] = ps[-(kB2StagePitch - 1) * r];
        }
        if (Erow0) {                                  // (ADJ may ask for the interior only)
#pragma unroll
            for (int q = 0; q < 8; ++q)
                if (FULL || r0 + q < rmax) dstp[(r0 + q) * pitch] = v[q];
        }
        if (Ei_row0) {
            // the same rows into the contiguous interior copy (pitch M, 128-byte aligned lines)
            float* d2 = Ei_row0 + tc * kTile + t;
#pragma unroll
            for (int q = 0; q < 8; ++q)
                if (FULL || r0 + q < rmax) d2[(r0 + q) * M] = v[q];
        }
    }
}

template <bool SWM, int RING, bool ADJ = false>
__global__ void __launch_bounds__(32) softdp_bwd3_kernel(BwdParams p) {
    static_assert(!(SWM && ADJ), "the adjoint sweeps cover the full range (sw.py:150-151,199-201)");
    constexpr int kSlotElems = kDiagElems * (ADJ ? 2 : 1);      // [Q tile | QdE tile]
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const int t = threadIdx.x, u = 31 - t;
    const int N = p.d.N, M = p.d.M, B = p.d.B;
    const int K = (N + 31) >> 5, T32 = M >> 5;
    const int TA = (K * M) >> 4;                  // 16-step blocks (= main Q tiles) per pair

    float* qring = reinterpret_cast<float*>(smem_raw);
    float* stage = qring + RING * kSlotElems;
    size_t off = (size_t)RING * kSlotElems * 4 + kB3StageBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + off);
    off = (off + (size_t)RING * 8 + 15) & ~(size_t)15;
    float* bnd = reinterpret_cast<float*>(smem_raw + off);
    float* zero_row = bnd + M;

    if (t == 0)
        for (int s = 0; s < RING; ++s) mbar_init(&bars[s], 1);
    if (t < 16) zero_row[t] = 0.f;
    fence_mbar_init();
    __syncthreads();

    const int grid = (int)gridDim.x, bid = (int)blockIdx.x;
    const int R = B / grid, rem = B - R * grid;
    const int npairs = R + ((((R & 1) ? grid - 1 - bid : bid) < rem) ? 1 : 0);
    if (npairs == 0) return;
    auto pair_of = [&](int r) { return r * grid + ((r & 1) ? grid - 1 - bid : bid); };
    const int G = npairs * K;                     // segments (strips) of this CTA
    const int NBLK = npairs * TA + 2;             // 16-step blocks until lane 0 is done
    const int NDRAIN = G * T32;                   // 32-column E tiles to drain
    const long long PS = p.ql.pair_stride;
    const long long Epair = (long long)(N + 2) * (M + 2);
    const int tlast = (N - 1) & 31;

    // ---- Q tile stream: block bb of pair idx = bb / TA needs main tile a = bb % TA of that
    // pair and, while a < 2 and idx >= 1, tail tile TA + a of the previous pair ---------------
    int iss_idx = 0, iss_a = 0, iss_sub = 0, iss_cnt = 0, con_cnt = 0;
    unsigned islot = 0, wslot = 0, phases = 0;
    auto next_tile = [&](int& pr, int& tile) -> bool {
        for (;;) {
            if (iss_idx > npairs || (iss_idx == npairs && iss_a >= 2)) return false;
            if (iss_sub == 0) {
                iss_sub = 1;
                if (iss_idx < npairs) {
                    pr = iss_idx;
                    tile = iss_a;
                    return true;
                }
            } else {
                const bool tail = iss_a < 2 && iss_idx >= 1;
                const int pi = iss_idx - 1, ta = TA + iss_a;
                iss_sub = 0;
                if (++iss_a == TA) {
                    iss_a = 0;
                    ++iss_idx;
                }
                if (tail) {
                    pr = pi;
                    tile = ta;
                    return true;
                }
            }
        }
    };
    auto wait_tile = [&]() -> const float* {
        mbar_wait(&bars[wslot], (phases >> wslot) & 1u);
        phases ^= 1u << wslot;
        const float* s = qring + wslot * kSlotElems;
        wslot = (wslot + 1 == RING) ? 0u : wslot + 1;
        return s;
    };

    // ---- sweep state -------------------------------------------------------------------------
    float zout = 0.f, dprev = 0.f, yprev = 0.f;
    int posS = 0, gL = 0, idxL = 0, kL = 0, a = 0;      // leading edge: segment, pair ordinal, strip ordinal (0 = bottom), block in pair
    int idxT = 0, kT = 0;                                // the segment before it
    int srow = 0;                                        // (16 b) mod 80
    int drained = 0, d_idx = 0, d_k = 0, d_tc = T32 - 1, d_segrow = 0;

    auto drain_ready = [&](int S0) {
        // the n-th tile in drain order is complete once the sweep has passed step 64 + 32 n
        while (drained < NDRAIN && 64 + 32 * drained <= S0) {
            const int pair = pair_of(d_idx);
            const int kb = K - 1 - d_k;
            float* Eb = p.E + (long long)pair * Epair;
            float* Er = Eb + (long long)(kb * kTile + 1) * (M + 2);
            const int rmax = N - kb * kTile;
            if (ADJ && !p.E) Er = nullptr;
            float* Eir = p.Ei ? p.Ei + ((long long)pair * N + kb * kTile) * M : nullptr;
            if (rmax >= kTile) bwd3_drain_tile<true>(stage, Er, d_tc, M, kTile, M + 2, t, d_segrow, Eir);
            else bwd3_drain_tile<false>(stage, Er, d_tc, M, rmax, M + 2, t, d_segrow, Eir);
            drained++;
            if (--d_tc < 0) {
                // strip complete: zero borders, E[N+1, M+1] = Et  (nw.py:125-127, 347)
                const int i = kb * kTile + t + 1;
                const bool padded = !ADJ || p.E != nullptr;
                if (padded && i <= N) {
                    Eb[(long long)i * (M + 2)] = 0.f;
                    Eb[(long long)i * (M + 2) + M + 1] = 0.f;
                }
                if (padded && kb == 0)
                    for (int col = t; col < M + 2; col += 32) Eb[col] = 0.f;
                if (padded && d_k == 0) {
                    const float et = ADJ ? 0.f : p.Et[(long long)pair * p.et_stride];
                    for (int col = t; col < M + 2; col += 32)
                        Eb[(long long)(N + 1) * (M + 2) + col] = (col == M + 1) ? et : 0.f;
                }
                d_tc = T32 - 1;
                d_segrow = (d_segrow + M) % kB2StageSteps;
                if (++d_k == K) {
                    d_k = 0;
                    ++d_idx;
                }
            }
        }
    };

    for (int b = 0; b < NBLK; ++b) {
        __syncwarp();
        {
            int pr, tile;
            while (iss_cnt - con_cnt < RING && next_tile(pr, tile)) {
                const long long po = (long long)pair_of(pr) * PS;
                q_tile_load<true>(qring + islot * kSlotElems, &bars[islot], p.Q + po, K * M + 15 - kDiagRows * tile, t,
                                  ADJ ? 2 : 1);
                if (ADJ)
                    q_tile_load<true>(qring + islot * kSlotElems + kDiagElems, &bars[islot], p.QdE + po,
                                      K * M + 15 - kDiagRows * tile, t, 0);
                islot = (islot + 1 == RING) ? 0u : islot + 1;
                iss_cnt++;
            }
        }
        const bool has_main = idxL < npairs;
        const bool has_tail = a < 2 && idxL >= 1;
        const float* qmain = nullptr;
        const float* qtail = nullptr;
        if (has_main) qmain = wait_tile();
        if (has_tail) qtail = wait_tile();
        if (!has_main) qmain = qtail;
        if (!has_tail) qtail = qmain;
        const int S0 = b * 16;
        drain_ready(S0);
        __syncwarp();

        const bool plain = posS >= 32;              // every lane is in segment gL
        const bool Lvalid = gL < G;
        const int kbL = K - 1 - kL;
        const bool fullL = (kbL + 1) * kTile <= N;
        const float* br = (Lvalid && kL > 0) ? bnd + posS : zero_row;
        float* st = stage + srow * kB2StagePitch + t;
        const bool sw_special = SWM && posS == M - 16;        // column 1 (sw.py: j >= 2) is swept in this block
        const bool sw_dead = SWM && kbL == 0 && t == 0;       // row 1 (sw.py: i >= 2): E = 0, nothing pushed

        // roll-over inside one pair, both strips complete: the steady step plus the row-start
        // resets (no tail tile, no partial rows, no seed)
        const bool simple_roll = !plain && Lvalid && kL > 0 && (kbL + 2) * kTile <= N;
        if ((plain && Lvalid && fullL && !sw_special) || simple_roll) {
            // ---- steady block (ROLL = false) / simple roll-over block (ROLL = true) ----------
            auto body = [&](auto roll_tag) {
                constexpr bool ROLL = decltype(roll_tag)::value;
                const float* qt = qmain + t;
                const int roll = u - posS;                    // ROLL: step at which the lane enters L
                float* bw = bnd + (posS - 31);
                float bv_[16];
#pragma unroll
                for (int q4 = 0; q4 < 4; ++q4) {
                    const float4 b4 = reinterpret_cast<const float4*>(br)[q4];
                    bv_[4 * q4] = b4.x;
                    bv_[4 * q4 + 1] = b4.y;
                    bv_[4 * q4 + 2] = b4.z;
                    bv_[4 * q4 + 3] = b4.w;
                }
                // the block's own Q (32 LDS) and the implied q_m up front: the staging / boundary
                // stores inside the loop may alias the Q ring as far as the compiler knows, so loads
                // left in the loop would be serialised behind them, one shared-memory latency per step
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
                    qm_[ss] = (1.f - qx_[ss]) - qy_[ss];      // >= 0 by the forward's clamp
                    if (ADJ) {
                        // a marked cell (Q == 0, sw.py first row / column) pushes nothing
                        const bool live = qx_[ss] >= 0.f;
                        qx_[ss] = live ? qx_[ss] : 0.f;
                        qy_[ss] = live ? qy_[ss] : 0.f;
                        qm_[ss] = live ? qm_[ss] : 0.f;
                        pm_[ss] = -(px_[ss] + py_[ss]);       // Qd sums to 0 over the states
                    }
                }
#pragma unroll
                for (int ss = 0; ss < 16; ++ss) {
                    float zin = __shfl_down_sync(kFull, zout, 1);
                    if (t == 31) zin = bv_[ss];
                    if (ROLL) {
                        // a lane that starts a new row has nothing to its right
                        const bool at = ss == roll;
                        yprev = at ? 0.f : yprev;
                        dprev = at ? 0.f : dprev;
                    }
                    float e = zin + yprev;
                    if (SWM) {
                        // sw.py sweeps i, j >= 2: row 1 (lane 0 of the top strip, once it is in L) and
                        // column 1 (the step before a lane starts its new row) hold E = 0 and push
                        // nothing; their Q marks are finite, 0 * mark = 0
                        const bool dead = ROLL ? ((sw_dead && ss >= roll) || ss == roll - 1) : sw_dead;
                        e = dead ? 0.f : e;
                    }
                    const float X = ADJ ? fmaf(qx_[ss], e, px_[ss]) : qx_[ss] * e;
                    const float Y = ADJ ? fmaf(qy_[ss], e, py_[ss]) : qy_[ss] * e;
                    const float D = ADJ ? fmaf(qm_[ss], e, pm_[ss]) : qm_[ss] * e;
                    st[ss * kB2StagePitch] = e;
                    zout = X + dprev;
                    dprev = D;
                    yprev = Y;
                    if (t == 0) {
                        if (ROLL) {
                            // lane 0 is 31 columns behind the leading edge: until it rolls over it
                            // still writes the boundary row of the previous segment
                            int bi = posS + ss - 31;
                            bi += (bi < 0) ? M : 0;
                            bnd[bi] = zout;
                        } else {
                            bw[ss] = zout;
                        }
                    }
                }
            };
            if (simple_roll) body(std::true_type{});
            else body(std::false_type{});
        } else {
            // ---- general block: two segments, start / end of the sequence, partial strips.
            // Q of cells outside the lattice was never written (arbitrary bits): products
            // are selected, not multiplied by zero. -----------------------------------------
            const bool Tvalid = plain ? Lvalid : (gL >= 1 && gL - 1 < G);
            const int kbT = K - 1 - (plain ? kL : kT);
            const int roll = plain ? -64 : (u - posS);         // step at which the lane enters L
            const bool okL = Lvalid && kbL * kTile + t < N;
            const bool okT = Tvalid && kbT * kTile + t < N;
            const bool seedL = Lvalid && kL == 0 && t == tlast;
            const float etL = (Lvalid && !ADJ) ? p.Et[(long long)pair_of(idxL) * p.et_stride] : 0.f;
            const int shat = 16 * a;                            // local step of the leading edge in pair idxL
#pragma unroll 4
            for (int ss = 0; ss < 16; ++ss) {
                float zin = __shfl_down_sync(kFull, zout, 1);
                if (t == 31) zin = br[ss];
                const bool inL = ss >= roll;
                const bool at = ss == roll;                     // last column of the lane's new row
                yprev = at ? 0.f : yprev;
                dprev = at ? 0.f : dprev;
                const bool ok = inL ? okL : okT;
                float e = zin + yprev;
                if (!ADJ && at && seedL) e = etL;               // E[N, M] = Et  (nw.py:125-127)
                bool comp = ok;
                if (SWM) {
                    const int o = posS + ss - u + (inL ? 0 : M);      // columns swept in this row so far
                    const int kb = inL ? kbL : kbT;
                    comp = ok && o != M - 1 && !(kb == 0 && t == 0);  // sw.py: i, j >= 2
                }
                e = comp ? e : 0.f;
                const float* qt = ((has_tail && u > shat + ss) ? qtail : qmain) + t + (15 - ss) * kStepFloats;
                const float qx = qt[0], qy = qt[kQY];
                float X, Y, D;
                if (ADJ) {
                    const bool live = comp && qx >= 0.f;        // inside the lattice and not a marked cell
                    const float px = qt[kDiagElems], py = qt[kDiagElems + kQY];
                    X = live ? fmaf(qx, e, px) : 0.f;
                    Y = live ? fmaf(qy, e, py) : 0.f;
                    D = live ? fmaf((1.f - qx) - qy, e, -(px + py)) : 0.f;
                } else {
                    X = comp ? qx * e : 0.f;
                    Y = comp ? qy * e : 0.f;
                    D = comp ? ((1.f - qx) - qy) * e : 0.f;
                }
                st[ss * kB2StagePitch] = e;
                zout = X + dprev;
                dprev = D;
                yprev = Y;
                if (t == 0) bnd[posS + ss - 31 + (inL ? 0 : M)] = zout;
            }
        }
        con_cnt += (has_main ? 1 : 0) + ((a < 2 && idxL >= 1) ? 1 : 0);
        srow += 16;
        if (srow == kB2StageSteps) srow = 0;
        posS += 16;
        if (posS == M) {
            posS = 0;
            idxT = idxL;
            kT = kL;
            gL++;
            if (++kL == K) kL = 0;
        }
        if (++a == TA) {
            a = 0;
            idxL++;
        }
    }
    __syncwarp();
    drain_ready(NBLK * 16);
    (void)idxT;
}

}  // namespace b200dp
