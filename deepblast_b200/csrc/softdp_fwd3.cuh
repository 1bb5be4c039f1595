// softdp_fwd3.cuh -- forward fill, CHAINED fast path for large batches of equal-size
// lattices (reference: deepblast/nw.py:46-62, sw.py:46-62; replaces
// deepblast/nw_cuda.py:46-79).
//
// Same cell arithmetic as softdp_fwd2.cuh (difference form, log2 units) and the same
// strip-major Q, but the wavefront never drains:
//   * one warp per CTA, no inter-warp hand-off at all.  The warp is dealt pairs round by
//     round (boustrophedon) and walks each pair as P = ceil(K / NCH) passes of NCH strips
//     side by side ("chains": lane t of chain c owns row 32 (NCH pass + c) + t + 1).  Chain c
//     runs 32 steps behind chain c-1, so its lane 0 takes the row above from lane 31 of
//     chain c-1 by the same rotate-shuffle that feeds the other lanes: two independent
//     dependent chains per lane (NCH = 2) to fill the issue slots of a latency-bound step.
//   * CHAINING.  All passes of all the CTA's pairs form one linear sequence of segments of
//     M columns.  At step S lane t of chain c sits at linear position L = S - t - 32 c,
//     i.e. column L mod M of segment L / M: when a lane reaches the end of its row it
//     starts the same row of the next segment on the very next step.  There are no ramps
//     between strips or pairs (only once at the start and the end of the CTA's life): at
//     256 x 256 the 32-lane wavefront is 100 % occupied instead of 89 %, and the expensive
//     ramp blocks of softdp_fwd2.cuh disappear.  Blocks of 16 steps in which every lane is
//     in the same segment run the predicate-free unrolled "steady" code; the 2 NCH blocks
//     per segment in which lanes roll over run the same step with per-lane selects.
//   * the strip-major Q of consecutive strips is itself chained (strip_stride = M steps),
//     so with NCH = 1 a pair is one sequential 256-byte-per-step stream end to end.
//   * theta/A: TMA 16x16 boxes per 16-row group exactly as in softdp_fwd2.cuh; group j
//     (16 j rows behind the leading edge) needs linear tile e - j in event e, so the
//     consumer's shared-memory addressing does not know about segments at all.
//   * the row above a pass (bottom row of the previous pass) lives in ONE boundary row of
//     M floats per warp: the writer (lane 31 of the last chain) is M - 32 NCH + 1 steps
//     ahead of the reader (lane 0 of chain 0) of the same column.
// Requirements (checked by the host): no per-pair lengths, M % 16 == 0, M >= 32 NCH + 32.
#pragma once
#include <type_traits>

#include "softdp_fwd2.cuh"

namespace b200dp {

// ADJ: the same chained sweep as the ADJOINT forward pass (nw.py:178-199, replaces
// nw_cuda.py:105-139) in difference form:
//   dx = za + hd[i-1,j], dy = za + vd[i,j-1]        (w_x - w_m, w_y - w_m of nw.py:188-192)
//   g  = q_x dx + q_y dy                             (sum_s q_s w_s - w_m; Q sums to 1)
//   Vd[i,j] - Vd[i-1,j-1] = ztheta + g,   Qd_x = q_x (dx - g),  Qd_y = q_y (dy - g)
// Operands per 16x16 box: ztheta, ZA (optional), E (optional) -- contiguous [B, N, M]
// tensors through TMA exactly like theta / A; the forward's Q arrives as a strip-major
// stream of 4 KB bulk-TMA tiles (main tile of the block, plus the tail tile of the previous
// pair in the two blocks after a pair boundary, as in softdp_bwd3.cuh); the output stream
// holds Qd * E (what the adjoint backward sweep consumes), or Qd when E is absent.  A cell
// whose Q carries the zero mark has Vd = ztheta (its diagonal predecessor is on the zero
// border) and Qd = 0.  Vtd = Vd[N, M] = sum_j hd[N, j].
template <int NCH, int RING, bool ADJ = false, int kAdj3QRing = 3>
__host__ __device__ inline size_t fwd3_smem_bytes(int M) {
    size_t b = (size_t)RING * NCH * (ADJ ? 6144 : 4096);   // [RING][2 NCH groups][theta, A (, E)][16][16] fp32
    if (ADJ) b += (size_t)kAdj3QRing * kDiagElems * 4 + (size_t)kAdj3QRing * 8;
    b += (size_t)RING * 8;                         // mbarriers
    b = (b + 15) & ~(size_t)15;
    b += (size_t)M * 4;                            // boundary row
    b += 128;                                      // 16 zeros + slack
    return b;
}

// One step of the adjoint forward sweep (see the ADJ note above); same calling convention
// as fwd2_step.  e = 1 when the E operand is absent, za = 0 when ZA is absent.
template <bool EDGE>
__device__ __forceinline__ float adj3_step(float zt, float za, float e, float qx, float qy, float hup, float& v,
                                           float* __restrict__ qp, bool store, bool comp) {
    const bool live = qx >= 0.f;                     // a marked cell: Q == 0
    qx = live ? qx : 0.f;
    qy = live ? qy : 0.f;
    const float dx = za + hup;
    const float dy = za + v;
    const float g = fmaf(qx, dx, qy * dy);
    const float ld = zt + g;
    float hn = ld - v;
    float vn = ld - hup;
    const float qdx = (qx * (dx - g)) * e;           // nw.py:30-43; Qd_m = -(Qd_x + Qd_y) is implied
    const float qdy = (qy * (dy - g)) * e;
    if (EDGE) {
        hn = comp ? hn : 0.f;
        vn = comp ? vn : 0.f;
    }
    if (!EDGE || store) {
        qp[0] = qdx;
        qp[kQY] = qdy;
    }
    v = vn;
    return hn;
}

template <bool SWM, int NCH, int RING, int DBG = 0, bool ADJ = false, int kAdj3QRing = 3>
__global__ void __launch_bounds__(32) softdp_fwd3_kernel(const __grid_constant__ CUtensorMap tm_theta,
                                                         const __grid_constant__ CUtensorMap tm_A,
                                                         const __grid_constant__ CUtensorMap tm_pf_theta,
                                                         const __grid_constant__ CUtensorMap tm_pf_A, FwdParams p) {
    static_assert(!ADJ || (NCH == 1 && !SWM && DBG == 0), "adjoint forward: one chain, full range");
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    constexpr int kGroupBytes = ADJ ? 3072 : 2048;       // [theta, A (, E)][16][16] fp32
    constexpr int kSlot = NCH * 2 * kGroupBytes;
    constexpr int kGroups = 2 * NCH;
    const CUtensorMap& tm_E = tm_pf_theta;               // ADJ: the third operand travels in the prefetch map's slot
    const int t = threadIdx.x, g = t >> 4, tp = t & 15;
    const int N = p.d.N, M = p.d.M, B = p.d.B;
    const int K = (N + 31) >> 5, P = (K + NCH - 1) / NCH, T16 = M >> 4;

    unsigned char* ring = smem_raw;
    constexpr size_t kQRingBytes = ADJ ? (size_t)kAdj3QRing * kDiagElems * 4 : 0;
    float* qring = reinterpret_cast<float*>(smem_raw + (size_t)RING * kSlot);          // ADJ: Q tiles
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + (size_t)RING * kSlot + kQRingBytes);
    uint64_t* qbars = bars + RING;
    constexpr size_t kBarBytes = (size_t)(RING + (ADJ ? kAdj3QRing : 0)) * 8;
    float* bnd = reinterpret_cast<float*>(smem_raw + (((size_t)RING * kSlot + kQRingBytes + kBarBytes + 15) & ~(size_t)15));
    float* zero_row = bnd + M;

    if (t == 0) {
        for (int s = 0; s < RING; ++s) mbar_init(&bars[s], 1);
        if (ADJ)
            for (int s = 0; s < kAdj3QRing; ++s) mbar_init(&qbars[s], 1);
        tma_prefetch_desc(&tm_theta);
        tma_prefetch_desc(&tm_A);
        if (p.pf_tiles > 0) {
            tma_prefetch_desc(&tm_pf_theta);
            tma_prefetch_desc(&tm_pf_A);
        }
    }
    if (t < 16) zero_row[t] = 0.f;
    fence_mbar_init();
    __syncthreads();

    // pairs dealt to this CTA: round r gives pair r*grid + (r odd ? grid-1-bid : bid)
    const int grid = (int)gridDim.x, bid = (int)blockIdx.x;
    const int R = B / grid, rem = B - R * grid;
    const int npairs = R + ((((R & 1) ? grid - 1 - bid : bid) < rem) ? 1 : 0);
    if (npairs == 0) return;
    auto pair_of = [&](int r) { return r * grid + ((r & 1) ? grid - 1 - bid : bid); };
    const int G = npairs * P;                 // segments (passes) of this CTA
    const int NT = G * T16;                   // linear 16-column tiles
    const int NEV = NT + kGroups - 1;         // events that carry data
    const int NBLK = NT + kGroups;            // 16-step blocks until the last lane is done
    const long long PS = p.ql.pair_stride, SS = p.ql.strip_stride;

    // ---- producer: event e = {group j: linear tile e - j} -------------------------------
    int c_ct[kGroups], c_pass[kGroups], c_idx[kGroups];
#pragma unroll
    for (int j = 0; j < kGroups; ++j) c_ct[j] = c_pass[j] = c_idx[j] = 0;
    int issued = 0;
    unsigned islot = 0, wslot = 0, phases = 0;
    // L2 prefetch cursor: wide boxes (pf_tiles x 16 columns, all 32 NCH rows of the pass)
    // pf_dist tiles ahead of the leading tile, so that DRAM is read in long contiguous
    // row pieces and the small 16x16 boxes of the ring hit in L2
    const int pf_tiles = p.pf_tiles, pf_dist = p.pf_dist;
    int pf_ct = 0, pf_pass = 0, pf_idx = 0, pf_tau = 0;
    auto pf_step = [&](bool leader) {
        if (pf_tau < NT) {
            if (leader && (pf_ct % pf_tiles) == 0) {
                const int pair = pair_of(pf_idx);
                tma_prefetch_l2_3d(&tm_pf_theta, pf_ct * kG, pf_pass * NCH * kTile, pair);
                tma_prefetch_l2_3d(&tm_pf_A, pf_ct * kG, pf_pass * NCH * kTile, pair);
            }
            pf_tau++;
            if (++pf_ct == T16) {
                pf_ct = 0;
                if (++pf_pass == P) {
                    pf_pass = 0;
                    ++pf_idx;
                }
            }
        }
    };
    if (pf_tiles > 0) {
        const bool leader = elect_one();
        for (int i = 0; i < pf_dist; ++i) pf_step(leader);
    }
    auto issue_event = [&](int e, unsigned slot) {
        unsigned bytes = 0;
        const unsigned gbytes = ADJ ? 1024u * (1u + (p.has_za ? 1u : 0u) + (p.has_e ? 1u : 0u)) : 2048u;
#pragma unroll
        for (int j = 0; j < kGroups; ++j)
            if (e - j >= 0 && e - j < NT) bytes += gbytes;
        const bool leader = !(DBG & 2) && elect_one();
        if (pf_tiles > 0) pf_step(leader);
        if (leader) mbar_expect_tx(&bars[slot], bytes);
#pragma unroll
        for (int j = 0; j < kGroups; ++j) {
            if (e - j >= 0 && e - j < NT) {
                if (leader) {
                    const int row0 = (c_pass[j] * NCH + (j >> 1)) * kTile + (j & 1) * kG;
                    const int pair = pair_of(c_idx[j]);
                    unsigned char* dst = ring + slot * kSlot + j * kGroupBytes;
                    tma_load_3d(dst, &tm_theta, &bars[slot], c_ct[j] * kG, row0, pair);
                    if (!ADJ || p.has_za) tma_load_3d(dst + 1024, &tm_A, &bars[slot], c_ct[j] * kG, row0, pair);
                    if (ADJ && p.has_e) tma_load_3d(dst + 2048, &tm_E, &bars[slot], c_ct[j] * kG, row0, pair);
                }
                if (++c_ct[j] == T16) {
                    c_ct[j] = 0;
                    if (++c_pass[j] == P) {
                        c_pass[j] = 0;
                        ++c_idx[j];
                    }
                }
            }
        }
    };

    // ---- consumer state -------------------------------------------------------------------
    float h[NCH], v[NCH], acc_hi[NCH], acc_lo[NCH];
#pragma unroll
    for (int c = 0; c < NCH; ++c) h[c] = v[c] = acc_hi[c] = acc_lo[c] = 0.f;
    int posL = 0, gL = 0, idxL = 0, passL = 0;      // segment of the leading edge (lane 0, chain 0)
    int idxT = 0, passT = 0;                        // the segment before it
    const int tlast = (N - 1) & 31, clast = ((N - 1) >> 5) % NCH;
    const int rot = (t + 31) & 31;
    const int lanebase = g * kGroupBytes + tp * 60; // group-in-chain, row, -4*tp column skew (bytes)
    unsigned slotA = 0, slotB = 0;

    // ---- ADJ: the forward's Q as a FIFO of 4 KB tiles (16 lines of the pair's stream).  Block
    // a = pass * T16 + pos / 16 of pair idx needs main tile a of that pair and, while a < 2 and
    // idx >= 1, tail tile TA + a of the previous pair (lanes that have not rolled over yet) ------
    const int TA = P * T16;
    int iss_idx = 0, iss_a = 0, iss_sub = 0, iss_cnt = 0, con_cnt = 0;
    unsigned qislot = 0, qwslot = 0, qphases = 0;
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
    auto wait_qtile = [&]() -> const float* {
        mbar_wait(&qbars[qwslot], (qphases >> qwslot) & 1u);
        qphases ^= 1u << qwslot;
        const float* sq = qring + qwslot * kDiagElems;
        qwslot = (qwslot + 1 == kAdj3QRing) ? 0u : qwslot + 1;
        return sq;
    };

    for (int b = 0; b < NBLK; ++b) {
        __syncwarp();
        while (issued <= b + RING - 2 && issued < NEV) {
            issue_event(issued, islot);
            issued++;
            islot = (islot + 1 == RING) ? 0u : islot + 1;
        }
        slotA = slotB;
        if (b < NEV) {
            if (!(DBG & 2)) mbar_wait(&bars[wslot], (phases >> wslot) & 1u);
            phases ^= 1u << wslot;
            slotB = wslot;
            wslot = (wslot + 1 == RING) ? 0u : wslot + 1;
        }
        const float* qmain = nullptr;
        const float* qtail = nullptr;
        bool has_tail = false;
        if (ADJ) {
            int pr, tile;
            while (iss_cnt - con_cnt < kAdj3QRing && next_tile(pr, tile)) {
                q_tile_load<true>(qring + qislot * kDiagElems, &qbars[qislot], p.Qin + (long long)pair_of(pr) * PS,
                                  kDiagRows * tile, t);
                qislot = (qislot + 1 == kAdj3QRing) ? 0u : qislot + 1;
                iss_cnt++;
            }
            const bool has_main = idxL < npairs;
            has_tail = passL == 0 && posL < 32 && idxL >= 1;
            if (has_main) qmain = wait_qtile();
            if (has_tail) qtail = wait_qtile();
            if (!has_main) qmain = qtail;
            if (!has_tail) qtail = qmain;
            con_cnt += (has_main ? 1 : 0) + (has_tail ? 1 : 0);
        }
        const bool plain = posL >= 32 * NCH;        // every lane of every chain is in segment gL
        const bool Lvalid = gL < G;
        const bool fullL = (passL + 1) * NCH * kTile <= N;
        const float* br = (Lvalid && passL > 0) ? bnd + posL : zero_row;
        const unsigned char* sA = ring + slotA * kSlot + lanebase + 64;
        const unsigned char* sB = ring + slotB * kSlot + lanebase;
        float part[NCH];
#pragma unroll
        for (int c = 0; c < NCH; ++c) part[c] = 0.f;

        // roll-over inside one pair (single chain), both strips complete: the steady step plus
        // the row-start resets; no pair boundary, no partial rows, no Vt
        const bool simple_roll = (NCH == 1) && !plain && Lvalid && passL > 0 && fullL;
        if ((plain && Lvalid && fullL) || simple_roll) {
            // ---- steady block (ROLL = false): one segment, every row inside the lattice;
            // simple roll-over block (ROLL = true): lanes t < ss + posL already in segment L ----
            auto body = [&](auto roll_tag) {
                constexpr bool ROLL = decltype(roll_tag)::value;
                const int pairL = pair_of(idxL);
                // (with one chain the storage of segment T continues into segment L: one pointer)
                float* qb = p.Q + (long long)pairL * PS + (long long)(passL * NCH) * SS + (long long)posL * kStepFloats + t;
                float* bw = bnd + (posL - 32 * NCH + 1);
                const int roll0 = t - posL;                   // ROLL: step at which the lane enters L
                float th_[NCH][16], a_[NCH][16], bv_[16];
                float e_[ADJ ? 16 : 1], qx_[ADJ ? 16 : 1], qy_[ADJ ? 16 : 1];
#pragma unroll
                for (int c = 0; c < NCH; ++c) {
#pragma unroll
                    for (int ss = 0; ss < 16; ++ss) {
                        const float* tb = reinterpret_cast<const float*>(((tp <= ss) ? sB : sA) + c * 2 * kGroupBytes);
                        th_[c][ss] = tb[ss];
                        a_[c][ss] = (!ADJ || p.has_za) ? tb[ss + 256] : 0.f;
                        if (ADJ) {
                            e_[ss] = p.has_e ? tb[ss + 512] : 1.f;
                            qx_[ss] = qmain[ss * kStepFloats + t];
                            qy_[ss] = qmain[ss * kStepFloats + kQY + t];
                        }
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
#pragma unroll
                for (int ss = 0; ss < 16; ++ss) {
                    float r[NCH];
#pragma unroll
                    for (int c = 0; c < NCH; ++c) r[c] = __shfl_sync(kFull, h[c], rot);
#pragma unroll
                    for (int c = 0; c < NCH; ++c) {
                        const float hup = (t == 0) ? (c == 0 ? bv_[ss] : r[c > 0 ? c - 1 : 0]) : r[c];
                        bool at = false;
                        if (ROLL) {
                            // first column of the lane's new row: V[i, 0] = 0, a new row sum
                            at = ss == roll0;
                            v[c] = at ? 0.f : v[c];
                            part[c] = at ? 0.f : part[c];
                        }
                        // chain c: strip passL*NCH + c, wavefront step posL + ss - 32 c of that strip
                        float* qp = qb + (long long)c * (SS - 32 * kStepFloats) + ss * kStepFloats;
                        // sw.py: row 1 (lane 0 of the pair's first strip) is below the origin -- V = 0, Q = 0
                        // (column 1 is always met in a roll-over block, i.e. in the general path)
                        // and column 1 (the first column of a new row, ROLL only) as well
                        const bool live = !(SWM && ((c == 0 && t == 0 && passL == 0) || at));
                        if (ADJ)
                            h[c] = adj3_step<false>(th_[c][ss], a_[c][ss], e_[ss], qx_[ss], qy_[ss], hup, v[c], qp, true, true);
                        else
                            h[c] = fwd2_step<false, SWM, DBG>(th_[c][ss], a_[c][ss], hup, v[c], qp, true, live);
                        part[c] += h[c];
                    }
                    if (t == 31) {
                        if (ROLL) {
                            // lane 31 is 31 columns behind the leading edge: until it rolls over it
                            // still writes the boundary row of the previous segment
                            int bi = posL + ss - 31;
                            bi += (bi < 0) ? M : 0;
                            bnd[bi] = h[NCH - 1];
                        } else {
                            bw[ss] = h[NCH - 1];
                        }
                    }
                }
                if (ROLL) {
                    // a lane that finished a row (not the pair's last: no Vt here) starts a new row sum
                    if (roll0 >= 0 && roll0 < 16) {
#pragma unroll
                        for (int c = 0; c < NCH; ++c) acc_hi[c] = acc_lo[c] = 0.f;
                    }
                }
            };
            if (simple_roll) body(std::true_type{});
            else body(std::false_type{});
        } else {
            // ---- general block: lanes may sit in two segments (T = the older one, L), be
            // before the start / past the end of the CTA's sequence, or below the lattice ----
            const bool Tvalid = plain ? Lvalid : (gL >= 1 && gL - 1 < G);
            const int iT = plain ? idxL : idxT, pT = plain ? passL : passT;
            const int pairL = Lvalid ? pair_of(idxL) : 0, pairT = Tvalid ? pair_of(iT) : 0;
            int roll[NCH];
            bool okL[NCH], okT[NCH];
            float *qL[NCH], *qT[NCH];
            float snap[NCH];
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                roll[c] = plain ? -64 : (t + 32 * c - posL);     // step at which the lane enters L
                okL[c] = Lvalid && (passL * NCH + c) * kTile + t < N;
                okT[c] = Tvalid && (pT * NCH + c) * kTile + t < N;
                qL[c] = p.Q + (long long)pairL * PS + (long long)(passL * NCH + c) * SS +
                        (long long)(posL - 32 * c) * kStepFloats + t;
                qT[c] = p.Q + (long long)pairT * PS + (long long)(pT * NCH + c) * SS +
                        (long long)(posL + M - 32 * c) * kStepFloats + t;
                snap[c] = 0.f;
            }
#pragma unroll 4
            for (int ss = 0; ss < 16; ++ss) {
                float r[NCH];
#pragma unroll
                for (int c = 0; c < NCH; ++c) r[c] = __shfl_sync(kFull, h[c], rot);
                const float bvs = br[ss];
#pragma unroll
                for (int c = 0; c < NCH; ++c) {
                    const float hup = (t == 0) ? (c == 0 ? bvs : r[c > 0 ? c - 1 : 0]) : r[c];
                    const bool inL = ss >= roll[c];
                    const bool at = ss == roll[c];             // first column of the lane's new row
                    snap[c] = at ? part[c] : snap[c];
                    part[c] = at ? 0.f : part[c];
                    v[c] = at ? 0.f : v[c];                    // V[i, 0] = 0
                    const bool ok = inL ? okL[c] : okT[c];
                    bool comp = ok;
                    if (SWM) comp = ok && !at && !(t == 0 && c == 0 && (inL ? passL : pT) == 0);
                    float* qp = (inL ? qL[c] : qT[c]) + ss * kStepFloats;
                    const float* tb = reinterpret_cast<const float*>(((tp <= ss) ? sB : sA) + c * 2 * kGroupBytes);
                    if (ADJ) {
                        // lanes that have not rolled over yet read the previous pair's tail tile
                        const float* qt = ((has_tail && !inL) ? qtail : qmain) + ss * kStepFloats + t;
                        h[c] = adj3_step<true>(tb[ss], p.has_za ? tb[ss + 256] : 0.f, p.has_e ? tb[ss + 512] : 1.f,
                                               ok ? qt[0] : 0.f, ok ? qt[kQY] : 0.f, hup, v[c], qp, ok, comp);
                    } else {
                        h[c] = fwd2_step<true, SWM, DBG>(tb[ss], tb[ss + 256], hup, v[c], qp, ok, comp);
                    }
                    part[c] += h[c];
                }
                if (t == 31) {
                    const bool inL = ss >= roll[NCH - 1];
                    bnd[posL + ss - 32 * NCH + 1 + (inL ? 0 : M)] = h[NCH - 1];
                }
            }
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                if (roll[c] >= 0 && roll[c] < 16) {
                    // the lane finished a row of segment T in this block:
                    // Vt = V[N, M] = ln 2 * sum_j h[N, j]
                    if (Tvalid && pT == P - 1 && c == clast && t == tlast)
                        p.Vt[pairT] = (acc_hi[c] + (acc_lo[c] + snap[c])) * (ADJ ? 1.f : kLn2);
                    acc_hi[c] = 0.f;
                    acc_lo[c] = 0.f;
                }
            }
        }
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
            // fold the block's partial row sum into the two-float accumulator (Fast2Sum)
            const float t1 = part[c] + acc_lo[c];
            const float nh = acc_hi[c] + t1;
            acc_lo[c] = t1 - (nh - acc_hi[c]);
            acc_hi[c] = nh;
        }
        posL += 16;
        if (posL == M) {
            posL = 0;
            idxT = idxL;
            passT = passL;
            gL++;
            if (++passL == P) {
                passL = 0;
                idxL++;
            }
        }
    }
}

}  // namespace b200dp
