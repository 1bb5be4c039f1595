// softdp_fwd2.cuh -- forward fill, fast path (reference: deepblast/nw.py:46-62,
// sw.py:46-62; replaces deepblast/nw_cuda.py:46-79).
//
// Same wavefront as softdp_fwd.cuh (lane t owns row 32k+t+1, one column per step, Q
// streamed out strip-major, 256 contiguous bytes per step) re-organised for instruction
// count, dependent-chain length and occupancy:
//   * DIFFERENCE FORM.  V itself is never formed.  Each lane carries v = V[i,j]-V[i-1,j]
//     and hands h = V[i,j]-V[i,j-1] to the lane below; with a = A[i-1,j-1]
//         dx = a + h[i-1,j],  dy = a + v[i,j-1],  l = theta + logsumexp(dx, 0, dy)
//         h[i,j] = l - v[i,j-1],  v[i,j] = l - h[i-1,j]          (nw.py:56-60 rearranged)
//     h and v stay O(1), so plain fp32 keeps |dQ| ~ 1e-6 at 1024^2 without the (hi, lo)
//     pairs of the general kernel: one shuffle per step instead of two and a dependent
//     chain of ~90 instead of ~130 cycles.  Everything is held in log2 units (theta and A
//     enter through FFMAs with log2 e) so ex2 / lg2 apply directly.
//     Vt = V[n,m] = ln 2 * sum_j h[n,j], accumulated per lane (block partial sums folded
//     into a two-float accumulator);
//   * steps run in blocks of 16; blocks in which every lane is inside the lattice are
//     fully unrolled with no per-lane predicates ("steady"), the ramps use the
//     predicated variant of the same step ("edge");
//   * theta/A are staged by TMA as 16x16 boxes for two 16-row groups per warp; group 1
//     trails group 0 by one tile, so the live window is a parallelogram and the ring is
//     12 KB per warp (3 events x {2 groups x 2 tensors x 1 KB});
//   * tile waits, boundary-row progress and ring bookkeeping happen once per block.
#pragma once
#include "softdp_fwd.cuh"

namespace b200dp {

constexpr int kG = 16;                       // rows per group == columns per tile
constexpr int kF2SlotBytes = 4096;           // [2 groups][2 tensors][16][16] fp32

// RING = events resident per warp: 2 live + (RING - 2) in flight (16 steps of lead each)
template <int RING>
__host__ __device__ inline size_t fwd2_smem_bytes(int W, int M) {
    size_t b = (size_t)W * RING * kF2SlotBytes;
    b += (size_t)W * RING * 8;
    b = (b + 15) & ~(size_t)15;
    b += (size_t)(2 * W + 1) * 8;
    b = (b + 15) & ~(size_t)15;
    b += (size_t)(W + 1) * (size_t)M * 4;
    b += 256;                                      // 16 zeros (row above a pair) + slack for ramp reads
    return b;
}

// One wavefront step of one lane in difference form (log2 units).  EDGE adds the
// lattice-membership predicates and the zero border cells; SWM adds the sw.py i, j >= 2
// rule.  hup = h of the row above at this column, v = the lane's own vertical difference
// at the previous column (in), this column (out); returns h of this cell.
// DBG: bit 0 = drop the Q stores but keep the arithmetic, bit 1 = no theta/A staging (both
// diagnostic builds only); bit 2 = score-only forward (no Q at all); bit 3 = memory skeleton (diagnostic
// builds only: the loads, the shuffle and the stores of a step with one FMA in place of the cell arithmetic).
template <bool EDGE, bool SWM, int DBG = 0>
__device__ __forceinline__ float fwd2_step(float th, float a, float hup, float& v, float* __restrict__ qp,
                                           bool store, bool comp) {
    if (DBG & 8) {
        const float hn = fmaf(v, 0.5f, th), vn = fmaf(hup, 0.5f, a);
        if (DBG & 1) {
            if (hn + vn == 12345.f) qp[0] = hn;
        } else if (!EDGE || store) {
            qp[0] = hn;
            qp[kQY] = vn;
        }
        v = vn;
        return hn;
    }
    const float dx = fmaf(a, kLog2e, hup);           // u_x - u_m   (nw.py:56-58), log2 units
    const float dy = fmaf(a, kLog2e, v);             // u_y - u_m
    // softmax / logsumexp over (dx, 0, dy), nw.py:10-27, relative to the maximum (one
    // FMNMX3): the largest term is ex2(0) = 1 exactly, S is in [1, 3].  Five XU operations
    // (3 ex2, rcp, lg2) and 15 FMA/ALU-pipe instructions per cell; the XU pipe runs at about
    // 30 % at the HBM-bound rate, so trading the selects and the Newton reciprocal of the
    // earlier two-exponential form for two more XU operations is a net win in issue slots.
    const float mx = fmaxf(fmaxf(dx, dy), 0.f);
    const float em = fast_ex2(-mx);
    const float ex = fast_ex2(dx - mx);
    const float ey = fast_ex2(dy - mx);
    const float S = (em + ex) + ey;
    const float r = fast_rcp(S);
    // q_m = em * r is implied, never stored: readers form (1 - q_x) - q_y, which the clamp
    // keeps >= 0 exactly (ex * r, or q_x + q_y, can round to 1 + 1 ulp when q_m underflows)
    float qx = fminf(ex * r, 1.f);
    float qy = fminf(ey * r, 1.f - qx);
    // l = theta + logsumexp = V[i,j] - V[i-1,j-1]   (nw.py:59-60)
    const float l = fast_lg2(S) + fmaf(th, kLog2e, mx);
    float hn = l - v;
    float vn = l - hup;
    if (EDGE || SWM) {
        // outside the lattice (or below the sw.py origin) V is 0, so are its differences;
        // Q there is only stored (as zeros) for sw.py's first row / column
        hn = comp ? hn : 0.f;
        vn = comp ? vn : 0.f;
        if (SWM) {
            qx = comp ? qx : kQZeroMark;             // Q == 0 below the sw.py origin: the mark
            qy = comp ? qy : kQZeroMark;
        }
    }
    if (DBG & 4) {
        // score-only forward (Vt alone, deepblast/alignment.py:127-137 under no_grad): no Q is
        // written and the compiler drops the probability arithmetic with the stores
    } else if (DBG & 1) {
        if (qx + qy == 12345.f) qp[0] = qx;          // keeps the math alive, never true
    } else if (!EDGE || store) {
        qp[0] = qx;
        qp[kQY] = qy;
    }
    v = vn;
    return hn;
}

template <bool SWM, int RING, int DBG = 0>
__global__ void __launch_bounds__(256) softdp_fwd2_kernel(const __grid_constant__ CUtensorMap tm_theta,
                                                          const __grid_constant__ CUtensorMap tm_A, FwdParams p) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const int W = blockDim.x >> 5, w = threadIdx.x >> 5, t = threadIdx.x & 31;
    const int NB = W + 1;
    const int Mcap = p.d.M;
    const int g = t >> 4, tp = t & 15;

    constexpr int kF2Ring = RING;
    constexpr int kF2WarpBytes = RING * kF2SlotBytes;
    unsigned char* ring = smem_raw + (size_t)w * kF2WarpBytes;
    size_t off = (size_t)W * kF2WarpBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + off) + w * kF2Ring;
    off += (size_t)W * kF2Ring * 8;
    off = (off + 15) & ~(size_t)15;
    unsigned long long* prog = reinterpret_cast<unsigned long long*>(smem_raw + off);
    unsigned long long* fin = prog + NB;      // per-warp finished-strip counters (run-ahead gate)
    off += (size_t)(NB + W) * 8;
    off = (off + 15) & ~(size_t)15;
    float* bnd = reinterpret_cast<float*>(smem_raw + off);

    if (t == 0) {
        for (int s = 0; s < kF2Ring; ++s) mbar_init(&bars[s], 1);
    }
    if ((int)threadIdx.x < NB) prog[threadIdx.x] = ~0ull;
    if ((int)threadIdx.x < W) fin[threadIdx.x] = 0ull;
    fence_mbar_init();
    __syncthreads();
    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tm_theta);
        tma_prefetch_desc(&tm_A);
    }

    // lane-constant part of the tile address: group, row, and the -4*tp column skew
    const int lanebase = g * 2048 + tp * 60;
    // 16 zeros: h of the "row above" of a pair's first strip (V[0, .] = 0)
    float* zero_row = bnd + (size_t)NB * Mcap;
    if (threadIdx.x < 16) zero_row[threadIdx.x] = 0.f;
    __syncthreads();

    Strip cur, nxt;
    strip_first(cur, p.d, w, W);
    nxt = cur;
    if (cur.valid) strip_next(nxt, p.d, w, W);

    TilePipe<kF2Ring, kF2Ring - 2> pipe;
    pipe.reset();

    // event e of a strip = {group 0: tile e, group 1: tile e-1} x {theta, A}
    auto issue = [&](const Strip& st, int e, unsigned slot) {
        if (!(DBG & 2) && elect_one()) {
            const int T16 = (st.m + kG - 1) / kG;
            unsigned char* dst = ring + slot * kF2SlotBytes;
            const bool g0 = e < T16, g1 = e >= 1;
            mbar_expect_tx(&bars[slot], (g0 ? 2048u : 0u) + (g1 ? 2048u : 0u));
            if (g0) {
                tma_load_3d(dst, &tm_theta, &bars[slot], e * kG, st.k * kTile, st.pair);
                tma_load_3d(dst + 1024, &tm_A, &bars[slot], e * kG, st.k * kTile, st.pair);
            }
            if (g1) {
                tma_load_3d(dst + 2048, &tm_theta, &bars[slot], (e - 1) * kG, st.k * kTile + kG, st.pair);
                tma_load_3d(dst + 3072, &tm_A, &bars[slot], (e - 1) * kG, st.k * kTile + kG, st.pair);
            }
        }
    };

    while (cur.valid) {
        strip_gate(fin, cur.q, w, W);
        const int n = cur.n, m = cur.m, k = cur.k;
        const int T16 = (m + kG - 1) / kG;
        const int NE = T16 + 1;                       // events of this strip
        const int NEn = nxt.valid ? (nxt.m + kG - 1) / kG + 1 : 0;
        const int NBk = (m + 31 + 15) / 16;           // blocks: steps s' = 0 .. m+30
        const int i = k * kTile + t + 1;
        const bool row_ok = i <= n;
        const bool rowcomp = row_ok && i >= p.i0;
        const bool full_rows = (k + 1) * kTile <= n;
        const bool has_up = k > 0;
        const bool feeds_down = (k + 1 < cur.K);
        const unsigned q = cur.q;
        const float* bnd_r = bnd + (size_t)((q + NB - 1) % NB) * Mcap;
        float* bnd_w = bnd + (size_t)(q % NB) * Mcap;
        const unsigned long long* prog_r = prog + ((q + NB - 1) % NB);
        unsigned long long* prog_w = prog + (q % NB);

        float v = 0.f, h = 0.f;                       // differences, log2 units
        float acc_hi = 0.f, acc_lo = 0.f;             // sum_j h[i, j] of the lane's row
        // cell (i, j = c+1), c = s' - t, is wavefront step sigma = s' of strip k
        float* qp = p.Q + (long long)cur.pair * p.ql.pair_stride + (long long)k * p.ql.strip_stride + t;
        float* bw31 = bnd_w - 31;

        int avail = 0;
        unsigned slotA = 0, slotB = 0;
        for (int b = 0; b < NBk; ++b) {
            __syncwarp();
            slotA = slotB;
            if (b < NE) {
                pipe.pump(b, NE, nxt.valid, NEn,
                          [&](bool from_next, int e, unsigned slot) { issue(from_next ? nxt : cur, e, slot); });
                if (DBG & 2) {
                    slotB = pipe.wslot;
                    pipe.wslot = (pipe.wslot + 1 == kF2Ring) ? 0u : pipe.wslot + 1;
                } else {
                    slotB = pipe.wait(bars);
                }
            }
            const int s0 = b * 16;
            if (W > 1 && has_up && s0 < m) {
                const int need = min(s0 + 16, m);
                if (avail < need) avail = progress_wait(prog_r, q - 1, need);
            }
            const float* baseA = reinterpret_cast<const float*>(ring + slotA * kF2SlotBytes + lanebase + 64);
            const float* baseB = reinterpret_cast<const float*>(ring + slotB * kF2SlotBytes + lanebase);
            const bool steady = full_rows && b >= 2 && s0 + 16 <= m && !(SWM && k == 0);
            float part = 0.f;
            if (steady) {
                const float* br = has_up ? bnd_r + s0 : zero_row;
                float* bw = bw31 + s0;
                // stage the block's operands in registers up front: 16 theta, 16 A and the 16 boundary
                // values (uniform address, broadcast) -- one shared-memory latency per block
                // instead of one per step on the dependent chain
                float th_[16], a_[16], bv_[16];
#pragma unroll
                for (int ss = 0; ss < 16; ++ss) {
                    const float* tb = (tp <= ss) ? baseB : baseA;
                    th_[ss] = tb[ss];
                    a_[ss] = tb[ss + 256];
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
                    float hup = __shfl_up_sync(kFull, h, 1);
                    hup = (t == 0) ? bv_[ss] : hup;
                    h = fwd2_step<false, false, DBG>(th_[ss], a_[ss], hup, v, qp + ss * kStepFloats, true, true);
                    part += h;
                    if (t == 31 && feeds_down) bw[ss] = h;
                }
                qp += 16 * kStepFloats;
            } else {
                // ramp blocks: the steady step with the lattice-membership selects
                const bool cap = (k + 1 == cur.K) && (((m - 1 + ((n - 1) & 31)) >> 4) == b);
                const float* br = has_up ? bnd_r + s0 : zero_row;
                const int brlim = has_up ? m - s0 : 16;      // entries of br that exist
#pragma unroll 4
                for (int ss = 0; ss < 16; ++ss) {
                    const int c = s0 + ss - t;
                    float hup = __shfl_up_sync(kFull, h, 1);
                    if (t == 0) hup = br[ss < brlim ? ss : 0];
                    const bool in = row_ok && (unsigned)c < (unsigned)m;
                    const bool comp = SWM ? (in && rowcomp && (c + 1) >= 2) : in;
                    const float* tb = (tp <= ss) ? baseB : baseA;
                    const float th = tb[ss];
                    const float a = tb[ss + 256];
                    h = fwd2_step<true, SWM, DBG>(th, a, hup, v, qp, in, comp);
                    part += h;
                    if (t == 31 && feeds_down && in) bnd_w[c] = h;
                    // Vt = V[n, m] = ln 2 * sum_j h[n, j]
                    if (cap && in && i == n && c == m - 1) p.Vt[cur.pair] = (acc_hi + (acc_lo + part)) * kLn2;
                    qp += kStepFloats;
                }
            }
            {
                // fold the block's partial row sum into the two-float accumulator (Fast2Sum)
                const float t1 = part + acc_lo;
                const float nh = acc_hi + t1;
                acc_lo = t1 - (nh - acc_hi);
                acc_hi = nh;
            }
            if (W > 1 && feeds_down && t == 31) {
                const int done = min(max(s0 + 16 - 31, 0), m);
                if (done > 0) st_release_u64(prog_w, ((unsigned long long)q << 32) | (unsigned)done);
            }
        }
        pipe.next_strip(NE);
        strip_done(fin, cur.q, w, W);
        cur = nxt;
        if (cur.valid) strip_next(nxt, p.d, w, W);
    }
}

}  // namespace b200dp
