// softdp_bwd2.cuh -- backward sweep, fast path (reference: deepblast/nw.py:120-135,
// sw.py:100-115; replaces deepblast/nw_cuda.py:82-102).
//
// Same push-form wavefront as softdp_bwd.cuh, re-organised like softdp_fwd2.cuh:
// 16-step blocks, fully unrolled predicate-free "steady" blocks (12 instructions per
// step: shuffle, 3 LDS of the lane's own Q from the bulk-TMA tile, 5 FP ops, 1 STS), the
// ramps through the predicated "edge" variant.  E is staged in shared memory in
// STEP-major order (row = step mod 80, pitch 33 floats) so the hot loop stores with
// immediate offsets and no bank conflicts; complete 32-column tiles are drained
// row-major (skewed, conflict-free read; coalesced 128-byte global stores).
#pragma once
#include "softdp_bwd.cuh"

namespace b200dp {

constexpr int kB2Ring = 3;                     // Q tiles: 1 live + 2 in flight (32 steps of lead)
constexpr int kB2StageSteps = 80;              // 62 steps of tile history + one 16-step block
constexpr int kB2StagePitch = 33;
constexpr int kB2StageFloats = kB2StageSteps * kB2StagePitch;
constexpr int kB2WarpBytes = kB2Ring * kDiagElems * 4 + ((kB2StageFloats * 4 + 127) / 128) * 128;

__host__ __device__ inline size_t bwd2_smem_bytes(int W, int M) {
    size_t b = (size_t)W * kB2WarpBytes;
    b += (size_t)W * kB2Ring * 8;
    b = (b + 15) & ~(size_t)15;
    b += (size_t)(2 * W + 1) * 8;
    b = (b + 15) & ~(size_t)15;
    b += (size_t)(W + 1) * (size_t)M * 4;
    b += 256;                                      // slack for the ramps' clamped boundary reads
    return b;
}

// Drain column tile tc (32 columns) of a strip from the step-major staging ring to the
// row-major E tensor: lane = column, 8 independent LDS then 8 coalesced 128-byte
// row-segment stores at a time.  Element (r, col) was produced at step (m-1-col) + (31-r).
__device__ __forceinline__ void bwd2_drain_tile(const float* __restrict__ stage, float* __restrict__ Erow0, int tc,
                                                int m, int rmax, int pitch, int t) {
    const int col = tc * kTile + t;
    if (col >= m) return;
    float* dstp = Erow0 + col + 1;
    const int sr0 = (m - 1 - col + 31) % kB2StageSteps;
#pragma unroll 1
    for (int r0 = 0; r0 < rmax; r0 += 8) {
        float v[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            int sr = sr0 - (r0 + q);
            sr += (sr < 0) ? kB2StageSteps : 0;
            v[q] = stage[sr * kB2StagePitch + r0 + q];
        }
#pragma unroll
        for (int q = 0; q < 8; ++q)
            if (r0 + q < rmax) dstp[(long long)(r0 + q) * pitch] = v[q];
    }
}

template <bool SWM>
__global__ void __launch_bounds__(256) softdp_bwd2_kernel(BwdParams p) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const int W = blockDim.x >> 5, w = threadIdx.x >> 5, t = threadIdx.x & 31;
    const int NB = W + 1;
    const int Mcap = p.d.M;

    float* qring = reinterpret_cast<float*>(smem_raw + (size_t)w * kB2WarpBytes);
    float* stage = qring + kB2Ring * kDiagElems;
    size_t off = (size_t)W * kB2WarpBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + off) + w * kB2Ring;
    off += (size_t)W * kB2Ring * 8;
    off = (off + 15) & ~(size_t)15;
    unsigned long long* prog = reinterpret_cast<unsigned long long*>(smem_raw + off);
    unsigned long long* fin = prog + NB;      // per-warp finished-strip counters (run-ahead gate)
    off += (size_t)(NB + W) * 8;
    off = (off + 15) & ~(size_t)15;
    float* bnd = reinterpret_cast<float*>(smem_raw + off);

    if (t == 0) {
        for (int s = 0; s < kB2Ring; ++s) mbar_init(&bars[s], 1);
    }
    if ((int)threadIdx.x < NB) prog[threadIdx.x] = ~0ull;
    if ((int)threadIdx.x < W) fin[threadIdx.x] = 0ull;
    fence_mbar_init();
    __syncthreads();

    const int N = p.d.N, M = p.d.M;
    const bool varlen = (p.d.xlen != nullptr) || (p.d.ylen != nullptr);
    const int u = 31 - t;

    Strip cur, nxt;
    strip_first(cur, p.d, w, W);
    nxt = cur;
    if (cur.valid) strip_next(nxt, p.d, w, W);

    TilePipe<kB2Ring, kB2Ring - 1> pipe;
    pipe.reset();

    // tile a covers sweep steps [16a, 16a+16) = wavefront steps sigma in
    // [m+15-16a, m+30-16a] of the strip, 4 KB contiguous; sweep step s reads row 15-(s&15)
    auto issue = [&](const Strip& st, int a, unsigned slot) {
        const int kb = st.K - 1 - st.k;
        const float* strip = p.Q + (long long)st.pair * p.ql.pair_stride + (long long)kb * p.ql.strip_stride;
        q_tile_load<true>(qring + slot * kDiagElems, &bars[slot], strip, st.m + 15 - kDiagRows * a, t);
    };

    while (cur.valid) {
        strip_gate(fin, cur.q, w, W);
        const int n = cur.n, m = cur.m;
        const int kb = cur.K - 1 - cur.k;             // row block, processed bottom-up
        const int NBk = (m + 31 + 15) / 16;           // blocks == Q tiles: steps 0 .. m+30
        const int NBn = nxt.valid ? (nxt.m + 31 + 15) / 16 : 0;
        const int i = kb * kTile + t + 1;
        const bool row_ok = i <= n;
        const bool rowcomp = row_ok && i >= p.i0;
        const bool full_rows = (kb + 1) * kTile <= n;
        const bool has_below = cur.k > 0;
        const bool feeds_up = kb > 0;
        const unsigned q = cur.q;
        const float* bnd_r = bnd + (size_t)((q + NB - 1) % NB) * Mcap;
        float* bnd_w = bnd + (size_t)(q % NB) * Mcap;
        const unsigned long long* prog_r = prog + ((q + NB - 1) % NB);
        unsigned long long* prog_w = prog + (q % NB);
        float* Eb = p.E + (long long)cur.pair * (N + 2) * (M + 2);
        const float et = p.Et[(long long)cur.pair * p.et_stride];

        int avail = 0;
        float zout = 0.f, dprev = 0.f, yprev = 0.f;
        int next_drain = (m - 1) >> 5;                // highest column tile not yet drained
        int srow = 0;                                 // (16 b) mod kB2StageSteps

        for (int b = 0; b < NBk; ++b) {
            __syncwarp();
            pipe.pump(b, NBk, nxt.valid, NBn,
                      [&](bool from_next, int a, unsigned slot) { issue(from_next ? nxt : cur, a, slot); });
            const unsigned dslot = pipe.wait(bars);
            const int s0 = b * 16;
            // drain every column tile completed before this block: tile tc is complete
            // once lane 0 has passed column 32 tc, i.e. after step m + 30 - 32 tc
            while (next_drain >= 0 && (m + 30 - 32 * next_drain) < s0) {
                bwd2_drain_tile(stage, Eb + (long long)(kb * kTile + 1) * (M + 2), next_drain, m,
                                min(kTile, n - kb * kTile), M + 2, t);
                next_drain--;
            }
            __syncwarp();
            if (W > 1 && has_below && s0 < m) {
                const int need = min(s0 + 16, m);
                if (avail < need) avail = progress_wait(prog_r, q - 1, need);
            }
            const float* qt = qring + dslot * kDiagElems + t;
            float* st = stage + srow * kB2StagePitch + t;
            // lane 31's column at step s is m-1-s; lane 0's is m+30-s
            const bool steady = full_rows && s0 >= 32 && s0 + 15 <= m - 1 - (SWM ? 1 : 0) && !(SWM && kb == 0);
            if (steady) {
                const float* br = bnd_r + (m - 1 - s0);
                float* bw = bnd_w + (m + 30 - s0);
#pragma unroll
                for (int ss = 0; ss < 16; ++ss) {
                    float zin = __shfl_down_sync(kFull, zout, 1);
                    if (t == 31) zin = has_below ? br[-ss] : 0.f;
                    const float e = zin + yprev;
                    const float qx = qt[(15 - ss) * kStepFloats], qy = qt[(15 - ss) * kStepFloats + kQY];
                    const float X = qx * e;
                    const float Y = qy * e;
                    const float D = ((1.f - qx) - qy) * e;    // implied q_m (>= 0 by the forward's clamp)
                    st[ss * kB2StagePitch] = e;
                    zout = X + dprev;
                    dprev = D;
                    yprev = Y;
                    if (t == 0 && feeds_up) bw[-ss] = zout;
                }
            } else {
                // ramp blocks: the steady step with the lattice-membership selects.  Q in
                // the ramps is never written by the forward (arbitrary bits), so the
                // products are selected, not multiplied by zero.
                const bool seed_blk = (cur.k == 0) && (s0 < 32);   // E[n, m] = Et lives here
#pragma unroll 4
                for (int ss = 0; ss < 16; ++ss) {
                    const int c = m - 1 - (s0 + ss - u);
                    float zin = __shfl_down_sync(kFull, zout, 1);
                    if (t == 31) zin = (has_below && c >= 0) ? bnd_r[c] : 0.f;
                    const bool in = row_ok && (unsigned)c < (unsigned)m;
                    const bool comp = SWM ? (in && rowcomp && (c + 1) >= 2) : in;
                    float e = zin + yprev;
                    if (seed_blk && i == n && c == m - 1) e = et;   // nw.py:125-127
                    e = comp ? e : 0.f;
                    const float qx = qt[(15 - ss) * kStepFloats], qy = qt[(15 - ss) * kStepFloats + kQY];
                    const float X = comp ? qx * e : 0.f;
                    const float Y = comp ? qy * e : 0.f;
                    const float D = comp ? ((1.f - qx) - qy) * e : 0.f;
                    st[ss * kB2StagePitch] = e;
                    zout = X + dprev;
                    dprev = D;
                    yprev = Y;
                    if (t == 0 && feeds_up && in) bnd_w[c] = zout;
                }
            }
            if (W > 1 && feeds_up && t == 0) {
                // lane 0 has finished columns m-1 .. m+30-(s0+15): count from the right
                const int done = min(max(s0 + 16 - 31, 0), m);
                if (done > 0) st_release_u64(prog_w, ((unsigned long long)q << 32) | (unsigned)done);
            }
            srow += 16;
            if (srow == kB2StageSteps) srow = 0;
        }
        __syncwarp();
        // drain what is left (all steps are done)
        while (next_drain >= 0) {
            bwd2_drain_tile(stage, Eb + (long long)(kb * kTile + 1) * (M + 2), next_drain, m,
                            min(kTile, n - kb * kTile), M + 2, t);
            next_drain--;
        }
        if (!varlen) {
            if (row_ok) {
                Eb[(long long)i * (M + 2)] = 0.f;
                Eb[(long long)i * (M + 2) + M + 1] = 0.f;
            }
            if (kb == 0)
                for (int col = t; col < M + 2; col += 32) Eb[col] = 0.f;
            if (cur.k == 0)
                for (int col = t; col < M + 2; col += 32)
                    Eb[(long long)(N + 1) * (M + 2) + col] = (col == M + 1) ? et : 0.f;
        } else if (cur.k == 0 && t == 0) {
            Eb[(long long)(N + 1) * (M + 2) + M + 1] = et;
        }
        pipe.next_strip(NBk);
        strip_done(fin, cur.q, w, W);
        cur = nxt;
        if (cur.valid) strip_next(nxt, p.d, w, W);
    }
}

}  // namespace b200dp
