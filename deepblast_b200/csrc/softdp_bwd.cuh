// softdp_bwd.cuh -- backward sweep E = dVt/dV (reference:
// deepblast/nw.py:120-135 _backward_pass_numba, deepblast/sw.py:100-115 with the
// loops stopping at 2; GPU counterpart being replaced: deepblast/nw_cuda.py:82-102).
//
// Push form of the reference's pull recurrence: once E[i,j] is known the cell emits
//   X = Q[i,j,x] E[i,j] -> (i-1, j),  D = Q[i,j,m] E[i,j] -> (i-1, j-1),
//   Y = Q[i,j,y] E[i,j] -> (i, j-1),
// so every lane multiplies by ITS OWN Q[i,j,:] (strip-major: the sweep reads Q back in
// exactly the reverse of the order the forward wrote it, 4 KB per 1-D bulk TMA copy)
// instead of its successors'.
// Lane t owns row 32kb+t+1 and walks right to left, lane 31 leading; the value a
// lane hands upward is Z = X(this step) + D(previous step), one shuffle per step.
// E is staged per 32x32 tile in shared memory and drained row-major, coalesced.
#pragma once
#include "softdp_pipes.cuh"

namespace b200dp {

struct BwdParams {
    const float* Et;      // [B], stride et_stride (0 for an expanded scalar)
    long long et_stride;
    const float* Q;       // strip-major storage base
    const float* QdE;     // adjoint backward only (softdp_bwd3_kernel<.., ADJ>): strip-major Qd * E
    float* E;             // [B, N+2, M+2] row-major (adjoint backward: Ed)
    float* Ei;            // optional second copy of the interior, contiguous [B, N, M] (softdp_bwd3 only):
                          // what the adjoint forward sweep reads through TMA if a double backward follows
    PairDims d;
    QLayout ql;
    int i0;
    int flags;
};

constexpr int kBwdWarpBytes = kDiagRing * kDiagElems * 4 + 2 * kTileElems * 4;   // Q ring + E staging

__host__ __device__ inline size_t bwd_smem_bytes(int W, int M) {
    size_t b = (size_t)W * kBwdWarpBytes;
    b += (size_t)W * kDiagRing * 8;
    b = (b + 15) & ~(size_t)15;
    b += (size_t)(2 * W + 1) * 8;
    b = (b + 15) & ~(size_t)15;
    b += (size_t)(W + 1) * (size_t)M * 4;          // boundary rows (Z)
    return b;
}

template <bool kTMA>
__global__ void __launch_bounds__(256) softdp_bwd_kernel(BwdParams p) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const int W = blockDim.x >> 5, w = threadIdx.x >> 5, t = threadIdx.x & 31;
    const int NB = W + 1;
    const int Mcap = p.d.M;

    float* qring = reinterpret_cast<float*>(smem_raw) + (size_t)w * (kBwdWarpBytes / 4);
    float* etile = qring + kDiagRing * kDiagElems;
    size_t off = (size_t)W * kBwdWarpBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + off) + w * kDiagRing;
    off += (size_t)W * kDiagRing * 8;
    off = (off + 15) & ~(size_t)15;
    unsigned long long* prog = reinterpret_cast<unsigned long long*>(smem_raw + off);
    unsigned long long* fin = prog + NB;      // per-warp finished-strip counters (run-ahead gate)
    off += (size_t)(NB + W) * 8;
    off = (off + 15) & ~(size_t)15;
    float* bnd = reinterpret_cast<float*>(smem_raw + off);

    if (t == 0) {
        for (int s = 0; s < kDiagRing; ++s) mbar_init(&bars[s], kTMA ? 1 : 32);
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

    TilePipe<kDiagRing, kDiagRing - 1> pipe;
    pipe.reset();

    // tile a of a strip covers sweep steps [16a, 16a+16); sweep step s touches wavefront
    // step sigma = m + 30 - s, so the tile is sigma in [m+15-16a, m+30-16a], stored with
    // sigma ascending: sweep step s reads tile row 15 - (s & 15).
    auto issue = [&](const Strip& st, int a, unsigned slot) {
        const int kb = st.K - 1 - st.k;
        const float* strip = p.Q + (long long)st.pair * p.ql.pair_stride + (long long)kb * p.ql.strip_stride;
        q_tile_load<kTMA>(qring + slot * kDiagElems, &bars[slot], strip, st.m + 15 - kDiagRows * a, t);
        if (!kTMA) cp_async_mbar_arrive_noinc(&bars[slot]);
    };

    while (cur.valid) {
        strip_gate(fin, cur.q, w, W);
        const int n = cur.n, m = cur.m;
        const int kb = cur.K - 1 - cur.k;           // row block, processed bottom-up
        const int Ta = (m + 31 + kDiagRows - 1) / kDiagRows;
        const int Tn = nxt.valid ? (nxt.m + 31 + kDiagRows - 1) / kDiagRows : 0;
        const int i = kb * kTile + t + 1;
        const bool row_ok = i <= n;
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
        unsigned dslot = 0;
        float zout = 0.f;      // X(this lane, last step) + D(this lane, two steps ago)
        float dprev = 0.f;     // D of the last step
        float yprev = 0.f;     // Y of the last step

        for (int s = 0; s <= m + 30; ++s) {
            if ((s & (kDiagRows - 1)) == 0) {
                const int a = s / kDiagRows;
                __syncwarp();
                pipe.pump(a, Ta, nxt.valid, Tn,
                          [&](bool from_next, int ti, unsigned slot) { issue(from_next ? nxt : cur, ti, slot); });
                dslot = pipe.wait(bars);
            }
            const int c = m - 1 - (s - u);          // 0-based lattice column of this lane
            // lane 31 needs Z[c31] from the strip below, c31 = m-1-s: (s+1) entries published
            if (has_below && s < m && avail < s + 1) avail = progress_wait(prog_r, q - 1, s + 1);

            float zin = __shfl_down_sync(kFull, zout, 1);
            if (t == 31) {
                zin = 0.f;
                if (has_below && c >= 0 && c < m) zin = bnd_r[c];
            }
            const bool in = row_ok && c >= 0 && c < m;
            const bool comp = in && i >= p.i0 && (c + 1) >= p.i0;
            float e = 0.f, X = 0.f, D = 0.f, Y = 0.f;
            if (comp) {
                const float* qt = qring + dslot * kDiagElems + (kDiagRows - 1 - (s & (kDiagRows - 1))) * kStepFloats + t;
                e = zin + yprev;
                if (i == n && c == m - 1) e = et;   // E[n, m] = Et (nw.py:125-127)
                // q_m = (1 - q_x) - q_y is implied (>= 0: the forward stores q_y <= 1 - q_x)
                const float qx = qt[0], qy = qt[kQY];
                X = qx * e;
                Y = qy * e;
                D = ((1.f - qx) - qy) * e;
            }
            if (c >= 0 && c < m) etile[((c >> 5) & 1) * kTileElems + t * kTile + (c & 31)] = e;
            zout = X + dprev;
            dprev = D;
            yprev = Y;
            if (t == 0 && feeds_up && c >= 0 && c < m) {
                bnd_w[c] = zout;
                const int done = m - c;
                if ((done & 7) == 0 || c == 0)
                    st_release_u64(prog_w, ((unsigned long long)q << 32) | (unsigned)done);
            }
            // lane 0 (the trailing lane) just finished column c0: drain a complete tile
            const int c0 = m + 30 - s;
            if (c0 >= 0 && (c0 & 31) == 0) {
                __syncwarp();
                const int tc = c0 >> 5;
                const float* src = etile + (tc & 1) * kTileElems;
                const int col = tc * kTile + t;
                if (col < m) {
                    float* dstp = Eb + (long long)(kb * kTile + 1) * (M + 2) + col + 1;
                    const int rmax = min(kTile, n - kb * kTile);
                    for (int r = 0; r < rmax; ++r) dstp[(long long)r * (M + 2)] = src[r * kTile + t];
                }
                __syncwarp();
            }
        }
        if (!varlen) {
            // zero borders of the padded tensor (nw.py:347 allocates zeros)
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
            Eb[(long long)(N + 1) * (M + 2) + M + 1] = et;   // caller pre-zeroed E
        }
        pipe.next_strip(Ta);
        strip_done(fin, cur.q, w, W);
        cur = nxt;
        if (cur.valid) strip_next(nxt, p.d, w, W);
    }
}

}  // namespace b200dp
