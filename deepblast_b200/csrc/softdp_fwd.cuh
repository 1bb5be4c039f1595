// softdp_fwd.cuh -- forward fill of the soft-DP lattice (reference:
// deepblast/nw.py:46-62 _forward_pass_numba, deepblast/sw.py:46-62 with the loops
// starting at 2; GPU counterpart being replaced: deepblast/nw_cuda.py:46-79).
//
// Wavefront: lane t of a warp owns lattice row 32k+t+1 of strip k and walks it left
// to right one column per step, one step behind lane t-1, so at any step the warp
// sits on one anti-diagonal.  V[i-1,j] arrives by one shuffle, V[i-1,j-1] is the
// previous step's shuffled value, V[i,j-1] is the lane's own previous value.  V is
// carried as an fp32 (hi, lo) pair and never stored: only differences of
// neighbouring V enter exp/log, and those are formed exactly (SURVEY.md section 7:
// the oracle accumulates V in fp64, nw.py:49).
#pragma once
#include "softdp_pipes.cuh"

namespace b200dp {

struct FwdParams {
    const float* theta;   // [B,N,M]
    const float* A;       // [B,N,M]
    float* Q;             // strip-major storage base
    float* Vt;            // [B]
    PairDims d;
    QLayout ql;
    int i0;               // 1 = Needleman-Wunsch, 2 = "Smith-Waterman" (sw.py:54-55)
    int flags;
    const float* Qin;     // softdp_fwd3 ADJ (adjoint forward): the forward's Q; `Q` is then the output Qd (* E)
    int has_za, has_e;    // ... ADJ: ZA / E operands present (else 0 / 1)
    int pf_tiles;         // softdp_fwd3: L2 prefetch box width in 16-column tiles (0 = off)
    int pf_dist;          // ... issued this many tiles ahead of the leading TMA tile
};

constexpr int kFwdWarpBytes = 2 * kRowRing * kTileElems * 4;   // theta ring + A ring

__host__ __device__ inline size_t fwd_smem_bytes(int W, int M) {
    size_t b = (size_t)W * kFwdWarpBytes;
    b += (size_t)W * kRowRing * 8;                 // mbarriers
    b = (b + 15) & ~(size_t)15;
    b += (size_t)(2 * W + 1) * 8;                      // progress words
    b = (b + 15) & ~(size_t)15;
    b += (size_t)(W + 1) * (size_t)M * 8;          // boundary rows (hi, lo)
    return b;
}

template <bool kTMA>
__global__ void __launch_bounds__(256) softdp_fwd_kernel(const __grid_constant__ CUtensorMap tm_theta,
                                                         const __grid_constant__ CUtensorMap tm_A, FwdParams p) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const int W = blockDim.x >> 5, w = threadIdx.x >> 5, t = threadIdx.x & 31;
    const int NB = W + 1;
    const int Mcap = p.d.M;

    float* tiles = reinterpret_cast<float*>(smem_raw) + (size_t)w * (kFwdWarpBytes / 4);
    size_t off = (size_t)W * kFwdWarpBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + off) + w * kRowRing;
    off += (size_t)W * kRowRing * 8;
    off = (off + 15) & ~(size_t)15;
    unsigned long long* prog = reinterpret_cast<unsigned long long*>(smem_raw + off);
    unsigned long long* fin = prog + NB;      // per-warp finished-strip counters (run-ahead gate)
    off += (size_t)(NB + W) * 8;
    off = (off + 15) & ~(size_t)15;
    float2* bnd = reinterpret_cast<float2*>(smem_raw + off);

    if (t == 0) {
        for (int s = 0; s < kRowRing; ++s) mbar_init(&bars[s], kTMA ? 1 : 32);
    }
    if ((int)threadIdx.x < NB) prog[threadIdx.x] = ~0ull;
    if ((int)threadIdx.x < W) fin[threadIdx.x] = 0ull;
    fence_mbar_init();
    __syncthreads();
    if (kTMA && threadIdx.x == 0) {
        tma_prefetch_desc(&tm_theta);
        tma_prefetch_desc(&tm_A);
    }

    const RowSrc s_theta{p.theta, (long long)p.d.N * p.d.M, p.d.M, p.d.N, p.d.M};
    const RowSrc s_A{p.A, (long long)p.d.N * p.d.M, p.d.M, p.d.N, p.d.M};
    Strip cur, nxt;
    strip_first(cur, p.d, w, W);
    nxt = cur;
    if (cur.valid) strip_next(nxt, p.d, w, W);

    TilePipe<kRowRing, kRowRing - 2> pipe;
    pipe.reset();

    auto issue = [&](const Strip& st, int tq, unsigned slot) {
        float* dth = tiles + slot * kTileElems;
        float* dA = tiles + (kRowRing + slot) * kTileElems;
        if (kTMA) {
            if (t == 0) {
                mbar_expect_tx(&bars[slot], 2 * kTileElems * 4);
                tma_load_3d(dth, &tm_theta, &bars[slot], tq * kTile, st.k * kTile, st.pair);
                tma_load_3d(dA, &tm_A, &bars[slot], tq * kTile, st.k * kTile, st.pair);
            }
        } else {
            row_tile_load_generic(dth, s_theta, st.pair, st.k, tq, t);
            row_tile_load_generic(dA, s_A, st.pair, st.k, tq, t);
            cp_async_mbar_arrive_noinc(&bars[slot]);
        }
    };

    while (cur.valid) {
        strip_gate(fin, cur.q, w, W);
        const int n = cur.n, m = cur.m, k = cur.k;
        const int T = (m + kTile - 1) / kTile;
        const int Tn = nxt.valid ? (nxt.m + kTile - 1) / kTile : 0;
        const int i = k * kTile + t + 1;            // padded lattice row of this lane
        const bool row_ok = i <= n;
        const bool has_up = k > 0;
        const bool feeds_down = (k + 1 < cur.K);
        const unsigned q = cur.q;
        const float2* bnd_r = bnd + (size_t)((q + NB - 1) % NB) * Mcap;   // boundary q-1
        float2* bnd_w = bnd + (size_t)(q % NB) * Mcap;                   // boundary q
        const unsigned long long* prog_r = prog + ((q + NB - 1) % NB);
        unsigned long long* prog_w = prog + (q % NB);

        int avail = 0;
        unsigned lslot = pipe.wslot;                 // slot of this strip's tile 0
        float vh = 0.f, vl = 0.f;                    // V[i, j-1]   (own previous)
        float dh = 0.f, dl = 0.f;                    // V[i-1, j-1] (previous shuffle)
        // cell (i, j), j = s - t, is wavefront step sigma = s - 1 of strip k
        float* qp = p.Q + (long long)cur.pair * p.ql.pair_stride + (long long)k * p.ql.strip_stride + t - kStepFloats;

        for (int s = 0; s <= m + 31; ++s) {
            if ((s & 31) == 0) {
                const int tq = s >> 5;
                if (tq < T) {
                    __syncwarp();
                    pipe.pump(tq, T, nxt.valid, Tn,
                              [&](bool from_next, int ti, unsigned slot) { issue(from_next ? nxt : cur, ti, slot); });
                    pipe.wait(bars);
                }
            }
            const int j = s - t;
            if (has_up && s >= 1 && s <= m && avail < s) avail = progress_wait(prog_r, q - 1, s);

            float uh = __shfl_up_sync(kFull, vh, 1);
            float ul = __shfl_up_sync(kFull, vl, 1);
            if (t == 0) {
                uh = 0.f;
                ul = 0.f;
                if (has_up && j >= 1 && j <= m) {
                    const float2 b = bnd_r[j - 1];
                    uh = b.x;
                    ul = b.y;
                }
            }
            const bool in = row_ok && j >= 1 && j <= m;
            const bool comp = in && i >= p.i0 && j >= p.i0;
            float qx = 0.f, qy = 0.f, nh = 0.f, nl = 0.f;
            if (comp) {
                const int c = j - 1;
                const int o = (int)lslot * kTileElems + t * kTile + (c & 31);
                const float th = tiles[o];
                const float a = tiles[kRowRing * kTileElems + o];
                // u_x - u_m and u_y - u_m (nw.py:56-58), differences formed in (hi, lo)
                const float dx = ((uh - dh) + (ul - dl)) + a;
                const float dy = ((vh - dh) + (vl - dl)) + a;
                const float mx = fmaxf(fmaxf(dx, dy), 0.f);
                const float ex = fast_ex2((dx - mx) * kLog2e);
                const float em = fast_ex2(-mx * kLog2e);
                const float ey = fast_ex2((dy - mx) * kLog2e);
                const float S = (ex + em) + ey;
                const float r = fast_rcp(S);
                qx = fminf(ex * r, 1.f);
                qy = fminf(ey * r, 1.f - qx);        // keeps the implied q_m = (1 - q_x) - q_y >= 0
                // V[i,j] = theta + V[i-1,j-1] + logsumexp(dx, 0, dy)   (nw.py:59-60)
                const float delta = th + fmaf(fast_lg2(S), kLn2, mx);
                const float t1 = delta + dl;
                nh = dh + t1;
                nl = t1 - (nh - dh);
            }
            if (in) {
                // two stored states; a cell below the sw.py origin (Q == 0) carries the mark
                qp[0] = comp ? qx : kQZeroMark;
                qp[kQY] = comp ? qy : kQZeroMark;
            }
            if (t == 31 && feeds_down && in) {
                bnd_w[j - 1] = make_float2(nh, nl);
                if ((j & 7) == 0 || j == m)
                    st_release_u64(prog_w, ((unsigned long long)q << 32) | (unsigned)j);
            }
            if (in && i == n && j == m) p.Vt[cur.pair] = nh + nl;
            if (j >= 1 && ((j - 1) & 31) == 31) lslot = (lslot + 1 == kRowRing) ? 0u : lslot + 1;
            dh = uh;
            dl = ul;
            vh = nh;
            vl = nl;
            qp += kStepFloats;
        }
        pipe.next_strip(T);
        strip_done(fin, cur.q, w, W);
        cur = nxt;
        if (cur.valid) strip_next(nxt, p.d, w, W);
    }
}

}  // namespace b200dp
