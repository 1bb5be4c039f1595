// softdp_adjoint.cuh -- the Hessian-vector (double-backward) pair of sweeps
// (reference: deepblast/nw.py:178-199 _adjoint_forward_pass_numba and
// nw.py:251-267 _adjoint_backward_pass_numba; sw.py uses the same full-range loops,
// sw.py:150-151,199-201; GPU counterparts replaced: nw_cuda.py:105-165).
//
// Same wavefront / strip / pipe machinery as the forward and backward kernels.
// Ztheta and E are padded row-major tensors whose row pitch (M+2)*4 B is not a
// multiple of 16 B, so they cannot be described by a TMA tensor map; their 32x32
// tiles are staged with 4-byte cp.async on the same mbarriers.  Q and Qd are
// strip-major (two stored states per cell, softdp_common.cuh) and arrive by 1-D bulk TMA,
// 4 KB per copy.
#pragma once
#include "softdp_pipes.cuh"

namespace b200dp {

struct AdjFwdParams {
    const float* Q;        // strip-major
    const float* Ztheta;   // [B, N+2, M+2]
    const float* ZA;       // [B, N, M]
    float* Vtd;            // [B]
    float* Qd;             // strip-major
    PairDims d;
    QLayout ql;
};

constexpr int kAdjFwdWarpBytes = 2 * kRowRing * kTileElems * 4 + kDiagRing * kDiagElems * 4;

__host__ __device__ inline size_t adj_fwd_smem_bytes(int W, int M) {
    size_t b = (size_t)W * kAdjFwdWarpBytes;
    b += (size_t)W * (kRowRing + kDiagRing) * 8;
    b = (b + 15) & ~(size_t)15;
    b += (size_t)(2 * W + 1) * 8;
    b = (b + 15) & ~(size_t)15;
    b += (size_t)(W + 1) * (size_t)M * 8;
    return b;
}

template <bool kTMA>
__global__ void __launch_bounds__(256) softdp_adj_fwd_kernel(AdjFwdParams p) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const int W = blockDim.x >> 5, w = threadIdx.x >> 5, t = threadIdx.x & 31;
    const int NB = W + 1;
    const int Mcap = p.d.M;
    const int N = p.d.N, M = p.d.M;

    float* rtiles = reinterpret_cast<float*>(smem_raw) + (size_t)w * (kAdjFwdWarpBytes / 4);
    float* qring = rtiles + 2 * kRowRing * kTileElems;
    size_t off = (size_t)W * kAdjFwdWarpBytes;
    uint64_t* rbars = reinterpret_cast<uint64_t*>(smem_raw + off) + w * (kRowRing + kDiagRing);
    uint64_t* qbars = rbars + kRowRing;
    off += (size_t)W * (kRowRing + kDiagRing) * 8;
    off = (off + 15) & ~(size_t)15;
    unsigned long long* prog = reinterpret_cast<unsigned long long*>(smem_raw + off);
    unsigned long long* fin = prog + NB;      // per-warp finished-strip counters (run-ahead gate)
    off += (size_t)(NB + W) * 8;
    off = (off + 15) & ~(size_t)15;
    float2* bnd = reinterpret_cast<float2*>(smem_raw + off);

    if (t == 0) {
        for (int s = 0; s < kRowRing; ++s) mbar_init(&rbars[s], 32);
        for (int s = 0; s < kDiagRing; ++s) mbar_init(&qbars[s], kTMA ? 1 : 32);
    }
    if ((int)threadIdx.x < NB) prog[threadIdx.x] = ~0ull;
    if ((int)threadIdx.x < W) fin[threadIdx.x] = 0ull;
    fence_mbar_init();
    __syncthreads();

    // lattice cell (r, c) 0-based of Ztheta is padded element (r+1, c+1)
    const RowSrc s_Zt{p.Ztheta + (M + 2) + 1, (long long)(N + 2) * (M + 2), M + 2, N, M};
    const RowSrc s_ZA{p.ZA, (long long)N * M, M, N, M};
    Strip cur, nxt;
    strip_first(cur, p.d, w, W);
    nxt = cur;
    if (cur.valid) strip_next(nxt, p.d, w, W);

    TilePipe<kRowRing, kRowRing - 2> rpipe;
    TilePipe<kDiagRing, kDiagRing - 1> qpipe;
    rpipe.reset();
    qpipe.reset();

    auto issue_row = [&](const Strip& st, int tq, unsigned slot) {
        row_tile_load_generic(rtiles + slot * kTileElems, s_Zt, st.pair, st.k, tq, t);
        row_tile_load_generic(rtiles + (kRowRing + slot) * kTileElems, s_ZA, st.pair, st.k, tq, t);
        cp_async_mbar_arrive_noinc(&rbars[slot]);
    };
    // ascending sweep: step s IS wavefront step sigma = s; tile a = sigma in [16a, 16a+16)
    auto issue_q = [&](const Strip& st, int a, unsigned slot) {
        const float* strip = p.Q + (long long)st.pair * p.ql.pair_stride + (long long)st.k * p.ql.strip_stride;
        q_tile_load<kTMA>(qring + slot * kDiagElems, &qbars[slot], strip, kDiagRows * a, t);
        if (!kTMA) cp_async_mbar_arrive_noinc(&qbars[slot]);
    };

    while (cur.valid) {
        strip_gate(fin, cur.q, w, W);
        const int n = cur.n, m = cur.m, k = cur.k;
        const int T = (m + kTile - 1) / kTile;
        const int Tn = nxt.valid ? (nxt.m + kTile - 1) / kTile : 0;
        const int Ta = (m + 31 + kDiagRows - 1) / kDiagRows;
        const int Tan = nxt.valid ? (nxt.m + 31 + kDiagRows - 1) / kDiagRows : 0;
        const int i = k * kTile + t + 1;
        const bool row_ok = i <= n;
        const bool has_up = k > 0;
        const bool feeds_down = (k + 1 < cur.K);
        const unsigned q = cur.q;
        const float2* bnd_r = bnd + (size_t)((q + NB - 1) % NB) * Mcap;
        float2* bnd_w = bnd + (size_t)(q % NB) * Mcap;
        const unsigned long long* prog_r = prog + ((q + NB - 1) % NB);
        unsigned long long* prog_w = prog + (q % NB);

        int avail = 0;
        unsigned lslot = rpipe.wslot;
        unsigned dslot = 0;
        float vh = 0.f, vl = 0.f, dh = 0.f, dl = 0.f;
        // cell (i, j), j = s - t + 1, is wavefront step sigma = s of strip k
        float* qdp = p.Qd + (long long)cur.pair * p.ql.pair_stride + (long long)k * p.ql.strip_stride + t;

        for (int s = 0; s <= m + 30; ++s) {
            if ((s & 31) == 0 && (s >> 5) < T) {
                __syncwarp();
                rpipe.pump(s >> 5, T, nxt.valid, Tn,
                           [&](bool fn, int ti, unsigned slot) { issue_row(fn ? nxt : cur, ti, slot); });
                rpipe.wait(rbars);
            }
            if ((s & (kDiagRows - 1)) == 0) {
                __syncwarp();
                qpipe.pump(s / kDiagRows, Ta, nxt.valid, Tan,
                           [&](bool fn, int ti, unsigned slot) { issue_q(fn ? nxt : cur, ti, slot); });
                dslot = qpipe.wait(qbars);
            }
            const int j = s - t + 1;
            if (has_up && s < m && avail < s + 1) avail = progress_wait(prog_r, q - 1, s + 1);

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
            float nh = 0.f, nl = 0.f;
            if (in) {
                const int c = j - 1;
                const int o = (int)lslot * kTileElems + t * kTile + (c & 31);
                const float zt = rtiles[o];
                const float za = rtiles[kRowRing * kTileElems + o];
                const float* qt = qring + dslot * kDiagElems + (s & (kDiagRows - 1)) * kStepFloats + t;
                const float qx = qt[0], qy = qt[kQY];
                const float qm = (1.f - qx) - qy;          // the implied state
                // w_x - w_m, w_y - w_m with w = (za + Vd[i-1,j], Vd[i-1,j-1], za + Vd[i,j-1]) (nw.py:188-192)
                const float dxm = ((uh - dh) + (ul - dl)) + za;
                const float dym = ((vh - dh) + (vl - dl)) + za;
                // eps = (qx + qm + qy) - 1, error-free: the stored fp32 Q do not sum to 1 exactly
                const float s1 = qx + qm;
                const float b1 = s1 - qx;
                const float e1 = (qx - (s1 - b1)) + (qm - b1);
                const float s2 = s1 + qy;
                const float b2 = s2 - s1;
                const float e2 = (s1 - (s2 - b2)) + (qy - b2);
                const float eps = (s2 - 1.f) + (e1 + e2);
                const float r = fmaf(qx, dxm, qy * dym);
                // g = tsum - w_m,  tsum = sum_s q_s w_s (nw.py:193-196)
                const float g = fmaf(eps, dh, fmaf(eps, dl, r));
                // nw.py:30-43; qd_m = q_m (-g) = -(qd_x + qd_y) up to rounding and is implied
                float qdx = qx * (dxm - g), qdy = qy * (dym - g);
                if (qx < 0.f) {
                    // marked cell: Q[i,j,:] == 0 (first row/column of the sw.py lattice): Vd = Ztheta
                    nh = zt;
                    nl = 0.f;
                    qdx = qdy = 0.f;
                } else {
                    const float delta = zt + g;
                    const float t1 = delta + dl;
                    nh = dh + t1;
                    nl = t1 - (nh - dh);
                }
                qdp[0] = qdx;
                qdp[kQY] = qdy;
            }
            if (t == 31 && feeds_down && in) {
                bnd_w[j - 1] = make_float2(nh, nl);
                if ((j & 7) == 0 || j == m)
                    st_release_u64(prog_w, ((unsigned long long)q << 32) | (unsigned)j);
            }
            if (in && i == n && j == m) p.Vtd[cur.pair] = nh + nl;
            if (j >= 1 && ((j - 1) & 31) == 31) lslot = (lslot + 1 == kRowRing) ? 0u : lslot + 1;
            dh = uh;
            dl = ul;
            vh = nh;
            vl = nl;
            qdp += kStepFloats;
        }
        rpipe.next_strip(T);
        qpipe.next_strip(Ta);
        strip_done(fin, cur.q, w, W);
        cur = nxt;
        if (cur.valid) strip_next(nxt, p.d, w, W);
    }
}

// ---------------------------------------------------------------------------
struct AdjBwdParams {
    const float* E;        // [B, N+2, M+2]
    const float* Q;        // strip-major
    const float* Qd;       // strip-major
    float* Ed;             // [B, N+2, M+2]
    PairDims d;
    QLayout ql;
};

constexpr int kAdjBwdWarpBytes =
    kRowRing * kTileElems * 4 + 2 * kDiagRing * kDiagElems * 4 + 2 * kTileElems * 4;

__host__ __device__ inline size_t adj_bwd_smem_bytes(int W, int M) {
    size_t b = (size_t)W * kAdjBwdWarpBytes;
    b += (size_t)W * (kRowRing + kDiagRing) * 8;
    b = (b + 15) & ~(size_t)15;
    b += (size_t)(2 * W + 1) * 8;
    b = (b + 15) & ~(size_t)15;
    b += (size_t)(W + 1) * (size_t)M * 4;
    return b;
}

template <bool kTMA>
__global__ void __launch_bounds__(256) softdp_adj_bwd_kernel(AdjBwdParams p) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const int W = blockDim.x >> 5, w = threadIdx.x >> 5, t = threadIdx.x & 31;
    const int NB = W + 1;
    const int Mcap = p.d.M;
    const int N = p.d.N, M = p.d.M;

    float* etiles = reinterpret_cast<float*>(smem_raw) + (size_t)w * (kAdjBwdWarpBytes / 4);
    float* qring = etiles + kRowRing * kTileElems;             // [slot][Q | Qd]
    float* otile = qring + 2 * kDiagRing * kDiagElems;         // Ed staging
    size_t off = (size_t)W * kAdjBwdWarpBytes;
    uint64_t* rbars = reinterpret_cast<uint64_t*>(smem_raw + off) + w * (kRowRing + kDiagRing);
    uint64_t* qbars = rbars + kRowRing;
    off += (size_t)W * (kRowRing + kDiagRing) * 8;
    off = (off + 15) & ~(size_t)15;
    unsigned long long* prog = reinterpret_cast<unsigned long long*>(smem_raw + off);
    unsigned long long* fin = prog + NB;      // per-warp finished-strip counters (run-ahead gate)
    off += (size_t)(NB + W) * 8;
    off = (off + 15) & ~(size_t)15;
    float* bnd = reinterpret_cast<float*>(smem_raw + off);

    if (t == 0) {
        for (int s = 0; s < kRowRing; ++s) mbar_init(&rbars[s], 32);
        for (int s = 0; s < kDiagRing; ++s) mbar_init(&qbars[s], kTMA ? 1 : 32);
    }
    if ((int)threadIdx.x < NB) prog[threadIdx.x] = ~0ull;
    if ((int)threadIdx.x < W) fin[threadIdx.x] = 0ull;
    fence_mbar_init();
    __syncthreads();

    const bool varlen = (p.d.xlen != nullptr) || (p.d.ylen != nullptr);
    const int u = 31 - t;
    const float* Ebase = p.E + (M + 2) + 1;     // lattice cell (0,0)
    const long long Epair = (long long)(N + 2) * (M + 2);

    Strip cur, nxt;
    strip_first(cur, p.d, w, W);
    nxt = cur;
    if (cur.valid) strip_next(nxt, p.d, w, W);

    TilePipe<kRowRing, kRowRing - 2> rpipe;
    TilePipe<kDiagRing, kDiagRing - 1> qpipe;
    rpipe.reset();
    qpipe.reset();

    // E tiles are anchored at the RIGHT edge of the pair's lattice: tile o covers
    // columns m-32(o+1) .. m-32o-1 (zero-filled below 0), so tile boundaries fall on
    // multiples of 32 steps of the right-to-left sweep.
    auto issue_row = [&](const Strip& st, int o, unsigned slot) {
        const int kb = st.K - 1 - st.k;
        const float* pb = Ebase + (long long)st.pair * Epair;
        const int col = st.m - kTile * (o + 1) + t;
        const bool cok = col >= 0;
        float* dst = etiles + slot * kTileElems;
#pragma unroll 8
        for (int r = 0; r < kTile; ++r) {
            const int row = kb * kTile + r;
            const bool ok = cok && row < st.n;
            const float* g = ok ? (pb + (long long)row * (M + 2) + col) : p.E;
            cp_async4_zfill(dst + r * kTile + t, g, ok);
        }
        cp_async_mbar_arrive_noinc(&rbars[slot]);
    };
    auto issue_q = [&](const Strip& st, int a, unsigned slot) {
        const int kb = st.K - 1 - st.k;
        const long long so = (long long)st.pair * p.ql.pair_stride + (long long)kb * p.ql.strip_stride;
        float* dq = qring + (2 * slot) * kDiagElems;
        q_tile_load<kTMA>(dq, &qbars[slot], p.Q + so, st.m + 15 - kDiagRows * a, t, 2);
        q_tile_load<kTMA>(dq + kDiagElems, &qbars[slot], p.Qd + so, st.m + 15 - kDiagRows * a, t, 0);
        if (!kTMA) cp_async_mbar_arrive_noinc(&qbars[slot]);
    };

    while (cur.valid) {
        strip_gate(fin, cur.q, w, W);
        const int n = cur.n, m = cur.m;
        const int kb = cur.K - 1 - cur.k;
        const int T = (m + kTile - 1) / kTile;
        const int Tn = nxt.valid ? (nxt.m + kTile - 1) / kTile : 0;
        const int Ta = (m + 31 + kDiagRows - 1) / kDiagRows;
        const int Tan = nxt.valid ? (nxt.m + 31 + kDiagRows - 1) / kDiagRows : 0;
        const int i = kb * kTile + t + 1;
        const bool row_ok = i <= n;
        const bool has_below = cur.k > 0;
        const bool feeds_up = kb > 0;
        const unsigned q = cur.q;
        const float* bnd_r = bnd + (size_t)((q + NB - 1) % NB) * Mcap;
        float* bnd_w = bnd + (size_t)(q % NB) * Mcap;
        const unsigned long long* prog_r = prog + ((q + NB - 1) % NB);
        unsigned long long* prog_w = prog + (q % NB);
        float* Edb = p.Ed + (long long)cur.pair * Epair;

        int avail = 0;
        unsigned lslot = rpipe.wslot;
        unsigned dslot = 0;
        float zout = 0.f, dprev = 0.f, yprev = 0.f;

        for (int s = 0; s <= m + 30; ++s) {
            if ((s & 31) == 0 && (s >> 5) < T) {
                __syncwarp();
                rpipe.pump(s >> 5, T, nxt.valid, Tn,
                           [&](bool fn, int ti, unsigned slot) { issue_row(fn ? nxt : cur, ti, slot); });
                rpipe.wait(rbars);
            }
            if ((s & (kDiagRows - 1)) == 0) {
                __syncwarp();
                qpipe.pump(s / kDiagRows, Ta, nxt.valid, Tan,
                           [&](bool fn, int ti, unsigned slot) { issue_q(fn ? nxt : cur, ti, slot); });
                dslot = qpipe.wait(qbars);
            }
            const int cr = s - u;                   // reverse column index of this lane
            const int c = m - 1 - cr;
            if (has_below && s < m && avail < s + 1) avail = progress_wait(prog_r, q - 1, s + 1);

            float zin = __shfl_down_sync(kFull, zout, 1);
            if (t == 31) {
                zin = 0.f;
                if (has_below && c >= 0 && c < m) zin = bnd_r[c];
            }
            const bool in = row_ok && c >= 0 && c < m;
            float ed = 0.f, X = 0.f, D = 0.f, Y = 0.f;
            if (in) {
                const float e = etiles[lslot * kTileElems + t * kTile + (31 - (cr & 31))];
                const float* qt =
                    qring + (2 * dslot) * kDiagElems + (kDiagRows - 1 - (s & (kDiagRows - 1))) * kStepFloats + t;
                const float* qdt = qt + kDiagElems;
                ed = zin + yprev;
                // nw.py:260-265, push form
                // implied states: q_m = (1 - q_x) - q_y, qd_m = -(qd_x + qd_y);
                // a marked cell (Q == 0, sw.py first row / column) pushes nothing
                const float qx = qt[0], qy = qt[kQY];
                if (qx >= 0.f) {
                    const float qdx = qdt[0], qdy = qdt[kQY];
                    X = fmaf(qdx, e, qx * ed);
                    Y = fmaf(qdy, e, qy * ed);
                    D = fmaf(-(qdx + qdy), e, ((1.f - qx) - qy) * ed);
                }
            }
            if (c >= 0 && c < m) otile[((c >> 5) & 1) * kTileElems + t * kTile + (c & 31)] = ed;
            zout = X + dprev;
            dprev = D;
            yprev = Y;
            if (t == 0 && feeds_up && c >= 0 && c < m) {
                bnd_w[c] = zout;
                const int done = m - c;
                if ((done & 7) == 0 || c == 0)
                    st_release_u64(prog_w, ((unsigned long long)q << 32) | (unsigned)done);
            }
            if (cr >= 0 && (cr & 31) == 31) lslot = (lslot + 1 == kRowRing) ? 0u : lslot + 1;
            const int c0 = m + 30 - s;
            if (c0 >= 0 && (c0 & 31) == 0) {
                __syncwarp();
                const int tc = c0 >> 5;
                const float* src = otile + (tc & 1) * kTileElems;
                const int col = tc * kTile + t;
                if (col < m) {
                    float* dstp = Edb + (long long)(kb * kTile + 1) * (M + 2) + col + 1;
                    const int rmax = min(kTile, n - kb * kTile);
                    for (int r = 0; r < rmax; ++r) dstp[(long long)r * (M + 2)] = src[r * kTile + t];
                }
                __syncwarp();
            }
        }
        if (!varlen) {
            if (row_ok) {
                Edb[(long long)i * (M + 2)] = 0.f;
                Edb[(long long)i * (M + 2) + M + 1] = 0.f;
            }
            if (kb == 0)
                for (int col = t; col < M + 2; col += 32) Edb[col] = 0.f;
            if (cur.k == 0)
                for (int col = t; col < M + 2; col += 32) Edb[(long long)(N + 1) * (M + 2) + col] = 0.f;
        }
        rpipe.next_strip(T);
        qpipe.next_strip(Ta);
        strip_done(fin, cur.q, w, W);
        cur = nxt;
        if (cur.valid) strip_next(nxt, p.d, w, W);
    }
}

}  // namespace b200dp
