// softdp_loss.cuh -- MatrixCrossEntropy over the expected alignment, fused (reference:
// deepblast/losses.py:9-48, called from trainer.py:154-171 on predA = decode(theta, A)).
//
// The reference loops over the batch in Python with two masked_selects per pair:
//   Yp    = clamp(Ypred, 3e-8, 1 - 3e-8)                    (upper bound: 1 - 2^-24 in fp32)
//   l_b   = -mean over {(i,j) : i < xlen_b, j < ylen_b, G[b,i,j] != 0} of
//            Ytrue log Yp + (1 - Ytrue) log(1 - Yp)
//   loss  = mean_b l_b
// Here one launch reduces every pair's sum and count in one pass over the three tensors (the
// step right after the DP backward: Ypred is read in place, any row stride), and the backward
// writes dloss/dYpred densely (zeros outside the mask) in one more pass.
#pragma once
#include "softdp_common.cuh"

namespace b200dp {

struct LossParams {
    const float* Ytrue;      // [B, N, M] contiguous
    const float* Ypred;      // [B, N, M], element strides (pb, pi, 1)
    const float* G;          // [B, N, M] contiguous, non-zero = counted; nullptr = all ones
    const int* xlen;         // [B] or nullptr (= N)
    const int* ylen;         // [B] or nullptr (= M)
    long long pb, pi;
    int B, N, M;
    float* pair_loss;        // [B]  l_b / B
    float* pair_count;       // [B]  number of counted cells
    // backward
    const float* gout;       // scalar upstream gradient
    float* grad;             // [B, N, M] contiguous
};

constexpr float kLossEps = 3e-8f;
constexpr float kLossMax = 0.99999994f;   // fp32(1 - 3e-8) = 1 - 2^-24, the largest float below 1

__device__ __forceinline__ float block_sum(float v, float* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    const int w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[w] = v;
    __syncthreads();
    float s = 0.f;
    for (int i = 0; i < nw; ++i) s += red[i];
    return s;
}

// Grid (row slabs, pairs): a CTA takes kLossRows rows of one pair, warp w the four rows
// 4w .. 4w+3 of the slab.  All loads of a step -- G, Ytrue and Ypred of two rows, 16 bytes per
// lane where the row pitches allow (M and the Ypred pitch multiples of 4, 16-byte aligned bases)
// -- are issued before anything is used, unconditionally (the mask is applied to the values,
// not to the loads), so a lane has 6 independent 16-byte loads in flight at four CTAs per SM
// (measured on B200: better than 12 loads per lane at two or three CTAs per SM): the kernels are
// pure streaming passes (12 B/cell forward, 16 B/cell backward) and need the memory-level
// parallelism.  Slabs add their partial (sum, count) to the pair's totals with atomics
// (zero-filled by b200dp_mxent_fwd), a small kernel then forms l_b / B.
constexpr int kLossRows = 32;

// Ytrue is an alignment indicator (0 or 1) in every caller: one of the two terms then has the factor 0 and
// its logarithm / quotient need not be formed (both are finite: q is clamped into [eps, 1 - 2^-24]), which
// takes the kernels off the issue limit -- logf and the IEEE division are ~20 and ~10 instructions.  The
// result is the same bit for bit (x * 1 + finite * 0 = x); fractional labels take the general expression.
// (BIN is decided once per 16 cells for the whole warp -- a vote -- so the vector loops carry no per-cell branch)
template <bool BIN>
__device__ __forceinline__ float mxent_term(float y, float q0, float g, float& c) {
    const bool on = g != 0.f;
    const float q = fminf(fmaxf(q0, kLossEps), kLossMax);
    c += on ? 1.f : 0.f;
    if (BIN) {
        const float l = logf(y == 1.f ? q : 1.f - q);
        return on ? l : 0.f;
    }
    return on ? y * logf(q) + (1.f - y) * logf(1.f - q) : 0.f;
}
__device__ __forceinline__ float mxent_term(float y, float q0, float g, float& c) {
    return (y == 1.f || y == 0.f) ? mxent_term<true>(y, q0, g, c) : mxent_term<false>(y, q0, g, c);
}
template <bool BIN>
__device__ __forceinline__ float mxent_grad(float y, float q0, float g, float scale) {
    // clamp passes the gradient only inside [eps, max] (torch.clamp backward)
    const bool on = g != 0.f && q0 >= kLossEps && q0 <= kLossMax;
    if (BIN) {
        const float d = (y == 1.f) ? 1.f / q0 : -(1.f / (1.f - q0));
        return on ? scale * d : 0.f;
    }
    return on ? scale * (y / q0 - (1.f - y) / (1.f - q0)) : 0.f;
}
__device__ __forceinline__ float mxent_grad(float y, float q0, float g, float scale) {
    return (y == 1.f || y == 0.f) ? mxent_grad<true>(y, q0, g, scale) : mxent_grad<false>(y, q0, g, scale);
}
__device__ __forceinline__ bool is_binary(const float4& y) {
    return (y.x == 1.f || y.x == 0.f) && (y.y == 1.f || y.y == 0.f) && (y.z == 1.f || y.z == 0.f) && (y.w == 1.f || y.w == 0.f);
}

#ifndef B200DP_LOSS_FMINB
#define B200DP_LOSS_FMINB 4
#endif
#ifndef B200DP_LOSS_FROWS
#define B200DP_LOSS_FROWS 2
#endif
constexpr int kLossFwdRows = B200DP_LOSS_FROWS;
template <bool VEC>
__global__ void __launch_bounds__(256, B200DP_LOSS_FMINB) softdp_mxent_fwd_kernel(LossParams p) {
    __shared__ float red[8];
    const int b = blockIdx.y;
    const int n = p.xlen ? min(max(p.xlen[b], 0), p.N) : p.N;
    const int m = p.ylen ? min(max(p.ylen[b], 0), p.M) : p.M;
    const int i0 = blockIdx.x * kLossRows;
    if (i0 >= n) return;
    const float* yt = p.Ytrue + (long long)b * p.N * p.M;
    const float* gm = p.G ? p.G + (long long)b * p.N * p.M : nullptr;
    const float* yp = p.Ypred + (long long)b * p.pb;
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int r0 = i0 + 4 * w;
    float s = 0.f, c = 0.f;
    if (VEC) {
        const int m4 = m >> 2;                           // whole float4 groups inside the pair's columns
        for (int c4 = lane; c4 < m4; c4 += 32) {
            // kLossFwdRows rows at a time (the order of the additions does not depend on it)
#pragma unroll
            for (int rr = 0; rr < 4; rr += kLossFwdRows) {
                float4 y4[kLossFwdRows], q4[kLossFwdRows], g4[kLossFwdRows];
#pragma unroll
                for (int r = 0; r < kLossFwdRows; ++r) {
                    const int i = min(r0 + rr + r, n - 1);       // rows past the pair: reload the last row, weight 0
                    y4[r] = reinterpret_cast<const float4*>(yt + (long long)i * p.M)[c4];
                    q4[r] = reinterpret_cast<const float4*>(yp + (long long)i * p.pi)[c4];
                    g4[r] = gm ? reinterpret_cast<const float4*>(gm + (long long)i * p.M)[c4] : make_float4(1.f, 1.f, 1.f, 1.f);
                }
                auto add = [&](auto bin_tag) {
                    constexpr bool BIN = decltype(bin_tag)::value;
#pragma unroll
                    for (int r = 0; r < kLossFwdRows; ++r) {
                        const float k = (r0 + rr + r < n) ? 1.f : 0.f;
                        s += mxent_term<BIN>(y4[r].x, q4[r].x, g4[r].x * k, c) + mxent_term<BIN>(y4[r].y, q4[r].y, g4[r].y * k, c) +
                             mxent_term<BIN>(y4[r].z, q4[r].z, g4[r].z * k, c) + mxent_term<BIN>(y4[r].w, q4[r].w, g4[r].w * k, c);
                    }
                };
                bool bin = true;
#pragma unroll
                for (int r = 0; r < kLossFwdRows; ++r) bin = bin && is_binary(y4[r]);
                if (__all_sync(__activemask(), bin)) add(std::true_type{});
                else add(std::false_type{});
            }
        }
        // the m % 4 tail columns
        for (int j = 4 * m4 + lane; j < m; j += 32)
            for (int r = 0; r < 4; ++r)
                if (r0 + r < n) {
                    const long long i = r0 + r;
                    s += mxent_term(yt[i * p.M + j], yp[i * p.pi + j], gm ? gm[i * p.M + j] : 1.f, c);
                }
    } else {
        for (int r = 0; r < 4; ++r) {
            const long long i = r0 + r;
            if (i >= n) break;
            for (int j = lane; j < m; j += 32)
                s += mxent_term(yt[i * p.M + j], yp[i * p.pi + j], gm ? gm[i * p.M + j] : 1.f, c);
        }
    }
    s = block_sum(s, red);
    c = block_sum(c, red);
    if (threadIdx.x == 0) {
        atomicAdd(p.pair_loss + b, s);               // raw sum here; softdp_mxent_fin_kernel turns it into l_b / B
        atomicAdd(p.pair_count + b, c);
    }
}

__global__ void softdp_mxent_fin_kernel(LossParams p) {
    for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < p.B; b += gridDim.x * blockDim.x)
        p.pair_loss[b] = -(p.pair_loss[b] / p.pair_count[b]) / (float)p.B;   // mean of an empty selection is NaN, as in torch
}

// The backward is a pure streaming pass (16 B/cell) with an IEEE division per cell: at 96 registers (two CTAs
// per SM) it ran at 0.210 ms for C2; four CTAs per SM with two rows (six 16-byte loads) in flight per lane:
// 0.168 ms (measured on B200, scripts/gpu_loss_perf.py).
#ifndef B200DP_LOSS_MINB
#define B200DP_LOSS_MINB 4
#endif
#ifndef B200DP_LOSS_BROWS
#define B200DP_LOSS_BROWS 2
#endif
constexpr int kLossBwdRows = B200DP_LOSS_BROWS;     // rows of a warp's four loaded together (1, 2 or 4)
template <bool VEC>
__global__ void __launch_bounds__(256, B200DP_LOSS_MINB) softdp_mxent_bwd_kernel(LossParams p) {
    const int b = blockIdx.y;
    const int n = p.xlen ? min(max(p.xlen[b], 0), p.N) : p.N;
    const int m = p.ylen ? min(max(p.ylen[b], 0), p.M) : p.M;
    const float* yt = p.Ytrue + (long long)b * p.N * p.M;
    const float* gm = p.G ? p.G + (long long)b * p.N * p.M : nullptr;
    const float* yp = p.Ypred + (long long)b * p.pb;
    float* gr = p.grad + (long long)b * p.N * p.M;
    const float scale = -p.gout[0] / (p.pair_count[b] * (float)p.B);
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int r0 = blockIdx.x * kLossRows + 4 * w;
    if (VEC) {
        const int M4 = p.M >> 2;
        // kLossBwdRows rows at a time: 3 * kLossBwdRows independent 16-byte loads in flight per lane
        for (int rr = 0; rr < 4; rr += kLossBwdRows) {
            for (int c4 = lane; c4 < M4; c4 += 32) {
                float4 y4[kLossBwdRows], q4[kLossBwdRows], g4[kLossBwdRows];
#pragma unroll
                for (int r = 0; r < kLossBwdRows; ++r) {
                    const int i = min(r0 + rr + r, p.N - 1);
                    y4[r] = reinterpret_cast<const float4*>(yt + (long long)i * p.M)[c4];
                    q4[r] = reinterpret_cast<const float4*>(yp + (long long)i * p.pi)[c4];
                    g4[r] = gm ? reinterpret_cast<const float4*>(gm + (long long)i * p.M)[c4] : make_float4(1.f, 1.f, 1.f, 1.f);
                }
                auto put = [&](auto bin_tag) {
                    constexpr bool BIN = decltype(bin_tag)::value;
#pragma unroll
                    for (int r = 0; r < kLossBwdRows; ++r) {
                        const int i = r0 + rr + r;
                        if (i >= p.N) break;
                        const int j = 4 * c4;
                        const float ki = i < n ? 1.f : 0.f;                      // zeros outside the pair's lattice
                        float4 o;
                        o.x = mxent_grad<BIN>(y4[r].x, q4[r].x, g4[r].x * (j < m ? ki : 0.f), scale);
                        o.y = mxent_grad<BIN>(y4[r].y, q4[r].y, g4[r].y * (j + 1 < m ? ki : 0.f), scale);
                        o.z = mxent_grad<BIN>(y4[r].z, q4[r].z, g4[r].z * (j + 2 < m ? ki : 0.f), scale);
                        o.w = mxent_grad<BIN>(y4[r].w, q4[r].w, g4[r].w * (j + 3 < m ? ki : 0.f), scale);
                        reinterpret_cast<float4*>(gr + (long long)i * p.M)[c4] = o;
                    }
                };
                bool bin = true;
#pragma unroll
                for (int r = 0; r < kLossBwdRows; ++r) bin = bin && is_binary(y4[r]);
                if (__all_sync(__activemask(), bin)) put(std::true_type{});
                else put(std::false_type{});
            }
        }
    } else {
        for (int r = 0; r < 4; ++r) {
            const long long i = r0 + r;
            if (i >= p.N) break;
            for (int j = lane; j < p.M; j += 32) {
                const float g = (i < n && j < m) ? (gm ? gm[i * p.M + j] : 1.f) : 0.f;
                gr[i * p.M + j] = mxent_grad(yt[i * p.M + j], yp[i * p.pi + j], g, scale);
            }
        }
    }
}

}  // namespace b200dp
