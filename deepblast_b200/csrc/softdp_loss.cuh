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

// Grid (row slabs, pairs): a CTA takes kLossRows rows of one pair, warp w rows w, w + 8, ... of
// the slab; lanes stride the columns, so every access is a coalesced 128-byte piece and no
// index needs a division.  Slabs add their partial (sum, count) to the pair's totals with
// atomics (pair_sum / pair_count zero-filled by the caller), a one-CTA kernel then forms
// l_b / B.  (One CTA per pair, as in round 1, left a B = 32 batch on 32 of 148 SMs.)
constexpr int kLossRows = 32;

__global__ void __launch_bounds__(256) softdp_mxent_fwd_kernel(LossParams p) {
    __shared__ float red[8];
    const int b = blockIdx.y;
    const int n = p.xlen ? min(max(p.xlen[b], 0), p.N) : p.N;
    const int m = p.ylen ? min(max(p.ylen[b], 0), p.M) : p.M;
    const int i0 = blockIdx.x * kLossRows, i1 = min(i0 + kLossRows, n);
    if (i0 >= n) return;
    const float* yt = p.Ytrue + (long long)b * p.N * p.M;
    const float* gm = p.G ? p.G + (long long)b * p.N * p.M : nullptr;
    const float* yp = p.Ypred + (long long)b * p.pb;
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    float s = 0.f, c = 0.f;
    for (int i = i0 + w; i < i1; i += nw) {
        const float* ytr = yt + (long long)i * p.M;
        const float* gmr = gm ? gm + (long long)i * p.M : nullptr;
        const float* ypr = yp + (long long)i * p.pi;
#pragma unroll 4
        for (int j = lane; j < m; j += 32) {
            if (!gmr || gmr[j] != 0.f) {
                const float y = ytr[j];
                const float q = fminf(fmaxf(ypr[j], kLossEps), kLossMax);
                s += y * logf(q) + (1.f - y) * logf(1.f - q);
                c += 1.f;
            }
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

__global__ void __launch_bounds__(256) softdp_mxent_bwd_kernel(LossParams p) {
    const int b = blockIdx.y;
    const int n = p.xlen ? min(max(p.xlen[b], 0), p.N) : p.N;
    const int m = p.ylen ? min(max(p.ylen[b], 0), p.M) : p.M;
    const float* yt = p.Ytrue + (long long)b * p.N * p.M;
    const float* gm = p.G ? p.G + (long long)b * p.N * p.M : nullptr;
    const float* yp = p.Ypred + (long long)b * p.pb;
    float* gr = p.grad + (long long)b * p.N * p.M;
    const float scale = -p.gout[0] / (p.pair_count[b] * (float)p.B);
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    const int i0 = blockIdx.x * kLossRows, i1 = min(i0 + kLossRows, p.N);
    for (int i = i0 + w; i < i1; i += nw) {
        const float* ytr = yt + (long long)i * p.M;
        const float* gmr = gm ? gm + (long long)i * p.M : nullptr;
        const float* ypr = yp + (long long)i * p.pi;
        float* grr = gr + (long long)i * p.M;
#pragma unroll 4
        for (int j = lane; j < p.M; j += 32) {
            float g = 0.f;
            if (i < n && j < m && (!gmr || gmr[j] != 0.f)) {
                const float q0 = ypr[j];
                // clamp passes the gradient only inside [eps, max] (torch.clamp backward)
                if (q0 >= kLossEps && q0 <= kLossMax) {
                    const float y = ytr[j];
                    g = scale * (y / q0 - (1.f - y) / (1.f - q0));
                }
            }
            grr[j] = g;
        }
    }
}

}  // namespace b200dp
