// softdp_traceback.cuh -- batched greedy traceback over an expected-alignment
// matrix (reference: deepblast/nw.py:401-444 "cpu" rule, deepblast/nw_cuda.py:273-317
// "cuda" rule).  One thread per pair walks at most n+m steps; the reference does the
// same walk in Python with three device->host syncs per step.
#pragma once
#include "softdp_common.cuh"

namespace b200dp {

struct TracebackParams {
    const float* grad;          // [B, N, M] with element strides (sb, si, sj)
    long long sb, si, sj;
    const int* xlen;
    const int* ylen;
    int B, N, M;
    int variant;                // 0: nw.py rule, 1: nw_cuda.py rule
    int32_t* out;               // [B, cap, 3] triples (i, j, state), final (reversed) order
    int cap;
    int32_t* len;               // [B]; -1 capacity exceeded, -2 reference would raise IndexError
};

__device__ __forceinline__ int tb_wrap(int idx, int n) { return idx < 0 ? idx + n : idx; }

constexpr int kTbWarps = 4;      // pairs (warps) per CTA

// One WARP per pair.  The walk is a chain of dependent reads (the next cell depends on the comparison of
// the three candidates): one thread per pair reading global memory pays an L2 / DRAM round trip per step
// (measured on B200: 0.71 ms for 64 pairs of 256 x 256 -- six times the forward and backward sweeps
// together; 0.058 ms with this kernel).  Here the warp loads the 32 x 32 block of the matrix that ends at the current
// cell into shared memory (coalesced rows), every lane then walks the block redundantly from shared memory
// (same addresses: broadcasts, uniform control flow) until the walk leaves it -- at least 32 steps later --
// and lane 0 records the path.  Same values, same comparisons, same first-index tie-break.  The block
// walk only runs while i >= 1 and j >= 1; the plain loop below (Python index semantics, wrap-around,
// the sentinels of both rules) takes over on row 0 / column 0.
__global__ void __launch_bounds__(32 * kTbWarps) softdp_traceback_kernel(TracebackParams p) {
    __shared__ float tile[kTbWarps][32][33];
    const int w = threadIdx.x >> 5, t = threadIdx.x & 31;
    const int b = blockIdx.x * kTbWarps + w;
    if (b >= p.B) return;
    // clamped like the reference's slice aln[b, :n, :m] (alignment.py:166-170): never past the tensor
    const int n = p.xlen ? min(max(p.xlen[b], 0), p.N) : p.N;
    const int m = p.ylen ? min(max(p.ylen[b], 0), p.M) : p.M;
    const float* g = p.grad + (long long)b * p.sb;
    int32_t* out = p.out + (long long)b * p.cap * 3;
    // sentinels of nw.py:418 / nw_cuda.py:291, compared after rounding to fp32
    const float sentinel = p.variant == 0 ? -100000.0f : -1e10f;
    int i = n - 1, j = m - 1, len = 0, status = 0;
    bool stopped = false;
    if (p.cap < 1 || n < 1 || m < 1) {
        if (t == 0) p.len[b] = -1;
        return;
    }
    if (t == 0) {
        out[0] = i;
        out[1] = j;
        out[2] = 1;
    }
    len = 1;
    float (*tl)[33] = tile[w];
    while (i >= 1 && j >= 1 && status == 0 && !stopped) {
        const int r0 = max(i - 31, 0), c0 = max(j - 31, 0);
        __syncwarp();
        for (int r = 0; r <= i - r0; ++r)
            tl[r][t] = (c0 + t <= j) ? g[(long long)(r0 + r) * p.si + (long long)(c0 + t) * p.sj] : 0.f;
        __syncwarp();
        while (i - 1 >= r0 && j - 1 >= c0) {
            const float left = tl[i - 1 - r0][j - c0], diag = tl[i - 1 - r0][j - 1 - c0], upper = tl[i - r0][j - 1 - c0];
            // (a matrix that holds the sentinel value itself stops the reference's walk: kept)
            if (p.variant == 0 ? (diag == sentinel && upper == sentinel && left == sentinel)
                               : (diag == sentinel || upper == sentinel || left == sentinel)) {
                stopped = true;
                break;
            }
            int ij = 0;                  // torch.argmax: first maximal index
            float best = left;
            if (diag > best) { best = diag; ij = 1; }
            if (upper > best) { best = upper; ij = 2; }
            i -= (ij != 2);
            j -= (ij != 0);
            if (len >= p.cap) { status = -1; break; }
            if (t == 0) { out[3 * len] = i; out[3 * len + 1] = j; out[3 * len + 2] = ij; }
            len++;
        }
    }
    for (; status == 0 && !stopped;) {
        float left, diag, upper;
        // every read uses Python index semantics: -n <= idx < n, negatives wrap, anything
        // else is an IndexError (j can go negative once a diagonal move wrapped column 0)
        if (i <= 0) left = sentinel;
        else {
            if (j < -m || j >= m) { status = -2; break; }
            left = g[(long long)(i - 1) * p.si + (long long)tb_wrap(j, m) * p.sj];
        }
        if (i <= 0 && j <= 0) diag = sentinel;
        else {
            if (i - 1 < -n || j - 1 < -m) { status = -2; break; }
            diag = g[(long long)tb_wrap(i - 1, n) * p.si + (long long)tb_wrap(j - 1, m) * p.sj];
        }
        if (j <= 0) upper = sentinel;
        else {
            if (i < -n) { status = -2; break; }
            upper = g[(long long)tb_wrap(i, n) * p.si + (long long)(j - 1) * p.sj];
        }
        const bool stop = p.variant == 0
                              ? (diag == sentinel && upper == sentinel && left == sentinel)
                              : (diag == sentinel || upper == sentinel || left == sentinel);
        if (stop) break;
        int ij = 0;                      // torch.argmax: first maximal index
        float best = left;
        if (diag > best) { best = diag; ij = 1; }
        if (upper > best) { best = upper; ij = 2; }
        if (ij == 0) i -= 1;
        else if (ij == 1) { i -= 1; j -= 1; }
        else j -= 1;
        if (len >= p.cap) { status = -1; break; }
        if (t == 0) { out[3 * len] = i; out[3 * len + 1] = j; out[3 * len + 2] = ij; }
        len++;
    }
    while (status == 0 && i > 0) {       // "take care of any outstanding gaps"
        i--;
        if (len >= p.cap) { status = -1; break; }
        if (t == 0) { out[3 * len] = i; out[3 * len + 1] = j; out[3 * len + 2] = 0; }
        len++;
    }
    while (status == 0 && j > 0) {
        j--;
        if (len >= p.cap) { status = -1; break; }
        if (t == 0) { out[3 * len] = i; out[3 * len + 1] = j; out[3 * len + 2] = 2; }
        len++;
    }
    if (status != 0) {
        if (t == 0) p.len[b] = status;
        return;
    }
    __syncwarp();                                        // lane 0's triples are visible to the warp
    for (int a = t; 2 * a + 1 < len; a += 32) {          // states[::-1], the lanes swap disjoint pairs
        const int z = len - 1 - a;
        for (int c = 0; c < 3; ++c) {
            const int32_t tmp = out[3 * a + c];
            out[3 * a + c] = out[3 * z + c];
            out[3 * z + c] = tmp;
        }
    }
    if (t == 0) p.len[b] = len;
}

}  // namespace b200dp
