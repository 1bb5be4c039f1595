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
constexpr int kTbW = 4;      // register window of the interior walk

__global__ void softdp_traceback_kernel(TracebackParams p) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
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
        p.len[b] = -1;
        return;
    }
    out[0] = i;
    out[1] = j;
    out[2] = 1;
    len = 1;
    // Interior of the lattice (no border, no wrap-around, no sentinel in reach): the walk is a chain of
    // dependent loads, about one L2 round trip per step.  A kTbW x kTbW register window w[a][c] =
    // grad[i-a, j-c] is kept around the current cell; after a move the window shifts and only its far
    // row / column is loaded -- cells the walk cannot need before kTbW - 2 more steps -- so the round
    // trips of consecutive steps overlap.  Same values, same comparisons: the decisions are those of the
    // plain loop below, which takes over as soon as the window would touch row 0 or column 0.
    if (i >= kTbW && j >= kTbW) {
        float w[kTbW][kTbW];
#pragma unroll
        for (int a = 0; a < kTbW; ++a)
#pragma unroll
            for (int c = 0; c < kTbW; ++c) w[a][c] = g[(long long)(i - a) * p.si + (long long)(j - c) * p.sj];
        while (i >= kTbW && j >= kTbW) {
            const float left = w[1][0], diag = w[1][1], upper = w[0][1];
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
            const int di = ij != 2, dj = ij != 0;
            i -= di;
            j -= dj;
            if (len >= p.cap) { status = -1; break; }
            out[3 * len] = i; out[3 * len + 1] = j; out[3 * len + 2] = ij; len++;
            // shift: w'[a][c] = w[a + di][c + dj]; the far row (di) / far column (dj) is loaded
            float nr[kTbW], nc[kTbW];
#pragma unroll
            for (int c = 0; c < kTbW; ++c) nr[c] = 0.f, nc[c] = 0.f;
            if (i >= kTbW - 1 && j >= kTbW - 1) {        // (the window of the new cell is inside the lattice)
                if (di) {
#pragma unroll
                    for (int c = 0; c < kTbW; ++c) nr[c] = g[(long long)(i - (kTbW - 1)) * p.si + (long long)(j - c) * p.sj];
                }
                if (dj) {
#pragma unroll
                    for (int a = 0; a < kTbW; ++a) nc[a] = g[(long long)(i - a) * p.si + (long long)(j - (kTbW - 1)) * p.sj];
                }
            }
            if (di) {
#pragma unroll
                for (int a = 0; a < kTbW - 1; ++a)
#pragma unroll
                    for (int c = 0; c < kTbW; ++c) w[a][c] = w[a + 1][c];
            }
            if (dj) {
#pragma unroll
                for (int a = 0; a < kTbW; ++a)
#pragma unroll
                    for (int c = 0; c < kTbW - 1; ++c) w[a][c] = w[a][c + 1];
            }
            if (di) {
#pragma unroll
                for (int c = 0; c < kTbW; ++c) w[kTbW - 1][c] = nr[c];
            }
            if (dj) {
#pragma unroll
                for (int a = 0; a < kTbW; ++a) w[a][kTbW - 1] = nc[a];
            }
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
        out[3 * len] = i; out[3 * len + 1] = j; out[3 * len + 2] = ij; len++;
    }
    while (status == 0 && i > 0) {       // "take care of any outstanding gaps"
        i--;
        if (len >= p.cap) { status = -1; break; }
        out[3 * len] = i; out[3 * len + 1] = j; out[3 * len + 2] = 0; len++;
    }
    while (status == 0 && j > 0) {
        j--;
        if (len >= p.cap) { status = -1; break; }
        out[3 * len] = i; out[3 * len + 1] = j; out[3 * len + 2] = 2; len++;
    }
    if (status != 0) {
        p.len[b] = status;
        return;
    }
    for (int a = 0, z = len - 1; a < z; ++a, --z) {      // states[::-1]
        for (int c = 0; c < 3; ++c) {
            const int32_t tmp = out[3 * a + c];
            out[3 * a + c] = out[3 * z + c];
            out[3 * z + c] = tmp;
        }
    }
    p.len[b] = len;
}

}  // namespace b200dp
