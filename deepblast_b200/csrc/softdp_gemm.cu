// softdp_gemm.cu -- the step right BEFORE the DP: the match / gap score
// matrices of `NeuralAligner.forward` (deepblast/alignment.py:122-123, also :134-135, :162-163)
//     theta = softplus  (einsum('bid,bjd->bij', zx, zy))
//     A     = logsigmoid(einsum('bid,bjd->bij', gx, gy))
// as ONE batched GEMM launch on the 5th-generation tensor cores with the activation fused into the
// epilogue -- the one tensor-core-shaped piece of the path (D = 1024 by default, trainer.py:351-353).
//
//   * tcgen05.mma (cta_group::1, kind::f16, M = 128, N = 128, K = 16), issued by one elected thread;
//     accumulators live in TMEM (128 fp32 columns per CTA), read back with tcgen05.ld for the epilogue.
//   * operands: TMA (cp.async.bulk.tensor, 128-byte swizzle) into a 3-stage shared-memory ring,
//     full / empty mbarriers, tcgen05.commit frees a stage when its MMAs have read it.
//   * fp32 accuracy from bf16 tensor cores: every fp32 embedding is split into two bf16 numbers
//     x = hi + lo (hi = bf16(x), lo = bf16(x - hi): 16 bits of mantissa together) by a small
//     pre-pass, and the product is accumulated in fp32 as  hi.hi + lo.hi + hi.lo  -- three passes over K
//     in the same accumulator (the dropped lo.lo term is 2^-16 of a product).  Plain bf16 or TF32
//     misses the 1e-4 bar of the north star on sums of 1024 products.
//   * epilogue: TMEM -> registers (32x32b.x32) -> softplus / logsigmoid (torch's definitions,
//     threshold 20) -> transposed through shared memory -> coalesced 128-byte row stores into theta / A
//     in the DP's operand layout: dense [B, Lx, Ly] or the PACKED ragged layout of plan.py (per-pair
//     offset and pitch), so the producer writes exactly what the strip-queue forward reads.
//   * warp roles (192 threads): warp 0 TMA producer, warp 1 TMEM allocation + MMA issue, warps 2..5
//     epilogue (TMEM lane quarter = warp % 4).  96 KB of shared memory per CTA: two CTAs per SM, so one
//     CTA's epilogue overlaps the other's main loop.
//
// Two kernels: version 2 (default, below: the hi/lo split inside the GEMM, no workspace) and version 1
// (a pre-pass writes bf16 copies, operands arrive through TMA; selected by passing a workspace).  Measured on
// B200, B = 1024, L = 256, D = 1024: 1.58 ms against 2.73 ms (torch fp32 einsum + activations: 4.9 ms) --
// the pre-pass alone moves 8.6 GB.  At this shape the producer is bound by reading the 4.3 GB of fp32
// embeddings (each 128-row block is read by two tiles: 8.6 GB through L2), not by the tensor cores.
//
// Feeding the forward's operand ring directly (no theta / A round trip through HBM) was the survey's
// second stage; it is deliberately not built: at D = 1024 the producer costs 2 x 3 x 2 x 1024 = 12288
// tensor-core FLOP per cell against the DP's 16 bytes per cell, i.e. the GEMM side is compute-bound at
// about the DP's own time, and the 8 B/cell round trip it would save is 4 % of the producer's time.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstring>
#include <mutex>
#include <string>

#include "../../include/b200dp.h"
#include "softdp_common.cuh"
#include "softdp_host.h"

using namespace b200dp;
using namespace b200dp_host;

namespace b200dp {

constexpr int kGM = 128, kGN = 128, kGK = 64;            // CTA tile; K block = 64 bf16 = one 128-byte swizzle row
constexpr int kGStages = 3;
constexpr int kGStageBytes = (kGM + kGN) * kGK * 2;        // 32 KB: A tile + B tile
constexpr int kGThreads = 192;
constexpr int kGSmem = kGStages * kGStageBytes + 1024 /*alignment slack*/ + 256;

struct GemmParams {
    float* theta;              // outputs
    float* A;
    const int* xlen;           // per-pair lengths or null
    const int* ylen;
    const long long* pair_off; // packed layout: element offset of pair b (null = dense [B, Lx, Ly])
    int B, Lx, Ly, D;
};

// ---- tcgen05 / TMEM primitives (PTX ISA 8.6+, sm_100a) -------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// D[tmem] (+)= A[smem] * B[smem]^T, both operands K-major
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive on an mbarrier once every tcgen05.mma issued so far by this thread has completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
// 32 lanes x 32 consecutive fp32 columns: thread l of the warp gets TMEM lane (base lane + l), columns c .. c+31
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// Shared-memory matrix descriptor of a K-major bf16 tile in the 128-byte-swizzle layout TMA writes
// (rows of 128 bytes, atoms of 8 rows = 1024 bytes): start address, stride between 8-row groups
// (SBO = 1024 B), descriptor version 1 (Blackwell), layout type 2 = SWIZZLE_128B.  The leading-dimension
// offset is not used by swizzled K-major layouts.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);           // bits [0,14): start address >> 4
    d |= (uint64_t)1 << 16;                                // bits [16,30): leading byte offset >> 4 (unused: 1)
    d |= (uint64_t)(1024 >> 4) << 32;                      // bits [32,46): stride byte offset >> 4
    d |= (uint64_t)1 << 46;                                // bits [46,48): version = 1
    d |= (uint64_t)2 << 61;                                // bits [61,64): SWIZZLE_128B
    return d;
}
// Instruction descriptor, kind::f16: D = F32, A = B = BF16, both K-major, N >> 3 at bit 17, M >> 4 at bit 24.
constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kGN >> 3) << 17) | ((uint32_t)(kGM >> 4) << 24);

__device__ __forceinline__ float softplus_f(float s) {      // torch.nn.functional.softplus (beta 1, threshold 20)
    return s > 20.f ? s : log1pf(expf(s));
}
__device__ __forceinline__ float logsigmoid_f(float s) {    // torch: min(s, 0) - log1p(exp(-|s|))
    return fminf(s, 0.f) - log1pf(expf(-fabsf(s)));
}

// grid (tiles along y, tiles along x, 2 B): z = 2 b + which; which 0: theta from (zx, zy), 1: A from (gx, gy).
// Tensor maps: [B, L, D] bf16, box 64 x 128 x 1, 128-byte swizzle; index [which][operand x/y][hi/lo].
struct GemmMaps {
    CUtensorMap m[2][2][2];
};

__global__ void __launch_bounds__(kGThreads) softdp_theta_a_kernel(const __grid_constant__ GemmMaps maps, GemmParams p) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    // 128-byte-swizzle tiles need 1024-byte alignment
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + kGStages * kGStageBytes);
    uint64_t* empty = full + kGStages;
    uint64_t* tmem_full = empty + kGStages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.z >> 1, which = blockIdx.z & 1;
    const int n = p.xlen ? min(max(p.xlen[b], 0), p.Lx) : p.Lx;
    const int m = p.ylen ? min(max(p.ylen[b], 0), p.Ly) : p.Ly;
    const int i0 = blockIdx.y * kGM, j0 = blockIdx.x * kGN;
    if (i0 >= n || j0 >= m) return;                      // tile outside this pair's lattice (whole CTA leaves)

    if (threadIdx.x == 0) {
        for (int s = 0; s < kGStages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        mbar_init(tmem_full, 1);
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, kGN);           // 128 fp32 columns x 128 lanes
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_acc = *tmem_slot;

    const int KB = p.D / kGK;                            // k blocks per pass
    const int NK = 3 * KB;                               // hi.hi, lo.hi, hi.lo
    if (warp == 0) {
        // ===== TMA producer =====
        if (elect_one()) {
            for (int kb = 0; kb < NK; ++kb) {
                const int s = kb % kGStages, pass = kb / KB, kk = (kb - pass * KB) * kGK;
                mbar_wait(&empty[s], ((kb / kGStages) & 1) ^ 1);
                unsigned char* sa = smem + s * kGStageBytes;
                unsigned char* sb = sa + kGM * kGK * 2;
                mbar_expect_tx(&full[s], kGStageBytes);
                tma_load_3d(sa, &maps.m[which][0][pass == 1 ? 1 : 0], &full[s], kk, i0, b);
                tma_load_3d(sb, &maps.m[which][1][pass == 2 ? 1 : 0], &full[s], kk, j0, b);
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        for (int kb = 0; kb < NK; ++kb) {
            const int s = kb % kGStages;
            mbar_wait(&full[s], (kb / kGStages) & 1);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t sa = smem_u32(smem + s * kGStageBytes);
                const uint32_t sb = sa + kGM * kGK * 2;
                const uint64_t ad = umma_desc_sw128(sa), bd = umma_desc_sw128(sb);
#pragma unroll
                for (int k = 0; k < kGK / 16; ++k)       // UMMA_K = 16 bf16 = 32 bytes inside the swizzle row: +2 (>> 4)
                    umma_bf16(tmem_acc, ad + 2 * k, bd + 2 * k, kIdesc, (kb | k) != 0);
                umma_commit(&empty[s]);                  // the stage is free once these MMAs have read it
                if (kb == NK - 1) umma_commit(tmem_full);
            }
            __syncwarp();
        }
    } else {
        // ===== epilogue: TMEM -> registers -> activation -> smem transpose -> coalesced stores =====
        const int q = warp & 3;                          // TMEM lane quarter this warp may read
        mbar_wait(tmem_full, 0);
        tc_fence_after();
        // (every TMA load and every MMA has completed: the operand ring is free, reuse stage memory)
        float* tbuf = reinterpret_cast<float*>(smem) + (warp - 2) * (32 * 33);
        float* out = which ? p.A : p.theta;
        const int pitch = p.pair_off ? ((m + 3) & ~3) : p.Ly;
        float* ob = out + (p.pair_off ? p.pair_off[b] : (long long)b * p.Lx * p.Ly);
#pragma unroll 1
        for (int c0 = 0; c0 < kGN; c0 += 32) {
            if (j0 + c0 >= m) break;
            float v[32];
            tmem_ld32(tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
#pragma unroll
            for (int c = 0; c < 32; ++c) tbuf[lane * 33 + c] = which ? logsigmoid_f(v[c]) : softplus_f(v[c]);
            __syncwarp();
            const int col = j0 + c0 + lane;
#pragma unroll 8
            for (int r = 0; r < 32; ++r) {
                const int row = i0 + q * 32 + r;
                if (row < n && col < m) ob[(long long)row * pitch + col] = tbuf[r * 33 + lane];
            }
            __syncwarp();
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_acc, kGN);
    }
}

// ---- version 2: the split happens INSIDE the GEMM (no pre-pass, no bf16 copies in HBM) -------------
// The pre-pass of version 1 reads every fp32 embedding once and writes hi and lo (8 B per element):
// at L = 256 that is as much time as the tensor-core work itself.  Here eight converter warps load
// the fp32 operand tiles straight from global memory (coalesced 16-byte loads, the next k block in
// flight in registers while the current one is multiplied), form hi = bf16(x), lo = bf16(x - hi) and
// store the four bf16 tiles (A hi, A lo, B hi, B lo) into shared memory in exactly the 128-byte-swizzle
// K-major layout TMA would have produced (16-byte chunk c of row r lands at chunk c ^ (r & 7)); one
// fence.proxy.async per thread makes the generic-proxy stores visible to the tensor core's async proxy,
// then the warp arrives on the `full` barrier.  One shared-memory stage per CTA (64 KB) and two CTAs
// per SM: while one CTA's twelve MMAs of a k block run (hi.hi, lo.hi, hi.lo x 4 k steps), the other
// CTA converts, and tcgen05.commit on `empty` hands the stage back.
constexpr int kG2Threads = 288;                            // warp 0: MMA issue + TMEM, warps 1..8: converters, then epilogue
constexpr int kG2Tile = kGM * kGK * 2;                     // one bf16 operand tile: 16 KB
constexpr int kG2Smem = 4 * kG2Tile + 1024 + 256;

struct Gemm2Params {
    const float* x[2];         // [which]: zx / gx   [B, Lx, D]
    const float* y[2];         // [which]: zy / gy   [B, Ly, D]
    float* theta;
    float* A;
    const int* xlen;
    const int* ylen;
    const long long* pair_off;
    int B, Lx, Ly, D;
};

__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
    const __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&v);
}
// 8 fp32 -> 8 bf16 hi (one 16-byte chunk) and 8 bf16 lo.  Only PACKED conversions (F2FP.BF16.PACK_AB,
// full-rate ALU pipe): the single-element F2F conversions run on the quarter-rate XU pipe and made
// the converter warps, not the tensor core, the bottleneck.  hi as a float is the packed word's half
// shifted back into place (exact).
__device__ __forceinline__ void split2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
    hi = pack_bf16x2(x0, x1);                                   // low half = bf16(x0), high half = bf16(x1)
    const float h0 = __uint_as_float(hi << 16), h1 = __uint_as_float(hi & 0xffff0000u);
    lo = pack_bf16x2(x0 - h0, x1 - h1);
}
__device__ __forceinline__ void split8(const float4& u, const float4& v, uint4& hi, uint4& lo) {
    split2(u.x, u.y, hi.x, lo.x);
    split2(u.z, u.w, hi.y, lo.y);
    split2(v.x, v.y, hi.z, lo.z);
    split2(v.z, v.w, hi.w, lo.w);
}

__global__ void __launch_bounds__(kG2Threads, 2) softdp_theta_a2_kernel(Gemm2Params p) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    // tiles: [0] A hi, [1] A lo, [2] B hi, [3] B lo
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + 4 * kG2Tile);
    uint64_t* empty = full + 1;
    uint64_t* tmem_full = full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(full + 3);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.z >> 1, which = blockIdx.z & 1;
    const int n = p.xlen ? min(max(p.xlen[b], 0), p.Lx) : p.Lx;
    const int m = p.ylen ? min(max(p.ylen[b], 0), p.Ly) : p.Ly;
    const int i0 = blockIdx.y * kGM, j0 = blockIdx.x * kGN;
    if (i0 >= n || j0 >= m) return;

    if (threadIdx.x == 0) {
        mbar_init(full, 8);                              // one arrival per converter warp
        mbar_init(empty, 1);                             // tcgen05.commit
        mbar_init(tmem_full, 1);
        fence_mbar_init();
    }
    if (warp == 0) tmem_alloc(tmem_slot, kGN);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_acc = *tmem_slot;
    const int KB = p.D / kGK;

    if (warp == 0) {
        // ===== MMA issuer =====
        for (int kb = 0; kb < KB; ++kb) {
            mbar_wait(full, kb & 1);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t s0 = smem_u32(smem);
                const uint64_t ah = umma_desc_sw128(s0), al = umma_desc_sw128(s0 + kG2Tile);
                const uint64_t bh = umma_desc_sw128(s0 + 2 * kG2Tile), bl = umma_desc_sw128(s0 + 3 * kG2Tile);
#pragma unroll
                for (int k = 0; k < kGK / 16; ++k) {
                    umma_bf16(tmem_acc, ah + 2 * k, bh + 2 * k, kIdesc, (kb | k) != 0);      // hi . hi
                    umma_bf16(tmem_acc, al + 2 * k, bh + 2 * k, kIdesc, 1);                   // lo . hi
                    umma_bf16(tmem_acc, ah + 2 * k, bl + 2 * k, kIdesc, 1);                   // hi . lo
                }
                umma_commit(empty);                      // the stage may be overwritten once these MMAs have read it
                if (kb == KB - 1) umma_commit(tmem_full);
            }
            __syncwarp();
        }
    } else {
        // ===== converters: global fp32 -> registers -> (hi, lo) bf16 -> swizzled shared memory =====
        const int ct = threadIdx.x - 32;                 // 0..255
        const float* xb = p.x[which] + (long long)b * p.Lx * p.D;
        const float* yb = p.y[which] + (long long)b * p.Ly * p.D;
        // unit u = ct + 256 i (i = 0..3) of each operand: row u / 8, 32-byte piece u % 8 of the 256-byte fp32 row
        // (row = ct / 8 + 32 i, piece = ct % 8: the swizzled destination of unit i is unit 0's plus i * 4096)
        const int row0 = ct >> 3, c8 = ct & 7;
        const uint32_t dst0 = row0 * 128 + ((c8 ^ (row0 & 7)) << 4);
        int soff[8];                                     // element offsets inside the pair's [L, D] operand
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int row = row0 + 32 * (i & 3);
            const int grow = i >= 4 ? min(j0 + row, p.Ly - 1) : min(i0 + row, p.Lx - 1);   // overhanging rows: any valid row
            soff[i] = grow * p.D + c8 * 8;
        }
        float4 r[16];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float4* s4 = reinterpret_cast<const float4*>((i >= 4 ? yb : xb) + soff[i]);
            r[2 * i] = __ldg(s4);
            r[2 * i + 1] = __ldg(s4 + 1);
        }
        for (int kb = 0; kb < KB; ++kb) {
            mbar_wait(empty, (kb & 1) ^ 1);              // the MMAs of k block kb-1 have read the stage
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                uint4 hi, lo;
                split8(r[2 * i], r[2 * i + 1], hi, lo);
                // rows 32 apart share (row & 7): the same swizzle, 4096 bytes further
                unsigned char* d = smem + dst0 + (i & 3) * 4096 + (i >= 4 ? 2 * kG2Tile : 0);
                *reinterpret_cast<uint4*>(d) = hi;
                *reinterpret_cast<uint4*>(d + kG2Tile) = lo;
            }
            fence_proxy_async_smem();                    // generic-proxy stores -> visible to tcgen05 (async proxy)
            __syncwarp();
            if (lane == 0) mbar_arrive(full);
            if (kb + 1 < KB) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float4* s4 = reinterpret_cast<const float4*>((i >= 4 ? yb : xb) + soff[i] + (kb + 1) * kGK);
                    r[2 * i] = __ldg(s4);
                    r[2 * i + 1] = __ldg(s4 + 1);
                }
            }
        }
        {
            // ===== epilogue (all eight converter warps): TMEM -> registers -> activation -> smem transpose ->
            // coalesced stores.  A warp may read the TMEM lane quarter warp % 4: warps 1..4 take columns 0..63 of
            // their quarter, warps 5..8 columns 64..127 =====
            const int q = warp & 3, half = (warp - 1) >> 2;
            mbar_wait(tmem_full, 0);
            tc_fence_after();
            float* tbuf = reinterpret_cast<float*>(smem) + (warp - 1) * (32 * 33);       // the stage is free now
            float* out = which ? p.A : p.theta;
            const int pitch = p.pair_off ? ((m + 3) & ~3) : p.Ly;
            float* ob = out + (p.pair_off ? p.pair_off[b] : (long long)b * p.Lx * p.Ly);
#pragma unroll 1
            for (int c0 = half * (kGN / 2); c0 < (half + 1) * (kGN / 2); c0 += 32) {
                if (j0 + c0 >= m) break;
                float v[32];
                tmem_ld32(tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
#pragma unroll
                for (int c = 0; c < 32; ++c) tbuf[lane * 33 + c] = which ? logsigmoid_f(v[c]) : softplus_f(v[c]);
                __syncwarp();
                const int col = j0 + c0 + lane;
#pragma unroll 8
                for (int rr = 0; rr < 32; ++rr) {
                    const int row = i0 + q * 32 + rr;
                    if (row < n && col < m) ob[(long long)row * pitch + col] = tbuf[rr * 33 + lane];
                }
                __syncwarp();
            }
            tc_fence_before();
        }
    }
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem_acc, kGN);
    }
}

// fp32 -> (hi, lo) bf16 split of up to four tensors in one launch
struct SplitParams {
    const float* src[4];
    __nv_bfloat16* hi[4];
    __nv_bfloat16* lo[4];
    long long n[4];
};
__global__ void __launch_bounds__(256) softdp_split_kernel(SplitParams p) {
    const int w = blockIdx.y;
    const long long n4 = p.n[w] >> 2;
    const float4* s = reinterpret_cast<const float4*>(p.src[w]);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const float4 x = s[i];
        const float xs[4] = {x.x, x.y, x.z, x.w};
        __nv_bfloat16 h[4], l[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            h[k] = __float2bfloat16_rn(xs[k]);
            l[k] = __float2bfloat16_rn(xs[k] - __bfloat162float(h[k]));
        }
        reinterpret_cast<uint2*>(p.hi[w])[i] = *reinterpret_cast<uint2*>(h);
        reinterpret_cast<uint2*>(p.lo[w])[i] = *reinterpret_cast<uint2*>(l);
    }
}

}  // namespace b200dp

namespace {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn gemm_encode() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    return fn;
}

// [B, L, D] bf16, box 64 (K) x 128 (rows) x 1, 128-byte swizzle, rows past L read as zeros
bool encode_bf16_map(CUtensorMap* map, const void* ptr, int B, int L, int D) {
    EncodeTiledFn enc = gemm_encode();
    if (!enc) return false;
    cuuint64_t dims[3] = {(cuuint64_t)D, (cuuint64_t)L, (cuuint64_t)B};
    cuuint64_t strides[2] = {(cuuint64_t)D * 2, (cuuint64_t)L * D * 2};
    cuuint32_t box[3] = {(cuuint32_t)kGK, (cuuint32_t)kGM, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

}  // namespace

extern "C" {

size_t b200dp_theta_a_workspace(int B, int Lx, int Ly, int D) {
    if (B < 1 || Lx < 1 || Ly < 1 || D < 1) return 0;
    // hi and lo bf16 copies of zx, gx ([B, Lx, D]) and zy, gy ([B, Ly, D])
    return 4 * align256((size_t)B * Lx * D * 2) + 4 * align256((size_t)B * Ly * D * 2);
}

int b200dp_theta_a(const float* zx, const float* zy, const float* gx, const float* gy, int B, int Lx, int Ly, int D,
                   const int32_t* xlen, const int32_t* ylen, const long long* pair_off, float* theta, float* A,
                   void* workspace, size_t workspace_bytes, void* stream) {
    if (B < 0 || Lx < 1 || Ly < 1 || D < 1) return fail(-1, "b200dp_theta_a: need B >= 0, Lx, Ly, D >= 1");
    if (B == 0) return 0;
    if (D % kGK != 0) return fail(-5, "b200dp_theta_a: the embedding dimension must be a multiple of 64");
    if (2 * (long long)B > 65535) return fail(-5, "b200dp_theta_a: batch too large for one launch (2 B <= 65535)");
    if (!zx || !zy || !gx || !gy || !theta || !A) return fail(-1, "b200dp_theta_a: null pointer");
    if (!workspace) {
        // version 2: the hi/lo split inside the GEMM, no workspace
        if (!aligned(zx, 16) || !aligned(zy, 16) || !aligned(gx, 16) || !aligned(gy, 16))
            return fail(-1, "b200dp_theta_a: embeddings must be 16-byte aligned");
        Gemm2Params q;
        q.x[0] = zx; q.x[1] = gx; q.y[0] = zy; q.y[1] = gy;
        q.theta = theta; q.A = A; q.xlen = xlen; q.ylen = ylen; q.pair_off = pair_off;
        q.B = B; q.Lx = Lx; q.Ly = Ly; q.D = D;
        if (int rc = set_smem(softdp_theta_a2_kernel, kG2Smem, "b200dp_theta_a")) return rc;
        const dim3 grid2((Ly + kGN - 1) / kGN, (Lx + kGM - 1) / kGM, 2 * B);
        softdp_theta_a2_kernel<<<grid2, kG2Threads, kG2Smem, reinterpret_cast<cudaStream_t>(stream)>>>(q);
        cudaError_t e2 = cudaGetLastError();
        if (e2 != cudaSuccess) return cuda_fail(e2, "b200dp_theta_a launch");
        return 0;
    }
    if (!aligned(zx, 16) || !aligned(zy, 16) || !aligned(gx, 16) || !aligned(gy, 16) || !aligned(workspace, 256))
        return fail(-1, "b200dp_theta_a: embeddings must be 16-byte, the workspace 256-byte aligned");
    if (workspace_bytes < b200dp_theta_a_workspace(B, Lx, Ly, D)) return fail(-1, "b200dp_theta_a: workspace too small");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    unsigned char* ws = static_cast<unsigned char*>(workspace);
    const size_t sx = align256((size_t)B * Lx * D * 2), sy = align256((size_t)B * Ly * D * 2);
    // order: zx, zy, gx, gy
    SplitParams sp;
    const float* srcs[4] = {zx, zy, gx, gy};
    const long long cnt[4] = {(long long)B * Lx * D, (long long)B * Ly * D, (long long)B * Lx * D, (long long)B * Ly * D};
    size_t off = 0;
    for (int w = 0; w < 4; ++w) {
        const size_t sz = (w & 1) ? sy : sx;
        sp.src[w] = srcs[w];
        sp.hi[w] = reinterpret_cast<__nv_bfloat16*>(ws + off);
        sp.lo[w] = reinterpret_cast<__nv_bfloat16*>(ws + off + sz);
        sp.n[w] = cnt[w];
        off += 2 * sz;
    }
    long long maxn = cnt[0] > cnt[1] ? cnt[0] : cnt[1];
    int gx_ = (int)((maxn / 4 + 255) / 256);
    if (gx_ > 148 * 16) gx_ = 148 * 16;
    if (gx_ < 1) gx_ = 1;
    softdp_split_kernel<<<dim3(gx_, 4), 256, 0, st>>>(sp);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "b200dp_theta_a split launch");

    GemmMaps maps;
    memset(&maps, 0, sizeof(maps));
    for (int which = 0; which < 2; ++which)
        for (int op = 0; op < 2; ++op) {
            const int w = 2 * which + op;
            const int L = op ? Ly : Lx;
            if (!encode_bf16_map(&maps.m[which][op][0], sp.hi[w], B, L, D) ||
                !encode_bf16_map(&maps.m[which][op][1], sp.lo[w], B, L, D))
                return fail(-2, "b200dp_theta_a: cuTensorMapEncodeTiled failed");
        }
    GemmParams p;
    p.theta = theta;
    p.A = A;
    p.xlen = xlen;
    p.ylen = ylen;
    p.pair_off = pair_off;
    p.B = B;
    p.Lx = Lx;
    p.Ly = Ly;
    p.D = D;
    if (int rc = set_smem(softdp_theta_a_kernel, kGSmem, "b200dp_theta_a")) return rc;
    const dim3 grid((Ly + kGN - 1) / kGN, (Lx + kGM - 1) / kGM, 2 * B);
    softdp_theta_a_kernel<<<grid, kGThreads, kGSmem, st>>>(maps, p);
    e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "b200dp_theta_a launch");
    return 0;
}

}  // extern "C"
