// micro_tma.cu -- how fast does one SM's TMA unit deliver the chained forward's operand boxes?  The kernel stages
// theta / A as 16 x 16 fp32 boxes (64-byte rows); with the cell arithmetic and the stores removed it still needs
// 0.111 ms for 537 MB at C2 (4.8 TB/s, scripts/gpu_x16.py).  One warp per CTA walks its pair strip by strip like the
// forward does, a ring of `ring_bytes` of boxes in flight, and only waits for the data (no consumption).
//   box widths 16 / 32 / 64 columns x 16 rows, two row groups x two tensors per event
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I deepblast_b200/csrc -o scripts/_bin/micro_tma scripts/micro_tma.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include "softdp_host.h"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) {
        asm volatile(
            "{\n\t.reg .pred P;\n\tmbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\tselp.u32 %0, 1, 0, P;\n\t}"
            : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    }
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
            smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

// events of one pair: strips x column tiles; one event = 2 row groups x 2 tensors boxes of BW x 16
// ST: 0 = loads only; 1 = per event the warp also stores the forward's Q bytes for those cells (BW steps x two 128 B
// STG) to its own contiguous stream; 2 = the same bytes as one bulk store per 16 steps from a shared-memory tile
template <int BW, int ST>
__global__ void __launch_bounds__(32) tma_walk(const __grid_constant__ CUtensorMap mT, const __grid_constant__ CUtensorMap mA,
                                               int npairs, int N, int M, int ring, float* sink, float* q, long long qstride) {
    extern __shared__ __align__(1024) unsigned char smem[];
    constexpr int kBox = BW * 16 * 4, kEvent = 4 * kBox;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)ring * kEvent);     // (ring <= 16)
    const int t = threadIdx.x;
    if (t == 0) {
        for (int i = 0; i < ring; ++i) mbar_init(&bars[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    const int K = N / 32, T = M / BW;
    float acc = 0.f;
    unsigned phases = 0;
    for (int pair = blockIdx.x; pair < npairs; pair += gridDim.x) {
        const int nev = K * T;
        int issued = 0;
        auto issue = [&](int e) {
            if (t == 0) {
                const int slot = e % ring, k = e / T, ct = e % T;
                unsigned char* dst = smem + (size_t)slot * kEvent;
                mbar_expect_tx(&bars[slot], kEvent);
                for (int g = 0; g < 2; ++g) {
                    tma_load_3d(dst + (2 * g) * kBox, &mT, &bars[slot], ct * BW, k * 32 + g * 16, pair);
                    tma_load_3d(dst + (2 * g + 1) * kBox, &mA, &bars[slot], ct * BW, k * 32 + g * 16, pair);
                }
            }
        };
        for (; issued < ring - 1 && issued < nev; ++issued) issue(issued);
        for (int e = 0; e < nev; ++e) {
            __syncwarp();
            if (issued < nev) { issue(issued); ++issued; }
            const int slot = e % ring;
            mbar_wait(&bars[slot], (phases >> slot) & 1u);
            phases ^= 1u << slot;
            acc += reinterpret_cast<const float*>(smem + (size_t)slot * kEvent)[t];
            if (ST == 1) {
                float* qp = q + (long long)pair * qstride + (long long)e * BW * 64 + t;
#pragma unroll
                for (int ss = 0; ss < BW; ++ss) {
                    qp[ss * 64] = acc;
                    qp[ss * 64 + 32] = acc;
                }
            } else if (ST == 2) {
                float* stage = reinterpret_cast<float*>(smem + (size_t)ring * kEvent + 128);
                for (int i = 0; i < BW / 16; ++i) {
                    float* sb = stage + ((e * (BW / 16) + i) & 1) * 1024;
                    if (t == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                    __syncwarp();
#pragma unroll
                    for (int ss = 0; ss < 16; ++ss) {
                        sb[ss * 64 + t] = acc;
                        sb[ss * 64 + 32 + t] = acc;
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (t == 0) {
                        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], 4096;" ::"l"(
                                         q + (long long)pair * qstride + ((long long)e * BW + i * 16) * 64),
                                     "r"(smem_u32(sb))
                                     : "memory");
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    }
                }
            }
        }
        if (ST == 2) {
            if (t == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
            __syncwarp();
        }
    }
    if (acc == 12345.f) sink[0] = acc;
}

// The backward's traffic alone: one warp per CTA takes strips (32 rows of a pair) round-robin, streams the strip's Q
// in as 4 KB bulk copies (ring of `ring` tiles) and writes E row-major as the drain does: per 32-column tile 32 row
// pieces of 128 B (WIDE = 1) or per 64 columns 32 pieces of 256 B (WIDE = 2).  No arithmetic.
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
template <int WIDE>
__global__ void __launch_bounds__(32) bwd_walk(const float* q, long long qstride, float* E, int B, int N, int M, int ring,
                                               float* sink) {
    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)ring * 4096);
    const int t = threadIdx.x;
    if (t == 0) {
        for (int i = 0; i < ring; ++i) mbar_init(&bars[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    const int K = N / 32, T = M / 16;
    unsigned phases = 0;
    float acc = 0.f;
    int ecount = 0;                                    // tiles consumed by this warp (ring position)
    for (int s = blockIdx.x; s < B * K; s += gridDim.x) {
        const int pair = s / K, k = s % K;
        const float* qs = q + (long long)pair * qstride + (long long)k * M * 64;
        int issued = 0;
        auto issue = [&](int e) {
            if (t == 0) {
                const int slot = (ecount + e) % ring;
                mbar_expect_tx(&bars[slot], 4096);
                bulk_load(smem + (size_t)slot * 4096, qs + (long long)e * 1024, 4096, &bars[slot]);
            }
        };
        for (; issued < ring - 1 && issued < T; ++issued) issue(issued);
        for (int e = 0; e < T; ++e) {
            __syncwarp();
            if (issued < T) { issue(issued); ++issued; }
            const int slot = (ecount + e) % ring;
            mbar_wait(&bars[slot], (phases >> slot) & 1u);
            phases ^= 1u << slot;
            acc += reinterpret_cast<const float*>(smem + (size_t)slot * 4096)[t];
            if ((e % (2 * WIDE)) == 2 * WIDE - 1) {
                float* er = E + ((long long)pair * N + k * 32) * M + (e / (2 * WIDE)) * 32 * WIDE;
#pragma unroll 8
                for (int r = 0; r < 32; ++r) {
                    if (WIDE == 1) er[(long long)r * M + t] = acc;
                    else reinterpret_cast<float2*>(er + (long long)r * M)[t] = make_float2(acc, acc);
                }
            }
        }
        ecount += T;
    }
    if (acc == 12345.f) sink[0] = acc;
}

int main(int argc, char** argv) {
    const int B = argc > 1 ? atoi(argv[1]) : 1024, N = argc > 2 ? atoi(argv[2]) : 256, M = argc > 3 ? atoi(argv[3]) : 256;
    const size_t n = (size_t)B * N * M;
    float *th, *a, *sink, *q;
    const long long qstride = (long long)N * M * 2 + 31 * 64;
    CK(cudaMalloc(&q, (size_t)B * qstride * 4));
    CK(cudaMalloc(&th, n * 4));
    CK(cudaMalloc(&a, n * 4));
    CK(cudaMalloc(&sink, 4));
    CK(cudaMemset(th, 0, n * 4));
    CK(cudaMemset(a, 0, n * 4));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    const double gb = 2.0 * n * 4 / 1e9;
    printf("theta + A: %d x %d x %d, %.1f MB\n", B, N, M, gb * 1e3);
    auto run = [&](auto kern, int bw, int ring_bytes, int grid, int st) {
        CUtensorMap mT, mA;
        if (!b200dp_host::encode_row_map(&mT, th, B, N, M, bw, 16) || !b200dp_host::encode_row_map(&mA, a, B, N, M, bw, 16)) {
            printf("encode failed\n");
            exit(1);
        }
        const int ev = 4 * bw * 16 * 4, ring = ring_bytes / ev;
        if (ring < 2) return;
        const size_t smem = (size_t)ring * ev + 128 + 8192;      // ring, barriers, two 4 KB store tiles
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        for (int i = 0; i < 3; ++i) kern<<<grid, 32, smem>>>(mT, mA, B, N, M, ring, sink, q, qstride);
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0));
        const int it = 20;
        for (int i = 0; i < it; ++i) kern<<<grid, 32, smem>>>(mT, mA, B, N, M, ring, sink, q, qstride);
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        ms /= it;
        printf("box %2d x 16  ring %2d events (%3d KB)  grid %4d  stores %d  %8.4f ms  %7.1f GB/s\n", bw, ring, ring_bytes >> 10,
               grid, st, ms, (st ? 2.0 : 1.0) * gb / (ms * 1e-3));
    };
    for (int grid : {B, B / 2}) {
        for (int rb : {16384, 32768}) {
            run(tma_walk<16, 0>, 16, rb, grid, 0);
            run(tma_walk<16, 1>, 16, rb, grid, 1);
            run(tma_walk<16, 2>, 16, rb, grid, 2);
            run(tma_walk<32, 0>, 32, rb, grid, 0);
            run(tma_walk<32, 1>, 32, rb, grid, 1);
            run(tma_walk<32, 2>, 32, rb, grid, 2);
            run(tma_walk<64, 1>, 64, rb, grid, 1);
            run(tma_walk<64, 2>, 64, rb, grid, 2);
        }
    }
    printf("backward traffic: Q in (4 KB bulk tiles), E out (row pieces)\n");
    auto runb = [&](auto kern, int wide, int ring, int grid) {
        const size_t smem = (size_t)ring * 4096 + 128;
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        for (int i = 0; i < 3; ++i) kern<<<grid, 32, smem>>>(q, qstride, th, B, N, M, ring, sink);
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0));
        const int it = 20;
        for (int i = 0; i < it; ++i) kern<<<grid, 32, smem>>>(q, qstride, th, B, N, M, ring, sink);
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        ms /= it;
        printf("E pieces %3d B  ring %d  grid %4d  %8.4f ms  %7.1f GB/s\n", 128 * wide, ring, grid, ms, 1.5 * gb / (ms * 1e-3));
    };
    for (int grid : {148 * 8, 148 * 11, 148 * 12, 148 * 16})
        for (int ring : {2, 3, 4}) {
            runb(bwd_walk<1>, 1, ring, grid);
            runb(bwd_walk<2>, 2, ring, grid);
        }
    return 0;
}
