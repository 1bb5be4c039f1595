// micro_mem.cu -- which WRITE patterns reach the DRAM write rate on B200?  The chained forward stores Q as
// 1024 concurrent streams (one per warp), 256 B per wavefront step (two 128-byte STG); with the cell
// arithmetic removed the kernel still takes 0.204 ms at C2 (scripts/gpu_x16.py), stores alone 0.106 ms
// (5.06 TB/s).  Modes (all write `streams` x `bytes` contiguous pieces, one warp per stream unless noted):
//   0  grid-stride float4 stores, 256-thread CTAs (the memset-like ceiling)
//   1  one warp per CTA per stream: two 128 B STG per step (the forward's pattern)
//   2  one 256 B STG.64 per step
//   3  one 512 B STG.128 per two steps
//   4  4 KB tiles staged in shared memory, one cp.async.bulk store per 16 steps (double-buffered)
//   5  as 1, the steps throttled by a dependent shuffle + FMA chain of `chain` links (the skeleton)
//   6  as 4, throttled like 5 (all streams advance together in small pieces: does the piece size matter then?)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/_bin/micro_mem scripts/micro_mem.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__global__ void w_ideal(float4* out, size_t n4) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const float4 v = make_float4(1.f, 2.f, 3.f, 4.f);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) out[i] = v;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int MODE>
__global__ void __launch_bounds__(32) w_stream(float* out, long long stream_floats, long long stride_floats, int nstreams,
                                              int chain) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int t = threadIdx.x;
    for (int s = blockIdx.x; s < nstreams; s += gridDim.x) {
        float* base = out + (long long)s * stride_floats;
        const int nsteps = (int)(stream_floats / 64);
        float h = (float)t, v = 1.f;
        if (MODE == 1 || MODE == 5) {
            for (int i = 0; i < nsteps; ++i) {
                if (MODE == 5) {
                    float r = __shfl_sync(0xffffffffu, h, (t + 31) & 31);
                    for (int k = 0; k < chain; ++k) r = fmaf(r, 0.5f, v);
                    h = r;
                    v = fmaf(v, 0.25f, 1.f);
                }
                base[i * 64 + t] = h;
                base[i * 64 + 32 + t] = v;
            }
        } else if (MODE == 2) {
            for (int i = 0; i < nsteps; ++i) reinterpret_cast<float2*>(base + i * 64)[t] = make_float2(h, v);
        } else if (MODE == 3) {
            for (int i = 0; i < nsteps; i += 2) reinterpret_cast<float4*>(base + i * 64)[t] = make_float4(h, v, h, v);
        } else if (MODE == 4 || MODE == 6) {
            float* stage = reinterpret_cast<float*>(smem);
            int buf = 0;
            for (int i = 0; i < nsteps; i += 16) {
                float* sb = stage + buf * 1024;
                // the buffer's previous bulk store must have READ its data
                if (t == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                __syncwarp();
#pragma unroll
                for (int ss = 0; ss < 16; ++ss) {
                    if (MODE == 6) {
                        float r = __shfl_sync(0xffffffffu, h, (t + 31) & 31);
                        for (int k = 0; k < chain; ++k) r = fmaf(r, 0.5f, v);
                        h = r;
                        v = fmaf(v, 0.25f, 1.f);
                    }
                    sb[ss * 64 + t] = h;
                    sb[ss * 64 + 32 + t] = v;
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (t == 0) {
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], 4096;" ::"l"(base + i * 64),
                                 "r"(smem_u32(sb))
                                 : "memory");
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
                buf ^= 1;
            }
            if (t == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
            __syncwarp();
        }
    }
}

int main(int argc, char** argv) {
    const int streams = argc > 1 ? atoi(argv[1]) : 1024;
    const long long stream_bytes = argc > 2 ? atoll(argv[2]) : 524288;       // C2: 8 strips x 256 steps x 256 B
    const long long pad_bytes = argc > 3 ? atoll(argv[3]) : 7936;            // the 31 ramp steps of the Q layout
    const long long stride_floats = (stream_bytes + pad_bytes) / 4, stream_floats = stream_bytes / 4;
    const size_t total = (size_t)streams * stride_floats * 4;
    float* buf;
    CK(cudaMalloc(&buf, total + (1 << 20)));
    CK(cudaMemset(buf, 0, total));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    const double gb = (double)streams * stream_bytes / 1e9;
    auto time = [&](const char* name, auto launch) {
        for (int i = 0; i < 3; ++i) launch();
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0));
        const int it = 20;
        for (int i = 0; i < it; ++i) launch();
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        ms /= it;
        printf("%-44s %8.4f ms  %7.1f GB/s\n", name, ms, gb / (ms * 1e-3));
    };
    printf("streams %d x %lld B (+%lld pad) = %.1f MB\n", streams, stream_bytes, pad_bytes, gb * 1e3);
    time("0 grid-stride float4 (ideal)", [&] { w_ideal<<<148 * 8, 256>>>(reinterpret_cast<float4*>(buf), (size_t)streams * stream_bytes / 16); });
    CK(cudaFuncSetAttribute(w_stream<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192));
    CK(cudaFuncSetAttribute(w_stream<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192));
    for (int grid : {streams}) {
        if (grid > streams) continue;
        char nm[96];
        // 18 KB of dynamic shared memory per CTA in modes 1-3 / 5 gives the forward's residency (about 11 per SM)
        snprintf(nm, sizeof nm, "1 2x128B STG per step, grid %d", grid);
        time(nm, [&] { w_stream<1><<<grid, 32, 0>>>(buf, stream_floats, stride_floats, streams, 0); });
        snprintf(nm, sizeof nm, "2 256B STG.64 per step, grid %d", grid);
        time(nm, [&] { w_stream<2><<<grid, 32, 0>>>(buf, stream_floats, stride_floats, streams, 0); });
        snprintf(nm, sizeof nm, "3 512B STG.128 per 2 steps, grid %d", grid);
        time(nm, [&] { w_stream<3><<<grid, 32, 0>>>(buf, stream_floats, stride_floats, streams, 0); });
        snprintf(nm, sizeof nm, "4 4KB bulk store per 16 steps, grid %d", grid);
        time(nm, [&] { w_stream<4><<<grid, 32, 8192>>>(buf, stream_floats, stride_floats, streams, 0); });
        for (int chain : {1, 4, 8}) {
            snprintf(nm, sizeof nm, "5 as 1, chain %d, grid %d", chain, grid);
            time(nm, [&] { w_stream<5><<<grid, 32, 0>>>(buf, stream_floats, stride_floats, streams, chain); });
            snprintf(nm, sizeof nm, "6 as 4, chain %d, grid %d", chain, grid);
            time(nm, [&] { w_stream<6><<<grid, 32, 8192>>>(buf, stream_floats, stride_floats, streams, chain); });
        }
    }
    return 0;
}
