// micro_lat.cu -- dependent-chain latencies of the instructions the soft-DP step is made of, and the
// per-step cost of the forward / backward cell recurrences for ONE warp with no memory traffic
// (what bounds a latency-bound strip).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3
// -o scripts/_bin/micro_lat scripts/micro_lat.cu ; run on the GPU box, prints cycles per iteration.
#include <cstdio>
#include <cuda_runtime.h>

#define FULL 0xffffffffu
__device__ __forceinline__ float ex2f(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lg2f(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcpf(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

constexpr int IT = 4096;

template <int WHICH>
__global__ void chain(float* out, long long* cyc, float seed) {
    float x = seed + threadIdx.x * 1e-3f, y = seed * 0.5f, z = 0.25f;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < IT; ++i) {
        if (WHICH == 0) x = x * 1.0001f + 0.5f;                                  // FFMA
        if (WHICH == 1) x = __shfl_up_sync(FULL, x, 1);                           // SHFL
        if (WHICH == 2) x = ex2f(x) - 1.0f;                                       // EX2 + FADD
        if (WHICH == 3) x = lg2f(x) + 3.0f;                                       // LG2 + FADD
        if (WHICH == 4) x = rcpf(x) + 0.5f;                                       // RCP + FADD
        if (WHICH == 5) x = fmaxf(x, y) + 0.001f;                                 // FMNMX + FADD
        if (WHICH == 6) { x = __shfl_up_sync(FULL, x, 1) + 0.001f; }              // SHFL + FADD
        if (WHICH == 7) { x = __shfl_sync(FULL, x, (threadIdx.x + 31) & 31) + 0.001f; }   // SHFL.IDX + FADD
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    out[blockIdx.x * 32 + threadIdx.x] = x + y + z;
}

// forward cell step (softdp_fwd2.cuh fwd2_step, steady form), NC independent chains per warp
template <int NC, int STORE>
__global__ void fwd_step(float* out, long long* cyc, const float* __restrict__ th, const float* __restrict__ aa, float* q) {
    q += (size_t)blockIdx.x * 1024 * 64;
    float h[NC], v[NC];
    for (int c = 0; c < NC; ++c) { h[c] = 0.1f * c; v[c] = 0.2f; }
    const float t_ = th[threadIdx.x], a_ = aa[threadIdx.x];
    float part = 0.f;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < IT; ++i) {
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            float hup = __shfl_up_sync(FULL, h[c], 1);
            hup = threadIdx.x == 0 ? 0.f : hup;
            const float dx = fmaf(a_, 1.4426950408889634f, hup);
            const float dy = fmaf(a_, 1.4426950408889634f, v[c]);
            const float mx = fmaxf(fmaxf(dx, dy), 0.f);
            const float em = ex2f(-mx), ex = ex2f(dx - mx), ey = ex2f(dy - mx);
            const float S = (em + ex) + ey;
            const float r = rcpf(S);
            float qx = fminf(ex * r, 1.f);
            float qy = fminf(ey * r, 1.f - qx);
            const float l = lg2f(S) + fmaf(t_, 1.4426950408889634f, mx);
            h[c] = l - v[c];
            v[c] = l - hup;
            part += h[c];
            if (STORE == 1) {
                q[(size_t)(i & 1023) * 64 + threadIdx.x] = qx;
                q[(size_t)(i & 1023) * 64 + 32 + threadIdx.x] = qy;
            } else if (STORE == 2) {
                reinterpret_cast<float2*>(q)[(size_t)(i & 1023) * 32 + threadIdx.x] = make_float2(qx, qy);
            } else {
                part += qx * 1e-9f + qy * 1e-9f;
            }
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    out[blockIdx.x * 32 + threadIdx.x] = part + h[0] + v[NC - 1];
}

// linear (ratio) form of the same recurrence: H = W[i,j]/W[i,j-1], V = W[i,j]/W[i-1,j]
template <int NC>
__global__ void lin_step(float* out, long long* cyc, const float* __restrict__ th, const float* __restrict__ aa) {
    float H[NC], V[NC], iV[NC];
    for (int c = 0; c < NC; ++c) { H[c] = 1.f + 0.1f * c; V[c] = 1.1f; iV[c] = 0.9f; }
    const float eT = ex2f(th[threadIdx.x]), eA = ex2f(aa[threadIdx.x]);
    float part = 0.f;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < IT; ++i) {
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            float Hup = __shfl_up_sync(FULL, H[c], 1);
            Hup = threadIdx.x == 0 ? 1.f : Hup;
            const float iHup = rcpf(Hup);
            const float S = fmaf(eA, Hup + V[c], 1.f);
            const float L = eT * S;
            H[c] = L * iV[c];
            const float Vn = L * iHup;
            const float iH = rcpf(H[c]);
            const float rS = (eT * iV[c]) * iH;
            const float tt = eA * rS;
            const float qx = fminf(tt * Hup, 1.f);
            const float qy = fminf(tt * V[c], 1.f - qx);
            iV[c] = rcpf(Vn);
            V[c] = Vn;
            part += lg2f(H[c]) + qx * 1e-9f + qy * 1e-9f;
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    out[blockIdx.x * 32 + threadIdx.x] = part + H[0] + V[NC - 1];
}

// backward cell step (softdp_sq.cuh steady form)
template <int NC>
__global__ void bwd_step(float* out, long long* cyc, const float* __restrict__ th, const float* __restrict__ aa) {
    float zout[NC], dprev[NC], yprev[NC];
    for (int c = 0; c < NC; ++c) { zout[c] = 0.1f; dprev[c] = 0.2f; yprev[c] = 0.3f; }
    const float qx = th[threadIdx.x] * 0.3f, qy = aa[threadIdx.x] * 0.3f, qm = (1.f - qx) - qy;
    float part = 0.f;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < IT; ++i) {
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            float zin = __shfl_down_sync(FULL, zout[c], 1);
            if (threadIdx.x == 31) zin = 0.01f;
            const float e = zin + yprev[c];
            zout[c] = fmaf(qx, e, dprev[c]);
            dprev[c] = qm * e;
            yprev[c] = qy * e;
            part += e;
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    out[blockIdx.x * 32 + threadIdx.x] = part + zout[0];
}

template <typename F>
static void run(const char* name, F launch, long long* dcyc, int ops = 1) {
    launch(1);
    cudaDeviceSynchronize();
    for (int warps : {1, 4, 8, 16}) {      // CTAs of one warp on ONE SM are not controllable; use many CTAs = all SMs
        launch(warps);
        cudaError_t e = cudaDeviceSynchronize();
        long long c = 0;
        cudaMemcpy(&c, dcyc, 8, cudaMemcpyDeviceToHost);
        printf("%-34s warps/SM %2d: %7.1f cycles/iter%s\n", name, warps, (double)c / IT / ops,
               e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
}

int main() {
    float *out, *th, *aa, *q;
    long long* cyc;
    cudaMalloc(&out, 148 * 64 * 32 * 4);
    cudaMalloc(&cyc, 148 * 64 * 8);
    cudaMalloc(&th, 128);
    cudaMalloc(&aa, 128);
    cudaMalloc(&q, (size_t)1024 * 64 * 4 * 148 * 16);
    float hth[32], haa[32];
    for (int i = 0; i < 32; ++i) { hth[i] = 0.3f + 0.01f * i; haa[i] = -0.5f - 0.01f * i; }
    cudaMemcpy(th, hth, 128, cudaMemcpyHostToDevice);
    cudaMemcpy(aa, haa, 128, cudaMemcpyHostToDevice);
    const int SM = 148;
#define CH(W, NAME) run(NAME, [&](int w) { chain<W><<<SM * w, 32>>>(out, cyc, 1.5f); }, cyc)
    CH(0, "FFMA chain");
    CH(1, "SHFL.UP chain");
    CH(2, "EX2+FADD chain");
    CH(3, "LG2+FADD chain");
    CH(4, "RCP+FADD chain");
    CH(5, "FMNMX+FADD chain");
    CH(6, "SHFL.UP+FADD chain");
    CH(7, "SHFL.IDX+FADD chain");
    run("fwd step log-domain, 1 chain", [&](int w) { fwd_step<1, 0><<<SM * w, 32>>>(out, cyc, th, aa, q); }, cyc);
    run("fwd step log-domain, 2 chains", [&](int w) { fwd_step<2, 0><<<SM * w, 32>>>(out, cyc, th, aa, q); }, cyc);
    run("fwd step log-domain +STG, 1 chain", [&](int w) { fwd_step<1, 1><<<SM * w, 32>>>(out, cyc, th, aa, q + (size_t)0); }, cyc);
    run("fwd step log-domain +STG.64, 1 chain", [&](int w) { fwd_step<1, 2><<<SM * w, 32>>>(out, cyc, th, aa, q + (size_t)0); }, cyc);
    run("fwd step linear ratio, 1 chain", [&](int w) { lin_step<1><<<SM * w, 32>>>(out, cyc, th, aa); }, cyc);
    run("fwd step linear ratio, 2 chains", [&](int w) { lin_step<2><<<SM * w, 32>>>(out, cyc, th, aa); }, cyc);
    run("bwd step, 1 chain", [&](int w) { bwd_step<1><<<SM * w, 32>>>(out, cyc, th, aa); }, cyc);
    run("bwd step, 2 chains", [&](int w) { bwd_step<2><<<SM * w, 32>>>(out, cyc, th, aa); }, cyc);
    return 0;
}
