// Micro-benchmark (tuning tool, not part of libssw): issue rate of the packed FP32 instructions of sm_100
// (add/mul/fma.f32x2 -> FADD2/FMUL2/FFMA2) against their scalar forms, per SM sub-partition.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_f32x2 ubench_f32x2.cu
// Prints flop/clk/SM for: scalar FFMA, packed FFMA2, scalar FADD, packed FADD2, and a mix that resembles the
// FFT butterflies (adds + fmas with shared-memory traffic in between is NOT modelled: pure issue/pipe rate).
#include <cstdio>
#include <cuda_runtime.h>

#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__device__ __forceinline__ unsigned long long pk(float a, float b) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ unsigned long long fadd2(unsigned long long a, unsigned long long b) {
    unsigned long long d;
    asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}

constexpr int ILP = 8, ITERS = 4096;

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, float s, long long* clk) {
    const long long t0 = clock64();
    float acc = 0.f;
    if constexpr (MODE == 0) {          // scalar FFMA, 2*ILP independent chains
        float v[2 * ILP];
#pragma unroll
        for (int i = 0; i < 2 * ILP; ++i) v[i] = threadIdx.x + i;
        for (int it = 0; it < ITERS; ++it) {
#pragma unroll
            for (int i = 0; i < 2 * ILP; ++i) v[i] = __fmaf_rn(v[i], s, 1.0f);
        }
#pragma unroll
        for (int i = 0; i < 2 * ILP; ++i) acc += v[i];
    } else if constexpr (MODE == 1) {   // packed FFMA2, ILP independent chains (same flops as mode 0)
        unsigned long long v[ILP];
        const unsigned long long ss = pk(s, s), one = pk(1.f, 1.f);
#pragma unroll
        for (int i = 0; i < ILP; ++i) v[i] = pk(threadIdx.x + i, threadIdx.x - i);
        for (int it = 0; it < ITERS; ++it) {
#pragma unroll
            for (int i = 0; i < ILP; ++i) v[i] = ffma2(v[i], ss, one);
        }
#pragma unroll
        for (int i = 0; i < ILP; ++i) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v[i])); acc += a + b; }
    } else if constexpr (MODE == 2) {   // scalar FADD
        float v[2 * ILP];
#pragma unroll
        for (int i = 0; i < 2 * ILP; ++i) v[i] = threadIdx.x + i;
        for (int it = 0; it < ITERS; ++it) {
#pragma unroll
            for (int i = 0; i < 2 * ILP; ++i) v[i] = __fadd_rn(v[i], s);
        }
#pragma unroll
        for (int i = 0; i < 2 * ILP; ++i) acc += v[i];
    } else {                            // packed FADD2
        unsigned long long v[ILP];
        const unsigned long long ss = pk(s, s);
#pragma unroll
        for (int i = 0; i < ILP; ++i) v[i] = pk(threadIdx.x + i, threadIdx.x - i);
        for (int it = 0; it < ITERS; ++it) {
#pragma unroll
            for (int i = 0; i < ILP; ++i) v[i] = fadd2(v[i], ss);
        }
#pragma unroll
        for (int i = 0; i < ILP; ++i) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v[i])); acc += a + b; }
    }
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

template <int MODE>
int run(const char* name, int warps_per_sm) {
    float* out; long long* clk;
    const int threads = 256, blocks_per_sm = warps_per_sm * 32 / threads;
    int sms; CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    const int blocks = sms * blocks_per_sm;
    CHECK(cudaMalloc(&out, sizeof(float) * blocks * threads));
    CHECK(cudaMalloc(&clk, sizeof(long long) * blocks));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<blocks, threads>>>(out, 0.999f, clk);
    CHECK(cudaDeviceSynchronize());
    cudaEventRecord(e0);
    k<MODE><<<blocks, threads>>>(out, 0.999f, clk);
    cudaEventRecord(e1);
    CHECK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long c0; CHECK(cudaMemcpy(&c0, clk, sizeof(c0), cudaMemcpyDeviceToHost));
    const bool fma = (MODE < 2);
    const double flop_per_thread = (double)ITERS * 2 * ILP * (fma ? 2 : 1);
    const double flops = flop_per_thread * threads * blocks;
    printf("%-12s warps/SM %2d: %8.3f ms  %7.2f TFLOP/s  %6.1f flop/clk/SM (CTA0 %lld clk)\n", name, warps_per_sm, ms,
           flops / ms * 1e-9, flop_per_thread * threads * blocks_per_sm / (double)c0, c0);
    cudaFree(out); cudaFree(clk);
    return 0;
}

int main() {
    for (int w : {8, 16, 32, 64}) {
        if (run<0>("FFMA", w)) return 1;
        if (run<1>("FFMA2", w)) return 1;
        if (run<2>("FADD", w)) return 1;
        if (run<3>("FADD2", w)) return 1;
    }
    return 0;
}
