// tools/mufu_probe.cu -- micro-benchmarks behind the SFU roofline of DESIGN.md (not part of the product).
// Measures, on the GPU it runs on: MUFU.EX2 / MUFU.LG2 / mixed throughput, and the same MUFU stream
// interleaved with shared-memory traffic and FMA-pipe filler in the proportions of the resident kernel
// (per edge: 3 MUFU, 3 LDS/STS, ~16 FMA/ALU).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/mufu_probe tools/mufu_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ float ex2a(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lg2a(float x) { float y; asm volatile("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

template <int MODE, int FILL>
__global__ void __launch_bounds__(256) probe(float* sink, int iters)
{
    extern __shared__ float sm[];
    float a[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) a[i] = 1.0f + 0.001f * (float)(threadIdx.x + i);
    float f0 = 1.f, f1 = 0.5f;
    const int t = threadIdx.x;
    sm[t] = 1.f; sm[t + 256] = 2.f;
    __syncthreads();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 6; i += 3) {
            if (MODE == 0) { a[i] = ex2a(-a[i]) + 1.f; a[i + 1] = ex2a(-a[i + 1]) + 1.f; a[i + 2] = ex2a(-a[i + 2]) + 1.f; }
            if (MODE == 1) { a[i] = lg2a(a[i]) + 2.f; a[i + 1] = lg2a(a[i + 1]) + 2.f; a[i + 2] = lg2a(a[i + 2]) + 2.f; }
            if (MODE >= 2) { a[i] = ex2a(-a[i]) + 1.f; a[i + 1] = lg2a(a[i + 1]) + 2.f; a[i + 2] = lg2a(a[i + 2]) + 2.f; }
            if (MODE == 3) {            // + 2 LDS + 1 STS per 3 MUFU
                float u = sm[(t + i) & 255], v = sm[256 + ((t + it) & 255)];
                sm[(t + 7 * i) & 255] = u + v;
                a[i] += u * 1e-9f;
            }
#pragma unroll
            for (int q = 0; q < FILL; ++q) { f0 = fmaf(f0, 1.0001f, f1); f1 = fmaf(f1, 0.9999f, f0 * 1e-9f); }
        }
    }
    float s = f0 + f1;
#pragma unroll
    for (int i = 0; i < 6; ++i) s += a[i];
    if (s == 123456.f) sink[0] = s;
}

template <int MODE, int FILL>
void run(const char* name, int sms, int blocks_per_sm, int threads)
{
    float* sink; cudaMalloc(&sink, 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 8192, grid = sms * blocks_per_sm;
    double best = 0;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        probe<MODE, FILL><<<grid, threads, 2048>>>(sink, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double ops = (double)grid * threads * 6.0 * iters / (ms * 1e-3);
        if (rep && ops > best) best = ops;
    }
    printf("%-44s warps/SM %3d  fill %2d : %8.1f G MUFU/s\n", name, blocks_per_sm * threads / 32, FILL * 2, best / 1e9);
    cudaFree(sink);
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    printf("%s, %d SMs, %d MHz\n", p.name, sms, p.clockRate / 1000);
    run<0, 0>("ex2 only", sms, 8, 256);
    run<1, 0>("lg2 only", sms, 8, 256);
    run<2, 0>("ex2 + 2 lg2", sms, 8, 256);
    run<2, 0>("ex2 + 2 lg2, 16 warps", sms, 2, 256);
    run<2, 0>("ex2 + 2 lg2, 8 warps", sms, 1, 256);
    run<2, 4>("ex2 + 2 lg2 + 8 FFMA per 3 MUFU", sms, 8, 256);
    run<2, 8>("ex2 + 2 lg2 + 16 FFMA per 3 MUFU", sms, 8, 256);
    run<2, 10>("ex2 + 2 lg2 + 20 FFMA per 3 MUFU", sms, 8, 256);
    run<3, 0>("ex2 + 2 lg2 + 2 LDS + STS", sms, 8, 256);
    run<3, 8>("ex2 + 2 lg2 + 2 LDS + STS + 16 FFMA", sms, 8, 256);
    run<3, 8>("same, 16 warps/SM", sms, 2, 256);
    run<3, 8>("same, 12 warps/SM", sms, 3, 128);
    run<3, 6>("ex2 + 2 lg2 + 2 LDS + STS + 12 FFMA", sms, 8, 256);
    return 0;
}
