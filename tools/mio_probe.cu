// tools/mio_probe.cu -- do MUFU (XU pipe) and shared-memory accesses (LSU) overlap, or do they serialise in the memory-IO
// (MIO) queue of a scheduler?  (not part of the product)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/mio_probe tools/mio_probe.cu
// Every mode runs the same loop body per step: NM x MUFU.EX2 and NL x (LDS.64 + STS.64 pair on conflict-free addresses),
// plus a fixed number of FFMA so that the loop is not issue-bound.  If the pipes overlapped, a step would cost
// max(8 NM, c NL) cycles; if they serialise, 8 NM + c NL.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ float2 lds64(uint32_t a) { float2 v; asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a)); return v; }
__device__ __forceinline__ void sts64(uint32_t a, float2 v) { asm volatile("st.shared.v2.f32 [%0], {%1, %2};" :: "r"(a), "f"(v.x), "f"(v.y) : "memory"); }
__device__ __forceinline__ float lds32(uint32_t a) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a)); return v; }
__device__ __forceinline__ void sts32(uint32_t a, float v) { asm volatile("st.shared.f32 [%0], %1;" :: "r"(a), "f"(v) : "memory"); }
__device__ __forceinline__ float ex2f(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

template <int NM, int NL, int WIDE>
__global__ void __launch_bounds__(256) probe(float* sink, const float* in, int iters)
{
    __shared__ float2 sh2[256 * 4];
    __shared__ float sh1[256 * 4];
    float a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = in[threadIdx.x + i];
    for (int i = threadIdx.x; i < 1024; i += 256) { sh2[i] = make_float2(1.f, 2.f); sh1[i] = 1.f; }
    __syncthreads();
    const uint32_t b2 = (uint32_t)__cvta_generic_to_shared(sh2), b1 = (uint32_t)__cvta_generic_to_shared(sh1);
    for (int it = 0; it < iters; ++it) {
        const int rot = (threadIdx.x + it) & 255;          // run-time address
#pragma unroll
        for (int i = 0; i < 8; ++i) {
#pragma unroll
            for (int q = 0; q < NM; ++q) a[(i + q) & 7] = ex2f(a[(i + q) & 7]);
#pragma unroll
            for (int q = 0; q < NL; ++q) {
                if (WIDE) {
                    float2 v = lds64(b2 + 8 * (rot + 256 * ((i + q) & 3)));
                    v.x += a[i];
                    sts64(b2 + 8 * (rot + 256 * ((i + q + 1) & 3)), v);
                } else {
                    float v = lds32(b1 + 4 * (rot + 256 * ((i + q) & 3)));
                    sts32(b1 + 4 * (rot + 256 * ((i + q + 1) & 3)), v + a[i]);
                }
            }
            a[i] = fmaf(a[i], 0.999f, 0.001f);
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i];
    s += sh2[threadIdx.x].x + sh1[threadIdx.x];
    if (s == 123456.f) sink[0] = s;
}

template <int NM, int NL, int WIDE>
void run(int sms, int bps)
{
    float *sink, *in; cudaMalloc(&sink, 4); cudaMalloc(&in, 4096); cudaMemset(in, 0x3f, 4096);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 2048, grid = sms * bps;
    double best = 1e30;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        probe<NM, NL, WIDE><<<grid, 256>>>(sink, in, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        // cycles per step and scheduler: time * clock / (steps per scheduler)
        double steps_per_sched = (double)bps * 8 /*warps*/ * 8 * iters / 4.0;
        double cyc = ms * 1e-3 * 1.965e9 / steps_per_sched;
        if (rep && cyc < best) best = cyc;
    }
    printf("MUFU x%d + (LDS+STS).%s x%d   warps/SM %3d : %6.2f clk per step and scheduler\n", NM, WIDE ? "64" : "32", NL, bps * 8, best);
    cudaFree(sink); cudaFree(in);
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    printf("%s, %d SMs, clock assumed 1.965 GHz\n", p.name, sms);
    for (int bps : {4, 2}) {
        run<2, 0, 0>(sms, bps);
        run<0, 2, 0>(sms, bps);
        run<0, 2, 1>(sms, bps);
        run<2, 2, 0>(sms, bps);
        run<2, 2, 1>(sms, bps);
        run<3, 1, 0>(sms, bps);
        run<3, 1, 1>(sms, bps);
        run<3, 2, 1>(sms, bps);
    }
    return 0;
}
