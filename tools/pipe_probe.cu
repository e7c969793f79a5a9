// tools/pipe_probe.cu -- issue-rate probes (not part of the product): 3-register FFMA, LOP3, FADD, and mixes.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/pipe_probe tools/pipe_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) probe(float* sink, const float* in, int iters)
{
    float a[8], b[8];
    unsigned u[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] = in[threadIdx.x + i]; b[i] = in[threadIdx.x + 8 + i]; u[i] = __float_as_uint(in[threadIdx.x + 16 + i]); }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int j = (i + 1) & 7, k = (i + 3) & 7;
            if (MODE == 0) a[i] = fmaf(a[i], b[j], b[k]);                           // FFMA, 3 distinct registers
            if (MODE == 1) u[i] = (u[i] ^ u[j]) & u[k];                             // LOP3, 3 registers (different lut per step)
            if (MODE == 2) a[i] = a[i] + b[j];                                      // FADD
            if (MODE == 3) { a[i] = fmaf(a[i], b[j], b[k]); u[i] = (u[i] ^ u[j]) & u[k]; }   // FFMA + LOP3 alternating
            if (MODE == 4) { a[i] = fmaf(a[i], b[j], b[k]); b[i] = fmaf(b[i], a[j], a[k]); } // 2 FFMA
            if (MODE == 5) { a[i] = fmaf(a[i], 1.0001f, b[k]); }                    // FFMA with immediate
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i] + b[i] + __uint_as_float(u[i]);
    if (s == 123456.f) sink[0] = s;
}

template <int MODE>
void run(const char* name, int sms, int bps, int per_iter)
{
    float *sink, *in; cudaMalloc(&sink, 4); cudaMalloc(&in, 4096); cudaMemset(in, 0x3f, 4096);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 8192, grid = sms * bps;
    double best = 0;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        probe<MODE><<<grid, 256>>>(sink, in, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double ipc = (double)grid * 8 /*warps*/ * per_iter * iters / (ms * 1e-3) / (sms * 4.0) / 1.965e9;
        if (rep && ipc > best) best = ipc;
    }
    printf("%-40s warps/SM %3d : %.3f warp-instr/clk/scheduler\n", name, bps * 8, best);
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    for (int bps : {8, 2}) {
        run<0>("FFMA 3-reg", sms, bps, 8);
        run<5>("FFMA reg,imm,reg", sms, bps, 8);
        run<1>("LOP3 3-reg", sms, bps, 8);
        run<2>("FADD", sms, bps, 8);
        run<3>("FFMA + LOP3 alternating", sms, bps, 16);
        run<4>("2 x FFMA", sms, bps, 16);
    }
    return 0;
}
