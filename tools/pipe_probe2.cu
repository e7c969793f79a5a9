// tools/pipe_probe2.cu -- round-2 issue-rate probes (not part of the product): packed fp32 (FFMA2 / FADD2 /
// FMUL2, sm_100a), 64-bit shared-memory accesses, and MUFU mixed with packed FMA work in the ratio of the
// two-frames-per-thread resident kernel (per check edge and PAIR of frames: 6 MUFU, ~8 FFMA2/FMUL2,
// ~3 FADD2, ~4 LOP3, 2 FMNMX, 3 LDS/STS.64, ~3 integer).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/pipe_probe2 tools/pipe_probe2.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ float ex2f(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lg2f(float x) { float y; asm volatile("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

template <int MODE>
__global__ void __launch_bounds__(256) probe(float* sink, const float* in, int iters)
{
    __shared__ float2 sh[256 * 4];
    float2 a[8], b[8];
    unsigned u[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        a[i] = make_float2(in[threadIdx.x + i], in[threadIdx.x + 32 + i]);
        b[i] = make_float2(in[threadIdx.x + 8 + i], in[threadIdx.x + 40 + i]);
        u[i] = __float_as_uint(in[threadIdx.x + 16 + i]);
    }
    for (int i = threadIdx.x; i < 1024; i += 256) sh[i] = make_float2(1.f, 2.f);
    __syncthreads();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int j = (i + 1) & 7, k = (i + 3) & 7;
            if (MODE == 0) a[i] = __ffma2_rn(a[i], b[j], b[k]);                 // FFMA2, 3 distinct register pairs
            if (MODE == 1) a[i] = __fadd2_rn(a[i], b[j]);                       // FADD2
            if (MODE == 2) a[i] = __fmul2_rn(a[i], b[j]);                       // FMUL2
            if (MODE == 3) { a[i].x = fmaf(a[i].x, b[j].x, b[k].x); a[i].y = fmaf(a[i].y, b[j].y, b[k].y); }   // 2 scalar FFMA
            if (MODE == 4) {                                                    // LDS.64 + FADD2 + STS.64 (the scatter)
                float2 v = sh[threadIdx.x + 256 * (i & 3)];
                v = __fadd2_rn(v, a[i]);
                sh[threadIdx.x + 256 * ((i + 1) & 3)] = v;
            }
            if (MODE == 5) {                                                    // MUFU only: 2 per step
                a[i].x = ex2f(a[i].x); a[i].y = lg2f(a[i].y);
            }
            if (MODE == 6) {                                                    // kernel-like mix per (edge, frame pair)
                // 6 MUFU, 8 FFMA2/FMUL2, 3 FADD2, 4 LOP3, 2 FMNMX
                float2 m = __fadd2_rn(a[i], b[j]);                              // mu = L - eps
                u[i] ^= __float_as_uint(m.x); u[j] ^= __float_as_uint(m.y);     // sign accumulate
                float2 x;
                x.x = ex2f(-fminf(fabsf(m.x), 50.5f)); x.y = ex2f(-fminf(fabsf(m.y), 50.5f));
                float2 fa = __ffma2_rn(b[k], x, a[j]);                          // prefix
                float2 fb = __ffma2_rn(a[j], x, b[k]);
                float2 A = __ffma2_rn(fa, a[k], __fmul2_rn(fb, b[j]));          // combine
                float2 B = __ffma2_rn(fa, b[j], __fmul2_rn(fb, a[k]));
                float2 sa = __ffma2_rn(b[j], x, a[k]);                          // suffix
                float2 sb = __ffma2_rn(a[k], x, b[j]);
                float2 mag = __fadd2_rn(make_float2(lg2f(A.x), lg2f(A.y)), make_float2(-lg2f(B.x), -lg2f(B.y)));
                mag.x = __uint_as_float(__float_as_uint(mag.x) | (u[i] & 0x80000000u));
                mag.y = __uint_as_float(__float_as_uint(mag.y) | (u[j] & 0x80000000u));
                a[i] = __fadd2_rn(mag, sa);
                b[i] = __fadd2_rn(b[i], sb);
            }
            if (MODE == 7) {                                                    // same mix, scalar (one frame): 3 MUFU, 8 FFMA/FMUL ...
                float m = a[i].x + b[j].x;
                u[i] ^= __float_as_uint(m);
                float x = ex2f(-fminf(fabsf(m), 50.5f));
                float fa = fmaf(b[k].x, x, a[j].x), fb = fmaf(a[j].x, x, b[k].x);
                float A = fmaf(fa, a[k].x, fb * b[j].x), B = fmaf(fa, b[j].x, fb * a[k].x);
                float sa = fmaf(b[j].x, x, a[k].x), sb = fmaf(a[k].x, x, b[j].x);
                float mag = lg2f(A) - lg2f(B);
                mag = __uint_as_float(__float_as_uint(mag) | (u[i] & 0x80000000u));
                a[i].x = mag + sa;
                b[i].x = b[i].x + sb;
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i].x + a[i].y + b[i].x + b[i].y + __uint_as_float(u[i]);
    if (MODE == 4) s += sh[threadIdx.x].x;
    if (s == 123456.f) sink[0] = s;
}

template <int MODE>
void run(const char* name, int sms, int bps, double units_per_step, const char* unit)
{
    float *sink, *in; cudaMalloc(&sink, 4); cudaMalloc(&in, 4096); cudaMemset(in, 0x3f, 4096);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 4096, grid = sms * bps;
    double best = 0;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        probe<MODE><<<grid, 256>>>(sink, in, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        // steps per clock per scheduler (a step = one body of the unrolled loop for one warp)
        double rate = (double)grid * 8 /*warps*/ * 8 /*steps*/ * iters / (ms * 1e-3) / (sms * 4.0) / 1.965e9;
        if (rep && rate > best) best = rate;
    }
    printf("%-52s warps/SM %3d : %.3f steps/clk/scheduler = %.3f %s/clk/scheduler  (%.2f clk per step)\n", name, bps * 8, best,
           best * units_per_step, unit, 1.0 / best);
    cudaFree(sink); cudaFree(in);
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    printf("%s, %d SMs, clock assumed 1.965 GHz\n", p.name, sms);
    for (int bps : {8, 2, 1}) {
        run<0>("FFMA2 3 register pairs", sms, bps, 1, "instr");
        run<1>("FADD2", sms, bps, 1, "instr");
        run<2>("FMUL2", sms, bps, 1, "instr");
        run<3>("2 x scalar FFMA 3-reg", sms, bps, 2, "instr");
        run<4>("LDS.64 + FADD2 + STS.64", sms, bps, 3, "instr");
        run<5>("MUFU ex2 + lg2", sms, bps, 2, "MUFU");
        run<6>("kernel-like mix, PAIR of frames (6 MUFU/step)", sms, bps, 2, "edge-frames");
        run<7>("kernel-like mix, one frame (3 MUFU/step)", sms, bps, 1, "edge-frames");
    }
    return 0;
}
