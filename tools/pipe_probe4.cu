// tools/pipe_probe4.cu -- what does ONE instruction of each kind cost a scheduler that is already busy?
// (not part of the product).  Background: 3 independent FFMA2 per step (saturates the scheduler's dispatch / register
// read bandwidth, 3 x 3.2 cycles); cost(X) = cycles(background + X) - cycles(background), 8 warps per scheduler.
// Every instruction is spelled as volatile inline PTX so that nothing is folded; SASS checked with cuobjdump.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/pipe_probe4 tools/pipe_probe4.cu
#include <cstdio>
#include <cuda_runtime.h>

#define BG(i)                                                                                                         \
    asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(A[i]) : "l"(B[(i + 1) & 7]), "l"(C[(i + 3) & 7]));             \
    asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(B[i]) : "l"(C[(i + 2) & 7]), "l"(A[(i + 5) & 7]));             \
    asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(C[i]) : "l"(A[(i + 6) & 7]), "l"(B[(i + 7) & 7]));

template <int MODE, int NX>
__global__ void __launch_bounds__(128) probe(float* sink, const float* in, int iters)
{
    __shared__ float2 sh[128 * 8];
    unsigned long long A[8], B[8], C[8];
    float f[8], g[8], h[8];
    unsigned u[8], v[8], w[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        A[i] = ((unsigned long long)__float_as_uint(in[threadIdx.x + i]) << 32) | __float_as_uint(in[threadIdx.x + 8 + i]);
        B[i] = ((unsigned long long)__float_as_uint(in[threadIdx.x + 16 + i]) << 32) | __float_as_uint(in[threadIdx.x + 24 + i]);
        C[i] = ((unsigned long long)__float_as_uint(in[threadIdx.x + 32 + i]) << 32) | __float_as_uint(in[threadIdx.x + 40 + i]);
        f[i] = in[threadIdx.x + 48 + i]; g[i] = in[threadIdx.x + 56 + i]; h[i] = in[threadIdx.x + 64 + i];
        u[i] = __float_as_uint(in[threadIdx.x + 72 + i]); v[i] = __float_as_uint(in[threadIdx.x + 80 + i]);
        w[i] = __float_as_uint(in[threadIdx.x + 88 + i]);
    }
    for (int i = threadIdx.x; i < 1024; i += 128) sh[i] = make_float2(1.f, 2.f);
    __syncthreads();
    const unsigned sbase = (unsigned)__cvta_generic_to_shared(sh) + threadIdx.x * 8;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            BG(i)
#pragma unroll
            for (int x = 0; x < NX; ++x) {
                const int j = (i + 1 + x) & 7, k = (i + 3 + x) & 7;
                if (MODE == 1) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(g[j]), "f"(h[k]));
                if (MODE == 2) asm volatile("fma.rn.f32 %0, %0, %1, 0f3f800000;" : "+f"(f[i]) : "f"(g[j]));
                if (MODE == 3) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(g[j]));
                if (MODE == 4) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(g[j]));
                if (MODE == 5) asm volatile("mul.rn.f32 %0, %0, 0f3f8ccccd;" : "+f"(f[i]));
                if (MODE == 6) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(A[i]) : "l"(B[j]));
                if (MODE == 7) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(A[i]) : "l"(B[j]));
                if (MODE == 8) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(A[i]) : "l"(B[j]), "l"(C[k]));
                if (MODE == 9) asm volatile("lop3.b32 %0, %0, %1, %2, 0x6a;" : "+r"(u[i]) : "r"(v[j]), "r"(w[k]));
                if (MODE == 10) asm volatile("lop3.b32 %0, %0, %1, 0x80000000, 0x78;" : "+r"(u[i]) : "r"(v[j]));
                if (MODE == 11) asm volatile("xor.b32 %0, %0, %1;" : "+r"(u[i]) : "r"(v[j]));
                if (MODE == 12) asm volatile("min.f32 %0, %0, 0f424a2979;" : "+f"(f[i]));
                if (MODE == 13) asm volatile("min.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(g[j]));
                if (MODE == 14) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(f[i]));
                if (MODE == 15) asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(f[i]) : "f"(g[j]));
                if (MODE == 15) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(g[i]) : "f"(f[i]));
                if (MODE == 16) { unsigned long long t; asm volatile("ld.volatile.shared.b64 %0, [%1];" : "=l"(t) : "r"(sbase + 1024 * i) : "memory");
                                  asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(A[i]) : "l"(t)); }
                if (MODE == 17) asm volatile("st.shared.v2.f32 [%0], {%1, %2};" :: "r"(sbase + 1024 * i), "f"(f[i]), "f"(g[j]) : "memory");
                if (MODE == 18) asm volatile("mov.b32 %0, %1;" : "=r"(u[i]) : "r"(v[j]));
                if (MODE == 19) asm volatile("add.s32 %0, %0, %1;" : "+r"(u[i]) : "r"(v[j]));
                if (MODE == 20) asm volatile("prmt.b32 %0, %0, %1, 0x7632;" : "+r"(u[i]) : "r"(v[j]));
                if (MODE == 21) asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(u[i]) : "r"(v[j]), "r"(w[k]));
                if (MODE == 22) asm volatile("fma.rn.bf16x2 %0, %0, %1, %2;" : "+r"(u[i]) : "r"(v[j]), "r"(w[k]));
                if (MODE == 23) asm volatile("add.rn.f32 %0, %0, 0f3f8ccccd;" : "+f"(f[i]));
                if (MODE == 24) asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(f[i]));
                if (MODE == 25) asm volatile("fma.rn.f32x2 %0, %0, %0, %1;" : "+l"(A[i]) : "l"(B[j]));
                if (MODE == 26) asm volatile("copysign.f32 %0, %1, %0;" : "+f"(f[i]) : "f"(g[j]));
                if (MODE == 27) asm volatile("abs.f32 %0, %0;" : "+f"(f[i]));
                if (MODE == 28) asm volatile("mul.rn.f16x2 %0, %0, %1;" : "+r"(u[i]) : "r"(v[j]));
                if (MODE == 29) asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(f[i]));
                if (MODE == 30) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(u[i]));
                if (MODE == 31) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(f[i]));
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i)
        s += __uint_as_float((unsigned)(A[i] ^ B[i] ^ C[i]) ^ (unsigned)((A[i] ^ B[i] ^ C[i]) >> 32)) + f[i] + g[i] + h[i] + __uint_as_float(u[i] ^ v[i] ^ w[i]);
    s += sh[threadIdx.x].x;
    if (s == 123456.f) sink[0] = s;
}

static double g_bg = 0;

template <int MODE, int NX>
void run(const char* name, int sms)
{
    float *sink, *in; cudaMalloc(&sink, 4); cudaMalloc(&in, 4096); cudaMemset(in, 0x3f, 4096);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 2048, bps = 8, grid = sms * bps;
    double best = 0;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        probe<MODE, NX><<<grid, 128>>>(sink, in, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double rate = (double)grid * 4 * 8 * iters / (ms * 1e-3) / (sms * 4.0) / 1.965e9;
        if (rep && rate > best) best = rate;
    }
    const double clk = 1.0 / best;
    if (MODE == 0) g_bg = clk;
    printf("%-44s %6.2f clk per step   -> %5.2f clk per instruction\n", name, clk, MODE == 0 ? clk / 3 : (clk - g_bg) / NX);
    cudaFree(sink); cudaFree(in);
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    printf("%s, %d SMs, clock assumed 1.965 GHz, 8 warps per scheduler; background = 3 FFMA2 per step\n", p.name, sms);
    run<0, 0>("background: 3 x FFMA2 r,r,r", sms);
    run<1, 2>("FFMA r,r,r", sms);
    run<24, 2>("FFMA r,r,r (one register three times)", sms);
    run<2, 2>("FFMA r,r,imm", sms);
    run<3, 2>("FADD r,r", sms);
    run<23, 2>("FADD r,imm", sms);
    run<4, 2>("FMUL r,r", sms);
    run<5, 2>("FMUL r,imm", sms);
    run<6, 2>("FADD2", sms);
    run<7, 2>("FMUL2", sms);
    run<8, 2>("FFMA2", sms);
    run<25, 2>("FFMA2 a,a,b", sms);
    run<9, 2>("LOP3 r,r,r", sms);
    run<10, 2>("LOP3 r,r,imm", sms);
    run<12, 1>("FMNMX r,imm", sms);
    run<13, 1>("FMNMX r,r", sms);
    run<26, 2>("copysign (LOP3)", sms);
    run<14, 1>("MUFU.EX2", sms);
    run<29, 1>("MUFU.LG2", sms);
    run<15, 1>("MUFU.RCP + FADD", sms);
    run<31, 1>("MUFU.TANH", sms);
    run<30, 1>("ex2.f16x2", sms);
    run<16, 1>("LDS.64 + FADD2", sms);
    run<17, 1>("STS.64", sms);
    run<19, 2>("IADD", sms);
    run<20, 2>("PRMT", sms);
    run<21, 2>("HFMA2 r,r,r", sms);
    run<22, 2>("HFMA2.BF16 r,r,r", sms);
    run<28, 2>("HMUL2 r,r", sms);
    return 0;
}
