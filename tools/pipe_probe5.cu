// tools/pipe_probe5.cu -- the floor of the resident gather kernel's instruction mix (not part of the product).
// One step = one check row of 6 edges for a PAIR of frames with exactly the hot-loop opcode counts of
// k_qc_gather<192, 2, false, WiMAX-2304 r1/2> (profiles/r2_sass_gather_wimax2304.txt, per edge pair: 3.79 FFMA2,
// 3.16 FMUL2, 2.32 FADD2, 5.05 LOP3 of which 1 with three registers, 2 FMNMX, 2 EX2 + 2 LG2 + 1.05 RCP, 1.87 LDS.64,
// 1.32 STS.64), as independent instruction streams (volatile inline PTX): no barriers, no dependent chains longer than the
// unrolled body, no tensor-memory traffic, no prologue.  What this runs at is what the mix costs the scheduler.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/pipe_probe5 tools/pipe_probe5.cu
#include <cstdio>
#include <cuda_runtime.h>

template <bool MUFU>
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
#define FFMA2(i, j, k) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(A[(i) & 7]) : "l"(B[(j) & 7]), "l"(C[(k) & 7]));
#define FMUL2(i, j) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(B[(i) & 7]) : "l"(C[(j) & 7]));
#define FADD2(i, j) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(C[(i) & 7]) : "l"(A[(j) & 7]));
#define LOP3R(i, j, k) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[(i) & 7]) : "r"(v[(j) & 7]), "r"(w[(k) & 7]));
#define LOP3I(i, j) asm volatile("lop3.b32 %0, %0, %1, 0x80000000, 0x78;" : "+r"(v[(i) & 7]) : "r"(w[(j) & 7]));
#define FMNMX(i) asm volatile("min.f32 %0, %0, 0f424a2979;" : "+f"(f[(i) & 7]));
#define EX2(i) if (MUFU) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(f[(i) & 7]));
#define LG2(i) if (MUFU) asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(g[(i) & 7]));
#define RCP(i) if (MUFU) asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(h[(i) & 7]) : "f"(g[(i + 1) & 7])); \
               if (MUFU) asm volatile("lop3.b32 %0, %0, %1, 0x80000000, 0x78;" : "+r"(w[(i) & 7]) : "r"(__float_as_uint(h[(i) & 7])));
#define LDS64(i) { unsigned long long t; asm volatile("ld.volatile.shared.b64 %0, [%1];" : "=l"(t) : "r"(sbase + 1024 * ((i) & 7)) : "memory"); \
                   asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(C[(i) & 7]) : "l"(t)); }
#define STS64(i) asm volatile("st.volatile.shared.b64 [%0], %1;" :: "r"(sbase + 1024 * ((i) & 7)), "l"(A[(i) & 7]) : "memory");
    for (int it = 0; it < iters; ++it) {
        // 6 edges x 2 frames.  The FADD2 that consumes each LDS.64 is one of the 14 FADD2 of the row (11 + 3 more).
#pragma unroll
        for (int e = 0; e < 6; ++e) {
            LDS64(e) FMNMX(e) FMNMX(e + 3) EX2(e) EX2(e + 3) LOP3R(e, e + 1, e + 2)
            FFMA2(e, e + 1, e + 2) FFMA2(e + 3, e + 4, e + 5) FFMA2(e + 5, e + 2, e + 7) FFMA2(e + 6, e + 1, e + 3)
            FMUL2(e, e + 1) FMUL2(e + 2, e + 5) FMUL2(e + 4, e + 3)
            LG2(e) LG2(e + 4) LOP3I(e, e + 1) LOP3I(e + 2, e + 3) LOP3I(e + 4, e + 5)
            STS64(e)
            if (e < 5) { LDS64(e + 4) }                 // 11 loads per row
            if (e & 1) { RCP(e) RCP(e + 1) }            // 6 reciprocals per row
            if (e < 3) { FADD2(e, e + 3) }              // 14 packed additions per row
            if (e < 2) { STS64(e + 5) }                 // 8 stores per row
            if (e == 5) { FMUL2(e + 1, e + 6) FFMA2(e + 2, e + 5, e + 1) }    // 19 packed multiplies, 23 FFMA2 per row ... (24)
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i)
        s += __uint_as_float((unsigned)(A[i] ^ B[i] ^ C[i]) ^ (unsigned)((A[i] ^ B[i] ^ C[i]) >> 32)) + f[i] + g[i] + h[i] + __uint_as_float(u[i] ^ v[i] ^ w[i]);
    s += sh[threadIdx.x].x;
    if (s == 123456.f) sink[0] = s;
}

template <bool MUFU>
void run(const char* name, int sms, int bps)
{
    float *sink, *in; cudaMalloc(&sink, 4); cudaMalloc(&in, 4096); cudaMemset(in, 0x3f, 4096);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 8192, grid = sms * bps;
    double best = 0;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        probe<MUFU><<<grid, 128>>>(sink, in, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double rate = (double)grid * 4 * iters / (ms * 1e-3) / (sms * 4.0) / 1.965e9;     // rows per clock per scheduler
        if (rep && rate > best) best = rate;
    }
    const double clk_edge = 1.0 / best / 6.0;
    // 7296 edges x 20 passes x 65536 frame pairs per launch on 148 x 4 schedulers
    const double ms_launch = clk_edge * 7296.0 * 20 * 65536 / 32.0 / (sms * 4.0) / 1.965e9 * 1e3;     // a warp instruction covers 32 check rows
    printf("%-40s warps/scheduler %2d : %6.2f clk per edge and frame pair  -> %6.2f ms per 131072-frame launch, %5.2f Gbit/s\n", name, bps,
           clk_edge, ms_launch, 131072.0 * 1152 / (ms_launch * 1e-3) / 1e9);
    cudaFree(sink); cudaFree(in);
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    printf("%s, %d SMs, clock assumed 1.965 GHz\n", p.name, sms);
    for (int bps : {3, 6, 8, 12}) {
        run<true>("gather-kernel mix", sms, bps);
        run<false>("gather-kernel mix without its MUFU", sms, bps);
    }
    return 0;
}
