// tools/pipe_probe3.cu -- do instructions of DIFFERENT pipes overlap inside one scheduler, or do their issue costs add up?
// (not part of the product).  Every mode runs N independent instruction streams per warp; a "step" is one unrolled body.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/pipe_probe3 tools/pipe_probe3.cu
// Reading: if cycles(A+B) ~ cycles(A) + cycles(B) the two share one resource (issue port / register read ports);
// if ~ max(...) they overlap.
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ float ex2f(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lg2f(float x) { float y; asm volatile("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcpf(float x) { float y; asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

template <int MODE>
__global__ void __launch_bounds__(128) probe(float* sink, const float* in, int iters)
{
    __shared__ float2 sh[128 * 4];
    float2 a[8], b[8], c[8];
    unsigned u[8], v[8], w[8];
    float m[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        a[i] = make_float2(in[threadIdx.x + i], in[threadIdx.x + 32 + i]);
        b[i] = make_float2(in[threadIdx.x + 8 + i], in[threadIdx.x + 40 + i]);
        c[i] = make_float2(in[threadIdx.x + 64 + i], in[threadIdx.x + 72 + i]);
        u[i] = __float_as_uint(in[threadIdx.x + 16 + i]);
        v[i] = __float_as_uint(in[threadIdx.x + 24 + i]);
        w[i] = __float_as_uint(in[threadIdx.x + 48 + i]);
        m[i] = in[threadIdx.x + 56 + i];
    }
    for (int i = threadIdx.x; i < 512; i += 128) sh[i] = make_float2(1.f, 2.f);
    __syncthreads();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int j = (i + 1) & 7, k = (i + 3) & 7;
            constexpr bool FFMA2 = MODE == 0 || MODE == 2 || MODE == 4 || MODE == 6;
            constexpr bool LOP = MODE == 1 || MODE == 2 || MODE == 5 || MODE == 7;
            constexpr bool FMNMX = MODE == 3 || MODE == 4;
            constexpr bool FMUL2 = MODE == 5 || MODE == 8;
            constexpr bool MUFU = MODE == 6 || MODE == 7 || MODE == 9;
            if (FFMA2) a[i] = __ffma2_rn(a[i], b[j], c[k]);
            if (LOP) u[i] = (u[i] & v[j]) ^ w[k];
            if (FMNMX) m[i] = fminf(fabsf(m[i]), 50.5f) + 0.f * m[j];
            if (FMUL2) a[i] = __fmul2_rn(a[i], b[j]);
            if (MUFU) m[i] = ex2f(m[i]);
            if (MODE == 10 || MODE == 11) {
                // the instruction mix of the gather kernel per edge and PAIR of frames (SASS histogram of round 2):
                // 3.8 FFMA2, 3.2 FMUL2, 2.3 FADD2, 5 LOP3, 2 FMNMX, 1.9 LDS.64, 1.3 STS.64, and (MODE 11) 5 MUFU
                float2 L = sh[threadIdx.x + 128 * (i & 3)];
                float2 mu = __fadd2_rn(L, b[j]);
                u[i] = u[i] ^ __float_as_uint(mu.x) ^ v[j];
                float2 x;
                if (MODE == 11) { x.x = ex2f(-fminf(fabsf(mu.x), 50.5f)); x.y = ex2f(-fminf(fabsf(mu.y), 50.5f)); }
                else { x.x = fminf(fabsf(mu.x), 50.5f); x.y = fminf(fabsf(mu.y), 50.5f); }
                float2 fa = __ffma2_rn(c[k], x, a[j]);
                float2 fb = __ffma2_rn(a[j], x, c[k]);
                float2 A = __ffma2_rn(fa, a[k], __fmul2_rn(fb, b[j]));
                float2 B = __ffma2_rn(fa, b[j], __fmul2_rn(fb, a[k]));
                float2 P = __fmul2_rn(B, c[j]);
                float2 rp = P;
                if (MODE == 11 && (i & 1)) { rp.x = rcpf(P.x); rp.y = rcpf(P.y); }
                float2 R = __fmul2_rn(A, __fmul2_rn(rp, c[i]));
                float2 mag = R;
                if (MODE == 11) { mag.x = lg2f(R.x); mag.y = lg2f(R.y); }
                unsigned s1 = (u[i] ^ __float_as_uint(mu.x)) & 0x80000000u, s2 = (v[i] ^ __float_as_uint(mu.y)) & 0x80000000u;
                mag.x = __uint_as_float(__float_as_uint(mag.x) | s1);
                mag.y = __uint_as_float(__float_as_uint(mag.y) | s2);
                sh[threadIdx.x + 128 * ((i + 1) & 3)] = mag;
                a[i] = __fadd2_rn(mag, a[i]);
                v[i] ^= __float_as_uint(mu.y);
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i].x + a[i].y + b[i].x + c[i].y + __uint_as_float(u[i] ^ v[i] ^ w[i]) + m[i];
    s += sh[threadIdx.x].x;
    if (s == 123456.f) sink[0] = s;
}

template <int MODE>
void run(const char* name, int sms, int bps)
{
    float *sink, *in; cudaMalloc(&sink, 4); cudaMalloc(&in, 4096); cudaMemset(in, 0x3f, 4096);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 2048, grid = sms * bps;
    double best = 0;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        probe<MODE><<<grid, 128>>>(sink, in, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double rate = (double)grid * 4 /*warps*/ * 8 /*steps*/ * iters / (ms * 1e-3) / (sms * 4.0) / 1.965e9;
        if (rep && rate > best) best = rate;
    }
    printf("%-56s warps/scheduler %2d : %6.2f clk per step\n", name, bps, 1.0 / best);
    cudaFree(sink); cudaFree(in);
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    printf("%s, %d SMs, clock assumed 1.965 GHz; one step = one instruction of each kind named\n", p.name, sms);
    for (int bps : {8, 3}) {
        run<0>("FFMA2 (3 register pairs)", sms, bps);
        run<1>("LOP3 (3 registers)", sms, bps);
        run<2>("FFMA2 + LOP3", sms, bps);
        run<3>("FMNMX |r|,imm (+ FFMA)", sms, bps);
        run<4>("FFMA2 + FMNMX (+ FFMA)", sms, bps);
        run<8>("FMUL2", sms, bps);
        run<5>("FMUL2 + LOP3", sms, bps);
        run<9>("MUFU.EX2", sms, bps);
        run<6>("MUFU.EX2 + FFMA2", sms, bps);
        run<7>("MUFU.EX2 + LOP3", sms, bps);
        run<10>("gather-kernel mix per edge pair, no MUFU", sms, bps);
        run<11>("gather-kernel mix per edge pair, 5 MUFU", sms, bps);
    }
    return 0;
}
