// tools/f32x2_probe.cu -- are the packed fp32 instructions of sm_100a (FFMA2 / FADD2 / FMUL2) bit-identical, lane by
// lane, to the scalar round-to-nearest FFMA / FADD / FMUL?  (not part of the product)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/f32x2_probe tools/f32x2_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t rng(uint32_t& s) { s = s * 1664525u + 1013904223u; return s; }

__global__ void probe(unsigned long long* bad, int mode, int iters)
{
    uint32_t s = blockIdx.x * 1024u + threadIdx.x + 12345u * (mode + 1);
    unsigned long long nf = 0, na = 0, nm = 0, nn = 0;
    for (int i = 0; i < iters; ++i) {
        float v[6];
#pragma unroll
        for (int q = 0; q < 6; ++q) {
            uint32_t b = rng(s) ^ (rng(s) << 13);
            if (mode == 0) {            // moderate magnitudes: exponent 100..150
                b = (b & 0x807fffffu) | ((100u + (rng(s) >> 8) % 50u) << 23);
            } else if (mode == 1) {     // anything, including denormals / inf / nan
            } else {                    // small values near the denormal range
                b = (b & 0x807fffffu) | (((rng(s) >> 8) % 40u) << 23);
            }
            v[q] = __uint_as_float(b);
        }
        const float2 a = make_float2(v[0], v[1]), b2 = make_float2(v[2], v[3]), c = make_float2(v[4], v[5]);
        const float2 f = __ffma2_rn(a, b2, c);
        const float2 ad = __fadd2_rn(a, b2);
        const float2 an = __fadd2_rn(a, make_float2(-b2.x, -b2.y));
        const float2 m = __fmul2_rn(a, b2);
        auto same = [](float x, float y) { return __float_as_uint(x) == __float_as_uint(y) || (x != x && y != y); };
        nf += !same(f.x, __fmaf_rn(a.x, b2.x, c.x)) + !same(f.y, __fmaf_rn(a.y, b2.y, c.y));
        na += !same(ad.x, __fadd_rn(a.x, b2.x)) + !same(ad.y, __fadd_rn(a.y, b2.y));
        nn += !same(an.x, __fadd_rn(a.x, -b2.x)) + !same(an.y, __fadd_rn(a.y, -b2.y));
        nm += !same(m.x, __fmul_rn(a.x, b2.x)) + !same(m.y, __fmul_rn(a.y, b2.y));
    }
    atomicAdd(&bad[0], nf); atomicAdd(&bad[1], na); atomicAdd(&bad[2], nn); atomicAdd(&bad[3], nm);
}

int main()
{
    unsigned long long* bad; cudaMalloc(&bad, 32);
    const char* names[3] = {"moderate magnitudes", "arbitrary bit patterns", "near the denormal range"};
    for (int mode = 0; mode < 3; ++mode) {
        cudaMemset(bad, 0, 32);
        probe<<<256, 256>>>(bad, mode, 4096);
        unsigned long long h[4]; cudaMemcpy(h, bad, 32, cudaMemcpyDeviceToHost);
        printf("%-26s lanes tested %llu: mismatches FFMA2 %llu  FADD2 %llu  FADD2(neg) %llu  FMUL2 %llu\n", names[mode],
               2ull * 256 * 256 * 4096, h[0], h[1], h[2], h[3]);
    }
    return 0;
}
