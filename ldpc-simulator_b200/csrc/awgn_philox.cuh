// awgn_philox.cuh -- counter-based BPSK-AWGN channel for the Monte-Carlo path.
//
// Replaces Channel.process mode 1 / modulation 1 (python_ldpc_app/channel.py:38-81):
//   symbol = -1 for bit 0, +1 for bit 1                         (:49)
//   noise  = N(0,1) * sigma^2   (the reference passes sigma**2 as the standard
//            deviation, :68) or N(0,1) * sigma with the quirk switched off
//   y      = symbol + noise                                      (:76)
//   LLR    = 2 y / sigma^2                                       (:80)
// The reference draws from a clock-seeded numpy RandomState (:30) and is not
// reproducible; here every value is a pure function of
// (seed, stream_id, frame, variable) through Philox4x32-10, so ranks and
// launches draw disjoint, replayable streams and the same LLRs can be
// regenerated for the oracle (ldpc_channel_llr).
//
// All arithmetic uses explicit _rn intrinsics so that the standalone channel
// kernel and the copy inlined into the resident decoder produce identical bits.
#pragma once
#include "qc_device.cuh"

namespace ldpc {

struct Philox4 { uint32_t x, y, z, w; };

__device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                 uint32_t k0, uint32_t k1)
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        c0 = hi1 ^ c1 ^ k0;
        c1 = lo1;
        c2 = hi0 ^ c3 ^ k1;
        c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return Philox4{c0, c1, c2, c3};
}

// Two independent unit normals from two 32-bit words (Box-Muller on the MUFU pipe).
// u1 in (0,1]: tail reaches sqrt(2*33*ln2) = 6.76 sigma.
__device__ __forceinline__ void box_muller(uint32_t a, uint32_t b, float& g0, float& g1)
{
    const float u1 = __fmaf_rn(__uint2float_rn(a), 2.3283064365386963e-10f, 1.1641532182693481e-10f);
    const float u2 = __fmaf_rn(__uint2float_rn(b), 2.3283064365386963e-10f, -0.5f);   // [-0.5, 0.5]
    // r = sqrt(-2 ln u1) = sqrt(-2 ln2 * lg2 u1)
    const float r = __fsqrt_rn(__fmul_rn(-1.3862943611198906f, __log2f(u1)));
    const float ang = __fmul_rn(6.283185307179586f, u2);                               // [-pi, pi]
    g0 = __fmul_rn(r, __cosf(ang));
    g1 = __fmul_rn(r, __sinf(ang));
}

struct ChannelConst {
    float noise_dev;   // sigma^2 (quirk) or sigma
    float llr_scale;   // 2 / sigma^2
    float amp;         // symbol amplitude: 1 (modulation 1, :49) or 0.7 (modulation 2, :51)
    // interference (channel modes 2 and 3, :83-100); a2 = 0 and hit_threshold = 0 in mode 1
    float a2;          // deviation of the second noise term when the bit is hit
    float l_hit;       // LLR scale of a hit bit (L_c2 / L_c3); llr_scale is the scale of the others (L_c1)
    uint32_t hit_threshold;   // a bit is hit iff a uniform 32-bit draw is below this (0xffffffff = always)
    uint32_t k0, k1;   // Philox key = seed
    uint32_t stream_id;
};

// LLRs of variables 4q .. 4q+3 of one frame.  bits4 packs the four transmitted
// bits in its low nibble (bit i = variable 4q+i).
__device__ __forceinline__ void channel_llr4(const ChannelConst& cc, uint64_t frame, uint32_t q,
                                             uint32_t bits4, float out[4])
{
    const Philox4 p = philox4x32_10(q, (uint32_t)frame, (uint32_t)(frame >> 32), cc.stream_id, cc.k0, cc.k1);
    float g[4];
    box_muller(p.x, p.y, g[0], g[1]);
    box_muller(p.z, p.w, g[2], g[3]);
    if (cc.hit_threshold == 0u) {                          // mode 1: AWGN only
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float sym = ((bits4 >> i) & 1u) ? cc.amp : -cc.amp;
            const float y = __fmaf_rn(cc.noise_dev, g[i], sym);
            out[i] = __fmul_rn(y, cc.llr_scale);
        }
        return;
    }
    // modes 2 / 3: a second Gaussian and the hit decision from two more Philox blocks of the same
    // (frame, stream): word index with bit 31 / bit 30 set (n < 2^30 variables)
    const Philox4 p2 = philox4x32_10(q | 0x80000000u, (uint32_t)frame, (uint32_t)(frame >> 32), cc.stream_id, cc.k0, cc.k1);
    const Philox4 pu = philox4x32_10(q | 0x40000000u, (uint32_t)frame, (uint32_t)(frame >> 32), cc.stream_id, cc.k0, cc.k1);
    float g2[4];
    box_muller(p2.x, p2.y, g2[0], g2[1]);
    box_muller(p2.z, p2.w, g2[2], g2[3]);
    const uint32_t u[4] = {pu.x, pu.y, pu.z, pu.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float sym = ((bits4 >> i) & 1u) ? cc.amp : -cc.amp;
        const bool hit = cc.hit_threshold == 0xffffffffu || u[i] < cc.hit_threshold;
        float y = __fmaf_rn(cc.noise_dev, g[i], sym);
        if (hit) y = __fmaf_rn(cc.a2, g2[i], y);
        out[i] = __fmul_rn(y, hit ? cc.l_hit : cc.llr_scale);
    }
}

}  // namespace ldpc
