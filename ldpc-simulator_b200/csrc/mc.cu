// mc.cu -- device-side channel fill and error-counter fold for the Monte-Carlo path.
//
// Replaces the per-frame bookkeeping of run_simulation / process_block
// (python_ldpc_app/main.py:295-339, 43-146): frame error = final syndrome != 0;
// info-bit errors are counted ONLY in failed frames, on the un-complemented
// decoder output (main.py:323-330); the convergence iteration is summed over
// converged frames (:336-339).
#include "ldpc_common.cuh"
#include "awgn_philox.cuh"

#include <cmath>

namespace ldpc {

namespace {

template <typename T>
__global__ void __launch_bounds__(256)
k_channel_fill(int n, int64_t frames, ChannelConst cc, uint64_t frame_offset,
               const uint8_t* __restrict__ codeword_base, long long cw_stride, T* __restrict__ llr)
{
    const int quads = (n + 3) / 4;
    const int64_t items = frames * quads;
    for (int64_t id = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; id < items;
         id += (int64_t)gridDim.x * blockDim.x) {
        const int64_t f = id / quads;
        const uint32_t q = (uint32_t)(id - f * quads);
        uint32_t bits = 0;
        if (codeword_base) {
            const uint8_t* codeword = codeword_base + f * cw_stride;
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if ((int)(4 * q + i) < n && codeword[4 * q + i]) bits |= 1u << i;
        }
        float v[4];
        channel_llr4(cc, frame_offset + (uint64_t)f, q, bits, v);
#pragma unroll
        for (int i = 0; i < 4; ++i)
            if ((int)(4 * q + i) < n) llr[f * n + 4 * q + i] = (T)v[i];
    }
}

// One warp per frame.
__global__ void __launch_bounds__(256)
k_count_errors(int n, int k_info, int64_t frames, const uint8_t* __restrict__ z,
               const uint8_t* __restrict__ ok, const int32_t* __restrict__ conv,
               const uint8_t* __restrict__ codeword_base, long long cw_stride,
               const uint8_t* __restrict__ info_mask, const float* __restrict__ norm, int k_norm,
               unsigned long long* __restrict__ counters)
{
    __shared__ unsigned long long acc[6];
    if (threadIdx.x < 6) acc[threadIdx.x] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    unsigned long long c_fr = 0, c_fe = 0, c_be = 0, c_cs = 0, c_cc = 0, c_nl = 0;
    for (int64_t f = warp0; f < frames; f += nwarps) {
        const bool good = ok[f] != 0;
        if (!good) {                                                     // main.py:326
            const uint8_t* codeword = codeword_base ? codeword_base + f * cw_stride : nullptr;
            int errs = 0;
            const int span = info_mask ? n : k_info;
            for (int j = lane; j < span; j += 32) {
                if (info_mask && !info_mask[j]) continue;
                const unsigned est = z[f * n + j] ^ 1u;                  // :328
                const unsigned sent = codeword ? codeword[j] : 0u;
                errs += (est != sent);
            }
#pragma unroll
            for (int o = 16; o; o >>= 1) errs += __shfl_xor_sync(0xffffffffu, errs, o);
            if (lane == 0) { c_be += (unsigned)errs; c_fe += 1; }
        }
        if (lane == 0) {
            c_fr += 1;
            const int ci = conv[f];
            if (ci >= 0) { c_cs += (unsigned)ci; c_cc += 1; }            // :336-339
            // :218-228,332-334: sign changes of the exit pass over the first k_norm bits, every frame
            if (norm) c_nl += (unsigned long long)__float2ll_rn(norm[f] * (float)k_norm);
        }
    }
    if (lane == 0) {
        atomicAdd(&acc[0], c_fr); atomicAdd(&acc[1], c_fe); atomicAdd(&acc[2], c_be);
        atomicAdd(&acc[3], c_cs); atomicAdd(&acc[4], c_cc);
        if (norm) atomicAdd(&acc[5], c_nl);
    }
    __syncthreads();
    if (threadIdx.x < (norm ? 6 : 5) && acc[threadIdx.x]) atomicAdd(&counters[threadIdx.x], acc[threadIdx.x]);
}

}  // namespace

// Channel.create_channel (channel.py:102-125) -> the constants the device generator needs.
//   mode 1: y = sym + N(0, sigma^2 or sigma^4) ; LLR = 2 y / sigma^2                      (:56-81)
//   mode 2: every bit is hit with probability P(a/n < p), a uniform in 0..n-1 (:86-88):
//           hit:  (sym + s2 g2 + s1 g1) L_c2 ;  else (sym + s1 g1) L_c1                     (:89-96)
//   mode 3: ((sym + s1 g1 + s2 g2) p + (sym + s1 g1)(1 - p)) L_c3 = (sym + s1 g1 + p s2 g2) L_c3   (:98-100)
// with s1, s2, L_c1..3 of :104-121 (modes 2/3 use sigma itself as the deviation, no sigma^2 quirk).
static int make_channel_const(const ldpc_channel& ch, int n, uint64_t seed, uint32_t stream_id, ChannelConst* out)
{
    if (!(ch.speed > 0.0)) { set_error("speed must be positive"); return LDPC_ERR_INVALID; }
    if (ch.mode < 1 || ch.mode > 3) { set_error("channel mode %d (1, 2 or 3)", ch.mode); return LDPC_ERR_INVALID; }
    if (ch.mode != 1 && !(ch.p > 0.0 && ch.p <= 1.0)) { set_error("interference share p must be in (0, 1]"); return LDPC_ERR_INVALID; }
    const double lin1 = std::pow(10.0, ch.snr_db * 0.1), lin2 = std::pow(10.0, ch.interference_snr_db * 0.1);
    const double sigma = 1.0 / std::sqrt(2.0 * ch.speed * lin1);                          // channel.py:113
    ChannelConst cc;
    cc.amp = ch.modulation == 2 ? 0.7f : 1.0f;                                            // channel.py:49,51
    cc.a2 = 0.f; cc.l_hit = 0.f; cc.hit_threshold = 0u;
    if (ch.mode == 1) {
        cc.noise_dev = (float)(ch.sigma_sq_quirk ? sigma * sigma : sigma);                // channel.py:68
        cc.llr_scale = (float)(2.0 / (sigma * sigma));                                    // channel.py:80
    } else if (ch.mode == 2) {
        cc.noise_dev = (float)sigma;
        cc.llr_scale = (float)(4.0 * ch.speed * lin1);                                                     // L_c1
        cc.a2 = (float)(1.0 / std::sqrt(2.0 * ch.speed * lin2 * ch.p));                                    // sigma2
        cc.l_hit = (float)(4.0 * ch.speed / (1.0 / lin1 + 1.0 / (lin2 * ch.p)));                           // L_c2
        const double hit = std::min(1.0, std::ceil(ch.p * n - 1e-12) / (double)n);                         // P(a/n < p)
        cc.hit_threshold = hit >= 1.0 ? 0xffffffffu : (uint32_t)std::max(1.0, hit * 4294967296.0);
    } else {
        cc.noise_dev = (float)sigma;
        cc.llr_scale = 0.f;
        cc.a2 = (float)(ch.p / std::sqrt(2.0 * ch.speed * lin2));                                          // p * sigma2
        cc.l_hit = (float)(4.0 * ch.p * ch.speed / (2.0 / lin2) + 4.0 * ch.speed * (1.0 - ch.p) * lin1);   // L_c3
        cc.hit_threshold = 0xffffffffu;
    }
    cc.k0 = (uint32_t)seed;
    cc.k1 = (uint32_t)(seed >> 32);
    cc.stream_id = stream_id;
    *out = cc;
    return LDPC_OK;
}

int channel_fill(int n, int dtype, int64_t frames, const ldpc_channel& ch, uint64_t seed,
                 uint32_t stream_id, uint64_t frame_offset, const uint8_t* codeword_dev, int64_t codeword_stride,
                 void* llr_dev, cudaStream_t stream)
{
    if (n <= 0 || frames < 0 || !llr_dev) { set_error("bad channel arguments"); return LDPC_ERR_INVALID; }
    ChannelConst cc;
    int rc = make_channel_const(ch, n, seed, stream_id, &cc);
    if (rc) return rc;
    if (frames == 0) return LDPC_OK;
    DeviceInfo di;
    rc = get_device_info(&di);
    if (rc) return rc;
    const int64_t items = frames * ((n + 3) / 4);
    const int grid = (int)std::min<int64_t>((items + 255) / 256, (int64_t)di.sm_count * 16);
    if (dtype == LDPC_F64)
        k_channel_fill<double><<<grid, 256, 0, stream>>>(n, frames, cc, frame_offset, codeword_dev, codeword_stride, (double*)llr_dev);
    else
        k_channel_fill<float><<<grid, 256, 0, stream>>>(n, frames, cc, frame_offset, codeword_dev, codeword_stride, (float*)llr_dev);
    LDPC_LAUNCH_CHECK();
    return LDPC_OK;
}

int channel_params(const ldpc_channel& ch, int n, uint64_t seed, uint32_t stream_id, McParams* mc)
{
    ChannelConst cc;
    const int rc = make_channel_const(ch, n, seed, stream_id, &cc);
    if (rc) return rc;
    mc->noise_dev = cc.noise_dev;
    mc->llr_scale = cc.llr_scale;
    mc->amp = cc.amp;
    mc->a2 = cc.a2;
    mc->l_hit = cc.l_hit;
    mc->hit_threshold = cc.hit_threshold;
    mc->seed = seed;
    mc->stream_id = stream_id;
    return LDPC_OK;
}

int count_errors(int n, int k_info, int64_t frames, const uint8_t* z_dev, const uint8_t* ok_dev,
                 const int32_t* conv_dev, const uint8_t* codeword_dev, int64_t codeword_stride,
                 const uint8_t* info_mask_dev, const float* norm_dev, int k_norm,
                 unsigned long long* counters_dev, cudaStream_t stream)
{
    if (frames == 0) return LDPC_OK;
    DeviceInfo di;
    int rc = get_device_info(&di);
    if (rc) return rc;
    const int grid = (int)std::min<int64_t>((frames * 32 + 255) / 256, (int64_t)di.sm_count * 8);
    k_count_errors<<<grid, 256, 0, stream>>>(n, k_info, frames, z_dev, ok_dev, conv_dev, codeword_dev, codeword_stride, info_mask_dev, norm_dev, k_norm, counters_dev);
    LDPC_LAUNCH_CHECK();
    return LDPC_OK;
}

}  // namespace ldpc
