// qc_device.cuh -- plain types shared by the host code and the resident kernels.
//
// Device-safe: no host headers, so that the same text compiles under nvcc (static registry,
// spa_qc_spec.cu) and under NVRTC (run-time specialisation, qc_jit.cu).
#pragma once
#ifdef __CUDACC_RTC__
typedef unsigned char uint8_t;
typedef short int16_t;
typedef int int32_t;
typedef unsigned int uint32_t;
typedef long long int64_t;
typedef unsigned long long uint64_t;
typedef unsigned long long uintptr_t;
#else
#include <stdint.h>
#endif

namespace ldpc {

// ---- resident quasi-cyclic path, spa_qc_resident.cu / qc_kernel.cuh ------
struct McParams {               // in-kernel channel (awgn_philox.cuh); enabled when active
    bool active = false;
    float noise_dev = 1.f;      // sigma^2 (reference quirk, channel.py:68) or sigma
    float llr_scale = 2.f;      // 2 / sigma^2 (channel.py:80)
    float amp = 1.f;            // symbol amplitude (channel.py:49,51)
    float a2 = 0.f;             // channel modes 2 / 3 (channel.py:83-100), see ChannelConst
    float l_hit = 0.f;
    uint32_t hit_threshold = 0;
    uint64_t seed = 0;
    uint32_t stream_id = 0;
    uint64_t frame_offset = 0;
    const uint8_t* codeword = nullptr;       // device, [n] (or [frames][n]) or null = all-zero
    long long codeword_stride = 0;           // bytes between the codewords of consecutive frames, 0 = one for all
    const uint8_t* info_mask = nullptr;      // device, [n] or null = first k_info positions
    int k_info = 0;
    unsigned long long* counters = nullptr;  // device uint64[5]
};

namespace qc {
// output pointers of the specialised resident kernels (any may be null)
struct Outputs {
    uint8_t* z;          // [F][n] or null
    uint32_t* zbits;     // [F][ceil(n/32)] or null
    int32_t* conv_it;    // [F] or null
    uint8_t* ok;         // [F] or null
    float* post;         // [F][n] or null
};
}  // namespace qc

}  // namespace ldpc
