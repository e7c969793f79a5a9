// ldpc_common.cuh -- shared declarations of the B200 LDPC decode library.
//
// Internal header: the public surface is include/ldpc_b200.h.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <string.h>
#include <vector>
#include <atomic>

#include "../../include/ldpc_b200.h"
#include "qc_device.cuh"

namespace ldpc {

// ---- error plumbing -------------------------------------------------------
void set_error(const char* fmt, ...);
extern std::atomic<uint64_t> g_launches;   // kernels launched by this library

#define LDPC_CUDA_TRY(expr)                                                           \
    do {                                                                              \
        cudaError_t _e = (expr);                                                      \
        if (_e != cudaSuccess) {                                                      \
            ::ldpc::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                              __FILE__, __LINE__);                                    \
            return LDPC_ERR_CUDA;                                                     \
        }                                                                             \
    } while (0)

#define LDPC_LAUNCH_CHECK()                                                           \
    do {                                                                              \
        ::ldpc::g_launches.fetch_add(1, std::memory_order_relaxed);                   \
        cudaError_t _e = cudaGetLastError();                                          \
        if (_e != cudaSuccess) {                                                      \
            ::ldpc::set_error("kernel launch failed: %s (%s:%d)",                     \
                              cudaGetErrorString(_e), __FILE__, __LINE__);            \
            return LDPC_ERR_CUDA;                                                     \
        }                                                                             \
    } while (0)

// ---- device properties (cached per process; one device per process) --------
struct DeviceInfo {
    int device = -1;
    int sm_count = 0;
    int max_smem_optin = 0;
    bool ok = false;
};
int get_device_info(DeviceInfo* out);   // returns LDPC_OK / LDPC_ERR_CUDA

// ---- quasi-cyclic description ----------------------------------------------
struct QcInfo {
    int z = 0, mb = 0, nb = 0;
    std::vector<int16_t> shift;          // mb*nb, -1 = zero block
};

}  // namespace ldpc

// The opaque handle of the C ABI.
struct ldpc_graph {
    int m = 0, n = 0;
    int64_t nnz = 0;
    int max_cdeg = 0, max_vdeg = 0;
    // host copies
    std::vector<int32_t> row_ptr, col_idx, col_ptr, csc_edge, edge_row;
    // device tables (edge numbering = CSR order)
    int32_t* d_row_ptr = nullptr;   // [m+1]
    int32_t* d_col_idx = nullptr;   // [nnz]  column of each edge
    int32_t* d_col_ptr = nullptr;   // [n+1]
    int32_t* d_csc_edge = nullptr;  // [nnz]  edges of each column, ascending row
    bool is_qc = false;
    ldpc::QcInfo qc;
    // resident-kernel tables (built lazily by spa_qc_resident.cu)
    void* d_qc_tables = nullptr;
    int device = -1;
    uint64_t serial = 0;            // unique per handle (keys of caches that must not alias a freed pointer)
    // kernel family per (TABLE_KERNEL, NO_JIT) flag combination, -1 = not determined yet (qc_resident_kind)
    mutable std::atomic<int> kind_cache[4] = {{-1}, {-1}, {-1}, {-1}};
};

namespace ldpc {

// ---- generic (any graph) path, spa_generic.cu ------------------------------
size_t generic_workspace_bytes(const ldpc_graph* g, int64_t frames, int dtype);
int generic_decode(const ldpc_graph* g, int dtype, int64_t frames, int max_iter, unsigned flags,
                   const void* llr_dev, uint8_t* z_dev, int32_t* conv_dev, uint8_t* ok_dev,
                   void* post_dev, float* norm_dev, int k_info, void* ws, size_t ws_bytes,
                   cudaStream_t stream);

// ---- resident quasi-cyclic path, spa_qc_resident.cu ------------------------
bool qc_resident_supported(const ldpc_graph* g);   // by the table-driven kernel
int qc_resident_decode(const ldpc_graph* g, int64_t frames, int max_iter, unsigned flags,
                       const float* llr_dev, uint8_t* z_dev, uint32_t* zbits_dev, int32_t* conv_dev,
                       uint8_t* ok_dev, float* post_dev, const McParams& mc, void* ws, size_t ws_bytes,
                       cudaStream_t stream);
void qc_resident_release(ldpc_graph* g);
// compile-time specialised kernels for registered base matrices, spa_qc_spec.cu
int qc_spec_find(const ldpc_graph* g);   // registry index or -1
int qc_spec_decode(int idx, const ldpc_graph* g, int64_t frames, int max_iter, unsigned flags,
                   const float* llr_dev, uint8_t* z_dev, uint32_t* zbits_dev, int32_t* conv_dev,
                   uint8_t* ok_dev, float* post_dev, const McParams& mc, void* ws, cudaStream_t stream);

// run-time specialised kernels for any other base matrix (NVRTC), qc_jit.cu
bool qc_jit_supported(const ldpc_graph* g);
int qc_jit_prepare(const ldpc_graph* g);
int qc_jit_decode(const ldpc_graph* g, int64_t frames, int max_iter, unsigned flags,
                  const float* llr_dev, uint8_t* z_dev, uint32_t* zbits_dev, int32_t* conv_dev,
                  uint8_t* ok_dev, float* post_dev, const McParams& mc, void* ws, cudaStream_t stream);
// kernel family of the resident path for (g, flags): an ldpc_kernel_kind, LDPC_KERNEL_GENERIC if none
int qc_resident_kind(const ldpc_graph* g, unsigned flags);

// ---- channel / counters, mc.cu ---------------------------------------------
int channel_fill(int n, int dtype, int64_t frames, const ldpc_channel& ch, uint64_t seed,
                 uint32_t stream_id, uint64_t frame_offset, const uint8_t* codeword_dev, int64_t codeword_stride,
                 void* llr_dev, cudaStream_t stream);
int channel_params(const ldpc_channel& ch, int n, uint64_t seed, uint32_t stream_id, McParams* mc);
int count_errors(int n, int k_info, int64_t frames, const uint8_t* z_dev, const uint8_t* ok_dev,
                 const int32_t* conv_dev, const uint8_t* codeword_dev, int64_t codeword_stride,
                 const uint8_t* info_mask_dev, const float* norm_dev, int k_norm,
                 unsigned long long* counters_dev, cudaStream_t stream);

}  // namespace ldpc
