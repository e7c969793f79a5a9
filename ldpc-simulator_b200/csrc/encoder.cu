// encoder.cu -- systematic encoder on the device, for Monte-Carlo runs with a fresh random codeword per frame.
//
// Replaces Generator.generate_bit_sequence + DataBuffer.encode (python_ldpc_app/generator.py:7-9,
// data_buffer.py:47-82) with G = [I_k | A^T] built from H_std = [A | I_m]
// (encoder_decoder_data.py:319-344): codeword (in H_std column order) = [u | A u mod 2].
// A is kept bit-packed ([m][ceil(k/64)] words, L2 resident); one warp computes one parity bit per lane
// step as popc(A_row & u) over the packed info word.  An optional output permutation writes the
// codeword in another column order (the raw ALIST order: position perm[j] = H_std column j).
#include "ldpc_common.cuh"
#include "awgn_philox.cuh"

struct ldpc_encoder {
    int m = 0, n = 0, k = 0, kw = 0;          // kw = 64-bit words per packed info vector
    uint64_t* d_a = nullptr;                  // [m][kw]
    int32_t* d_pos = nullptr;                 // [n] output position of H_std column j (identity if null)
};

namespace ldpc {
namespace {

// one CTA per frame: pack the info bits into shared memory, then every warp sweeps parity rows
__global__ void __launch_bounds__(256)
k_encode(int m, int n, int k, int kw, const uint64_t* __restrict__ a_bits, const int32_t* __restrict__ pos,
         int64_t frames, const uint8_t* __restrict__ data_in, uint32_t k0, uint32_t k1, uint32_t stream_id,
         uint64_t frame_offset, uint8_t* __restrict__ data_out, uint8_t* __restrict__ cw)
{
    extern __shared__ unsigned long long s_u[];      // kw packed info words
    for (int64_t f = blockIdx.x; f < frames; f += gridDim.x) {
        __syncthreads();
        for (int w = threadIdx.x; w < kw; w += blockDim.x) {
            unsigned long long word = 0;
            if (data_in) {
                for (int b = 0; b < 64 && w * 64 + b < k; ++b)
                    word |= (unsigned long long)(data_in[f * k + w * 64 + b] & 1u) << b;
            } else {
                // random info bits: Philox counter (word/2, frame, ~stream) -- disjoint from the noise counters,
                // which use the plain stream id (awgn_philox.cuh)
                const Philox4 p = philox4x32_10((uint32_t)(w >> 1), (uint32_t)(frame_offset + f),
                                                (uint32_t)((frame_offset + f) >> 32), ~stream_id, k0, k1);
                word = (w & 1) ? ((unsigned long long)p.w << 32 | p.z) : ((unsigned long long)p.y << 32 | p.x);
                const int rem = k - w * 64;
                if (rem < 64) word &= (rem <= 0) ? 0ull : ((1ull << rem) - 1ull);
            }
            s_u[w] = word;
        }
        __syncthreads();
        uint8_t* out = cw + f * n;
        for (int j = threadIdx.x; j < k; j += blockDim.x) {         // systematic part
            const uint8_t bit = (uint8_t)((s_u[j >> 6] >> (j & 63)) & 1ull);
            out[pos ? pos[j] : j] = bit;
            if (data_out) data_out[f * k + j] = bit;
        }
        for (int i = threadIdx.x; i < m; i += blockDim.x) {         // parity part: A u mod 2
            const uint64_t* row = a_bits + (size_t)i * kw;
            unsigned acc = 0;
            for (int w = 0; w < kw; ++w) acc ^= (unsigned)__popcll(row[w] & s_u[w]);
            out[pos ? pos[k + i] : k + i] = (uint8_t)(acc & 1u);
        }
    }
}

}  // namespace
}  // namespace ldpc

using namespace ldpc;

extern "C" int ldpc_encoder_create(int m, int n, const uint64_t* h_std_bits, const int32_t* out_pos, ldpc_encoder** out)
{
    if (!out) { set_error("null out"); return LDPC_ERR_INVALID; }
    *out = nullptr;
    if (m <= 0 || n <= m || !h_std_bits) { set_error("bad standard-form matrix (m=%d, n=%d)", m, n); return LDPC_ERR_INVALID; }
    DeviceInfo di;
    int rc = get_device_info(&di);
    if (rc) return rc;
    const int k = n - m, W = (n + 63) / 64, kw = (k + 63) / 64;
    // keep only the A part (first k columns) of every row; check the identity part
    std::vector<uint64_t> a((size_t)m * kw, 0);
    for (int i = 0; i < m; ++i) {
        const uint64_t* row = h_std_bits + (size_t)i * W;
        for (int j = 0; j < k; ++j)
            if ((row[j >> 6] >> (j & 63)) & 1) a[(size_t)i * kw + (j >> 6)] |= (uint64_t)1 << (j & 63);
        for (int j = k; j < n; ++j) {
            const bool bit = (row[j >> 6] >> (j & 63)) & 1;
            if (bit != (j - k == i)) { set_error("matrix is not in [A | I] form at row %d, column %d", i, j); return LDPC_ERR_INVALID; }
        }
    }
    ldpc_encoder* e = new (std::nothrow) ldpc_encoder();
    if (!e) { set_error("out of host memory"); return LDPC_ERR_NOMEM; }
    e->m = m; e->n = n; e->k = k; e->kw = kw;
    if (cudaMalloc((void**)&e->d_a, a.size() * 8) != cudaSuccess ||
        cudaMemcpy(e->d_a, a.data(), a.size() * 8, cudaMemcpyHostToDevice) != cudaSuccess) {
        set_error("cannot upload the encoder matrix: %s", cudaGetErrorString(cudaGetLastError()));
        ldpc_encoder_destroy(e);
        return LDPC_ERR_CUDA;
    }
    if (out_pos) {
        std::vector<char> seen(n, 0);
        for (int j = 0; j < n; ++j) {
            if (out_pos[j] < 0 || out_pos[j] >= n || seen[out_pos[j]]) { set_error("out_pos is not a permutation"); ldpc_encoder_destroy(e); return LDPC_ERR_INVALID; }
            seen[out_pos[j]] = 1;
        }
        if (cudaMalloc((void**)&e->d_pos, sizeof(int32_t) * n) != cudaSuccess ||
            cudaMemcpy(e->d_pos, out_pos, sizeof(int32_t) * n, cudaMemcpyHostToDevice) != cudaSuccess) {
            set_error("cannot upload the output permutation: %s", cudaGetErrorString(cudaGetLastError()));
            ldpc_encoder_destroy(e);
            return LDPC_ERR_CUDA;
        }
    }
    *out = e;
    return LDPC_OK;
}

extern "C" void ldpc_encoder_destroy(ldpc_encoder* e)
{
    if (!e) return;
    cudaFree(e->d_a);
    cudaFree(e->d_pos);
    delete e;
}

extern "C" int ldpc_encode_batch(const ldpc_encoder* e, int64_t frames, const uint8_t* data_dev, uint64_t seed,
                                 uint32_t stream_id, uint64_t frame_offset, uint8_t* data_out_dev,
                                 uint8_t* codeword_dev, void* stream)
{
    if (!e || !codeword_dev || frames < 0) { set_error("bad encoder arguments"); return LDPC_ERR_INVALID; }
    if (frames == 0) return LDPC_OK;
    DeviceInfo di;
    int rc = get_device_info(&di);
    if (rc) return rc;
    const int grid = (int)std::min<int64_t>(frames, (int64_t)di.sm_count * 8);
    k_encode<<<grid, 256, sizeof(unsigned long long) * e->kw, (cudaStream_t)stream>>>(
        e->m, e->n, e->k, e->kw, e->d_a, e->d_pos, frames, data_dev, (uint32_t)seed, (uint32_t)(seed >> 32), stream_id,
        frame_offset, data_out_dev, codeword_dev);
    LDPC_LAUNCH_CHECK();
    return LDPC_OK;
}
