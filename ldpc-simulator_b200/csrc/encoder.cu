// encoder.cu -- systematic encoder on the device, for Monte-Carlo runs with a fresh random codeword per frame.
//
// Replaces Generator.generate_bit_sequence + DataBuffer.encode (python_ldpc_app/generator.py:7-9,
// data_buffer.py:47-82) with G = [I_k | A^T] built from H_std = [A | I_m]
// (encoder_decoder_data.py:319-344): codeword (in H_std column order) = [u | A u mod 2].
// A is kept bit-packed ([m][ceil(k/64)] words, L2 resident); a thread computes parity bit i of a tile of 32 frames
// from bit-sliced look-up tables (below).  An optional output permutation writes the
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

// One CTA encodes a TILE of 32 frames, bit-sliced ("method of the four Russians"):
//   1. the packed info words of the tile go to shared memory, frame-major (Philox counters are per frame, so a frame's
//      bits do not depend on how the batch is tiled);
//   2. they are transposed into slices U_j = bit j of all 32 frames (one warp ballot per info bit);
//   3. for every group of four info bits a 16-entry table of the XOR combinations of its slices is built;
//   4. parity bit i of all 32 frames is then the XOR of one table entry per nibble of row i of A: 288 look-ups for
//      WiMAX-2304 where the popc(A_row & u) formulation needs 18 words x 32 frames.  The look-ups of a warp hit 16
//      consecutive words per group -- 16 banks, identical words broadcast -- so they are conflict free.
// History (tools/mc_et_probe.py, Monte-Carlo point at 4 dB with early termination, 262 144 frames): one frame per CTA
// re-read A from L2 for every frame (166 KB per frame, 6 ms = the L2 bandwidth; the point took 11.9 ms); a 32-frame tile
// with popc 4.1 ms (10.6 ms); assembling the rows in shared memory for coalesced stores was slower (13.4 ms).
constexpr int kTile = 32;

__global__ void __launch_bounds__(256)
k_encode(int m, int n, int k, int kw, const uint64_t* __restrict__ a_bits, const int32_t* __restrict__ pos,
         int64_t frames, const uint8_t* __restrict__ data_in, uint32_t k0, uint32_t k1, uint32_t stream_id,
         uint64_t frame_offset, uint8_t* __restrict__ data_out, uint8_t* __restrict__ cw)
{
    extern __shared__ unsigned long long s_u[];                  // [kTile][kw] packed info words, frame-major
    const int kpad = kw * 64;                                     // info bits rounded up to whole words
    uint32_t* s_slice = reinterpret_cast<uint32_t*>(s_u + (size_t)kTile * kw);     // [kpad] bit j of the 32 frames
    uint32_t* s_tab = s_slice + kpad;                             // [kpad / 4][16] XOR combinations of four slices
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, warps = blockDim.x >> 5;
    const int64_t tiles = (frames + kTile - 1) / kTile;
    for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const int64_t f0 = tile * kTile;
        const int nf = (int)((frames - f0 < kTile) ? (frames - f0) : kTile);
        __syncthreads();
        for (int q = threadIdx.x; q < kTile * kw; q += blockDim.x) {
            const int fl = q / kw, w = q - fl * kw;
            const int64_t f = f0 + fl;
            unsigned long long word = 0;
            if (fl < nf) {
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
            }
            s_u[q] = word;
        }
        __syncthreads();
        for (int j = warp; j < kpad; j += warps) {                   // transpose: lane = frame, ballot = slice of bit j
            const unsigned bit = (unsigned)((s_u[lane * kw + (j >> 6)] >> (j & 63)) & 1ull);
            const unsigned slice = __ballot_sync(0xffffffffu, bit != 0);
            if (lane == 0) s_slice[j] = slice;
        }
        __syncthreads();
        for (int q = threadIdx.x; q < (kpad / 4) * 16; q += blockDim.x) {     // tables: entry v of group g
            const int g = q >> 4, v = q & 15;
            uint32_t x = 0;
            if (v & 1) x ^= s_slice[4 * g];
            if (v & 2) x ^= s_slice[4 * g + 1];
            if (v & 4) x ^= s_slice[4 * g + 2];
            if (v & 8) x ^= s_slice[4 * g + 3];
            s_tab[q] = x;
        }
        for (int q = threadIdx.x; q < nf * k; q += blockDim.x) {    // systematic part
            const int fl = q / k, j = q - fl * k;
            const uint8_t bit = (uint8_t)((s_u[fl * kw + (j >> 6)] >> (j & 63)) & 1ull);
            cw[(f0 + fl) * n + (pos ? pos[j] : j)] = bit;
            if (data_out) data_out[(f0 + fl) * k + j] = bit;
        }
        __syncthreads();
        for (int i = threadIdx.x; i < m; i += blockDim.x) {         // parity part: A u mod 2, row i for all frames of the tile
            const uint64_t* row = a_bits + (size_t)i * kw;
            uint32_t acc = 0;
            const uint32_t* tab = s_tab;
            for (int w = 0; w < kw; ++w, tab += 256) {
                const unsigned long long r = row[w];
                const uint32_t lo = (uint32_t)r, hi = (uint32_t)(r >> 32);
#pragma unroll
                for (int q = 0; q < 8; ++q) acc ^= tab[16 * q + ((lo >> (4 * q)) & 15u)];
#pragma unroll
                for (int q = 0; q < 8; ++q) acc ^= tab[128 + 16 * q + ((hi >> (4 * q)) & 15u)];
            }
            const int64_t col = pos ? pos[k + i] : k + i;
#pragma unroll
            for (int fl = 0; fl < kTile; ++fl)
                if (fl < nf) cw[(f0 + fl) * n + col] = (uint8_t)((acc >> fl) & 1u);
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
    // shared memory of a tile: packed info words (32 x kw x 8) + slices (4 per info bit) + tables (16 per info bit)
    const size_t smem = sizeof(unsigned long long) * e->kw * kTile + (size_t)e->kw * 64 * 4 + (size_t)e->kw * 64 * 16;
    if (smem > (size_t)di.max_smem_optin) { set_error("encoder: %d info bits exceed the shared-memory tile", e->k); return LDPC_ERR_UNSUPPORTED; }
    if (smem > 48 * 1024) LDPC_CUDA_TRY(cudaFuncSetAttribute(k_encode, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int tile_frames = kTile;
    const int64_t tiles = (frames + tile_frames - 1) / tile_frames;
    const int grid = (int)std::min<int64_t>(tiles, (int64_t)di.sm_count * 8);
    k_encode<<<grid, 256, smem, (cudaStream_t)stream>>>(
        e->m, e->n, e->k, e->kw, e->d_a, e->d_pos, frames, data_dev, (uint32_t)seed, (uint32_t)(seed >> 32), stream_id,
        frame_offset, data_out_dev, codeword_dev);
    LDPC_LAUNCH_CHECK();
    return LDPC_OK;
}
