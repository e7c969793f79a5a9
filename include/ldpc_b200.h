/*
 * ldpc_b200.h -- C ABI of the B200-native LDPC sum-product (SPA) decode path.
 *
 * This is the drop-in boundary for the one hot path of omkuprin7/ldpc-simulator
 * (python_ldpc_app/): SPA_Decoder.decode and the Monte-Carlo loop around it.
 * The reference has no FFI of its own (it is pure Python), so every entry
 * point below names the reference interface whose work it replaces
 * (file:line relative to python_ldpc_app/).  The Python-side binding is
 * ldpc-simulator_b200/_native.py (ctypes); INTEGRATION.md shows the stub a
 * maintainer of the reference would add.
 *
 * Conventions
 *   - plain C types only: pointers, sizes, ints.  No torch / CUDA types; a
 *     CUDA stream is passed as void* (cudaStream_t), NULL = default stream.
 *   - every function returns 0 on success or a negative ldpc_status; the
 *     message for the calling thread is available from ldpc_last_error().
 *     Nothing throws or longjmps across the boundary.
 *   - "dev" pointers are device memory owned by the caller (e.g. torch
 *     tensors); the library owns only the opaque ldpc_graph handles.
 *   - a graph handle is immutable after creation and may be shared between
 *     streams and threads; a workspace belongs to one in-flight call.
 *   - a graph handle lives on the CUDA device that was current when it was
 *     created (its tables, run-time compiled modules and the staging
 *     pipelines of ldpc_decode_batch_host).  Compute calls made while another
 *     device is current fail with LDPC_ERR_INVALID; a process that drives
 *     several GPUs creates one handle per device.
 *   - there is no CPU fallback: without a CUDA device every compute entry
 *     point fails with LDPC_ERR_CUDA.  The ldpc_host_* helpers are pure host
 *     integer code (graph analysis) and need no device.
 *
 * Bit/LLR conventions are the reference's: LLR < 0 means bit 0,
 * z = (posterior < 0) is the COMPLEMENT of the decided bit and is what
 * SPA_Decoder.decode leaves in _decoded_data (spa_decoder.py:188,233,245).
 */
#ifndef LDPC_B200_H
#define LDPC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LDPC_B200_ABI_VERSION 4

typedef enum ldpc_status {
    LDPC_OK = 0,
    LDPC_ERR_INVALID = -1,     /* bad argument (mirrors Result.INVALID_INPUT, enums.py:6) */
    LDPC_ERR_CUDA = -2,        /* CUDA runtime error / no device */
    LDPC_ERR_WORKSPACE = -3,   /* workspace too small */
    LDPC_ERR_UNSUPPORTED = -4, /* e.g. resident kernel requested for a non-QC graph */
    LDPC_ERR_NOMEM = -5
} ldpc_status;

/* Arithmetic of the decoder. */
typedef enum ldpc_dtype {
    LDPC_F64 = 0,      /* parity mode: the reference's fp64 formulas literally (spa_decoder.py:133-168) */
    LDPC_F32 = 1,      /* same formulas in fp32 with accurate tanhf/atanhf, any graph */
    LDPC_F32_FAST = 2  /* throughput mode, fp32 with MUFU approximations: the SM-resident kernel on quasi-cyclic
                          graphs, the generic kernels on every other graph (with a MUFU check node when no
                          check has more than 24 edges, else identical to LDPC_F32) */
} ldpc_dtype;

/* Flags for ldpc_decode_batch / ldpc_mc_run. */
#define LDPC_FLAG_EARLY_TERM    0x1u /* stop a frame at its first zero syndrome (reference behaviour,
                                        spa_decoder.py:231-241).  Without it every frame runs exactly
                                        max_iter passes and the syndrome is taken after the last one. */
#define LDPC_FLAG_COMPACT       0x2u /* generic kernels: compact converged frames out of the active list */
#define LDPC_FLAG_FIX_ODD_SIGN  0x4u /* NOT the reference: negate extrinsics of odd-degree checks, which
                                        makes the inverted LLR convention self-consistent */
#define LDPC_FLAG_FORCE_GENERIC 0x8u /* never pick the resident QC kernel */
#define LDPC_FLAG_TABLE_KERNEL  0x10u /* resident path: use the table-driven kernel even when a kernel
                                         specialised at build time for this base matrix exists (tests) */
#define LDPC_FLAG_NO_JIT        0x20u /* resident path: never specialise the kernel at run time (NVRTC);
                                         unregistered base matrices then use the table-driven kernel */

#define LDPC_FLAG_NO_REPLAY     0x80u /* ldpc_decode_batch_host: do not replay tiny calls (<= 32 frames, generic kernels)
                                         from a captured CUDA graph; launch the kernels one by one (tests) */
#define LDPC_FLAG_ONE_FRAME     0x100u /* resident path: use the one-frame-per-thread kernel even for batches that would
                                          run the two-frames-per-thread kernel (identical results; tests, A/B timing) */
#define LDPC_FLAG_PAIR_REGS     0x200u /* resident path, two-frames-per-thread kernel: keep the messages in registers
                                          instead of tensor memory (identical results; tests, A/B timing) */
#define LDPC_FLAG_PAIR_SCATTER  0x400u /* resident path, two-frames-per-thread kernel: accumulate the posterior in place
                                          (one barrier per group of block rows) instead of the barrier-free check-node
                                          phase + gather (identical results; tests, A/B timing) */
#define LDPC_FLAG_PAIR_GATHER   0x800u /* resident path: use the two-frames-per-thread gather kernel also with early
                                          termination (default there: the one-frame kernel) */
#define LDPC_FLAG_LLR_F16       0x1000u /* ldpc_decode_batch_host only, LDPC_F32 / LDPC_F32_FAST: llr_host holds IEEE half precision
                                           values (2 bytes per LLR over PCIe instead of 4); they are widened to fp32 on the
                                           device before decoding.  Not the reference's input type: decisions are those of the
                                           rounded LLRs (measured agreement: bench.py e2e_f16_ingest) */
#define LDPC_FLAG_LLR_I8        0x2000u /* ldpc_decode_batch_host only, fp32 precisions: llr_host holds int8 fixed-point values q,
                                           LLR = q / 4 (range +-31.75, step 0.25; 1 byte per LLR over PCIe), widened on the
                                           device.  A quantised input, labelled as such wherever it is measured */
#define LDPC_FLAG_ONE_GATHER    0x4000u /* resident path: one frame per thread with the gather structure and tensor-memory
                                           messages (four CTAs per SM); identical results */
#define LDPC_FLAG_NORM_LLR      0x40u /* ldpc_mc_run: also accumulate the "normalized LLR" metric (spa_decoder.py:210-228)
                                         in counters[5]; runs the generic kernels, which carry the metric */

/* Channel flags (the sigma_sq_quirk argument of ldpc_mc_run / ldpc_channel_llr). */
#define LDPC_CHANNEL_SIGMA_SQ   0x1
#define LDPC_CHANNEL_AMP_07     0x2

/* Which kernel family a (graph, dtype, flags) combination runs on (ldpc_graph_prepare). */
typedef enum ldpc_kernel_kind {
    LDPC_KERNEL_GENERIC = 0,       /* frame-minor streaming kernels, any graph */
    LDPC_KERNEL_QC_TABLE = 1,      /* SM-resident, shift tables read at run time */
    LDPC_KERNEL_QC_REGISTERED = 2, /* SM-resident, specialised when the library was built */
    LDPC_KERNEL_QC_JIT = 3         /* SM-resident, specialised at run time with NVRTC */
} ldpc_kernel_kind;

typedef struct ldpc_graph ldpc_graph;

/* ------------------------------------------------------------------------- *
 * Host-side graph analysis (no device needed).
 * ------------------------------------------------------------------------- */

/*
 * Edge index of a parity-check pattern given as CSR (columns ascending within a
 * row).  Replaces SPA_Decoder._init_neighbor_structures (spa_decoder.py:44-61)
 * and EncoderDecoderData._init_decoder_structures (encoder_decoder_data.py:
 * 718-747): edges are numbered in CSR order (the reference's COO order);
 * col_ptr[n+1] / csc_edge[nnz] list, for every column, its edge numbers in
 * ascending row order; edge_row[nnz] is the row of each edge.
 */
int ldpc_host_edge_index(int m, int n, const int32_t* row_ptr, const int32_t* col_idx,
                         int32_t* col_ptr, int32_t* csc_edge, int32_t* edge_row);

/*
 * Quasi-cyclic structure detection (new; north_star item 1): finds the largest
 * z such that H is an (m/z) x (n/z) array of z x z blocks that are each zero
 * or a single cyclic permutation.  Returns 1 and fills z/mb/nb/shift (row-major
 * mb*nb, -1 = zero block, else s with H[r, (r+s) mod z] = 1) when found, 0 when
 * the matrix is not quasi-cyclic, <0 on error.  shift_cap = capacity of shift.
 */
int ldpc_host_detect_qc(int m, int n, const int32_t* row_ptr, const int32_t* col_idx,
                        int* z, int* mb, int* nb, int16_t* shift, int64_t shift_cap);

/*
 * Standard form [A | I] by bit-packed GF(2) Gauss-Jordan.  Replaces
 * gaussian_elimination + create_standart_parity_check_matrix
 * (encoder_decoder_data.py:13-183, 269-317) with identical results: same pivot
 * rule, same rank repair, same column permutation.
 *   h_std_bits  [m][words] uint64, words = (n+63)/64, bit c of row r set iff
 *               H_std[r][c] = 1 (only the first *rank rows are meaningful)
 *   perm        [n]  H_std[:, c] = H_reduced[:, perm[c]]
 */
int ldpc_host_standard_form(int m, int n, const int32_t* row_ptr, const int32_t* col_idx,
                            uint64_t* h_std_bits, int32_t* perm, int32_t* rank);

/* ------------------------------------------------------------------------- *
 * Graph handles (device resident tables).
 * ------------------------------------------------------------------------- */

/* Any parity-check pattern (raw ALIST H or the dense H_std).  Replaces
 * SPA_Decoder.__init__ (spa_decoder.py:16-42).  Quasi-cyclic structure is
 * detected automatically and enables LDPC_F32_FAST. */
int ldpc_graph_create_csr(int m, int n, int64_t nnz, const int32_t* row_ptr,
                          const int32_t* col_idx, ldpc_graph** out);

/* Directly from a shift table (row-major mb*nb, -1 = zero block). */
int ldpc_graph_create_qc(int z, int mb, int nb, const int16_t* shift, ldpc_graph** out);

/* Query. Any out pointer may be NULL. */
int ldpc_graph_info(const ldpc_graph* g, int* m, int* n, int64_t* nnz, int* max_check_degree,
                    int* max_var_degree, int* qc_z, int* qc_mb, int* qc_nb);

/* Copies the detected shift table (capacity in entries); returns 1 if QC, 0 if not. */
int ldpc_graph_qc_shifts(const ldpc_graph* g, int16_t* shift, int64_t shift_cap);

/*
 * Selects -- and, for LDPC_KERNEL_QC_JIT, compiles and loads -- the kernels ldpc_decode_batch /
 * ldpc_mc_run will use for this graph, so that the first decode call does not pay for it.  Optional:
 * the decode calls do the same lazily.  *kind (may be NULL) receives an ldpc_kernel_kind.  A base
 * matrix that is not in the build-time registry is specialised with NVRTC (about 2 s once per code and
 * machine; cubins are cached under $LDPC_JIT_CACHE, default ~/.cache/ldpc_b200; the NVRTC library is
 * found through $LDPC_NVRTC_LIB, the loader path or /usr/local/cuda/lib64).
 */
int ldpc_graph_prepare(const ldpc_graph* g, int dtype, unsigned flags, int* kind);

/*
 * Host only (no device needed): the block-row schedule the specialised resident kernel uses for a base
 * matrix, written as the C++ type text the kernel is instantiated with (type_text, may be NULL), and,
 * when cubin_bytes is not NULL, a trial NVRTC compilation for sm_100a reporting the size of the cubin.
 */
int ldpc_host_jit_compile(int z, int mb, int nb, const int16_t* shift, char* type_text, size_t type_cap,
                          size_t* cubin_bytes);

void ldpc_graph_destroy(ldpc_graph* g);

/* ------------------------------------------------------------------------- *
 * Decoding.
 * ------------------------------------------------------------------------- */

/* Bytes of workspace that let ldpc_decode_batch process `frames` frames in one
 * chunk.  A smaller workspace is accepted (frames are then processed in
 * several chunks) down to ldpc_workspace_bytes(g, 32, dtype). */
size_t ldpc_workspace_bytes(const ldpc_graph* g, int64_t frames, int dtype);
/* The same for a call with these flags / with a norm_llr_dev output: LDPC_FLAG_FORCE_GENERIC, LDPC_FLAG_NO_JIT,
 * LDPC_FLAG_TABLE_KERNEL and the normalized-LLR output move an LDPC_F32_FAST call from the resident kernel
 * (256 bytes) to the generic kernels.  ldpc_workspace_bytes(g, F, t) == ldpc_workspace_bytes_ex(g, F, t, 0, 0). */
size_t ldpc_workspace_bytes_ex(const ldpc_graph* g, int64_t frames, int dtype, unsigned flags, int want_norm);

/*
 * Decode `frames` frames.  Replaces SPA_Decoder.decode (spa_decoder.py:63-280),
 * one call per batch instead of one per frame.  Asynchronous on `stream`.
 *
 *   llr_dev       [frames][n] row-major channel LLRs in graph-column order
 *                 (DataBuffer._channel_data, spa_decoder.py:88); double for
 *                 LDPC_F64, float otherwise
 *   z_dev         [frames][n] uint8, z = (posterior < 0)   (:188,233,245)
 *   conv_iter_dev [frames] int32, 0-based pass index at convergence, -1 = not
 *                 converged (SPA_Decoder.convergence_iteration, :65,232)
 *   ok_dev        [frames] uint8, 1 = Result.OK, 0 = DATA_TRANSFER_NOT_OK (:241,253)
 *   post_dev      [frames][n] posterior LLRs of the exit pass (same type as llr)
 *                 or NULL
 *   norm_llr_dev  [frames] float "normalized LLR" of the exit pass
 *                 (_d_summarize_normalized_llr, :210-228,237-251) or NULL;
 *                 k_info = number of leading bits it is taken over (n - m)
 */
int ldpc_decode_batch(const ldpc_graph* g, int dtype, int64_t frames, int max_iter, unsigned flags,
                      const void* llr_dev, uint8_t* z_dev, int32_t* conv_iter_dev, uint8_t* ok_dev,
                      void* post_dev, float* norm_llr_dev, int k_info,
                      void* workspace_dev, size_t workspace_bytes, void* stream);

/*
 * Same, with HOST buffers: the library stages chunks through pinned memory and
 * overlaps host<->device copies with decoding on its own streams, then blocks
 * until the results are in the host arrays.  This is the end-to-end call
 * behind SPA_Decoder.decode / decode_batch.  With LDPC_FLAG_LLR_F16 llr_host is
 * [frames][n] half precision.  z_host may be NULL when only
 * zbits_host ([frames][4*ceil(n/32)] bytes, bit j of a frame = z[j], LSB first) is wanted.
 */
int ldpc_decode_batch_host(const ldpc_graph* g, int dtype, int64_t frames, int max_iter, unsigned flags,
                           const void* llr_host, uint8_t* z_host, uint8_t* zbits_host,
                           int32_t* conv_iter_host, uint8_t* ok_host, void* post_host,
                           float* norm_llr_host, int k_info);

/*
 * Monte-Carlo kernel: generate `frames` BPSK-AWGN frames on the device
 * (Philox4x32-10 + Box-Muller), decode them and fold the error counters.
 * Replaces the per-frame loop of run_simulation / process_block
 * (main.py:295-339, 43-146) together with Channel.process mode 1
 * (channel.py:38-81).
 *
 *   speed, snr_db      sigma = 1/sqrt(2*speed*10^(snr_db/10))       (channel.py:113)
 *   sigma_sq_quirk     channel flags: LDPC_CHANNEL_SIGMA_SQ (1) = noise stddev sigma^2 as the reference
 *                      does (channel.py:68), else sigma; LDPC_CHANNEL_AMP_07 (2) = symbols +-0.7
 *                      ("modulation 2", channel.py:50-51) instead of +-1
 *   seed, stream_id, frame_offset
 *                      Philox key = seed; counter = (stream_id, frame_offset + frame, word) so that
 *                      ranks / launches draw disjoint streams
 *   codeword_dev       uint8 transmitted codeword(s) in graph-column order, NULL = all-zero
 *   codeword_stride    0: one codeword [n] for every frame; >= n: frame f sends codeword_dev + f*stride
 *                      (e.g. the output of ldpc_encode_batch)
 *   info_mask_dev      [n] uint8, 1 = information position; NULL = the first k_info positions
 *   k_info             number of information bits; BER is taken over them (main.py:323-330)
 *   counters_dev       uint64[5], ACCUMULATED: frames, failed frames, info-bit errors counted in
 *                      failed frames only, sum of convergence iterations, converged frames
 *                      (main.py:314-339).  With LDPC_FLAG_NORM_LLR the array is uint64[6]:
 *                      counters[5] += sum over ALL frames of the exit pass' sign-change count over
 *                      the first k_info bits with |L| <= 7 (spa_decoder.py:218-228); the reference's
 *                      avg_normalized_llr of a point is counters[5] / (k_info * frames) (main.py:332-334,357)
 */
int ldpc_mc_run(const ldpc_graph* g, int dtype, int64_t frames, int max_iter, unsigned flags,
                double speed, double snr_db, int sigma_sq_quirk,
                uint64_t seed, uint32_t stream_id, uint64_t frame_offset,
                const uint8_t* codeword_dev, int64_t codeword_stride, const uint8_t* info_mask_dev, int k_info,
                uint64_t* counters_dev, void* workspace_dev, size_t workspace_bytes, void* stream);

/*
 * The channel of Channel.create_channel(speed, sn1, sn2, mode, p, mod) (channel.py:102-125) for the device generator.
 * mode 1: AWGN; mode 2: AWGN + interference in a share p of the band (every bit is hit with probability
 * P(a/n < p), a uniform in 0..n-1, :86-96); mode 3: AWGN + counter interference weighted with p (:98-100).
 * The Gaussian terms come from Philox, not from the reference's Park-Miller generators: same distribution,
 * not the same numbers (the host-side Channel.process reproduces the reference's stream bit for bit).
 */
typedef struct ldpc_channel {
    int mode;                    /* 1, 2 or 3 */
    int modulation;              /* 1: symbols +-1, 2: +-0.7 (channel.py:49-51) */
    int sigma_sq_quirk;          /* mode 1 only: noise deviation sigma^2 as in the reference (channel.py:68) */
    double speed;                /* plays the role of the code rate in sigma (channel.py:113) */
    double snr_db;               /* sn1 */
    double interference_snr_db;  /* sn2, modes 2 and 3 */
    double p;                    /* modes 2 and 3 */
} ldpc_channel;

/* ldpc_mc_run / ldpc_channel_llr with a full channel description. */
int ldpc_mc_run_ex(const ldpc_graph* g, int dtype, int64_t frames, int max_iter, unsigned flags,
                   const ldpc_channel* channel, uint64_t seed, uint32_t stream_id, uint64_t frame_offset,
                   const uint8_t* codeword_dev, int64_t codeword_stride, const uint8_t* info_mask_dev, int k_info,
                   uint64_t* counters_dev, void* workspace_dev, size_t workspace_bytes, void* stream);
int ldpc_channel_llr_ex(int n, int dtype, int64_t frames, const ldpc_channel* channel,
                        uint64_t seed, uint32_t stream_id, uint64_t frame_offset,
                        const uint8_t* codeword_dev, int64_t codeword_stride, void* llr_dev, void* stream);

/* Workspace for ldpc_mc_run (it also holds the generated LLRs and decoder outputs). */
size_t ldpc_mc_workspace_bytes(const ldpc_graph* g, int64_t frames, int dtype);
size_t ldpc_mc_workspace_bytes_ex(const ldpc_graph* g, int64_t frames, int dtype, unsigned flags);   /* for a call with these flags */

/* Fill llr_dev [frames][n] (float, or double when dtype == LDPC_F64) with the channel
 * output ldpc_mc_run would decode -- used by the tests to feed identical frames to the oracle. */
int ldpc_channel_llr(int n, int dtype, int64_t frames, double speed, double snr_db, int sigma_sq_quirk,
                     uint64_t seed, uint32_t stream_id, uint64_t frame_offset,
                     const uint8_t* codeword_dev, int64_t codeword_stride, void* llr_dev, void* stream);

/* ------------------------------------------------------------------------- *
 * Systematic encoder on the device (random-codeword Monte-Carlo).
 * ------------------------------------------------------------------------- */
typedef struct ldpc_encoder ldpc_encoder;

/*
 * Encoder for H_std = [A | I_m] as produced by ldpc_host_standard_form (h_std_bits [m][(n+63)/64]).
 * Replaces create_generator_matrix + DataBuffer.encode (encoder_decoder_data.py:319-344,
 * data_buffer.py:47-82): codeword in H_std column order = [u | A u mod 2].  out_pos (may be NULL)
 * gives, for every H_std column j, the position it is written to -- pass the column permutation of
 * the standard form to obtain codewords in raw ALIST order.
 */
int ldpc_encoder_create(int m, int n, const uint64_t* h_std_bits, const int32_t* out_pos, ldpc_encoder** out);
void ldpc_encoder_destroy(ldpc_encoder* e);

/*
 * Encode `frames` frames.  data_dev [frames][k] uint8 info bits, or NULL to draw them on the device
 * (Philox, replaces Generator.generate_bit_sequence, generator.py:7-9; key = seed, counters disjoint
 * from the noise of the same stream_id / frame).  data_out_dev [frames][k] (may be NULL) receives the
 * info bits; codeword_dev [frames][n] the codewords.  Asynchronous on `stream`.
 */
int ldpc_encode_batch(const ldpc_encoder* e, int64_t frames, const uint8_t* data_dev, uint64_t seed,
                      uint32_t stream_id, uint64_t frame_offset, uint8_t* data_out_dev,
                      uint8_t* codeword_dev, void* stream);

/* Number of kernels this library launched since it was loaded (bench.py's gpu_launches). */
uint64_t ldpc_kernel_launch_count(void);

/* Saturating MUFU micro-benchmark: returns MUFU ops/s of this device in *ops_per_s
 * (roofline denominator for the resident kernel). */
int ldpc_measure_mufu_peak(double* ops_per_s, void* stream);

const char* ldpc_last_error(void);
int ldpc_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* LDPC_B200_H */
