// spa_qc_resident.cu -- SM-resident sum-product decoder for quasi-cyclic codes.
//
// Throughput path (LDPC_F32_FAST) for SPA_Decoder.decode
// (python_ldpc_app/spa_decoder.py:63-280) on graphs whose parity-check matrix is
// an mb x nb array of z x z circulants (WiMAX 802.16e, 802.11n, ...).
//
// One CTA decodes one frame at a time and keeps ALL of its state on the SM for
// every iteration; HBM only sees the channel LLRs going in and the packed
// decisions coming out:
//   * thread r (0 <= r < z) owns check row r of every block row, and keeps the
//     check->variable messages of its rows in REGISTERS (eps[MB][DC]);
//   * shared memory holds three n-vectors: the channel LLRs and the posterior
//     of the previous / current pass (flooding schedule = the reference's,
//     spa_decoder.py:104-276: every check reads posteriors of the previous pass);
//   * the circulant shift is an index rotation (r + s) mod z, so a warp's 32
//     consecutive rows touch 32 consecutive (mod z) words of a column block:
//     conflict-free shared-memory accesses;
//   * the shift table is a __grid_constant__ kernel parameter: after full
//     unrolling every entry is an immediate constant-bank operand.
//
// Arithmetic.  The reference's check node, E = 2 atanh(prod tanh(M/2)) with
// |M| clipped at 35.03 (spa_decoder.py:133-168), is evaluated in fp32 in the
// "likelihood-ratio" form that has no cancellation and no saturation:
//   messages are stored as log2-likelihood ratios m = L * log2(e);
//   x = 2^-|m|                      (1 MUFU.EX2; x = (1-|t|)/(1+|t|), t = tanh(L/2))
//   prod_k (1 + x_k e), e^2 = 1  =  A + B e   (A, B sums of positive terms, FMA pipe)
//   |E| = lg2(A') - lg2(B')         (2 MUFU.LG2) with (A', B') the product over the
//                                    OTHER edges (prefix/suffix products, no division)
//   sign(E) = product of the other signs -- exactly the reference's formula,
//   including its behaviour on odd-degree checks (DESIGN.md "sign convention").
// The reference's two clips become min(|m|, 35.032 * log2 e) on the way in; its
// |tanh| <= 1e-10 branch is the natural x = 1 case here (no division anywhere).
//
// Roofline: 3 MUFU operations per edge and pass; no tensor cores --
// SPA is a sparse gather/scatter, not a contraction.
#include "ldpc_common.cuh"
#include "awgn_philox.cuh"

#include <algorithm>

namespace ldpc {

namespace {

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
// 2*atanh(0.99999999999999878) = 35.0320... LLR units (spa_decoder.py:140-146,167), in bits:
constexpr float kClipBits = 50.5405f;

template <int MB, int DC>
struct QcParams {
    // slot word: bits 0..15 column-block base (bc*z), bits 16..30 shift, bit 31 = first
    // block row touching this column block (starts the posterior accumulation)
    uint32_t slot[MB][DC];
    int deg[MB];      // circulants in each block row (0 for unused rows)
    int z, mb, n;
};

__device__ __forceinline__ float ex2_approx(float x)
{
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float lg2_approx(float x)
{
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

struct Outputs {
    uint8_t* z;          // [F][n] or null
    uint32_t* zbits;     // [F][ceil(n/32)] or null
    int32_t* conv_it;    // [F] or null
    uint8_t* ok;         // [F] or null
    float* post;         // [F][n] or null
};

// EARLY: stop a frame at its first zero syndrome (reference semantics).  The syndrome
// of pass p is obtained for free during pass p+1 (every check reads the signs of the
// previous posterior anyway), so a converged frame costs one extra check-node pass.
template <int MB, int DC, bool EARLY, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB)
k_qc_resident(const __grid_constant__ QcParams<MB, DC> p, const float* __restrict__ llr, Outputs out,
              long long frames, int max_iter, int fix_odd, McParams mc, unsigned long long* __restrict__ work_counter)
{
    extern __shared__ __align__(16) float smem[];
    float* lam_ch = smem;             // channel log2-LRs
    float* bufA = smem + p.n;
    float* bufB = smem + 2 * p.n;
    __shared__ long long s_frame;
    __shared__ unsigned long long s_cnt[5];
    __shared__ int s_err;

    const int r = threadIdx.x;
    const bool row_ok = r < p.z;
    const int n = p.n;
    if (threadIdx.x < 5) s_cnt[threadIdx.x] = 0;

    ChannelConst cc;
    cc.noise_dev = mc.noise_dev; cc.llr_scale = mc.llr_scale; cc.amp = mc.amp;
    cc.a2 = mc.a2; cc.l_hit = mc.l_hit; cc.hit_threshold = mc.hit_threshold;
    cc.k0 = (uint32_t)mc.seed; cc.k1 = (uint32_t)(mc.seed >> 32); cc.stream_id = mc.stream_id;

    for (long long f = blockIdx.x;; ) {
        if (EARLY) {      // frames differ in cost: dynamic schedule = compaction of the active set
            __syncthreads();
            if (threadIdx.x == 0) s_frame = (long long)atomicAdd(work_counter, 1ull);
            __syncthreads();
            f = s_frame;
        }
        if (f >= frames) break;

        // ---- prologue: channel LLRs -> shared memory, scaled to log2 units ----
        if (mc.active) {
            for (int q = threadIdx.x; q < (n + 3) / 4; q += THREADS) {
                uint32_t bits = 0;
                if (mc.codeword) {
                    const uint8_t* cw = mc.codeword + f * mc.codeword_stride;
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        if (4 * q + i < n && cw[4 * q + i]) bits |= 1u << i;
                }
                float v[4];
                channel_llr4(cc, mc.frame_offset + (uint64_t)f, (uint32_t)q, bits, v);
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    if (4 * q + i < n) lam_ch[4 * q + i] = v[i] * kLog2e;
            }
        } else {
            const float* src = llr + (size_t)f * n;
            if ((n & 3) == 0) {
                const float4* src4 = reinterpret_cast<const float4*>(src);
                for (int q = threadIdx.x; q < n / 4; q += THREADS) {
                    float4 v = __ldg(src4 + q);
                    v.x *= kLog2e; v.y *= kLog2e; v.z *= kLog2e; v.w *= kLog2e;
                    reinterpret_cast<float4*>(lam_ch)[q] = v;
                }
            } else {
                for (int j = threadIdx.x; j < n; j += THREADS) lam_ch[j] = __ldg(src + j) * kLog2e;
            }
        }
        __syncthreads();

        float eps[MB][DC];
#pragma unroll
        for (int b = 0; b < MB; ++b)
#pragma unroll
            for (int c = 0; c < DC; ++c) eps[b][c] = 0.f;

        const float* prev = lam_ch;
        float* nxt = bufA;
        int conv = -1;

        for (int it = 0; it < max_iter; ++it) {
            bool unsat = false;
#pragma unroll
            for (int b = 0; b < MB; ++b) {
                const int deg = p.deg[b];
                if (deg > 0 && row_ok) {
                    float x[DC];
                    uint32_t mbits[DC];
                    int off[DC];
                    uint32_t sgn = (fix_odd && (deg & 1)) ? 0x80000000u : 0u;
                    bool par = (deg & 1) != 0;      // parity of the estimates (L >= 0), spa_decoder.py:188-195
#pragma unroll
                    for (int c = 0; c < DC; ++c) {
                        x[c] = 0.f; mbits[c] = 0; off[c] = 0;
                        if (c < deg) {
                            const uint32_t w = p.slot[b][c];
                            int idx = r + (int)((w >> 16) & 0x7fffu);
                            idx = min((unsigned)idx, (unsigned)(idx - p.z));
                            off[c] = (int)(w & 0xffffu) + idx;
                            const float L = prev[off[c]];
                            const float mu = L - eps[b][c];                  // :260-268
                            if (EARLY) par ^= (L < 0.f);
                            mbits[c] = __float_as_uint(mu);
                            sgn ^= mbits[c];
                            x[c] = ex2_approx(-fminf(fabsf(mu), kClipBits)); // :133-146
                        }
                    }
                    if (EARLY) unsat |= par;
                    // prefix products fa + fb*e over slots 0..k
                    float fa[DC], fb[DC];
                    fa[0] = 1.f; fb[0] = x[0];
#pragma unroll
                    for (int k = 1; k < DC - 1; ++k) {
                        fa[k] = fmaf(fb[k - 1], x[k], fa[k - 1]);
                        fb[k] = fmaf(fa[k - 1], x[k], fb[k - 1]);
                    }
                    float sa = 1.f, sb = 0.f;      // suffix product over slots > k
#pragma unroll
                    for (int k = DC - 1; k >= 0; --k) {
                        float A, B;
                        if (k == DC - 1) { A = fa[DC - 2]; B = fb[DC - 2]; }
                        else if (k == 0) { A = sa; B = sb; }
                        else {
                            A = fmaf(fa[k - 1], sa, fb[k - 1] * sb);
                            B = fmaf(fa[k - 1], sb, fb[k - 1] * sa);
                        }
                        if (k < deg) {
                            const float mag = lg2_approx(A) - lg2_approx(B);  // :151-168
                            const uint32_t sbit = (sgn ^ mbits[k]) & 0x80000000u;
                            const float e = __uint_as_float(__float_as_uint(mag) | sbit);
                            eps[b][k] = e;
                            const bool first = (p.slot[b][k] >> 31) != 0;
                            const float base = first ? lam_ch[off[k]] : nxt[off[k]];
                            nxt[off[k]] = base + e;                           // :173-185
                        }
                        const float na = fmaf(sb, x[k], sa);
                        sb = fmaf(sa, x[k], sb);
                        sa = na;
                    }
                }
                __syncthreads();
            }
            if (EARLY && it > 0) {
                // the posterior of pass it-1 (in prev) satisfied every check -> it is the exit pass
                if (!__syncthreads_or(unsat)) { conv = it - 1; break; }
            }
            prev = nxt;
            nxt = (nxt == bufA) ? bufB : bufA;
        }

        // ---- exit: syndrome of the last posterior if it has not been seen yet ----
        // (prev now points at the posterior of the exit pass)
        if (conv < 0) {
            bool unsat = false;
            if (row_ok) {
#pragma unroll
                for (int b = 0; b < MB; ++b) {
                    const int deg = p.deg[b];
                    bool par = (deg & 1) != 0;
#pragma unroll
                    for (int c = 0; c < DC; ++c) {
                        if (c < deg) {
                            const uint32_t w = p.slot[b][c];
                            int idx = r + (int)((w >> 16) & 0x7fffu);
                            idx = min((unsigned)idx, (unsigned)(idx - p.z));
                            par ^= (prev[(int)(w & 0xffffu) + idx] < 0.f);
                        }
                    }
                    unsat |= par;
                }
            }
            if (!__syncthreads_or(unsat)) conv = max_iter - 1;
        }

        // ---- outputs ----
        const bool good = conv >= 0;
        if (threadIdx.x == 0) {
            if (out.conv_it) out.conv_it[f] = conv;
            if (out.ok) out.ok[f] = good ? 1 : 0;
        }
        if (out.zbits) {
            const int words = (n + 31) / 32;
            for (int base = 0; base < words * 32; base += THREADS) {
                const int j = base + threadIdx.x;
                const bool neg = (j < n) && (prev[j] < 0.f);
                const unsigned m = __ballot_sync(0xffffffffu, neg);
                if ((threadIdx.x & 31) == 0 && j < words * 32) out.zbits[(size_t)f * words + (j >> 5)] = m;
            }
        }
        if (out.z) {
            uint8_t* dst = out.z + (size_t)f * n;
            for (int j = threadIdx.x; j < n; j += THREADS) dst[j] = (uint8_t)(prev[j] < 0.f);   // :188
        }
        if (out.post) {
            float* dst = out.post + (size_t)f * n;
            for (int j = threadIdx.x; j < n; j += THREADS) dst[j] = prev[j] * kLn2;
        }
        if (mc.active) {
            // main.py:314-339: bit errors only in failed frames, on the un-complemented output
            if (threadIdx.x == 0) s_err = 0;
            __syncthreads();
            if (!good) {
                int errs = 0;
                const int span = mc.info_mask ? n : mc.k_info;
                for (int j = threadIdx.x; j < span; j += THREADS) {
                    if (mc.info_mask && !mc.info_mask[j]) continue;
                    const unsigned est = (prev[j] < 0.f) ? 0u : 1u;
                    const unsigned sent = mc.codeword ? mc.codeword[f * mc.codeword_stride + j] : 0u;
                    errs += (est != sent);
                }
#pragma unroll
                for (int o = 16; o; o >>= 1) errs += __shfl_xor_sync(0xffffffffu, errs, o);
                if ((threadIdx.x & 31) == 0 && errs) atomicAdd(&s_err, errs);
            }
            __syncthreads();
            if (threadIdx.x == 0) {
                s_cnt[0] += 1;
                if (!good) { s_cnt[1] += 1; s_cnt[2] += (unsigned)s_err; }
                else { s_cnt[3] += (unsigned)conv; s_cnt[4] += 1; }
            }
        }
        __syncthreads();     // shared buffers are reused by the next frame
        if (!EARLY) f += gridDim.x;
    }
    if (mc.active && mc.counters) {
        __syncthreads();
        if (threadIdx.x < 5 && s_cnt[threadIdx.x]) atomicAdd(&mc.counters[threadIdx.x], s_cnt[threadIdx.x]);
    }
}

// ---------------------------------------------------------------------------
template <int MB, int DC>
bool build_params(const ldpc_graph* g, QcParams<MB, DC>* p)
{
    const QcInfo& q = g->qc;
    if (!g->is_qc || q.mb > MB || q.z > 1024 || g->n > 65535) return false;
    memset(p, 0, sizeof(*p));
    p->z = q.z; p->mb = q.mb; p->n = g->n;
    std::vector<char> seen(q.nb, 0);
    for (int b = 0; b < q.mb; ++b) {
        int d = 0;
        for (int bc = 0; bc < q.nb; ++bc) {
            const int s = q.shift[(size_t)b * q.nb + bc];
            if (s < 0) continue;
            if (d >= DC) return false;
            uint32_t w = (uint32_t)(bc * q.z) | ((uint32_t)s << 16);
            if (!seen[bc]) { w |= 0x80000000u; seen[bc] = 1; }
            p->slot[b][d++] = w;
        }
        if (d < 2) return false;   // the prefix/suffix scheme needs degree >= 2
        p->deg[b] = d;
    }
    for (int bc = 0; bc < q.nb; ++bc)
        if (!seen[bc]) return false;      // a column block with no check: posterior never formed
    return true;
}

struct Shape { int mb, dc; };
// instantiated (MB, DC) shapes, smallest first
constexpr Shape kShapes[] = { {12, 7}, {6, 15}, {4, 22} };

int pick_shape(const ldpc_graph* g)
{
    if (!g->is_qc) return -1;
    int maxw = 0;
    for (int b = 0; b < g->qc.mb; ++b) {
        int d = 0;
        for (int bc = 0; bc < g->qc.nb; ++bc) d += g->qc.shift[(size_t)b * g->qc.nb + bc] >= 0;
        maxw = std::max(maxw, d);
        if (d < 2) return -1;
    }
    for (size_t i = 0; i < sizeof(kShapes) / sizeof(kShapes[0]); ++i)
        if (g->qc.mb <= kShapes[i].mb && maxw <= kShapes[i].dc) return (int)i;
    return -1;
}

template <int MB, int DC, bool EARLY, int THREADS, int MINB>
int launch(const ldpc_graph* g, int64_t frames, int max_iter, unsigned flags, const float* llr,
           const Outputs& out, const McParams& mc, void* ws, cudaStream_t stream)
{
    QcParams<MB, DC> p;
    if (!build_params<MB, DC>(g, &p)) { set_error("graph does not fit the resident kernel shape"); return LDPC_ERR_UNSUPPORTED; }
    DeviceInfo di;
    int rc = get_device_info(&di);
    if (rc) return rc;
    auto kern = k_qc_resident<MB, DC, EARLY, THREADS, MINB>;
    const size_t smem = sizeof(float) * 3 * (size_t)g->n;
    if ((int)smem > di.max_smem_optin) { set_error("n=%d needs %zu bytes of shared memory", g->n, smem); return LDPC_ERR_UNSUPPORTED; }
    static thread_local const void* configured = nullptr;      // function attributes are per device: the key holds the ordinal
    static thread_local size_t configured_smem = 0;
    static thread_local int configured_dev = -1;
    if (configured != (const void*)kern || configured_smem < smem || configured_dev != di.device) {
        LDPC_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        LDPC_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        configured = (const void*)kern; configured_smem = smem; configured_dev = di.device;
    }
    int per_sm = 0;
    LDPC_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, THREADS, smem));
    if (per_sm < 1) { set_error("resident kernel does not fit on an SM"); return LDPC_ERR_UNSUPPORTED; }
    const int grid = (int)std::min<int64_t>(frames, (int64_t)per_sm * di.sm_count);
    unsigned long long* counter = nullptr;
    if (EARLY) {
        // the work counter lives in the caller's workspace; reset it on the stream
        if (!ws) { set_error("early termination needs a workspace (work counter)"); return LDPC_ERR_WORKSPACE; }
        counter = (unsigned long long*)ws;
        LDPC_CUDA_TRY(cudaMemsetAsync(counter, 0, sizeof(unsigned long long), stream));
    }
    kern<<<grid, THREADS, smem, stream>>>(p, llr, out, (long long)frames, max_iter,
                                           (flags & LDPC_FLAG_FIX_ODD_SIGN) ? 1 : 0, mc, counter);
    LDPC_LAUNCH_CHECK();
    return LDPC_OK;
}

template <int MB, int DC>
int launch_shape(const ldpc_graph* g, int64_t frames, int max_iter, unsigned flags, const float* llr,
                 const Outputs& out, const McParams& mc, void* ws, cudaStream_t stream)
{
    const bool early = (flags & LDPC_FLAG_EARLY_TERM) != 0;
    const int z = g->qc.z;
#define LDPC_QC_GO(T, B)                                                                                  \
    return early ? launch<MB, DC, true, T, B>(g, frames, max_iter, flags, llr, out, mc, ws, stream)      \
                 : launch<MB, DC, false, T, B>(g, frames, max_iter, flags, llr, out, mc, ws, stream)
    if (z <= 32) { LDPC_QC_GO(32, 12); }
    if (z <= 64) { LDPC_QC_GO(64, 8); }
    if (z <= 96) { LDPC_QC_GO(96, 5); }
    set_error("z=%d is larger than the resident kernel supports (96)", z);
    return LDPC_ERR_UNSUPPORTED;
#undef LDPC_QC_GO
}

}  // namespace

bool qc_resident_supported(const ldpc_graph* g)
{
    if (!g || pick_shape(g) < 0 || g->qc.z > 96 || g->n > 65535) return false;
    std::vector<char> seen(g->qc.nb, 0);
    for (int b = 0; b < g->qc.mb; ++b)
        for (int bc = 0; bc < g->qc.nb; ++bc)
            if (g->qc.shift[(size_t)b * g->qc.nb + bc] >= 0) seen[bc] = 1;
    for (char s : seen) if (!s) return false;
    return (size_t)g->n * 12 <= 227 * 1024;
}

int qc_resident_decode(const ldpc_graph* g, int64_t frames, int max_iter, unsigned flags,
                       const float* llr_dev, uint8_t* z_dev, uint32_t* zbits_dev, int32_t* conv_dev,
                       uint8_t* ok_dev, float* post_dev, const McParams& mc, void* ws, size_t ws_bytes,
                       cudaStream_t stream)
{
    const int kind = qc_resident_kind(g, flags);
    if (kind == LDPC_KERNEL_GENERIC) { set_error("graph is not supported by the resident QC kernels"); return LDPC_ERR_UNSUPPORTED; }
    if (frames == 0) return LDPC_OK;
    if (ws && ws_bytes < 256) ws = nullptr;
    if (kind == LDPC_KERNEL_QC_REGISTERED)
        return qc_spec_decode(qc_spec_find(g), g, frames, max_iter, flags, llr_dev, z_dev, zbits_dev, conv_dev, ok_dev,
                              post_dev, mc, ws, stream);
    if (kind == LDPC_KERNEL_QC_JIT) {
        const int rc = qc_jit_decode(g, frames, max_iter, flags, llr_dev, z_dev, zbits_dev, conv_dev, ok_dev, post_dev,
                                     mc, ws, stream);
        if (rc != LDPC_ERR_UNSUPPORTED) return rc;
        // The specialisation failed (no NVRTC, compile error; remembered per code by qc_jit.cu): downgrade the memoised
        // kernel family of this handle, so that ldpc_workspace_bytes_ex / ldpc_graph_prepare / later calls see what will
        // really run -- the table-driven kernel where it has a shape for the code, else the generic kernels.
        const bool table = qc_resident_supported(g);
        g->kind_cache[0].store(table ? LDPC_KERNEL_QC_TABLE : LDPC_KERNEL_GENERIC, std::memory_order_relaxed);
        if (!table) return rc;             // decode_device falls through to the generic kernels
    }
    Outputs out{z_dev, zbits_dev, conv_dev, ok_dev, post_dev};
    switch (pick_shape(g)) {
        case 0: return launch_shape<12, 7>(g, frames, max_iter, flags, llr_dev, out, mc, ws, stream);
        case 1: return launch_shape<6, 15>(g, frames, max_iter, flags, llr_dev, out, mc, ws, stream);
        case 2: return launch_shape<4, 22>(g, frames, max_iter, flags, llr_dev, out, mc, ws, stream);
        default: break;
    }
    set_error("no resident kernel shape for this graph");
    return LDPC_ERR_UNSUPPORTED;
}

int qc_resident_kind(const ldpc_graph* g, unsigned flags)
{
    if (!g || !g->is_qc) return LDPC_KERNEL_GENERIC;
    std::atomic<int>& memo = g->kind_cache[((flags & LDPC_FLAG_TABLE_KERNEL) ? 1 : 0) | ((flags & LDPC_FLAG_NO_JIT) ? 2 : 0)];
    int kind = memo.load(std::memory_order_relaxed);
    if (kind >= 0) return kind;                       // asked on every decode call: decided once per handle
    const bool table = qc_resident_supported(g);
    if (flags & LDPC_FLAG_TABLE_KERNEL) kind = table ? LDPC_KERNEL_QC_TABLE : LDPC_KERNEL_GENERIC;
    else if (qc_spec_find(g) >= 0) kind = LDPC_KERNEL_QC_REGISTERED;
    else if (!(flags & LDPC_FLAG_NO_JIT) && qc_jit_supported(g)) kind = LDPC_KERNEL_QC_JIT;
    else kind = table ? LDPC_KERNEL_QC_TABLE : LDPC_KERNEL_GENERIC;
    memo.store(kind, std::memory_order_relaxed);
    return kind;
}

void qc_resident_release(ldpc_graph* g)
{
    if (g && g->d_qc_tables) { cudaFree(g->d_qc_tables); g->d_qc_tables = nullptr; }
}

}  // namespace ldpc
