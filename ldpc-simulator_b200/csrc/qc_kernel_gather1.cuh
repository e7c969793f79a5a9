// qc_kernel_gather1.cuh -- ONE frame per thread, barrier-free check-node phase + gather, messages and channel values in
// tensor memory: the structure of qc_kernel_gather.cuh with scalar arithmetic (row_front / row_back of qc_kernel.cuh).
//
// Why: the pair kernel needs 165 registers, so only two CTAs (12 warps) fit an SM, and three in-order warps per
// scheduler overlap MUFU and FMA work poorly (profiles/r2_tuning.md).  With the messages in tensor memory and one frame
// per thread the working set of a check row is ~70 registers: four CTAs = 24 warps per SM (6 per scheduler), at the
// price of more instructions per edge and frame.  Shared memory per CTA: posterior n + edge buffer E z + TMA stage n
// floats (47.6 KB for WiMAX-2304 r1/2); tensor memory per thread: GROUPS rows of 8 columns + 16 columns of channel
// values (128 columns per CTA, four CTAs fill the SM's 512).  Bit-identical to every other resident variant (tested).
#pragma once
#include "qc_kernel_gather.cuh"

namespace ldpc {
namespace qc {

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]) :: "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8])
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}

template <class C>
struct Gather1Shape {
    static constexpr int TZ = (C::Z + 31) / 32 * 32;
    static constexpr int THREADS = TZ * C::TEAMS;
    static constexpr int WARPS = THREADS / 32;
    static constexpr int NB = C::N / C::Z;
    static constexpr int OWN = (NB + C::TEAMS - 1) / C::TEAMS;
    static constexpr int ROWS = C::GROUPS;
    static constexpr int ROW_COLS = 8;
    static constexpr int CH_COL = ROWS * ROW_COLS;
    static constexpr int STACK = CH_COL + 16;
    static constexpr int NEED = (WARPS + 3) / 4 * STACK;
    static constexpr int COLS = NEED <= 32 ? 32 : NEED <= 64 ? 64 : NEED <= 128 ? 128 : NEED <= 256 ? 256 : 512;
    static constexpr int MAX_CTAS = 512 / COLS;
    static constexpr int EDGES = GatherShape<C>::EDGES;
    static constexpr bool FITS = NEED <= 512 && C::MAXDEG <= 8 && OWN <= 16;
    static constexpr size_t SMEM = sizeof(float) * ((size_t)C::N + (size_t)EDGES * C::Z + (size_t)C::N);
    static constexpr int B0 = 65536 / (THREADS * 80);
    static constexpr int B1 = B0 < 1 ? 1 : B0;
    static constexpr int MINB = B1 > MAX_CTAS ? MAX_CTAS : B1;
};

// one block row of the check-node phase: old messages from tensor memory, new ones to tensor memory and the edge buffer
template <int Z, int TEAM, int TBASE, int EOFF, int ROW, bool EARLY, class G0, class... Rest>
__device__ __forceinline__ void cn1_rows(const uint32_t taddr0, const float* __restrict__ post, float* __restrict__ ebuf, const int r,
                                         const bool fix_odd, const bool act, bool& unsat, G0, Rest... rest)
{
    using R = typename TeamRow<TEAM, G0>::type;
    constexpr int D = R::D;
    if constexpr (D > 0) {
        float m[D];
        uint32_t raw[8];
        tmem_ld8(taddr0 + ROW * 8, raw);
#pragma unroll
        for (int k = 0; k < D; ++k) m[k] = __uint_as_float(raw[k]);
        RowFront<D> f;
        row_front<Z, 0, EARLY>(R(), m, post, r, 0, fix_odd, act, unsat, f);
        row_back<Z, 0>(R(), m, f);
        if (act) {
#pragma unroll
            for (int k = 0; k < D; ++k) ebuf[(TBASE + EOFF + k) * Z + r] = m[k];
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) raw[k] = k < D ? __float_as_uint(m[k]) : 0u;
        tmem_st8(taddr0 + ROW * 8, raw);
    }
    if constexpr (sizeof...(Rest) > 0)
        cn1_rows<Z, TEAM, TBASE, EOFF + D, ROW + (D > 0 ? 1 : 0), EARLY>(taddr0, post, ebuf, r, fix_odd, act, unsat, rest...);
}

template <int Z, int CB, class... S>
__device__ __forceinline__ void gather1_row(Row<S...>, float& acc, const float* __restrict__ ebuf, const int t, const int slot0)
{
    constexpr int D = sizeof...(S);
    if constexpr (D == 0) return;
    constexpr int DD = D > 0 ? D : 1;
    constexpr int COLB[DD] = {S::colb...};
    constexpr int SH[DD] = {S::shift...};
#pragma unroll
    for (int k = 0; k < D; ++k) {
        if (COLB[k] == CB) {
            int idx = t + (Z - SH[k]);
            idx = (int)min((unsigned)idx, (unsigned)(idx - Z));
            acc = __fadd_rn(acc, ebuf[(slot0 + k) * Z + idx]);
        }
    }
}

template <class C, int CB, int B0, int B1, int B2, int B3>
__device__ __forceinline__ void gather1_groups(float&, const float* __restrict__, const int) {}

template <class C, int CB, int B0, int B1, int B2, int B3, class G0, class... Rest>
__device__ __forceinline__ void gather1_groups(float& acc, const float* __restrict__ ebuf, const int t, G0, Rest... rest)
{
    using R0 = typename TeamRow<0, G0>::type;
    using R1 = typename TeamRow<1, G0>::type;
    using R2 = typename TeamRow<2, G0>::type;
    using R3 = typename TeamRow<3, G0>::type;
    gather1_row<C::Z, CB>(R0(), acc, ebuf, t, TeamBase<C, 0>::value + B0);
    if constexpr (C::TEAMS > 1) gather1_row<C::Z, CB>(R1(), acc, ebuf, t, TeamBase<C, 1>::value + B1);
    if constexpr (C::TEAMS > 2) gather1_row<C::Z, CB>(R2(), acc, ebuf, t, TeamBase<C, 2>::value + B2);
    if constexpr (C::TEAMS > 3) gather1_row<C::Z, CB>(R3(), acc, ebuf, t, TeamBase<C, 3>::value + B3);
    gather1_groups<C, CB, B0 + R0::D, B1 + R1::D, B2 + R2::D, B3 + R3::D>(acc, ebuf, t, rest...);
}

template <class C, int TEAM, int K, class... G>
__device__ __forceinline__ void vn1_columns(const uint32_t (&ch)[16], float* __restrict__ post, const float* __restrict__ ebuf, const int t)
{
    constexpr int NB = C::N / C::Z;
    constexpr int CB = K * C::TEAMS + TEAM;
    if constexpr (CB < NB) {
        float acc = __uint_as_float(ch[K]);
        gather1_groups<C, CB, 0, 0, 0, 0>(acc, ebuf, t, G()...);
        post[CB * C::Z + t] = acc;
        vn1_columns<C, TEAM, K + 1, G...>(ch, post, ebuf, t);
    }
}

template <int THREADS, bool EARLY, int Z, int N, class... G>
__device__ __forceinline__ void decode_gather1(Code<Z, N, G...>, const float* __restrict__ llr, const Outputs& out,
                                               long long frames, int max_iter, int fix_odd, const McParams& mc,
                                               unsigned long long* __restrict__ work_counter)
{
    using C = Code<Z, N, G...>;
    using SH = Gather1Shape<C>;
    constexpr int TZ = SH::TZ;
    static_assert(THREADS == TZ * C::TEAMS, "CTA size = teams x ceil32(z)");
    extern __shared__ __align__(16) float sm[];
    float* const post = sm;                                   // [N]
    float* const ebuf = sm + N;                               // [EDGES][Z]
    float* const stage_f = sm + N + SH::EDGES * Z;            // [N]
    __shared__ long long s_frame;
    __shared__ unsigned long long s_cnt[5];
    __shared__ int s_err;
    __shared__ __align__(8) unsigned long long s_tma_bar;
    __shared__ uint32_t s_tmem;

    const int team = (C::TEAMS == 1) ? 0 : (int)(threadIdx.x / TZ);
    const int r = (C::TEAMS == 1) ? (int)threadIdx.x : (int)(threadIdx.x - team * TZ);
    const bool row_ok = (Z == TZ) ? true : (r < Z);
    if (threadIdx.x < 5) s_cnt[threadIdx.x] = 0;
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     :: "r"((uint32_t)__cvta_generic_to_shared(&s_tmem)), "n"(SH::COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    const uint32_t tma_bar = (uint32_t)__cvta_generic_to_shared(&s_tma_bar);
    if (threadIdx.x == 0) {
        mbar_init(tma_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    uint32_t tma_phase = 0;
    const bool use_tma = !mc.active && (N % 4 == 0) && ((reinterpret_cast<uintptr_t>(llr) & 15) == 0);
    const uint32_t stage = (uint32_t)__cvta_generic_to_shared(stage_f);
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t warp = threadIdx.x >> 5;
    const uint32_t taddr0 = s_tmem + (((warp & 3u) * 32u) << 16) + (warp >> 2) * (uint32_t)SH::STACK;
    const uint32_t ch_addr = taddr0 + SH::CH_COL;

    ChannelConst cc;
    cc.noise_dev = mc.noise_dev; cc.llr_scale = mc.llr_scale; cc.amp = mc.amp;
    cc.a2 = mc.a2; cc.l_hit = mc.l_hit; cc.hit_threshold = mc.hit_threshold;
    cc.k0 = (uint32_t)mc.seed; cc.k1 = (uint32_t)(mc.seed >> 32); cc.stream_id = mc.stream_id;

    long long f = blockIdx.x;
    if (EARLY) {
        if (threadIdx.x == 0) s_frame = (long long)atomicAdd(work_counter, 1ull);
        __syncthreads();
        f = s_frame;
    }
    if (use_tma && threadIdx.x == 0 && f < frames) tma_load_row(stage, llr + (size_t)f * N, N * 4, tma_bar);
    while (f < frames) {
        long long f_next = f + gridDim.x;
        if (EARLY) {
            __syncthreads();
            if (threadIdx.x == 0) s_frame = (long long)atomicAdd(work_counter, 1ull);
            __syncthreads();
            f_next = s_frame;
        }
        // ---- prologue: the raw LLR row into the stage buffer ----
        if (mc.active) {
            for (int q = threadIdx.x; q < (N + 3) / 4; q += THREADS) {
                uint32_t bits = 0;
                if (mc.codeword) {
                    const uint8_t* cw = mc.codeword + f * mc.codeword_stride;
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        if (4 * q + i < N && cw[4 * q + i]) bits |= 1u << i;
                }
                float v[4];
                channel_llr4(cc, mc.frame_offset + (uint64_t)f, (uint32_t)q, bits, v);
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    if (4 * q + i < N) stage_f[4 * q + i] = v[i];
            }
            __syncthreads();
        } else if (use_tma) {
            mbar_wait(tma_bar, tma_phase & 1u);
            ++tma_phase;
        } else {
            const float* src = llr + (size_t)f * N;
            for (int j = threadIdx.x; j < N; j += THREADS) stage_f[j] = __ldg(src + j);
            __syncthreads();
        }
        {
            uint32_t chv[16], z8[8];
#pragma unroll
            for (int q = 0; q < 16; ++q) chv[q] = 0u;
#pragma unroll
            for (int q = 0; q < 8; ++q) z8[q] = 0u;
            if (row_ok) {
#pragma unroll
                for (int k = 0; k < SH::OWN; ++k) {
                    const int cb = k * C::TEAMS + team;
                    if (cb < SH::NB) {
                        const int j = cb * Z + r;
                        const float v = stage_f[j] * kLog2e;
                        post[j] = v;
                        chv[k] = __float_as_uint(v);
                    }
                }
            }
            tmem_wait_st();
            tmem_st16(ch_addr, chv);
#pragma unroll
            for (int row = 0; row < SH::ROWS; ++row) tmem_st8(taddr0 + row * 8, z8);      // messages of "pass -1"
            tmem_wait_st();
        }
        __syncthreads();
        if (use_tma && threadIdx.x == 0 && f_next < frames) tma_load_row(stage, llr + (size_t)f_next * N, N * 4, tma_bar);

        int conv = -1;
        for (int it = 0; it < max_iter; ++it) {
            bool unsat = false;
            tmem_wait_st();
            if (team == 0) cn1_rows<Z, 0, TeamBase<C, 0>::value, 0, 0, EARLY>(taddr0, post, ebuf, r, fix_odd != 0, row_ok, unsat, G()...);
            if constexpr (C::TEAMS > 1) { if (team == 1) cn1_rows<Z, 1, TeamBase<C, 1>::value, 0, 0, EARLY>(taddr0, post, ebuf, r, fix_odd != 0, row_ok, unsat, G()...); }
            if constexpr (C::TEAMS > 2) { if (team == 2) cn1_rows<Z, 2, TeamBase<C, 2>::value, 0, 0, EARLY>(taddr0, post, ebuf, r, fix_odd != 0, row_ok, unsat, G()...); }
            if constexpr (C::TEAMS > 3) { if (team == 3) cn1_rows<Z, 3, TeamBase<C, 3>::value, 0, 0, EARLY>(taddr0, post, ebuf, r, fix_odd != 0, row_ok, unsat, G()...); }
            if (!row_ok) unsat = false;
            bool nearly = it == 0;               // (see decode_frames in qc_kernel.cuh: the eager syndrome test)
            if (EARLY && it > 0) {
                const int open_rows = __syncthreads_count(unsat);
                if (open_rows == 0) { conv = it - 1; break; }
                nearly = open_rows <= kEagerSyndromeThreads;
            } else {
                __syncthreads();
            }
            uint32_t ch[16];
            tmem_ld16(ch_addr, ch);
            tmem_wait_ld16(ch);
            if (row_ok) {
                if (team == 0) vn1_columns<C, 0, 0, G...>(ch, post, ebuf, r);
                if constexpr (C::TEAMS > 1) { if (team == 1) vn1_columns<C, 1, 0, G...>(ch, post, ebuf, r); }
                if constexpr (C::TEAMS > 2) { if (team == 2) vn1_columns<C, 2, 0, G...>(ch, post, ebuf, r); }
                if constexpr (C::TEAMS > 3) { if (team == 3) vn1_columns<C, 3, 0, G...>(ch, post, ebuf, r); }
            }
            __syncthreads();
            if (EARLY && nearly && it + 1 < max_iter) {
                // the previous pass left few open checks (or this is the first pass): test this pass' posterior at once
                // instead of finding out during the next pass
                bool u = false;
                if (row_ok) {
                    if (team == 0) u = team_unsat<Z, 0>(post, r, 0, G()...);
                    if constexpr (C::TEAMS > 1) { if (team == 1) u = team_unsat<Z, 1>(post, r, 0, G()...); }
                    if constexpr (C::TEAMS > 2) { if (team == 2) u = team_unsat<Z, 2>(post, r, 0, G()...); }
                    if constexpr (C::TEAMS > 3) { if (team == 3) u = team_unsat<Z, 3>(post, r, 0, G()...); }
                }
                if (!__syncthreads_or(u)) { conv = it; break; }
            }
        }
        if (conv < 0) {
            bool unsat = false;
            if (row_ok) {
                if (team == 0) unsat = team_unsat<Z, 0>(post, r, 0, G()...);
                if constexpr (C::TEAMS > 1) { if (team == 1) unsat = team_unsat<Z, 1>(post, r, 0, G()...); }
                if constexpr (C::TEAMS > 2) { if (team == 2) unsat = team_unsat<Z, 2>(post, r, 0, G()...); }
                if constexpr (C::TEAMS > 3) { if (team == 3) unsat = team_unsat<Z, 3>(post, r, 0, G()...); }
            }
            if (!__syncthreads_or(unsat)) conv = max_iter - 1;
        }
        const float* fin = post;
        const bool good = conv >= 0;
        if (threadIdx.x == 0) {
            if (out.conv_it) out.conv_it[f] = conv;
            if (out.ok) out.ok[f] = good ? 1 : 0;
        }
        if (out.zbits) {
            constexpr int words = (N + 31) / 32;
            for (int base = 0; base < words * 32; base += THREADS) {
                const int j = base + threadIdx.x;
                const bool neg = (j < N) && (fin[j] < 0.f);
                const unsigned mm = __ballot_sync(0xffffffffu, neg);
                if ((threadIdx.x & 31) == 0 && j < words * 32) out.zbits[(size_t)f * words + (j >> 5)] = mm;
            }
        }
        if (out.z) {
            uint8_t* dst = out.z + (size_t)f * N;
            for (int j = threadIdx.x; j < N; j += THREADS) dst[j] = (uint8_t)(fin[j] < 0.f);
        }
        if (out.post) {
            float* dst = out.post + (size_t)f * N;
            for (int j = threadIdx.x; j < N; j += THREADS) dst[j] = fin[j] * kLn2;
        }
        if (mc.active) {
            if (threadIdx.x == 0) s_err = 0;
            __syncthreads();
            if (!good) {
                int errs = 0;
                const int span = mc.info_mask ? N : mc.k_info;
                for (int j = threadIdx.x; j < span; j += THREADS) {
                    if (mc.info_mask && !mc.info_mask[j]) continue;
                    const unsigned est = (fin[j] < 0.f) ? 0u : 1u;
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
        __syncthreads();
        f = f_next;
    }
    if (mc.active && mc.counters) {
        __syncthreads();
        if (threadIdx.x < 5 && s_cnt[threadIdx.x]) atomicAdd(&mc.counters[threadIdx.x], s_cnt[threadIdx.x]);
    }
    tmem_wait_st();
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(s_tmem), "n"(SH::COLS) : "memory");
}

template <int THREADS, int MINB, bool EARLY, class C>
__global__ void __launch_bounds__(THREADS, MINB)
k_qc_gather1(const float* __restrict__ llr, Outputs out, long long frames, int max_iter, int fix_odd, McParams mc,
             unsigned long long* __restrict__ work_counter)
{
    decode_gather1<THREADS, EARLY>(C(), llr, out, frames, max_iter, fix_odd, mc, work_counter);
}

}  // namespace qc
}  // namespace ldpc
