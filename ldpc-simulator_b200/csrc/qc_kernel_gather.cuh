// qc_kernel_gather.cuh -- two frames per thread, BARRIER-FREE check-node phase.
//
// The pair kernel of qc_kernel_pair.cuh accumulates the new posterior in place (read-modify-write of the
// posterior words by the check rows), which orders the block rows that share a column block: one CTA-wide
// barrier per group of rows, six per pass for WiMAX r1/2.  The ncu source view of that kernel shows the price:
// 13 % of all warp samples sit in the barrier's wait loop, another 6 % on the first posterior load behind it,
// and the warps of a CTA are phase-locked, so their MUFU bursts collide (MIO throttle 16 %) while the MUFU
// pipe idles in between (65 % busy).
//
// Here a pass is split the way the reference writes it (spa_decoder.py:112-185):
//   CN phase   every thread walks its block rows back to back WITHOUT synchronisation: previous posterior
//              (LDS.64, rotated) - own previous message (LDS.64 of the thread's own word of the edge buffer: the
//              variable-node phase only reads it, so it still holds the messages of the previous pass) ->
//              likelihood-ratio check node -> new messages to the edge buffer E[slot][row] in shared memory
//              (STS.64, the thread's own word: no address arithmetic, no conflict);
//   barrier
//   VN phase   every thread owns column r of some of the block columns: posterior = channel value (tensor
//              memory) + the messages of the column read from E in schedule order (rotated index, LDS.64),
//              written to the ONE posterior buffer;
//   barrier
// Two barriers per pass instead of six, none inside the MUFU-heavy part.  The additions of a column happen in
// the same order as in the scatter kernels, so the results are bit-identical to them (tested).
//
// Shared memory per CTA (float2 units): posterior [n] | edge buffer [edges of the base matrix][z] | TMA stage
// (2 n floats); the channel values of the columns a thread owns live in tensor memory (32 columns per thread, one
// tcgen05.ld per pass).  Round 2 first kept the messages in tensor memory as well (one tcgen05.ld/st of 16 columns per
// row): reading them back from the edge buffer costs one LDS.64 per edge more but drops the LDTM/STTM, their waits
// and the register shuffling around the 16-register blocks -- 16.81 instead of 17.09 ms (profiles/r2_tuning.md).
#pragma once
#include "qc_kernel_pair.cuh"

namespace ldpc {
namespace qc {

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                   "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr));
    // valid after tcgen05.wait::ld; in/out operands keep every use below the wait (two statements: 30-operand limit)
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                   "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
                 :: "memory");
    asm volatile("" : "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]),
                      "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
                 :: "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32])
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
                 "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
                 :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
                    "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
                    "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
                    "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]) : "memory");
}

// Shapes of the gather kernel for a code.
template <class C>
struct GatherShape {
    static constexpr int TZ = (C::Z + 31) / 32 * 32;
    static constexpr int THREADS = TZ * C::TEAMS;
    static constexpr int WARPS = THREADS / 32;
    static constexpr int NB = C::N / C::Z;                                   // block columns
    static constexpr int OWN = (NB + C::TEAMS - 1) / C::TEAMS;               // block columns a thread owns in the VN phase
    static constexpr int STACK = 32;                                          // TMEM columns per thread: its channel values
    static constexpr int NEED = (WARPS + 3) / 4 * STACK;
    static constexpr int COLS = NEED <= 32 ? 32 : NEED <= 64 ? 64 : NEED <= 128 ? 128 : NEED <= 256 ? 256 : 512;
    static constexpr int MAX_CTAS = 512 / COLS;
    static constexpr int EDGES = (C::template TeamSlots<0>::value + C::template TeamSlots<1>::value +
                                  C::template TeamSlots<2>::value + C::template TeamSlots<3>::value);
    static constexpr bool FITS = NEED <= 512 && C::MAXDEG <= 8 && 2 * OWN <= 32;
    static constexpr size_t SMEM = sizeof(float2) * ((size_t)C::N + (size_t)EDGES * C::Z + (size_t)C::N);
    static constexpr int B0 = 65536 / (THREADS * 168);
    static constexpr int B1 = B0 < 1 ? 1 : B0;
    static constexpr int MINB = B1 > MAX_CTAS ? MAX_CTAS : B1;
};

// first edge-buffer slot of team T (teams are laid out one after the other)
template <class C, int T> struct TeamBase { static constexpr int value = TeamBase<C, T - 1>::value + C::template TeamSlots<T - 1>::value; };
template <class C> struct TeamBase<C, 0> { static constexpr int value = 0; };

// ---- CN phase: the rows of a team back to back (the overlap of consecutive rows is left to ptxas, see cn_pipe) ----
// A row's messages of the PREVIOUS pass: the edge buffer still holds them (the variable-node phase only reads it), and the
// words are the thread's own -- one LDS.64 per edge at an immediate offset, no tensor-memory round trip, no store.
template <int Z, class... S>
__device__ __forceinline__ void row_prev(Row<S...>, RowMsg<sizeof...(S)>& m, const float2* __restrict__ ebuf, const int r, const int slot0)
{
    constexpr int D = sizeof...(S);
#pragma unroll
    for (int k = 0; k < D; ++k) m.v[k] = sh_ld2(ebuf, (slot0 + k) * Z + r);
}

template <int Z, class... S>
__device__ __forceinline__ void row_emit(Row<S...>, const RowMsg<sizeof...(S)>& m, float2* __restrict__ ebuf, const int r,
                                         const int slot0, const bool act)
{
    constexpr int D = sizeof...(S);
    if constexpr (D == 0) return;
    if (!act) return;
#ifndef LDPC_EXP_NOSMEM
#pragma unroll
    for (int k = 0; k < D; ++k) sh_st2(ebuf, (slot0 + k) * Z + r, m.v[k]);
#endif
}

template <int Z, int TEAM, int TBASE, int EOFF, bool EARLY, class GCUR>
__device__ __forceinline__ void cn_pipe(const float2* __restrict__ post, float2* __restrict__ ebuf, const int r,
                                        const bool fix_odd, const bool act, bool& ua, bool& ub,
                                        RowMsg<TeamRow<TEAM, GCUR>::type::D>& mcur,
                                        const RowFront2<TeamRow<TEAM, GCUR>::type::D>& fcur, GCUR)
{
    using R = typename TeamRow<TEAM, GCUR>::type;
    row_back2(R(), mcur, fcur);
    row_emit<Z>(R(), mcur, ebuf, r, TBASE + EOFF, act);
}

template <int Z, int TEAM, int TBASE, int EOFF, bool EARLY, class GCUR, class GNEXT, class... Rest>
__device__ __forceinline__ void cn_pipe(const float2* __restrict__ post, float2* __restrict__ ebuf, const int r,
                                        const bool fix_odd, const bool act, bool& ua, bool& ub,
                                        RowMsg<TeamRow<TEAM, GCUR>::type::D>& mcur,
                                        const RowFront2<TeamRow<TEAM, GCUR>::type::D>& fcur, GCUR, GNEXT gn, Rest... rest)
{
    using R = typename TeamRow<TEAM, GCUR>::type;
    using RN = typename TeamRow<TEAM, GNEXT>::type;
    // Source order: the whole current row first, then the loads and the front half of the next one.  ptxas interleaves the
    // two rows on its own, and how well depends on this order (one call, ms per launch: front of the next row before the
    // back of the current one 16.75, back / front / emit 17.58, back / emit / front 16.40, this order 16.37; the front
    // half split into its loads before and its exponentials after the current row 17.12 -- profiles/r2_tuning.md 5.10).
    row_back2(R(), mcur, fcur);
    row_emit<Z>(R(), mcur, ebuf, r, TBASE + EOFF, act);
    RowMsg<RN::D> mnext;
    row_prev<Z>(RN(), mnext, ebuf, r, TBASE + EOFF + R::D);
    RowFront2<RN::D> fnext;
    row_front2<Z, EARLY>(RN(), mnext, post, r, 0, fix_odd, act, ua, ub, fnext);
    cn_pipe<Z, TEAM, TBASE, EOFF + R::D, EARLY>(post, ebuf, r, fix_odd, act, ua, ub, mnext, fnext, gn, rest...);
}

template <int Z, int TEAM, int TBASE, bool EARLY, class G0, class... Rest>
__device__ __forceinline__ void cn_phase(const float2* __restrict__ post, float2* __restrict__ ebuf, const int r,
                                         const bool fix_odd, const bool act, bool& ua, bool& ub,
                                         G0 g0, Rest... rest)
{
    using R0 = typename TeamRow<TEAM, G0>::type;
    RowMsg<R0::D> m0;
    row_prev<Z>(R0(), m0, ebuf, r, TBASE);
    RowFront2<R0::D> f0;
    row_front2<Z, EARLY>(R0(), m0, post, r, 0, fix_odd, act, ua, ub, f0);
    cn_pipe<Z, TEAM, TBASE, 0, EARLY>(post, ebuf, r, fix_odd, act, ua, ub, m0, f0, g0, rest...);
}

// ---- VN phase: posterior of column t of block column CB = channel value + its messages in schedule order ----
template <int Z, int CB, class... S>
__device__ __forceinline__ void gather_row(Row<S...>, float2& acc, const float2* __restrict__ ebuf, const int t, const int slot0)
{
    constexpr int D = sizeof...(S);
    if constexpr (D == 0) return;
    constexpr int DD = D > 0 ? D : 1;
    constexpr int COLB[DD] = {S::colb...};
    constexpr int SH[DD] = {S::shift...};
#pragma unroll
    for (int k = 0; k < D; ++k) {
        if (COLB[k] == CB) {                                     // resolved at compile time
            int idx = t + (Z - SH[k]);                           // check row that owns this column: (t - shift) mod z
            idx = (int)min((unsigned)idx, (unsigned)(idx - Z));
            acc = f2add(acc, sh_ld2(ebuf, (slot0 + k) * Z + idx));
        }
    }
}

// the running message offsets of the teams (B0..B3) advance group by group
template <class C, int CB, int B0, int B1, int B2, int B3>
__device__ __forceinline__ void gather_groups(float2&, const float2* __restrict__, const int) {}

template <class C, int CB, int B0, int B1, int B2, int B3, class G0, class... Rest>
__device__ __forceinline__ void gather_groups(float2& acc, const float2* __restrict__ ebuf, const int t, G0, Rest... rest)
{
    using R0 = typename TeamRow<0, G0>::type;
    using R1 = typename TeamRow<1, G0>::type;
    using R2 = typename TeamRow<2, G0>::type;
    using R3 = typename TeamRow<3, G0>::type;
    gather_row<C::Z, CB>(R0(), acc, ebuf, t, TeamBase<C, 0>::value + B0);
    if constexpr (C::TEAMS > 1) gather_row<C::Z, CB>(R1(), acc, ebuf, t, TeamBase<C, 1>::value + B1);
    if constexpr (C::TEAMS > 2) gather_row<C::Z, CB>(R2(), acc, ebuf, t, TeamBase<C, 2>::value + B2);
    if constexpr (C::TEAMS > 3) gather_row<C::Z, CB>(R3(), acc, ebuf, t, TeamBase<C, 3>::value + B3);
    gather_groups<C, CB, B0 + R0::D, B1 + R1::D, B2 + R2::D, B3 + R3::D>(acc, ebuf, t, rest...);
}

template <class C, int TEAM, int K, class... G>
__device__ __forceinline__ void vn_columns(const uint32_t (&ch)[32], float2* __restrict__ post, const float2* __restrict__ ebuf, const int t)
{
    constexpr int NB = C::N / C::Z;
    constexpr int CB = K * C::TEAMS + TEAM;                      // block columns are dealt to the teams round robin
    if constexpr (CB < NB) {
        float2 acc = f2(__uint_as_float(ch[2 * K]), __uint_as_float(ch[2 * K + 1]));
        gather_groups<C, CB, 0, 0, 0, 0>(acc, ebuf, t, G()...);
        sh_st2(post, CB * C::Z + t, acc);
        vn_columns<C, TEAM, K + 1, G...>(ch, post, ebuf, t);
    }
}

template <int THREADS, bool EARLY, int Z, int N, class... G>
__device__ __forceinline__ void decode_gather(Code<Z, N, G...>, const float* __restrict__ llr, const Outputs& out,
                                              long long frames, int max_iter, int fix_odd, const McParams& mc,
                                              unsigned long long* __restrict__ work_counter)
{
    using C = Code<Z, N, G...>;
    using SH = GatherShape<C>;
    constexpr int TZ = SH::TZ;
    static_assert(THREADS == TZ * C::TEAMS, "CTA size = teams x ceil32(z)");
    extern __shared__ __align__(16) float2 sm2[];
    float2* const post = sm2;                                 // [N]
    float2* const ebuf = sm2 + N;                             // [EDGES][Z]
    float* const stage_f = reinterpret_cast<float*>(sm2 + N + SH::EDGES * Z);    // 2 N floats
    __shared__ long long s_pair;
    __shared__ unsigned long long s_cnt[5];
    __shared__ int s_err[2];
    __shared__ unsigned s_or[3];
    __shared__ __align__(8) unsigned long long s_tma_bar;
    __shared__ uint32_t s_tmem;

    const int team = (C::TEAMS == 1) ? 0 : (int)(threadIdx.x / TZ);     // warp-uniform
    const int r = (C::TEAMS == 1) ? (int)threadIdx.x : (int)(threadIdx.x - team * TZ);
    const bool row_ok = (Z == TZ) ? true : (r < Z);
    if (threadIdx.x < 5) s_cnt[threadIdx.x] = 0;
    if (threadIdx.x < 3) s_or[threadIdx.x] = 0;
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
    unsigned or_use = 0;
    const long long pairs = (frames + 1) / 2;
    const bool use_tma = !mc.active && (N % 4 == 0) && ((reinterpret_cast<uintptr_t>(llr) & 15) == 0);
    const uint32_t stage = (uint32_t)__cvta_generic_to_shared(stage_f);
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // a thread owns the TMEM lane of its warp quarter (.32x32b: lane i of the warp <-> TMEM lane 32 * (warp % 4) + i)
    const uint32_t ch_addr = s_tmem + ((((threadIdx.x >> 5) & 3u) * 32u) << 16) + (threadIdx.x >> 7) * (uint32_t)SH::STACK;

    ChannelConst cc;
    cc.noise_dev = mc.noise_dev; cc.llr_scale = mc.llr_scale; cc.amp = mc.amp;
    cc.a2 = mc.a2; cc.l_hit = mc.l_hit; cc.hit_threshold = mc.hit_threshold;
    cc.k0 = (uint32_t)mc.seed; cc.k1 = (uint32_t)(mc.seed >> 32); cc.stream_id = mc.stream_id;

    auto rows_of = [&](long long p) -> uint32_t { return (2 * p + 1 < frames) ? 2u : 1u; };

    long long p = blockIdx.x;
    if (EARLY) {
        if (threadIdx.x == 0) s_pair = (long long)atomicAdd(work_counter, 1ull);
        __syncthreads();
        p = s_pair;
    }
    if (use_tma && threadIdx.x == 0 && p < pairs) tma_load_row(stage, llr + (size_t)(2 * p) * N, rows_of(p) * N * 4, tma_bar);
    while (p < pairs) {
        long long p_next = p + gridDim.x;
        if (EARLY) {
            __syncthreads();
            if (threadIdx.x == 0) s_pair = (long long)atomicAdd(work_counter, 1ull);
            __syncthreads();
            p_next = s_pair;
        }
        const long long fa = 2 * p, fb = 2 * p + 1;
        const bool has_b = fb < frames;

        // ---- prologue: raw LLR rows of both frames into the stage buffer (TMA / Philox / plain loads) ----
        if (mc.active) {
            for (int q = threadIdx.x; q < (N + 3) / 4; q += THREADS) {
                uint32_t bits_a = 0, bits_b = 0;
                if (mc.codeword) {
                    const uint8_t* cwa = mc.codeword + fa * mc.codeword_stride;
                    const uint8_t* cwb = mc.codeword + (has_b ? fb : fa) * mc.codeword_stride;
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        if (4 * q + i < N) {
                            if (cwa[4 * q + i]) bits_a |= 1u << i;
                            if (cwb[4 * q + i]) bits_b |= 1u << i;
                        }
                }
                float va[4], vb[4];
                channel_llr4(cc, mc.frame_offset + (uint64_t)fa, (uint32_t)q, bits_a, va);
                channel_llr4(cc, mc.frame_offset + (uint64_t)(has_b ? fb : fa), (uint32_t)q, bits_b, vb);
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    if (4 * q + i < N) { stage_f[4 * q + i] = va[i]; stage_f[N + 4 * q + i] = vb[i]; }
            }
            __syncthreads();
        } else if (use_tma) {
            mbar_wait(tma_bar, tma_phase & 1u);          // this pair's rows have landed in the stage buffer
            ++tma_phase;
        } else {
            const float* sa_ = llr + (size_t)fa * N;
            const float* sb_ = llr + (size_t)(has_b ? fb : fa) * N;
            for (int j = threadIdx.x; j < N; j += THREADS) { stage_f[j] = __ldg(sa_ + j); stage_f[N + j] = __ldg(sb_ + j); }
            __syncthreads();
        }
        // every thread takes the channel values of the columns it owns in the VN phase: posterior of "pass -1"
        // (shared memory) and tensor memory (log2 units)
        {
            uint32_t chv[32];
#pragma unroll
            for (int q = 0; q < 32; ++q) chv[q] = 0u;
            if (row_ok) {
                const int off_b = (has_b || mc.active) ? N : 0;      // an odd last frame: lane B copies lane A
#pragma unroll
                for (int k = 0; k < SH::OWN; ++k) {
                    const int cb = k * C::TEAMS + team;
                    if (cb < SH::NB) {
                        const int j = cb * Z + r;
                        const float2 v = f2(stage_f[j] * kLog2e, stage_f[off_b + j] * kLog2e);
                        post[j] = v;
                        chv[2 * k] = __float_as_uint(v.x);
                        chv[2 * k + 1] = __float_as_uint(v.y);
                    }
                }
            }
            tmem_wait_st();
            tmem_st32(ch_addr, chv);
            if (row_ok) {                                // the messages of "pass -1" are zero
                const int tb = team == 0 ? TeamBase<C, 0>::value : team == 1 ? TeamBase<C, 1>::value : team == 2 ? TeamBase<C, 2>::value : TeamBase<C, 3>::value;
                const int ts = team == 0 ? C::template TeamSlots<0>::value : team == 1 ? C::template TeamSlots<1>::value
                             : team == 2 ? C::template TeamSlots<2>::value : C::template TeamSlots<3>::value;
                for (int s = 0; s < ts; ++s) ebuf[(tb + s) * Z + r] = f2(0.f, 0.f);
            }
        }
        __syncthreads();                                 // posterior complete, stage consumed -> prefetch the next pair
        if (use_tma && threadIdx.x == 0 && p_next < pairs)
            tma_load_row(stage, llr + (size_t)(2 * p_next) * N, rows_of(p_next) * N * 4, tma_bar);

        int conv_a = -1, conv_b = -1;
        bool done_a = false, done_b = !has_b;
        for (int it = 0; it < max_iter; ++it) {
            bool ua = false, ub = false;
            if (team == 0) cn_phase<Z, 0, TeamBase<C, 0>::value, EARLY>(post, ebuf, r, fix_odd != 0, row_ok, ua, ub, G()...);
            if constexpr (C::TEAMS > 1) { if (team == 1) cn_phase<Z, 1, TeamBase<C, 1>::value, EARLY>(post, ebuf, r, fix_odd != 0, row_ok, ua, ub, G()...); }
            if constexpr (C::TEAMS > 2) { if (team == 2) cn_phase<Z, 2, TeamBase<C, 2>::value, EARLY>(post, ebuf, r, fix_odd != 0, row_ok, ua, ub, G()...); }
            if constexpr (C::TEAMS > 3) { if (team == 3) cn_phase<Z, 3, TeamBase<C, 3>::value, EARLY>(post, ebuf, r, fix_odd != 0, row_ok, ua, ub, G()...); }
            if (!row_ok) ua = ub = false;
            if (EARLY && it > 0) {
                // the posterior of pass it-1 is still in place (the VN phase of this pass has not run): a frame whose
                // checks it satisfied exits with it
                const unsigned u = block_or2((ua ? 1u : 0u) | (ub ? 2u : 0u), s_or, or_use);
                bool stored = false;                     // (CTA-uniform: u, done_a, done_b are)
                if (!done_a && !(u & 1u)) { conv_a = it - 1; done_a = stored = true; store_frame<THREADS, N>(out, post, 0, fa, conv_a); }
                if (!done_b && !(u & 2u)) { conv_b = it - 1; done_b = stored = true; store_frame<THREADS, N>(out, post, 1, fb, conv_b); }
                if (done_a && done_b) break;
                if (stored) __syncthreads();             // the outputs were read from the posterior the VN phase overwrites
            } else {
#ifndef LDPC_EXP_NOBAR
                __syncthreads();                         // every message of this pass is in the edge buffer
#endif
            }
            uint32_t ch[32];
            tmem_wait_st();
            tmem_ld32(ch_addr, ch);                      // (warp collective: also the lanes without a row)
#if defined(LDPC_EXP_NOVN) || defined(LDPC_EXP_NOSMEM)
            if (row_ok && max_iter > 1000) {
#else
            if (row_ok) {
#endif
                if (team == 0) vn_columns<C, 0, 0, G...>(ch, post, ebuf, r);
                if constexpr (C::TEAMS > 1) { if (team == 1) vn_columns<C, 1, 0, G...>(ch, post, ebuf, r); }
                if constexpr (C::TEAMS > 2) { if (team == 2) vn_columns<C, 2, 0, G...>(ch, post, ebuf, r); }
                if constexpr (C::TEAMS > 3) { if (team == 3) vn_columns<C, 3, 0, G...>(ch, post, ebuf, r); }
            }
#ifndef LDPC_EXP_NOBAR
            __syncthreads();                             // the new posterior is complete
#endif
        }

        // ---- exit: syndrome of the last posterior for the frames that have not converged yet ----
        if (!(done_a && done_b)) {
            bool ua = false, ub = false;
            if (row_ok) {
                if (team == 0) team_unsat2<Z, 0>(post, r, 0, ua, ub, G()...);
                if constexpr (C::TEAMS > 1) { if (team == 1) team_unsat2<Z, 1>(post, r, 0, ua, ub, G()...); }
                if constexpr (C::TEAMS > 2) { if (team == 2) team_unsat2<Z, 2>(post, r, 0, ua, ub, G()...); }
                if constexpr (C::TEAMS > 3) { if (team == 3) team_unsat2<Z, 3>(post, r, 0, ua, ub, G()...); }
            }
            const unsigned u = block_or2((ua ? 1u : 0u) | (ub ? 2u : 0u), s_or, or_use);
            if (!done_a) { if (!(u & 1u)) conv_a = max_iter - 1; store_frame<THREADS, N>(out, post, 0, fa, conv_a); }
            if (!done_b) { if (!(u & 2u)) conv_b = max_iter - 1; store_frame<THREADS, N>(out, post, 1, fb, conv_b); }
        }

        if (mc.active) {
            // main.py:314-339: bit errors only in failed frames, on the un-complemented output.  A failed frame ran
            // every pass, so its exit posterior is the one in place.
            if (threadIdx.x < 2) s_err[threadIdx.x] = 0;
            __syncthreads();
            const bool bad_a = conv_a < 0, bad_b = has_b && conv_b < 0;
            if (bad_a || bad_b) {
                int ea = 0, eb = 0;
                const int span = mc.info_mask ? N : mc.k_info;
                for (int j = threadIdx.x; j < span; j += THREADS) {
                    if (mc.info_mask && !mc.info_mask[j]) continue;
                    const float2 L = post[j];
                    const unsigned sent_a = mc.codeword ? mc.codeword[fa * mc.codeword_stride + j] : 0u;
                    const unsigned sent_b = (mc.codeword && has_b) ? mc.codeword[fb * mc.codeword_stride + j] : 0u;
                    ea += (((L.x < 0.f) ? 0u : 1u) != sent_a);
                    eb += (((L.y < 0.f) ? 0u : 1u) != sent_b);
                }
#pragma unroll
                for (int o = 16; o; o >>= 1) { ea += __shfl_xor_sync(0xffffffffu, ea, o); eb += __shfl_xor_sync(0xffffffffu, eb, o); }
                if ((threadIdx.x & 31) == 0) {
                    if (bad_a && ea) atomicAdd(&s_err[0], ea);
                    if (bad_b && eb) atomicAdd(&s_err[1], eb);
                }
            }
            __syncthreads();
            if (threadIdx.x == 0) {
                s_cnt[0] += has_b ? 2 : 1;
                if (bad_a) { s_cnt[1] += 1; s_cnt[2] += (unsigned)s_err[0]; }
                else { s_cnt[3] += (unsigned)conv_a; s_cnt[4] += 1; }
                if (has_b) {
                    if (bad_b) { s_cnt[1] += 1; s_cnt[2] += (unsigned)s_err[1]; }
                    else { s_cnt[3] += (unsigned)conv_b; s_cnt[4] += 1; }
                }
            }
        }
        __syncthreads();     // shared buffers are reused by the next pair
        p = p_next;
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
k_qc_gather(const float* __restrict__ llr, Outputs out, long long frames, int max_iter, int fix_odd, McParams mc,
            unsigned long long* __restrict__ work_counter)
{
    decode_gather<THREADS, EARLY>(C(), llr, out, frames, max_iter, fix_odd, mc, work_counter);
}

}  // namespace qc
}  // namespace ldpc
