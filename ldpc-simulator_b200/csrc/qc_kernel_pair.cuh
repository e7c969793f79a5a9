// qc_kernel_pair.cuh -- the SM-resident quasi-cyclic SPA kernel with TWO FRAMES PER THREAD.
//
// Same algorithm, schedule (groups / teams / split-phase mbarrier) and arithmetic as qc_kernel.cuh; what
// changes is the data layout: every thread owns check row r of its block rows for a PAIR of frames, and
// everything that is per (edge, frame) is stored as a float2 {frame A, frame B}:
//   * shared memory: channel values and the two posteriors are float2[n]  -> one LDS.64 / STS.64, one
//     rotated index and one address serve both frames;
//   * registers: check->variable messages are float2 (two consecutive registers), so the
//     variable->check subtraction, the prefix / suffix products of the likelihood-ratio check node and
//     the posterior accumulation run on the packed fp32 instructions of sm_100a (FADD2 / FMUL2 / FFMA2:
//     one issue slot and one set of 64-bit register reads for two frames).
// Why: tools/pipe_probe.cu shows that the one-frame kernel is bound by instruction issue / register
// operand bandwidth next to the MUFU pipe (3-register FFMA 0.66, LOP3 0.50 instructions per clock per
// scheduler); per (edge, frame) the pair kernel issues ~16 instead of ~21 instructions, and the two
// frames of a thread are independent dependency chains, so MUFU work of one overlaps FMA work of the
// other inside a warp.  The MUFU count is unchanged (1 EX2 + 2 LG2 per edge and frame; sm_100a has no
// packed MUFU).  Results are bit-identical to the one-frame kernel: every packed instruction is the
// IEEE round-to-nearest operation of its two lanes (tested).
//
// Frame pairs: pair p = frames (2p, 2p+1); their LLR rows are adjacent in memory, so ONE bulk copy
// (TMA, cp.async.bulk) of 2n floats prefetches the next pair while the current one decodes.  An odd
// last frame runs with lane B switched off (its outputs are not written).
#pragma once
#include "qc_kernel.cuh"

namespace ldpc {
namespace qc {

__device__ __forceinline__ float2 f2(float a, float b) { return make_float2(a, b); }
__device__ __forceinline__ float2 f2neg(float2 a) { return make_float2(-a.x, -a.y); }
__device__ __forceinline__ float2 f2add(float2 a, float2 b) { return __fadd2_rn(a, b); }
// a - b.  In the NVRTC build (LDPC_SMEM_ASM) the negation of __fadd2_rn(a, -b) is not folded into the FADD2 behind the
// inline-PTX loads (two scalar FADD per edge pair): there it is spelled sub.rn.f32x2, which always is one FADD2.  The nvcc
// build keeps the intrinsic form: the PTX spelling is 24 instructions per pass shorter but schedules 2 % slower
// (16.8 -> 17.1 ms, profiles/r2_tuning.md).
#ifndef LDPC_SMEM_ASM
__device__ __forceinline__ float2 f2sub(float2 a, float2 b) { return __fadd2_rn(a, f2neg(b)); }
#else
__device__ __forceinline__ float2 f2sub(float2 a, float2 b)
{
    unsigned long long ua, ub, ur;
    float2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(ua) : "f"(a.x), "f"(a.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(ub) : "f"(b.x), "f"(b.y));
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(ur) : "l"(ua), "l"(ub));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(ur));
    return r;
}
#endif
__device__ __forceinline__ float2 f2mul(float2 a, float2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ float2 f2fma(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }

// ---- where the check->variable messages of a thread live between passes ---------------------------------
// The kernel works on one block row at a time: row_front2 turns the row's old messages into the
// variable->check messages (parked in the same D float2 registers), row_back2 overwrites them with the new
// check->variable messages, row_scatter2 adds those to the posterior.  Between passes the rows are kept by
// one of two stores:
//   RegMsgs   all rows in registers (2 x MAXSLOT registers per thread: one CTA per SM for WiMAX r1/2);
//   TmemMsgs  all rows in TENSOR MEMORY (sm_100a: 512 columns x 128 lanes x 32 bit per SM).  A thread owns the
//             TMEM lane of its warp quarter; a row is 16 (D <= 8) or 32 (D <= 16) consecutive columns, read and
//             written with ONE tcgen05.ld / tcgen05.st (.32x32b: lane i of the warp <-> TMEM lane 32*(warp%4)+i).
//             TMEM is used here as a software-managed extension of the register file -- no MMA involved; it frees
//             2 x MAXSLOT registers per thread, so twice as many CTAs (warps) fit on an SM.
template <int D> struct RowMsg { float2 v[D > 0 ? D : 1]; };

template <int NS>
struct RegMsgs {
    float2 eps[NS];
    __device__ __forceinline__ void begin_frame()
    {
#pragma unroll
        for (int s = 0; s < NS; ++s) eps[s] = f2(0.f, 0.f);
    }
    template <int ROWS> __device__ __forceinline__ void begin_frame_rows() {}
    __device__ __forceinline__ void begin_pass() const {}
    template <int EOFF, int ROW, int D>
    __device__ __forceinline__ void load(RowMsg<D>& m, bool) const
    {
#pragma unroll
        for (int k = 0; k < D; ++k) m.v[k] = eps[EOFF + k];
    }
    template <int D> __device__ __forceinline__ void load_done(RowMsg<D>&) const {}
    template <int EOFF, int ROW, int D>
    __device__ __forceinline__ void store(const RowMsg<D>& m)
    {
#pragma unroll
        for (int k = 0; k < D; ++k) eps[EOFF + k] = m.v[k];
    }
};

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16])
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
                 :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
                    "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
// The loaded registers are valid only after tcgen05.wait::ld: they are passed through the wait as in/out operands so
// that no use of them can be scheduled above it.
__device__ __forceinline__ void tmem_wait_ld16(uint32_t (&r)[16])
{
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                   "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
                 :: "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// Rows of at most 8 edges (16 columns per row).  taddr0 = TMEM address of this thread's first row
// (lane quarter of the warp in bits 31:16, column in bits 15:0).
struct TmemMsgs {
    static constexpr int ROW_COLS = 16;
    uint32_t taddr0;
    // the messages of "pass -1" are zero: written once per frame (ROWS stores), so that every pass simply loads
    template <int ROWS>
    __device__ __forceinline__ void begin_frame_rows()
    {
        uint32_t z[16];
#pragma unroll
        for (int q = 0; q < 16; ++q) z[q] = 0u;
#pragma unroll
        for (int row = 0; row < ROWS; ++row) tmem_st16(taddr0 + row * ROW_COLS, z);
        tmem_wait_st();
    }
    __device__ __forceinline__ void begin_frame() {}
    // once per pass, before the first load: this thread's stores of the previous pass have landed
    __device__ __forceinline__ void begin_pass() const { tmem_wait_st(); }
    template <int EOFF, int ROW, int D>
    __device__ __forceinline__ void load(RowMsg<D>& m, bool) const
    {
        static_assert(D <= 8, "a row is one 16-column TMEM access");
        uint32_t r[16];
        tmem_ld16(taddr0 + ROW * ROW_COLS, r);
        tmem_wait_ld16(r);
#pragma unroll
        for (int k = 0; k < D; ++k) m.v[k] = f2(__uint_as_float(r[2 * k]), __uint_as_float(r[2 * k + 1]));
    }
    template <int D> __device__ __forceinline__ void load_done(RowMsg<D>&) const {}
    template <int EOFF, int ROW, int D>
    __device__ __forceinline__ void store(const RowMsg<D>& m)
    {
        uint32_t r[16];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            r[2 * k] = k < D ? __float_as_uint(m.v[k].x) : 0u;
            r[2 * k + 1] = k < D ? __float_as_uint(m.v[k].y) : 0u;
        }
        tmem_st16(taddr0 + ROW * ROW_COLS, r);
    }
};

// What the front half of a row leaves for the back half (the variable->check messages themselves are
// parked in the row's message registers until the back half overwrites them).
template <int D>
struct RowFront2 {
    float2 x[D > 0 ? D : 1];        // 2^-|m| of every edge of the check, both frames
    uint32_t sgn_a, sgn_b;          // xor of the message bit patterns (product of the signs) + odd-degree fix-up
};

// Front half: rotated index, previous posterior (LDS.64), variable->check message (FADD2), x = 2^-|m|.
// m: in = the row's check->variable messages of the previous pass, out = its variable->check messages.
template <int Z, bool EARLY, class... S>
__device__ __forceinline__ void row_front2(Row<S...>, RowMsg<sizeof...(S)>& m, const float2* __restrict__ sm, const int r,
                                           const int po, const bool fix_odd, const bool act, bool& unsat_a, bool& unsat_b,
                                           RowFront2<sizeof...(S)>& f)
{
    constexpr int D = sizeof...(S);
    if constexpr (D == 0) return;
    static_assert(D == 0 || D >= 2, "the prefix/suffix scheme needs check degree >= 2");
    constexpr int CB[D > 0 ? D : 1] = {(S::colb * Z)...};
    constexpr int SH[D > 0 ? D : 1] = {S::shift...};
    f.sgn_a = f.sgn_b = (fix_odd && (D & 1)) ? 0x80000000u : 0u;
    bool par_a = (D & 1) != 0, par_b = (D & 1) != 0;      // parity of the estimates, spa_decoder.py:188-195
    // By kind of operation -- all loads, then all subtractions and sign products, then all exponentials -- rather than edge
    // by edge: same instructions, but ptxas schedules this order 0.5-1.5 % faster (16.29 against 16.37-16.54 ms per launch,
    // tools/r2_call54.sh; profiles/r2_tuning.md 5.10).
    float2 Lv[D > 0 ? D : 1];
#pragma unroll
    for (int c = 0; c < D; ++c) {
        int idx = r + SH[c];
        idx = (int)min((unsigned)idx, (unsigned)(idx - Z));      // (r + shift) mod z
#ifdef LDPC_EXP_NOSMEM      // timing experiment (wrong results): no shared-memory traffic in the check-node phase
        Lv[c] = f2((float)(idx + c) * 0.01f, (float)(idx - c) * 0.02f);
#else
        Lv[c] = act ? sh_ld2(sm, po + CB[c] + idx) : f2(1.f, 1.f);
#endif
    }
#pragma unroll
    for (int c = 0; c < D; ++c) {
        const float2 mu = f2sub(Lv[c], m.v[c]);                  // variable->check messages, :260-268
        if (EARLY) { par_a ^= (Lv[c].x < 0.f); par_b ^= (Lv[c].y < 0.f); }
        m.v[c] = mu;                                             // parked until row_back2 (signs)
        f.sgn_a ^= __float_as_uint(mu.x);
        f.sgn_b ^= __float_as_uint(mu.y);
    }
#pragma unroll
    for (int c = 0; c < D; ++c)
        f.x[c] = f2(ex2_approx(-fminf(fabsf(m.v[c].x), kClipBits)),      // x = 2^-|m|, both clips of :133-146,167
                    ex2_approx(-fminf(fabsf(m.v[c].y), kClipBits)));
    if (EARLY) { unsat_a |= par_a; unsat_b |= par_b; }
}

// Back half: leave-one-out products with packed FMAs, |E| = lg2(A / B) with one reciprocal per two edges, signs.
// (The expressions are those of row_back in qc_kernel.cuh, lane by lane.)
template <class... S>
__device__ __forceinline__ void row_back2(Row<S...>, RowMsg<sizeof...(S)>& m, const RowFront2<sizeof...(S)>& f)
{
    constexpr int D = sizeof...(S);
    if constexpr (D == 0) return;
    constexpr int DD = D > 0 ? D : 1;
    const float2 one = f2(1.f, 1.f);
    // prefix products (fa + fb e) over slots 0..k, e^2 = 1
    float2 fa[DD], fb[DD];
    fa[0] = one;
    fb[0] = f.x[0];
#pragma unroll
    for (int k = 1; k < D - 1; ++k) {
        if (k == 1) {
            fa[1] = f2fma(f.x[0], f.x[1], one);
            fb[1] = f2add(f.x[0], f.x[1]);
        } else {
            fa[k] = f2fma(fb[k - 1], f.x[k], fa[k - 1]);
            fb[k] = f2fma(fa[k - 1], f.x[k], fb[k - 1]);
        }
    }
    float2 sa = one, sb = f2(0.f, 0.f);      // suffix product over slots > k
    float2 heldA = one, heldB = one;         // first edge of a pair, waiting for its partner
    int heldk = 0;
#pragma unroll
    for (int k = D - 1; k >= 0; --k) {
        float2 A, B;
        if (k == D - 1) { A = fa[D - 2]; B = fb[D - 2]; }
        else if (k == 0) { A = sa; B = sb; }
        else if (k == D - 2) {               // suffix = (1 + x[D-1] e)
            A = f2fma(fb[k - 1], sb, fa[k - 1]);
            B = f2fma(fa[k - 1], sb, fb[k - 1]);
        } else if (k == 1) {                 // prefix = (1 + x[0] e)
            A = f2fma(fb[0], sb, sa);
            B = f2fma(fb[0], sa, sb);
        } else {
            A = f2fma(fa[k - 1], sa, f2mul(fb[k - 1], sb));
            B = f2fma(fa[k - 1], sb, f2mul(fb[k - 1], sa));
        }
        // |E| = lg2(A / B) (:151-168).  The quotients of TWO edges share one reciprocal, 1 / (B_h B): 2.5 instead of 3
        // MUFU per edge (B >= 6e-16 by the clip on |m|, so the product of two stays a normal number).  Edges are
        // paired in the order they are finished (k = D-1 with D-2, D-3 with D-4, ...); an odd row leaves k = 0 alone.
        {
            const bool second = ((D - 1 - k) & 1) != 0;
            if (!second && k > 0) { heldA = A; heldB = B; heldk = k; }
            else {
                float2 Rh = one, R;
                if (second) {
                    const float2 P = f2mul(heldB, B);
                    const float2 rp = f2(rcp_approx(P.x), rcp_approx(P.y));
                    Rh = f2mul(heldA, f2mul(rp, B));
                    R = f2mul(A, f2mul(rp, heldB));
                    const float2 mg = f2(lg2_approx(Rh.x), lg2_approx(Rh.y));
                    const uint32_t s1 = (f.sgn_a ^ __float_as_uint(m.v[heldk].x)) & 0x80000000u;
                    const uint32_t s2 = (f.sgn_b ^ __float_as_uint(m.v[heldk].y)) & 0x80000000u;
                    m.v[heldk] = f2(__uint_as_float(__float_as_uint(mg.x) | s1), __uint_as_float(__float_as_uint(mg.y) | s2));
                } else {
                    R = f2mul(A, f2(rcp_approx(B.x), rcp_approx(B.y)));
                }
                const float2 mag = f2(lg2_approx(R.x), lg2_approx(R.y));
                const uint32_t sa_bit = (f.sgn_a ^ __float_as_uint(m.v[k].x)) & 0x80000000u;
                const uint32_t sb_bit = (f.sgn_b ^ __float_as_uint(m.v[k].y)) & 0x80000000u;
                m.v[k] = f2(__uint_as_float(__float_as_uint(mag.x) | sa_bit), __uint_as_float(__float_as_uint(mag.y) | sb_bit));
            }
        }
        if (k == D - 1) { sb = f.x[k]; }     // sa stays 1
        else if (k == D - 2) {
            sa = f2fma(sb, f.x[k], one);
            sb = f2add(sb, f.x[k]);
        } else if (k > 0) {
            const float2 na = f2fma(sb, f.x[k], sa);
            sb = f2fma(sa, f.x[k], sb);
            sa = na;
        }
    }
}

// Scatter half: add the new messages into the posterior being built (:173-185); LDS.64 / FADD2 / STS.64.
template <int Z, class... S>
__device__ __forceinline__ void row_scatter2(Row<S...>, const RowMsg<sizeof...(S)>& m, float2* __restrict__ sm, const int r,
                                             const int no, const bool act)
{
    constexpr int D = sizeof...(S);
    if constexpr (D == 0) return;
    constexpr int DD = D > 0 ? D : 1;
    constexpr int CB[DD] = {(S::colb * Z)...};
    constexpr int SH[DD] = {S::shift...};
    constexpr bool FI[DD] = {S::first...};
    if (!act) return;
    int off[DD];
    float2 base[DD];
#pragma unroll
    for (int k = 0; k < D; ++k) {
        int idx = r + SH[k];
        idx = (int)min((unsigned)idx, (unsigned)(idx - Z));
        off[k] = CB[k] + idx;
        base[k] = FI[k] ? sm[off[k]] : sm[no + off[k]];
    }
#pragma unroll
    for (int k = 0; k < D; ++k) sm[no + off[k]] = f2add(base[k], m.v[k]);
}

// Software-pipelined sweep over the groups of a pass (see team_pipe in qc_kernel.cuh).  ROW counts the rows of
// this team (its TMEM row slot), EOFF the message registers before the row.
template <int Z, int TEAM, int EOFF, int ROW, bool EARLY, class MS, class GCUR>
__device__ __forceinline__ void team_pipe2(MS& ms, float2* __restrict__ sm, const int r, const int po,
                                           const int no, const bool fix_odd, const bool act, const bool first_pass,
                                           bool& ua, bool& ub, const uint32_t bar, uint32_t& phase, const bool first_group,
                                           RowMsg<TeamRow<TEAM, GCUR>::type::D>& mcur,
                                           const RowFront2<TeamRow<TEAM, GCUR>::type::D>& fcur, GCUR)
{
    using R = typename TeamRow<TEAM, GCUR>::type;
    row_back2(R(), mcur, fcur);
    if (!first_group) mbar_wait(bar, (phase - 1) & 1u);
    row_scatter2<Z>(R(), mcur, sm, r, no, act);
    if constexpr (R::D > 0) ms.template store<EOFF, ROW>(mcur);
    mbar_arrive(bar);
    ++phase;
}

template <int Z, int TEAM, int EOFF, int ROW, bool EARLY, class MS, class GCUR, class GNEXT, class... Rest>
__device__ __forceinline__ void team_pipe2(MS& ms, float2* __restrict__ sm, const int r, const int po,
                                           const int no, const bool fix_odd, const bool act, const bool first_pass,
                                           bool& ua, bool& ub, const uint32_t bar, uint32_t& phase, const bool first_group,
                                           RowMsg<TeamRow<TEAM, GCUR>::type::D>& mcur,
                                           const RowFront2<TeamRow<TEAM, GCUR>::type::D>& fcur, GCUR, GNEXT gn, Rest... rest)
{
    using R = typename TeamRow<TEAM, GCUR>::type;
    using RN = typename TeamRow<TEAM, GNEXT>::type;
    constexpr int NROW = ROW + (R::D > 0 ? 1 : 0);
    RowMsg<RN::D> mnext;
    if constexpr (RN::D > 0) ms.template load<EOFF + R::D, NROW>(mnext, first_pass);
    row_back2(R(), mcur, fcur);
    RowFront2<RN::D> fnext;
    row_front2<Z, EARLY>(RN(), mnext, sm, r, po, fix_odd, act, ua, ub, fnext);
    if (!first_group) mbar_wait(bar, (phase - 1) & 1u);
    row_scatter2<Z>(R(), mcur, sm, r, no, act);
    if constexpr (R::D > 0) ms.template store<EOFF, ROW>(mcur);
    mbar_arrive(bar);
    ++phase;
    team_pipe2<Z, TEAM, EOFF + R::D, NROW, EARLY>(ms, sm, r, po, no, fix_odd, act, first_pass, ua, ub, bar, phase, false,
                                                  mnext, fnext, gn, rest...);
}

template <int Z, int TEAM, bool EARLY, class MS, class G0, class... Rest>
__device__ __forceinline__ void team_pass2(MS& ms, float2* __restrict__ sm, const int r, const int po,
                                           const int no, const bool fix_odd, const bool act, const bool first_pass,
                                           bool& ua, bool& ub, const uint32_t bar, uint32_t& phase, G0 g0, Rest... rest)
{
    using R0 = typename TeamRow<TEAM, G0>::type;
    RowMsg<R0::D> m0;
    ms.begin_pass();
    if constexpr (R0::D > 0) ms.template load<0, 0>(m0, first_pass);
    if (phase) mbar_wait(bar, (phase - 1) & 1u);         // the previous pass' posterior is complete
    RowFront2<R0::D> f0;
    row_front2<Z, EARLY>(R0(), m0, sm, r, po, fix_odd, act, ua, ub, f0);
    team_pipe2<Z, TEAM, 0, 0, EARLY>(ms, sm, r, po, no, fix_odd, act, first_pass, ua, ub, bar, phase, true, m0, f0, g0, rest...);
}

// syndrome of one block row on the posterior at float2 offset po, both frames
template <int Z, class... S>
__device__ __forceinline__ void row_unsat2(Row<S...>, const float2* __restrict__ sm, const int r, const int po, bool& ua, bool& ub)
{
    constexpr int D = sizeof...(S);
    if constexpr (D == 0) return;
    constexpr int CB[D > 0 ? D : 1] = {(S::colb * Z)...};
    constexpr int SH[D > 0 ? D : 1] = {S::shift...};
    bool pa = (D & 1) != 0, pb = (D & 1) != 0;
#pragma unroll
    for (int c = 0; c < D; ++c) {
        int idx = r + SH[c];
        idx = (int)min((unsigned)idx, (unsigned)(idx - Z));
        const float2 L = sm[po + CB[c] + idx];
        pa ^= (L.x < 0.f);
        pb ^= (L.y < 0.f);
    }
    ua |= pa;
    ub |= pb;
}

template <int Z, int TEAM, class... G>
__device__ __forceinline__ void team_unsat2(const float2* __restrict__ sm, const int r, const int po, bool& ua, bool& ub, G...)
{
    (row_unsat2<Z>(typename TeamRow<TEAM, G>::type(), sm, r, po, ua, ub), ...);
}

// CTA-wide OR of a 2-bit value: shared-memory atomicOr into one of three rotating slots + one barrier.
// Slot (k+2)%3 is cleared after barrier k: its last readers passed barrier k-1 and arrived at barrier k,
// its next writers come after barrier k+1.
__device__ __forceinline__ unsigned block_or2(unsigned bits, unsigned* s_or, unsigned& use)
{
    const unsigned slot = use % 3u;
    if (bits) atomicOr(&s_or[slot], bits);
    __syncthreads();
    const unsigned v = s_or[slot];
    if (threadIdx.x == 0) s_or[(use + 2u) % 3u] = 0u;
    ++use;
    return v;
}

// Outputs of one frame of the pair (lane 0 = .x, lane 1 = .y) from the posterior at fin.
template <int THREADS, int N>
__device__ __forceinline__ void store_frame(const Outputs& out, const float2* __restrict__ fin, const int lane,
                                            const long long f, const int conv)
{
    const float* fl = reinterpret_cast<const float*>(fin) + lane;        // element j at fl[2 j]
    if (threadIdx.x == 0) {
        if (out.conv_it) out.conv_it[f] = conv;
        if (out.ok) out.ok[f] = conv >= 0 ? 1 : 0;
    }
    if (out.zbits) {
        constexpr int words = (N + 31) / 32;
        for (int base = 0; base < words * 32; base += THREADS) {
            const int j = base + threadIdx.x;
            const bool neg = (j < N) && (fl[2 * j] < 0.f);
            const unsigned m = __ballot_sync(0xffffffffu, neg);
            if ((threadIdx.x & 31) == 0 && j < words * 32) out.zbits[(size_t)f * words + (j >> 5)] = m;
        }
    }
    if (out.z) {
        uint8_t* dst = out.z + (size_t)f * N;
        for (int j = threadIdx.x; j < N; j += THREADS) dst[j] = (uint8_t)(fl[2 * j] < 0.f);   // :188
    }
    if (out.post) {
        float* dst = out.post + (size_t)f * N;
        for (int j = threadIdx.x; j < N; j += THREADS) dst[j] = fl[2 * j] * kLn2;
    }
}

template <bool TM, int NS> struct MsgStore { using type = RegMsgs<NS>; };
template <int NS> struct MsgStore<true, NS> { using type = TmemMsgs; };

// TMEM columns a CTA of the TMEM variant allocates: the warps that share a lane quarter (warp % 4) stack their rows.
template <class C>
struct TmemShape {
    static constexpr int TZ = (C::Z + 31) / 32 * 32;
    static constexpr int WARPS = TZ * C::TEAMS / 32;
    static constexpr int ROWS = C::GROUPS;                       // rows per thread (one per group)
    static constexpr int NEED = (WARPS + 3) / 4 * ROWS * TmemMsgs::ROW_COLS;
    static constexpr int COLS = NEED <= 32 ? 32 : NEED <= 64 ? 64 : NEED <= 128 ? 128 : NEED <= 256 ? 256 : 512;
    static constexpr int MAX_CTAS = 512 / COLS;                  // per SM; the launcher caps the occupancy to this
    static constexpr bool FITS = NEED <= 512 && C::MAXDEG <= 8;
};

template <int THREADS, bool EARLY, bool TM, int Z, int N, class... G>
__device__ __forceinline__ void decode_pairs(Code<Z, N, G...>, const float* __restrict__ llr, const Outputs& out,
                                             long long frames, int max_iter, int fix_odd, const McParams& mc,
                                             unsigned long long* __restrict__ work_counter)
{
    using C = Code<Z, N, G...>;
    constexpr int NS = C::MAXSLOT;
    constexpr int TZ = (Z + 31) / 32 * 32;            // threads of one team
    static_assert(THREADS == TZ * C::TEAMS, "CTA size = teams x ceil32(z)");
    // float2 units: [0,N) channel | [N,2N) | [2N,3N) posteriors | [3N,4N) = 2N floats of TMA stage (two LLR rows)
    extern __shared__ __align__(16) float2 sm2[];
    __shared__ long long s_pair;
    __shared__ unsigned long long s_cnt[5];
    __shared__ int s_err[2];
    __shared__ unsigned s_or[3];
    __shared__ __align__(8) unsigned long long s_bar;
    __shared__ __align__(8) unsigned long long s_tma_bar;
    __shared__ uint32_t s_tmem;

    const int team = (C::TEAMS == 1) ? 0 : (int)(threadIdx.x / TZ);     // warp-uniform
    const int r = (C::TEAMS == 1) ? (int)threadIdx.x : (int)(threadIdx.x - team * TZ);
    const bool row_ok = (Z == TZ) ? true : (r < Z);
    if (threadIdx.x < 5) s_cnt[threadIdx.x] = 0;
    if (threadIdx.x < 3) s_or[threadIdx.x] = 0;
    if constexpr (TM) {
        // one warp allocates the CTA's TMEM columns (power of two) and gives the allocation permit back at once
        if (threadIdx.x < 32) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                         :: "r"((uint32_t)__cvta_generic_to_shared(&s_tmem)), "n"(TmemShape<C>::COLS) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&s_bar);
    const uint32_t tma_bar = (uint32_t)__cvta_generic_to_shared(&s_tma_bar);
    if (threadIdx.x == 0) {
        mbar_init(bar, THREADS);
        mbar_init(tma_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    uint32_t phase = 0;                      // barrier phases completed so far (uniform across the CTA)
    uint32_t tma_phase = 0;                  // LLR row pairs received so far
    unsigned or_use = 0;
    const long long pairs = (frames + 1) / 2;
    using MS = typename MsgStore<TM, NS>::type;
    MS ms;
    // host-fed LLR rows are prefetched by TMA one pair ahead (rows are 16-byte multiples and aligned)
    const bool use_tma = !mc.active && (N % 4 == 0) && ((reinterpret_cast<uintptr_t>(llr) & 15) == 0);
    float* stage_f = reinterpret_cast<float*>(sm2 + 3 * N);
    const uint32_t stage = (uint32_t)__cvta_generic_to_shared(stage_f);
    __syncthreads();
    if constexpr (TM) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t warp = threadIdx.x >> 5;
        // lane quarter of this warp in bits 31:16; the warps of one quarter stack their rows along the columns
        ms.taddr0 = s_tmem + (((warp & 3u) * 32u) << 16) + (warp >> 2) * (uint32_t)(TmemShape<C>::ROWS * TmemMsgs::ROW_COLS);
    }

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

        // ---- prologue: channel LLRs of both frames -> shared memory (float2, log2 units) ----
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
                    if (4 * q + i < N) sm2[4 * q + i] = f2(va[i] * kLog2e, vb[i] * kLog2e);
            }
        } else if (use_tma) {
            mbar_wait(tma_bar, tma_phase & 1u);          // this pair's rows have landed in the stage buffer
            ++tma_phase;
            const float4* ra = reinterpret_cast<const float4*>(stage_f);
            const float4* rb = reinterpret_cast<const float4*>(stage_f + (has_b ? N : 0));
            for (int q = threadIdx.x; q < N / 4; q += THREADS) {
                const float4 a = ra[q], b = rb[q];
                float4* dst = reinterpret_cast<float4*>(sm2 + 4 * q);
                dst[0] = make_float4(a.x * kLog2e, b.x * kLog2e, a.y * kLog2e, b.y * kLog2e);
                dst[1] = make_float4(a.z * kLog2e, b.z * kLog2e, a.w * kLog2e, b.w * kLog2e);
            }
            __syncthreads();                             // stage fully consumed -> prefetch the next pair
            if (threadIdx.x == 0 && p_next < pairs)
                tma_load_row(stage, llr + (size_t)(2 * p_next) * N, rows_of(p_next) * N * 4, tma_bar);
        } else {
            const float* sa_ = llr + (size_t)fa * N;
            const float* sb_ = llr + (size_t)(has_b ? fb : fa) * N;
            for (int j = threadIdx.x; j < N; j += THREADS) sm2[j] = f2(__ldg(sa_ + j) * kLog2e, __ldg(sb_ + j) * kLog2e);
        }
        __syncthreads();

        ms.begin_frame();
        ms.template begin_frame_rows<C::GROUPS>();

        int po = 0, no = N;          // pass 0 reads the channel values as the "previous posterior"
        int conv_a = -1, conv_b = -1;
        bool done_a = false, done_b = !has_b;
        for (int it = 0; it < max_iter; ++it) {
            bool ua = false, ub = false;
            if (team == 0) team_pass2<Z, 0, EARLY>(ms, sm2, r, po, no, fix_odd != 0, row_ok, it == 0, ua, ub, bar, phase, G()...);
            if constexpr (C::TEAMS > 1) { if (team == 1) team_pass2<Z, 1, EARLY>(ms, sm2, r, po, no, fix_odd != 0, row_ok, it == 0, ua, ub, bar, phase, G()...); }
            if constexpr (C::TEAMS > 2) { if (team == 2) team_pass2<Z, 2, EARLY>(ms, sm2, r, po, no, fix_odd != 0, row_ok, it == 0, ua, ub, bar, phase, G()...); }
            if constexpr (C::TEAMS > 3) { if (team == 3) team_pass2<Z, 3, EARLY>(ms, sm2, r, po, no, fix_odd != 0, row_ok, it == 0, ua, ub, bar, phase, G()...); }
            if (!row_ok) ua = ub = false;
            if (EARLY && it > 0) {
                // a frame whose posterior of pass it-1 (at po) satisfied every check exits with that posterior; its
                // outputs are written now, because the pair keeps iterating until both frames are done
                const unsigned u = block_or2((ua ? 1u : 0u) | (ub ? 2u : 0u), s_or, or_use);
                if (!done_a && !(u & 1u)) { conv_a = it - 1; done_a = true; store_frame<THREADS, N>(out, sm2 + po, 0, fa, conv_a); }
                if (!done_b && !(u & 2u)) { conv_b = it - 1; done_b = true; store_frame<THREADS, N>(out, sm2 + po, 1, fb, conv_b); }
                if (done_a && done_b) break;
            }
            po = no;
            no = (no == N) ? 2 * N : N;
        }

        __syncthreads();     // every scatter of the last executed pass has landed
        // ---- exit: syndrome of the last posterior for the frames that have not converged yet (po = exit posterior) ----
        if (!(done_a && done_b)) {
            bool ua = false, ub = false;
            if (row_ok) {
                if (team == 0) team_unsat2<Z, 0>(sm2, r, po, ua, ub, G()...);
                if constexpr (C::TEAMS > 1) { if (team == 1) team_unsat2<Z, 1>(sm2, r, po, ua, ub, G()...); }
                if constexpr (C::TEAMS > 2) { if (team == 2) team_unsat2<Z, 2>(sm2, r, po, ua, ub, G()...); }
                if constexpr (C::TEAMS > 3) { if (team == 3) team_unsat2<Z, 3>(sm2, r, po, ua, ub, G()...); }
            }
            const unsigned u = block_or2((ua ? 1u : 0u) | (ub ? 2u : 0u), s_or, or_use);
            if (!done_a) { if (!(u & 1u)) conv_a = max_iter - 1; store_frame<THREADS, N>(out, sm2 + po, 0, fa, conv_a); }
            if (!done_b) { if (!(u & 2u)) conv_b = max_iter - 1; store_frame<THREADS, N>(out, sm2 + po, 1, fb, conv_b); }
        }

        if (mc.active) {
            // main.py:314-339: bit errors only in failed frames, on the un-complemented output.  A failed frame ran
            // every pass, so its exit posterior is the one at po.
            if (threadIdx.x < 2) s_err[threadIdx.x] = 0;
            __syncthreads();
            const float2* fin = sm2 + po;
            const bool bad_a = conv_a < 0, bad_b = has_b && conv_b < 0;
            if (bad_a || bad_b) {
                int ea = 0, eb = 0;
                const int span = mc.info_mask ? N : mc.k_info;
                for (int j = threadIdx.x; j < span; j += THREADS) {
                    if (mc.info_mask && !mc.info_mask[j]) continue;
                    const float2 L = fin[j];
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
    if constexpr (TM) {
        tmem_wait_st();
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (threadIdx.x < 32)
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(s_tmem), "n"(TmemShape<C>::COLS) : "memory");
    }
}

// CTA shape of the pair kernel: same threads as the one-frame kernel, twice the message registers, so
// one CTA per SM unless the code is small.
template <class C>
struct PairShape {
    static constexpr int TZ = (C::Z + 31) / 32 * 32;
    static constexpr int THREADS = TZ * C::TEAMS;
    static constexpr bool ENABLED = 2 * C::MAXSLOT <= 96;      // the messages of both frames must fit the register file without spills
    static constexpr int NEED = 4 * C::MAXSLOT + 100;          // registers a thread wants
    static constexpr int REGS = NEED > 255 ? 255 : (NEED < 128 ? 128 : NEED);
    static constexpr int B0 = 65536 / (THREADS * REGS);
    static constexpr int MINB = B0 < 1 ? 1 : (B0 > 32 ? 32 : B0);
    // TMEM variant: ~168 registers per thread (two rows of messages + the working set of a check row), at most
    // TmemShape::MAX_CTAS CTAs per SM (TMEM columns)
    static constexpr bool TM_ENABLED = TmemShape<C>::FITS;
    static constexpr int TB0 = 65536 / (THREADS * 168);
    static constexpr int TB1 = TB0 < 1 ? 1 : TB0;
    static constexpr int TM_MINB = TB1 > TmemShape<C>::MAX_CTAS ? TmemShape<C>::MAX_CTAS : TB1;
};

template <int THREADS, int MINB, bool EARLY, class C>
__global__ void __launch_bounds__(THREADS, MINB)
k_qc_pair(const float* __restrict__ llr, Outputs out, long long frames, int max_iter, int fix_odd, McParams mc,
          unsigned long long* __restrict__ work_counter)
{
    decode_pairs<THREADS, EARLY, false>(C(), llr, out, frames, max_iter, fix_odd, mc, work_counter);
}

// Messages in tensor memory: the registers hold two rows at a time, so MINB is twice that of k_qc_pair.
template <int THREADS, int MINB, bool EARLY, class C>
__global__ void __launch_bounds__(THREADS, MINB)
k_qc_pair_tmem(const float* __restrict__ llr, Outputs out, long long frames, int max_iter, int fix_odd, McParams mc,
               unsigned long long* __restrict__ work_counter)
{
    decode_pairs<THREADS, EARLY, true>(C(), llr, out, frames, max_iter, fix_odd, mc, work_counter);
}

}  // namespace qc
}  // namespace ldpc
