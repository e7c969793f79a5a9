// f64_math.cuh -- the three fp64 functions of the reference's check node (spa_decoder.py:133-168), hand-rolled.
//
// The parity-grade (fp64) generic kernels spent 345 instructions per edge, almost all of them inside CUDA's libm
// tanh / atanh and the IEEE division (profiles/r1_ncu_generic_kernels.txt).  The check node only ever evaluates
//     t = tanh(m / 2)       for |m| <= 35          (the reference clips beyond, :140-146)
//     r = P / t             a plain quotient       (:157)
//     E = 2 atanh(r)        for |r| <= 1 - 1.22e-15 (:167-168)
// so the general-purpose special-case handling (NaN / Inf / denormal arguments, huge arguments) is dead weight, and the
// exponential inside tanh and the logarithm inside atanh can share their reciprocals with the quotients around them:
//     tanh(m/2)   one exp core (2^k * (1 + p(r)), Taylor degree 13 on |r| <= ln2/2), ONE reciprocal y = 1/(1 + w):
//                 t = 1 - 2 w y for |m| > 2.5 (correctly rounded near saturation, where the reference's results hinge on the
//                 last bit of t) and t = -expm1(-|m|) y for small |m| (relative accuracy down to m -> 0);
//     P / t       reciprocal by Newton's iteration from rcp.approx.f64 + one residual correction (the IEEE result
//                 except in rare double-rounding cases);
//     2 atanh(r)  |r| <= 0.17: 2 r P(r^2) directly; otherwise ln((1+r)/(1-r)) with BOTH range reductions folded into one
//                 quotient s = (mN - mD) / (mN + mD), |s| <= 0.1716, and the same polynomial P.
// Accuracy (tools/f64_math_check.cpp against binary128, 4e6 points each): tanh <= 2.6 ulp, 0.50 ulp (correctly rounded)
// where |m| > 20; 2 atanh <= 3.6 ulp; the quotient equals the IEEE one in all 4e6 cases.  The same source is compiled by
// g++ for that check.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

// Under nvcc the functions are device code and their coefficients sit in constant memory: an fp64 instruction takes a
// constant-bank operand directly, whereas a literal costs two UMOV per use (52 of the first version's 245 instructions
// per edge were such moves).  g++ (tools/f64_math_check.cpp) sees plain inline functions and a static table.
#ifdef __CUDACC__
#define LDPC_HD __device__ __forceinline__
#define LDPC_TABLE static __constant__
#else
#define LDPC_HD inline
#define LDPC_TABLE static const
#endif

namespace ldpc {
namespace f64 {

// The same coefficients as literals (LIT functions below): inside a rolled loop over a long check row the compiler
// re-loads table entries on every trip (LDC, a variable-latency load), whereas literals are uniform-register moves.
struct Lit {
    static LDPC_HD double expm1(int i)
    {
        const double c[12] = {1.6059043836821613e-10, 2.08767569878681e-09, 2.505210838544172e-08, 2.755731922398589e-07,
                              2.7557319223985893e-06, 2.48015873015873e-05, 1.984126984126984e-04, 1.388888888888889e-03,
                              8.333333333333333e-03, 4.1666666666666664e-02, 1.6666666666666666e-01, 0.5};
        return c[i];
    }
    static LDPC_HD double atanh(int i)
    {
        const double c[10] = {1.0 / 21.0, 1.0 / 19.0, 1.0 / 17.0, 1.0 / 15.0, 1.0 / 13.0, 1.0 / 11.0, 1.0 / 9.0, 1.0 / 7.0, 1.0 / 5.0, 1.0 / 3.0};
        return c[i];
    }
    static LDPC_HD double misc(int i)
    {
        const double c[6] = {6755399441055744.0, 1.4426950408889634, -6.93147180369123816490e-01, -1.90821492927058770002e-10,
                             1.4142135623730951, 0.17};
        return c[i];
    }
};

LDPC_TABLE double kExpm1[12] = {                                 // 1/13! ... 1/2!
    1.6059043836821613e-10, 2.08767569878681e-09, 2.505210838544172e-08, 2.755731922398589e-07,
    2.7557319223985893e-06, 2.48015873015873e-05, 1.984126984126984e-04, 1.388888888888889e-03,
    8.333333333333333e-03, 4.1666666666666664e-02, 1.6666666666666666e-01, 0.5};
LDPC_TABLE double kAtanh[10] = {                                 // 1/21 ... 1/3
    1.0 / 21.0, 1.0 / 19.0, 1.0 / 17.0, 1.0 / 15.0, 1.0 / 13.0, 1.0 / 11.0, 1.0 / 9.0, 1.0 / 7.0, 1.0 / 5.0, 1.0 / 3.0};
LDPC_TABLE double kMisc[6] = {
    6755399441055744.0,                                          // 1.5 * 2^52: the low word of (v + magic) is rint(v)
    1.4426950408889634,                                          // log2(e)
    -6.93147180369123816490e-01, -1.90821492927058770002e-10,    // -ln2, split so that k * hi is exact
    1.4142135623730951, 0.17};

LDPC_HD double from_bits(uint64_t u)
{
#ifdef __CUDA_ARCH__
    return __longlong_as_double((long long)u);
#else
    double d; std::memcpy(&d, &u, 8); return d;
#endif
}
LDPC_HD uint64_t to_bits(double d)
{
#ifdef __CUDA_ARCH__
    return (uint64_t)__double_as_longlong(d);
#else
    uint64_t u; std::memcpy(&u, &d, 8); return u;
#endif
}

// 1 / d for a normal d, to within an ulp: seed (2^-23) + two Newton steps.
LDPC_HD double recip(double d)
{
    double y;
#ifdef __CUDA_ARCH__
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
#else
    y = (double)(1.0f / (float)d);
#endif
    double e = fma(-d, y, 1.0);
    y = fma(y, e, y);
    e = fma(-d, y, 1.0);
    y = fma(y, e, y);
    return y;
}

// a / b with the residual correction of the classic FMA division: the correctly rounded quotient unless a * y falls
// within a rounding error of a tie.
LDPC_HD double divide(double a, double b)
{
    const double y = recip(b);
    const double q = a * y;
    const double rem = fma(-b, q, a);
    return fma(rem, y, q);
}

// tanh(m / 2) with the reference's clip: beyond |m / 2| = 17.5 it returns +-0.99999999999999878 (:140-146), which IS the
// correctly rounded tanh(17.5) -- clamping the argument gives the same values without a branch.
template <bool TAB = true>
LDPC_HD double tanh_half(double m)
{
#define LDPC_C(tab, lit, i) (TAB ? tab[i] : Lit::lit(i))
    const double x = fmax(-fabs(m), -35.0);                      // e^x = e^(-2 |m/2|)
    // x = k ln2 + r, |r| <= ln2 / 2
    const double magic = LDPC_C(kMisc, misc, 0);
    const double kt = fma(x, LDPC_C(kMisc, misc, 1), magic);
    const double kf = kt - magic;
    const int k = (int)(uint32_t)to_bits(kt);
    double r = fma(kf, LDPC_C(kMisc, misc, 2), x);
    r = fma(kf, LDPC_C(kMisc, misc, 3), r);
    // p = e^r - 1 = r + r^2 (1/2! + r/3! + ... + r^11/13!)
    double q = LDPC_C(kExpm1, expm1, 0);
#pragma unroll
    for (int i = 1; i < 12; ++i) q = fma(q, r, LDPC_C(kExpm1, expm1, i));
    const double p = fma(r * r, q, r);
    const double s = from_bits((uint64_t)(uint32_t)(k + 1023) << 52);          // 2^k, -51 <= k <= 0
    const double w = fma(s, p, s);                               // e^x
    const double em = fma(s, p, s - 1.0);                        // e^x - 1
    const double y = recip(1.0 + w);
    const double big = fma(-2.0 * w, y, 1.0);                    // 1 - 2 w / (1 + w): one rounding near saturation
    const double d1 = 1.0 + w;
    const double sq = -em * y;                                   // relative accuracy as m -> 0 ...
    const double small = fma(fma(-d1, sq, -em), y, sq);          // ... with the residual correction of a division
    return copysign(x < -2.5 ? big : small, m);
}

// 2 atanh(clip(r)) = ln((1 + r) / (1 - r)), |r| clipped to 1 - 1.22e-15 as the reference does (:167).
template <bool TAB = true>
LDPC_HD double two_atanh(double r)
{
    const double a = fmin(fabs(r), 0.99999999999999878);
    double s = a, shift_hi = 0.0, shift_lo = 0.0;
    if (a > LDPC_C(kMisc, misc, 5)) {
        double n = 1.0 + a;                                      // in (1.17, 2)
        const double d = 1.0 - a;                                // exact for a >= 1/2; >= 1.2e-15
        const uint64_t db = to_bits(d);
        int e = (int)(db >> 52) - 1023;                          // d = 2^e md, md in [1, 2)
        double md = from_bits((db & 0x000fffffffffffffull) | 0x3ff0000000000000ull);
        if (n > LDPC_C(kMisc, misc, 4) * md) { md += md; e -= 1; }             // keep n / md inside [1/sqrt2, sqrt2]
        else if (n * LDPC_C(kMisc, misc, 4) < md) { md *= 0.5; e += 1; }
        s = divide(n - md, n + md);                              // (n - md) is exact (Sterbenz)
        const double ef = (double)e;
        shift_hi = ef * LDPC_C(kMisc, misc, 2);                  // exact: 32 trailing zero bits in ln2_hi
        shift_lo = ef * LDPC_C(kMisc, misc, 3);
    }
    // 2 atanh(s) = 2 s (1 + z/3 + z^2/5 + ... + z^10/21), z = s^2 <= 0.0295
    const double z = s * s;
    double q = LDPC_C(kAtanh, atanh, 0);
#pragma unroll
    for (int i = 1; i < 10; ++i) q = fma(q, z, LDPC_C(kAtanh, atanh, i));
    const double s2 = s + s;
    const double tail = fma(s2 * z, q, shift_lo);                // 2 s (P - 1) + low part of the exponent term
    return copysign(shift_hi + (s2 + tail), r);
}

#undef LDPC_C

}  // namespace f64
}  // namespace ldpc
