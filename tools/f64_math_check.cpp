// tools/f64_math_check.cpp -- accuracy of csrc/f64_math.cuh against binary128 (libquadmath), on the CPU.
//   g++ -O2 -ffp-contract=off -std=c++17 -o /tmp/f64_math_check tools/f64_math_check.cpp -lquadmath && /tmp/f64_math_check
#include <quadmath.h>
#include <cstdio>
#include <cstdlib>
#include <random>
#include "../ldpc-simulator_b200/csrc/f64_math.cuh"

static double ulp_of(double x) { x = fabs(x); return nextafter(x, INFINITY) - x; }

int main(int argc, char** argv)
{
    std::mt19937_64 rng(12345);
    std::uniform_real_distribution<double> U(0.0, 1.0);
    const long N = argc > 1 ? atol(argv[1]) : 4000000;
    double worst_t = 0, worst_t_sat = 0, worst_a = 0, worst_d = 0, arg_t = 0, arg_a = 0;
    long differ_libm_t = 0, differ_libm_a = 0, differ_div = 0;
    for (long i = 0; i < N; ++i) {
        // tanh(m/2): |m| log-uniform in [1e-13, 35], one third of the points uniform in [20, 35] (saturation)
        double m = (i % 3 == 0) ? 20.0 + 15.0 * U(rng) : exp(log(1e-13) + U(rng) * (log(35.0) - log(1e-13)));
        if (i & 1) m = -m;
        const double t = ldpc::f64::tanh_half(m);
        const __float128 tq = tanhq((__float128)m / 2);
        const double err = (double)fabsq(((__float128)t - tq)) / ulp_of((double)tq);
        if (fabs(m) > 20) { if (err > worst_t_sat) worst_t_sat = err; }
        else if (err > worst_t) { worst_t = err; arg_t = m; }
        if (t != tanh(m / 2)) ++differ_libm_t;
        // 2 atanh(r): r = 1 - 2^-u (saturation side) or log-uniform small
        double r = (i % 2 == 0) ? 1.0 - exp2(-1.0 - 48.7 * U(rng)) : exp(log(1e-13) + U(rng) * (log(0.999) - log(1e-13)));
        if (r > 0.99999999999999878) r = 0.99999999999999878;
        if (i & 2) r = -r;
        const double a = ldpc::f64::two_atanh(r);
        const __float128 aq = 2 * atanhq((__float128)r);
        const double erra = (double)fabsq(((__float128)a - aq)) / ulp_of((double)aq);
        if (erra > worst_a) { worst_a = erra; arg_a = r; }
        if (a != 2.0 * atanh(r)) ++differ_libm_a;
        // quotient
        const double num = (U(rng) - 0.5) * 2, den = t == 0 ? 1.0 : t;
        const double qd = ldpc::f64::divide(num, den);
        if (qd != num / den) ++differ_div;
        const double errd = fabs(qd - num / den) / ulp_of(num / den);
        if (errd > worst_d) worst_d = errd;
    }
    printf("tanh_half : max error %.3f ulp for |m| <= 20 (at m = %.17g), %.3f ulp for |m| > 20; differs from libm tanh in %.4f %% of the points\n",
           worst_t, arg_t, worst_t_sat, 100.0 * differ_libm_t / N);
    printf("two_atanh : max error %.3f ulp (at r = %.17g); differs from 2*libm atanh in %.4f %% of the points\n", worst_a, arg_a,
           100.0 * differ_libm_a / N);
    printf("divide    : differs from the IEEE quotient in %.5f %% of the points (max %.2f ulp)\n", 100.0 * differ_div / N, worst_d);
    // the reference's own constants
    printf("MAXULP tanh %.4f tanh_sat %.4f atanh %.4f div_diff %ld\n", worst_t, worst_t_sat, worst_a, differ_div);
    printf("tanh_half(35) = %.17g (clip constant 0.99999999999999878), two_atanh(clip) = %.17g (libm %.17g)\n", ldpc::f64::tanh_half(35.0),
           ldpc::f64::two_atanh(0.99999999999999878), 2 * atanh(0.99999999999999878));
    return 0;
}
