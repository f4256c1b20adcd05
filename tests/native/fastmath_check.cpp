// Host-side accuracy check of csrc/fastmath.cuh (the header builds for the host with the same sequence of
// IEEE operations as on the device).  Prints "name max_ulp_error samples" lines; tests/test_fastmath.py
// asserts the bounds.  Reference values: glibc long double (64-bit mantissa).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include "fastmath.cuh"
using namespace b200i::fm;

static double ulp_err(double got, long double want)
{
    if (want == 0.0L) return got == 0.0 ? 0.0 : 1e30;
    int e;
    frexpl(want, &e);                         // want = f * 2^e, f in [0.5,1)
    const long double ulp = ldexpl(1.0L, e - 53);
    return (double)(fabsl((long double)got - want) / ulp);
}

int main(int argc, char **argv)
{
    const long n = argc > 1 ? atol(argv[1]) : 2000000;
    std::mt19937_64 g(12345);
    std::uniform_real_distribution<double> U(0.0, 1.0);
    double e_div = 0, e_rcp = 0, e_log = 0, e_logsmall = 0, e_exp = 0, e_expwide = 0, e_cbrt = 0, e_d15 = 0;
    long bad_d15 = 0, bad_div = 0;
    const double K = 14137.166941154068;
    LogTabEntry tab[LOGT_SIZE];
    for (int j = 0; j < LOGT_SIZE; ++j) tab[j] = log_table_entry(j);
    const LogTabK ltk = log_tab_consts();
    const double logK = std::log(K);
    double e_logtab_abs = 0, e_logtab_ulp = 0;
    for (long i = 0; i < n; ++i) {
        const double a = std::exp((U(g) - 0.5) * 60.0), b = std::exp((U(g) - 0.5) * 60.0);
        const double q = div_fast(a, b);
        e_div = std::fmax(e_div, ulp_err(q, (long double)a / b));
        bad_div += (q != a / b);
        e_rcp = std::fmax(e_rcp, ulp_err(rcp_fast(b), 1.0L / b));
        // tumour volumes: 1e-12 .. 1150.35 against K
        const double V = std::exp(std::log(1e-12) + U(g) * (std::log(1150.3465) - std::log(1e-12)));
        e_log = std::fmax(e_log, ulp_err(log_ratio(K, V), logl((long double)K / V)));
        {   // table-driven log(K / den) over the projected volumes' range (den = V + 1e-7, V up to 1e4)
            const double den = std::exp(std::log(1e-7) + U(g) * (std::log(1e4) - std::log(1e-7)));
            const long double want = logl((long double)K / den);
            const double got = base_minus_log_tab(ltk, tab, logK, den);
            e_logtab_abs = std::fmax(e_logtab_abs, (double)fabsl((long double)got - want));
            if (fabsl(want) >= 1.0L) e_logtab_ulp = std::fmax(e_logtab_ulp, ulp_err(got, want));
        }
        // generic ratios with |log| >= 1
        if (std::fabs(std::log(a / b)) >= 1.0)
            e_logsmall = std::fmax(e_logsmall, ulp_err(log_ratio(a, b), logl((long double)a / b)));
        const double x = (U(g) - 0.5) * 16.0;
        e_exp = std::fmax(e_exp, ulp_err(exp_fast(x), expl((long double)x)));
        const double xw = (U(g) - 0.5) * 1400.0;
        e_expwide = std::fmax(e_expwide, ulp_err(exp_fast(xw), expl((long double)xw)));
        const double c = std::exp(std::log(1e-14) + U(g) * (std::log(1e3) - std::log(1e-14)));
        e_cbrt = std::fmax(e_cbrt, ulp_err(cbrt_fast(c), cbrtl((long double)c)));
        const int nn = 1 + (int)(U(g) * 15.0) % 15;
        const double s = U(g) * 200.0;
        const double d = div_small(s, (double)nn, kInvN[nn]);
        bad_d15 += (d != s / (double)nn);
        e_d15 = std::fmax(e_d15, ulp_err(d, (long double)s / nn));
    }
    printf("div_fast %.4f %ld\n", e_div, n);
    printf("div_fast_not_correctly_rounded %ld %ld\n", bad_div, n);
    printf("rcp_fast %.4f %ld\n", e_rcp, n);
    printf("log_ratio_path %.4f %ld\n", e_log, n);
    printf("log_ratio_generic %.4f %ld\n", e_logsmall, n);
    printf("log_tab_ulp %.4f %ld\n", e_logtab_ulp, n);
    printf("log_tab_abs %.4g %ld\n", e_logtab_abs, n);
    printf("exp_fast %.4f %ld\n", e_exp, n);
    printf("exp_fast_wide %.4f %ld\n", e_expwide, n);
    printf("cbrt_fast %.4f %ld\n", e_cbrt, n);
    printf("div_small %.4f %ld\n", e_d15, n);
    printf("div_small_not_correctly_rounded %ld %ld\n", bad_d15, n);
    return 0;
}
