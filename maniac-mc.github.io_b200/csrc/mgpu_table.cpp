// mgpu_table.cpp -- piecewise-polynomial table of g(s) = erfc(alpha sqrt(s)) / sqrt(s).
//
// The reference evaluates q_i q_j erfc(alpha r) / r with the libm erfc for every pair
// (src/pairwise_energy_utils.f90:175).  On the GPU that one call is ~60 FP64 instructions;
// here g is fitted per interval of s = r^2 (intervals = top MGPU_TAB_K mantissa bits within
// each binary octave) by a truncated Chebyshev series of degree 6 computed in long double
// from 28 Chebyshev nodes, and converted to a monomial in u = s - centre.
// The fit error is checked in tests/test_host_logic.py (no GPU needed) through
// mgpu_coulomb_table_check.
#include <cmath>
#include <cstring>
#include "mgpu_table.h"

namespace {
inline long double g_exact(long double alpha, long double s)
{
    const long double r = sqrtl(s);
    return erfcl(alpha * r) / r;
}
}

void mgpu_build_coulomb_table(double alpha, double r_lo, double r_hi, int *emin_out, int *noct_out, std::vector<double> &tab)
{
    const int K = MGPU_TAB_K, D = 6, NC = D + 1, NI = 1 << K;
    int emin = (int)std::floor(std::log2(r_lo * r_lo));
    int emax = (int)std::floor(std::log2(r_hi * r_hi)) + 1;          // exclusive
    if (emax - emin > MGPU_TAB_MAXOCT) emax = emin + MGPU_TAB_MAXOCT;
    if (emax <= emin) emax = emin + 1;
    const int noct = emax - emin;
    tab.assign((size_t)noct * NI * MGPU_TAB_ROW, 0.0);
    const int NN = 4 * NC;                                            // Chebyshev nodes per interval
    const long double PI_L = 3.14159265358979323846264338327950288L;
    std::vector<long double> f(NN), a(NC), T0(NC), T1(NC), T2(NC), b(NC);
    for (int e = 0; e < noct; ++e)
        for (int j = 0; j < NI; ++j) {
            const long double base = ldexpl(1.0L, emin + e);
            const long double c = base * (1.0L + ((long double)j + 0.5L) / NI), h = base * 0.5L / NI;
            for (int i = 0; i < NN; ++i) f[i] = g_exact(alpha, c + h * cosl(PI_L * (i + 0.5L) / NN));
            for (int n = 0; n < NC; ++n) {
                long double s = 0.0L;
                for (int i = 0; i < NN; ++i) s += f[i] * cosl(PI_L * n * (i + 0.5L) / NN);
                a[n] = 2.0L * s / NN;
            }
            a[0] *= 0.5L;
            // sum_n a_n T_n(t) -> monomials b_m t^m
            for (int m = 0; m < NC; ++m) { b[m] = 0.0L; T0[m] = 0.0L; T1[m] = 0.0L; }
            T0[0] = 1.0L; T1[1] = 1.0L;
            for (int m = 0; m < NC; ++m) b[m] += a[0] * T0[m] + a[1] * T1[m];
            for (int n = 2; n < NC; ++n) {
                for (int m = 0; m < NC; ++m) T2[m] = (m > 0 ? 2.0L * T1[m - 1] : 0.0L) - T0[m];
                for (int m = 0; m < NC; ++m) b[m] += a[n] * T2[m];
                T0 = T1; T1 = T2;
            }
            // t = u / h; c5, c6 stored in single precision, packed into the sixth double
            double *row = tab.data() + (size_t)(e * NI + j) * MGPU_TAB_ROW;
            long double hp = 1.0L;
            float cf[2] = { 0.f, 0.f };
            for (int m = 0; m < NC; ++m) {
                if (m < 5) row[m] = (double)(b[m] / hp); else cf[m - 5] = (float)(b[m] / hp);
                hp *= h;
            }
            std::memcpy(&row[5], cf, 8);
        }
    *emin_out = emin; *noct_out = noct;
}

bool mgpu_eval_coulomb_table(const std::vector<double> &tab, int emin, int noct, double s, double *g)
{
    const int K = MGPU_TAB_K;
    uint64_t bits; std::memcpy(&bits, &s, 8);
    const int32_t hi = (int32_t)(bits >> 32);
    const int idx = (hi >> (20 - K)) - ((1023 + emin) << K);
    if (idx < 0 || idx >= (noct << K)) return false;
    const int32_t chi = (hi & ~((1 << (20 - K)) - 1)) | (1 << (19 - K));
    const uint64_t cb = (uint64_t)(uint32_t)chi << 32;
    double c; std::memcpy(&c, &cb, 8);
    const double u = s - c;
    const double *row = tab.data() + (size_t)idx * MGPU_TAB_ROW;
    float cf[2]; std::memcpy(cf, &row[5], 8);
    const float uf = (float)u;
    double p = (double)std::fmaf(cf[1], uf, cf[0]);
    for (int k = 4; k >= 0; --k) p = std::fma(p, u, row[k]);
    *g = p;
    return true;
}

extern "C" int mgpu_coulomb_table_check(double alpha, double r_lo, double r_hi, int32_t n, double *max_rel_err, double *max_abs_err_times_r)
{
    std::vector<double> tab; int emin = 0, noct = 0;
    mgpu_build_coulomb_table(alpha, r_lo, r_hi, &emin, &noct, tab);
    double er = 0.0, ea = 0.0;
    int covered = 0;
    for (int i = 0; i < n; ++i) {
        const double r = r_lo * std::exp(std::log(r_hi / r_lo) * (i + 0.5) / n);
        const double s = r * r;
        double g;
        if (!mgpu_eval_coulomb_table(tab, emin, noct, s, &g)) continue;
        ++covered;
        const long double ref = g_exact(alpha, s);
        const double d = (double)fabsl((long double)g - ref);
        if (ref > 1e-290L) er = std::fmax(er, d / (double)ref);
        ea = std::fmax(ea, d * r);                                    // g <= 1/r: error relative to the pair's Coulomb scale
    }
    if (max_rel_err) *max_rel_err = er;
    if (max_abs_err_times_r) *max_abs_err_times_r = ea;
    return covered;
}
