/* CPU check of the division sequence of meshlesshydro_b200/csrc/mlh_internal.cuh (mlh_div_known): a/b from the correctly
 * rounded reciprocal y = RN(1/b) by a multiplication and two FMA corrections must equal the IEEE quotient bit for bit.
 * Returns the number of mismatches over `trials` pseudo-random quotients per divisor. */
#include <math.h>
#include <stdint.h>

static double div_known(double a, double b, double y) {
    double q = a * y;
    q = fma(fma(-b, q, a), y, q);
    return fma(fma(-b, q, a), y, q);
}

static uint64_t s_state;
static uint64_t rnd(void) { s_state ^= s_state << 13; s_state ^= s_state >> 7; s_state ^= s_state << 17; return s_state; }

long div_known_mismatches(const double *divisors, int ndiv, long trials, uint64_t seed) {
    long bad = 0;
    s_state = seed ? seed : 88172645463325252ull;
    for (int ib = 0; ib < ndiv; ++ib) {
        const double b = divisors[ib], y = 1.0 / b;
        for (long k = 0; k < trials; ++k) {
            double a;
            if (k & 1) a = (double)(rnd() >> 11) / 9007199254740992.0 * 4.0 * b;                              /* r in [0, 2h) */
            else a = ldexp(1.0 + (double)(rnd() >> 12) / 4503599627370496.0, (int)(rnd() % 40) - 30) * b;     /* quotient in [2^-30, 2^10) */
            if (div_known(a, b, y) != a / b) ++bad;
        }
        if (div_known(0.0, b, y) != 0.0) ++bad;
        if (div_known(b, b, y) != 1.0) ++bad;
    }
    return bad;
}
