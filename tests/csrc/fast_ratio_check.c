/* Exhaustive CPU check of the division-free arithmetic in csrc/qratio.cu:qratio_from_lcs and
 * csrc/jaccard.cu:div_counts.  Test infrastructure only.
 *
 * The kernels replace  a / b  by  r = RN(1 / b);  q0 = RN(a * r);  q = fma(fma(-q0, b, a), r, q0).
 * This program evaluates that sequence with the C library's correctly rounded fma() and compares
 * it, bit for bit, with the plain IEEE division for EVERY input the kernels' fast paths accept:
 *   1. dist / lensum            for 0 <= dist <= lensum < LEN_LIMIT           (both kernels)
 *   2. (norm_sim * 100) / 100   for every norm_sim = 1 - dist / lensum of 1.   (qratio only)
 * Prints the number of mismatches (0 expected).  Build: gcc -O1 -ffp-contract=off ... -lm */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static double fast_div(double a, double b, double r) {
    const double q0 = a * r;
    return fma(fma(-q0, b, a), r, q0);
}

static int same(double x, double y) { return memcmp(&x, &y, sizeof x) == 0; }

int main(int argc, char **argv) {
    const int limit = argc > 1 ? atoi(argv[1]) : 1024;
    const volatile double hundred = 100.0;
    const double r100 = 1.0 / hundred;
    long bad1 = 0, bad2 = 0, n = 0;
    for (int u = 1; u < limit; ++u) {
        const volatile double b = (double)u;
        const double r = 1.0 / b;
        for (int i = 0; i <= u; ++i) {
            const volatile double a = (double)i;
            const double want = a / b;
            const double got = fast_div(a, b, r);
            if (!same(want, got)) ++bad1;
            /* QRatio / 100 = ((1.0 - norm_dist) * 100) / 100 */
            const volatile double x = (1.0 - want) * hundred;
            const double want2 = x / hundred;
            const double got2 = fast_div(x, hundred, r100);
            if (!same(want2, got2)) ++bad2;
            ++n;
        }
    }
    printf("%ld %ld %ld\n", n, bad1, bad2);
    return 0;
}
