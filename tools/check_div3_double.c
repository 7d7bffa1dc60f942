// Check behind hf6d::div3_rn (csrc/gather.cuh): a / 3 in double as q = a * RN(1/3), r = fma(-3, q, a), q' = fma(r, RN(1/3), q)
// equals the IEEE division a / 3.0 bit for bit.  The arguments the gather kernel feeds it are squares of floats (exact in
// double, at most 48 significant bits): 2 * 10^9 of those from random floats in [1e-8, 4) plus 10^9 random doubles.
//   gcc -O2 -fopenmp -mfma -ffp-contract=off -o /tmp/div3check tools/check_div3_double.c -lm && /tmp/div3check
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <omp.h>
static inline uint64_t mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
int main() {
    const double y = 0x1.5555555555555p-2;
    long long bad = 0;
#pragma omp parallel for reduction(+ : bad) schedule(static)
    for (long long i = 0; i < 3000000000LL; ++i) {
        const uint64_t z = mix64((uint64_t)i * 0x9E3779B97F4A7C15ULL + 1);
        double a;
        if (i < 2000000000LL) {
            uint32_t u = (uint32_t)z;
            u = (u & 0x007FFFFFu) | ((uint32_t)(100 + (z >> 40) % 29) << 23);  /* exponent 2^-27 .. 2^1 */
            float f;
            memcpy(&f, &u, 4);
            a = (double)f * (double)f;
        } else {
            uint64_t u = (z & 0x000FFFFFFFFFFFFFULL) | ((uint64_t)(1023 - 60 + (z >> 53) % 64) << 52);
            memcpy(&a, &u, 8);
        }
        const double q = a * y;
        const double q2 = fma(fma(-3.0, q, a), y, q);
        const double ref = a / 3.0;
        if (memcmp(&q2, &ref, 8) != 0) ++bad;
    }
    printf("a/3: %lld mismatches in 3e9 samples\n", bad);
    return bad != 0;
}
