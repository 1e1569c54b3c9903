// Randomised check of common.cuh div_by(): (float)((double)a * (1.0 / (double)b)) == a / b (IEEE fp32 division)
// for finite operands of moderate magnitude (the proof is in common.cuh; this is the belt to its braces).
//   gcc -O2 -fopenmp -ffp-contract=off tools/divby_check.c -o tools/divby_check -lm && tools/divby_check
#include <math.h>
#include <stdlib.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
static inline uint64_t rng(uint64_t *s) { *s ^= *s << 13; *s ^= *s >> 7; *s ^= *s << 17; return *s; }
int main(int argc, char **argv) {
    long long bad = 0, n = 0;
    const long long per_thread = argc > 1 ? atoll(argv[1]) : 400000000LL;
#pragma omp parallel reduction(+ : bad, n)
    {
        uint64_t s = 88172645463325252ULL;
#ifdef _OPENMP
        extern int omp_get_thread_num(void);
        s += 0x9E3779B97F4A7C15ULL * (uint64_t)(omp_get_thread_num() + 1);
#endif
        for (long long it = 0; it < per_thread; it++) {
            uint64_t r = rng(&s);
            uint32_t ua = (uint32_t)r, ub = (uint32_t)(r >> 32);
            // exponents limited to 2^-40 .. 2^40, random sign and significand
            ua = (ua & 0x807fffffu) | ((uint32_t)(87 + (ua >> 23) % 81) << 23);
            ub = (ub & 0x807fffffu) | ((uint32_t)(87 + (ub >> 23) % 81) << 23);
            float a, b;
            memcpy(&a, &ua, 4);
            memcpy(&b, &ub, 4);
            volatile float ref = a / b;
            const double inv = 1.0 / (double)b;
            const float got = (float)((double)a * inv);
            n++;
            if (ref != got) { bad++; if (bad < 10) printf("MISMATCH a=%a b=%a ref=%a got=%a\n", a, b, ref, got); }
        }
    }
    printf("checked %lld divisions, mismatches: %lld\n", n, bad);
    return bad != 0;
}
