// Exhaustive check of the table-driven exp used by the CUDA kernels (common.cuh: cr_expf_neg) against the
// oracle's definition (float)exp((double)x) with this host's libm, over EVERY fp32 x in [-16, 0].
// The C code below performs exactly the IEEE double operations of the device function (explicit fma, no
// contraction: build with -ffp-contract=off), so agreement here is agreement on the device.
//   gcc -O2 -fopenmp -ffp-contract=off -mfma tools/exp_check.c -o tools/exp_check -lm && tools/exp_check
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

static const double T32[32] = {
    0x1p+0, 0x1.059b0d3158574p+0, 0x1.0b5586cf9890fp+0, 0x1.11301d0125b51p+0,
    0x1.172b83c7d517bp+0, 0x1.1d4873168b9aap+0, 0x1.2387a6e756238p+0, 0x1.29e9df51fdee1p+0,
    0x1.306fe0a31b715p+0, 0x1.371a7373aa9cbp+0, 0x1.3dea64c123422p+0, 0x1.44e086061892dp+0,
    0x1.4bfdad5362a27p+0, 0x1.5342b569d4f82p+0, 0x1.5ab07dd485429p+0, 0x1.6247eb03a5585p+0,
    0x1.6a09e667f3bcdp+0, 0x1.71f75e8ec5f74p+0, 0x1.7a11473eb0187p+0, 0x1.82589994cce13p+0,
    0x1.8ace5422aa0dbp+0, 0x1.93737b0cdc5e5p+0, 0x1.9c49182a3f09p+0, 0x1.a5503b23e255dp+0,
    0x1.ae89f995ad3adp+0, 0x1.b7f76f2fb5e47p+0, 0x1.c199bdd85529cp+0, 0x1.cb720dcef9069p+0,
    0x1.d5818dcfba487p+0, 0x1.dfc97337b9b5fp+0, 0x1.ea4afa2a490dap+0, 0x1.f50765b6e454p+0};

static inline float exp_tab(float xf) {
    const double x = (double)xf;
    const double z = fma(x, 0x1.71547652b82fep+5, 0x1.8p52);  // k = rint(x * 32/ln2) in the low word
    uint64_t zb;
    memcpy(&zb, &z, 8);
    const int k = (int)(uint32_t)zb;
    const double kd = z - 0x1.8p52;
    double r = fma(kd, -0x1.62e42feep-6, x);
    r = fma(kd, -0x1.a39ef358p-38, r);
    double q = fma(r, 1.0 / 720.0, 1.0 / 120.0);
    q = fma(r, q, 1.0 / 24.0);
    q = fma(r, q, 1.0 / 6.0);
    q = fma(r, q, 0.5);
    const double p = fma(r * r, q, r);
    double s = T32[k & 31];
    uint64_t sb;
    memcpy(&sb, &s, 8);
    sb += (uint64_t)(int64_t)(k >> 5) << 52;
    memcpy(&s, &sb, 8);
    return (float)fma(s, p, s);
}

int main(void) {
    long long bad = 0, n = 0;
    uint32_t lo_bits, hi_bits;
    float lo = -16.0f, hi = -0.0f;
    memcpy(&lo_bits, &lo, 4);
    memcpy(&hi_bits, &hi, 4);
    // negative floats: bit patterns 0x80000000 (-0) .. bits(-16)
#pragma omp parallel for reduction(+ : bad, n) schedule(static, 1 << 20)
    for (long long b = hi_bits; b <= (long long)lo_bits; b++) {
        uint32_t u = (uint32_t)b;
        float x;
        memcpy(&x, &u, 4);
        const float ref = (float)exp((double)x);
        const float got = exp_tab(x);
        n++;
        if (memcmp(&ref, &got, 4) != 0) {
            bad++;
            if (bad < 20) printf("MISMATCH x=%a ref=%a got=%a\n", x, ref, got);
        }
    }
    printf("checked %lld inputs in [-16, -0], mismatches: %lld\n", n, bad);
    return bad != 0;
}
