/*
 * oracle/oracle_math.h -- TEST INFRASTRUCTURE ONLY (CPU oracle; never linked into the product).
 *
 * Elementary functions of the "deterministic float32 arithmetic contract" (DESIGN.md section 3).
 * The reference computes exp/log/sin/cos/tan/atan/atan2 through XLA (third-party, absent), whose
 * float32 polynomials are not reproducible offline.  The oracle therefore fixes ONE float32
 * algorithm per function, built only from IEEE-754 round-to-nearest +,-,*,/,sqrt,fma and integer
 * bit operations, so that a second implementation (the CUDA kernels) can reproduce every bit.
 * The algorithms are the classic Cody-Waite / minimax single-precision forms (Cephes family);
 * tests/test_cpu_oracle.py bounds their error against glibc (<= 3 ulp; 4 ulp for sin/cos of |x| up to 100).
 *
 * Compile with -ffp-contract=off: every fused multiply-add below is an explicit fmaf().
 */
#ifndef ORACLE_MATH_H
#define ORACLE_MATH_H
#include <math.h>
#include <stdint.h>
#include <string.h>

static inline uint32_t om_f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float om_u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

/* exp(x): x clamped to [-87, 88] (so exp(x < -87) = exp(-87) ~ 1.6e-38, exp(x > 88) = exp(88)). */
static inline float om_exp(float x) {
    if (x != x) return x;
    if (x < -87.0f) x = -87.0f;
    if (x > 88.0f) x = 88.0f;
    const float MAGIC = 12582912.0f; /* 1.5 * 2^23: adding it rounds to nearest integer */
    float t = fmaf(x, 1.44269504088896341f, MAGIC);
    float nf = t - MAGIC;
    int32_t n = (int32_t)om_f2u(t) - (int32_t)om_f2u(MAGIC);
    float r = fmaf(nf, -0.693359375f, x);
    r = fmaf(nf, 2.12194440e-4f, r);
    float p = 1.9875691500e-4f;
    p = fmaf(p, r, 1.3981999507e-3f);
    p = fmaf(p, r, 8.3334519073e-3f);
    p = fmaf(p, r, 4.1665795894e-2f);
    p = fmaf(p, r, 1.6666665459e-1f);
    p = fmaf(p, r, 5.0000001201e-1f);
    float z = r * r;
    float y = fmaf(p, z, r) + 1.0f;
    /* scale by 2^n in two steps so that n = -126..127 never leaves the normal range */
    return y * om_u2f((uint32_t)(n + 127) << 23);
}

/* Laplace kernel entry k(d; sigma) = 2^(-(d * s2)) of the reduced-set inner CEM (the contract's form of exp(-d / sigma), kernel_computation.py:31-37;
 * DESIGN.md section 3, D1 -- restated, csrc/dmath.cuh dm::lap_ is the same operations):
 *   per bandwidth: rinv = 1 / sigma, s2 = rinv * log2(e), dcap = 125 / s2;
 *   per distance:  dc = min(d, dcap) (NaN propagates), t = fma(dc, -s2, MAGIC), n = bits(t) - bits(MAGIC), f = fma(dc, -s2, MAGIC - t) in [-1/2, 1/2],
 *                  p = P6(f) ~ 2^f (degree 6, constant term 1, interpolating 2^f at the Chebyshev extrema of the interval: tools/lap_poly.py), k = p * 2^n.  <= 1 ulp of 2^(-(d s2)); k(0) = 1 exactly. */
typedef struct { float ns2, dcap; } om_lapscale_t;
static inline om_lapscale_t om_lap_scale(float sigma) {
    float rinv = 1.0f / sigma;
    float s2 = rinv * 1.44269504088896341f;
    om_lapscale_t L; L.ns2 = -s2; L.dcap = 125.0f / s2;
    return L;
}
static inline float om_lap(float d, om_lapscale_t L) {
    const float MAGIC = 12582912.0f;
    float dc = (d != d || L.dcap != L.dcap) ? NAN : (d < L.dcap ? d : L.dcap);      /* min.NaN.f32 */
    float t = fmaf(dc, L.ns2, MAGIC);
    float f = fmaf(dc, L.ns2, MAGIC - t);
    float p = 0.00015469731971976444f;
    p = fmaf(p, f, 0.0013410000965866657f);
    p = fmaf(p, f, 0.009618030782528724f);
    p = fmaf(p, f, 0.05550297314200181f);
    p = fmaf(p, f, 0.24022651084117067f);
    p = fmaf(p, f, 0.6931472253950105f);
    p = fmaf(p, f, 1.0f);
    return p * om_u2f((om_f2u(t) << 23) + 0x3f800000u);
}

/* log(x): x<0 -> NaN, 0 -> -inf, inf -> inf, denormals handled by pre-scaling. */
static inline float om_log(float x) {
    if (x != x) return x;
    if (x < 0.0f) return NAN;
    if (x == 0.0f) return -INFINITY;
    if (x == INFINITY) return x;
    int32_t e = 0;
    if (x < 1.17549435e-38f) { x = x * 8388608.0f; e = -23; }
    uint32_t u = om_f2u(x);
    e += (int32_t)((u >> 23) & 0xffu) - 126;
    float m = om_u2f((u & 0x807fffffu) | 0x3f000000u); /* [0.5, 1) */
    if (m < 0.707106781186547524f) { e -= 1; m = (m + m) - 1.0f; } else { m = m - 1.0f; }
    float z = m * m;
    float p = 7.0376836292e-2f;
    p = fmaf(p, m, -1.1514610310e-1f);
    p = fmaf(p, m, 1.1676998740e-1f);
    p = fmaf(p, m, -1.2420140846e-1f);
    p = fmaf(p, m, 1.4249322787e-1f);
    p = fmaf(p, m, -1.6668057665e-1f);
    p = fmaf(p, m, 2.0000714765e-1f);
    p = fmaf(p, m, -2.4999993993e-1f);
    p = fmaf(p, m, 3.3333331174e-1f);
    float fe = (float)e;
    float y = (m * z) * p;
    y = fmaf(fe, -2.12194440e-4f, y);
    y = fmaf(-0.5f, z, y);
    float r = m + y;
    return fmaf(fe, 0.693359375f, r);
}

/* log1p(x) = log(u) * x / (u - 1), u = 1 + x  (exact-compensation form). */
static inline float om_log1p(float x) {
    float u = 1.0f + x;
    if (u == 1.0f) return x;
    return (om_log(u) * x) / (u - 1.0f);
}

/* shared Cody-Waite reduction to r in [-pi/4, pi/4] and octant j (even), for |x| < 8192 */
static inline float om_trig_reduce(float ax, int32_t *jout) {
    int32_t j = (int32_t)(ax * 1.27323954473516f);
    if (j & 1) j += 1;
    float y = (float)j;
    float r = fmaf(y, -0.78515625f, ax);
    r = fmaf(y, -2.4187564849853515625e-4f, r);
    r = fmaf(y, -3.77489497744594108e-8f, r);
    *jout = j;
    return r;
}
static inline float om_sin_poly(float r) {
    float z = r * r;
    float p = -1.9515295891e-4f;
    p = fmaf(p, z, 8.3321608736e-3f);
    p = fmaf(p, z, -1.6666654611e-1f);
    return fmaf(p * z, r, r);
}
static inline float om_cos_poly(float r) {
    float z = r * r;
    float p = 2.443315711809948e-5f;
    p = fmaf(p, z, -1.388731625493765e-3f);
    p = fmaf(p, z, 4.166664568298827e-2f);
    return fmaf(p, z * z, fmaf(-0.5f, z, 1.0f));
}
static inline float om_sin(float x) {
    if (x != x || fabsf(x) == INFINITY) return NAN;
    int32_t j; float ax = fabsf(x);
    float r = om_trig_reduce(ax, &j);
    int neg = x < 0.0f;
    j &= 7;
    if (j > 3) { neg = !neg; j -= 4; }
    float y = (j == 2) ? om_cos_poly(r) : om_sin_poly(r); /* j in {0,2} (j is even) */
    return neg ? -y : y;
}
static inline float om_cos(float x) {
    if (x != x || fabsf(x) == INFINITY) return NAN;
    int32_t j; float ax = fabsf(x);
    float r = om_trig_reduce(ax, &j);
    int neg = 0;
    j &= 7;
    if (j > 3) { neg = !neg; j -= 4; }
    if (j > 1) neg = !neg;
    float y = (j == 2) ? om_sin_poly(r) : om_cos_poly(r);
    return neg ? -y : y;
}
static inline float om_tan(float x) {
    if (x != x || fabsf(x) == INFINITY) return NAN;
    int32_t j; float ax = fabsf(x);
    float r = om_trig_reduce(ax, &j);
    float z = r * r;
    float p = 9.38540185543e-3f;
    p = fmaf(p, z, 3.11992232697e-3f);
    p = fmaf(p, z, 2.44301354525e-2f);
    p = fmaf(p, z, 5.34112807005e-2f);
    p = fmaf(p, z, 1.33387994085e-1f);
    p = fmaf(p, z, 3.33331568548e-1f);
    float y = fmaf(p * z, r, r);
    if (j & 2) y = -1.0f / y;
    return (x < 0.0f) ? -y : y;
}
static inline float om_atan(float x) {
    if (x != x) return x;
    float ax = fabsf(x), y0;
    if (ax > 2.414213562373095f) { y0 = 1.5707963267948966f; ax = -1.0f / ax; }
    else if (ax > 0.4142135623730950f) { y0 = 0.7853981633974483f; ax = (ax - 1.0f) / (ax + 1.0f); }
    else y0 = 0.0f;
    float z = ax * ax;
    float p = 8.05374449538e-2f;
    p = fmaf(p, z, -1.38776856032e-1f);
    p = fmaf(p, z, 1.99777106478e-1f);
    p = fmaf(p, z, -3.33329491539e-1f);
    float y = y0 + fmaf(p * z, ax, ax);
    return (x < 0.0f) ? -y : y;
}
/* atan2(y, x) with numpy/XLA quadrant conventions (atan2(0,0) = 0; signed zeros ignored). */
static inline float om_atan2(float y, float x) {
    if (x != x || y != y) return NAN;
    const float PI = 3.14159265358979323846f, PIO2 = 1.5707963267948966f;
    if (x == 0.0f) {
        if (y > 0.0f) return PIO2;
        if (y < 0.0f) return -PIO2;
        return 0.0f;
    }
    float z = om_atan(y / x);
    if (x < 0.0f) return (y < 0.0f) ? z - PI : z + PI;
    return z;
}
#endif
