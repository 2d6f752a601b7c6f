/*
 * oracle/oracle_rng.h -- TEST INFRASTRUCTURE ONLY (CPU oracle; never linked into the product).
 *
 * Restatement of the JAX 0.3.23 PRNG protocol the reference relies on.  JAX is a third-party
 * dependency that is NOT vendored under /root/reference (requirements.txt:1 pins jax==0.3.23,
 * jaxlib unpinned), so this file restates its published algorithm and is pinned by known-answer
 * tests (tests/test_oracle_rng.py): Random123 Threefry-2x32 KAT, jax.random.split(PRNGKey(0)),
 * jax.random.normal values from the JAX documentation.  The gamma/beta sampler has no offline KAT
 * ("parity unpinned" for beta noise): it follows jax/_src/random.py::_gamma_one of 0.3.23.
 *
 * Reference call sites (S/ = synthetic_static_obs/):
 *   PRNGKey            S/optimizer/cem.py:114,225  S/optimizer/cem_helper.py:86  S/compute_beta.py:25
 *   split              S/optimizer/cem.py:254,302  S/optimizer/cem_helper.py:125,409,430,438,473,495,503
 *                      S/compute_beta.py:44,54,108,131
 *   multivariate_normal S/optimizer/cem_helper.py:126,292,406,411,439,470,475,504  S/compute_beta.py:46,63
 *   beta               S/optimizer/cem_helper.py:427,432,492,497
 */
#ifndef ORACLE_RNG_H
#define ORACLE_RNG_H
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "oracle_math.h"

typedef struct { uint32_t k0, k1; } okey_t;

/* jax.random.PRNGKey(seed) for a 32-bit seed: [seed >> 32, seed & 0xffffffff] = [0, seed]. */
static inline okey_t rng_key(uint32_t seed) { okey_t k = {0u, seed}; return k; }

static inline uint32_t rotl32(uint32_t x, int d) { return (x << d) | (x >> (32 - d)); }

/* Threefry-2x32, 20 rounds (jax/_src/prng.py::_threefry2x32_lowering; Random123). */
static inline void threefry2x32(okey_t key, uint32_t x0, uint32_t x1, uint32_t *o0, uint32_t *o1) {
    static const int R0[4] = {13, 15, 26, 6}, R1[4] = {17, 29, 16, 24};
    uint32_t ks[3] = {key.k0, key.k1, key.k0 ^ key.k1 ^ 0x1BD11BDAu};
    x0 += ks[0]; x1 += ks[1];
    for (int i = 0; i < 5; i++) {
        const int *R = (i & 1) ? R1 : R0;
        for (int r = 0; r < 4; r++) { x0 += x1; x1 = rotl32(x1, R[r]); x1 ^= x0; }
        x0 += ks[(i + 1) % 3];
        x1 += ks[(i + 2) % 3] + (uint32_t)(i + 1);
    }
    *o0 = x0; *o1 = x1;
}

/* threefry_random_bits(key, 32, (n,)): counters iota(n), padded to even, first half / second half
 * are the two Threefry words; outputs concatenated; pad dropped (prng.py::threefry_2x32). */
static inline void rng_bits(okey_t key, size_t n, uint32_t *out) {
    size_t np2 = n + (n & 1), half = np2 / 2;
    for (size_t i = 0; i < half; i++) {
        uint32_t c0 = (uint32_t)i, c1 = (uint32_t)(i + half);
        if (i + half >= n) c1 = 0u; /* the pad element is a literal 0 */
        uint32_t a, b;
        threefry2x32(key, c0, c1, &a, &b);
        out[i] = a;
        if (i + half < n) out[i + half] = b;
    }
}

/* random.split(key, m) = bits(key, 2m).reshape(m, 2). */
static inline void rng_split(okey_t key, int m, okey_t *out) {
    uint32_t stackbuf[16];
    uint32_t *b = (m <= 8) ? stackbuf : (uint32_t *)malloc(sizeof(uint32_t) * 2 * (size_t)m);
    rng_bits(key, 2 * (size_t)m, b);
    for (int i = 0; i < m; i++) { out[i].k0 = b[2 * i]; out[i].k1 = b[2 * i + 1]; }
    if (b != stackbuf) free(b);
}
/* the reference always keeps row 0:  key, _ = split(key) */
static inline okey_t rng_split0(okey_t key) { okey_t o[2]; rng_split(key, 2, o); return o[0]; }

static inline float bits_to_unit_float(uint32_t b) {
    uint32_t u = (b >> 9) | 0x3F800000u; float f; memcpy(&f, &u, 4); return f - 1.0f;
}
/* random.uniform(key, (n,), f32, minval, maxval): max(minval, f*(maxval-minval)+minval). */
static inline void rng_uniform(okey_t key, size_t n, float minval, float maxval, float *out) {
    uint32_t *b = (uint32_t *)malloc(sizeof(uint32_t) * (n ? n : 1));
    rng_bits(key, n, b);
    float scale = maxval - minval;
    for (size_t i = 0; i < n; i++) {
        float t = bits_to_unit_float(b[i]) * scale; /* compiled with -ffp-contract=off */
        float v = t + minval;
        out[i] = v > minval ? v : minval;
    }
    free(b);
}

/* XLA ErfInv32 (xla/client/lib/math.cc): Giles' single-precision polynomial, w = -log1p(-x*x). */
static inline float xla_erfinv32(float x) {
    static const float lt5[9] = {2.81022636e-08f, 3.43273939e-07f, -3.5233877e-06f, -4.39150654e-06f,
                                 0.00021858087f, -0.00125372503f, -0.00417768164f, 0.246640727f, 1.50140941f};
    static const float ge5[9] = {-0.000200214257f, 0.000100950558f, 0.00134934322f, -0.00367342844f,
                                 0.00573950773f, -0.0076224613f, 0.00943887047f, 1.00167406f, 2.83297682f};
    if (fabsf(x) == 1.0f) return x * INFINITY;
    float xx = x * x;
    float w = -om_log1p(-xx);
    const float *c = (w < 5.0f) ? lt5 : ge5;
    w = (w < 5.0f) ? (w - 2.5f) : (sqrtf(w) - 3.0f);
    float p = c[0];
    for (int i = 1; i < 9; i++) { float t = p * w; p = c[i] + t; }
    return p * x;
}

/* random.normal(key, (n,), f32) = sqrt(2) * erf_inv(uniform(lo=nextafter(-1,0), hi=1)). */
static inline void rng_normal(okey_t key, size_t n, float *out) {
    const float lo = nextafterf(-1.0f, 0.0f);
    rng_uniform(key, n, lo, 1.0f, out);
    const float s2 = (float)1.4142135623730951;
    for (size_t i = 0; i < n; i++) out[i] = s2 * xla_erfinv32(out[i]);
}
/* scalar draws (shape ()): bits(key, 1) = first word of threefry(key, (0, 0)) -- no heap traffic */
static inline uint32_t rng_bits1(okey_t key) { uint32_t a, b; threefry2x32(key, 0u, 0u, &a, &b); return a; }
static inline float rng_uniform1(okey_t key) {
    float v = bits_to_unit_float(rng_bits1(key)) * 1.0f + 0.0f;
    return v > 0.0f ? v : 0.0f;
}
static inline float rng_normal1(okey_t key) {
    const float lo = nextafterf(-1.0f, 0.0f);
    float t = bits_to_unit_float(rng_bits1(key)) * (1.0f - lo);
    float v = t + lo;
    v = v > lo ? v : lo;
    return (float)1.4142135623730951 * xla_erfinv32(v);
}

/* jax/_src/random.py::_gamma_one(key, alpha, log_space=True) of 0.3.23 (Marsaglia-Tsang with
 * the alpha<1 boost in log space).  Returns log(Gamma(alpha) sample). */
static inline float rng_loggamma_one(okey_t key, float alpha) {
    const float one_over_three = (float)(1.0 / 3.0), squeeze_const = 0.0331f;
    int boost_mask = alpha >= 1.0f;
    float alpha_orig = alpha;
    alpha = boost_mask ? alpha : alpha + 1.0f;
    float d = alpha - one_over_three;
    float c = one_over_three / sqrtf(d);
    okey_t ks[2]; rng_split(key, 2, ks);           /* key, subkey = _split(key) */
    key = ks[0];
    okey_t subkey = ks[1];
    float X = 0.0f, V = 1.0f, U = 2.0f;            /* initial state makes _cond_fn true */
    for (;;) {
        float xx = squeeze_const * (X * X);
        int c1 = U >= 1.0f - xx;
        float t1 = X * 0.5f, t2 = d * ((1.0f - V) + om_log(V));
        int c2 = om_log(U) >= t1 + t2;
        if (!(c1 && c2)) break;
        okey_t k3[3]; rng_split(key, 3, k3);       /* key, x_key, U_key = _split(key, 3) */
        key = k3[0];
        okey_t xk = k3[1];
        float x = 0.0f, v = -1.0f;
        while (v <= 0.0f) {                        /* _next_kxv */
            okey_t s[2]; rng_split(xk, 2, s);
            xk = s[0];
            x = rng_normal1(s[1]);
            float xc = x * c;
            v = 1.0f + xc;
        }
        X = x * x;
        float vv = v * v;
        V = vv * v;
        U = rng_uniform1(k3[2]);
    }
    /* log_samples = -exponential(subkey) = log1p(-uniform(subkey)) */
    float u = rng_uniform1(subkey);
    float log_samples = -(-om_log1p(-u));
    float log_boost;
    if (boost_mask || log_samples == 0.0f) log_boost = 0.0f;
    else log_boost = log_samples * (1.0f / alpha_orig);
    return (om_log(d) + om_log(V)) + log_boost;
}

/* random.beta(key, a, b, shape) with a,b already broadcast to n elements (random.py::_beta):
 * key_a,key_b = split(key); per-element keys = split(key_x, n). */
static inline void rng_beta(okey_t key, const float *a, const float *b, size_t n, float *out) {
    okey_t kab[2]; rng_split(key, 2, kab);
    okey_t *ka = (okey_t *)malloc(sizeof(okey_t) * n), *kb = (okey_t *)malloc(sizeof(okey_t) * n);
    rng_split(kab[0], (int)n, ka);
    rng_split(kab[1], (int)n, kb);
    for (size_t i = 0; i < n; i++) {
        float lga = rng_loggamma_one(ka[i], a[i]);
        float lgb = rng_loggamma_one(kb[i], b[i]);
        float m = lga > lgb ? lga : lgb;
        if (isnan(lga) || isnan(lgb)) m = NAN;
        float ga = om_exp(lga - m), gb = om_exp(lgb - m);
        out[i] = ga / (ga + gb);
    }
    free(ka); free(kb);
}
#endif
