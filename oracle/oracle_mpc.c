/*
 * oracle/oracle_mpc.c -- TEST INFRASTRUCTURE ONLY (CPU oracle; never linked into the product).
 *
 * Plain-C float32 restatement of the MPC-MMD trajectory-optimizer inner loop
 * (reference: Basant1861/MPC-MMD, S/ = synthetic_static_obs/, D/ = synthetic_dynamic_obs/):
 *     S/optimizer/cem.py            CEM.compute_cem_{mmd_opt,mmd_random,cvar,saa}     (:201-714)
 *     S/optimizer/cem_helper.py     Helper.*                                          (:122-564)
 *     S/optimizer/projection.py     Projection.compute_projection                     (:276-323)
 *     S/optimizer/costs.py          Costs.*                                           (:50-234)
 *     S/kernel_computation.py       kernel_matrix.compute_mmd / compute_kernel        (:19-87)
 *     S/compute_beta.py             beta_cem.compute_cem                              (:93-157)
 * Each function below cites the lines it follows.
 *
 * PARITY PINS.  The reference has no tests or golden vectors and its runtime
 * (jax==0.3.23 + jaxlib, XLA, LAPACK) cannot be installed here, so the oracle is pinned by
 *   (a) RNG known answers (Random123 Threefry KAT, JAX-docs split/normal values) -- tests/test_cpu_oracle.py,
 *   (b) the Bernstein basis computed by the reference's own bernstein_coeff_order10_arbitinterval.py,
 *   (c) tests/golden/ref_stages.npz: inputs/outputs of EVERY stage method of the reference's optimizer,
 *       recorded while its own unmodified source files ran on a NumPy float32 stand-in for the JAX API
 *       (tests/golden/jax_shim, tests/golden/make_golden_ref.py); tests/test_reference_stages.py holds the
 *       oracle (and the CUDA stage entry points) to 1e-4 relative on those vectors.
 * Still restated, not pinned: the third-party JAX runtime itself -- XLA's float32 kernels (covered only
 * to round-off by (c)) and jax.random.beta's Marsaglia-Tsang rejection sampler, which (c) exercises
 * through this oracle's own restatement (oracle_rng.h).  "Parity unpinned" applies to that sampler only.
 *
 * Arithmetic contract: every operation is IEEE float32 round-to-nearest in
 * the association written here; fused multiply-adds appear only as explicit fmaf(); long sums use
 * either an ascending sequential chain or the 32-partial "lane" order of lane_sum_sq(); the
 * transcendental functions are those of oracle_math.h.  Build with -ffp-contract=off.
 * Documented deviations from the literal reference formulation (all below float32 round-off of
 * the reference's own solves):
 *   D1  inner-CEM Laplace kernel uses exp(-(d * (1/sigma))) instead of exp(-d/sigma)
 *   D2  constant linear solves (x_guess KKT, projection KKT, ridge fit) use matrices inverted once
 *       in float64 on the host and rounded to float32
 *   D3  the (nr+1)x(nr+1) beta KKT system is solved by Cholesky block elimination, not LU
 *   D4  0*cost_des_lane and 0*cost_lane terms (exact zeros for finite inputs) are dropped
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include "oracle_math.h"
#include "oracle_rng.h"

#define T_ 100   /* num,  S/optimizer/cem.py:38 */
#define NV 11    /* nvar, S/optimizer/cem.py:50 */
#define NL 99    /* num-1 lane rows per side, S/optimizer/cem.py:126-134 */
#define NP_ 8    /* num_params, S/optimizer/cem.py:136 */

enum { COST_MMD_OPT = 0, COST_MMD_RANDOM = 1, COST_CVAR = 2, COST_SAA = 3 };
enum { NOISE_GAUSSIAN = 0, NOISE_BETA = 1 };

typedef struct {
    int32_t B;            /* num_batch            cem.py:137 */
    int32_t np;           /* num_prime            cem.py:52  */
    int32_t nr;           /* num_reduced          cem.py:142 */
    int32_t nm;           /* num_mother = nr^2    cem.py:143 */
    int32_t O;            /* num_obs */
    int32_t iters;        /* maxiter_cem = 20     cem.py:89  */
    int32_t n_el;         /* ellite_num = 5       cem.py:138 */
    int32_t n_el_cost;    /* ellite_num_cost = 20 cem.py:140 */
    int32_t noise_kind;
    int32_t S_in;         /* num_samples_cem = 100   compute_beta.py:14 */
    int32_t iters_in;     /* maxiter_beta_cem = 20   compute_beta.py:15 */
    int32_t n_el_in;      /* num_ellite_beta = 11    compute_beta.py:26 */
    int32_t naive;        /* 1: recompute the 22-dim L1 distances for every beta sample like the
                             reference does (compute_beta.py:120-127); 0: look them up in the
                             per-chain nm x nm table.  Bit-identical results either way. */
    int32_t pad_;
    float sigma_acc, sigma_steer, ksig_steer, acc_const, steer_const, beta_a, beta_b;
    float v_min, v_max, a_max, b_lane_ub, b_lane_lb, y_lb, y_ub, a2_obs, b2_obs;
    float wheel_base, dt, steer_max, steer_rate_pen, alpha_quant, ker_wt;
    float w_obs;          /* weight_{mmd,cvar,saa}_obs  cem.py:161-163 */
    float lam_inv;        /* 1/lamda,  cem_helper.py:285 */
    float one_m_alpha_mean, alpha_mean, one_m_alpha_cov, alpha_cov;
    float sigma_clip, inv_nm, m2_inv_nm, beta_del, sigma_random;
    const float *P, *Pd, *Pdd;      /* (100,11) row-major */
    const float *Gx, *Gy;           /* x_guess affine maps (11,7), (11,8)      [D2] */
    const float *Kx, *Ky;           /* projection KKT inverse rows (11,14), (11,15) [D2] */
    const float *Wfit;              /* (11,np) ridge fit (P'^T P' + .05 I)^-1 P'^T  [D2] */
    const float *z_init;            /* (B,8)   normal(split0(PRNGKey(0)))   cem_helper.py:125-126 */
    const float *theta0;            /* (S_in,nm+1) initial inner samples     compute_beta.py:41-49 */
    const float *zb_iter;           /* (iters_in, S_in-n_el_in, nm+1)        compute_beta.py:131,63 */
} ocfg_t;

/* ------------------------------------------------------------------------------------------ */
/* small helpers of the arithmetic contract                                                    */

static inline float clipf(float x, float lo, float hi) { /* jnp.clip = minimum(maximum(x,lo),hi) */
    float m = (x != x) ? x : (x > lo ? x : lo);
    return (m != m) ? m : (m < hi ? m : hi);
}
static inline float nmax0(float x) { return (x != x) ? x : (x > 0.0f ? x : 0.0f); } /* jnp.maximum(0,x) */
static inline float nmaxf(float a, float b) { if (a != a) return a; if (b != b) return b; return a > b ? a : b; }

/* sum of squares of the concatenation of up to 3 arrays in "lane order": partial l takes the
 * elements l, l+32, ... of each array in turn (fma accumulate), then a xor-butterfly 16,8,4,2,1. */
static float lane_sum_sq(const float *a0, int n0, const float *a1, int n1) {
    float p[32], q[32];
    for (int l = 0; l < 32; l++) {
        float acc = 0.0f;
        for (int i = l; i < n0; i += 32) acc = fmaf(a0[i], a0[i], acc);
        for (int i = l; i < n1; i += 32) acc = fmaf(a1[i], a1[i], acc);
        p[l] = acc;
    }
    for (int off = 16; off >= 1; off >>= 1) {
        for (int l = 0; l < 32; l++) q[l] = p[l] + p[l ^ off];
        memcpy(p, q, sizeof p);
    }
    return p[0];
}
static inline float lane_norm(const float *a0, int n0, const float *a1, int n1) {
    return sqrtf(lane_sum_sq(a0, n0, a1, n1));
}
static inline float dot11(const float *row, const float *c) {
    float acc = 0.0f;
    for (int k = 0; k < NV; k++) acc = fmaf(row[k], c[k], acc);
    return acc;
}
/* (M^T r)[j] for M (100,11): ascending-t fma chain */
static inline float dotT(const float *M, int j, const float *r) {
    float acc = 0.0f;
    for (int t = 0; t < T_; t++) acc = fmaf(M[t * NV + j], r[t], acc);
    return acc;
}
/* (A_lane_bound^T w)[j], A_lane_bound = [P[1:]; -P[1:]]  (cem.py:126-134, gamma = 1) */
static inline float dot_lane(const float *P, int j, const float *w) {
    float acc = 0.0f;
    for (int i = 0; i < NL; i++) acc = fmaf(P[(i + 1) * NV + j], w[i], acc);
    for (int i = 0; i < NL; i++) acc = fmaf(-P[(i + 1) * NV + j], w[NL + i], acc);
    return acc;
}
/* jnp.unwrap(p) along time, period 2*pi (jax/_src/numpy/lax_numpy.py::unwrap) */
static void unwrap100(const float *p, float *up) {
    const float PI = 3.14159265358979323846f, TWO_PI = 6.28318530717958647692f;
    float cs = 0.0f;
    up[0] = p[0];
    for (int t = 1; t < T_; t++) {
        float dd = p[t] - p[t - 1];
        float corr = 0.0f;
        if (!(fabsf(dd) < PI)) {
            float a = dd + PI;
            float r = fmodf(a, TWO_PI);
            if (r != 0.0f && r < 0.0f) r += TWO_PI;   /* jnp.mod: sign of the divisor */
            float ddmod = r - PI;
            if (ddmod == -PI && dd > 0.0f) ddmod = PI;
            corr = ddmod - dd;
        }
        cs = cs + corr;
        up[t] = p[t] + cs;
    }
}
/* stable ascending argsort, NaN last (jnp.argsort) via ranks */
static inline int lt_nanlast(float a, float b) { return (a == a && b != b) || a < b; }
static void argsort_stable(const float *v, int n, int *perm) {
    /* insertion sort: an element only moves past strictly greater ones => stable */
    for (int i = 0; i < n; i++) {
        int j = i;
        while (j > 0 && lt_nanlast(v[i], v[perm[j - 1]])) { perm[j] = perm[j - 1]; j--; }
        perm[j] = i;
    }
}
/* Cholesky (lower, row-major n x n, leading dim ld) with reciprocal pivots; returns L in place of
 * the lower triangle of A.  acc starts at A[i][j] and subtracts L[i][k]*L[j][k] for ascending k
 * as fma(-L[i][k], L[j][k], acc);  L[j][j] = sqrt(acc), rd = 1/L[j][j], L[i][j] = acc * rd. */
static void chol_inplace(float *A, int n, int ld, float *rd) {
    for (int j = 0; j < n; j++) {
        float acc = A[j * ld + j];
        for (int k = 0; k < j; k++) acc = fmaf(-A[j * ld + k], A[j * ld + k], acc);
        float d = sqrtf(acc);
        A[j * ld + j] = d;
        rd[j] = 1.0f / d;
        for (int i = j + 1; i < n; i++) {
            float a = A[i * ld + j];
            for (int k = 0; k < j; k++) a = fmaf(-A[i * ld + k], A[j * ld + k], a);
            A[i * ld + j] = a * rd[j];
        }
    }
}
/* mean + L z with ascending-k fma chain over k <= i  (jax.random.multivariate_normal, cholesky) */
static inline float mvn_elem(const float *L, int ld, int i, const float *z, float mean) {
    float acc = 0.0f;
    for (int k = 0; k <= i; k++) acc = fmaf(L[i * ld + k], z[k], acc);
    return mean + acc;
}

/* ------------------------------------------------------------------------------------------ */
/* Stage A: x_guess + projection + controls + state-cost terms, one CEM sample                 */
/* cem_helper.py:169-230 (x_guess), projection.py:276-323, cem_helper.py:540-551, :232-262      */

typedef struct {
    float cx[NV], cy[NV];
    float xd[T_], yd[T_], xdd[T_], ydd[T_], y[T_];
    float res_norm;
    float acc[T_], steer[T_];
    float cost_base;      /* compute_cost without the obstacle/lane terms */
} oproj_out_t;

void oracle_project(const ocfg_t *c, const float *param, const float *beq_x, const float *beq_y, float v_des,
                    float *lam_x, float *lam_y, float *s_lane, oproj_out_t *o) {
    const float *P = c->P, *Pd = c->Pd, *Pdd = c->Pdd;
    float cbx[NV], cby[NV];
    /* x_guess (cem_helper.py:169-230) folded to an affine map [D2] */
    {
        float ux[7] = {param[0], param[1], param[2], param[3], beq_x[0], beq_x[1], beq_x[2]};
        float uy[8] = {param[4], param[5], param[6], param[7], beq_y[0], beq_y[1], beq_y[2], beq_y[3]};
        for (int i = 0; i < NV; i++) {
            float a = 0.0f;
            for (int k = 0; k < 7; k++) a = fmaf(c->Gx[i * 7 + k], ux[k], a);
            cbx[i] = a;
            a = 0.0f;
            for (int k = 0; k < 8; k++) a = fmaf(c->Gy[i * 8 + k], uy[k], a);
            cby[i] = a;
        }
    }
    /* projection.py:282-289 guess derivatives (x_guess,y_guess feed only the dead obstacle terms) */
    float xdg[T_], ydg[T_], xddg[T_], yddg[T_];
    for (int t = 0; t < T_; t++) {
        xdg[t] = dot11(Pd + t * NV, cbx);
        ydg[t] = dot11(Pd + t * NV, cby);
        xddg[t] = dot11(Pdd + t * NV, cbx);
        yddg[t] = dot11(Pdd + t * NV, cby);
    }
    /* initial_alpha_d_obs (projection.py:73-119) */
    float av_raw[T_], aa_raw[T_], av[T_], aa[T_], dv[T_], da[T_];
    for (int t = 0; t < T_; t++) { av_raw[t] = om_atan2(ydg[t], xdg[t]); aa_raw[t] = om_atan2(yddg[t], xddg[t]); }
    unwrap100(av_raw, av);
    unwrap100(aa_raw, aa);
    float r_ax[T_], r_ay[T_], r_vx[T_], r_vy[T_];
    for (int t = 0; t < T_; t++) {
        float cv = om_cos(av[t]), sv = om_sin(av[t]);
        float c1 = cv * cv + sv * sv;
        float c2 = xdg[t] * cv + ydg[t] * sv;
        dv[t] = clipf(c2 / c1, c->v_min, c->v_max);
        float ca = om_cos(aa[t]), sa = om_sin(aa[t]);
        c1 = ca * ca + sa * sa;
        c2 = xddg[t] * ca + yddg[t] * sa;
        da[t] = clipf(c2 / c1, 0.0f, c->a_max);
        r_ax[t] = xddg[t] - da[t] * ca;
        r_ay[t] = yddg[t] - da[t] * sa;
        r_vx[t] = xdg[t] - dv[t] * cv;
        r_vy[t] = ydg[t] - dv[t] * sv;
    }
    for (int j = 0; j < NV; j++) {   /* projection.py:115-119 */
        lam_x[j] = (lam_x[j] - dotT(Pdd, j, r_ax)) - dotT(Pd, j, r_vx);
        lam_y[j] = (lam_y[j] - dotT(Pdd, j, r_ay)) - dotT(Pd, j, r_vy);
    }
    /* compute_x (projection.py:123-185) */
    float b_ax[T_], b_ay[T_], b_vx[T_], b_vy[T_], b_aug[2 * NL];
    for (int t = 0; t < T_; t++) {
        float cv = om_cos(av[t]), sv = om_sin(av[t]), ca = om_cos(aa[t]), sa = om_sin(aa[t]);
        b_ax[t] = da[t] * ca; b_ay[t] = da[t] * sa;
        b_vx[t] = dv[t] * cv; b_vy[t] = dv[t] * sv;
    }
    for (int i = 0; i < NL; i++) { b_aug[i] = c->b_lane_ub - s_lane[i]; b_aug[NL + i] = c->b_lane_lb - s_lane[NL + i]; }
    float rhs_x[14], rhs_y[15];
    for (int j = 0; j < NV; j++) {
        float lx = ((-lam_x[j] - cbx[j]) - dotT(Pdd, j, b_ax)) - dotT(Pd, j, b_vx);
        float ly = (((-lam_y[j] - cby[j]) - dotT(Pdd, j, b_ay)) - dotT(Pd, j, b_vy)) - dot_lane(P, j, b_aug);
        rhs_x[j] = -lx;
        rhs_y[j] = -ly;
    }
    for (int k = 0; k < 3; k++) rhs_x[NV + k] = beq_x[k];
    for (int k = 0; k < 4; k++) rhs_y[NV + k] = beq_y[k];
    for (int i = 0; i < NV; i++) {   /* KKT solve with the constant inverse [D2], projection.py:145-171 */
        float a = 0.0f;
        for (int k = 0; k < 14; k++) a = fmaf(c->Kx[i * 14 + k], rhs_x[k], a);
        o->cx[i] = a;
        a = 0.0f;
        for (int k = 0; k < 15; k++) a = fmaf(c->Ky[i * 15 + k], rhs_y[k], a);
        o->cy[i] = a;
    }
    for (int t = 0; t < T_; t++) {   /* projection.py:173-180 */
        o->xd[t] = dot11(Pd + t * NV, o->cx);
        o->xdd[t] = dot11(Pdd + t * NV, o->cx);
        o->y[t] = dot11(P + t * NV, o->cy);
        o->yd[t] = dot11(Pd + t * NV, o->cy);
        o->ydd[t] = dot11(Pdd + t * NV, o->cy);
    }
    float r_lane[2 * NL];
    for (int i = 0; i < NL; i++) {   /* projection.py:182-183 */
        float Ay = o->y[i + 1];
        s_lane[i] = nmax0(-Ay + c->b_lane_ub);
        r_lane[i] = (Ay - c->b_lane_ub) + s_lane[i];
        Ay = -o->y[i + 1];
        s_lane[NL + i] = nmax0(-Ay + c->b_lane_lb);
        r_lane[NL + i] = (Ay - c->b_lane_lb) + s_lane[NL + i];
    }
    /* compute_alph_d (projection.py:217-272): no unwrap here */
    for (int t = 0; t < T_; t++) {
        float a_v = om_atan2(o->yd[t], o->xd[t]);
        float cv = om_cos(a_v), sv = om_sin(a_v);
        float c1 = cv * cv + sv * sv;
        float c2 = o->xd[t] * cv + o->yd[t] * sv;
        float d_v = clipf(c2 / c1, c->v_min, c->v_max);
        float a_a = om_atan2(o->ydd[t], o->xdd[t]);
        float ca = om_cos(a_a), sa = om_sin(a_a);
        c1 = ca * ca + sa * sa;
        c2 = o->xdd[t] * ca + o->ydd[t] * sa;
        float d_a = clipf(c2 / c1, 0.0f, c->a_max);
        r_ax[t] = o->xdd[t] - d_a * ca;
        r_ay[t] = o->ydd[t] - d_a * sa;
        r_vx[t] = o->xd[t] - d_v * cv;
        r_vy[t] = o->yd[t] - d_v * sv;
    }
    o->res_norm = (lane_norm(r_ax, T_, r_ay, T_) + lane_norm(r_vx, T_, r_vy, T_)) + lane_norm(r_lane, 2 * NL, NULL, 0);
    for (int j = 0; j < NV; j++) {   /* projection.py:267-272 */
        lam_x[j] = (lam_x[j] - dotT(Pdd, j, r_ax)) - dotT(Pd, j, r_vx);
        lam_y[j] = ((lam_y[j] - dotT(Pdd, j, r_ay)) - dotT(Pd, j, r_vy)) - dot_lane(P, j, r_lane);
    }
    /* compute_controls (cem_helper.py:540-551); acc[99] = 0 because v is padded with its last value */
    float v[T_], dvv[T_], pen[T_], sv1[T_], sv2[T_], penv[T_];
    for (int t = 0; t < T_; t++) v[t] = sqrtf(o->xd[t] * o->xd[t] + o->yd[t] * o->yd[t]);
    for (int t = 0; t < T_; t++) {
        float vn = (t + 1 < T_) ? v[t + 1] : v[T_ - 1];
        o->acc[t] = (vn - v[t]) / c->dt;
        float s2 = o->xd[t] * o->xd[t] + o->yd[t] * o->yd[t];
        float curv = (o->ydd[t] * o->xd[t] - o->yd[t] * o->xdd[t]) / (s2 * sqrtf(s2));
        o->steer[t] = om_atan(curv * c->wheel_base);
    }
    /* compute_cost terms that do not depend on the risk (cem_helper.py:232-262) [D4] */
    for (int t = 0; t < T_; t++) { dvv[t] = v[t] - v_des; pen[t] = nmax0(fabsf(o->steer[t]) - c->steer_max); }
    for (int t = 0; t < T_ - 1; t++) { sv1[t] = o->steer[t + 1] - o->steer[t]; penv[t] = nmax0(fabsf(sv1[t]) - c->steer_rate_pen); }
    for (int t = 0; t < T_ - 2; t++) sv2[t] = sv1[t + 1] - sv1[t];
    float n_v = lane_norm(dvv, T_, NULL, 0);
    float c_s = lane_norm(o->steer, T_, NULL, 0), c_sv = lane_norm(sv1, T_ - 1, NULL, 0), c_sa = lane_norm(sv2, T_ - 2, NULL, 0);
    float p1 = lane_norm(pen, T_, NULL, 0), p2 = lane_norm(penv, T_ - 1, NULL, 0);
    float n_ydd = lane_norm(o->ydd, T_, NULL, 0), n_xdd = lane_norm(o->xdd, T_, NULL, 0);
    o->cost_base = ((((o->res_norm + 0.1f * n_v) + 0.1f * ((c_s + c_sv) + c_sa)) + 0.1f * (p1 + p2)) + 0.02f * n_ydd) + 0.02f * n_xdd;
}

/* ------------------------------------------------------------------------------------------ */
/* risk functionals                                                                            */

/* costs.py:50-60 + the max over (obs,time) of :178-179 / :211-212 */
static float fbar_max(const ocfg_t *c, const float *xr, const float *yr, const float *x_obs, const float *y_obs) {
    float best = 0.0f; int first = 1;
    for (int o = 0; o < c->O; o++)
        for (int t = 0; t < c->np; t++) {
            float wc = xr[t] - x_obs[o * T_ + t], ws = yr[t] - y_obs[o * T_ + t];
            float cost = (-(wc * wc) / c->a2_obs - (ws * ws) / c->b2_obs) + 1.0f;
            float cb = nmax0(cost);
            best = first ? cb : nmaxf(best, cb);
            first = 0;
        }
    return best;
}
/* costs.py:62-71 + max over time */
static void lane_max(const ocfg_t *c, const float *yr, float *lb, float *ub) {
    float l = 0.0f, u = 0.0f;
    for (int t = 0; t < c->np; t++) {
        float cl = nmax0(-yr[t] + c->y_lb), cu = nmax0(yr[t] - c->y_ub);
        l = t ? nmaxf(l, cl) : cl;
        u = t ? nmaxf(u, cu) : cu;
    }
    *lb = l; *ub = u;
}
/* kernel_computation.py:67-87 compute_mmd (Laplace kernel :31-39) */
static float mmd_cost(const ocfg_t *c, const float *beta, const float *cost, float sigma) {
    int nr = c->nr;
    float s1 = 0.0f, s2 = 0.0f;
    for (int i = 0; i < nr; i++) {
        float t = 0.0f;
        for (int j = 0; j < nr; j++) t = fmaf(om_exp(-fabsf(cost[i] - cost[j]) / sigma), beta[j], t);
        s1 = fmaf(beta[i], t, s1);
        float e = om_exp(-fabsf(cost[i] - 0.0f) / sigma), u = 0.0f;
        for (int j = 0; j < nr; j++) u = fmaf(e, c->beta_del, u);
        s2 = fmaf(beta[i], u, s2);
    }
    return c->ker_wt * (s1 - 2.0f * s2);
}
/* jnp.quantile(v, q) linear interpolation + the CVaR mean of costs.py:213-220 */
static float cvar_cost(const ocfg_t *c, const float *v) {
    int nr = c->nr, perm[64];
    argsort_stable(v, nr, perm);
    float q = c->alpha_quant * (float)(nr - 1);
    float lo = floorf(q), hi = ceilf(q);
    float hw = q - lo, lw = 1.0f - hw;
    int ilo = (int)lo, ihi = (int)hi;
    if (ilo < 0) ilo = 0; if (ilo > nr - 1) ilo = nr - 1;
    if (ihi < 0) ihi = 0; if (ihi > nr - 1) ihi = nr - 1;
    float var = v[perm[ilo]] * lw + v[perm[ihi]] * hw;
    float s = 0.0f; int n = 0;
    for (int i = 0; i < nr; i++) if (v[i] >= var) { s = s + v[i]; n++; }
    return n > 0 ? s / (float)n : 0.0f;
}
static float saa_cost(const ocfg_t *c, const float *v) {
    float s = 0.0f;
    for (int i = 0; i < c->nr; i++) s = s + (v[i] > 0.0f ? 1.0f : 0.0f);
    return s / (float)c->nr;
}

/* ------------------------------------------------------------------------------------------ */
/* noisy rollouts (cem_helper.py:380-400, 402-538)                                             */

typedef struct {           /* per (episode, outer iteration) noise, shared by all B samples [Q5] */
    const float *z1, *z2, *z3;   /* (nr,np) normals: acc, steer, common-mode */
    okey_t k1, k2;               /* beta noise keys (cem_helper.py:427,432 / 492,497) */
    const float *b1, *b2;        /* optional (nr,np): the two jax.random.beta draws themselves, INJECTED (NULL = draw them from k1 / k2).  Lets a machine
                                    that has the reference's jax==0.3.23 feed its own samples and so take the restated sampler out of the comparison */
} onoise_t;

static void rollout_one(const ocfg_t *c, const float *a, const float *s, const float *st0, float *xr, float *yr) {
    float x = st0[0], y = st0[1], vx = st0[2], vy = st0[3], psi = st0[4];
    for (int t = 0; t < c->np; t++) {     /* cem_helper.py:451-458: record the state before the step [Q11] */
        xr[t] = x; yr[t] = y;
        float v = sqrtf(vx * vx + vy * vy);     /* cem_helper.py:380-400 */
        v = v + a[t] * c->dt;
        float psidot = (v * om_tan(s[t])) / c->wheel_base;
        psi = psi + psidot * c->dt;
        vx = v * om_cos(psi);
        vy = v * om_sin(psi);
        x = x + vx * c->dt;
        y = y + vy * c->dt;
    }
}
/* perturbed controls, (nr,np) each */
static void noisy_controls(const ocfg_t *c, const float *acc, const float *steer, const onoise_t *nz, float *an, float *sn) {
    int nr = c->nr, np = c->np, n = nr * np;
    float *pa = (float *)malloc(sizeof(float) * n), *ps = (float *)malloc(sizeof(float) * n);
    if (c->noise_kind == NOISE_GAUSSIAN) {   /* cem_helper.py:405-415 */
        for (int r = 0; r < nr; r++)
            for (int t = 0; t < np; t++) {
                pa[r * np + t] = (c->sigma_acc * fabsf(acc[t])) * nz->z1[r * np + t];
                ps[r * np + t] = (c->sigma_steer * fabsf(steer[t])) * nz->z2[r * np + t];
            }
    } else {                                 /* cem_helper.py:427-436 */
        float *a = (float *)malloc(sizeof(float) * n), *b = (float *)malloc(sizeof(float) * n);
        float *smp = (float *)malloc(sizeof(float) * n);
        for (int r = 0; r < nr; r++) for (int t = 0; t < np; t++) { a[r * np + t] = c->beta_a * fabsf(acc[t]); b[r * np + t] = c->beta_b * fabsf(acc[t]); }
        if (nz->b1) memcpy(smp, nz->b1, sizeof(float) * n); else rng_beta(nz->k1, a, b, (size_t)n, smp);
        for (int i = 0; i < n; i++) pa[i] = c->sigma_acc * (2.0f * smp[i] - 1.0f);
        for (int r = 0; r < nr; r++) for (int t = 0; t < np; t++) { a[r * np + t] = c->beta_a * fabsf(steer[t]); b[r * np + t] = c->beta_b * fabsf(steer[t]); }
        if (nz->b2) memcpy(smp, nz->b2, sizeof(float) * n); else rng_beta(nz->k2, a, b, (size_t)n, smp);
        for (int i = 0; i < n; i++) ps[i] = c->ksig_steer * (2.0f * smp[i] - 1.0f);
        free(a); free(b); free(smp);
    }
    for (int r = 0; r < nr; r++)             /* cem_helper.py:438-443 */
        for (int t = 0; t < np; t++) {
            an[r * np + t] = (acc[t] + pa[r * np + t]) + c->acc_const * nz->z3[r * np + t];
            sn[r * np + t] = (steer[t] + ps[r * np + t]) + c->steer_const * nz->z3[r * np + t];
        }
    free(pa); free(ps);
}

/* ------------------------------------------------------------------------------------------ */
/* reduced-set inner CEM (compute_beta.py:93-157) on one sample's mother rollouts               */

typedef struct { float beta[64]; float sigma; int idx[64]; float res[64]; } oinner_out_t;

static void top_abs(const float *th, int nm, int nr, int *idx) {
    /* argsort(|theta|)[nm-nr:nm]  (compute_beta.py:117-118), stable */
    float a[4096]; int perm[4096];
    for (int m = 0; m < nm; m++) a[m] = fabsf(th[m]);
    argsort_stable(a, nm, perm);
    for (int i = 0; i < nr; i++) idx[i] = perm[nm - nr + i];
}
static float l1_dist(const float *Fa, const float *Fb) {  /* kernel_computation.py:31-33 */
    float d = 0.0f;
    for (int f = 0; f < 2 * NV; f++) d = d + fabsf(Fa[f] - Fb[f]);
    return d;
}
/* compute_beta_reduced (compute_beta.py:70-91) [D3] */
static float beta_qp(const ocfg_t *c, const float *Kred, const float *rowsum, float *beta) {
    int nr = c->nr, ld = nr;
    float A[64 * 64], rd[64], kbar[64], q[64], u[64], w[64];
    for (int i = 0; i < nr; i++) {
        for (int j = 0; j <= i; j++) A[i * ld + j] = Kred[i * nr + j];
        A[i * ld + i] = Kred[i * nr + i] + 0.05f;
        kbar[i] = c->inv_nm * rowsum[i];
        q[i] = c->m2_inv_nm * rowsum[i];
    }
    chol_inplace(A, nr, ld, rd);
    for (int i = 0; i < nr; i++) {          /* forward: L y = b for b = kbar and b = 1 */
        float a = kbar[i], b = 1.0f;
        for (int k = 0; k < i; k++) { a = fmaf(-A[i * ld + k], u[k], a); b = fmaf(-A[i * ld + k], w[k], b); }
        u[i] = a * rd[i]; w[i] = b * rd[i];
    }
    for (int i = nr - 1; i >= 0; i--) {     /* backward: L^T x = y */
        float a = u[i], b = w[i];
        for (int k = i + 1; k < nr; k++) { a = fmaf(-A[k * ld + i], u[k], a); b = fmaf(-A[k * ld + i], w[k], b); }
        u[i] = a * rd[i]; w[i] = b * rd[i];
    }
    float su = 0.0f, sw = 0.0f;
    for (int i = 0; i < nr; i++) { su = su + u[i]; sw = sw + w[i]; }
    float nu = (su - 1.0f) / sw;
    for (int i = 0; i < nr; i++) beta[i] = fmaf(-nu, w[i], u[i]);
    float s1 = 0.0f, s2 = 0.0f;
    for (int i = 0; i < nr; i++) {
        float t = 0.0f;
        for (int j = 0; j < nr; j++) t = fmaf(Kred[i * nr + j], beta[j], t);
        s1 = fmaf(beta[i], t, s1);
        s2 = fmaf(q[i], beta[i], s2);
    }
    return s1 + s2;
}

void oracle_inner_cem(const ocfg_t *c, const float *F /* (nm,22) */, oinner_out_t *o) {
    int nm = c->nm, nr = c->nr, S = c->S_in, d = nm + 1, ne = c->n_el_in;
    float *th = (float *)malloc(sizeof(float) * S * d), *thn = (float *)malloc(sizeof(float) * S * d);
    float *D = (float *)malloc(sizeof(float) * nm * nm);
    float *cost = (float *)malloc(sizeof(float) * S), *betas = (float *)malloc(sizeof(float) * S * nr);
    int *idxs = (int *)malloc(sizeof(int) * S * nr), *perm = (int *)malloc(sizeof(int) * S);
    float *C = (float *)malloc(sizeof(float) * d * d), *rd = (float *)malloc(sizeof(float) * d);
    float *mean = (float *)malloc(sizeof(float) * d), *xc = (float *)malloc(sizeof(float) * ne * d);
    float *Kmix = (float *)malloc(sizeof(float) * nr * nm), *Kred = (float *)malloc(sizeof(float) * nr * nr);
    memcpy(th, c->theta0, sizeof(float) * S * d);
    if (!c->naive)
        for (int a = 0; a < nm; a++) for (int b = 0; b < nm; b++) D[a * nm + b] = l1_dist(F + a * 2 * NV, F + b * 2 * NV);
    for (int it = 0; it < c->iters_in; it++) {
        for (int s = 0; s < S; s++) {
            const float *row = th + s * d;
            int *idx = idxs + s * nr;
            float sigma = row[nm], rowsum[64];
            top_abs(row, nm, nr, idx);
            om_lapscale_t ls = om_lap_scale(sigma);                              /* [D1] */
            for (int i = 0; i < nr; i++) {
                float rs = 0.0f;
                for (int m = 0; m < nm; m++) {
                    float dist = c->naive ? l1_dist(F + idx[i] * 2 * NV, F + m * 2 * NV) : D[idx[i] * nm + m];
                    float k = om_lap(dist, ls);
                    Kmix[i * nm + m] = k;
                    rs = rs + k;
                }
                rowsum[i] = rs;
            }
            for (int i = 0; i < nr; i++)
                for (int j = 0; j < nr; j++) {
                    if (c->naive) { float dist = l1_dist(F + idx[i] * 2 * NV, F + idx[j] * 2 * NV); Kred[i * nr + j] = om_lap(dist, ls); }
                    else Kred[i * nr + j] = Kmix[i * nm + idx[j]];
                }
            cost[s] = beta_qp(c, Kred, rowsum, betas + s * nr);
        }
        /* compute_mean_cov_beta (compute_beta.py:51-68) */
        argsort_stable(cost, S, perm);
        int imin = perm[0];                                      /* argmin (no NaN) */
        for (int e = 0; e < ne; e++) memcpy(thn + e * d, th + perm[e] * d, sizeof(float) * d);
        for (int i = 0; i < d; i++) {
            float s = 0.0f;
            for (int e = 0; e < ne; e++) s = s + thn[e * d + i];
            mean[i] = s / (float)ne;
        }
        for (int e = 0; e < ne; e++) for (int i = 0; i < d; i++) xc[e * d + i] = thn[e * d + i] - mean[i];
        for (int i = 0; i < d; i++)
            for (int j = 0; j <= i; j++) {
                float a = 0.0f;
                for (int e = 0; e < ne; e++) a = fmaf(xc[e * d + i], xc[e * d + j], a);
                a = a / (float)(ne - 1);
                if (i == j) a = a + 0.05f;
                C[i * d + j] = a;
            }
        chol_inplace(C, d, d, rd);
        const float *z = c->zb_iter + (size_t)it * (S - ne) * d;
        for (int r = 0; r < S - ne; r++)
            for (int i = 0; i < d; i++) thn[(ne + r) * d + i] = mvn_elem(C, d, i, z + r * d, mean[i]);
        for (int s = 0; s < S; s++) { float v = thn[s * d + nm]; thn[s * d + nm] = (v != v) ? v : (v > c->sigma_clip ? v : c->sigma_clip); }
        /* compute_beta.py:136-142; sigma is read from the RESAMPLED array [Q7] */
        o->res[it] = cost[imin];
        if (it == c->iters_in - 1) {
            for (int i = 0; i < nr; i++) { o->beta[i] = betas[imin * nr + i]; o->idx[i] = idxs[imin * nr + i]; }
            o->sigma = thn[imin * d + nm];
        }
        float *tmp = th; th = thn; thn = tmp;
    }
    free(th); free(thn); free(D); free(cost); free(betas); free(idxs); free(perm); free(C); free(rd); free(mean); free(xc); free(Kmix); free(Kred);
}

/* ------------------------------------------------------------------------------------------ */
/* Stage B: rollouts + risk for one CEM sample                                                 */

typedef struct {
    float risk, lane;          /* mmd_obs / cvar_obs / saa_obs and the matching lane value */
    float beta[64], sigma, res_beta[64];
    float red_cost[64];        /* c_i of the (reduced) rollout set, for stage tests */
    int   red_idx[64];
} orisk_out_t;

void oracle_risk(const ocfg_t *c, int cost_kind, const float *acc, const float *steer, const float *st0,
                 const onoise_t *nz, const float *x_obs, const float *y_obs, orisk_out_t *o,
                 float *x_roll_out, float *y_roll_out /* optional (R,np) */) {
    int nr = c->nr, np = c->np, nm = c->nm;
    int R = (cost_kind == COST_MMD_OPT) ? nm : nr;
    float *an = (float *)malloc(sizeof(float) * nr * np), *sn = (float *)malloc(sizeof(float) * nr * np);
    float *xr = (float *)malloc(sizeof(float) * R * np), *yr = (float *)malloc(sizeof(float) * R * np);
    noisy_controls(c, acc, steer, nz, an, sn);
    for (int m = 0; m < R; m++) {
        /* mmd_opt: acc rows repeated, steer rows tiled (cem_helper.py:510-511): m = i*nr + j */
        int ia = (cost_kind == COST_MMD_OPT) ? m / nr : m, is = (cost_kind == COST_MMD_OPT) ? m % nr : m;
        rollout_one(c, an + ia * np, sn + is * np, st0, xr + m * np, yr + m * np);
    }
    if (x_roll_out) memcpy(x_roll_out, xr, sizeof(float) * R * np);
    if (y_roll_out) memcpy(y_roll_out, yr, sizeof(float) * R * np);
    float cst[64], lb[64], ub[64];
    memset(o, 0, sizeof *o);
    if (cost_kind == COST_MMD_OPT) {
        /* compute_coeff (cem_helper.py:553-564) [D2]: features = [W x_roll ; W y_roll] */
        float *F = (float *)malloc(sizeof(float) * nm * 2 * NV);
        for (int m = 0; m < nm; m++)
            for (int k = 0; k < NV; k++) {
                float ax = 0.0f, ay = 0.0f;
                for (int t = 0; t < np; t++) { ax = fmaf(c->Wfit[k * np + t], xr[m * np + t], ax); ay = fmaf(c->Wfit[k * np + t], yr[m * np + t], ay); }
                F[m * 2 * NV + k] = ax; F[m * 2 * NV + NV + k] = ay;
            }
        oinner_out_t in;
        oracle_inner_cem(c, F, &in);
        free(F);
        for (int i = 0; i < nr; i++) {
            int m = in.idx[i];
            o->red_idx[i] = m; o->beta[i] = in.beta[i];
            cst[i] = fbar_max(c, xr + m * np, yr + m * np, x_obs, y_obs);
            lane_max(c, yr + m * np, &lb[i], &ub[i]);
        }
        o->sigma = in.sigma;
        for (int i = 0; i < c->iters_in; i++) o->res_beta[i] = in.res[i];
        o->risk = mmd_cost(c, o->beta, cst, o->sigma);                 /* costs.py:173-186 */
        o->lane = mmd_cost(c, o->beta, lb, o->sigma) + mmd_cost(c, o->beta, ub, o->sigma);  /* costs.py:121-135 */
    } else {
        for (int i = 0; i < nr; i++) {
            o->red_idx[i] = i;
            cst[i] = fbar_max(c, xr + i * np, yr + i * np, x_obs, y_obs);
            lane_max(c, yr + i * np, &lb[i], &ub[i]);
        }
        if (cost_kind == COST_MMD_RANDOM) {      /* cem.py:355-356, 404-424: beta = 1/nr, sigma = 0.01, lane = 0 */
            for (int i = 0; i < nr; i++) o->beta[i] = c->beta_del;
            o->sigma = c->sigma_random;
            o->risk = mmd_cost(c, o->beta, cst, o->sigma);
            o->lane = 0.0f;
        } else if (cost_kind == COST_CVAR) {     /* costs.py:206-221, 137-158 */
            o->risk = cvar_cost(c, cst);
            o->lane = cvar_cost(c, lb) + cvar_cost(c, ub);
        } else {                                 /* costs.py:223-234, 160-171 */
            o->risk = saa_cost(c, cst);
            float sl = 0.0f, su = 0.0f;
            for (int i = 0; i < nr; i++) { sl = sl + (lb[i] > 0.0f ? 1.0f : 0.0f); su = su + (ub[i] > 0.0f ? 1.0f : 0.0f); }
            o->lane = (sl + su) / (float)nr;
        }
    }
    for (int i = 0; i < nr; i++) o->red_cost[i] = cst[i];
    free(an); free(sn); free(xr); free(yr);
}

/* ------------------------------------------------------------------------------------------ */
/* Stage C: elite selection + mean/covariance update + resampling, one episode                  */
/* cem.py:233-315, cem_helper.py:264-314                                                        */

typedef struct { int sel; float cost_min; int top[32]; int elite[8]; } oselect_out_t;

void oracle_sample_params(const ocfg_t *c, const float *mean, const float *cov, const float *z, int rows, float *out) {
    /* multivariate_normal(cholesky) + clip of the 4 speed columns (cem_helper.py:126-148 / 292-307) */
    float L[NP_ * NP_], rd[NP_];
    memcpy(L, cov, sizeof L);
    chol_inplace(L, NP_, NP_, rd);
    for (int r = 0; r < rows; r++)
        for (int i = 0; i < NP_; i++) {
            float v = mvn_elem(L, NP_, i, z + r * NP_, mean[i]);
            out[r * NP_ + i] = (i < 4) ? clipf(v, c->v_min, c->v_max) : v;
        }
}

void oracle_select(const ocfg_t *c, const float *res_norm, const float *risk, const float *cost_base,
                   const float *params /* (B,8) in/out */, float *params_next, float *mean, float *cov,
                   const float *z_cem /* (B-n_el,8) */, oselect_out_t *o) {
    int B = c->B, n20 = c->n_el_cost, n5 = c->n_el;
    int *p1 = (int *)malloc(sizeof(int) * B), *p2 = (int *)malloc(sizeof(int) * B);
    float *r1 = (float *)malloc(sizeof(float) * B);
    argsort_stable(res_norm, B, p1);                       /* cem.py:233  (keeps all B rows [Q4]) */
    for (int p = 0; p < B; p++) r1[p] = risk[p1[p]];
    argsort_stable(r1, B, p2);                             /* cem.py:264 */
    float cost20[32]; int s20[32], p3[32];
    for (int q = 0; q < n20; q++) {
        s20[q] = p1[p2[q]];
        o->top[q] = s20[q];
        cost20[q] = cost_base[s20[q]] + c->w_obs * risk[s20[q]];      /* cem_helper.py:253-261 [D4] */
    }
    argsort_stable(cost20, n20, p3);                       /* cem_helper.py:267 */
    float ce[8], w[8], th[8][NP_];
    for (int e = 0; e < n5; e++) {
        ce[e] = cost20[p3[e]];
        o->elite[e] = s20[p3[e]];
        memcpy(th[e], params + s20[p3[e]] * NP_, sizeof(float) * NP_);
    }
    /* compute_shifted_samples (cem_helper.py:280-314) */
    float wmin = ce[0]; int imin = 0;
    for (int e = 1; e < n5; e++) if (ce[e] < wmin) { wmin = ce[e]; imin = e; }
    float sum_w = 0.0f;
    for (int e = 0; e < n5; e++) { w[e] = om_exp((-c->lam_inv) * (ce[e] - wmin)); sum_w = sum_w + w[e]; }
    float mean_new[NP_], cov_new[NP_ * NP_], dif[8][NP_];
    for (int i = 0; i < NP_; i++) {
        float s = 0.0f;
        for (int e = 0; e < n5; e++) s = s + th[e][i] * w[e];
        mean_new[i] = c->one_m_alpha_mean * mean[i] + c->alpha_mean * (s / sum_w);
    }
    for (int e = 0; e < n5; e++) for (int i = 0; i < NP_; i++) dif[e][i] = th[e][i] - mean_new[i];
    for (int i = 0; i < NP_; i++)
        for (int j = 0; j < NP_; j++) {
            float s = 0.0f;
            for (int e = 0; e < n5; e++) s = s + w[e] * (dif[e][i] * dif[e][j]);
            float v = c->one_m_alpha_cov * cov[i * NP_ + j] + c->alpha_cov * (s / sum_w);
            cov_new[i * NP_ + j] = (i == j) ? v + 0.01f : v;
        }
    memcpy(mean, mean_new, sizeof mean_new);
    memcpy(cov, cov_new, sizeof cov_new);
    for (int e = 0; e < n5; e++) memcpy(params_next + e * NP_, th[e], sizeof(float) * NP_);
    oracle_sample_params(c, mean, cov, z_cem, B - n5, params_next + n5 * NP_);
    /* cem.py:308-315 [Q1]: idx_min indexes the 5 sorted elite costs but is applied to the 20 risk-sorted rows */
    o->sel = s20[imin];
    o->cost_min = wmin;
    free(p1); free(p2); free(r1);
}

/* ------------------------------------------------------------------------------------------ */
/* noise tables for one (idx_mpc, iteration)  (cem.py:225,254,302; cem_helper.py:405-443)        */

void oracle_noise_tables(const ocfg_t *c, int32_t idx_mpc, int32_t iter, float *z1, float *z2, float *z3, float *z_cem, uint32_t *keys /* [k1.k0,k1.k1,k2.k0,k2.k1] */) {
    int n = c->nr * c->np;
    okey_t key = rng_key((uint32_t)(3 * idx_mpc + 5 * iter + 7));
    okey_t k1 = rng_split0(key), k2 = rng_split0(k1), k3 = rng_split0(k2);
    rng_normal(k1, (size_t)n, z1);
    rng_normal(k2, (size_t)n, z2);
    rng_normal(k3, (size_t)n, z3);
    rng_normal(k2, (size_t)(c->B - c->n_el) * NP_, z_cem);      /* [Q6] same key as the steer noise */
    if (keys) { keys[0] = k1.k0; keys[1] = k1.k1; keys[2] = k2.k0; keys[3] = k2.k1; }
}

/* ------------------------------------------------------------------------------------------ */
/* the solve (cem.py:201-333 / 335-462 / 464-588 / 590-714)                                      */

typedef struct {
    float cx[NV], cy[NV], cost_lane, cost_obs, beta[64], sigma, res_beta[64];
    int32_t sel_last;
} osolve_out_t;

/* optional per-iteration trace for stage-level (teacher-forced) tests */
typedef struct {
    float *params;     /* (iters+1, B, 8) batch entering each iteration (+ final) */
    float *res_norm;   /* (iters, B) */
    float *risk;       /* (iters, B) */
    float *lane;       /* (iters, B) */
    float *cost_base;  /* (iters, B) */
    float *mean;       /* (iters+1, 8) */
    float *cov;        /* (iters+1, 64) */
    int32_t *sel;      /* (iters,) emitted sample index */
    float *cxy;        /* (iters, 22) emitted cx,cy */
} otrace_t;

static int g_oracle_threads = 1;
void oracle_set_threads(int n) { g_oracle_threads = n; }
typedef struct {
    const ocfg_t *c; int cost_kind; const float *params, *beq_x, *beq_y; float v_des; float *lam_x, *lam_y, *s_lane;
    oproj_out_t *pr; orisk_out_t *rk; const float *st0; const onoise_t *nz; const float *x_obs, *y_obs; int b0, b1;
} osample_job_t;
static void *osample_worker(void *arg) {
    osample_job_t *j = (osample_job_t *)arg;
    for (int b = j->b0; b < j->b1; b++) {
        oracle_project(j->c, j->params + b * NP_, j->beq_x, j->beq_y, j->v_des, j->lam_x + b * NV, j->lam_y + b * NV, j->s_lane + b * 2 * NL, &j->pr[b]);
        oracle_risk(j->c, j->cost_kind, j->pr[b].acc, j->pr[b].steer, j->st0, j->nz, j->x_obs, j->y_obs, &j->rk[b], NULL, NULL);
    }
    return NULL;
}

int oracle_solve(const ocfg_t *c, int cost_kind, int32_t idx_mpc, const float *init_state, const float *mean0,
                 const float *cov0, const float *x_obs, const float *y_obs, float v_des, osolve_out_t *out, otrace_t *tr) {
    int B = c->B, np = c->np, nr = c->nr;
    float beq_x[3] = {init_state[0], init_state[2], init_state[4]};            /* cem_helper.py:152-167 */
    float beq_y[4] = {init_state[1], init_state[3], init_state[5], 0.0f};
    float st0[5] = {init_state[0], init_state[1], init_state[2], init_state[3], om_atan2(init_state[3], init_state[2])}; /* cem.py:218-219 */
    float mean[NP_], cov[NP_ * NP_];
    memcpy(mean, mean0, sizeof mean); memcpy(cov, cov0, sizeof cov);
    float *params = (float *)malloc(sizeof(float) * B * NP_), *params_next = (float *)malloc(sizeof(float) * B * NP_);
    float *lam_x = (float *)calloc((size_t)B * NV, sizeof(float)), *lam_y = (float *)calloc((size_t)B * NV, sizeof(float));
    float *s_lane = (float *)calloc((size_t)B * 2 * NL, sizeof(float));                 /* cem.py:206-208 */
    oproj_out_t *pr = (oproj_out_t *)malloc(sizeof(oproj_out_t) * B);
    orisk_out_t *rk = (orisk_out_t *)malloc(sizeof(orisk_out_t) * B);
    float *res_norm = (float *)malloc(sizeof(float) * B), *risk = (float *)malloc(sizeof(float) * B), *base = (float *)malloc(sizeof(float) * B);
    int n = nr * np;
    float *z1 = (float *)malloc(sizeof(float) * n), *z2 = (float *)malloc(sizeof(float) * n), *z3 = (float *)malloc(sizeof(float) * n);
    float *z_cem = (float *)malloc(sizeof(float) * (B - c->n_el) * NP_);
    oracle_sample_params(c, mean, cov, c->z_init, B, params);                          /* cem.py:213 */
    for (int it = 0; it < c->iters; it++) {
        uint32_t keys[4];
        oracle_noise_tables(c, idx_mpc, it, z1, z2, z3, z_cem, keys);
        onoise_t nz = {z1, z2, z3, {keys[0], keys[1]}, {keys[2], keys[3]}, NULL, NULL};
        if (tr && tr->params) memcpy(tr->params + (size_t)it * B * NP_, params, sizeof(float) * B * NP_);
        if (tr && tr->mean) memcpy(tr->mean + it * NP_, mean, sizeof mean);
        if (tr && tr->cov) memcpy(tr->cov + it * NP_ * NP_, cov, sizeof cov);
        {   /* the B samples of one iteration are independent: split them over host threads (results do not depend on the split) */
            int nt = g_oracle_threads < 1 ? 1 : (g_oracle_threads > 256 ? 256 : g_oracle_threads);
            if (nt > B) nt = B;
            pthread_t th[256]; osample_job_t jobs[256];
            for (int k = 0; k < nt; k++) {
                osample_job_t j = {c, cost_kind, params, beq_x, beq_y, v_des, lam_x, lam_y, s_lane, pr, rk, st0, &nz, x_obs, y_obs,
                                   (int)((long)B * k / nt), (int)((long)B * (k + 1) / nt)};
                jobs[k] = j;
                if (nt > 1) pthread_create(&th[k], NULL, osample_worker, &jobs[k]); else osample_worker(&jobs[k]);
            }
            if (nt > 1) for (int k = 0; k < nt; k++) pthread_join(th[k], NULL);
            for (int b = 0; b < B; b++) { res_norm[b] = pr[b].res_norm; risk[b] = rk[b].risk; base[b] = pr[b].cost_base; }
        }
        oselect_out_t so;
        oracle_select(c, res_norm, risk, base, params, params_next, mean, cov, z_cem, &so);
        if (tr) {
            if (tr->res_norm) memcpy(tr->res_norm + (size_t)it * B, res_norm, sizeof(float) * B);
            if (tr->risk) memcpy(tr->risk + (size_t)it * B, risk, sizeof(float) * B);
            if (tr->cost_base) memcpy(tr->cost_base + (size_t)it * B, base, sizeof(float) * B);
            if (tr->lane) for (int b = 0; b < B; b++) tr->lane[(size_t)it * B + b] = rk[b].lane;
            if (tr->sel) tr->sel[it] = so.sel;
            if (tr->cxy) { memcpy(tr->cxy + it * 2 * NV, pr[so.sel].cx, sizeof(float) * NV); memcpy(tr->cxy + it * 2 * NV + NV, pr[so.sel].cy, sizeof(float) * NV); }
        }
        if (it == c->iters - 1) {                                                       /* [Q2] */
            int s = so.sel;
            memset(out, 0, sizeof *out);
            memcpy(out->cx, pr[s].cx, sizeof out->cx); memcpy(out->cy, pr[s].cy, sizeof out->cy);
            out->cost_lane = rk[s].lane; out->cost_obs = rk[s].risk;
            for (int i = 0; i < nr; i++) out->beta[i] = rk[s].beta[i];
            out->sigma = rk[s].sigma;
            for (int i = 0; i < c->iters_in; i++) out->res_beta[i] = rk[s].res_beta[i];
            out->sel_last = s;
        }
        float *tmp = params; params = params_next; params_next = tmp;
    }
    if (tr && tr->params) memcpy(tr->params + (size_t)c->iters * B * NP_, params, sizeof(float) * B * NP_);
    if (tr && tr->mean) memcpy(tr->mean + c->iters * NP_, mean, sizeof mean);
    if (tr && tr->cov) memcpy(tr->cov + c->iters * NP_ * NP_, cov, sizeof cov);
    free(params); free(params_next); free(lam_x); free(lam_y); free(s_lane); free(pr); free(rk);
    free(res_norm); free(risk); free(base); free(z1); free(z2); free(z3); free(z_cem);
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* vector entry points for the math / RNG known-answer tests                                    */

void oracle_math_vec(int fn, const float *x, const float *y, float *out, int n) {
    for (int i = 0; i < n; i++) {
        switch (fn) {
            case 0: out[i] = om_exp(x[i]); break;
            case 1: out[i] = om_log(x[i]); break;
            case 2: out[i] = om_log1p(x[i]); break;
            case 3: out[i] = om_sin(x[i]); break;
            case 4: out[i] = om_cos(x[i]); break;
            case 5: out[i] = om_tan(x[i]); break;
            case 6: out[i] = om_atan(x[i]); break;
            case 7: out[i] = om_atan2(y[i], x[i]); break;
            case 8: out[i] = xla_erfinv32(x[i]); break;
            case 12: out[i] = om_lap(x[i], om_lap_scale(y[i])); break;          /* Laplace kernel entry k(d = x; sigma = y) */
            default: out[i] = NAN;
        }
    }
}
void oracle_threefry(uint32_t k0, uint32_t k1, uint32_t x0, uint32_t x1, uint32_t *o) { okey_t k = {k0, k1}; threefry2x32(k, x0, x1, &o[0], &o[1]); }
void oracle_rng_split(uint32_t k0, uint32_t k1, int m, uint32_t *o) {
    okey_t k = {k0, k1};
    okey_t *ks = (okey_t *)malloc(sizeof(okey_t) * m);
    rng_split(k, m, ks);
    for (int i = 0; i < m; i++) { o[2 * i] = ks[i].k0; o[2 * i + 1] = ks[i].k1; }
    free(ks);
}
void oracle_rng_bits(uint32_t k0, uint32_t k1, int n, uint32_t *o) { okey_t k = {k0, k1}; rng_bits(k, (size_t)n, o); }
void oracle_rng_normal(uint32_t k0, uint32_t k1, int n, float *o) { okey_t k = {k0, k1}; rng_normal(k, (size_t)n, o); }
void oracle_rng_uniform(uint32_t k0, uint32_t k1, int n, float lo, float hi, float *o) { okey_t k = {k0, k1}; rng_uniform(k, (size_t)n, lo, hi, o); }
void oracle_rng_beta(uint32_t k0, uint32_t k1, const float *a, const float *b, int n, float *o) { okey_t k = {k0, k1}; rng_beta(k, a, b, (size_t)n, o); }
int oracle_sizeof_cfg(void) { return (int)sizeof(ocfg_t); }
int oracle_sizeof_proj(void) { return (int)sizeof(oproj_out_t); }
int oracle_sizeof_risk(void) { return (int)sizeof(orisk_out_t); }
int oracle_sizeof_solve(void) { return (int)sizeof(osolve_out_t); }
