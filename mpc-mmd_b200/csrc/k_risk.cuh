// k_risk.cuh -- noisy control rollouts, reduced-set selection and the risk functionals.
//
// Replaces, per CEM sample:
//   Helper.compute_rollout_complete_baseline / _opt + compute_rollout_one_step   S/optimizer/cem_helper.py:402-538, 380-400
//   Helper.compute_coeff                                                         S/optimizer/cem_helper.py:553-564
//   beta_cem.compute_cem (+ compute_mean_cov_beta, compute_beta_reduced)          S/compute_beta.py:93-157, 51-91
//   kernel_matrix.compute_kernel / compute_mmd (Laplace kernel)                   S/kernel_computation.py:19-87
//   Costs.compute_f_bar / compute_lane_bar / compute_{mmd,cvar,saa}_{obs,lane}     S/optimizer/costs.py:50-71, 121-234
// k_rollouts : noisy controls + rollouts for a group of samples; finishes cvar / saa / mmd_random.
// k_inner_cem: one CTA per sample (mmd_opt: the inner reduced-set CEM, the dominant cost of a solve,
//              plus the MMD risk of the chosen set), all state in shared memory.
#pragma once
#include "common.cuh"

struct RiskArgs {
    int n_samples, B, cost_kind;
    const float *acc, *steer;        // [n][100]
    const float* state0;             // [E][5]
    const float *z1, *z2, *z3;       // episode e at z + e*z_stride, (nr,np)
    size_t z_stride;
    const uint32_t* keys;            // episode e at keys + e*key_stride, 4 words
    size_t key_stride;
    const float* btab;               // episode e at btab + e*btab_stride: [4][GT_FIELDS][nr*np] (beta noise; null -> direct sampler)
    size_t btab_stride;
    const float *x_obs, *y_obs;      // [E][O][100]
    float *risk, *lane;              // [n]
    float *beta, *sigma, *res_beta;  // [n][nr], [n], [n][iters_in]
};

// Euler bicycle rollout, records the state BEFORE each step  [cem_helper.py:380-400, 451-458]
__device__ __forceinline__ void rollout_one(const DCfg& c, const float* a, const float* s, const float* st0, float* xr, float* yr) {
    float x = st0[0], y = st0[1], vx = st0[2], vy = st0[3], psi = st0[4];
    for (int t = 0; t < c.np; t++) {
        const float at = a[t], st = s[t];       // read before the stores: xr / yr may alias a / s (num_reduced-rollout costs)
        xr[t] = x; yr[t] = y;
        float v = sqrtf(vx * vx + vy * vy);
        v = v + at * c.dt;
        float psidot = (v * dm::tan_(st)) / c.wheel_base;
        psi = psi + psidot * c.dt;
        float sp, cp; dm::sincos_(psi, sp, cp);
        vx = v * cp; vy = v * sp;
        x = x + vx * c.dt; y = y + vy * c.dt;
    }
}
// obstacle indicator at one (rollout point, obstacle point)  [costs.py:50-60]
__device__ __forceinline__ float fbar(const DCfg& c, float x, float y, float xo, float yo) {
    float wc = x - xo, ws = y - yo;
    float cost = (-(wc * wc) / c.a2_obs - (ws * ws) / c.b2_obs) + 1.0f;
    return dm::max0_(cost);
}
// Laplace-kernel MMD of nr scalar costs against the zero cost  [kernel_computation.py:67-87]
__device__ __noinline__ float mmd_cost(const DCfg& c, const float* beta, const float* cost, float sigma) {
    const int nr = c.nr;
    float s1 = 0.0f, s2 = 0.0f;
    for (int i = 0; i < nr; i++) {
        float t = 0.0f;
        for (int j = 0; j < nr; j++) t = fmaf(dm::exp_(-fabsf(cost[i] - cost[j]) / sigma), beta[j], t);
        s1 = fmaf(beta[i], t, s1);
        float e = dm::exp_(-fabsf(cost[i] - 0.0f) / sigma), u = 0.0f;
        for (int j = 0; j < nr; j++) u = fmaf(e, c.beta_del, u);
        s2 = fmaf(beta[i], u, s2);
    }
    return c.ker_wt * (s1 - 2.0f * s2);
}
// the same value computed by a whole warp (num_reduced <= 64): lane l evaluates the kernel rows i = l, l + 32 as the same ascending-j
// fma chains, then every lane folds them in ascending i -- bit-identical to mmd_cost, nr^2 / 32 exps per lane instead of nr^2 on one
__device__ __forceinline__ float mmd_cost_warp(const DCfg& c, const float* beta, const float* cost, float sigma, int lane) {
    const int nr = c.nr;
    float tv[2] = {0.0f, 0.0f}, uv[2] = {0.0f, 0.0f};
#pragma unroll
    for (int h = 0; h < 2; h++) {
        const int i = lane + 32 * h;
        if (i < nr) {
            float t = 0.0f;
            for (int j = 0; j < nr; j++) t = fmaf(dm::exp_(-fabsf(cost[i] - cost[j]) / sigma), beta[j], t);
            const float e = dm::exp_(-fabsf(cost[i] - 0.0f) / sigma);
            float u = 0.0f;
            for (int j = 0; j < nr; j++) u = fmaf(e, c.beta_del, u);
            tv[h] = t; uv[h] = u;
        }
    }
    float s1 = 0.0f, s2 = 0.0f;
    for (int i = 0; i < nr; i++) {
        const float t = __shfl_sync(FULL, i < 32 ? tv[0] : tv[1], i & 31), u = __shfl_sync(FULL, i < 32 ? uv[0] : uv[1], i & 31);
        s1 = fmaf(beta[i], t, s1);
        s2 = fmaf(beta[i], u, s2);
    }
    return c.ker_wt * (s1 - 2.0f * s2);
}
// jnp.quantile (linear interpolation) + mean of the tail  [costs.py:213-220]
__device__ __noinline__ float cvar_cost(const DCfg& c, const float* v) {
    const int nr = c.nr;
    int perm[MPCMMD_MAX_NR_DEV];
    for (int i = 0; i < nr; i++) {            // stable insertion sort, NaN last
        int j = i;
        while (j > 0 && dm::lt_nanlast(v[i], v[perm[j - 1]])) { perm[j] = perm[j - 1]; j--; }
        perm[j] = i;
    }
    float q = c.alpha_quant * (float)(nr - 1);
    float lo = floorf(q), hi = ceilf(q);
    float hw = q - lo, lw = 1.0f - hw;
    int ilo = (int)lo, ihi = (int)hi;
    ilo = ilo < 0 ? 0 : (ilo > nr - 1 ? nr - 1 : ilo);
    ihi = ihi < 0 ? 0 : (ihi > nr - 1 ? nr - 1 : ihi);
    float var = v[perm[ilo]] * lw + v[perm[ihi]] * hw;
    float s = 0.0f; int n = 0;
    for (int i = 0; i < nr; i++) if (v[i] >= var) { s = s + v[i]; n++; }
    return n > 0 ? s / (float)n : 0.0f;
}

// ---------------------------------------------------------------------------------------------
// k_rollouts: noisy controls + Euler rollouts for a group of samples per CTA.  cost kinds with num_reduced rollouts
// (cvar / saa / mmd_random) finish here with their risk functional; mmd_opt writes the num_reduced^2 mother rollouts and
// their ridge-fit features for k_inner_cem.  Splitting the solve this way keeps each kernel's instruction footprint
// inside the 32 KB L1.5 instruction cache (v1 of this file, one fused kernel of 13.7k SASS instructions, spent
// 6.7 issue slots stalled on instruction fetch per instruction issued -- profiles/r01_v2_summary.md).
#define ROLL_THREADS 128
struct RollArgs {
    RiskArgs r;
    int spb;                 // samples per CTA
    int R;                   // rollouts per sample (nr or nr*nr)
    float *xroll, *yroll;    // [n][R][np]   (mmd_opt only)
    float* feat;             // [n][nm][22]  (mmd_opt only)
    float* stash;            // [persistent CTAs][S][32]  row stash of k_inner_cem_warp
    int* ridx;               // [n][nr]  reduced set chosen by k_inner_cem_fast, read by k_opt_risk
    float* bscratch;         // [n][S][nr + 1]  per-row beta vectors and packed reduced-set indices of k_inner_cem_fast's current iteration
};
__host__ __device__ inline int roll_tail_floats(int nr) { return nr <= 16 ? 16 : ((nr + 3) & ~3); }       // per-rollout cost / lane-lb / lane-ub slots
// per sample: noisy controls (2 x nr x np), rollouts (2 x R x np), cost slots.  With R == nr (cvar / saa / mmd_random) rollout r reads only
// control row r, so the rollouts overwrite the controls in place and the sample needs half the space (twice the resident CTAs).
__host__ __device__ inline int roll_smem_floats(int spb, int nr, int np, int R) { return spb * (2 * nr * np + (R == nr ? 0 : 2 * R * np) + 4 * roll_tail_floats(nr)); }

__global__ void __launch_bounds__(ROLL_THREADS) k_rollouts(DCfg c, RollArgs ra) {
    extern __shared__ __align__(16) float sm[];
    const RiskArgs& a = ra.r;
    const int tid = threadIdx.x, nt = ROLL_THREADS, warp = tid >> 5, lane = tid & 31;
    const int nr = c.nr, np = c.np, n = nr * np, R = ra.R, spb = ra.spb;
    const int g0 = blockIdx.x * spb;
    const int ns = min(spb, a.n_samples - g0);                 // samples in this CTA
    if (ns <= 0) return;
    const bool opt = a.cost_kind == 0;
    const int tail = roll_tail_floats(nr), xoff = opt ? 2 * n : 0, yoff = opt ? 2 * n + R * np : n, coff = opt ? 2 * n + 2 * R * np : 2 * n;
    const int per = coff + 4 * tail;
    // ---- perturbed controls  [cem_helper.py:405-443 / 470-508]
#pragma unroll 1
    for (int i = tid; i < ns * n; i += nt) {
        const int ls = i / n, el = i % n, t = el % np, g = g0 + ls, e = g / a.B;
        float* an = sm + ls * per; float* sn = an + n;
        const float av = a.acc[(size_t)g * T_ + t], sv = a.steer[(size_t)g * T_ + t];
        const float* z3 = a.z3 + e * a.z_stride;
        float pa, ps;
        if (c.noise_kind == 0) {
            pa = (c.sigma_acc * fabsf(av)) * (a.z1 + e * a.z_stride)[el];
            ps = (c.sigma_steer * fabsf(sv)) * (a.z2 + e * a.z_stride)[el];
        } else {
            const uint32_t* keys = a.keys + e * a.key_stride;
            dr::Key k1, k2; k1.k0 = keys[0]; k1.k1 = keys[1]; k2.k0 = keys[2]; k2.k1 = keys[3];
            float b1, b2;
            if (a.btab) {
                const float* bt = a.btab + e * a.btab_stride;
                b1 = dr::beta_replay(bt, k1, (uint32_t)n, (uint32_t)el, c.beta_a * fabsf(av), c.beta_b * fabsf(av));
                b2 = dr::beta_replay(bt + (size_t)2 * GT_FIELDS * n, k2, (uint32_t)n, (uint32_t)el, c.beta_a * fabsf(sv), c.beta_b * fabsf(sv));
            } else {
                b1 = dr::beta_elem(k1, (uint32_t)n, (uint32_t)el, c.beta_a * fabsf(av), c.beta_b * fabsf(av));
                b2 = dr::beta_elem(k2, (uint32_t)n, (uint32_t)el, c.beta_a * fabsf(sv), c.beta_b * fabsf(sv));
            }
            pa = c.sigma_acc * (2.0f * b1 - 1.0f);
            ps = c.ksig_steer * (2.0f * b2 - 1.0f);
        }
        an[el] = (av + pa) + c.acc_const * z3[el];
        sn[el] = (sv + ps) + c.steer_const * z3[el];
    }
    __syncthreads();
    // ---- rollouts: thread (sample, rollout); mmd_opt mother sample m = i*nr + j uses acc noise i, steer noise j  [cem_helper.py:510-511]
#pragma unroll 1
    for (int i = tid; i < ns * R; i += nt) {
        const int ls = i / R, m = i % R, g = g0 + ls, e = g / a.B;
        float* an = sm + ls * per; float* sn = an + n; float* xr = an + xoff; float* yr = an + yoff;
        const int ia = opt ? m / nr : m, is = opt ? m % nr : m;
        rollout_one(c, an + ia * np, sn + is * np, a.state0 + e * 5, xr + m * np, yr + m * np);
    }
    __syncthreads();
    if (opt) {
        // ---- mother rollouts + ridge-fit features to global  [cem_helper.py:553-564, folded]
        const int nm = R;
#pragma unroll 1
        for (int i = tid; i < ns * nm * np; i += nt) {
            const int ls = i / (nm * np), rem = i % (nm * np);
            const float* xr = sm + ls * per + 2 * n; const float* yr = xr + R * np;
            ra.xroll[(size_t)(g0 + ls) * nm * np + rem] = xr[rem];
            ra.yroll[(size_t)(g0 + ls) * nm * np + rem] = yr[rem];
        }
#pragma unroll 1
        for (int i = tid; i < ns * nm * 2 * NV; i += nt) {
            const int ls = i / (nm * 2 * NV), rem = i % (nm * 2 * NV), m = rem / (2 * NV), k = rem % (2 * NV);
            const float* xr = sm + ls * per + 2 * n; const float* yr = xr + R * np;
            const float* src = (k < NV) ? xr + m * np : yr + m * np;
            const float* W = c.Wfit + (k < NV ? k : k - NV) * np;
            float acc = 0.0f;
            for (int t = 0; t < np; t++) acc = fmaf(__ldg(W + t), src[t], acc);
            ra.feat[(size_t)(g0 + ls) * nm * 2 * NV + rem] = acc;
        }
        return;
    }
    // ---- risk of the nr rollouts (one warp per sample at a time)  [costs.py:50-71, 137-234; cem.py:355-356]
#pragma unroll 1
    for (int ls = warp; ls < ns; ls += nt / 32) {
        const int g = g0 + ls, e = g / a.B;
        const float* xr = sm + ls * per + xoff; const float* yr = sm + ls * per + yoff;
        float* cst = sm + ls * per + coff; float* lb = cst + tail; float* ub = lb + tail; float* bet = ub + tail;
        const float* xo = a.x_obs + (size_t)e * c.O * T_; const float* yo = a.y_obs + (size_t)e * c.O * T_;
#pragma unroll 1
        for (int r = 0; r < nr; r++) {
            float m = 0.0f, l = 0.0f, u = 0.0f;
            for (int t = lane; t < np; t += 32) {          // lane owns timesteps; the maxima are order independent (NaN propagates either way)
                const float x = xr[r * np + t], y = yr[r * np + t];
                _Pragma("unroll 4") for (int o = 0; o < c.O; o++) m = dm::nmax_(m, fbar(c, x, y, xo[o * T_ + t], yo[o * T_ + t]));
                l = dm::nmax_(l, dm::max0_(-y + c.y_lb));
                u = dm::nmax_(u, dm::max0_(y - c.y_ub));
            }
            m = warp_nmax(m); l = warp_nmax(l); u = warp_nmax(u);
            if (lane == 0) { cst[r] = m; lb[r] = l; ub[r] = u; }
        }
        __syncwarp();
        if (a.cost_kind == 1) {                           // mmd_random: beta = 1/nr, sigma = 0.01, lane = 0  [cem.py:355-356, 404-424]
            for (int i = lane; i < nr; i += 32) { bet[i] = c.beta_del; a.beta[(size_t)g * nr + i] = c.beta_del; }
            __syncwarp();
            const float risk = mmd_cost_warp(c, bet, cst, c.sigma_random, lane);
            if (lane == 0) { a.sigma[g] = c.sigma_random; a.risk[g] = risk; a.lane[g] = 0.0f; }
        } else if (lane == 0) {
            float risk, lanec;
            if (a.cost_kind == 2) {                       // cvar  [costs.py:206-221, 137-158]
                risk = cvar_cost(c, cst);
                lanec = cvar_cost(c, lb) + cvar_cost(c, ub);
            } else {                                      // saa  [costs.py:223-234, 160-171]
                float s = 0.0f, sl = 0.0f, su = 0.0f;
                for (int i = 0; i < nr; i++) { s = s + (cst[i] > 0.0f ? 1.0f : 0.0f); sl = sl + (lb[i] > 0.0f ? 1.0f : 0.0f); su = su + (ub[i] > 0.0f ? 1.0f : 0.0f); }
                risk = s / (float)nr;
                lanec = (sl + su) / (float)nr;
            }
            a.risk[g] = risk; a.lane[g] = lanec;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// k_inner_cem (mmd_opt): one CTA (3 warps) per sample: distance table, the reduced-set CEM, MMD risk of the chosen set.
// 89 = S - n_elite new beta samples are evaluated per inner iteration (the 11 elites keep last iteration's cost: same
// row => same arithmetic => same bits), so 96 threads are ~93 % busy in the sample stage and in the row-per-thread
// resampling stage.
#define RISKO_THREADS 96

struct OptLayout {          // shared-memory carve-up (in floats); every offset is a multiple of 4 floats (16 B)
    int F, D, small, red, th, cost, betas, idxs, key64, perm, C, rd, mean, eth, xc, ecost, ebetas, eidxs;
    int ldc, total;
};
__host__ __device__ inline int al4(int x) { return (x + 3) & ~3; }
__host__ __device__ inline OptLayout opt_layout(int nr, int np, int S, int ne) {
    OptLayout L; const int nm = nr * nr, d = nm + 1;
    (void)np;
    L.ldc = al4(d);
    int q = 0;
    L.F = q; q += al4(nm * 2 * NV); L.D = q; q += al4(nm * nm); L.small = q; q += 64;
    L.red = q; q += al4(3 * 16 * (RISKO_THREADS / 32));
    L.th = q; q += al4(S * d); L.cost = q; q += al4(S); L.betas = q; q += al4(S * nr); L.idxs = q; q += al4(S * nr);
    L.key64 = q; q += al4(2 * S); L.perm = q; q += al4(S); L.C = q; q += al4(d * L.ldc); L.rd = q; q += al4(d); L.mean = q; q += al4(d);
    L.eth = q; q += al4(ne * d); L.xc = q; q += al4(ne * d); L.ecost = q; q += al4(ne); L.ebetas = q; q += al4(ne * nr); L.eidxs = q; q += al4(ne * nr);
    L.total = q;
    return L;
}

// one beta sample of the inner CEM: choose the top-NR |theta| (stable, ascending), build the Laplace kernels from the
// chain's distance table, solve the equality-constrained QP by Cholesky block elimination, return the MMD cost
// [compute_beta.py:113-129, 70-91].  |theta| is compared through its bit pattern (non-negative floats order like
// integers and NaN patterns sort last), which is the jnp.argsort order.
template <int NR>
__device__ __forceinline__ float beta_sample(const DCfg& c, const float* __restrict__ row, const float* __restrict__ D,
                                             float* __restrict__ beta_out, int* __restrict__ idx_out) {
    constexpr int nm = NR * NR;
    int tv[NR], ti[NR];
#pragma unroll
    for (int i = 0; i < NR; i++) { tv[i] = -1; ti[i] = -1; }
#pragma unroll 1
    for (int m = 0; m < nm; m++) {                 // branch-free insertion: replace the smallest, bubble it up past <= elements
        const int v = (int)(dm::f2u(row[m]) & 0x7fffffffu);
        bool mv = v >= tv[0];
        tv[0] = mv ? v : tv[0]; ti[0] = mv ? m : ti[0];
#pragma unroll
        for (int p = 0; p < NR - 1; p++) {
            mv = mv && (tv[p] >= tv[p + 1]);
            const int a0 = tv[p], a1 = tv[p + 1], b0 = ti[p], b1 = ti[p + 1];
            tv[p] = mv ? a1 : a0; tv[p + 1] = mv ? a0 : a1; ti[p] = mv ? b1 : b0; ti[p + 1] = mv ? b0 : b1;
        }
    }
    const float sigma = row[nm];
    const float rinv = 1.0f / sigma;
    float rowsum[NR];
#pragma unroll
    for (int i = 0; i < NR; i++) rowsum[i] = 0.0f;
#pragma unroll 1
    for (int m = 0; m < nm; m++) {                 // D is symmetric bit for bit: read column m (bank-conflict free across threads)
        const float* Dm = D + m * nm;
#pragma unroll
        for (int i = 0; i < NR; i++) rowsum[i] = rowsum[i] + dm::exp_nonpos(-(Dm[ti[i]] * rinv));
    }
    float K[NR][NR];                               // ker_red (symmetric bit for bit); diagonal: exp(-(0 * rinv)) = 1
#pragma unroll
    for (int i = 0; i < NR; i++) {
        K[i][i] = 1.0f;
#pragma unroll
        for (int j = 0; j < i; j++) { const float k = dm::exp_nonpos(-(D[ti[i] * nm + ti[j]] * rinv)); K[i][j] = k; K[j][i] = k; }
    }
    // A = ker_red + 0.05 I, Cholesky with reciprocal pivots
    float Lm[NR][NR], rd[NR], u[NR], w[NR];
#pragma unroll
    for (int j = 0; j < NR; j++) {
        float acc = K[j][j] + 0.05f;
#pragma unroll
        for (int k = 0; k < j; k++) acc = fmaf(-Lm[j][k], Lm[j][k], acc);
        const float dd = sqrtf(acc);
        Lm[j][j] = dd; rd[j] = 1.0f / dd;
#pragma unroll
        for (int i = j + 1; i < NR; i++) {
            float aa = K[i][j];
#pragma unroll
            for (int k = 0; k < j; k++) aa = fmaf(-Lm[i][k], Lm[j][k], aa);
            Lm[i][j] = aa * rd[j];
        }
    }
#pragma unroll
    for (int i = 0; i < NR; i++) {
        float aa = c.inv_nm * rowsum[i], bb = 1.0f;
#pragma unroll
        for (int k = 0; k < i; k++) { aa = fmaf(-Lm[i][k], u[k], aa); bb = fmaf(-Lm[i][k], w[k], bb); }
        u[i] = aa * rd[i]; w[i] = bb * rd[i];
    }
#pragma unroll
    for (int i = NR - 1; i >= 0; i--) {
        float aa = u[i], bb = w[i];
#pragma unroll
        for (int k = i + 1; k < NR; k++) { aa = fmaf(-Lm[k][i], u[k], aa); bb = fmaf(-Lm[k][i], w[k], bb); }
        u[i] = aa * rd[i]; w[i] = bb * rd[i];
    }
    float su = 0.0f, sw = 0.0f;
#pragma unroll
    for (int i = 0; i < NR; i++) { su = su + u[i]; sw = sw + w[i]; }
    const float nu = (su - 1.0f) / sw;
    float beta[NR];
#pragma unroll
    for (int i = 0; i < NR; i++) beta[i] = fmaf(-nu, w[i], u[i]);
    float s1 = 0.0f, s2 = 0.0f;
#pragma unroll
    for (int i = 0; i < NR; i++) {
        float t = 0.0f;
#pragma unroll
        for (int j = 0; j < NR; j++) t = fmaf(K[i][j], beta[j], t);
        s1 = fmaf(beta[i], t, s1);
        s2 = fmaf(c.m2_inv_nm * rowsum[i], beta[i], s2);
    }
#pragma unroll
    for (int i = 0; i < NR; i++) { beta_out[i] = beta[i]; idx_out[i] = ti[i]; }
    return s1 + s2;
}

// float -> int64 key whose signed order is "ascending float, -0 == +0, NaN last", tie-broken by the index in the low bits
__device__ __forceinline__ long long sort_key64(float x, int idx) {
    const float xz = x + 0.0f;                                   // -0 -> +0
    const uint32_t ubits = dm::f2u(xz);
    int k = (ubits & 0x80000000u) ? (int)(~ubits ^ 0x80000000u) : (int)ubits;
    if (xz != xz) k = 0x7fffffff;
    return ((long long)k << 12) | (long long)idx;
}

template <int NR>
__global__ void __launch_bounds__(RISKO_THREADS, (NR <= 5) ? 7 : 1) k_inner_cem(DCfg c, RollArgs ra) {
    extern __shared__ __align__(16) float sm[];
    const RiskArgs& a = ra.r;
    const int g = blockIdx.x;
    if (g >= a.n_samples) return;
    constexpr int nm = NR * NR, d = nm + 1;
    constexpr bool SMALL = d <= 32;
    const int tid = threadIdx.x, nt = RISKO_THREADS, warp = tid >> 5, lane = tid & 31;
    const int e = g / a.B, np = c.np, S = c.S_in, ne = c.n_el_in;
    const OptLayout L = opt_layout(NR, np, S, ne);
    const int ldc = L.ldc;
    float* F = sm + L.F; float* D = sm + L.D; float* small = sm + L.small;
    // ---- distance table of the mother features  [kernel_computation.py:31-33]
    {
        const float* Fg = ra.feat + (size_t)g * nm * 2 * NV;
#pragma unroll 1
        for (int i = tid; i < nm * 2 * NV; i += nt) F[i] = Fg[i];
    }
    __syncthreads();
#pragma unroll 1
    for (int i = tid; i < nm * nm; i += nt) {
        const float* Fa = F + (i / nm) * 2 * NV; const float* Fb = F + (i % nm) * 2 * NV;
        float dist = 0.0f;
#pragma unroll
        for (int f = 0; f < 2 * NV; f++) dist = dist + fabsf(Fa[f] - Fb[f]);
        D[i] = dist;
    }
    // ---- inner CEM  [compute_beta.py:93-157]
    float* th = sm + L.th; float* cost = sm + L.cost; float* betas = sm + L.betas; int* idxs = (int*)(sm + L.idxs);
    long long* key64 = (long long*)(sm + L.key64); int* perm = (int*)(sm + L.perm);
    float* C = sm + L.C; float* mean = sm + L.mean; float* eth = sm + L.eth; float* xc = sm + L.xc;
    float* ecost = sm + L.ecost; float* ebetas = sm + L.ebetas; int* eidxs = (int*)(sm + L.eidxs);
#pragma unroll 1
    for (int i = tid; i < S * d; i += nt) th[i] = __ldg(c.theta0 + i);
    __syncthreads();
    float* resb = a.res_beta + (size_t)g * c.iters_in;
#pragma unroll 1
    for (int it = 0; it < c.iters_in; it++) {
        // -- evaluate the new samples (all of them in iteration 0; afterwards rows 0..ne-1 are last iteration's elites)
#pragma unroll 1
        for (int s = (it == 0 ? 0 : ne) + tid; s < S; s += nt) cost[s] = beta_sample<NR>(c, th + s * d, D, betas + s * NR, idxs + s * NR);
        __syncthreads();
#pragma unroll 1
        for (int s = tid; s < S; s += nt) key64[s] = sort_key64(cost[s], s);
        __syncthreads();
#pragma unroll 1
        for (int s = tid; s < S; s += nt) {        // stable argsort by rank counting; only the ne best are needed
            const long long ks = key64[s];
            int rank = 0;
#pragma unroll 4
            for (int j = 0; j < S; j++) rank += (key64[j] < ks) ? 1 : 0;
            if (rank < ne) perm[rank] = s;
        }
        __syncthreads();
        // -- gather the elites (rank order) and their mean  [compute_beta.py:56-60]
#pragma unroll 1
        for (int i = tid; i < ne * d; i += nt) eth[i] = th[perm[i / d] * d + (i % d)];
#pragma unroll 1
        for (int i = tid; i < ne * NR; i += nt) { const int src = perm[i / NR] * NR + (i % NR); ebetas[i] = betas[src]; eidxs[i] = idxs[src]; }
#pragma unroll 1
        for (int i = tid; i < ne; i += nt) ecost[i] = cost[perm[i]];
#pragma unroll 1
        for (int i = tid; i < d; i += nt) {
            float s = 0.0f;
            for (int el = 0; el < ne; el++) s = s + th[perm[el] * d + i];
            mean[i] = s / (float)ne;
        }
        __syncthreads();
#pragma unroll 1
        for (int i = tid; i < ne * d; i += nt) { const float v = eth[i]; th[i] = v; xc[i] = v - mean[i % d]; }
#pragma unroll 1
        for (int i = tid; i < ne * NR; i += nt) { betas[i] = ebetas[i]; idxs[i] = eidxs[i]; }
#pragma unroll 1
        for (int i = tid; i < ne; i += nt) cost[i] = ecost[i];
        __syncthreads();
#pragma unroll 1
        for (int i = tid; i < d * d; i += nt) {    // jnp.cov (ddof = 1) + 0.05 I, lower triangle  [compute_beta.py:61]
            const int r = i / d, q = i % d;
            if (q <= r) {
                float acc = 0.0f;
                for (int el = 0; el < ne; el++) acc = fmaf(xc[el * d + r], xc[el * d + q], acc);
                acc = acc / (float)(ne - 1);
                if (r == q) acc = acc + 0.05f;
                C[r * ldc + q] = acc;
            }
        }
        __syncthreads();
        if constexpr (SMALL) {
            // -- Cholesky by warp 0, one row per lane, left-looking: acc = a_ij - sum_k<j L_ik L_jk as an ascending-k fma chain
            if (warp == 0) {
#pragma unroll 1
                for (int j = 0; j < d; j++) {
                    float acc = 0.0f;
                    if (lane >= j && lane < d) {
                        acc = C[lane * ldc + j];
                        for (int k = 0; k < j; k++) acc = fmaf(-C[lane * ldc + k], C[j * ldc + k], acc);
                    }
                    const float ajj = __shfl_sync(FULL, acc, j);
                    const float dd = sqrtf(ajj);
                    const float rdj = 1.0f / dd;
                    if (lane == j) C[j * ldc + j] = dd;
                    else if (lane > j && lane < d) C[lane * ldc + j] = acc * rdj;
                    __syncwarp();
                }
            }
            __syncthreads();
            // -- resample: one thread per new row, normals in registers, L broadcast from shared memory  [compute_beta.py:63-66]
            const int nrow = S - ne;
            const float* zT = c.zb_iterT + (size_t)it * d * nrow;
#pragma unroll 1
            for (int r = tid; r < nrow; r += nt) {
                float z[d];
#pragma unroll
                for (int k = 0; k < d; k++) z[k] = __ldg(zT + k * nrow + r);
                float* dst = th + (ne + r) * d;
#pragma unroll
                for (int q = 0; q < d; q++) {
                    const float4* Cq = reinterpret_cast<const float4*>(C + q * ldc);
                    float acc = 0.0f;
#pragma unroll
                    for (int k4 = 0; k4 <= q / 4; k4++) {
                        const float4 l = Cq[k4];
                        if (4 * k4 + 0 <= q) acc = fmaf(l.x, z[4 * k4 + 0], acc);
                        if (4 * k4 + 1 <= q) acc = fmaf(l.y, z[4 * k4 + 1], acc);
                        if (4 * k4 + 2 <= q) acc = fmaf(l.z, z[4 * k4 + 2], acc);
                        if (4 * k4 + 3 <= q) acc = fmaf(l.w, z[4 * k4 + 3], acc);
                    }
                    float v = mean[q] + acc;
                    if (q == nm) v = (v != v) ? v : (v > c.sigma_clip ? v : c.sigma_clip);
                    dst[q] = v;
                }
            }
        } else {
            float* rd = sm + L.rd;
#pragma unroll 1
            for (int j = 0; j < d; j++) {          // block-wide Cholesky, column by column; contract order: ascending-k fma chain
                for (int r = j + tid; r < d; r += nt) {
                    float acc = C[r * ldc + j];
                    for (int k = 0; k < j; k++) acc = fmaf(-C[r * ldc + k], C[j * ldc + k], acc);
                    if (r == j) { const float dd = sqrtf(acc); C[j * ldc + j] = dd; rd[j] = 1.0f / dd; }
                    else C[r * ldc + j] = acc;                  // provisional, scaled by 1/L[j][j] after the barrier
                }
                __syncthreads();
                for (int r = j + 1 + tid; r < d; r += nt) C[r * ldc + j] = C[r * ldc + j] * rd[j];
                __syncthreads();
            }
            const float* z = c.zb_iter + (size_t)it * (S - ne) * d;
#pragma unroll 1
            for (int i = tid; i < (S - ne) * d; i += nt) {
                const int r = i / d, q = i % d;
                float v = mvn_elem(C, ldc, q, z + r * d, mean[q]);
                if (q == nm) v = (v != v) ? v : (v > c.sigma_clip ? v : c.sigma_clip);
                th[(ne + r) * d + q] = v;
            }
        }
        __syncthreads();
        if (tid == 0) {
            resb[it] = ecost[0];
            if (it == c.iters_in - 1) {            // beta / reduced set of the best sample; sigma from the RESAMPLED array [Q7]
                for (int i = 0; i < NR; i++) { small[i] = ebetas[i]; ((int*)small)[16 + i] = eidxs[i]; }
                small[48] = th[perm[0] * d + nm];
            }
        }
        __syncthreads();
    }
    // ---- risk of the chosen reduced set (its rollouts come back from global memory)  [costs.py:173-186, 121-135]
    const int* ridx = (const int*)small + 16;
    const float* xg = ra.xroll + (size_t)g * nm * np; const float* yg = ra.yroll + (size_t)g * nm * np;
    const float* xo = a.x_obs + (size_t)e * c.O * T_; const float* yo = a.y_obs + (size_t)e * c.O * T_;
    constexpr int NW = RISKO_THREADS / 32;
    float* red = sm + L.red;  // per-warp partial maxima, 3 * NR * NW floats
#pragma unroll 1
    for (int r = 0; r < NR; r++) {
        const float* xred = xg + ridx[r] * np; const float* yred = yg + ridx[r] * np;
        float m = 0.0f, l = 0.0f, u = 0.0f;
        for (int i = tid; i < c.O * np; i += nt) {
            const int o = i / np, t = i % np;
            m = dm::nmax_(m, fbar(c, xred[t], yred[t], xo[o * T_ + t], yo[o * T_ + t]));
        }
        for (int t = tid; t < np; t += nt) {
            l = dm::nmax_(l, dm::max0_(-yred[t] + c.y_lb));
            u = dm::nmax_(u, dm::max0_(yred[t] - c.y_ub));
        }
        m = warp_nmax(m); l = warp_nmax(l); u = warp_nmax(u);
        if (lane == 0) { red[(r * 3 + 0) * NW + warp] = m; red[(r * 3 + 1) * NW + warp] = l; red[(r * 3 + 2) * NW + warp] = u; }
    }
    __syncthreads();
    if (tid == 0) {
        float cs[NR], lbv[NR], ubv[NR], beta[NR];
        for (int r = 0; r < NR; r++) {
            float m = red[(r * 3 + 0) * NW], l = red[(r * 3 + 1) * NW], u = red[(r * 3 + 2) * NW];
            for (int wv = 1; wv < NW; wv++) { m = dm::nmax_(m, red[(r * 3 + 0) * NW + wv]); l = dm::nmax_(l, red[(r * 3 + 1) * NW + wv]); u = dm::nmax_(u, red[(r * 3 + 2) * NW + wv]); }
            cs[r] = m; lbv[r] = l; ubv[r] = u; beta[r] = small[r];
            a.beta[(size_t)g * NR + r] = small[r];
        }
        const float sigma = small[48];
        a.sigma[g] = sigma;
        a.risk[g] = mmd_cost(c, beta, cs, sigma);
        a.lane[g] = mmd_cost(c, beta, lbv, sigma) + mmd_cost(c, beta, ubv, sigma);
    }
}
