// k_risk.cuh -- noisy control rollouts, reduced-set selection and the risk functionals.
//
// Replaces, per CEM sample:
//   Helper.compute_rollout_complete_baseline / _opt + compute_rollout_one_step   S/optimizer/cem_helper.py:402-538, 380-400
//   Helper.compute_coeff                                                         S/optimizer/cem_helper.py:553-564
//   beta_cem.compute_cem (+ compute_mean_cov_beta, compute_beta_reduced)          S/compute_beta.py:93-157, 51-91
//   kernel_matrix.compute_kernel / compute_mmd (Laplace kernel)                   S/kernel_computation.py:19-87
//   Costs.compute_f_bar / compute_lane_bar / compute_{mmd,cvar,saa}_{obs,lane}     S/optimizer/costs.py:50-71, 121-234
// k_risk_base: one warp per sample (cvar / saa / mmd_random: num_reduced rollouts).
// k_risk_opt : one CTA per sample (mmd_opt: num_reduced^2 mother rollouts + the inner reduced-set CEM,
//              the dominant cost of a solve), all state in shared memory.
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
    const float *x_obs, *y_obs;      // [E][O][100]
    float *risk, *lane;              // [n]
    float *beta, *sigma, *res_beta;  // [n][nr], [n], [n][iters_in]
};

// perturbed controls (nr,np)  [cem_helper.py:405-443 / 470-508]; threads tid, tid+nt, ...
__device__ __forceinline__ void noisy_controls(const DCfg& c, const float* acc, const float* steer, const float* z1, const float* z2,
                                               const float* z3, const uint32_t* keys, float* an, float* sn, int tid, int nt) {
    const int np = c.np, n = c.nr * np;
    for (int i = tid; i < n; i += nt) {
        const int t = i % np;
        const float a = acc[t], s = steer[t];
        float pa, ps;
        if (c.noise_kind == 0) {
            pa = (c.sigma_acc * fabsf(a)) * z1[i];
            ps = (c.sigma_steer * fabsf(s)) * z2[i];
        } else {
            dr::Key k1, k2; k1.k0 = keys[0]; k1.k1 = keys[1]; k2.k0 = keys[2]; k2.k1 = keys[3];
            float b1 = dr::beta_elem(k1, (uint32_t)n, (uint32_t)i, c.beta_a * fabsf(a), c.beta_b * fabsf(a));
            float b2 = dr::beta_elem(k2, (uint32_t)n, (uint32_t)i, c.beta_a * fabsf(s), c.beta_b * fabsf(s));
            pa = c.sigma_acc * (2.0f * b1 - 1.0f);
            ps = c.ksig_steer * (2.0f * b2 - 1.0f);
        }
        an[i] = (a + pa) + c.acc_const * z3[i];
        sn[i] = (s + ps) + c.steer_const * z3[i];
    }
}
// Euler bicycle rollout, records the state BEFORE each step  [cem_helper.py:380-400, 451-458]
__device__ __forceinline__ void rollout_one(const DCfg& c, const float* a, const float* s, const float* st0, float* xr, float* yr) {
    float x = st0[0], y = st0[1], vx = st0[2], vy = st0[3], psi = st0[4];
    for (int t = 0; t < c.np; t++) {
        xr[t] = x; yr[t] = y;
        float v = sqrtf(vx * vx + vy * vy);
        v = v + a[t] * c.dt;
        float psidot = (v * dm::tan_(s[t])) / c.wheel_base;
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
__device__ float mmd_cost(const DCfg& c, const float* beta, const float* cost, float sigma) {
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
// jnp.quantile (linear interpolation) + mean of the tail  [costs.py:213-220]
__device__ float cvar_cost(const DCfg& c, const float* v) {
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
#define RISKB_WARPS 4
__host__ __device__ inline int riskb_warp_floats(int nr, int np) { return 4 * nr * np + 3 * 16; }

__global__ void __launch_bounds__(RISKB_WARPS * 32) k_risk_base(DCfg c, RiskArgs a) {
    extern __shared__ float sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = blockIdx.x * RISKB_WARPS + warp;
    if (g >= a.n_samples) return;
    const int e = g / a.B, nr = c.nr, np = c.np, n = nr * np;
    float* an = sm + warp * riskb_warp_floats(nr, np); float* sn = an + n; float* xr = sn + n; float* yr = xr + n;
    float* cst = yr + n; float* lb = cst + 16; float* ub = lb + 16;
    noisy_controls(c, a.acc + (size_t)g * T_, a.steer + (size_t)g * T_, a.z1 + e * a.z_stride, a.z2 + e * a.z_stride,
                   a.z3 + e * a.z_stride, a.keys + e * a.key_stride, an, sn, lane, 32);
    __syncwarp();
    for (int r = lane; r < nr; r += 32) rollout_one(c, an + r * np, sn + r * np, a.state0 + e * 5, xr + r * np, yr + r * np);
    __syncwarp();
    const float* xo = a.x_obs + (size_t)e * c.O * T_; const float* yo = a.y_obs + (size_t)e * c.O * T_;
    for (int r = 0; r < nr; r++) {
        float m = 0.0f, l = 0.0f, u = 0.0f;
        for (int i = lane; i < c.O * np; i += 32) {
            const int o = i / np, t = i % np;
            m = dm::nmax_(m, fbar(c, xr[r * np + t], yr[r * np + t], xo[o * T_ + t], yo[o * T_ + t]));
        }
        for (int t = lane; t < np; t += 32) {
            l = dm::nmax_(l, dm::max0_(-yr[r * np + t] + c.y_lb));
            u = dm::nmax_(u, dm::max0_(yr[r * np + t] - c.y_ub));
        }
        m = warp_nmax(m); l = warp_nmax(l); u = warp_nmax(u);
        if (lane == 0) { cst[r] = m; lb[r] = l; ub[r] = u; }
    }
    __syncwarp();
    if (lane == 0) {
        float risk, lanec;
        if (a.cost_kind == 1) {                       // mmd_random: beta = 1/nr, sigma = 0.01, lane = 0  [cem.py:355-356, 404-424]
            float beta[MPCMMD_MAX_NR_DEV];
            for (int i = 0; i < nr; i++) { beta[i] = c.beta_del; a.beta[(size_t)g * nr + i] = c.beta_del; }
            a.sigma[g] = c.sigma_random;
            risk = mmd_cost(c, beta, cst, c.sigma_random);
            lanec = 0.0f;
        } else if (a.cost_kind == 2) {                // cvar  [costs.py:206-221, 137-158]
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

// ---------------------------------------------------------------------------------------------
// mmd_opt: one CTA per sample
#define RISKO_THREADS 128

struct OptLayout {          // shared-memory carve-up (in floats)
    int an, sn, F, D, small, red, uni;   // persistent region offsets; `uni` = start of the aliased region
    int xr, yr;                          // phase 1 (mother rollouts)
    int th, thn, cost, betas, idxs, perm, C, rd, mean, xc;   // phase 2 (inner CEM)
    int xred, yred;                      // phase 3 (rollouts of the chosen reduced set)
    int total;
};
__host__ __device__ inline OptLayout opt_layout(int nr, int np, int S, int ne) {
    OptLayout L; const int nm = nr * nr, d = nm + 1;
    int o = 0;
    L.an = o; o += nr * np; L.sn = o; o += nr * np;
    L.F = o; o += nm * 2 * NV; L.D = o; o += nm * nm;
    L.small = o; o += 64;
    L.red = o; o += 3 * MPCMMD_MAX_NR_DEV * (RISKO_THREADS / 32);
    L.uni = o;
    L.xr = o; L.yr = o + nm * np; int p1 = o + 2 * nm * np;
    int q = o;
    L.th = q; q += S * d; L.thn = q; q += S * d; L.cost = q; q += S; L.betas = q; q += S * nr; L.idxs = q; q += S * nr;
    L.perm = q; q += S; L.C = q; q += d * d; L.rd = q; q += d; L.mean = q; q += d; L.xc = q; q += ne * d;
    L.xred = o; L.yred = o + nr * np; int p3 = o + 2 * nr * np;
    L.total = p1 > q ? p1 : q; if (p3 > L.total) L.total = p3;
    return L;
}

// one beta sample of the inner CEM: choose the top-NR |theta|, build the Laplace kernels, solve the
// equality-constrained QP by Cholesky block elimination, return the MMD cost  [compute_beta.py:113-129, 70-91]
template <int NR>
__device__ __forceinline__ float beta_sample(const DCfg& c, const float* row, const float* D, float* beta_out, int* idx_out) {
    const int nm = NR * NR;
    float tv[NR]; int ti[NR];
#pragma unroll
    for (int i = 0; i < NR; i++) { tv[i] = -1.0f; ti[i] = -1; }
    for (int m = 0; m < nm; m++) {                 // stable "last NR of argsort(|theta|)"
        const float v = fabsf(row[m]);
        if (!dm::lt_nanlast(v, tv[0])) {
            tv[0] = v; ti[0] = m;
            bool mv = true;
#pragma unroll
            for (int p = 0; p < NR - 1; p++) {
                mv = mv && !dm::lt_nanlast(tv[p], tv[p + 1]);
                if (mv) { float fv = tv[p]; tv[p] = tv[p + 1]; tv[p + 1] = fv; int iv = ti[p]; ti[p] = ti[p + 1]; ti[p + 1] = iv; }
            }
        }
    }
    const float sigma = row[nm];
    const float rinv = 1.0f / sigma;
    float rowsum[NR];
    float K[NR][NR];                               // ker_red, full (symmetric bit for bit)
#pragma unroll
    for (int i = 0; i < NR; i++) {
        const float* Di = D + ti[i] * nm;
        float rs = 0.0f;
        for (int m = 0; m < nm; m++) rs = rs + dm::exp_nonpos(-(Di[m] * rinv));
        rowsum[i] = rs;
#pragma unroll
        for (int j = 0; j <= i; j++) { float k = dm::exp_nonpos(-(Di[ti[j]] * rinv)); K[i][j] = k; K[j][i] = k; }
    }
    // A = ker_red + 0.05 I, Cholesky with reciprocal pivots
    float Lm[NR][NR], rd[NR], u[NR], w[NR];
#pragma unroll
    for (int j = 0; j < NR; j++) {
        float acc = K[j][j] + 0.05f;
#pragma unroll
        for (int k = 0; k < j; k++) acc = fmaf(-Lm[j][k], Lm[j][k], acc);
        float dd = sqrtf(acc);
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

template <int NR>
__global__ void __launch_bounds__(RISKO_THREADS) k_risk_opt(DCfg c, RiskArgs a) {
    extern __shared__ float sm[];
    const int g = blockIdx.x;
    if (g >= a.n_samples) return;
    const int tid = threadIdx.x, nt = RISKO_THREADS;
    const int e = g / a.B, np = c.np, nm = NR * NR, d = nm + 1, S = c.S_in, ne = c.n_el_in;
    const OptLayout L = opt_layout(NR, np, S, ne);
    float* an = sm + L.an; float* sn = sm + L.sn; float* F = sm + L.F; float* D = sm + L.D; float* small = sm + L.small;
    float* xr = sm + L.xr; float* yr = sm + L.yr;
    // ---- phase 1: controls, mother rollouts, features, distance table
    noisy_controls(c, a.acc + (size_t)g * T_, a.steer + (size_t)g * T_, a.z1 + e * a.z_stride, a.z2 + e * a.z_stride,
                   a.z3 + e * a.z_stride, a.keys + e * a.key_stride, an, sn, tid, nt);
    __syncthreads();
    const float* st0 = a.state0 + e * 5;
    for (int m = tid; m < nm; m += nt)             // mother sample m = i*nr + j: acc noise i, steer noise j  [cem_helper.py:510-511]
        rollout_one(c, an + (m / NR) * np, sn + (m % NR) * np, st0, xr + m * np, yr + m * np);
    __syncthreads();
    for (int i = tid; i < nm * 2 * NV; i += nt) {  // ridge-fit features  [cem_helper.py:553-564, folded]
        const int m = i / (2 * NV), k = i % (2 * NV);
        const float* src = (k < NV) ? xr + m * np : yr + m * np;
        const float* W = c.Wfit + (k < NV ? k : k - NV) * np;
        float acc = 0.0f;
        for (int t = 0; t < np; t++) acc = fmaf(W[t], src[t], acc);
        F[i] = acc;
    }
    __syncthreads();
    for (int i = tid; i < nm * nm; i += nt) {      // L1 distances between mother features  [kernel_computation.py:31-33]
        const float* Fa = F + (i / nm) * 2 * NV; const float* Fb = F + (i % nm) * 2 * NV;
        float dist = 0.0f;
#pragma unroll
        for (int f = 0; f < 2 * NV; f++) dist = dist + fabsf(Fa[f] - Fb[f]);
        D[i] = dist;
    }
    __syncthreads();                               // xr / yr are dead from here on (aliased by the CEM state)
    // ---- phase 2: inner CEM  [compute_beta.py:93-157]
    float* th = sm + L.th; float* thn = sm + L.thn; float* cost = sm + L.cost; float* betas = sm + L.betas;
    int* idxs = (int*)(sm + L.idxs); int* perm = (int*)(sm + L.perm);
    float* C = sm + L.C; float* rd = sm + L.rd; float* mean = sm + L.mean; float* xc = sm + L.xc;
    for (int i = tid; i < S * d; i += nt) th[i] = c.theta0[i];
    __syncthreads();
    float* resb = a.res_beta + (size_t)g * c.iters_in;
    for (int it = 0; it < c.iters_in; it++) {
        for (int s = tid; s < S; s += nt) cost[s] = beta_sample<NR>(c, th + s * d, D, betas + s * NR, idxs + s * NR);
        __syncthreads();
        for (int s = tid; s < S; s += nt) {        // stable argsort by rank counting (NaN last)
            const float v = cost[s];
            int rank = 0;
            for (int j = 0; j < S; j++) { const float u = cost[j]; rank += (dm::lt_nanlast(u, v) || (!dm::lt_nanlast(v, u) && j < s)) ? 1 : 0; }
            perm[rank] = s;
        }
        __syncthreads();
        for (int i = tid; i < ne * d; i += nt) thn[i] = th[perm[i / d] * d + (i % d)];      // elites keep their rank order
        __syncthreads();
        for (int i = tid; i < d; i += nt) {        // mean over the elites  [compute_beta.py:60]
            float s = 0.0f;
            for (int el = 0; el < ne; el++) s = s + thn[el * d + i];
            mean[i] = s / (float)ne;
        }
        __syncthreads();
        for (int i = tid; i < ne * d; i += nt) xc[i] = thn[i] - mean[i % d];
        __syncthreads();
        for (int i = tid; i < d * d; i += nt) {    // jnp.cov (ddof = 1) + 0.05 I, lower triangle  [compute_beta.py:61]
            const int r = i / d, q = i % d;
            if (q <= r) {
                float acc = 0.0f;
                for (int el = 0; el < ne; el++) acc = fmaf(xc[el * d + r], xc[el * d + q], acc);
                acc = acc / (float)(ne - 1);
                if (r == q) acc = acc + 0.05f;
                C[i] = acc;
            }
        }
        __syncthreads();
        for (int j = 0; j < d; j++) {              // Cholesky, column by column (rows in parallel); contract order: ascending-k fma chain
            for (int r = j + tid; r < d; r += nt) {
                float acc = C[r * d + j];
                for (int k = 0; k < j; k++) acc = fmaf(-C[r * d + k], C[j * d + k], acc);
                if (r == j) { const float dd = sqrtf(acc); C[j * d + j] = dd; rd[j] = 1.0f / dd; }
                else C[r * d + j] = acc;                        // provisional, scaled by 1/L[j][j] after the barrier
            }
            __syncthreads();
            for (int r = j + 1 + tid; r < d; r += nt) C[r * d + j] = C[r * d + j] * rd[j];
            __syncthreads();
        }
        const float* z = c.zb_iter + (size_t)it * (S - ne) * d;
        for (int i = tid; i < (S - ne) * d; i += nt) {         // resample  [compute_beta.py:63-64]
            const int r = i / d, q = i % d;
            thn[(ne + r) * d + q] = mvn_elem(C, d, q, z + r * d, mean[q]);
        }
        __syncthreads();
        for (int s = tid; s < S; s += nt) { float v = thn[s * d + nm]; thn[s * d + nm] = (v != v) ? v : (v > c.sigma_clip ? v : c.sigma_clip); }
        __syncthreads();
        if (tid == 0) {
            const int imin = perm[0];
            resb[it] = cost[imin];
            if (it == c.iters_in - 1) {            // beta / reduced set of the best sample; sigma from the RESAMPLED array [Q7]
                for (int i = 0; i < NR; i++) { small[i] = betas[imin * NR + i]; ((int*)small)[16 + i] = idxs[imin * NR + i]; }
                small[48] = thn[imin * d + nm];
            }
        }
        __syncthreads();
        float* tmp = th; th = thn; thn = tmp;
    }
    // ---- phase 3: risk of the chosen reduced set  [costs.py:173-186, 121-135]
    float* xred = sm + L.xred; float* yred = sm + L.yred;
    const int* ridx = (const int*)small + 16;
    for (int i = tid; i < NR; i += nt) {
        const int m = ridx[i];
        rollout_one(c, an + (m / NR) * np, sn + (m % NR) * np, st0, xred + i * np, yred + i * np);
    }
    __syncthreads();
    const float* xo = a.x_obs + (size_t)e * c.O * T_; const float* yo = a.y_obs + (size_t)e * c.O * T_;
    const int warp = tid >> 5, lane = tid & 31;
    float* red = sm + L.red;  // per-warp partial maxima, 3 * NR * (threads/32) floats
    for (int r = 0; r < NR; r++) {
        float m = 0.0f, l = 0.0f, u = 0.0f;
        for (int i = tid; i < c.O * np; i += nt) {
            const int o = i / np, t = i % np;
            m = dm::nmax_(m, fbar(c, xred[r * np + t], yred[r * np + t], xo[o * T_ + t], yo[o * T_ + t]));
        }
        for (int t = tid; t < np; t += nt) {
            l = dm::nmax_(l, dm::max0_(-yred[r * np + t] + c.y_lb));
            u = dm::nmax_(u, dm::max0_(yred[r * np + t] - c.y_ub));
        }
        m = warp_nmax(m); l = warp_nmax(l); u = warp_nmax(u);
        if (lane == 0) { red[(r * 3 + 0) * 4 + warp] = m; red[(r * 3 + 1) * 4 + warp] = l; red[(r * 3 + 2) * 4 + warp] = u; }
    }
    __syncthreads();
    if (tid == 0) {
        float cs[NR], lbv[NR], ubv[NR], beta[NR];
        for (int r = 0; r < NR; r++) {
            float m = red[(r * 3 + 0) * 4], l = red[(r * 3 + 1) * 4], u = red[(r * 3 + 2) * 4];
            for (int wv = 1; wv < RISKO_THREADS / 32; wv++) { m = dm::nmax_(m, red[(r * 3 + 0) * 4 + wv]); l = dm::nmax_(l, red[(r * 3 + 1) * 4 + wv]); u = dm::nmax_(u, red[(r * 3 + 2) * 4 + wv]); }
            cs[r] = m; lbv[r] = l; ubv[r] = u; beta[r] = small[r];
            a.beta[(size_t)g * NR + r] = small[r];
        }
        const float sigma = small[48];
        a.sigma[g] = sigma;
        a.risk[g] = mmd_cost(c, beta, cs, sigma);
        a.lane[g] = mmd_cost(c, beta, lbv, sigma) + mmd_cost(c, beta, ubv, sigma);
    }
}
