// k_select.cuh -- elite selection, CEM mean/covariance update and resampling, one CTA per episode.
//
// Replaces the tail of one CEM iteration:
//   jnp.argsort(res_norm_batch) / jnp.argsort(mmd_obs) gathers     reference S/optimizer/cem.py:233-289
//   the obstacle-risk term of Helper.compute_cost                   S/optimizer/cem_helper.py:253-261
//   Helper.compute_ellite_samples                                   S/optimizer/cem_helper.py:264-271
//   Helper.compute_shifted_samples (+ comp_prod)                    S/optimizer/cem_helper.py:273-314
//   the emitted "best" row                                          S/optimizer/cem.py:308-315
// The two stable argsorts compose into ONE lexicographic order (risk, res_norm, index); ranks
// are obtained by counting, which is exact and needs no data movement.
#pragma once
#include "common.cuh"

struct SelArgs {
    int n_ep, B, it, nr, iters_in;
    float w_obs;                                   // weight_{mmd,cvar,saa}_obs  cem.py:161-163
    const float *res_norm, *risk, *lane, *cost_base;   // [E][B]
    float* params;                                 // [E][B][8]   replaced by the next batch
    float *mean, *cov;                             // [E][8], [E][64]
    const float* zcem; size_t z_stride;            // episode e at zcem + e*z_stride, (B-n_el, 8)
    const float *cx, *cy;                          // [E][B][11]
    const float *beta, *sigma, *res_beta;          // [E][B][nr], [E][B], [E][B][iters_in]
    float *o_cx, *o_cy, *o_lane, *o_obs, *o_beta, *o_sigma, *o_res_beta;
    int32_t* o_sel; int sel_stride;                // o_sel[e*sel_stride + it]
};

#define SEL_THREADS 128

// (risk, res_norm, index) lexicographic "j before i", NaN last in each key
__device__ __forceinline__ bool sel_before(float rj, float nj, int j, float ri, float ni, int i) {
    if (dm::lt_nanlast(rj, ri)) return true;
    if (dm::lt_nanlast(ri, rj)) return false;
    if (dm::lt_nanlast(nj, ni)) return true;
    if (dm::lt_nanlast(ni, nj)) return false;
    return j < i;
}

__global__ void __launch_bounds__(SEL_THREADS) k_select(DCfg c, SelArgs a) {
    extern __shared__ float sm[];                  // risk[B], res[B]
    const int e = blockIdx.x;
    if (e >= a.n_ep) return;
    const int B = a.B, tid = threadIdx.x, n20 = c.n_el_cost, n5 = c.n_el;
    float* srisk = sm; float* sres = sm + B;
    __shared__ int top[32], el[8];
    __shared__ float cost20[32], th[8][NPAR], L[NPAR * NPAR], mean_new[NPAR];
    __shared__ int s_sel;
    const float* risk = a.risk + (size_t)e * B; const float* res = a.res_norm + (size_t)e * B;
    for (int i = tid; i < B; i += SEL_THREADS) { srisk[i] = risk[i]; sres[i] = res[i]; }
    __syncthreads();
    for (int i = tid; i < B; i += SEL_THREADS) {   // position of sample i after the two stable argsorts
        const float ri = srisk[i], ni = sres[i];
        int rank = 0;
        for (int j = 0; j < B; j++) rank += sel_before(srisk[j], sres[j], j, ri, ni, i) ? 1 : 0;
        if (rank < n20) top[rank] = i;
    }
    __syncthreads();
    if (tid < n20) cost20[tid] = a.cost_base[(size_t)e * B + top[tid]] + a.w_obs * srisk[top[tid]];   // [cem_helper.py:253-261]
    __syncthreads();
    if (tid < n20) {                               // stable argsort of the 20 costs -> 5 elites  [cem_helper.py:264-271]
        const float v = cost20[tid];
        int rank = 0;
        for (int j = 0; j < n20; j++) { const float u = cost20[j]; rank += (dm::lt_nanlast(u, v) || (!dm::lt_nanlast(v, u) && j < tid)) ? 1 : 0; }
        if (rank < n5) el[rank] = tid;
    }
    __syncthreads();
    float* params = a.params + (size_t)e * B * NPAR;
    if (tid < n5 * NPAR) th[tid / NPAR][tid % NPAR] = params[top[el[tid / NPAR]] * NPAR + tid % NPAR];
    __syncthreads();
    if (tid == 0) {                                // compute_shifted_samples  [cem_helper.py:280-291]
        float ce[8], w[8], dif[8][NPAR];
        for (int k = 0; k < n5; k++) ce[k] = cost20[el[k]];
        float wmin = ce[0]; int imin = 0;
        for (int k = 1; k < n5; k++) if (ce[k] < wmin) { wmin = ce[k]; imin = k; }
        float sum_w = 0.0f;
        for (int k = 0; k < n5; k++) { w[k] = dm::exp_((-c.lam_inv) * (ce[k] - wmin)); sum_w = sum_w + w[k]; }
        float* mean = a.mean + e * NPAR; float* cov = a.cov + e * NPAR * NPAR;
        for (int i = 0; i < NPAR; i++) {
            float s = 0.0f;
            for (int k = 0; k < n5; k++) s = s + th[k][i] * w[k];
            mean_new[i] = c.one_m_alpha_mean * mean[i] + c.alpha_mean * (s / sum_w);
        }
        for (int k = 0; k < n5; k++) for (int i = 0; i < NPAR; i++) dif[k][i] = th[k][i] - mean_new[i];
        for (int i = 0; i < NPAR; i++)
            for (int j = 0; j < NPAR; j++) {
                float s = 0.0f;
                for (int k = 0; k < n5; k++) s = s + w[k] * (dif[k][i] * dif[k][j]);
                float v = c.one_m_alpha_cov * cov[i * NPAR + j] + c.alpha_cov * (s / sum_w);
                v = (i == j) ? v + 0.01f : v;
                L[i * NPAR + j] = v;
            }
        for (int i = 0; i < NPAR; i++) mean[i] = mean_new[i];
        for (int i = 0; i < NPAR * NPAR; i++) cov[i] = L[i];
        float rd[NPAR];
        chol_serial(L, NPAR, NPAR, rd);
        s_sel = top[imin];                         // [Q1] idx_min of the 5 sorted costs applied to the 20 risk-sorted rows
    }
    __syncthreads();
    // next batch = [elites ; mean + L z]  [cem_helper.py:292-312]
    if (tid < n5 * NPAR) params[tid] = th[tid / NPAR][tid % NPAR];
    sample_batch(c, mean_new, L, a.zcem + e * a.z_stride, B - n5, params + n5 * NPAR);
    // emitted row (only the last iteration's survives, cem.py:324-331)
    const int s = s_sel;
    const size_t gs = (size_t)e * B + s;
    if (tid < NV) { a.o_cx[e * NV + tid] = a.cx[gs * NV + tid]; a.o_cy[e * NV + tid] = a.cy[gs * NV + tid]; }
    if (tid == 0) { a.o_lane[e] = a.lane[gs]; a.o_obs[e] = srisk[s]; a.o_sigma[e] = a.sigma[gs]; a.o_sel[e * a.sel_stride + a.it] = s; }
    if (tid < a.nr) a.o_beta[e * a.nr + tid] = a.beta[gs * a.nr + tid];
    if (tid < a.iters_in) a.o_res_beta[e * a.iters_in + tid] = a.res_beta[gs * a.iters_in + tid];
}
