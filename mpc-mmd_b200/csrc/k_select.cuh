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
#define SEL_THREADS_BIG 1024
#define SEL_RANK_MAX 1024   // up to this batch size the 64-bit keys are staged in shared memory and ranked by counting

// float -> uint32 whose unsigned order is "ascending float, -0 == +0, NaN last" (jnp.argsort order)
__device__ __forceinline__ uint32_t sel_key32(float x) {
    const float xz = x + 0.0f;
    const uint32_t u = dm::f2u(xz);
    uint32_t k = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
    if (xz != xz) k = 0xffffffffu;
    return k;
}

__global__ void __launch_bounds__(1024) k_select(DCfg c, SelArgs a) {      // SEL_THREADS threads, SEL_THREADS_BIG for batches above SEL_RANK_MAX
    extern __shared__ __align__(128) float sm[];    // 64-bit composite keys [B]
    const int e = blockIdx.x;
    if (e >= a.n_ep) return;
    const int B = a.B, tid = threadIdx.x, nthr = blockDim.x, n20 = c.n_el_cost, n5 = c.n_el;
    unsigned long long* skey = reinterpret_cast<unsigned long long*>(sm);
    __shared__ int top[32], el[8];
    __shared__ float cost20[32], th[8][NPAR], L[NPAR * NPAR], mean_new[NPAR], wts[8], s_sumw;
    __shared__ int s_sel;
    const float* risk = a.risk + (size_t)e * B; const float* res = a.res_norm + (size_t)e * B;
    // position of sample i after the two stable argsorts = rank of (risk, res_norm, index): one 64-bit compare per pair
    if (B <= SEL_RANK_MAX) {
        for (int i = tid; i < B; i += nthr) skey[i] = ((unsigned long long)sel_key32(risk[i]) << 32) | (unsigned long long)sel_key32(res[i]);
        __syncthreads();
        for (int i = tid; i < B; i += nthr) {
            const unsigned long long ki = skey[i];
            int rank = 0;
#pragma unroll 4
            for (int j = 0; j < B; j++) { const unsigned long long kj = skey[j]; rank += (kj < ki || (kj == ki && j < i)) ? 1 : 0; }
            if (rank < n20) top[rank] = i;
        }
    } else {
        // large batches (scaled configurations): only the first n_el_cost positions are needed, so take them by n_el_cost rounds of
        // "smallest (key, index) above the previous winner" -- O(n_el_cost * B) instead of the O(B^2) rank count; keys are recomputed
        // from global memory (L2-resident) instead of being staged in shared memory
        __shared__ unsigned long long wk[32];
        __shared__ int wi[32];
        __shared__ unsigned long long last_k; __shared__ int last_i;
        if (tid == 0) { last_k = 0ull; last_i = -1; }
        __syncthreads();
        for (int r = 0; r < n20; r++) {
            const unsigned long long lk = last_k; const int li = last_i;
            unsigned long long bk = ~0ull; int bi = 0x7fffffff;
            for (int j = tid; j < B; j += nthr) {
                const unsigned long long kj = ((unsigned long long)sel_key32(risk[j]) << 32) | (unsigned long long)sel_key32(res[j]);
                const bool above = (r == 0) || kj > lk || (kj == lk && j > li);
                if (above && (kj < bk || (kj == bk && j < bi))) { bk = kj; bi = j; }
            }
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) {
                const unsigned long long ok = __shfl_xor_sync(FULL, bk, off); const int oi = __shfl_xor_sync(FULL, bi, off);
                if (ok < bk || (ok == bk && oi < bi)) { bk = ok; bi = oi; }
            }
            if ((tid & 31) == 0) { wk[tid >> 5] = bk; wi[tid >> 5] = bi; }
            __syncthreads();
            if (tid == 0) {
                for (int w2 = 1; w2 < nthr / 32; w2++) if (wk[w2] < bk || (wk[w2] == bk && wi[w2] < bi)) { bk = wk[w2]; bi = wi[w2]; }
                top[r] = bi; last_k = bk; last_i = bi;
            }
            __syncthreads();
        }
    }
    __syncthreads();
    if (tid < n20) cost20[tid] = a.cost_base[(size_t)e * B + top[tid]] + a.w_obs * risk[top[tid]];   // [cem_helper.py:253-261]
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
    // compute_shifted_samples  [cem_helper.py:280-291]: weights (one thread: 5 exps, sequential sum), then mean / covariance in parallel
    if (tid == 0) {
        float ce[8];
        for (int k = 0; k < n5; k++) ce[k] = cost20[el[k]];
        float wmin = ce[0]; int imin = 0;
        for (int k = 1; k < n5; k++) if (ce[k] < wmin) { wmin = ce[k]; imin = k; }
        float sum_w = 0.0f;
        for (int k = 0; k < n5; k++) { const float w = dm::exp_((-c.lam_inv) * (ce[k] - wmin)); wts[k] = w; sum_w = sum_w + w; }
        s_sumw = sum_w;
        s_sel = top[imin];                         // [Q1] idx_min of the 5 sorted costs applied to the 20 risk-sorted rows
    }
    __syncthreads();
    float* mean = a.mean + e * NPAR; float* cov = a.cov + e * NPAR * NPAR;
    if (tid < NPAR) {
        float sacc = 0.0f;
        for (int k = 0; k < n5; k++) sacc = sacc + th[k][tid] * wts[k];
        mean_new[tid] = c.one_m_alpha_mean * mean[tid] + c.alpha_mean * (sacc / s_sumw);
    }
    __syncthreads();
    if (tid < NPAR * NPAR) {
        const int i = tid / NPAR, j = tid % NPAR;
        float sacc = 0.0f;
        for (int k = 0; k < n5; k++) sacc = sacc + wts[k] * ((th[k][i] - mean_new[i]) * (th[k][j] - mean_new[j]));
        float v = c.one_m_alpha_cov * cov[tid] + c.alpha_cov * (sacc / s_sumw);
        v = (i == j) ? v + 0.01f : v;
        L[tid] = v;
        cov[tid] = v;
    }
    if (tid < NPAR) mean[tid] = mean_new[tid];
    __syncthreads();
    // Cholesky of the 8x8 covariance by 8 lanes of warp 0, left-looking; same operation order per entry as chol_serial
    if (tid < 32) {
        const int i = tid < NPAR ? tid : NPAR - 1;
#pragma unroll 1
        for (int j = 0; j < NPAR; j++) {
            float acc = L[i * NPAR + j];
            for (int k = 0; k < j; k++) acc = fmaf(-L[i * NPAR + k], L[j * NPAR + k], acc);
            const float ajj = __shfl_sync(FULL, acc, j);
            const float dd = sqrtf(ajj);
            const float rd = 1.0f / dd;
            __syncwarp();
            if (tid == j) L[j * NPAR + j] = dd;
            else if (tid > j && tid < NPAR) L[tid * NPAR + j] = acc * rd;
            __syncwarp();
        }
    }
    __syncthreads();
    // next batch = [elites ; mean + L z]  [cem_helper.py:292-312]
    if (tid < n5 * NPAR) params[tid] = th[tid / NPAR][tid % NPAR];
    sample_batch(c, mean_new, L, a.zcem + e * a.z_stride, B - n5, params + n5 * NPAR);
    // emitted row (only the last iteration's survives, cem.py:324-331)
    const int s = s_sel;
    const size_t gs = (size_t)e * B + s;
    if (tid < NV) { a.o_cx[e * NV + tid] = a.cx[gs * NV + tid]; a.o_cy[e * NV + tid] = a.cy[gs * NV + tid]; }
    if (tid == 0) { a.o_lane[e] = a.lane[gs]; a.o_obs[e] = risk[s]; a.o_sigma[e] = a.sigma[gs]; a.o_sel[e * a.sel_stride + a.it] = s; }
    if (tid < a.nr) a.o_beta[e * a.nr + tid] = a.beta[gs * a.nr + tid];
    if (tid < a.iters_in) a.o_res_beta[e * a.iters_in + tid] = a.res_beta[gs * a.iters_in + tid];
}
