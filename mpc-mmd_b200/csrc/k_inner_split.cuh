// k_inner_split.cuh -- the reduced-set inner CEM of mmd_opt (num_reduced <= 5) as a PHASE-SPLIT pipeline: per inner iteration one
// embarrassingly parallel evaluation kernel and one warp-per-chain update kernel, chain state in an L2-resident global block.
//
//   k_icem_dist    once per chain: distance table D of the mother features                       [kernel_computation.py:31-33]
//   k_icem_eval    (it): thread per NEW beta sample: resample the row from (LT, mean, z) in registers -- it is never stored --, top-num_reduced,
//                  Laplace kernels, KKT solve, cost.  D and [LT | mean] arrive by TMA bulk copies.      [compute_beta.py:113-129, 70-91, 63-66]
//   k_icem_update  (it): ONE WARP per chain: stable selection of the elites, elite rows (new ones re-derived from the old factor: same
//                  operations => same bits), mean, covariance, panel Cholesky, outputs.                  [compute_beta.py:51-68, 131-157]
//
// Why: the fused one-CTA-per-chain kernel (k_inner_cem_fast) parks two of its three warps for the one-warp phases -- 37 % of all warp samples
// were barrier stalls, 22 % on the Cholesky alone (profiles/r01_v12_summary.md; ncu source view of v15) -- and its 56-register x 96-thread CTAs
// cap an SM at 12 chains.  Here no warp ever waits for another: the evaluation kernel is barrier-free after its operand load, the update kernel has
// no block barrier at all, and a launch of n chains exposes 89 n evaluation threads / n update warps, so small launches (the 8-GPU shard of
// the 200-episode sweep: 2500 chains) still fill the machine.  Arithmetic is the same contract, element for element, as k_inner_cem_fast.
#pragma once
#include "k_inner_cem.cuh"

struct SplitLayout {        // per-chain global state, in floats; every offset is a multiple of 4 floats (16 bytes: TMA bulk granularity)
    int D, LT, mean, eth, ecost, eb, ei, cost, total, ldc;
};
__host__ __device__ inline SplitLayout split_layout(int nr, int S, int ne) {
    SplitLayout L; const int nm = nr * nr, d = nm + 1;
    L.ldc = al4(d);
    int q = 0;
    L.D = q; q += al4(nm * nm);
    L.LT = q; q += d * L.ldc;                // transposed Cholesky factor LT[k][q] = L[q][k] (zeros for q < k) ...
    L.mean = q; q += L.ldc;                  // ... followed by the mean: ONE bulk copy brings both
    L.eth = q; q += ICF_MAX_NE * L.ldc;      // elite rows (rank order)
    L.ecost = q; q += al4(ICF_MAX_NE); L.eb = q; q += al4(ICF_MAX_NE * nr); L.ei = q; q += al4(ICF_MAX_NE);
    L.cost = q; q += al4(S);                 // costs of this iteration's new rows (eval -> update)
    L.total = al4(q + 31) & ~31;             // 128-byte multiple
    (void)ne;
    return L;
}
struct SplitArgs {
    int n_chains;            // chains of this launch; chain g = g0 + blockIdx-derived index
    int g0;
    float* state;            // [n][SplitLayout::total]
    const float* feat;       // [n][nm][22]
    float* bscratch;         // [n][S][nr + 1] per-row beta vectors + packed reduced-set indices of the current iteration
    float *beta, *sigma, *res_beta;   // outputs [n][nr], [n], [n][iters_in]
    int* ridx;               // [n][nr]
};

// ---- distance table: one thread per (chain, i, j)
template <int NR>
__global__ void __launch_bounds__(256) k_icem_dist(DCfg c, SplitArgs sa) {
    constexpr int nm = NR * NR;
    const SplitLayout L = split_layout(NR, c.S_in, c.n_el_in);
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)sa.n_chains * nm * nm) return;
    const int g = sa.g0 + (int)(idx / (nm * nm)), ij = (int)(idx % (nm * nm));
    const float* Fa = sa.feat + ((size_t)g * nm + ij / nm) * 2 * NV; const float* Fb = sa.feat + ((size_t)g * nm + ij % nm) * 2 * NV;
    float dist = 0.0f;
#pragma unroll
    for (int f = 0; f < 2 * NV; f++) dist = dist + fabsf(__ldg(Fa + f) - __ldg(Fb + f));
    sa.state[(size_t)g * L.total + L.D + ij] = dist;
}

// top-NR |theta| + evaluation of one beta sample whose row lives in registers (same keys / same exact path as beta_sample_fast)
template <int NR>
__device__ __forceinline__ float beta_sample_regs(const DCfg& c, const float (&row)[NR * NR + 1], const float* __restrict__ D,
                                                  float* __restrict__ beta_out, int* __restrict__ idx_out) {
    constexpr int nm = NR * NR;
    static_assert(nm <= 32, "index must fit in the 5 low key bits");
    int tk[NR + 1];
#pragma unroll
    for (int p = 0; p <= NR; p++) tk[p] = 0;
#pragma unroll
    for (int m = 0; m < nm; m++) {
        int v = (int)((dm::f2u(row[m]) & 0x7fffffe0u) | (uint32_t)m);
        tk[0] = max(tk[0], v);
#pragma unroll
        for (int p = 0; p < NR; p++) { const int lo = min(tk[p], tk[p + 1]), hi = max(tk[p], tk[p + 1]); tk[p] = lo; tk[p + 1] = hi; }
    }
    int ti[NR];
    bool near = false;
#pragma unroll
    for (int p = 0; p < NR; p++) { ti[p] = tk[p + 1] & 31; near |= ((tk[p] ^ tk[p + 1]) < 32); }
    if (near) {                                   // rare: two candidates share their upper 26 value bits -> exact (value, index) order
        float tmp[nm];
#pragma unroll
        for (int m = 0; m < nm; m++) tmp[m] = row[m];
        const int pkd = top_abs_exact<NR>(tmp);
#pragma unroll
        for (int p = 0; p < NR; p++) ti[p] = (pkd >> (5 * p)) & 31;
    }
    return beta_eval<NR>(c, ti, row[nm], D, beta_out, idx_out);
}

#define ICE_THREADS 96
// evaluation of the new rows of inner iteration `it` (it = 0: the S rows of the constant theta0 table; else the S - ne rows resampled from
// the factor / mean the previous update left in the chain block).  One CTA per chain.
template <int NR>
__global__ void __launch_bounds__(128, 9) k_icem_eval(DCfg c, SplitArgs sa, int it) {
    constexpr int nm = NR * NR, d = nm + 1, NG = (d + 3) / 4, NPAIR = (d + 1) / 2;
    extern __shared__ __align__(128) float sm[];
    __shared__ __align__(8) unsigned long long bar;
    const SplitLayout L = split_layout(NR, c.S_in, c.n_el_in);
    const int ldc = L.ldc, tid = threadIdx.x, S = c.S_in, ne = c.n_el_in;
    const int g = sa.g0 + blockIdx.x;
    float* st = sa.state + (size_t)g * L.total;
    float* D = sm; float* LT = sm + al4(nm * nm); float* mean = LT + d * ldc;
    if (tid == 0) {                                    // ONE arrival with the total byte count, then one or two bulk copies completing on it
        const uint32_t bD = (uint32_t)(((nm * nm + 3) & ~3) * sizeof(float));
        const uint32_t bL = (uint32_t)((d + 1) * ldc * sizeof(float));
        mbar_init(&bar, 1);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(&bar)), "r"(bD + (it > 0 ? bL : 0u)) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     :: "r"(smem_u32(D)), "l"(st + L.D), "r"(bD), "r"(smem_u32(&bar)) : "memory");
        if (it > 0)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         :: "r"(smem_u32(LT)), "l"(st + L.LT), "r"(bL), "r"(smem_u32(&bar)) : "memory");
    }
    __syncthreads();                                   // barrier initialised before anyone polls it
    mbar_wait(&bar, 0);
    const int n_new = it == 0 ? S : S - ne;
    float* betas = sa.bscratch + (size_t)g * S * (NR + 1); int* idxs = (int*)(betas + S * NR);
    float* cost = st + L.cost;
#pragma unroll 1
    for (int s = tid; s < n_new; s += blockDim.x) {
        float row[d];
        if (it == 0) {
#pragma unroll
            for (int k = 0; k < d; k++) row[k] = __ldg(c.theta0T + k * S + s);
        } else {
            const int nrow = S - ne;
            pk::f2 acc[2 * NG];
            icf_mvn_row_unrolled<d>(LT, ldc, c.zb_iterT + (size_t)(it - 1) * d * nrow, nrow, s, acc);
#pragma unroll
            for (int p = 0; p < NPAIR; p++) {
                float x0, x1; pk::unpack(acc[p], x0, x1);
                float v0 = mean[2 * p] + x0;
                if (2 * p == nm) v0 = (v0 != v0) ? v0 : (v0 > c.sigma_clip ? v0 : c.sigma_clip);
                row[2 * p] = v0;
                if (2 * p + 1 < d) {
                    float v1 = mean[2 * p + 1] + x1;
                    if (2 * p + 1 == nm) v1 = (v1 != v1) ? v1 : (v1 > c.sigma_clip ? v1 : c.sigma_clip);
                    row[2 * p + 1] = v1;
                }
            }
        }
        cost[s] = beta_sample_regs<NR>(c, row, D, betas + s * NR, idxs + s);
    }
}

// element q of resampled row r, re-derived from the factor: the operations of icf_mvn_row_unrolled on column q (k ascending through the end of q's
// group of four -- the k > q terms multiply stored zeros, exactly as there), then mean + acc and the sigma clip
template <int d>
__device__ __forceinline__ float icu_row_elem(const float* __restrict__ LT, int ldc, const float* __restrict__ mean, const float* __restrict__ zT,
                                              int nrow, int r, int q, float sigma_clip) {
    const int kend = min(4 * (q >> 2) + 3, d - 1);
    float acc = 0.0f;
#pragma unroll 4
    for (int k = 0; k <= kend; k++) acc = fmaf(LT[k * ldc + q], __ldg(zT + k * nrow + r), acc);
    float v = mean[q] + acc;
    if (q == d - 1) v = (v != v) ? v : (v > sigma_clip ? v : sigma_clip);
    return v;
}

#define ICU_WARPS 4
struct UpdLayout { int C, mean, xc, ecost, perm, total; };
__host__ __device__ inline UpdLayout upd_layout(int nr) {
    UpdLayout U; const int d = nr * nr + 1, ldc = al4(d);
    int q = 0;
    U.C = q; q += d * ldc; U.mean = q; q += ldc;       // contiguous like the chain block: [LT | mean]
    U.xc = q; q += ICF_MAX_NE * ldc; U.ecost = q; q += al4(ICF_MAX_NE); U.perm = q; q += al4(ICF_MAX_NE);
    U.total = q;
    return U;
}
// selection of the elites + distribution update of inner iteration `it`, one warp per chain  [compute_beta.py:51-68]
template <int NR>
__global__ void __launch_bounds__(ICU_WARPS * 32) k_icem_update(DCfg c, SplitArgs sa, int it) {
    constexpr int nm = NR * NR, d = nm + 1;
    extern __shared__ __align__(128) float sm[];
    const SplitLayout L = split_layout(NR, c.S_in, c.n_el_in);
    const UpdLayout U = upd_layout(NR);
    const int ldc = L.ldc, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, S = c.S_in, ne = c.n_el_in;
    const int gl = blockIdx.x * ICU_WARPS + warp;
    if (gl >= sa.n_chains) return;                      // whole warp; no block-wide barrier below
    const int g = sa.g0 + gl;
    float* st = sa.state + (size_t)g * L.total;
    float* W = sm + warp * U.total;
    float* C = W + U.C; float* mean = W + U.mean; float* xc = W + U.xc; float* ecost = W + U.ecost; int* perm = (int*)(W + U.perm);
    const int n_old = it == 0 ? 0 : ne, nrow = S - ne;
    const float* betas = sa.bscratch + (size_t)g * S * (NR + 1); const int* idxs = (const int*)(betas + S * NR);
    float* eth = st + L.eth; float* eb = st + L.eb; int* ei = (int*)(st + L.ei);
    // ---- old elite costs; old factor + mean (the rows evaluated this iteration were drawn from them)
    if (lane < ne) ecost[lane] = it == 0 ? 0.0f : st[L.ecost + lane];
    if (it > 0) {
        const float4* src = reinterpret_cast<const float4*>(st + L.LT); float4* dst = reinterpret_cast<float4*>(C);
#pragma unroll 1
        for (int i = lane; i < (d + 1) * ldc / 4; i += 32) dst[i] = src[i];
    }
    __syncwarp();
    // ---- stable argsort of the S candidate costs, first ne entries  [compute_beta.py:56]
    icf_select(lane, S, n_old, ne, ecost, st + L.cost, perm, ecost);
    __syncwarp();
    // ---- elite rows (rank order), mean, centered rows  [compute_beta.py:56-61]: lane q owns column q
    float v[ICF_MAX_NE]; float mu = 0.0f;
    {
        const int q = lane < d ? lane : d - 1;
        const float* zT = c.zb_iterT + (size_t)(it > 0 ? it - 1 : 0) * d * nrow;
        float s = 0.0f;
#pragma unroll
        for (int el = 0; el < ICF_MAX_NE; el++) {
            if (el < ne) {
                const int p = perm[el];
                float x;
                if (p < n_old) x = eth[p * ldc + q];
                else if (it == 0) x = __ldg(c.theta0 + p * d + q);
                else x = icu_row_elem<d>(C, ldc, mean, zT, nrow, p - n_old, q, c.sigma_clip);
                v[el] = x; s = s + x;
            }
        }
        mu = s / (float)ne;
    }
    // elite beta vectors / packed index words: entry i = (elite i / NR, component i % NR), two per lane (ne * NR <= 64)
    float gb[2] = {0.0f, 0.0f}; int gi = 0;
#pragma unroll
    for (int h = 0; h < 2; h++) {
        const int i = lane + 32 * h;
        if (i < ne * NR) { const int p = perm[i / NR], k = i % NR; gb[h] = p < n_old ? eb[p * NR + k] : betas[(p - n_old) * NR + k]; }
    }
    if (lane < ne) { const int p = perm[lane]; gi = p < n_old ? ei[p] : idxs[p - n_old]; }
    __syncwarp();                                      // every read of the old elites / old factor precedes the writes below
    if (lane < d) {
        mean[lane] = mu;
#pragma unroll
        for (int el = 0; el < ICF_MAX_NE; el++) if (el < ne) { eth[el * ldc + lane] = v[el]; xc[el * ldc + lane] = v[el] - mu; }
    }
#pragma unroll
    for (int h = 0; h < 2; h++) { const int i = lane + 32 * h; if (i < ne * NR) eb[i] = gb[h]; }
    if (lane < ne) { ei[lane] = gi; st[L.ecost + lane] = ecost[lane]; }
    __syncwarp();
    // ---- jnp.cov (ddof = 1) + 0.05 I, lower triangle, four columns per task  [compute_beta.py:61]
    {   // task t -> (row r, column group q4 <= r / 4): rows 4b .. 4b+3 have b + 1 tasks each, 2 b (b + 1) tasks precede block b
#pragma unroll 1
        for (int t = lane; t < 2 * ((d + 3) / 4) * ((d + 3) / 4 + 1); t += 32) {
            int b = 0;
#pragma unroll
            for (int bb = 1; bb < (d + 3) / 4; bb++) if (t >= 2 * bb * (bb + 1)) b = bb;
            const int rem = t - 2 * b * (b + 1), r = 4 * b + rem / (b + 1), q4 = rem % (b + 1);
            if (r < d) icf_cov_task(xc, C, ldc, ne, r, q4);
        }
    }
    __syncwarp();
    // ---- Cholesky, left-looking by panels of four columns, factor transposed in place  [compute_beta.py:63]
    icf_chol_panel<d>(C, ldc, lane);
    {   // [LT | mean] back to the chain block (read by the next evaluation and by the next update's row re-derivation)
        const float4* src = reinterpret_cast<const float4*>(C); float4* dst = reinterpret_cast<float4*>(st + L.LT);
#pragma unroll 1
        for (int i = lane; i < (d + 1) * ldc / 4; i += 32) dst[i] = src[i];
    }
    if (lane == 0) sa.res_beta[(size_t)g * c.iters_in + it] = ecost[0];
    if (it == c.iters_in - 1) {
        // beta / reduced set of the best sample; sigma from the RESAMPLED array [Q7]: candidate index perm[0] applied to the arrays of the NEXT
        // iteration (elites = this iteration's winners, rows = the rows the new factor would produce)
        if (lane < NR) { sa.beta[(size_t)g * NR + lane] = gb[0]; }       // entries 0..NR-1 of the gather = elite 0
        const int gi0 = __shfl_sync(FULL, gi, 0);
        if (lane < NR) sa.ridx[(size_t)g * NR + lane] = (gi0 >> (5 * lane)) & 31;
        const int p0 = perm[0];                                        // warp-uniform
        float sg = 0.0f;
        if (p0 < ne) {                                                  // column nm of the new elite of rank p0 (lane nm holds it)
            float pick = 0.0f;
#pragma unroll
            for (int el = 0; el < ICF_MAX_NE; el++) if (el == p0) pick = v[el];
            sg = __shfl_sync(FULL, pick, nm);
        } else {
            __syncwarp();
            if (lane == 0) sg = icu_row_elem<d>(C, ldc, mean, c.zb_iterT + (size_t)it * d * nrow, nrow, p0 - ne, nm, c.sigma_clip);
        }
        if (lane == 0) sa.sigma[g] = sg;
    }
}
