// k_inner_pipe.cuh -- the reduced-set inner CEM of mmd_opt (num_reduced <= 5) as a WARP-SPECIALISED two-stage pipeline: one CTA = two chains,
// three EVALUATION warps + one SERIAL warp.  While the serial warp runs the one-warp phases of chain A (stable elite selection, elite gather,
// covariance, panel Cholesky), the evaluation warps resample + evaluate the 89 new beta samples of chain B; then the roles swap chains.  The two
// stages take about the same time (7 us each when alone on an SM), so no warp waits for another.
//
// Why (ncu of the one-CTA-per-chain kernel k_inner_cem_fast, profiles/r01_v12_summary.md + source view of v15): 37 % of all warp samples were
// barrier stalls -- two of its three warps parked behind the one-warp Cholesky (22 %) and selection -- so only ~23 of the 36 resident warps
// of an SM were ever runnable and issue-active stalled at 65 %.  The phase-split alternative (k_inner_split.cuh: evaluation and update as separate
// kernels, state through L2) removes the idle warps as well but pays for it with launches and a row re-derivation; measured slower.
//
// Also new here: the resampled rows are never stored in shared memory.  A thread draws its row into registers (mean + L z, k ascending),
// evaluates it, and writes it to an L2 scratch only if its cost beats the worst surviving elite -- any other row cannot be selected
// (stable sort: the 11 old elites precede it).  Chain state shrinks from 16.9 KB to 8.9 KB, so eight or nine 4-warp CTAs = 16-18 chains are in flight per SM.
//
// Arithmetic: the same contract, element for element, as k_inner_cem_fast (tests hold every variant to the oracle bit for bit).
// Replaces beta_cem.compute_cem (S/compute_beta.py:93-157) + kernel_matrix.compute_kernel (S/kernel_computation.py:19-65).
#pragma once
#include "k_inner_split.cuh"

#define ICP_THREADS 128
#define ICP_EVAL_THREADS 96
#define ICP_TH_LD 96          // row stride of the L2 row scratch thT[chain][column][row]

struct PipeLayout {            // per-chain shared-memory state in floats; offsets are multiples of 4 floats
    int D, LT, mean, eth, xc, ecost, perm, eb, ei, cost, small, chain, ldc;
};
__host__ __device__ inline PipeLayout pipe_layout(int nr, int S, int ne) {
    PipeLayout L; const int nm = nr * nr, d = nm + 1;
    L.ldc = al4(d);
    int q = 0;
    L.D = q; q += al4(nm * nm);
    L.LT = q; q += d * L.ldc; L.mean = q; q += L.ldc;
    L.eth = q; q += ne * L.ldc;
    L.xc = q; q += ne * L.ldc;
    L.ecost = q; q += al4(ICF_MAX_NE); L.perm = q; q += al4(ICF_MAX_NE); L.eb = q; q += al4(ICF_MAX_NE * nr); L.ei = q; q += al4(ICF_MAX_NE);
    L.cost = q; q += al4(S);
    L.small = q;
    L.chain = q;
    return L;
}
// + ONE row buffer per CTA (the evaluation warps work on one chain at a time): S - ne rows of odd stride (conflict-free row-per-thread access)
__host__ __device__ inline size_t pipe_smem_bytes(int nr, int S, int ne) {
    return (2 * (size_t)pipe_layout(nr, S, ne).chain + (size_t)al4((S - ne) * ((nr * nr + 1) | 1))) * sizeof(float);
}

// barrier ids are IMMEDIATES: with register operands ptxas reserves all 16 hardware barriers per CTA and the SM's pool of 64 caps it at 4 CTAs
// (measured: launch__occupancy_limit_barriers = 4, profiles/r02_pipe_v1.md)
template <int ID> __device__ __forceinline__ void icp_bar_sync() { asm volatile("bar.sync %0, %1;" :: "n"(ID), "n"(ICP_THREADS) : "memory"); }
template <int ID> __device__ __forceinline__ void icp_bar_arrive() { asm volatile("bar.arrive %0, %1;" :: "n"(ID), "n"(ICP_THREADS) : "memory"); }


// ---- evaluation stage of chain g, inner iteration it: thread s < n_new draws row s (mean + L z, k ascending) into the CTA's row buffer and
//      evaluates it  [compute_beta.py:63-66, 113-129, 70-91]
template <int NR>
__device__ __forceinline__ void icp_eval(const DCfg& c, const RollArgs& ra, float* __restrict__ thT, int g, int it, int tid, float* W, float* th, const PipeLayout& L) {
    constexpr int nm = NR * NR, d = nm + 1, NG = (d + 3) / 4, NPAIR = (d + 1) / 2, ldt = d | 1;
    const int S = c.S_in, ne = c.n_el_in, ldc = L.ldc;
    const float* D = W + L.D; const float* LT = W + L.LT; const float* mean = W + L.mean; float* cost = W + L.cost;
    float* betas = ra.bscratch + (size_t)g * S * (NR + 1); int* idxs = (int*)(betas + S * NR);
    const int n_new = it == 0 ? S : S - ne;
    // a new row can enter the elite set only if it sorts strictly before the worst old elite (stable argsort: the ne old elites come first on ties)
    const uint32_t gate = it == 0 ? 0u : sort_key32(W[L.ecost + ne - 1]);
#pragma unroll 1
    for (int s = tid; s < n_new; s += ICP_EVAL_THREADS) {
        const float* row;
        if (it == 0) row = c.theta0 + s * d;             // iteration 0 evaluates the constant table in place
        else {
            const int nrow = S - ne;
            pk::f2 acc[2 * NG];
            icf_mvn_row_unrolled<d>(LT, ldc, c.zb_iterT + (size_t)(it - 1) * d * nrow, nrow, s, acc);
            float* dst = th + s * ldt;
#pragma unroll
            for (int p = 0; p < NPAIR; p++) {
                float x0, x1; pk::unpack(acc[p], x0, x1);
                float v0 = mean[2 * p] + x0;
                if (2 * p == nm) v0 = (v0 != v0) ? v0 : (v0 > c.sigma_clip ? v0 : c.sigma_clip);
                dst[2 * p] = v0;
                if (2 * p + 1 < d) {
                    float v1 = mean[2 * p + 1] + x1;
                    if (2 * p + 1 == nm) v1 = (v1 != v1) ? v1 : (v1 > c.sigma_clip ? v1 : c.sigma_clip);
                    dst[2 * p + 1] = v1;
                }
            }
            row = dst;
        }
        float bt[NR]; int pidx;
        const float cs = beta_sample_fast<NR>(c, row, D, bt, &pidx);
        cost[s] = cs;
        if (it == 0 || sort_key32(cs) < gate) {          // candidate elite: publish its beta vector / reduced set (and the row itself when it is not a table row)
#pragma unroll
            for (int i = 0; i < NR; i++) betas[s * NR + i] = bt[i];
            idxs[s] = pidx;
            if (it > 0) {
                float* dst = thT + (size_t)g * d * ICP_TH_LD + s;
#pragma unroll
                for (int k = 0; k < d; k++) dst[k * ICP_TH_LD] = row[k];
            }
        }
    }
}

// ---- serial stage of chain g, inner iteration it, by ONE warp: selection, elite gather, mean, covariance, Cholesky  [compute_beta.py:51-68]
template <int NR>
__device__ __forceinline__ void icp_serial(const DCfg& c, const RollArgs& ra, const float* __restrict__ thT, int g, int it, int lane, float* W, const PipeLayout& L) {
    constexpr int nm = NR * NR, d = nm + 1;
    const RiskArgs& a = ra.r;
    const int S = c.S_in, ne = c.n_el_in, ldc = L.ldc;
    float* C = W + L.LT; float* mean = W + L.mean; float* eth = W + L.eth; float* xc = W + L.xc; float* ecost = W + L.ecost;
    int* perm = (int*)(W + L.perm); float* eb = W + L.eb; int* ei = (int*)(W + L.ei); const float* cost = W + L.cost;
    const int n_old = it == 0 ? 0 : ne, nrow = S - ne;
    const float* betas = ra.bscratch + (size_t)g * S * (NR + 1); const int* idxs = (const int*)(betas + S * NR);
    icf_select(lane, S, n_old, ne, ecost, cost, perm, ecost);
    __syncwarp();
    float v[ICF_MAX_NE]; float mu = 0.0f;
    {
        const int q = lane < d ? lane : d - 1;
        const float* rows = thT + ((size_t)g * d + q) * ICP_TH_LD;
        float s = 0.0f;
#pragma unroll
        for (int el = 0; el < ICF_MAX_NE; el++) {
            if (el < ne) {
                const int p = perm[el];
                const float x = p < n_old ? eth[p * ldc + q] : (it == 0 ? __ldg(c.theta0 + p * d + q) : __ldcg(rows + (p - n_old)));
                v[el] = x; s = s + x;
            }
        }
        mu = s / (float)ne;
    }
    float gb[2] = {0.0f, 0.0f}; int gi = 0;
#pragma unroll
    for (int h = 0; h < 2; h++) {
        const int i = lane + 32 * h;
        if (i < ne * NR) { const int p = perm[i / NR], k = i % NR; gb[h] = p < n_old ? eb[p * NR + k] : __ldcg(betas + (p - n_old) * NR + k); }
    }
    if (lane < ne) { const int p = perm[lane]; gi = p < n_old ? ei[p] : __ldcg(idxs + (p - n_old)); }
    __syncwarp();                                      // every read of the old elites precedes the writes below
    if (lane < d) {
        mean[lane] = mu;
#pragma unroll
        for (int el = 0; el < ICF_MAX_NE; el++) if (el < ne) { eth[el * ldc + lane] = v[el]; xc[el * ldc + lane] = v[el] - mu; }
    }
#pragma unroll
    for (int h = 0; h < 2; h++) { const int i = lane + 32 * h; if (i < ne * NR) eb[i] = gb[h]; }
    if (lane < ne) ei[lane] = gi;
    __syncwarp();
    {   // jnp.cov (ddof = 1) + 0.05 I: task t -> (row r, column group q4 <= r / 4); rows 4b .. 4b+3 have b + 1 tasks each, 2 b (b + 1) tasks precede block b
        constexpr int NB = (d + 3) / 4;
#pragma unroll 1
        for (int t = lane; t < 2 * NB * (NB + 1); t += 32) {
            int b = 0;
#pragma unroll
            for (int bb = 1; bb < NB; bb++) if (t >= 2 * bb * (bb + 1)) b = bb;
            const int rem = t - 2 * b * (b + 1), r = 4 * b + rem / (b + 1), q4 = rem % (b + 1);
            if (r < d) icf_cov_task(xc, C, ldc, ne, r, q4);
        }
    }
    __syncwarp();
    icf_chol_panel<d>(C, ldc, lane);
    if (lane == 0) a.res_beta[(size_t)g * c.iters_in + it] = ecost[0];
    if (it == c.iters_in - 1) {
        // beta / reduced set of the best sample; sigma from the RESAMPLED array [Q7]: candidate index perm[0] applied to the arrays of the NEXT
        // iteration (elites = this iteration's winners, rows = what the new factor would produce)
        if (lane < NR) a.beta[(size_t)g * NR + lane] = gb[0];            // gather entries 0 .. NR-1 = elite 0
        const int gi0 = __shfl_sync(FULL, gi, 0);
        if (lane < NR) ra.ridx[(size_t)g * NR + lane] = (gi0 >> (5 * lane)) & 31;
        const int p0 = perm[0];                                          // warp-uniform
        float sg = 0.0f;
        if (p0 < ne) {
            float pick = 0.0f;
#pragma unroll
            for (int el = 0; el < ICF_MAX_NE; el++) if (el == p0) pick = v[el];
            sg = __shfl_sync(FULL, pick, nm);
        } else if (lane == 0) {
            sg = icu_row_elem<d>(C, ldc, mean, c.zb_iterT + (size_t)it * d * nrow, nrow, p0 - ne, nm, c.sigma_clip);
        }
        if (lane == 0) a.sigma[g] = sg;
    }
}

template <int NR, int MINB>
__global__ void __launch_bounds__(ICP_THREADS, MINB) k_inner_cem_pipe(DCfg c, RollArgs ra, float* __restrict__ thT) {
    extern __shared__ __align__(128) float sm[];
    constexpr int nm = NR * NR;
    const RiskArgs& a = ra.r;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g0 = 2 * blockIdx.x;
    if (g0 >= a.n_samples) return;
    const int nch = (g0 + 1 < a.n_samples) ? 2 : 1;
    const PipeLayout L = pipe_layout(NR, c.S_in, c.n_el_in);
    // ---- distance tables of both chains [kernel_computation.py:31-33]; the mother features pass through the row buffer, one chain at a time
    float* th = sm + 2 * L.chain;
#pragma unroll 1
    for (int x = 0; x < nch; x++) {
        const float* Fg = ra.feat + (size_t)(g0 + x) * nm * 2 * NV;
#pragma unroll 1
        for (int i = tid; i < nm * 2 * NV; i += ICP_THREADS) th[i] = Fg[i];
        __syncthreads();
        float* D = sm + x * L.chain + L.D;
#pragma unroll 1
        for (int i = tid; i < nm * nm; i += ICP_THREADS) {
            const float* Fa = th + (i / nm) * 2 * NV; const float* Fb = th + (i % nm) * 2 * NV;
            float dist = 0.0f;
#pragma unroll 2
            for (int f = 0; f < 2 * NV; f++) dist = dist + fabsf(Fa[f] - Fb[f]);
            D[i] = dist;
        }
        __syncthreads();
    }
    // ---- the pipeline.  Named barriers (128 participants each): 1 + x = "evaluation of chain x finished" (evaluation warps arrive, serial warp waits),
    //      3 + x = "distribution of chain x updated" (serial warp arrives, evaluation warps wait)
    const int iters = c.iters_in;
    if (warp < 3) {
#pragma unroll 1
        for (int it = 0; it < iters; it++) {
#pragma unroll 1
            for (int x = 0; x < nch; x++) {          // x is CTA-uniform; ONE copy of the stage code serves both chains (instruction cache)
                if (it > 0) { if (x == 0) icp_bar_sync<3>(); else icp_bar_sync<4>(); }
                icp_eval<NR>(c, ra, thT, g0 + x, it, tid, sm + x * L.chain, th, L);
                __threadfence_block();
                __syncwarp();                      // aligned barrier instructions: the warp is converged again after the row loop
                if (x == 0) icp_bar_arrive<1>(); else icp_bar_arrive<2>();
            }
        }
    } else {
#pragma unroll 1
        for (int it = 0; it < iters; it++) {
#pragma unroll 1
            for (int x = 0; x < nch; x++) {
                if (x == 0) icp_bar_sync<1>(); else icp_bar_sync<2>();
                icp_serial<NR>(c, ra, thT, g0 + x, it, lane, sm + x * L.chain, L);
                __threadfence_block();
                __syncwarp();
                if (it + 1 < iters) { if (x == 0) icp_bar_arrive<3>(); else icp_bar_arrive<4>(); }
            }
        }
    }
}
