// k_inner_cem_warp.cuh -- throughput form of the reduced-set inner CEM (mmd_opt, num_reduced <= 5): ONE WARP PER CHAIN.
//
// k_inner_cem_fast (one 3-warp CTA per chain) spends a third of its warp-time at __syncthreads while one warp runs the serial
// stages (elite selection, Cholesky) -- profiles/r01_v5_summary.md: 4.0 barrier-stalled warps per issued instruction, 53 % issue.
// Here a chain lives entirely in one warp: no block barriers, the serial stages of one chain overlap the parallel stages of the
// ~20 other chains resident on the SM, and the chain's shared-memory state shrinks to 10.7 KB because the resampled rows are
// never stored in shared memory: a lane draws its row (multivariate normal, packed FP32), evaluates it at once from registers
// and parks theta / beta / indices in a 128-byte row of an L2-resident stash (one stash per persistent CTA, 36 MB in total)
// from which only the 11 new elites are read back.  CTAs are persistent (grid = SMs x resident warps) and stride over the chains.
// Arithmetic is identical, operation for operation, to k_inner_cem_fast / k_inner_cem / oracle_inner_cem.
#pragma once
#include "k_inner_cem.cuh"

#define ICW_STASH_LD 32        // floats per stashed row: theta [0,d), beta [26,26+NR), packed indices [31]
#define ICW_BETA_OFF 26

struct WarpLayout { int D, cost, eth, ecost, ebetas, eidxs, perm, xc, C, mean, small, total; int lde, ldc; };
__host__ __device__ inline WarpLayout warp_layout(int nr, int S, int ne) {
    WarpLayout L; const int nm = nr * nr, d = nm + 1;
    L.ldc = al4(d); L.lde = al4(d);
    int q = 0;
    L.D = q; q += al4(nm * nm);
    L.cost = q; q += al4(S);
    L.eth = q; q += 2 * al4(ne * L.lde); L.ecost = q; q += 2 * al4(ne); L.ebetas = q; q += 2 * al4(ne * nr); L.eidxs = q; q += 2 * al4(ne);
    L.perm = q; q += al4(ne);
    L.xc = q; q += ICF_MAX_NE * L.ldc;       // centered elites; with C it also stages the mother features while D is built
    L.C = q; q += d * L.ldc;
    L.mean = q; q += L.ldc;
    L.small = q; q += 64;
    L.total = q;
    return L;
}

template <int NR>
__global__ void __launch_bounds__(32, 20) k_inner_cem_warp(DCfg c, RollArgs ra) {
    extern __shared__ __align__(128) float sm[];
    const RiskArgs& a = ra.r;
    constexpr int nm = NR * NR, d = nm + 1, NG = (d + 3) / 4, NPAIR = (d + 1) / 2;
    static_assert(d <= ICW_BETA_OFF && ICW_BETA_OFF + NR <= ICW_STASH_LD - 1, "stash row layout");
    const int lane = threadIdx.x;
    const int np = c.np, S = c.S_in, ne = c.n_el_in, nrow = S - ne;
    const WarpLayout L = warp_layout(NR, S, ne);
    const int ldc = L.ldc, lde = L.lde;
    float* D = sm + L.D; float* cost = sm + L.cost; int* perm = (int*)(sm + L.perm); float* xc = sm + L.xc; float* C = sm + L.C; float* LT = C;
    float* mean = sm + L.mean; float* small = sm + L.small;
    const int eth_sz = al4(ne * lde), ecost_sz = al4(ne), eb_sz = al4(ne * NR), ei_sz = al4(ne);
    float* stash = ra.stash + (size_t)blockIdx.x * S * ICW_STASH_LD;
    // covariance tasks of this lane: task t = lane + 32u -> (row, column group), lower triangle in groups of four columns
    int ctask[4] = {-1, -1, -1, -1};
    {
        int t = 0;
#pragma unroll 1
        for (int r = 0; r < d; r++) {
#pragma unroll 1
            for (int q4 = 0; q4 <= r / 4; q4++) {
#pragma unroll
                for (int u = 0; u < 4; u++) if (t == lane + 32 * u) ctask[u] = r | (q4 << 8);
                t++;
            }
        }
    }
#pragma unroll 1
    for (int g = blockIdx.x; g < a.n_samples; g += gridDim.x) {
        const int e = g / a.B;
        // ---- distance table of the mother features  [kernel_computation.py:31-33]; the features are staged in the xc / C region
        {
            float* F = xc;
            const float* Fg = ra.feat + (size_t)g * nm * 2 * NV;
#pragma unroll 1
            for (int i = lane; i < nm * 2 * NV; i += 32) F[i] = Fg[i];
            __syncwarp();
#pragma unroll 1
            for (int i = lane; i < nm * nm; i += 32) {
                const float* Fa = F + (i / nm) * 2 * NV; const float* Fb = F + (i % nm) * 2 * NV;
                float dist = 0.0f;
#pragma unroll 2
                for (int f = 0; f < 2 * NV; f++) dist = dist + fabsf(Fa[f] - Fb[f]);
                D[i] = dist;
            }
            __syncwarp();
#pragma unroll 1
            for (int i = lane; i < ICF_MAX_NE * ldc; i += 32) xc[i] = 0.0f;            // pad columns / rows stay zero
            __syncwarp();
        }
        float* resb = a.res_beta + (size_t)g * c.iters_in;
#pragma unroll 1
        for (int it = 0; it < c.iters_in; it++) {
            const int cur = it & 1, nxt = cur ^ 1;
            const int n_old = it == 0 ? 0 : ne, n_new = S - n_old;
            float* eth_c = sm + L.eth + cur * eth_sz; float* eth_n = sm + L.eth + nxt * eth_sz;
            float* ecost_c = sm + L.ecost + cur * ecost_sz; float* ecost_n = sm + L.ecost + nxt * ecost_sz;
            float* eb_c = sm + L.ebetas + cur * eb_sz; float* eb_n = sm + L.ebetas + nxt * eb_sz;
            int* ei_c = (int*)(sm + L.eidxs) + cur * ei_sz; int* ei_n = (int*)(sm + L.eidxs) + nxt * ei_sz;
            // ---- draw (iteration 0: the constant theta0 table; later: mean + L z of the previous elites, z table it-1) and evaluate
            //      the new rows, 32 at a time, straight from registers  [compute_beta.py:63-66, 113-129]
            const float* zT = c.zb_iterT + (size_t)(it > 0 ? it - 1 : 0) * d * nrow;
#pragma unroll 1
            for (int base = 0; base < n_new; base += 32) {
                const bool act = base + lane < n_new;
                const int row = act ? base + lane : n_new - 1;
                pk::f2 acc[2 * NG];
                if (it == 0) {
#pragma unroll
                    for (int p = 0; p < 2 * NG; p++) {
                        const float x0 = 2 * p < d ? __ldg(c.theta0T + (2 * p) * S + row) : 0.0f;
                        const float x1 = 2 * p + 1 < d ? __ldg(c.theta0T + (2 * p + 1) * S + row) : 0.0f;
                        acc[p] = pk::pack(x0, x1);
                    }
                } else {
                    icf_mvn_row<d>(LT, ldc, zT, nrow, row, acc);
#pragma unroll
                    for (int p = 0; p < NPAIR; p++) {
                        float x0, x1; pk::unpack(acc[p], x0, x1);
                        x0 = mean[2 * p] + x0;
                        if (2 * p == nm) x0 = (x0 != x0) ? x0 : (x0 > c.sigma_clip ? x0 : c.sigma_clip);
                        if (2 * p + 1 < d) {
                            x1 = mean[2 * p + 1] + x1;
                            if (2 * p + 1 == nm) x1 = (x1 != x1) ? x1 : (x1 > c.sigma_clip ? x1 : c.sigma_clip);
                        }
                        acc[p] = pk::pack(x0, x1);
                    }
                }
                // park theta in the stash row and feed the packed top-NR selection (see beta_sample_fast)
                float* srow = stash + (size_t)row * ICW_STASH_LD;
                int tk[NR + 1];
#pragma unroll
                for (int p = 0; p <= NR; p++) tk[p] = 0;
                float sigma = 0.0f;
#pragma unroll
                for (int g4 = 0; g4 < NG; g4++) {
                    float4 v; pk::unpack(acc[2 * g4], v.x, v.y); pk::unpack(acc[2 * g4 + 1], v.z, v.w);
                    if (act) *reinterpret_cast<float4*>(srow + 4 * g4) = v;
                    const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        const int m = 4 * g4 + u;
                        if (m < nm) {
                            const int key = (int)((dm::f2u(vv[u]) & 0x7fffffe0u) | (uint32_t)m);
                            tk[0] = max(tk[0], key);
#pragma unroll
                            for (int p = 0; p < NR; p++) { const int lo = min(tk[p], tk[p + 1]), hi = max(tk[p], tk[p + 1]); tk[p] = lo; tk[p + 1] = hi; }
                        } else if (m == nm) sigma = vv[u];
                    }
                }
                int ti[NR];
                bool near = false;
#pragma unroll
                for (int p = 0; p < NR; p++) { ti[p] = tk[p + 1] & 31; near |= ((tk[p] ^ tk[p + 1]) < 32); }
                if (near) {                                   // exact path on the row this lane just parked (own writes are visible to itself)
                    const int pkd = top_abs_exact<NR>(srow);
#pragma unroll
                    for (int p = 0; p < NR; p++) ti[p] = (pkd >> (5 * p)) & 31;
                }
                float bt[NR]; int pidx;
                const float cst = beta_eval<NR>(c, ti, sigma, D, bt, &pidx);
                if (act) {
                    cost[row] = cst;
#pragma unroll
                    for (int i = 0; i < NR; i++) srow[ICW_BETA_OFF + i] = bt[i];
                    ((int*)srow)[ICW_STASH_LD - 1] = pidx;
                }
            }
            __threadfence_block();                            // the stash rows are read back by other lanes of this warp
            __syncwarp();
            // ---- stable argsort, first ne entries  [compute_beta.py:56]
            icf_select(lane, S, n_old, ne, ecost_c, cost, perm, ecost_n);
            __syncwarp();
            // ---- gather the elites (rank order) from the previous elites / the stash, their mean and the centered rows  [compute_beta.py:56-61]
            if (lane < d) {
                float s = 0.0f;
                float v[ICF_MAX_NE];
#pragma unroll
                for (int el = 0; el < ICF_MAX_NE; el++) {
                    v[el] = 0.0f;
                    if (el < ne) { const int p = perm[el]; v[el] = p < n_old ? eth_c[p * lde + lane] : __ldcg(stash + (size_t)(p - n_old) * ICW_STASH_LD + lane); }
                }
#pragma unroll
                for (int el = 0; el < ICF_MAX_NE; el++) if (el < ne) { eth_n[el * lde + lane] = v[el]; s = s + v[el]; }
                const float mu = s / (float)ne;
                mean[lane] = mu;
#pragma unroll
                for (int el = 0; el < ICF_MAX_NE; el++) if (el < ne) xc[el * ldc + lane] = v[el] - mu;
            }
#pragma unroll 1
            for (int i = lane; i < ne * NR; i += 32) {
                const int p = perm[i / NR], k = i % NR;
                eb_n[i] = p < n_old ? eb_c[p * NR + k] : __ldcg(stash + (size_t)(p - n_old) * ICW_STASH_LD + ICW_BETA_OFF + k);
            }
            if (lane < ne) { const int p = perm[lane]; ei_n[lane] = p < n_old ? ei_c[p] : __ldcg((const int*)stash + (size_t)(p - n_old) * ICW_STASH_LD + ICW_STASH_LD - 1); }
            __syncwarp();
            // ---- jnp.cov (ddof = 1) + 0.05 I  [compute_beta.py:61], then its Cholesky factor
#pragma unroll 1
            for (int u = 0; u < 4; u++) {                     // rolled: one copy of the task body in the instruction cache
                const int tsk = u == 0 ? ctask[0] : u == 1 ? ctask[1] : u == 2 ? ctask[2] : ctask[3];
                if (tsk >= 0) icf_cov_task(xc, C, ldc, ne, tsk & 0xff, tsk >> 8);
            }
            __syncwarp();
            icf_chol<d>(C, ldc, lane);
            if (lane == 0) resb[it] = ecost_n[0];
        }
        // ---- outputs of the chain: beta / reduced set of the best sample of the last iteration; sigma is read from the array RESAMPLED after
        //      that iteration [Q7]: an elite row, or one element of the (otherwise unused) last multivariate-normal draw
        const int last = ((c.iters_in - 1) & 1) ^ 1;
        const float* eth_l = sm + L.eth + last * eth_sz; const float* eb_l = sm + L.ebetas + last * eb_sz; const int* ei_l = (const int*)(sm + L.eidxs) + last * ei_sz;
        if (lane == 0) {
            for (int i = 0; i < NR; i++) { small[i] = eb_l[i]; ((int*)small)[16 + i] = (ei_l[0] >> (5 * i)) & 31; }
            const int p0 = perm[0];
            float sg;
            if (p0 < ne) sg = eth_l[p0 * lde + nm];
            else {
                const float* zl = c.zb_iterT + (size_t)(c.iters_in - 1) * d * nrow;
                float accv = 0.0f;
                for (int k = 0; k <= nm; k++) accv = fmaf(LT[k * ldc + nm], __ldg(zl + k * nrow + (p0 - ne)), accv);
                sg = mean[nm] + accv;
                sg = (sg != sg) ? sg : (sg > c.sigma_clip ? sg : c.sigma_clip);
            }
            small[48] = sg;
        }
        __syncwarp();
        // ---- risk of the chosen reduced set (its rollouts come back from global memory)  [costs.py:173-186, 121-135]
        const int* ridx = (const int*)small + 16;
        const float* xg = ra.xroll + (size_t)g * nm * np; const float* yg = ra.yroll + (size_t)g * nm * np;
        const float* xo = a.x_obs + (size_t)e * c.O * T_; const float* yo = a.y_obs + (size_t)e * c.O * T_;
        float cs[NR], lbv[NR], ubv[NR];
#pragma unroll 1
        for (int r = 0; r < NR; r++) {
            const float* xred = xg + ridx[r] * np; const float* yred = yg + ridx[r] * np;
            float m = 0.0f, l = 0.0f, u = 0.0f;
            for (int i = lane; i < c.O * np; i += 32) {
                const int o = i / np, t = i % np;
                m = dm::nmax_(m, fbar(c, xred[t], yred[t], xo[o * T_ + t], yo[o * T_ + t]));
            }
            for (int t = lane; t < np; t += 32) {
                l = dm::nmax_(l, dm::max0_(-yred[t] + c.y_lb));
                u = dm::nmax_(u, dm::max0_(yred[t] - c.y_ub));
            }
            m = warp_nmax(m); l = warp_nmax(l); u = warp_nmax(u);
            if (lane == 0) { small[24 + r] = m; small[32 + r] = l; small[40 + r] = u; }
        }
        __syncwarp();
        if (lane == 0) {
            float beta[NR];
            for (int r = 0; r < NR; r++) { cs[r] = small[24 + r]; lbv[r] = small[32 + r]; ubv[r] = small[40 + r]; beta[r] = small[r]; a.beta[(size_t)g * NR + r] = small[r]; }
            const float sigma = small[48];
            a.sigma[g] = sigma;
            a.risk[g] = mmd_cost(c, beta, cs, sigma);
            a.lane[g] = mmd_cost(c, beta, lbv, sigma) + mmd_cost(c, beta, ubv, sigma);
        }
        __syncwarp();
    }
}
