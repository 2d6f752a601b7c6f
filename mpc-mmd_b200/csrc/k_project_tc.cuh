// k_project_tc.cuh -- tensor-core variant of the projection kernel (opt-in: MPCMMD_PROJ=tc).
//
// Same stage as k_project (Helper.compute_x_guess, Projection.compute_projection, Helper.compute_controls and the
// risk-independent cost terms; reference S/optimizer/cem_helper.py:169-230, :540-551, :232-262, S/optimizer/projection.py:276-323),
// but the matrix products run on the 5th-generation tensor cores:
//   * a CTA owns 128 samples; sample r is TMEM lane r, so `tcgen05.ld 32x32b` hands thread r its own row of every product;
//   * "expansion" products  [128 samples x 16 coefficients] x [16 x 32 knots]  (guess derivatives, projected trajectories):
//     A = the coefficients, written K-major into shared memory by the owning thread, B = the Bernstein images P / Pdot / Pddot;
//   * "reduction" products  [128 samples x 8 knots] x [8 x 16 coefficients]  (P^T r for the multiplier and linear-cost updates):
//     A = residuals / clipped targets of 8 knots, B = a transposed copy of the image (an MN-major descriptor on the first copy
//     returned zeros for kind::tf32 with SWIZZLE_NONE on B200, measured; both operands are therefore K-major);
//   * every product is 4 x kind::tf32 MMAs on an error-free round-to-nearest split x = hi + lo of both operands
//     (lo*lo, lo*hi, hi*lo, hi*hi, smallest first; fp32 accumulation in TMEM), i.e. ~2^-22 relative per product.
// The element-wise work between the products (polar clip / lane slacks / controls / cost norms) stays on the CUDA cores with one
// thread per sample; finite differences are sequential carries in registers.  The polar step needs no trigonometry here: the
// reference's atan2 -> unwrap -> cos / sin chain is the unit vector (x, y) / r (see polar_fast), which a tolerance-parity kernel may use.
// Roles: warps 0-3 compute (thread = sample); one lane of warp 4 issues every MMA.  A step = the compute threads write their operand
// rows, `fence.proxy.async`, arrive on bar_f (128 arrivals); the issuer waits for bar_f, issues the step's MMAs and `tcgen05.commit`s to bar_m;
// the compute threads wait for bar_m before they overwrite the operand buffer or read the accumulators.  112 KB of shared memory (80 KB of
// constant images, one `cp.async.bulk`) and 256 TMEM columns per CTA: two CTAs per SM.
// This variant does NOT reproduce the ascending fma chains of the arithmetic contract: it is tested against the oracle at the
// tolerance north_star states (1e-4), not bit for bit, and is therefore not the default path.
#pragma once
#include "k_project.cuh"

namespace ptc {
constexpr int THREADS = 128;                               // compute threads (one per sample); warp 4 issues the MMAs
constexpr int CTA_THREADS = 160;
constexpr int NROW = 104;                                  // knots padded to 13 x 8
constexpr uint32_t KSTR = NROW * 16;                       // 1664: bytes between the 16-byte coefficient chunks of an image
constexpr uint32_t IMG = 4 * KSTR;                         // 6656: one [4 chunks][104 knots][4 coefficients] tf32 image
constexpr uint32_t RSTR = 16 * 16;                         // 256: bytes between the 4-knot chunks of a transposed image
constexpr uint32_t RIMG = 26 * RSTR;                       // 6656: one [26 chunks][16 coefficients][4 knots] tf32 image (knots padded to 104)
constexpr uint32_t OFF_R = 6 * IMG;                        // P_hi P_lo Pd_hi Pd_lo Pdd_hi Pdd_lo, then the same six transposed
constexpr uint32_t B_BYTES = 6 * IMG + 6 * RIMG;
constexpr int SMALL = 77 + 88 + 154 + 165 + 36;            // Gx Gy Kx Ky | fp32 rows P[0] Pd[0] Pdd[0] (+3 pad)
constexpr uint32_t CONST_BYTES = B_BYTES + SMALL * 4;      // 81952
constexpr uint32_t MAT = 4096;                             // one [2 chunks][128 rows][4] part of an A operand (8 knots)
constexpr uint32_t ABUF = 8 * MAT;                         // 4 matrices x (hi, lo); also holds the coefficient operand (4 x 8 KB)
constexpr uint32_t OFF_A = (CONST_BYTES + 127) / 128 * 128;
constexpr uint32_t SMEM_BYTES = OFF_A + ABUF + 64;
constexpr uint32_t TMEM_COLS = 256;
constexpr uint32_t COL_UW = 160;                           // columns 0..159: the five products (xd, yd, xdd, ydd, y) of one 32-knot block; 160..223: the four 16-column reductions Ux, Wx, Uy, Wy
// knot blocks of the expansion products: 32, 32, 32 and a 16-knot block at 88 whose upper half (knots 96..103) is the one consumed
// (an M = 128 MMA needs N % 16 == 0 and the images end at knot 103)
static_assert(CONST_BYTES % 16 == 0, "bulk copy size");

__host__ __device__ constexpr uint32_t idesc(int n, int b_mn) {      // kind::tf32, fp32 accumulate, M = 128, A K-major
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)b_mn << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ uint64_t sdesc(uint32_t addr, uint32_t lbo, uint32_t sbo) {   // SWIZZLE_NONE shared-memory matrix descriptor
    return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ void mma(uint32_t d, uint64_t ad, uint64_t bd, uint32_t id, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 :: "r"(d), "l"(ad), "l"(bd), "r"(id), "r"(accumulate) : "memory");
}
// D (+)= (A_hi + A_lo) (B_hi + B_lo), one K = 8 step
__device__ __forceinline__ void mma4(uint32_t d, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi, uint32_t b_lo, uint32_t b_lbo, uint32_t b_sbo,
                                     uint32_t id, bool first) {
    const uint64_t ah = sdesc(a_hi, 2048, 128), al = sdesc(a_lo, 2048, 128);
    const uint64_t bh = sdesc(b_hi, b_lbo, b_sbo), bl = sdesc(b_lo, b_lbo, b_sbo);
    mma(d, al, bl, id, first ? 0u : 1u);
    mma(d, al, bh, id, 1u);
    mma(d, ah, bl, id, 1u);
    mma(d, ah, bh, id, 1u);
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void commit(unsigned long long* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tld8(uint32_t taddr, float (&v)[8]) {
    uint32_t r[8];
    __syncwarp();                                                     // .sync.aligned: the warp must be converged (thread 0 issues MMAs on its own)
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr));
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// error-free split into two tf32 values (round to nearest): x = hi + lo up to 2^-22 |x|
__device__ __forceinline__ void split(float x, float& hi, float& lo) {
    uint32_t h, l;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(x));
    hi = __uint_as_float(h);
    const float d = x - hi;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(l) : "f"(d));
    lo = __uint_as_float(l);
}
// 8 knots of one A matrix: slot m of the buffer, (hi, lo) parts, chunk q at q*2048 + row*16
__device__ __forceinline__ void put8(unsigned char* abuf, int m, int row, const float (&v)[8]) {
    float h[8], l[8];
#pragma unroll
    for (int i = 0; i < 8; i++) split(v[i], h[i], l[i]);
    unsigned char* p = abuf + (2 * m) * MAT + row * 16;
    *reinterpret_cast<float4*>(p) = make_float4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<float4*>(p + 2048) = make_float4(h[4], h[5], h[6], h[7]);
    *reinterpret_cast<float4*>(p + MAT) = make_float4(l[0], l[1], l[2], l[3]);
    *reinterpret_cast<float4*>(p + MAT + 2048) = make_float4(l[4], l[5], l[6], l[7]);
}
// the 2 x 11 coefficients of a sample as the K = 16 operand of the expansion products: x_hi | x_lo | y_hi | y_lo, 8 KB each
__device__ __forceinline__ void put_coef(unsigned char* abuf, int row, const float (&cf)[22]) {
#pragma unroll
    for (int ax = 0; ax < 2; ax++) {
#pragma unroll
        for (int q = 0; q < 4; q++) {
            float h[4], l[4];
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int j = 4 * q + i;
                if (j < NV) split(cf[ax * NV + j], h[i], l[i]); else { h[i] = 0.0f; l[i] = 0.0f; }
            }
            unsigned char* p = abuf + ax * 16384 + q * 2048 + row * 16;
            *reinterpret_cast<float4*>(p) = make_float4(h[0], h[1], h[2], h[3]);
            *reinterpret_cast<float4*>(p + 8192) = make_float4(l[0], l[1], l[2], l[3]);
        }
    }
}
// expansion products of one knot block: D[q] = coef(axis) x image^T for q = xd, yd, xdd, ydd (, y)
__device__ __forceinline__ void issue_expand(uint32_t tmem, uint32_t sb, uint32_t ab, int kb, int n, int nq) {
    const uint32_t id = idesc(n, 0);
    for (int q = 0; q < nq; q++) {
        const uint32_t img = sb + (q < 2 ? 2u : q < 4 ? 4u : 0u) * IMG + (uint32_t)kb * 16;   // Pd, Pdd, P (hi; lo = +IMG)
        const uint32_t ax = ab + ((q == 1 || q == 3 || q == 4) ? 16384u : 0u);
#pragma unroll
        for (int s = 0; s < 2; s++)
            mma4(tmem + q * 32, ax + s * 4096, ax + 8192 + s * 4096, img + s * 2 * KSTR, img + IMG + s * 2 * KSTR, KSTR, 128, id, s == 0);
    }
}
// reduction product of one 8-knot chunk: D[acc] (+)= A(slot m) x transposed image[0..15][t0 .. t0+7]
__device__ __forceinline__ void issue_reduce(uint32_t tmem, uint32_t sb, uint32_t ab, int m, int img_idx, int t0, int acc, bool first) {
    const uint32_t img = sb + OFF_R + 2u * img_idx * RIMG + (uint32_t)(t0 >> 2) * RSTR;
    mma4(tmem + COL_UW + acc * 16, ab + 2 * m * MAT, ab + (2 * m + 1) * MAT, img, img + RIMG, RSTR, 128, idesc(16, 0), first);
}
__device__ __forceinline__ float sq(float x) { return x * x; }
// polar re-parametrisation + clip of one (x, y) pair without trigonometry: with alpha = atan2(y, x) (plus any multiple of 2 pi from
// jnp.unwrap) cos(alpha) = x / r and sin(alpha) = y / r, so the reference's  d = clip((x cos + y sin) / (cos^2 + sin^2))  is clip(r) and the
// target point is (x, y) * clip(r) / r  [projection.py:73-99, 217-243].  atan2(0, 0) = 0 makes the direction (1, 0) at the origin.
__device__ __forceinline__ void polar_fast(float x, float y, float lo, float hi, float& bx, float& by) {
    const float r = sqrtf(x * x + y * y);
    const float d = dm::clip_(r, lo, hi);
    const bool origin = r == 0.0f;
    const float k = d / r;
    bx = origin ? d : k * x;
    by = origin ? 0.0f : k * y;
}
}  // namespace ptc

__global__ void __launch_bounds__(ptc::CTA_THREADS, 2) k_project_tc(DCfg c, ProjArgs a) {
    using namespace ptc;
    extern __shared__ __align__(128) unsigned char smraw[];
    const uint32_t sb = smem_u32(smraw);
    const float* sGx = reinterpret_cast<const float*>(smraw + B_BYTES);
    const float* sGy = sGx + 77; const float* sKx = sGy + 88; const float* sKy = sKx + 154;
    unsigned char* abuf = smraw + OFF_A;
    const uint32_t ab = sb + OFF_A;
    unsigned long long* bar_c = reinterpret_cast<unsigned long long*>(smraw + OFF_A + ABUF);
    unsigned long long* bar_m = bar_c + 1;                                // MMA group retired (tcgen05.commit)
    unsigned long long* bar_f = bar_c + 2;                                // operands of a step written by all 128 compute threads
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_c + 3);
    const int tid = threadIdx.x, warp = tid >> 5;

    if (tid == 0) { mbar_init(bar_c, 1); mbar_init(bar_m, 1); mbar_init(bar_f, THREADS); }
    __syncthreads();
    if (tid == 0) bulk_g2s(smraw, c.proj_tc_const, CONST_BYTES, bar_c);
    if (warp == 0) {
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t tl = tmem + ((uint32_t)(warp * 32) << 16);         // this warp's 32 TMEM lanes
    mbar_wait(bar_c, 0);

    if (warp == 4) {
        // ---- MMA issuer: one elected lane replays the step sequence of the compute warps; each step waits until all 128 compute
        // threads have written their operands (bar_f), issues the step's MMAs and commits them to bar_m
        if (tid == 4 * 32) {
            uint32_t fpar = 0;
#define PTC_STEP() do { mbar_wait(bar_f, fpar); fpar ^= 1u; fence_after(); } while (0)
            for (int rd = 0; rd < 4; rd++) {                                   // lane term of the linear cost
                const int nch = rd < 3 ? 4 : 1;
                PTC_STEP();
                for (int m = 0; m < nch; m++) issue_reduce(tmem, sb, ab, m, 0, 8 * (4 * rd + m), 3, rd == 0 && m == 0);   // Wy = P^T (LA_ub - LA_lb)
                commit(bar_m);
            }
            for (int pass = 0; pass < 2; pass++) {
                for (int blk = 0; blk < 4; blk++) {
                    const int kb = blk < 3 ? 32 * blk : 88, n = blk < 3 ? 32 : 16, nsc = blk < 3 ? 4 : 1, c0 = blk < 3 ? 0 : 8;
                    PTC_STEP();
                    issue_expand(tmem, sb, ab, kb, n, pass == 0 ? 4 : 5);
                    commit(bar_m);
                    for (int sc = 0; sc < nsc; sc++) {
                        const int t0 = kb + c0 + 8 * sc;
                        const bool first = t0 == 0;
                        PTC_STEP();
                        if (pass == 0) {
                            issue_reduce(tmem, sb, ab, 0, 1, t0, 0, first);     // Ux  = Pd^T r_vx
                            issue_reduce(tmem, sb, ab, 1, 1, t0, 1, first);     // Wx  = Pd^T b_vx
                            issue_reduce(tmem, sb, ab, 2, 1, t0, 2, first);     // Uy  = Pd^T r_vy
                            issue_reduce(tmem, sb, ab, 3, 1, t0, 3, false);     // Wy += Pd^T b_vy
                        } else {
                            issue_reduce(tmem, sb, ab, 0, 1, t0, 0, first);     // Ux  = Pd^T r_vx
                            issue_reduce(tmem, sb, ab, 1, 1, t0, 2, first);     // Uy  = Pd^T r_vy
                            issue_reduce(tmem, sb, ab, 2, 0, t0, 2, false);     // Uy += P^T (r_lane_ub - r_lane_lb)
                        }
                        commit(bar_m);
                        PTC_STEP();
                        if (pass == 0) {
                            issue_reduce(tmem, sb, ab, 0, 2, t0, 0, false);     // Ux += Pdd^T r_ax
                            issue_reduce(tmem, sb, ab, 1, 2, t0, 1, false);     // Wx += Pdd^T b_ax
                            issue_reduce(tmem, sb, ab, 2, 2, t0, 2, false);     // Uy += Pdd^T r_ay
                            issue_reduce(tmem, sb, ab, 3, 2, t0, 3, false);     // Wy += Pdd^T b_ay
                        } else {
                            issue_reduce(tmem, sb, ab, 0, 2, t0, 0, false);     // Ux += Pdd^T r_ax
                            issue_reduce(tmem, sb, ab, 1, 2, t0, 2, false);     // Uy += Pdd^T r_ay
                        }
                        commit(bar_m);
                    }
                }
            }
#undef PTC_STEP
        }
    } else {
    const int g0 = blockIdx.x * THREADS + tid;
    const bool live0 = g0 < a.n_samples;
    const int g = live0 ? g0 : a.n_samples - 1;                         // idle rows shadow the last sample and store nothing
    const int e = g / a.B;
    const float* bqx = a.beq_x + e * 3; const float* bqy = a.beq_y + e * 4;
    uint32_t par = 0; bool pending = false;
#define PTC_WAIT() do { if (pending) { mbar_wait(bar_m, par); par ^= 1u; pending = false; } } while (0)
#define PTC_PUBLISH() do { fence_async_smem(); fence_before(); mbar_arrive(bar_f); } while (0)      // operands written: hand the step to the issuer warp

    // ---- x_guess (same fma chains as k_project)  [cem_helper.py:169-230]
    float cf[22];
    {
        float in_x[7], in_y[8];
        const float4 p0 = *reinterpret_cast<const float4*>(a.params + (size_t)g * NPAR), p1 = *reinterpret_cast<const float4*>(a.params + (size_t)g * NPAR + 4);
        in_x[0] = p0.x; in_x[1] = p0.y; in_x[2] = p0.z; in_x[3] = p0.w; in_x[4] = bqx[0]; in_x[5] = bqx[1]; in_x[6] = bqx[2];
        in_y[0] = p1.x; in_y[1] = p1.y; in_y[2] = p1.z; in_y[3] = p1.w; in_y[4] = bqy[0]; in_y[5] = bqy[1]; in_y[6] = bqy[2]; in_y[7] = bqy[3];
#pragma unroll
        for (int j = 0; j < NV; j++) {
            float s = 0.0f;
#pragma unroll
            for (int k = 0; k < 7; k++) s = fmaf(sGx[j * 7 + k], in_x[k], s);
            cf[j] = s;
            s = 0.0f;
#pragma unroll
            for (int k = 0; k < 8; k++) s = fmaf(sGy[j * 8 + k], in_y[k], s);
            cf[NV + j] = s;
        }
    }
    const float* slr = a.s_lane + (size_t)g * 2 * NL;

    // ---- pass 1: guess derivatives -> polar clip (unwrap does not change cos / sin, see polar_fast) -> P^T r, P^T b  [projection.py:73-131, 158-166]
    {
        // lane term of the linear cost first: Wy = P^T (LA_ub - LA_lb), four 8-knot chunks per round  [projection.py:127-131]
        for (int rd = 0; rd < 4; rd++) {
            const int nch = rd < 3 ? 4 : 1;
            // all loads of the round in flight before the first use, as 8-byte accesses (a row of s_lane is 8-byte aligned: 198 floats).
            // knot t reads s_lane[t - 1] (upper bound rows) and s_lane[98 + t] (lower bound rows); fu / fl hold indices 32 rd - 2 .. and 32 rd + 98 ..
            float fu[34], fl[32];
#pragma unroll
            for (int k = 0; k < 17; k++) {
                const int idx = 32 * rd - 2 + 2 * k;
                float2 v = make_float2(0.0f, 0.0f);
                if (idx >= 0 && idx < 2 * NL && 2 * k < 8 * nch + 2) v = *reinterpret_cast<const float2*>(slr + idx);
                fu[2 * k] = v.x; fu[2 * k + 1] = v.y;
            }
#pragma unroll
            for (int k = 0; k < 16; k++) {
                const int idx = 32 * rd + 98 + 2 * k;
                float2 v = make_float2(0.0f, 0.0f);
                if (idx < 2 * NL && 2 * k < 8 * nch) v = *reinterpret_cast<const float2*>(slr + idx);
                fl[2 * k] = v.x; fl[2 * k + 1] = v.y;
            }
            PTC_WAIT();
#pragma unroll
            for (int m = 0; m < 4; m++) {
                if (m < nch) {
                    float dl[8];
#pragma unroll
                    for (int i = 0; i < 8; i++) {
                        const int t = 32 * rd + 8 * m + i;
                        dl[i] = (t >= 1 && t < T_) ? (c.b_lane_ub - fu[8 * m + i + 1]) - (c.b_lane_lb - fl[8 * m + i]) : 0.0f;
                    }
                    put8(abuf, m, tid, dl);
                }
            }
            PTC_PUBLISH();
            pending = true;
        }
        for (int blk = 0; blk < 4; blk++) {
            const int kb = blk < 3 ? 32 * blk : 88, n = blk < 3 ? 32 : 16, nsc = blk < 3 ? 4 : 1, c0 = blk < 3 ? 0 : 8;
            PTC_WAIT();
            put_coef(abuf, tid, cf);
            PTC_PUBLISH();
            pending = true;
            PTC_WAIT();
            fence_after();
            for (int sc = 0; sc < nsc; sc++) {
                const int t0 = kb + c0 + 8 * sc;
                float gx[8], gy[8], r_x[8], b_x[8], r_y[8], b_y[8];
                tld8(tl + 0 * 32 + c0 + 8 * sc, gx); tld8(tl + 1 * 32 + c0 + 8 * sc, gy);
                tld_wait();
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    const int t = t0 + i;
                    float bx, by;
                    polar_fast(gx[i], gy[i], c.v_min, c.v_max, bx, by);
                    const bool ok = t < T_;
                    r_x[i] = ok ? gx[i] - bx : 0.0f; b_x[i] = ok ? bx : 0.0f;
                    r_y[i] = ok ? gy[i] - by : 0.0f; b_y[i] = ok ? by : 0.0f;
                    if ((a.dbg & 3) == 1 && live0 && ok) { a.acc[(size_t)g * T_ + t] = gx[i]; a.steer[(size_t)g * T_ + t] = gy[i]; }
                }
                PTC_WAIT();
                put8(abuf, 0, tid, r_x); put8(abuf, 1, tid, b_x); put8(abuf, 2, tid, r_y); put8(abuf, 3, tid, b_y);
                PTC_PUBLISH();
                pending = true;
                tld8(tl + 2 * 32 + c0 + 8 * sc, gx); tld8(tl + 3 * 32 + c0 + 8 * sc, gy);
                tld_wait();
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    const int t = t0 + i;
                    float bx, by;
                    polar_fast(gx[i], gy[i], 0.0f, c.a_max, bx, by);
                    const bool ok = t < T_;
                    r_x[i] = ok ? gx[i] - bx : 0.0f; b_x[i] = ok ? bx : 0.0f;
                    r_y[i] = ok ? gy[i] - by : 0.0f; b_y[i] = ok ? by : 0.0f;
                    if ((a.dbg & 3) == 2 && live0 && ok) { a.acc[(size_t)g * T_ + t] = gx[i]; a.steer[(size_t)g * T_ + t] = gy[i]; }
                }
                PTC_WAIT();
                put8(abuf, 0, tid, r_x); put8(abuf, 1, tid, b_x); put8(abuf, 2, tid, r_y); put8(abuf, 3, tid, b_y);
                PTC_PUBLISH();
                pending = true;
            }
        }
    }
    PTC_WAIT();
    fence_after();

    // ---- multiplier update, linear cost, KKT solve  [projection.py:115-119, 158-171]
    {
        float ux[8], ux2[8], wx[8], wx2[8], uy[8], uy2[8], wy[8], wy2[8];
        tld8(tl + COL_UW + 0, ux); tld8(tl + COL_UW + 8, ux2); tld8(tl + COL_UW + 16, wx); tld8(tl + COL_UW + 24, wx2);
        tld8(tl + COL_UW + 32, uy); tld8(tl + COL_UW + 40, uy2); tld8(tl + COL_UW + 48, wy); tld8(tl + COL_UW + 56, wy2);
        tld_wait();
        float rhs_x[14], rhs_y[15];
        float* lxg = a.lam_x + (size_t)g * NV; float* lyg = a.lam_y + (size_t)g * NV;
#pragma unroll
        for (int j = 0; j < NV; j++) {
            const float u_x = j < 8 ? ux[j] : ux2[j - 8], w_x = j < 8 ? wx[j] : wx2[j - 8];
            const float u_y = j < 8 ? uy[j] : uy2[j - 8], w_y = j < 8 ? wy[j] : wy2[j - 8];
            const float lx = lxg[j] - u_x, ly = lyg[j] - u_y;
            rhs_x[j] = (lx + cf[j]) + w_x;
            rhs_y[j] = (ly + cf[NV + j]) + w_y;
            if (live0) { lxg[j] = a.dbg ? u_x : lx; lyg[j] = a.dbg ? u_y : ly; }
            if (a.dbg && live0) { a.cx[(size_t)g * NV + j] = w_x; a.cy[(size_t)g * NV + j] = w_y; }
        }
        rhs_x[11] = bqx[0]; rhs_x[12] = bqx[1]; rhs_x[13] = bqx[2];
        rhs_y[11] = bqy[0]; rhs_y[12] = bqy[1]; rhs_y[13] = bqy[2]; rhs_y[14] = bqy[3];
#pragma unroll
        for (int j = 0; j < NV; j++) {
            float s = 0.0f;
#pragma unroll
            for (int k = 0; k < 14; k++) s = fmaf(sKx[j * 14 + k], rhs_x[k], s);
            cf[j] = s;
            s = 0.0f;
#pragma unroll
            for (int k = 0; k < 15; k++) s = fmaf(sKy[j * 15 + k], rhs_y[k], s);
            cf[NV + j] = s;
        }
        if (live0 && !a.dbg) {
#pragma unroll
            for (int j = 0; j < NV; j++) { a.cx[(size_t)g * NV + j] = cf[j]; a.cy[(size_t)g * NV + j] = cf[NV + j]; }
        }
    }

    const bool live = live0 && !a.dbg;
    // ---- pass 2: projected trajectories -> lane slacks, residuals, P^T r, controls, cost terms  [projection.py:173-272, cem_helper.py:540-551, :232-262]
    float q_acc = 0.0f, q_vel = 0.0f, q_lane = 0.0f;                       // squared residual norms
    float q_v = 0.0f, q_s = 0.0f, q_sv = 0.0f, q_sa = 0.0f, q_p1 = 0.0f, q_p2 = 0.0f, q_ydd = 0.0f, q_xdd = 0.0f;
    {
        const float vdes = a.v_des[e];
        float v_prev = 0.0f, st_prev = 0.0f, sv_prev = 0.0f, s1_carry = 0.0f;
        float apend[8];                                                      // acc[t0 - 8 .. t0 - 1] being assembled (acc[t - 1] needs the speed at knot t)
#pragma unroll
        for (int i = 0; i < 8; i++) apend[i] = 0.0f;
        float* slw = a.s_lane + (size_t)g * 2 * NL;
        float* accg = a.acc + (size_t)g * T_; float* steerg = a.steer + (size_t)g * T_;
        for (int blk = 0; blk < 4; blk++) {
            const int kb = blk < 3 ? 32 * blk : 88, n = blk < 3 ? 32 : 16, nsc = blk < 3 ? 4 : 1, c0 = blk < 3 ? 0 : 8;
            PTC_WAIT();
            put_coef(abuf, tid, cf);
            PTC_PUBLISH();
            pending = true;
            PTC_WAIT();
            fence_after();
            for (int sc = 0; sc < nsc; sc++) {
                const int t0 = kb + c0 + 8 * sc;
                float xd[8], yd[8], xdd[8], ydd[8], yy[8];
                tld8(tl + 0 * 32 + c0 + 8 * sc, xd); tld8(tl + 1 * 32 + c0 + 8 * sc, yd); tld8(tl + 2 * 32 + c0 + 8 * sc, xdd);
                tld8(tl + 3 * 32 + c0 + 8 * sc, ydd); tld8(tl + 4 * 32 + c0 + 8 * sc, yy);
                tld_wait();
                if (t0 == 0) {
                    // knot 0 is pinned by the boundary conditions: with the vehicle's lateral velocity and acceleration at rest the exact steering
                    // there is 0, where the reference's beta noise model Beta(a|steer|, b|steer|) (cem_helper.py:427-436) is singular.  The
                    // reference (and k_project) get a rounding-level value from the fp32 dot product; evaluate this one knot the same way.
                    const float* r0 = sKy + 165;
                    xd[0] = dot11(r0 + NV, cf); yd[0] = dot11(r0 + NV, cf + NV); xdd[0] = dot11(r0 + 2 * NV, cf); ydd[0] = dot11(r0 + 2 * NV, cf + NV);
                    yy[0] = dot11(r0, cf + NV);
                }
                float rvx[8], rvy[8], rax[8], ray[8], dlb[8], st8[8], s1v[8], s2v[8];
#pragma unroll
                for (int i = 0; i < 8; i++) { s1v[i] = 0.0f; s2v[i] = 0.0f; }
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    const int t = t0 + i;
                    const bool ok = t < T_;
                    // velocity residuals of the re-projected point  [projection.py:217-255]
                    float bvx, bvy;
                    polar_fast(xd[i], yd[i], c.v_min, c.v_max, bvx, bvy);
                    rvx[i] = ok ? xd[i] - bvx : 0.0f; rvy[i] = ok ? yd[i] - bvy : 0.0f;
                    q_vel = fmaf(rvx[i], rvx[i], q_vel); q_vel = fmaf(rvy[i], rvy[i], q_vel);
                    // lane slacks and residuals at knot t (row t-1 of A_lane_bound)  [projection.py:182-183]
                    float d = 0.0f;
                    if (t >= 1 && ok) {
                        const float Ay = yy[i];
                        const float s1 = dm::max0_(-Ay + c.b_lane_ub), r1 = (Ay - c.b_lane_ub) + s1;
                        const float s2 = dm::max0_(Ay + c.b_lane_lb), r2 = (-Ay - c.b_lane_lb) + s2;
                        s1v[i] = s1; s2v[i] = s2;
                        q_lane = fmaf(r1, r1, q_lane); q_lane = fmaf(r2, r2, q_lane);
                        d = r1 - r2;
                    }
                    dlb[i] = d;
                }
                if (live) {
                    // slacks as 8-byte stores: upper-bound rows s_lane[t - 1] in pairs (t0 - 2, t0 - 1), (t0, t0 + 1), .. with the slack of knot
                    // t0 - 1 carried from the previous chunk; lower-bound rows s_lane[98 + t] in pairs starting at the even index 98 + t0
                    // (the pair of chunk 0 touches index 98, which the last chunk rewrites with the upper-bound slack of knot 99)
                    if (t0 >= 8) *reinterpret_cast<float2*>(slw + t0 - 2) = make_float2(s1_carry, s1v[0]);
                    if (t0 + 2 < T_) *reinterpret_cast<float2*>(slw + t0) = make_float2(s1v[1], s1v[2]);
                    if (t0 + 4 < T_) { *reinterpret_cast<float2*>(slw + t0 + 2) = make_float2(s1v[3], s1v[4]); *reinterpret_cast<float2*>(slw + t0 + 4) = make_float2(s1v[5], s1v[6]); }
                    else slw[t0 + 2] = s1v[3];                                                        // knot 99 -> index 98
                    *reinterpret_cast<float2*>(slw + 98 + t0) = make_float2(s2v[0], s2v[1]);
                    *reinterpret_cast<float2*>(slw + 100 + t0) = make_float2(s2v[2], s2v[3]);
                    if (t0 + 4 < T_) { *reinterpret_cast<float2*>(slw + 102 + t0) = make_float2(s2v[4], s2v[5]); *reinterpret_cast<float2*>(slw + 104 + t0) = make_float2(s2v[6], s2v[7]); }
                }
                s1_carry = s1v[7];
                PTC_WAIT();
                put8(abuf, 0, tid, rvx); put8(abuf, 1, tid, rvy); put8(abuf, 2, tid, dlb);
                PTC_PUBLISH();
                pending = true;
                float anew[7];
#pragma unroll
                for (int i = 0; i < 7; i++) anew[i] = 0.0f;
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    const int t = t0 + i;
                    const bool ok = t < T_;
                    // acceleration residuals
                    float bax, bay;
                    polar_fast(xdd[i], ydd[i], 0.0f, c.a_max, bax, bay);
                    rax[i] = ok ? xdd[i] - bax : 0.0f; ray[i] = ok ? ydd[i] - bay : 0.0f;
                    q_acc = fmaf(rax[i], rax[i], q_acc); q_acc = fmaf(ray[i], ray[i], q_acc);
                    // controls and cost terms  [cem_helper.py:540-551, :232-262]
                    const float s2v = xd[i] * xd[i] + yd[i] * yd[i];
                    const float v = sqrtf(s2v);
                    const float curv = (ydd[i] * xd[i] - yd[i] * xdd[i]) / (s2v * sqrtf(s2v));
                    const float st = dm::atan_(curv * c.wheel_base);
                    st8[i] = st;
                    if (ok) {
                        if (t >= 1) {
                            const float av = (v - v_prev) / c.dt;
                            if (i == 0) apend[7] = av; else anew[i - 1] = av;
                            const float sv = st - st_prev;
                            q_sv = fmaf(sv, sv, q_sv);
                            q_p2 = fmaf(sq(dm::max0_(fabsf(sv) - c.steer_rate_pen)), 1.0f, q_p2);
                            if (t >= 2) q_sa = fmaf(sq(sv - sv_prev), 1.0f, q_sa);
                            sv_prev = sv;
                        }
                        q_v = fmaf(sq(v - vdes), 1.0f, q_v);
                        q_s = fmaf(st, st, q_s);
                        q_p1 = fmaf(sq(dm::max0_(fabsf(st) - c.steer_max)), 1.0f, q_p1);
                        q_ydd = fmaf(ydd[i], ydd[i], q_ydd); q_xdd = fmaf(xdd[i], xdd[i], q_xdd);
                        v_prev = v; st_prev = st;
                    }
                }
                if (live && t0 >= 8) {
                    *reinterpret_cast<float4*>(accg + t0 - 8) = make_float4(apend[0], apend[1], apend[2], apend[3]);
                    *reinterpret_cast<float4*>(accg + t0 - 4) = make_float4(apend[4], apend[5], apend[6], apend[7]);
                }
#pragma unroll
                for (int i = 0; i < 7; i++) apend[i] = anew[i];
                if (live) {
                    *reinterpret_cast<float4*>(steerg + t0) = make_float4(st8[0], st8[1], st8[2], st8[3]);
                    if (t0 + 4 < T_) *reinterpret_cast<float4*>(steerg + t0 + 4) = make_float4(st8[4], st8[5], st8[6], st8[7]);
                }
                PTC_WAIT();
                put8(abuf, 0, tid, rax); put8(abuf, 1, tid, ray);
                PTC_PUBLISH();
                pending = true;
            }
        }
        if (live) *reinterpret_cast<float4*>(accg + T_ - 4) = make_float4(apend[0], apend[1], apend[2], (v_prev - v_prev) / c.dt);     // acc[96..98], acc[99] = 0
    }
    PTC_WAIT();
    fence_after();
    {
        float ux[8], ux2[8], uy[8], uy2[8];
        tld8(tl + COL_UW + 0, ux); tld8(tl + COL_UW + 8, ux2); tld8(tl + COL_UW + 32, uy); tld8(tl + COL_UW + 40, uy2);
        tld_wait();
        if (live) {
            float* lxg = a.lam_x + (size_t)g * NV; float* lyg = a.lam_y + (size_t)g * NV;
#pragma unroll
            for (int j = 0; j < NV; j++) {
                lxg[j] = lxg[j] - (j < 8 ? ux[j] : ux2[j - 8]);
                lyg[j] = lyg[j] - (j < 8 ? uy[j] : uy2[j - 8]);
            }
            const float res_norm = (sqrtf(q_acc) + sqrtf(q_vel)) + sqrtf(q_lane);
            const float base = ((((res_norm + 0.1f * sqrtf(q_v)) + 0.1f * ((sqrtf(q_s) + sqrtf(q_sv)) + sqrtf(q_sa))) + 0.1f * (sqrtf(q_p1) + sqrtf(q_p2))) +
                                0.02f * sqrtf(q_ydd)) + 0.02f * sqrtf(q_xdd);
            a.res_norm[g] = res_norm; a.cost_base[g] = base;
        }
    }
#undef PTC_WAIT
#undef PTC_PUBLISH
    }   // compute warps
    fence_before();
    __syncthreads();
    if (warp == 0) {
        __syncwarp();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(TMEM_COLS) : "memory");
    }
}
