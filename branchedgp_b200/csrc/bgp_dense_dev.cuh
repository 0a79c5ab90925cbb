// Device helpers shared by the one-CTA-per-item dense pipeline (bgp_dense.cu) and the multi-CTA pipeline (bgp_dense_wide.cu).
#pragma once
#include "bgp_dense.cuh"
#include "bgp_gemm.cuh"

namespace bgp {

#define DENSE_SETUP() DENSE_SETUP_IDX(blockIdx.x)
#define DENSE_SETUP_IDX(BIDX)                                                                      \
    extern __shared__ __align__(128) unsigned char smem[];                                         \
    const int tid = threadIdx.x;                                                                   \
    const int slot = P.slot_map ? P.slot_map[(BIDX)] : (int)(BIDX);                            \
    const int item = P.item0 + slot;                                                               \
    const DenseDims& dm = P.d;                                                                     \
    const VecLayout VL = vec_layout(dm.Mp, dm.nM, dm.D);                                           \
    double* const vec = P.vec + (size_t)slot * P.slot_stride;                                       \
    double* const LC = P.LC + (size_t)slot * P.slot_stride;                                         \
    double* const RX = P.RX + (size_t)slot * P.slot_stride;                                         \
    double* const PB = P.PB + (size_t)slot * P.slot_stride;                                         \
    double* const LB = P.LB + (size_t)slot * P.slot_stride;                                         \
    double* const LCe = P.LCe + (size_t)slot * P.slot_stride;                                 \
    double* const LinvD = P.LinvD + (size_t)slot * P.slot_stride;                        \
    double* const RinvD = P.RinvD + (size_t)slot * P.slot_stride;                        \
    double* const SCR = P.SCR + (size_t)slot * P.slot_stride;                                \
    const double ell = P.hyp[(size_t)item * 3 + 0];                                                \
    const double kvar = P.hyp[(size_t)item * 3 + 1];                                               \
    const double likvar = P.hyp[(size_t)item * 3 + 2];                                             \
    (void)tid; (void)vec; (void)LC; (void)RX; (void)PB; (void)LB; (void)LCe; (void)LinvD; (void)RinvD; (void)SCR; \
    (void)ell; (void)kvar; (void)likvar; (void)VL; (void)smem;

// ------------------------------------------------------------------------------------------------ covariance
// K[(t,f),(t',f')] for cell-major rows 3*cell + f  (VBHelperFunctions.py:170-194; branch_kernParamGPflow.py:117-198).
__device__ __forceinline__ double kgen(const DenseParams& P, const double* __restrict__ ka, const double* __restrict__ KT, int gi,
                                       int gj, double kinv) {
    const int M = P.d.M;
    if (gi >= M || gj >= M) return gi == gj ? 1.0 : 0.0;   // identity padding
    if (P.KConst != nullptr) return P.KConst[(size_t)gi * M + gj];
    const int ci = gi / 3, fi = gi - 3 * ci, cj = gj / 3, fj = gj - 3 * cj;
    double k;
    if (fi == fj) k = KT[(size_t)ci * P.d.N + cj];
    else k = (ka[ci] * kinv) * ka[cj];
    if (gi == gj) k += P.square_noise;
    return k;
}

// ------------------------------------------------------------------------------------------------ 128x128 helpers in shared memory
// The 128x128 diagonal macro blocks are factored (chol128) and inverted (trinv128) in shared memory, row-major with leading dimension
// DLD; the upper triangle of the block is scratch.  The 32x32 single-warp routines below serve the batched sampler (bgp_sample.cu).

// In-place lower Cholesky of the 32x32 tile T (leading dimension ld) by ONE warp: lane = row, left-looking (column j is formed
// from the finished columns to its left, so the inner loop only loads; a fully unrolled register version ran out of instruction
// cache).  One reciprocal square root per pivot instead of a square root and a division on the critical path.
__device__ __forceinline__ void chol32_warp(double* T, int ld, int& fail, int base_index, int lane) {
    double* row = T + lane * ld;
    for (int j = 0; j < TS; ++j) {
        const double* rj = T + j * ld;
        double s0 = row[j], s1 = 0.0;
        for (int k = 0; k + 1 < j; k += 2) {
            s0 = fma(-row[k], rj[k], s0);
            s1 = fma(-row[k + 1], rj[k + 1], s1);
        }
        if (j & 1) s0 = fma(-row[j - 1], rj[j - 1], s0);
        const double a = s0 + s1;                          // A_ij - sum_k L_ik L_jk
        double d = __shfl_sync(0xffffffffu, a, j);
        if (!(d > 0.0)) {
            if (fail == 0) fail = base_index + j + 1;
            d = 1.0;
        }
        const double rinv = rsqrt(d);
        __syncwarp();
        if (lane >= j) row[j] = (lane == j) ? d * rinv : a * rinv;
        __syncwarp();
    }
}

// X = inverse of the lower-triangular 32x32 tile L (upper part of L ignored, upper part of X zeroed) by ONE warp: lane c solves
// column c by forward substitution (it only re-reads what it wrote itself).
__device__ __forceinline__ void trinv32_warp(const double* L, int ldl, double* X, int ldx, int lane) {
    const double dinv = 1.0 / L[lane * ldl + lane];   // all 32 divisions in parallel
    for (int i = 0; i < TS; ++i) {
        const double* li = L + i * ldl;
        double s0 = (lane == i) ? 1.0 : 0.0, s1 = 0.0;
        for (int k = 0; k + 1 < i; k += 2) {
            s0 = fma(-li[k], X[k * ldx + lane], s0);
            s1 = fma(-li[k + 1], X[(k + 1) * ldx + lane], s1);
        }
        if (i & 1) s0 = fma(-li[i - 1], X[(i - 1) * ldx + lane], s0);
        const double rii = __shfl_sync(0xffffffffu, dinv, i);
        X[i * ldx + lane] = (lane <= i) ? (s0 + s1) * rii : 0.0;
    }
    __syncwarp();
}

// In-place lower Cholesky of Dm (row-major, leading dimension DLD).  Non-positive pivots are replaced by 1 and reported.
// Right-looking over 8-column blocks (16 steps of three CTA barriers).  A measurement of the multi-CTA pipeline
// (tools/wide_timers.py) put the previous version -- 32-column panels, the 32x32 diagonal tile factored AND inverted by one warp
// through shared memory -- at 90 us per block, 60 % of the critical path of the tile Cholesky.  Here the serial part of a step
// is an 8x8 diagonal block in the registers of one warp (pivots and multipliers travel by shuffles, one rsqrt per pivot, no
// inverse: the panel rows are forward-substituted, one thread per row) and the rank-8 trailing update runs as 8x8 DMMA tiles
// over all sixteen warps.
static __device__ void chol128(double* Dm, int& fail, int base_index) {
    __shared__ int fail_s;
    __shared__ double rinv_s[2][8];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, c = lane & 3;
    constexpr int NB8 = MB / 8;
    // 8x8 diagonal block J in the registers of the calling warp: lane = row (lanes 8..31 carry identity rows); pivots and
    // multipliers travel by shuffles, one rsqrt per pivot
    auto factor_diag = [&](int J) {
        const int j0 = 8 * J;
        double a[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) a[k] = (lane < 8 && k <= lane) ? Dm[(j0 + lane) * DLD + j0 + k] : ((k == lane) ? 1.0 : 0.0);
        double rinv[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            double sacc = a[j];
#pragma unroll
            for (int k = 0; k < j; ++k) sacc = fma(-a[k], __shfl_sync(0xffffffffu, a[k], j), sacc);
            double d = __shfl_sync(0xffffffffu, sacc, j);
            if (!(d > 0.0)) {
                if (lane == 0 && fail_s == 0) fail_s = base_index + j0 + j + 1;
                d = 1.0;
            }
            rinv[j] = rsqrt(d);
            a[j] = (lane < j) ? 0.0 : ((lane == j) ? d * rinv[j] : sacc * rinv[j]);
        }
        if (lane < 8) {
#pragma unroll
            for (int k = 0; k < 8; ++k)
                if (k <= lane) Dm[(j0 + lane) * DLD + j0 + k] = a[k];
        }
        if (lane == 0) {
#pragma unroll
            for (int k = 0; k < 8; ++k) rinv_s[J & 1][k] = rinv[k];
        }
    };
    // one trailing 8x8 tile (rt, ct) of step J: D -= L_rt L_ct^T (two DMMAs)
    auto update_tile = [&](int J, int rt, int ct) {
        const int j0 = 8 * J;
        double* dp = Dm + (8 * rt + g) * DLD + 8 * ct + 2 * c;
        double d0 = dp[0], d1 = dp[1];
        const double* ar = Dm + (8 * rt + g) * DLD + j0 + c;
        const double* br = Dm + (8 * ct + g) * DLD + j0 + c;
        dmma(d0, d1, -ar[0], br[0]);
        dmma(d0, d1, -ar[4], br[4]);
        dp[0] = d0;
        dp[1] = d1;
    };
    if (tid == 0) fail_s = 0;
    __syncthreads();
    if (warp == 0) factor_diag(0);
    for (int J = 0; J < NB8; ++J) {
        const int j0 = 8 * J;
        __syncthreads();
        // panel row i:  L_i L_JJ^T = A_i  by forward substitution
        {
            const int i = j0 + 8 + tid;
            if (i < MB) {
                double* row = Dm + i * DLD + j0;
                const double* Ld = Dm + j0 * DLD + j0;
                double l[8];
#pragma unroll
                for (int cc = 0; cc < 8; ++cc) {
                    double sacc = row[cc];
#pragma unroll
                    for (int k = 0; k < cc; ++k) sacc = fma(-l[k], Ld[cc * DLD + k], sacc);
                    l[cc] = sacc * rinv_s[J & 1][cc];
                }
#pragma unroll
                for (int cc = 0; cc < 8; ++cc) row[cc] = l[cc];
            }
        }
        __syncthreads();
        // trailing 8x8 tiles (rt, ct), J < ct <= rt.  Lookahead: warp 0 updates the next diagonal tile first and factors it at once
        // (the serial part of the next step) while the other warps work through the rest of the update.
        const int T0 = J + 1, nrem = NB8 - T0;
        if (nrem > 0) {
            if (warp == 0) {
                update_tile(J, T0, T0);
                __syncwarp();
                factor_diag(T0);
            } else {
                int rr = 0, cr = warp;   // tile index 0 is (T0, T0): taken by warp 0
                for (int t8 = warp; t8 < nrem * (nrem + 1) / 2; t8 += GEMM_WARPS - 1, cr += GEMM_WARPS - 1) {
                    while (cr > rr) { cr -= rr + 1; ++rr; }
                    update_tile(J, T0 + rr, T0 + cr);
                }
            }
        }
    }
    __syncthreads();
    if (fail == 0) fail = fail_s;
    __syncthreads();
}

// In-place inverse of the lower-triangular Dm (row-major, leading dimension DLD) by recursive doubling on the fp64 tensor pipe.
// Level 0: the sixteen 8x8 diagonal blocks, one warp each, in registers (lane = column of the inverse, forward substitution
// through shuffles).  Level s = 8, 16, 32, 64: every pair of adjacent inverted s x s diagonal blocks A (top) and B (bottom) with
// the original off-diagonal block C between them becomes the inverse of the 2s x 2s block:  C <- -B (C A).  T = C A is parked
// in the mirror block above the diagonal (the upper triangle of Dm is scratch), so no second buffer is needed; both products
// run as 8x8 DMMA tiles dealt to the sixteen warps, with the k ranges cut to the triangles of A and B.  Two CTA barriers per
// level.  (The version this replaces -- 32x32 diagonal tiles inverted by forward substitution through shared memory, then
// three levels of tile products on the CUDA cores -- took 23 us per block on the chain of the tile Cholesky.)
static __device__ void trinv128(double* Dm) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, c = lane & 3;
    __syncthreads();
    {
        // 8x8 diagonal block `warp`: lanes 0..7 hold the rows of L, then lane = column of X = inv(L)
        const int j0 = 8 * warp;
        double a[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) a[k] = (lane < 8 && k <= lane) ? Dm[(j0 + lane) * DLD + j0 + k] : ((k == lane) ? 1.0 : 0.0);
        double dl = 1.0;   // this lane's own diagonal entry, then its reciprocal: one division per lane, all in parallel
#pragma unroll
        for (int k = 0; k < 8; ++k)
            if (lane == k) dl = a[k];
        const double rd = 1.0 / dl;
        double x[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            double sacc = (i == lane) ? 1.0 : 0.0;
#pragma unroll
            for (int k = 0; k < i; ++k) sacc = fma(-__shfl_sync(0xffffffffu, a[k], i), x[k], sacc);
            const double rii = __shfl_sync(0xffffffffu, rd, i);   // outside the select: every lane must take part in the shuffle
            x[i] = (i >= lane) ? sacc * rii : 0.0;
        }
        __syncwarp();
        if (lane < 8) {
#pragma unroll
            for (int i = 0; i < 8; ++i)
                if (i >= lane) Dm[(j0 + i) * DLD + j0 + lane] = x[i];
        }
    }
    for (int sz = 8; sz < MB; sz *= 2) {
        const int npair = MB / (2 * sz), tps = sz / 8;       // pairs of blocks, 8x8 tiles per block side
        const int ntile = npair * tps * tps;
        __syncthreads();
        // T = C A  -> mirror block (rows of A, columns of B).  T[i][j] = sum_{k >= j} C[i][k] A[k][j]  (A lower triangular)
        for (int t = warp; t < ntile; t += GEMM_WARPS) {
            const int pr = t / (tps * tps), r = t % (tps * tps), ti = r / tps, tj = r % tps;
            const int o = pr * 2 * sz;                           // top-left corner of the pair
            const double* Cb = Dm + (o + sz) * DLD + o;          // C: rows o+sz.., columns o..
            const double* Ab = Dm + o * DLD + o;
            // two accumulator chains over alternating 8-wide k blocks, the four fragments of a block loaded before its MMAs
            double d0 = 0.0, d1 = 0.0, e0 = 0.0, e1 = 0.0;
            const double* ap = Cb + (8 * ti + g) * DLD + c;
            const int jc = 8 * tj + g;
            for (int k0 = 8 * tj; k0 < sz; k0 += 8) {
                const double a0 = ap[k0], a1 = ap[k0 + 4];
                const double b0 = (k0 + c >= jc) ? Ab[(k0 + c) * DLD + jc] : 0.0;
                const double b1 = (k0 + 4 + c >= jc) ? Ab[(k0 + 4 + c) * DLD + jc] : 0.0;
                dmma(d0, d1, a0, b0);
                dmma(e0, e1, a1, b1);
            }
            d0 += e0; d1 += e1;
            double* Tb = Dm + o * DLD + o + sz;                  // mirror block: rows o.., columns o+sz..
            Tb[(8 * ti + g) * DLD + 8 * tj + 2 * c] = d0;
            Tb[(8 * ti + g) * DLD + 8 * tj + 2 * c + 1] = d1;
        }
        __syncthreads();
        // C <- -B T.  C[i][j] = -sum_{k <= i} B[i][k] T[k][j]  (B lower triangular)
        for (int t = warp; t < ntile; t += GEMM_WARPS) {
            const int pr = t / (tps * tps), r = t % (tps * tps), ti = r / tps, tj = r % tps;
            const int o = pr * 2 * sz;
            const double* Bb = Dm + (o + sz) * DLD + o + sz;
            const double* Tb = Dm + o * DLD + o + sz;
            double d0 = 0.0, d1 = 0.0, e0 = 0.0, e1 = 0.0;
            const int ir = 8 * ti + g;
            const double* bp = Tb + c * DLD + 8 * tj + g;
            for (int k0 = 0; k0 < 8 * ti + 8; k0 += 8) {
                const double a0 = (k0 + c <= ir) ? -Bb[ir * DLD + k0 + c] : 0.0;
                const double a1 = (k0 + 4 + c <= ir) ? -Bb[ir * DLD + k0 + 4 + c] : 0.0;
                const double b0 = bp[k0 * DLD], b1 = bp[(k0 + 4) * DLD];
                dmma(d0, d1, a0, b0);
                dmma(e0, e1, a1, b1);
            }
            d0 += e0; d1 += e1;
            double* Cw = Dm + (o + sz) * DLD + o;
            Cw[(8 * ti + g) * DLD + 8 * tj + 2 * c] = d0;
            Cw[(8 * ti + g) * DLD + 8 * tj + 2 * c + 1] = d1;
        }
    }
    __syncthreads();
}

// Lower tiles (a >= b) of the 128x128 smem block -> global tiles.  packed != nullptr: block-lower matrix position
// (4J+a, 4J+b);  block != nullptr: 4x4 tile block.  Upper triangle of diagonal tiles is written as zero.
// dalt (optional): diagonal tiles with `dadd` added on the diagonal.
static __device__ void store_lower_block(const double* Dm, int J, double* packed, double* block, double* dalt, double dadd) {
    const int tid = threadIdx.x;
    for (int a = 0; a < TPM; ++a)
        for (int b = 0; b <= a; ++b) {
            double* tp = packed ? packed + packed_tile(TPM * J + a, TPM * J + b) * TILE_ELEMS : nullptr;
            double* tb = block ? block + (size_t)(a * TPM + b) * TILE_ELEMS : nullptr;
            double* td = (dalt && a == b) ? dalt + (size_t)(TPM * J + a) * TILE_ELEMS : nullptr;
            for (int e = tid; e < TILE_ELEMS; e += GEMM_THREADS) {
                const int i = e >> 5, j = e & 31;
                double v = Dm[(TS * a + i) * DLD + TS * b + j];
                if (a == b && j > i) v = 0.0;
                const int o = tswz(i, j);
                if (tp) tp[o] = v;
                if (tb) tb[o] = v;
                if (td) td[o] = v + ((i == j) ? dadd : 0.0);
            }
        }
}

// sum over the stored lower triangle of Ktilde o dK/dtheta.  w carries 2*Kbar_ij off the diagonal and Kbar_ii on it.
__device__ __forceinline__ void hyper_accum(const DenseParams& P, const double* __restrict__ ka, const double* __restrict__ dkal,
                                            const double* __restrict__ dkab, const double* __restrict__ KT, int gi, int gj, double w,
                                            double kvar, double kinv, double& gl, double& gv, double& gb) {
    const int M = P.d.M;
    if (gi >= M || gj >= M || gj > gi) return;
    if (gi == gj) w *= 0.5;
    const int ci = gi / 3, fi = gi - 3 * ci, cj = gj / 3, fj = gj - 3 * cj;
    if (fi == fj) {
        const size_t e = (size_t)ci * P.d.N + cj;
        gl += w * KT[(size_t)P.d.N * P.d.N + e];
        gv += w * KT[e] / kvar;
    } else {
        const double a = ka[ci], b = ka[cj];
        const double cr = (a * kinv) * b;
        gl += w * (dkal[ci] * b + a * dkal[cj]) * kinv;
        gv += w * cr * (2.0 / kvar - kinv);
        gb += w * (dkab[ci] * b + a * dkab[cj]) * kinv;
    }
}

}  // namespace bgp
