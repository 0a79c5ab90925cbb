// Inducing-point AssignGP bounds and gradients (sparse BGP, MBGP multi-output), sm_100a.
//
//   Kuu = K(Z,Z) + (noise + jitter) I,  L = chol(Kuu),  V = L^-1 K(Z,X),  P = V Delta V^T + I,  R = chol(P),
//   c = R^-1 V (Phi^T Y) / s,  bound = trace - N D/2 log(2 pi s) - D/2 sum log R_ii^2 - sum Y^2 / (2 s) + sum c^2 / 2 - KL
//   (assigngp_denseSparse.py:59-107; MBGP/assigngp.py:492-558 sums this over outputs with per-output b_d and masks);
//   gradient = SURVEY.md App. B.2.
// Matrices are Mz x Mz (Mz = 30 for BGP, 180 for MBGP): they are factored / inverted in shared memory (packed lower
// triangle) by one CTA per sub-problem, while the 3N columns of X are streamed in slabs of 32 with K(Z, x_j) generated
// in-kernel.  Every GEMM-shaped piece (slab products, blocked Cholesky / inverse, the triangular products of the Kuu adjoint) runs on
// the fp64 tensor pipe (mma.sync.m8n8k4.f64, SASS DMMA).  The O(N * 3N) assignment passes (bgp_phi.cu) dominate the sparse BGP sizes.
#include "bgp_sparse.cuh"
#include "bgp_dense.cuh"
#include "bgp_launch.h"

namespace bgp {

#define SP_SETUP()                                                                                            \
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;                                            \
    const int slot = blockIdx.y;                                                                              \
    const int item = P.item0 + slot / P.SD;                                                                   \
    const int sd = slot % P.SD;                                                                               \
    const int Mz = P.Mz, Mp = P.Mp, M = P.M, N = P.N, D = P.D;                                                \
    const SparseLayout& SLy = P.lay;                                                                          \
    double* const ws = P.ws + (size_t)slot * P.slot_stride;                                                   \
    const double ell = P.hyp[(size_t)item * 3 + 0], kvar = P.hyp[(size_t)item * 3 + 1], likvar = P.hyp[(size_t)item * 3 + 2]; \
    const double bpt = P.bv[(size_t)item * P.SD + sd];                                                        \
    const double kinv = 1.0 / (kvar + JITTER);                                                                \
    (void)tid; (void)warp; (void)lane; (void)sd; (void)M; (void)N; (void)D; (void)ell; (void)kvar; (void)likvar; (void)bpt; (void)kinv; (void)ws;

__device__ __forceinline__ int pk(int i, int j) { return i * (i + 1) / 2 + j; }   // packed lower, j <= i

// ---- Cholesky / inverse of a packed lower-triangular matrix in shared memory on the fp64 tensor pipe (8 x 8 blocks = one DMMA tile; the packed layout is addressed element by element,
// which is all a DMMA fragment needs).  `dinv` receives the inverses of the diagonal blocks ([blocks][64], zero above the
// diagonal and in the padding of a ragged last block); the inverse routine reuses them.
__device__ __forceinline__ int pk_blocks(int n) { return (n + 7) >> 3; }

// Right-looking: diagonal block factored + inverted by warp 0, panel = A_panel inv(L_JJ)^T with a thread per row, trailing
// lower tiles updated by rank-8 MMAs.  Three CTA barriers per 8 columns instead of three per column.
__device__ void chol_packed_blocked(double* A, int n, int& fail, double* dinv) {
    __shared__ int fail_s;
    const int tid = threadIdx.x, nt = blockDim.x, warp = tid >> 5, lane = tid & 31, nwarps = nt >> 5;
    const int g = lane >> 2, c = lane & 3;
    const int nblk = pk_blocks(n);
    if (tid == 0) fail_s = 0;
    for (int J = 0; J < nblk; ++J) {
        const int j0 = 8 * J, w = min(8, n - j0);
        double* Dv = dinv + 64 * J;
        __syncthreads();
        if (warp == 0) {
            // the 8 x 8 diagonal block in the registers of lanes 0..7 (lane = row; rows beyond a ragged edge are identity rows):
            // pivots and multipliers travel by shuffles, no shared-memory round trip on the critical path of the whole CTA
            double a[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) a[k] = (lane < w && k <= lane) ? A[pk(j0 + lane, j0 + k)] : ((k == lane) ? 1.0 : 0.0);
            double rinv[8];   // reciprocal pivots: one rsqrt per column, no square root or division on this serial path
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                double sacc = a[j];
#pragma unroll
                for (int k = 0; k < j; ++k) sacc = fma(-a[k], __shfl_sync(0xffffffffu, a[k], j), sacc);
                double d = __shfl_sync(0xffffffffu, sacc, j);
                if (!(d > 0.0)) {
                    if (lane == 0 && fail_s == 0) fail_s = j0 + j + 1;
                    d = 1.0;
                }
                rinv[j] = rsqrt(d);
                a[j] = (lane < j) ? 0.0 : ((lane == j) ? d * rinv[j] : sacc * rinv[j]);
            }
            double x[8];   // lane = column of the inverse
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                double sacc = (i == lane) ? 1.0 : 0.0;
#pragma unroll
                for (int k = 0; k < i; ++k) sacc = fma(-__shfl_sync(0xffffffffu, a[k], i), x[k], sacc);
                x[i] = (i >= lane) ? sacc * rinv[i] : 0.0;
            }
            if (lane < 8) {
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    if (lane < w && k <= lane) A[pk(j0 + lane, j0 + k)] = a[k];
                    Dv[k * 8 + lane] = (k < w && lane < w) ? x[k] : 0.0;
                }
            }
        }
        __syncthreads();
        for (int i = j0 + w + tid; i < n; i += nt) {   // panel row i:  L_i = A_i inv(L_JJ)^T
            double a[8], l[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) a[k] = (k < w) ? A[pk(i, j0 + k)] : 0.0;
#pragma unroll
            for (int cc = 0; cc < 8; ++cc) {
                double sacc = 0.0;
#pragma unroll
                for (int k = 0; k <= cc; ++k) sacc = fma(a[k], Dv[cc * 8 + k], sacc);
                l[cc] = sacc;
            }
#pragma unroll
            for (int cc = 0; cc < 8; ++cc)
                if (cc < w) A[pk(i, j0 + cc)] = l[cc];
        }
        __syncthreads();
        const int T0 = J + 1, nrem = nblk - T0;   // trailing tiles (rt, ct), T0 <= ct <= rt < nblk
        int rr = 0, cr = warp;
        for (int tix = warp; tix < nrem * (nrem + 1) / 2; tix += nwarps, cr += nwarps) {
            while (cr > rr) { cr -= rr + 1; ++rr; }
            const int m = 8 * (T0 + rr) + g, nrow = 8 * (T0 + cr) + g, n0 = 8 * (T0 + cr) + 2 * c;
            const bool vm = m < n;
            double d0 = (vm && n0 <= m) ? A[pk(m, n0)] : 0.0;
            double d1 = (vm && n0 + 1 <= m) ? A[pk(m, n0 + 1)] : 0.0;
#pragma unroll
            for (int kk = 0; kk < 8; kk += 4) {
                const bool vk = kk + c < w;
                const double a = (vm && vk) ? -A[pk(m, j0 + kk + c)] : 0.0;
                const double b = (nrow < n && vk) ? A[pk(nrow, j0 + kk + c)] : 0.0;
                dmma(d0, d1, a, b);
            }
            if (vm && n0 <= m) A[pk(m, n0)] = d0;
            if (vm && n0 + 1 <= m) A[pk(m, n0 + 1)] = d1;
        }
    }
    __syncthreads();
    if (fail == 0) fail = fail_s;
    __syncthreads();
}

// In-place inverse of the packed lower-triangular factor, one 8-row block at a time:  X_IJ = -inv(L_II) sum_{J <= K < I} L_IK X_KJ.
// Row block I of L is read by every warp before any of it is overwritten (results wait in registers across the barrier).
// `wscr`: 8 x 9 doubles per warp.  Needs (blocks - 1) <= 4 x warps.
__device__ void trinv_packed_blocked(double* A, int n, const double* dinv, double* wscr) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nwarps = blockDim.x >> 5;
    const int g = lane >> 2, c = lane & 3;
    const int nblk = pk_blocks(n);
    double* S = wscr + warp * 72;
    for (int I = 0; I < nblk; ++I) {
        const int i0 = 8 * I;
        const double* Dv = dinv + 64 * I;
        const int m = i0 + g;
        const bool vm = m < n;
        double res[4][2];
        __syncthreads();
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int J = warp + q * nwarps;
            res[q][0] = 0.0; res[q][1] = 0.0;
            if (J < I) {
                double d0 = 0.0, d1 = 0.0, e0 = 0.0, e1 = 0.0;
                const int bcol = 8 * J + g;
                for (int K = J; K < I; ++K) {
                    const int r0 = 8 * K + c, r1 = r0 + 4;
                    const double a0 = vm ? A[pk(m, r0)] : 0.0, a1 = vm ? A[pk(m, r1)] : 0.0;
                    const double b0 = (bcol <= r0) ? A[pk(r0, bcol)] : 0.0, b1 = (bcol <= r1) ? A[pk(r1, bcol)] : 0.0;
                    dmma(d0, d1, a0, b0);
                    dmma(e0, e1, a1, b1);
                }
                S[g * 9 + 2 * c] = d0 + e0;
                S[g * 9 + 2 * c + 1] = d1 + e1;
                __syncwarp();
                double x0 = 0.0, x1 = 0.0;
#pragma unroll
                for (int kk = 0; kk < 8; kk += 4) dmma(x0, x1, -Dv[g * 8 + kk + c], S[(kk + c) * 9 + g]);
                res[q][0] = x0; res[q][1] = x1;
                __syncwarp();
            }
        }
        __syncthreads();
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int J = warp + q * nwarps;
            if (J < I && vm) { A[pk(m, 8 * J + 2 * c)] = res[q][0]; A[pk(m, 8 * J + 2 * c + 1)] = res[q][1]; }
        }
        if (warp == nwarps - 1)
            for (int e = lane; e < 64; e += 32) {
                const int r = e >> 3, cc = e & 7;
                if (cc <= r && i0 + r < n) A[pk(i0 + r, i0 + cc)] = Dv[e];
            }
    }
    __syncthreads();
}

// ---------------------------------------------------------------------------------------------------- k_sp_reduce_partials
// A_j, Delta_j, v_jd from the row-chunk partials of the assignment pass; grid (column blocks, slots) so that the
// nch x Mp partial arrays (300 MB per item at N = 10 000) are read with the whole GPU, in a fixed order (deterministic).
__global__ void __launch_bounds__(NTHREADS) k_sp_reduce_partials(SparseParams P) {
    const int slot = blockIdx.y;
    const int item = P.item0 + slot / P.SD;
    const int Mp = P.Mp, M = P.M, D = P.D, nch = P.nch;
    const SparseLayout& SLy = P.lay;
    double* const ws = P.ws + (size_t)slot * P.slot_stride;
    const double likvar = P.hyp[(size_t)item * 3 + 2];
    const int j = blockIdx.x * NTHREADS + threadIdx.x;
    if (j >= Mp) return;
    double a = 0.0;
    if (j < M)
        for (int ch = 0; ch < nch; ++ch) a += ws[SLy.Apart + (size_t)ch * Mp + j];
    ws[SLy.A + j] = a;
    ws[SLy.Delta + j] = a / likvar;
    for (int d = 0; d < D; ++d) {
        double s = 0.0;
        if (j < M)
            for (int ch = 0; ch < nch; ++ch) s += ws[SLy.vpart + ((size_t)ch * D + d) * Mp + j];
        ws[SLy.v + (size_t)d * Mp + j] = s;
    }
}

// ---------------------------------------------------------------------------------------------------- k_sp_prep
__global__ void __launch_bounds__(NTHREADS) k_sp_prep(SparseParams P) {
    SP_SETUP();
    __shared__ double red[32];
    const int nch = P.nch;
    double suma = 0.0;
    for (int j = tid; j < Mp; j += NTHREADS) suma += ws[SLy.A + j];   // A, Delta, v were reduced by k_sp_reduce_partials
    for (int i = tid; i < N; i += NTHREADS) {
        double k, dl, dd;
        kbase_grads(P.t[i] - bpt, ell, kvar, P.kind, k, dl, dd);
        ws[SLy.kax + i] = k; ws[SLy.dkalx + i] = dl; ws[SLy.dkabx + i] = -dd;
    }
    for (int m = tid; m < Mz; m += NTHREADS) {
        double k, dl, dd;
        kbase_grads(P.Z[2 * m] - bpt, ell, kvar, P.kind, k, dl, dd);
        ws[SLy.kaz + m] = k; ws[SLy.dkalz + m] = dl; ws[SLy.dkabz + m] = -dd;
    }
    suma = block_sum(suma, red);
    double kl = 0.0;
    for (int ch = tid; ch < nch; ch += NTHREADS) kl += ws[SLy.klpart + ch];
    kl = block_sum(kl, red);
    const double* Yg = P.Y + (size_t)P.item_gene[item] * N * P.Dtot + (P.SD > 1 ? sd : 0);
    double y2 = 0.0;
    for (int e = tid; e < N * D; e += NTHREADS) { const double y = Yg[(size_t)(e / D) * P.Dtot + (e % D)]; y2 += y * y; }
    y2 = block_sum(y2, red);
    if (tid == 0) {
        ws[SLy.scal + SS_KL] = kl; ws[SLy.scal + SS_SUMY2] = y2; ws[SLy.scal + SS_SUMA] = suma;
        if (sd == 0) P.status[item] = 0;
    }
}

// Slab products on the fp64 tensor pipe (mma.sync.m8n8k4.f64, SASS DMMA): a slab is [Mz][SP_LDS] in shared memory (SP_SL = 32
// columns of X; SP_LDS = 36 makes both fragment patterns below bank-conflict free: 36 = 4 mod 16 eight-byte banks).
// lane = 4 g + c:  a = A[g][c],  b = B[c][g],  d = {D[g][2c], D[g][2c + 1]}.
constexpr int SP_LDS = SP_SL + 4;
constexpr int SP_ZBAR_D = 4;
constexpr int SP_CELLS = (SP_SL + 2) / 3 + 2;   // cells touched by a slab of SP_SL columns (<= 12)
#ifndef BGP_SP_T
#define BGP_SP_T 512
#endif
constexpr int SP_T = BGP_SP_T;             // threads of the slab kernels (k_sp_fwd, k_sp_bwd) and of k_sp_kuu / k_sp_mid
constexpr int SP_WARPS = SP_T / 32;

// Cs[m][0..32) = sum_k A[m][k] Bs[k][.],  A = Mz x Mz row-major in global memory (L1 / L2 resident, explicit zeros outside its
// triangle).  TRI: 0 full, 1 lower (k <= m), 2 upper (k >= m) -- only narrows the k range.
template <int TRI>
__device__ __forceinline__ void slab_mm(const double* __restrict__ A, const double* __restrict__ Bs, double* __restrict__ Cs, int Mz,
                                        int warp, int lane) {
    const int g = lane >> 2, c = lane & 3;
    const int nrt = (Mz + 7) >> 3;
    // 8-row tiles dealt to the warps in snake order (w, 15 - w, 16 + w, ...): the triangular k ranges add up evenly
    for (int i = 0; (i >> 1) * 2 * SP_WARPS < nrt; ++i) {
        const int base = (i >> 1) * 2 * SP_WARPS;
        const int js = (i & 1) ? base + 2 * SP_WARPS - 1 - warp : base + warp;
        if (js >= nrt) continue;
        const int rt = (TRI == 1) ? nrt - 1 - js : js;   // longest k ranges first: k <= m grows with the row tile, k >= m shrinks
        const int m = rt * 8 + g;
        const bool vm = m < Mz;
        const double* pa = A + (size_t)min(m, Mz - 1) * Mz + c;
        double acc[4][2];
#pragma unroll
        for (int t = 0; t < 4; ++t) { acc[t][0] = 0.0; acc[t][1] = 0.0; }
        const int k0 = (TRI == 2) ? rt * 8 : 0;
        const int k1 = (TRI == 1) ? min(Mz, rt * 8 + 8) : Mz;
        // the A fragments come from global memory (L2, ~600 cycles): a block of 8 k steps (32 MMAs) is loaded while the previous
        // block is multiplied
        constexpr int PF = 8;
        double nxt[PF];
#pragma unroll
        for (int q = 0; q < PF; ++q) nxt[q] = (vm && k0 + 4 * q + c < Mz) ? pa[k0 + 4 * q] : 0.0;
        for (int kb = k0; kb < k1; kb += 4 * PF) {
            double cur[PF];
#pragma unroll
            for (int q = 0; q < PF; ++q) cur[q] = nxt[q];
            if (kb + 4 * PF < k1) {
#pragma unroll
                for (int q = 0; q < PF; ++q) nxt[q] = (vm && kb + 4 * PF + 4 * q + c < Mz) ? pa[kb + 4 * PF + 4 * q] : 0.0;
            }
#pragma unroll
            for (int q = 0; q < PF; ++q) {
                const int k = kb + 4 * q;
                if (k < k1) {
                    const int kk = k + c;
                    const bool vk = kk < Mz;
                    const double* bk = Bs + (size_t)min(kk, Mz - 1) * SP_LDS + g;
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        const double b = vk ? bk[8 * t] : 0.0;
                        dmma(acc[t][0], acc[t][1], cur[q], b);
                    }
                }
            }
        }
        if (vm) {
#pragma unroll
            for (int t = 0; t < 4; ++t) { Cs[(size_t)m * SP_LDS + 8 * t + 2 * c] = acc[t][0]; Cs[(size_t)m * SP_LDS + 8 * t + 2 * c + 1] = acc[t][1]; }
        }
    }
}

// dst[m][n] (+)= sign * sum_jj As[m][jj] Bs[n][jj] for n <= m (dst = Mz x Mz row-major in global memory, lower triangle).  The 8 x 8
// tiles of the lower triangle are dealt to the warps one by one; two accumulator chains (even / odd k steps) per tile.
__device__ __forceinline__ void slab_outer_lower(const double* __restrict__ As, const double* __restrict__ Bs, double* __restrict__ dst,
                                                 int Mz, bool first, double sign, int warp, int lane) {
    const int g = lane >> 2, c = lane & 3;
    const int nt = (Mz + 7) >> 3;
    const int ntiles = nt * (nt + 1) / 2;
    int rt = 0, ct = warp;   // tile (rt, ct) of the lower triangle, advanced by SP_WARPS tiles per trip
    for (int tix = warp; tix < ntiles; tix += SP_WARPS, ct += SP_WARPS) {
        while (ct > rt) { ct -= rt + 1; ++rt; }
        const int m = rt * 8 + g, nr = ct * 8 + g;
        const bool vm = m < Mz, vn = nr < Mz;
        const double* pa = As + (size_t)min(m, Mz - 1) * SP_LDS + c;
        const double* pb = Bs + (size_t)min(nr, Mz - 1) * SP_LDS + c;
        double d0 = 0.0, d1 = 0.0, e0 = 0.0, e1 = 0.0;
#pragma unroll
        for (int k = 0; k < SP_SL; k += 8) {
            const double a0 = vm ? pa[k] : 0.0, b0 = vn ? pb[k] : 0.0;
            const double a1 = vm ? pa[k + 4] : 0.0, b1 = vn ? pb[k + 4] : 0.0;
            dmma(d0, d1, a0, b0);
            dmma(e0, e1, a1, b1);
        }
        d0 += e0; d1 += e1;
        if (vm) {
            const int n = ct * 8 + 2 * c;
            double* d = dst + (size_t)m * Mz + n;
            // every element is owned by this thread for the whole kernel: a plain store on the first slab, then reductions
            // (RED.ADD.F64, nothing to wait for) in program order -- deterministic
            if (first) {
                if (n <= m) d[0] = sign * d0;
                if (n + 1 <= m) d[1] = sign * d1;
            } else {
                if (n <= m) atomicAdd(&d[0], sign * d0);
                if (n + 1 <= m) atomicAdd(&d[1], sign * d1);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------- k_sp_kuu
__global__ void __launch_bounds__(SP_T) k_sp_kuu(SparseParams P) {
    SP_SETUP();
    extern __shared__ __align__(16) double sm[];
    int fail = 0;
    for (int e = tid; e < Mz * Mz; e += SP_T) {
        const int m = e / Mz, n = e % Mz;
        if (n > m) continue;
        const int fm = (int)P.Z[2 * m + 1], fn = (int)P.Z[2 * n + 1];
        double k = (fm == fn) ? kbase(P.Z[2 * m] - P.Z[2 * n], ell, kvar, P.kind) : (ws[SLy.kaz + m] * kinv) * ws[SLy.kaz + n];
        if (m == n) k += P.square_noise + JITTER;   // White / noise_level 1e-6 + jitter 1e-6 (assigngp_denseSparse.py:72-75)
        sm[pk(m, n)] = k;
    }
    __syncthreads();
    double* const dinv = sm + (size_t)Mz * (Mz + 1) / 2;          // inverse diagonal blocks, then per-warp scratch
    double* const wscr = dinv + 64 * pk_blocks(Mz);
    chol_packed_blocked(sm, Mz, fail, dinv);
    for (int e = tid; e < Mz * Mz; e += SP_T) {
        const int m = e / Mz, n = e % Mz;
        ws[SLy.L + e] = (n <= m) ? sm[pk(m, n)] : 0.0;
    }
    __syncthreads();
    trinv_packed_blocked(sm, Mz, dinv, wscr);
    for (int e = tid; e < Mz * Mz; e += SP_T) {
        const int m = e / Mz, n = e % Mz;
        const double x = (n <= m) ? sm[pk(m, n)] : 0.0;
        ws[SLy.Linv + e] = x;
        ws[SLy.LinvT + (size_t)n * Mz + m] = x;
    }
    if (tid == 0 && fail != 0) atomicCAS(&P.status[item], 0, fail);
}

// ---------------------------------------------------------------------------------------------------- k_sp_fwd
// grid (nsplit, slots).  V = Linv Kuf for this CTA's slabs; partial P = V Delta V^T, z = V v, trace.
__global__ void __launch_bounds__(SP_T) k_sp_fwd(SparseParams P) {
    SP_SETUP();
    extern __shared__ __align__(16) double sm[];
    __shared__ double red[32];
    constexpr int LDS_ = SP_LDS;
    double* Ks = sm;                       // [Mz][33]  K(Z, x_j) of the slab
    double* Vs = sm + (size_t)Mz * LDS_;   // [Mz][33]  V of the slab
    double* Dls = Vs + (size_t)Mz * LDS_;  // [32] Delta of the slab
    // per-row constants of the inducing inputs and per-cell values of the slab, staged once instead of being fetched from global
    // memory for every entry (ncu: a quarter of the stall samples sat on those dependent loads)
    double* vsl_s = Dls + SP_SL;           // [SP_ZBAR_D][32] v of the slab (outputs beyond SP_ZBAR_D are read from global memory)
    double* zt_s = vsl_s + SP_ZBAR_D * SP_SL;   // [Mz] inducing time
    double* kaz_s = zt_s + Mz;             // [Mz] k(z_m, b) / (k(b,b) + jitter)
    int* fz_s = reinterpret_cast<int*>(kaz_s + Mz);   // [Mz] function index 0..2
    double* tc_s = kaz_s + Mz + (Mz + 1) / 2;         // [SP_CELLS] cell time
    double* kax_s = tc_s + SP_CELLS;                  // [SP_CELLS] k(t_cell, b)
    for (int m = tid; m < Mz; m += SP_T) {
        zt_s[m] = P.Z[2 * m];
        fz_s[m] = (int)P.Z[2 * m + 1] - 1;
        kaz_s[m] = ws[SLy.kaz + m] * kinv;
    }
    const int split = blockIdx.x, nsplit = P.nsplit;
    const double* Linv = ws + SLy.Linv;
    double* Pp = ws + SLy.Ppart + (size_t)split * Mz * Mz;
    double* zp = ws + SLy.zpart + (size_t)split * D * Mz;
    const int nslab = (M + SP_SL - 1) / SP_SL;
    double tr = 0.0;
    bool first = true;
    for (int sl = split; sl < nslab || first; sl += nsplit) {
        const bool empty = sl >= nslab;   // a split without slabs still zero-initialises its partials
        const int j0 = sl * SP_SL;
        __syncthreads();
        // K(Z, x_j) of the slab (branch_kernParamGPflow.py:117-198) in two divergence-free sweeps: every (inducing row, cell) pair has
        // ONE same-function column 3 cell + f_z that needs the base kernel (an exp); all other entries are products of the
        // per-row / per-cell kernels against the branching point
        const int cell0 = j0 / 3, ncell = empty ? 0 : (min(j0 + SP_SL, M) + 2) / 3 - cell0;
        if (tid < ncell) { tc_s[tid] = P.t[cell0 + tid]; kax_s[tid] = ws[SLy.kax + cell0 + tid]; }
        if (tid < SP_SL) Dls[tid] = (!empty && j0 + tid < M) ? ws[SLy.Delta + j0 + tid] : 0.0;
        if (tid >= 64 && tid < 64 + SP_ZBAR_D * SP_SL) {
            const int d = (tid - 64) / SP_SL, jj = (tid - 64) % SP_SL;
            vsl_s[d * SP_SL + jj] = (d < D && !empty && j0 + jj < M) ? ws[SLy.v + (size_t)d * Mp + j0 + jj] : 0.0;
        }
        __syncthreads();
        if (!(P.dbg_skip & 4))
        for (int e = tid; e < Mz * SP_SL; e += SP_T) {
            const int m = e / SP_SL, jj = e % SP_SL, j = j0 + jj;
            const int cell = j / 3, f = j - 3 * cell;
            if (empty || j >= M) Ks[m * LDS_ + jj] = 0.0;
            else if (fz_s[m] != f) Ks[m * LDS_ + jj] = kaz_s[m] * kax_s[cell - cell0];
        }
        if (!(P.dbg_skip & 4))
        for (int e = tid; e < Mz * ncell; e += SP_T) {
            const int m = e / ncell, lc = e - m * ncell;
            const int jj = 3 * (cell0 + lc) + fz_s[m] - j0;
            if (jj >= 0 && jj < SP_SL && j0 + jj < M) Ks[m * LDS_ + jj] = kbase(zt_s[m] - tc_s[lc], ell, kvar, P.kind);
        }
        __syncthreads();
        if (!(P.dbg_skip & 1)) slab_mm<1>(Linv, Ks, Vs, Mz, warp, lane);   // V = Linv K(Z, X_slab)
        __syncthreads();
        for (int e = tid; e < Mz * SP_SL; e += SP_T) {
            const int m = e / SP_SL, jj = e % SP_SL;
            const double x = Vs[m * LDS_ + jj];
            if (!empty && j0 + jj < Mp) ws[SLy.V + (size_t)m * Mp + j0 + jj] = x;
            tr += Dls[jj] * x * x;
        }
        // Ks <- Delta_j V (scaled copy) so that P += Ks Vs^T
        for (int e = tid; e < Mz * SP_SL; e += SP_T) {
            const int m = e / SP_SL, jj = e % SP_SL;
            Ks[m * LDS_ + jj] = Dls[jj] * Vs[m * LDS_ + jj];
        }
        __syncthreads();
        if (!(P.dbg_skip & 2)) slab_outer_lower(Ks, Vs, Pp, Mz, first, 1.0, warp, lane);
        for (int e = tid; e < D * Mz; e += SP_T) {
            const int d = e / Mz, m = e % Mz;
            double s = 0.0;
            if (!empty) {
                if (d < SP_ZBAR_D) {
#pragma unroll 8
                    for (int jj = 0; jj < SP_SL; ++jj) s += Vs[m * LDS_ + jj] * vsl_s[d * SP_SL + jj];   // (columns >= M hold zeros)
                } else {
                    for (int jj = 0; jj < SP_SL; ++jj)
                        if (j0 + jj < M) s += Vs[m * LDS_ + jj] * ws[SLy.v + (size_t)d * Mp + j0 + jj];
                }
            }
            zp[e] = first ? s : (zp[e] + s);
        }
        first = false;
    }
    tr = block_sum(tr, red);
    if (tid == 0) ws[SLy.hpart + (size_t)split * 8 + 0] = tr;
}

// ---------------------------------------------------------------------------------------------------- k_sp_mid
__global__ void __launch_bounds__(SP_T) k_sp_mid(SparseParams P) {
    SP_SETUP();
    extern __shared__ __align__(16) double sm[];
    __shared__ double red[32];
    int fail = 0;
    for (int e = tid; e < Mz * Mz; e += SP_T) {
        const int m = e / Mz, n = e % Mz;
        if (n > m) continue;
        double s = (m == n) ? 1.0 : 0.0;
        for (int sp = 0; sp < P.nsplit; ++sp) s += ws[SLy.Ppart + (size_t)sp * Mz * Mz + e];
        sm[pk(m, n)] = s;
    }
    __syncthreads();
    double* const dinv = sm + (size_t)Mz * (Mz + 1) / 2;          // inverse diagonal blocks, then per-warp scratch
    double* const wscr = dinv + 64 * pk_blocks(Mz);
    chol_packed_blocked(sm, Mz, fail, dinv);
    double ld = 0.0;
    for (int m = tid; m < Mz; m += SP_T) { const double r = sm[pk(m, m)]; ld += log(r * r); }
    ld = block_sum(ld, red);
    for (int e = tid; e < Mz * Mz; e += SP_T) {
        const int m = e / Mz, n = e % Mz;
        ws[SLy.R + e] = (n <= m) ? sm[pk(m, n)] : 0.0;
    }
    __syncthreads();
    trinv_packed_blocked(sm, Mz, dinv, wscr);   // sm = R^-1 (packed)
    for (int e = tid; e < Mz * Mz; e += SP_T) {
        const int m = e / Mz, n = e % Mz;
        ws[SLy.Rinv + e] = (n <= m) ? sm[pk(m, n)] : 0.0;
    }
    // z, c = R^-1 z / s, q = R^-T c, zbar = q / s
    double c2 = 0.0;
    for (int d = 0; d < D; ++d) {
        double* z = ws + SLy.z + (size_t)d * Mz;
        double* cv = ws + SLy.c + (size_t)d * Mz;
        double* q = ws + SLy.q + (size_t)d * Mz;
        double* zb = ws + SLy.zbar + (size_t)d * Mz;
        for (int m = tid; m < Mz; m += SP_T) {
            double s = 0.0;
            for (int sp = 0; sp < P.nsplit; ++sp) s += ws[SLy.zpart + ((size_t)sp * D + d) * Mz + m];
            z[m] = s;
        }
        __syncthreads();
        for (int m = tid; m < Mz; m += SP_T) {
            double s = 0.0;
            for (int k = 0; k <= m; ++k) s += sm[pk(m, k)] * z[k];
            s /= likvar;
            cv[m] = s;
            c2 += s * s;
        }
        __syncthreads();
        for (int m = tid; m < Mz; m += SP_T) {
            double s = 0.0;
            for (int k = m; k < Mz; ++k) s += sm[pk(k, m)] * cv[k];
            q[m] = s;
            zb[m] = s / likvar;
        }
        __syncthreads();
    }
    c2 = block_sum(c2, red);
    // Pbar = -D/2 R^-T R^-1 - 1/2 q q^T   (full symmetric): 8 x 8 tiles of the lower triangle, X = R^-1 read from the packed matrix
    {
        const double hD = 0.5 * (double)D;
        const int g = lane >> 2, c = lane & 3;
        const int nblk = pk_blocks(Mz);
        int rt = 0, ct = warp;
        for (int tix = warp; tix < nblk * (nblk + 1) / 2; tix += SP_WARPS, ct += SP_WARPS) {
            while (ct > rt) { ct -= rt + 1; ++rt; }
            const int ma = 8 * rt + g, nb = 8 * ct + g;   // columns of X feeding the A / B fragments
            double d0 = 0.0, d1 = 0.0, e0 = 0.0, e1 = 0.0;
            for (int K = rt; K < nblk; ++K) {
                const int r0 = 8 * K + c, r1 = r0 + 4;
                const double a0 = (r0 < Mz && ma <= r0) ? sm[pk(r0, ma)] : 0.0, a1 = (r1 < Mz && ma <= r1) ? sm[pk(r1, ma)] : 0.0;
                const double b0 = (r0 < Mz && nb <= r0) ? sm[pk(r0, nb)] : 0.0, b1 = (r1 < Mz && nb <= r1) ? sm[pk(r1, nb)] : 0.0;
                dmma(d0, d1, a0, b0);
                dmma(e0, e1, a1, b1);
            }
            d0 += e0; d1 += e1;
            const int m = 8 * rt + g;
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int n = 8 * ct + 2 * c + q;
                if (m >= Mz || n > m) continue;
                double qq = 0.0;
                for (int d = 0; d < D; ++d) qq += ws[SLy.q + (size_t)d * Mz + m] * ws[SLy.q + (size_t)d * Mz + n];
                const double pb = -hD * (q ? d1 : d0) - 0.5 * qq;
                ws[SLy.Pbar + (size_t)m * Mz + n] = pb;
                ws[SLy.Pbar + (size_t)n * Mz + m] = pb;
            }
        }
    }
    if (tid == 0) {
        ws[SLy.scal + SS_LOGDET] = ld;
        ws[SLy.scal + SS_SUMC2] = c2;
        if (fail != 0) atomicCAS(&P.status[item], 0, fail + Mz);
    }
}

// ---------------------------------------------------------------------------------------------------- k_sp_bwd
// grid (nsplit, slots).  Per slab: U = Pbar V, delta, Abar, vbar, Vbar, Kuf_bar = Linv^T Vbar, partial Lbar, hyper-gradient of Kuf.
__global__ void __launch_bounds__(SP_T) k_sp_bwd(SparseParams P) {
    SP_SETUP();
    extern __shared__ __align__(16) double sm[];
    __shared__ double red[32];
    constexpr int LDS_ = SP_LDS;
    double* Vs = sm;                        // V slab
    double* Us = sm + (size_t)Mz * LDS_;    // U, then Vbar
    double* Gs = Us + (size_t)Mz * LDS_;    // Kuf_bar
    double* Dls = Gs + (size_t)Mz * LDS_;   // Delta slab [32]
    // staged per-row / per-cell constants (see k_sp_fwd)
    double* vsl_s = Dls + SP_SL;            // [SP_ZBAR_D][32] v of the slab
    double* zt_s = vsl_s + SP_ZBAR_D * SP_SL;   // [Mz]
    double* kaz_s = zt_s + Mz;              // [Mz] k(z_m, b)   (unscaled)
    double* dkalz_s = kaz_s + Mz;           // [Mz]
    double* dkabz_s = dkalz_s + Mz;         // [Mz]
    int* fz_s = reinterpret_cast<int*>(dkabz_s + Mz);     // [Mz]
    double* tc_s = dkabz_s + Mz + (Mz + 1) / 2;           // [SP_CELLS]
    double* kax_s = tc_s + SP_CELLS;
    double* dkalx_s = kax_s + SP_CELLS;
    double* dkabx_s = dkalx_s + SP_CELLS;
    double* zbar_st = dkabx_s + SP_CELLS;                  // [D][Mz], staged for D <= SP_ZBAR_D outputs only (shared memory budget)
    for (int m = tid; m < Mz; m += SP_T) {
        zt_s[m] = P.Z[2 * m];
        fz_s[m] = (int)P.Z[2 * m + 1] - 1;
        kaz_s[m] = ws[SLy.kaz + m]; dkalz_s[m] = ws[SLy.dkalz + m]; dkabz_s[m] = ws[SLy.dkabz + m];
    }
    if (D <= SP_ZBAR_D)
        for (int e = tid; e < D * Mz; e += SP_T) zbar_st[e] = ws[SLy.zbar + e];
    const double* zbar_s = (D <= SP_ZBAR_D) ? zbar_st : ws + SLy.zbar;
    const int split = blockIdx.x, nsplit = P.nsplit;
    const double* Pbar = ws + SLy.Pbar;
    const double* LinvT = ws + SLy.LinvT;
    double* Lp = ws + SLy.Lbpart + (size_t)split * Mz * Mz;
    const int nslab = (M + SP_SL - 1) / SP_SL;
    const double kdiag = kvar + P.square_noise;
    double adel = 0.0, gl = 0.0, gv = 0.0, gb = 0.0;
    bool first = true;
    for (int sl = split; sl < nslab || first; sl += nsplit) {
        const bool empty = sl >= nslab;
        const int j0 = sl * SP_SL;
        __syncthreads();
        for (int e = tid; e < Mz * SP_SL; e += SP_T) {
            const int m = e / SP_SL, jj = e % SP_SL, j = j0 + jj;
            Vs[m * LDS_ + jj] = (!empty && j < M) ? ws[SLy.V + (size_t)m * Mp + j] : 0.0;
        }
        if (tid < SP_SL) Dls[tid] = (!empty && j0 + tid < M) ? ws[SLy.Delta + j0 + tid] : 0.0;
        const int cell0 = j0 / 3, ncell = empty ? 0 : (min(j0 + SP_SL, M) + 2) / 3 - cell0;
        if (tid >= 64 && tid < 64 + SP_ZBAR_D * SP_SL) {
            const int d = (tid - 64) / SP_SL, jj = (tid - 64) % SP_SL;
            vsl_s[d * SP_SL + jj] = (d < D && !empty && j0 + jj < M) ? ws[SLy.v + (size_t)d * Mp + j0 + jj] : 0.0;
        }
        if (tid >= SP_SL && tid < SP_SL + ncell) {
            const int lc = tid - SP_SL;
            tc_s[lc] = P.t[cell0 + lc]; kax_s[lc] = ws[SLy.kax + cell0 + lc];
            dkalx_s[lc] = ws[SLy.dkalx + cell0 + lc]; dkabx_s[lc] = ws[SLy.dkabx + cell0 + lc];
        }
        __syncthreads();
        if (!(P.dbg_skip & 1)) slab_mm<0>(Pbar, Vs, Us, Mz, warp, lane);   // U = Pbar V
        __syncthreads();
        // column sums over the Mz rows: warp w takes rows w, w + 8, ...; the partials are combined through Gs (free until the
        // next product) when its Mz rows can hold them, else one warp runs over all rows
        const bool wide = 2 * SP_WARPS <= Mz;
        const int mstep = wide ? SP_WARPS : 1;
        if (wide || warp == 0) {
            double vu = 0.0, vv = 0.0;
            for (int m = wide ? warp : 0; m < Mz; m += mstep) { const double x = Vs[m * LDS_ + lane]; vu += x * Us[m * LDS_ + lane]; vv += x * x; }
            if (wide) { Gs[warp * LDS_ + lane] = vu; Gs[(SP_WARPS + warp) * LDS_ + lane] = vv; }
            else { Gs[lane] = vu; Gs[LDS_ + lane] = vv; }   // Mz >= 2 rows are always there
        }
        __syncthreads();
        if (tid < SP_SL && !empty && j0 + tid < M) {
            const int j = j0 + tid;
            double vu = 0.0, vv = 0.0;
            for (int w = 0; w < (wide ? SP_WARPS : 1); ++w) { vu += Gs[w * LDS_ + tid]; vv += Gs[((wide ? SP_WARPS : 1) + w) * LDS_ + tid]; }
            const double dlt = vu + 0.5 * vv - 0.5 * kdiag;
            ws[SLy.Abar + j] = dlt / likvar;
            adel += ws[SLy.A + j] * dlt;
        }
        __syncthreads();
        for (int d = 0; d < D; ++d) {   // vbar = V^T zbar
            if (wide || warp == 0) {
                double sacc = 0.0;
                for (int m = wide ? warp : 0; m < Mz; m += mstep) sacc += Vs[m * LDS_ + lane] * zbar_s[d * Mz + m];
                Gs[(wide ? warp : 0) * LDS_ + lane] = sacc;
            }
            __syncthreads();
            if (tid < SP_SL && !empty && j0 + tid < M) {
                double sacc = 0.0;
                for (int w = 0; w < (wide ? SP_WARPS : 1); ++w) sacc += Gs[w * LDS_ + tid];
                ws[SLy.vbar + (size_t)d * Mp + j0 + tid] = sacc;
            }
            __syncthreads();
        }
        // Vbar = 2 Delta U + zbar v^T + Delta V
        for (int e = tid; e < Mz * SP_SL; e += SP_T) {
            const int m = e / SP_SL, jj = e % SP_SL, j = j0 + jj;
            double zv = 0.0;
            if (!empty && j < M)
                for (int d = 0; d < D; ++d) zv += zbar_s[d * Mz + m] * (d < SP_ZBAR_D ? vsl_s[d * SP_SL + jj] : ws[SLy.v + (size_t)d * Mp + j]);
            Us[m * LDS_ + jj] = Dls[jj] * (2.0 * Us[m * LDS_ + jj] + Vs[m * LDS_ + jj]) + zv;
        }
        __syncthreads();
        // Kuf_bar[k][j] = sum_{m >= k} Linv[m][k] Vbar[m][j]
        if (!(P.dbg_skip & 1)) slab_mm<2>(LinvT, Us, Gs, Mz, warp, lane);   // LinvT rows are zero left of the diagonal
        __syncthreads();
        // hyper-gradient through K(Z, X): cross-function entries (products), then the same-function entry of every (row, cell) pair
        if (!empty) {
            const double gvf = 2.0 / kvar - kinv;
            for (int e = tid; e < Mz * SP_SL; e += SP_T) {
                const int m = e / SP_SL, jj = e % SP_SL, j = j0 + jj;
                if (j >= M) continue;
                const int cell = j / 3, f = j - 3 * cell, lc = cell - cell0;
                if (fz_s[m] == f) continue;
                const double w = Gs[m * LDS_ + jj] * kinv;
                const double a = kaz_s[m], b = kax_s[lc];
                gl += w * (dkalz_s[m] * b + a * dkalx_s[lc]);
                gv += w * (a * b) * gvf;
                gb += w * (dkabz_s[m] * b + a * dkabx_s[lc]);
            }
            for (int e = tid; e < Mz * ncell; e += SP_T) {
                const int m = e / ncell, lc = e - m * ncell;
                const int jj = 3 * (cell0 + lc) + fz_s[m] - j0;
                if (jj < 0 || jj >= SP_SL || j0 + jj >= M) continue;
                double k, dl, dd;
                kbase_grads(zt_s[m] - tc_s[lc], ell, kvar, P.kind, k, dl, dd);
                const double w = Gs[m * LDS_ + jj];
                gl += w * dl;
                gv += w * k / kvar;
            }
        }
        // Lbar partial = -tril(Kuf_bar V^T)
        if (!(P.dbg_skip & 2)) slab_outer_lower(Gs, Vs, Lp, Mz, first, -1.0, warp, lane);
        first = false;
    }
    adel = block_sum(adel, red);
    gl = block_sum(gl, red);
    gv = block_sum(gv, red);
    gb = block_sum(gb, red);
    if (tid == 0) {
        double* hp = ws + SLy.hpart + (size_t)split * 8;
        hp[1] = adel; hp[2] = gl; hp[3] = gv; hp[4] = gb;
    }
}

// ---------------------------------------------------------------------------------------------------- k_sp_end
// C = op(A) B on Mz x Mz row-major matrices in global memory (L2 resident) with the fp64 tensor pipe.  A warp owns an 8 x 32 strip
// of C (four accumulator tiles); fragments are plain 8-byte loads, one k step ahead of the MMAs, and the kernel runs 32 warps so
// that the L2 latency of the other warps' loads is covered.  KR selects the k range from the triangular structure of the operands:
//   0:  T = Phi(L^T Lbar)   A = L read transposed, k >= m;  output columns n <= m, diagonal halved, zeros right of it
//   1:  W = T Linv          n <= k <= m;  output tiles on / below the diagonal (entries right of the diagonal come out as 0)
//   2:  S = Linv^T W        A = LinvT,  k >= max(m, n);  all tiles
template <int KR>
__device__ __forceinline__ void gmm_tri(const double* __restrict__ A, const double* __restrict__ B, double* __restrict__ C, int Mz) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int g = lane >> 2, c = lane & 3;
    const int nrt = (Mz + 7) >> 3, nst = (Mz + 31) >> 5;
    for (int u = warp; u < nrt * nst; u += nwarps) {
        const int rt = u / nst, stp = u - rt * nst;
        const int m0 = rt * 8, n0 = stp * 32;
        if (KR != 2 && n0 > m0 + 7) continue;
        const int k0 = (KR == 0) ? m0 : (KR == 1 ? n0 : max(m0, n0));
        const int k1 = (KR == 1) ? min(Mz, m0 + 8) : Mz;
        const int m = m0 + g;
        const bool vm = m < Mz;
        double acc[4][2];
#pragma unroll
        for (int t = 0; t < 4; ++t) { acc[t][0] = 0.0; acc[t][1] = 0.0; }
        auto lda = [&](int k) -> double {
            const int kk = k + c;
            if (!vm || kk >= Mz) return 0.0;
            return (KR == 0) ? A[(size_t)kk * Mz + m] : A[(size_t)m * Mz + kk];
        };
        auto ldb = [&](int k, int t) -> double {
            const int kk = k + c, n = n0 + 8 * t + g;
            return (kk < Mz && n < Mz) ? B[(size_t)kk * Mz + n] : 0.0;
        };
        double an = lda(k0), bn[4];   // one k step ahead (three steps ahead measured slower: 0.77 against 0.66 ms at Mz = 180)
#pragma unroll
        for (int t = 0; t < 4; ++t) bn[t] = ldb(k0, t);
        for (int k = k0; k < k1; k += 4) {
            const double a = an;
            double b[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) b[t] = bn[t];
            if (k + 4 < k1) {
                an = lda(k + 4);
#pragma unroll
                for (int t = 0; t < 4; ++t) bn[t] = ldb(k + 4, t);
            }
#pragma unroll
            for (int t = 0; t < 4; ++t) dmma(acc[t][0], acc[t][1], a, b[t]);
        }
        if (vm) {
#pragma unroll
            for (int t = 0; t < 4; ++t)
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const int n = n0 + 8 * t + 2 * c + q;
                    if (n >= Mz) continue;
                    double v = acc[t][q];
                    if (KR == 0) v = (n > m) ? 0.0 : (n == m ? 0.5 * v : v);
                    C[(size_t)m * Mz + n] = v;
                }
        }
    }
}

// Cholesky adjoint of Kuu (Kuu_bar = sym(Linv^T Phi(L^T Lbar) Linv)), hyper-gradient through K(Z,Z), assembly of the
// sub-problem's bound and gradient.  outputs: sub_out[slot][8] = elbo, g_ell, g_var, g_lik, g_b
constexpr int SP_END_THREADS = 512;
__global__ void __launch_bounds__(SP_END_THREADS) k_sp_end(SparseParams P, double* sub_out) {
    SP_SETUP();
    __shared__ double red[32];
    const int NT = blockDim.x;
    const double* L = ws + SLy.L;
    const double* Linv = ws + SLy.Linv;
    const double* LinvT = ws + SLy.LinvT;
    double* T = ws + SLy.T;
    double* W = ws + SLy.W;
    double* S = ws + SLy.S;
    double gl = 0.0, gv = 0.0, gb = 0.0;
    if (P.want_grad) {
    // Lbar = sum of the column splits' partials (lower triangle; what lies above it is never part of a result)
    double* Lb = ws + SLy.Lbpart;
    if (P.nsplit > 1) {
        for (int e = tid; e < Mz * Mz; e += NT) {
            if (e % Mz > e / Mz) continue;
            double lb = Lb[e];
            for (int sp = 1; sp < P.nsplit; ++sp) lb += Lb[(size_t)sp * Mz * Mz + e];
            Lb[e] = lb;
        }
        __syncthreads();
    }
    gmm_tri<0>(L, Lb, T, Mz);        // T = Phi(L^T Lbar)
    __syncthreads();
    gmm_tri<1>(T, Linv, W, Mz);      // W = T Linv
    __syncthreads();
    gmm_tri<2>(LinvT, W, S, Mz);     // S = Linv^T W
    __syncthreads();
    // hyper-gradient through K(Z,Z): Kuu_bar = (S + S^T)/2, full sum
    for (int e = tid; e < Mz * Mz; e += NT) {
        const int m = e / Mz, n = e % Mz;
        const double w = 0.5 * (S[e] + S[(size_t)n * Mz + m]);
        const int fm = (int)P.Z[2 * m + 1], fn = (int)P.Z[2 * n + 1];
        if (fm == fn) {
            double k, dl, dd;
            kbase_grads(P.Z[2 * m] - P.Z[2 * n], ell, kvar, P.kind, k, dl, dd);
            gl += w * dl;
            gv += w * k / kvar;
        } else {
            const double a = ws[SLy.kaz + m], b = ws[SLy.kaz + n];
            gl += w * (ws[SLy.dkalz + m] * b + a * ws[SLy.dkalz + n]) * kinv;
            gv += w * ((a * kinv) * b) * (2.0 / kvar - kinv);
            gb += w * (ws[SLy.dkabz + m] * b + a * ws[SLy.dkabz + n]) * kinv;
        }
    }
    }
    gl = block_sum(gl, red);
    gv = block_sum(gv, red);
    gb = block_sum(gb, red);
    if (tid == 0) {
        double tr = 0.0, adel = 0.0;
        for (int sp = 0; sp < P.nsplit; ++sp) {
            const double* hp = ws + SLy.hpart + (size_t)sp * 8;
            tr += hp[0];
            if (P.want_grad) { adel += hp[1]; gl += hp[2]; gv += hp[3]; gb += hp[4]; }
        }
        const double s = likvar, ND = (double)N * (double)D;
        const double suma = ws[SLy.scal + SS_SUMA], y2 = ws[SLy.scal + SS_SUMY2], c2 = ws[SLy.scal + SS_SUMC2];
        const double kdiag = kvar + P.square_noise;
        const double trace = -0.5 * kdiag * suma / s + 0.5 * tr;
        const double elbo = trace - 0.5 * ND * log(2.0 * 3.14159265358979323846 * s) - 0.5 * (double)D * ws[SLy.scal + SS_LOGDET] -
                            0.5 * y2 / s + 0.5 * c2 - P.kl_weight * ws[SLy.scal + SS_KL];
        double* o = sub_out + (size_t)slot * 8;
        o[0] = elbo;
        o[1] = gl;
        o[2] = gv - 0.5 * suma / s;
        o[3] = -c2 / s - adel / (s * s) - 0.5 * ND / s + 0.5 * y2 / (s * s);
        o[4] = gb;
    }
}

// sum the SD sub-problems of every item of the wave
__global__ void k_sp_reduce(SparseParams P, const double* sub_out, int nitems) {
    const int li = blockIdx.x * blockDim.x + threadIdx.x;
    if (li >= nitems) return;
    const int item = P.item0 + li;
    double e = 0.0, gl = 0.0, gv = 0.0, gs = 0.0;
    for (int sd = 0; sd < P.SD; ++sd) {
        const double* o = sub_out + ((size_t)li * P.SD + sd) * 8;
        e += o[0]; gl += o[1]; gv += o[2]; gs += o[3];
        if (P.grad_b) P.grad_b[(size_t)item * P.SD + sd] = o[4];
    }
    P.elbo[item] = e;
    if (P.grad_hyp) {
        P.grad_hyp[(size_t)item * 3 + 0] = gl;
        P.grad_hyp[(size_t)item * 3 + 1] = gv;
        P.grad_hyp[(size_t)item * 3 + 2] = gs;
    }
}

// ---------------------------------------------------------------------------------------------------- k_sp_predict
// AssignGPSparse.predict_f (assigngp_denseSparse.py:109-155) for one sub-problem whose forward state (Linv, Rinv, c) is in
// the workspace.  scratch: 3 * Mz * T doubles.  Xnew[T][2] = (time, function).  full_cov: var_out is [T][T], else [T].
__device__ __forceinline__ double k_pair(const SparseParams& P, double t1, int f1, double t2, int f2, double ell, double kvar,
                                         double kinv, double bpt) {
    if (f1 == f2) return kbase(t1 - t2, ell, kvar, P.kind);
    return (kbase(t1 - bpt, ell, kvar, P.kind) * kinv) * kbase(t2 - bpt, ell, kvar, P.kind);
}

__global__ void __launch_bounds__(NTHREADS) k_sp_predict(SparseParams P, const double* Xnew, int T, int full_cov, double* scratch,
                                                         double* mean_out, double* var_out) {
    SP_SETUP();
    double* Kus = scratch;
    double* t1 = scratch + (size_t)Mz * T;
    double* t2 = t1 + (size_t)Mz * T;
    const double* Linv = ws + SLy.Linv;
    const double* Rinv = ws + SLy.Rinv;
    for (int e = tid; e < Mz * T; e += NTHREADS) {
        const int m = e / T, s = e % T;
        Kus[e] = k_pair(P, P.Z[2 * m], (int)P.Z[2 * m + 1], Xnew[2 * s], (int)Xnew[2 * s + 1], ell, kvar, kinv, bpt);
    }
    __syncthreads();
    for (int e = tid; e < Mz * T; e += NTHREADS) {
        const int m = e / T, s = e % T;
        double a = 0.0;
        for (int k = 0; k <= m; ++k) a += Linv[(size_t)m * Mz + k] * Kus[(size_t)k * T + s];
        t1[e] = a;
    }
    __syncthreads();
    for (int e = tid; e < Mz * T; e += NTHREADS) {
        const int m = e / T, s = e % T;
        double a = 0.0;
        for (int k = 0; k <= m; ++k) a += Rinv[(size_t)m * Mz + k] * t1[(size_t)k * T + s];
        t2[e] = a;
    }
    __syncthreads();
    for (int e = tid; e < T * D; e += NTHREADS) {
        const int s = e / D, d = e % D;
        double a = 0.0;
        for (int m = 0; m < Mz; ++m) a += t2[(size_t)m * T + s] * ws[SLy.c + (size_t)d * Mz + m];
        mean_out[e] = a;
    }
    if (!full_cov) {
        for (int s = tid; s < T; s += NTHREADS) {
            double a = kvar + P.square_noise;
            for (int m = 0; m < Mz; ++m) { const double x = t2[(size_t)m * T + s], y = t1[(size_t)m * T + s]; a += x * x - y * y; }
            var_out[s] = a;
        }
    } else {
        for (int e = tid; e < T * T; e += NTHREADS) {
            const int s = e / T, r = e % T;
            double a = k_pair(P, Xnew[2 * s], (int)Xnew[2 * s + 1], Xnew[2 * r], (int)Xnew[2 * r + 1], ell, kvar, kinv, bpt);
            if (s == r) a += P.square_noise;
            for (int m = 0; m < Mz; ++m) a += t2[(size_t)m * T + s] * t2[(size_t)m * T + r] - t1[(size_t)m * T + s] * t1[(size_t)m * T + r];
            var_out[e] = a;
        }
    }
}

void sparse_launch_predict(const SparseParams& P, const double* Xnew, int T, int full_cov, double* scratch, double* mean_out,
                           double* var_out, cudaStream_t st) {
    k_sp_predict<<<dim3(1, 1), NTHREADS, 0, st>>>(P, Xnew, T, full_cov, scratch, mean_out, var_out);
}

// ---------------------------------------------------------------------------------------------------- launchers
static size_t packed_bytes(int Mz) {   // packed lower triangle + inverse diagonal blocks + 8 x 9 scratch per warp
    return ((size_t)Mz * (Mz + 1) / 2 + 64 * (size_t)((Mz + 7) / 8) + 72 * (SP_T / 32)) * 8;
}
// slab buffers + Delta + staged vectors (k_sp_fwd: 2 per-row vectors + fz; k_sp_bwd: 4 + fz + zbar; 4 per-cell vectors)
static size_t slab_bytes(int Mz, int nbuf, int D) { return ((size_t)nbuf * Mz * SP_LDS + SP_SL + 5 * (size_t)Mz + (size_t)(D <= SP_ZBAR_D ? D : 0) * Mz + 4 * SP_CELLS + SP_ZBAR_D * SP_SL + 8) * 8; }

cudaError_t sparse_init_kernels() {
    cudaError_t e;
    const int big = 200 * 1024;
    if ((e = cudaFuncSetAttribute((const void*)k_sp_kuu, cudaFuncAttributeMaxDynamicSharedMemorySize, big)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute((const void*)k_sp_mid, cudaFuncAttributeMaxDynamicSharedMemorySize, big)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute((const void*)k_sp_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, big)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute((const void*)k_sp_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, big)) != cudaSuccess) return e;
    return cudaSuccess;
}

// phase: 0 prep, 1 Kuu, 2 forward slabs, 3 mid, 4 backward slabs, 5 end + reduce
void sparse_launch_phase(const SparseParams& P, int nslots, int phase, double* sub_out, cudaStream_t st) {
    dim3 g1(1, nslots), gs(P.nsplit, nslots);
    switch (phase) {
        case 0:
            k_sp_reduce_partials<<<dim3((P.Mp + NTHREADS - 1) / NTHREADS, nslots), NTHREADS, 0, st>>>(P);
            k_sp_prep<<<g1, NTHREADS, 0, st>>>(P);
            break;
        case 1: k_sp_kuu<<<g1, SP_T, packed_bytes(P.Mz), st>>>(P); break;
        case 2: k_sp_fwd<<<gs, SP_T, slab_bytes(P.Mz, 2, P.D), st>>>(P); break;
        case 3: k_sp_mid<<<g1, SP_T, packed_bytes(P.Mz), st>>>(P); break;
        case 4: k_sp_bwd<<<gs, SP_T, slab_bytes(P.Mz, 3, P.D), st>>>(P); break;
        case 5: {
            k_sp_end<<<g1, SP_END_THREADS, 0, st>>>(P, sub_out);
            const int nitems = nslots / P.SD;
            k_sp_reduce<<<(nitems + 127) / 128, 128, 0, st>>>(P, sub_out, nitems);
            break;
        }
        default: break;
    }
}

}  // namespace bgp
