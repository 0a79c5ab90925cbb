// Dense AssignGP bound + gradient, one CTA per work item (gene x branching point), sm_100a.
//
// Restates assigngp_dense.py:151-199 (forward) and the hand-derived adjoint of SURVEY.md App. B.1 as a chain of
// kernels that all share the tile GEMM core of bgp_gemm.cuh.  A work item owns one workspace slot; kernels of a
// wave are launched back to back on one stream with grid = items in the wave:
//
//   k_prep    A, Delta, v = Phi^T Y (from the Phi pass partials), k(t_i, b) vectors
//   k_chol<0> Lc = chol(K + 1e-6 I), K generated in the epilogue (branch_kernParamGPflow.py:117-198), left-looking
//   k_syrk    P = (Lc + eps I)^T Delta (Lc + eps I) + I                      (assigngp_dense.py:165-172)
//   k_chol<1> R = chol(P), log-det                                            (:173)
//   k_solve   z = L^T v, c = R^-1 z / s, q = R^-T c                           (:174-181)
//   k_trtri   X = R^-1 in place;  k_lauum  Pbar = -D/2 X^T X - q q^T / 2
//   k_trmm    G = L Pbar (lower), Lbar = 2 Delta G + v zbar^T, delta = rowsum(G o L)
//   k_post    Abar, vbar, d/d likvar, ELBO
//   k_adj     gather form of the blocked Cholesky adjoint, fused with sum(Ktilde o dK/dtheta)
#include "bgp_dense_dev.cuh"
#include "bgp_launch.h"

namespace bgp {

// ------------------------------------------------------------------------------------------------ k_prep
struct PrepArgs {
    const double* Apart;      // [slot][nch][Mp]
    const double* vpart;      // [slot][nch][D][Mp]
    const double* klpart;     // [slot][nch]
    const double* Y;          // [nY][N][D]
    const int* item_gene;     // [count]
};

__global__ void __launch_bounds__(NTHREADS) k_prep(DenseParams P, PrepArgs Q) {
    DENSE_SETUP();
    __shared__ double red[32];
    const int Mp = dm.Mp, M = dm.M, D = dm.D, nch = dm.nch, N = dm.N;
    const double b = P.bv[item];
    for (int j = tid; j < Mp; j += NTHREADS) {
        double a = 0.0;
        if (j < M)
            for (int ch = 0; ch < nch; ++ch) a += Q.Apart[(size_t)slot * P.slot_stride + (size_t)ch * Mp + j];
        vec[VL.A + j] = a;
        vec[VL.Delta + j] = a / likvar;
        for (int d = 0; d < D; ++d) {
            double s = 0.0;
            if (j < M)
                for (int ch = 0; ch < nch; ++ch) s += Q.vpart[(size_t)slot * P.slot_stride + ((size_t)ch * D + d) * Mp + j];
            vec[VL.v + (size_t)d * Mp + j] = s;
        }
        // per-cell kernel vectors against the branching point (cross-covariance factors and their derivatives)
        double k = 0.0, dl = 0.0, dd = 0.0;
        if (j < N) kbase_grads(P.t[j] - b, ell, kvar, P.kind, k, dl, dd);
        vec[VL.ka + j] = k;
        vec[VL.dkal + j] = dl;
        vec[VL.dkab + j] = -dd;   // d/db k(t - b) = -k'(t - b)
    }
    double kl = 0.0;
    for (int ch = tid; ch < nch; ch += NTHREADS) kl += Q.klpart[(size_t)slot * P.slot_stride + ch];
    kl = block_sum(kl, red);
    const double* Yg = Q.Y + (size_t)Q.item_gene[item] * N * D;
    double y2 = 0.0;
    for (int e = tid; e < N * D; e += NTHREADS) y2 += Yg[e] * Yg[e];
    y2 = block_sum(y2, red);
    if (tid == 0) {
        vec[VL.scal + SC_KL] = kl;
        vec[VL.scal + SC_SUMY2] = y2;
        vec[VL.scal + SC_LOGDET] = 0.0;
        P.status[item] = 0;
    }
}

// Base-kernel table: every same-function entry K[(i,f),(j,f)] = k(t_i - t_j) is needed 3 times by the covariance epilogue
// and 3 times by the hyper-gradient; one exp per cell pair here replaces one per matrix entry there.
__global__ void __launch_bounds__(NTHREADS) k_ktable(DenseParams P) {
    const int slot = blockIdx.y, item = P.item0 + slot;
    const int N = P.d.N;
    const double ell = P.hyp[(size_t)item * 3 + 0], kvar = P.hyp[(size_t)item * 3 + 1];
    double* KT = P.KT + (size_t)slot * P.slot_stride;
    const long long nn = (long long)N * N;
    for (long long e = (long long)blockIdx.x * NTHREADS + threadIdx.x; e < nn; e += (long long)gridDim.x * NTHREADS) {
        const int i = (int)(e / N), j = (int)(e % N);
        double k, dl, dd;
        kbase_grads(P.t[i] - P.t[j], ell, kvar, P.kind, k, dl, dd);
        KT[e] = k;
        KT[nn + e] = dl;
    }
}

// ------------------------------------------------------------------------------------------------ k_chol
// SRC == 0: factor K (generated on the fly) into LC, also LCe and LinvD.   SRC == 1: factor P (stored in RX) in place, RinvD, log-det.
template <int SRC>
__global__ void __launch_bounds__(GEMM_THREADS, 1) k_chol(DenseParams P) {
    DENSE_SETUP();
    Pipe pipe;
    pipe_init(pipe, smem);
    pipe.row_limit = dm.nTv;   // tile rows beyond the last one that holds a row < M are identity padding
    double* red = smem_red(smem);
    double* Dm = pipe.stages;
    double* dst = SRC == 0 ? LC : RX;
    double* invD = SRC == 0 ? LinvD : RinvD;
    const double* ka = vec + VL.ka;
    const double* KT = P.KT + (size_t)slot * P.slot_stride;
    const double kinv = 1.0 / (kvar + JITTER);   // inv(k(b,b) + jitter), branch_kernParamGPflow.py:168-189
    const int nM = dm.nM;
    int fail = 0;
    double logdet = 0.0;
    const Opnd opL = opnd_packed(dst, ACC_ROWK);
    Acc acc;
    for (int J = 0; J < nM; ++J) {
        // ---- diagonal macro block
        acc_zero(acc);
        gemm_macro(acc, pipe, opL, TPM * J, opL, TPM * J, 0, TPM * J, nullptr, 2);   // only the lower warp tiles
        __syncthreads();   // the ring doubles as the 128x128 scratch from here on
        acc_foreach(acc, J, J, [&](int it, int jt, int i, int j, double v0, double v1) {
            if (it < jt) return;
            const int gi = it * TS + i, gj = jt * TS + j;
            double s0, s1;
            if (SRC == 0) {
                s0 = kgen(P, ka, KT, gi, gj, kinv);
                s1 = kgen(P, ka, KT, gi, gj + 1, kinv);
            } else {
                const double2 p = tile_load2(dst + packed_tile(it, jt) * TILE_ELEMS, i, j);
                s0 = p.x; s1 = p.y;
            }
            double* o = Dm + (gi - MB * J) * DLD + (gj - MB * J);
            o[0] = s0 - v0;
            o[1] = s1 - v1;
        }, 2);
        __syncthreads();
#ifdef BGP_TIMING
        const long long tc0 = clock64();
#endif
        chol128(Dm, fail, MB * J);
#ifdef BGP_TIMING
        if (tid == 0) atomicAdd(&g_gemm_timers[6], (unsigned long long)(clock64() - tc0));
#endif
        if (SRC == 1) {
            double l = 0.0;
            if (tid < MB) { const double r = Dm[tid * DLD + tid]; l = log(r * r); }
            logdet += block_sum(l, red);
        }
        store_lower_block(Dm, J, dst, nullptr, SRC == 0 ? LCe : nullptr, JITTER);
        __syncthreads();
#ifdef BGP_TIMING
        const long long tc1 = clock64();
#endif
        trinv128(Dm);
#ifdef BGP_TIMING
        if (tid == 0) atomicAdd(&g_gemm_timers[7], (unsigned long long)(clock64() - tc1));
#endif
        store_lower_block(Dm, J, nullptr, invD + (size_t)J * 16 * TILE_ELEMS, nullptr, 0.0);
        fence_proxy_async();
        __syncthreads();
        // ---- panel below the diagonal block
        const Opnd opInv = opnd_block(invD + (size_t)J * 16 * TILE_ELEMS, ACC_ROWK, TPM * J, TPM * J, 1);
        for (int I = J + 1; I < nM; ++I) {
            acc_zero(acc);
            gemm_macro(acc, pipe, opL, TPM * I, opL, TPM * J, 0, TPM * J, nullptr);
            acc_foreach(acc, I, J, [&](int it, int jt, int i, int j, double v0, double v1) {
                double* tp = dst + packed_tile(it, jt) * TILE_ELEMS;
                double s0, s1;
                if (SRC == 0) {
                    const int gi = it * TS + i, gj = jt * TS + j;
                    s0 = kgen(P, ka, KT, gi, gj, kinv);
                    s1 = kgen(P, ka, KT, gi, gj + 1, kinv);
                } else {
                    const double2 p = tile_load2(tp, i, j);
                    s0 = p.x; s1 = p.y;
                }
                tile_store2(tp, i, j, s0 - v0, s1 - v1);
            });
            fence_proxy_async();
            // X = C * inv(L_JJ)^T, in place
            acc_zero(acc);
            gemm_macro(acc, pipe, opL, TPM * I, opInv, TPM * J, TPM * J, TPM * J + TPM, nullptr, 0);   // the skipped operand is B
            acc_foreach(acc, I, J, [&](int it, int jt, int i, int j, double v0, double v1) {
                tile_store2(dst + packed_tile(it, jt) * TILE_ELEMS, i, j, v0, v1);
            }, 0);
            fence_proxy_async();
        }
    }
    if (tid == 0) {
        if (SRC == 1) vec[VL.scal + SC_LOGDET] = logdet;
        if (fail != 0 && P.status[item] == 0) P.status[item] = fail + (SRC == 1 ? dm.Mp : 0);
    }
}

// ------------------------------------------------------------------------------------------------ k_syrk:  P = L^T Delta L + I  (L = Lc + eps I)
__global__ void __launch_bounds__(GEMM_THREADS, 1) k_syrk(DenseParams P) {
    DENSE_SETUP();
    Pipe pipe;
    pipe_init(pipe, smem);
    pipe.row_limit = dm.nTv;   // tile rows beyond the last one that holds a row < M are identity padding
    const Opnd opL = opnd_packed(LC, ACC_COLK, LCe);
    const double* Delta = vec + VL.Delta;
    Acc acc;
    for (int I = 0; I < dm.nM; ++I)
        for (int J = 0; J <= I; ++J) {
            acc_zero(acc);
            const int rowmap = (I == J) ? 2 : 1;
            gemm_macro(acc, pipe, opL, TPM * I, opL, TPM * J, TPM * I, dm.nTv, Delta, rowmap);
            acc_foreach(acc, I, J, [&](int it, int jt, int i, int j, double v0, double v1) {
                if (it < jt) return;
                const int gi = it * TS + i, gj = jt * TS + j;
                if (gi == gj) v0 += 1.0;
                if (gi == gj + 1) v1 += 1.0;
                tile_store2(RX + packed_tile(it, jt) * TILE_ELEMS, i, j, v0, v1);
            }, rowmap);
        }
}

// ------------------------------------------------------------------------------------------------ k_solve: z, c, q
constexpr int SOLVE_THREADS = 1024;   // memory-bound vector kernel: many warps = many tile loads in flight
constexpr int SOLVE_WARPS = SOLVE_THREADS / 32;
constexpr int SOLVE_NH = SOLVE_WARPS / 4;   // warps sharing one 32-row (or 32-column) strip of a macro block

__global__ void __launch_bounds__(SOLVE_THREADS) k_solve(DenseParams P) {
    DENSE_SETUP();
    __shared__ double part[SOLVE_THREADS];
    __shared__ double yv[MB];
    __shared__ double red[32];
    const int warp = tid >> 5, lane = tid & 31;
    const int Mp = dm.Mp, nT = dm.nT, nM = dm.nM;
    double sumc2 = 0.0;
    for (int d = 0; d < dm.D; ++d) {
        const double* v = vec + VL.v + (size_t)d * Mp;
        double* z = vec + VL.z + (size_t)d * Mp;
        double* cv = vec + VL.c + (size_t)d * Mp;
        double* q = vec + VL.q + (size_t)d * Mp;
        // z = (Lc + eps I)^T v : z[j] = sum_{i >= j} L[i][j] v[i]; long and short tile columns are interleaved over the warps
        for (int idx = warp; idx < nT; idx += SOLVE_WARPS) {
            const int jt = (idx & 1) ? (nT - 1 - (idx >> 1)) : (idx >> 1);
            double s = 0.0;
            for (int it = jt; it < nT; ++it) {
                const double* tile = (it == jt) ? LCe + (size_t)it * TILE_ELEMS : LC + packed_tile(it, jt) * TILE_ELEMS;
                const double* vv = v + it * TS;
#pragma unroll
                for (int i = 0; i < TS; ++i) s += tile[tswz(i, lane)] * vv[i];
            }
            z[jt * TS + lane] = s;
        }
        __syncthreads();
        // forward substitution by macro block rows: cu_J = inv(R_JJ) (z_J - sum_{k<J} R[J,k] cu_k)   (cu = R^-1 z)
        for (int J = 0; J < nM; ++J) {
            {
                const int a = warp & 3, h = warp >> 2;
                const int it = TPM * J + a;
                double s = 0.0;
                for (int kt = h; kt < TPM * J; kt += SOLVE_NH) {
                    const double* row = RX + packed_tile(it, kt) * TILE_ELEMS + lane * TS;
                    const double* cc = cv + kt * TS;
                    const int sw = (lane & 3) << 2;
#pragma unroll
                    for (int k = 0; k < TS; ++k) s += row[k ^ sw] * cc[k];
                }
                part[tid] = s;
            }
            __syncthreads();
            if (tid < MB) {
                double s = z[MB * J + tid];
#pragma unroll
                for (int h = 0; h < SOLVE_NH; ++h) s -= part[tid + MB * h];
                yv[tid] = s;
            }
            __syncthreads();
            if (tid < MB) {
                const double* blk = RinvD + (size_t)J * 16 * TILE_ELEMS;
                const int a = tid >> 5, i = tid & 31;
                double s = 0.0;
                for (int b = 0; b <= a; ++b) {
                    const double* row = blk + (size_t)(a * TPM + b) * TILE_ELEMS + i * TS;
                    const int sw = (i & 3) << 2;
#pragma unroll
                    for (int k = 0; k < TS; ++k) s += row[k ^ sw] * yv[b * TS + k];
                }
                cv[MB * J + tid] = s;
            }
            __syncthreads();
        }
        // c = cu / s
        for (int j = tid; j < Mp; j += SOLVE_THREADS) {
            const double x = cv[j] / likvar;
            cv[j] = x;
            sumc2 += x * x;
        }
        __syncthreads();
        // backward substitution: q_J = inv(R_JJ)^T (c_J - sum_{I>J} R[I,J]^T q_I)
        for (int J = nM - 1; J >= 0; --J) {
            {
                const int b = warp & 3, h = warp >> 2;
                const int jt = TPM * J + b;
                double s = 0.0;
                for (int it = TPM * (J + 1) + h; it < nT; it += SOLVE_NH) {
                    const double* tile = RX + packed_tile(it, jt) * TILE_ELEMS;
                    const double* qq = q + it * TS;
#pragma unroll
                    for (int i = 0; i < TS; ++i) s += tile[tswz(i, lane)] * qq[i];
                }
                part[tid] = s;
            }
            __syncthreads();
            if (tid < MB) {
                double s = cv[MB * J + tid];
#pragma unroll
                for (int h = 0; h < SOLVE_NH; ++h) s -= part[tid + MB * h];
                yv[tid] = s;
            }
            __syncthreads();
            if (tid < MB) {
                const double* blk = RinvD + (size_t)J * 16 * TILE_ELEMS;
                const int b = tid >> 5, j = tid & 31;
                double s = 0.0;
                for (int a = b; a < TPM; ++a) {
                    const double* tile = blk + (size_t)(a * TPM + b) * TILE_ELEMS;
#pragma unroll
                    for (int k = 0; k < TS; ++k) s += tile[tswz(k, j)] * yv[a * TS + k];
                }
                q[MB * J + tid] = s;
            }
            __syncthreads();
        }
    }
    sumc2 = block_sum(sumc2, red);
    if (tid == 0) vec[VL.scal + SC_SUMC2] = sumc2;
}

// ------------------------------------------------------------------------------------------------ k_trtri:  RX <- inv(R), in place
__global__ void __launch_bounds__(GEMM_THREADS, 1) k_trtri(DenseParams P) {
    DENSE_SETUP();
    Pipe pipe;
    pipe_init(pipe, smem);
    pipe.row_limit = dm.nTv;   // tile rows beyond the last one that holds a row < M are identity padding
    const int nM = dm.nM;
    const Opnd opRow = opnd_packed(RX, ACC_ROWK);
    const Opnd opCol = opnd_packed(RX, ACC_COLK);
    Acc acc;
    for (int J = nM - 1; J >= 0; --J) {
        const double* blk = RinvD + (size_t)J * 16 * TILE_ELEMS;
        const Opnd opInv = opnd_block(blk, ACC_COLK, TPM * J, TPM * J, 1);
        for (int I = nM - 1; I > J; --I) {
            // acc = sum_{J<k<=I} X[I,k] R[k,J]
            acc_zero(acc);
            gemm_macro(acc, pipe, opRow, TPM * I, opCol, TPM * J, TPM * (J + 1), min(TPM * I + TPM, dm.nTv), nullptr);
            acc_foreach(acc, I, J, [&](int it, int jt, int i, int j, double v0, double v1) {
                tile_store2(RX + packed_tile(it, jt) * TILE_ELEMS, i, j, v0, v1);
            });
            fence_proxy_async();
            // X[I,J] = -acc * inv(R_JJ)
            acc_zero(acc);
            gemm_macro(acc, pipe, opRow, TPM * I, opInv, TPM * J, TPM * J, TPM * J + TPM, nullptr, 0);   // the skipped operand is B
            acc_foreach(acc, I, J, [&](int it, int jt, int i, int j, double v0, double v1) {
                tile_store2(RX + packed_tile(it, jt) * TILE_ELEMS, i, j, -v0, -v1);
            }, 0);
            fence_proxy_async();
        }
        // X[J,J] = inv(R_JJ)
        __syncthreads();
        for (int a = 0; a < TPM; ++a)
            for (int b = 0; b <= a; ++b) {
                const double* src = blk + (size_t)(a * TPM + b) * TILE_ELEMS;
                double* dstp = RX + packed_tile(TPM * J + a, TPM * J + b) * TILE_ELEMS;
                for (int e = tid; e < TILE_ELEMS; e += GEMM_THREADS) dstp[e] = src[e];
            }
        fence_proxy_async();
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------ k_lauum:  PB = -D/2 X^T X - 1/2 q q^T
__global__ void __launch_bounds__(GEMM_THREADS, 1) k_lauum(DenseParams P) {
    DENSE_SETUP();
    Pipe pipe;
    pipe_init(pipe, smem);
    pipe.row_limit = dm.nTv;   // tile rows beyond the last one that holds a row < M are identity padding
    const Opnd opX = opnd_packed(RX, ACC_COLK);
    const double* q = vec + VL.q;
    const int Mp = dm.Mp, D = dm.D;
    const double hD = 0.5 * (double)D;
    Acc acc;
    for (int I = 0; I < dm.nM; ++I)
        for (int J = 0; J <= I; ++J) {
            acc_zero(acc);
            const int rowmap = (I == J) ? 2 : 1;
            gemm_macro(acc, pipe, opX, TPM * I, opX, TPM * J, TPM * I, dm.nTv, nullptr, rowmap);
            acc_foreach(acc, I, J, [&](int it, int jt, int i, int j, double v0, double v1) {
                if (it < jt) return;
                const int gi = it * TS + i, gj = jt * TS + j;
                double qq0 = 0.0, qq1 = 0.0;
                for (int d = 0; d < D; ++d) {
                    const double qi = q[(size_t)d * Mp + gi];
                    qq0 += qi * q[(size_t)d * Mp + gj];
                    qq1 += qi * q[(size_t)d * Mp + gj + 1];
                }
                tile_store2(PB + packed_tile(it, jt) * TILE_ELEMS, i, j, -hD * v0 - 0.5 * qq0, -hD * v1 - 0.5 * qq1);
            }, rowmap);
        }
    pipe_report(pipe);
}

// ------------------------------------------------------------------------------------------------ k_trmm:  G = L Pbar (lower), Lbar, delta partials
__global__ void __launch_bounds__(GEMM_THREADS, 1) k_trmm(DenseParams P) {
    DENSE_SETUP();
    Pipe pipe;
    pipe_init(pipe, smem);
    pipe.row_limit = dm.nTv;   // tile rows beyond the last one that holds a row < M are identity padding
    double* red = smem_red(smem);
    const Opnd opL = opnd_packed(LC, ACC_ROWK, LCe);
    const Opnd opP = opnd_packed(PB, ACC_SYM);
    const double* Delta = vec + VL.Delta;
    const double* v = vec + VL.v;
    const double* q = vec + VL.q;
    double* dpart = vec + VL.dpart;
    const int Mp = dm.Mp, D = dm.D;
    const int warp = tid >> 5, lane = tid & 31, g = lane >> 2, c = lane & 3;
    Acc acc;
    for (int I = 0; I < dm.nM; ++I)
        for (int J = 0; J <= I; ++J) {
            acc_zero(acc);
            const int rowmap = (I == J) ? 2 : 1;
            gemm_macro(acc, pipe, opL, TPM * I, opP, TPM * J, 0, min(TPM * I + TPM, dm.nTv), nullptr, rowmap);
            double rs[4] = {0.0, 0.0, 0.0, 0.0};
            acc_foreach(acc, I, J, [&](int it, int jt, int i, int j, double v0, double v1) {
                if (it < jt) return;
                const int gi = it * TS + i, gj = jt * TS + j;
                const double* lt = (it == jt) ? LCe + (size_t)it * TILE_ELEMS : LC + packed_tile(it, jt) * TILE_ELEMS;
                const double2 l = tile_load2(lt, i, j);
                double vz0 = 0.0, vz1 = 0.0;
                for (int d = 0; d < D; ++d) {
                    const double vi = v[(size_t)d * Mp + gi];
                    vz0 += vi * q[(size_t)d * Mp + gj];
                    vz1 += vi * q[(size_t)d * Mp + gj + 1];
                }
                const double dl = 2.0 * Delta[gi];
                double o0 = 0.0, o1 = 0.0, r = 0.0;
                if (gi >= gj) { o0 = dl * v0 + vz0 / likvar; r += v0 * l.x; }
                if (gi >= gj + 1) { o1 = dl * v1 + vz1 / likvar; r += v1 * l.y; }
                rs[i >> 3] += r;
                tile_store2(LB + packed_tile(it, jt) * TILE_ELEMS, i, j, o0, o1);
            }, rowmap);
            // delta partial of this macro tile: row sums over its 128 columns
            if (rowmap == 2)   // upper warp tiles are not computed on diagonal macro tiles: their partials are zero
                for (int e = tid; e < WCOLS * MB; e += GEMM_THREADS) red[e] = 0.0;
            __syncthreads();
            {
                int wa_, wb_;
                bool idle_;
                warp_tile(warp, rowmap, wa_, wb_, idle_);
#pragma unroll
                for (int mi = 0; mi < 4; ++mi) {
                    double r = rs[mi];
                    r += __shfl_xor_sync(0xffffffffu, r, 1);
                    r += __shfl_xor_sync(0xffffffffu, r, 2);
                    if (c == 0 && !idle_) red[wb_ * MB + wa_ * TS + 8 * mi + g] = r;   // one partial per column group of warps
                }
            }
            __syncthreads();
            if (tid < MB) {
                double r = 0.0;
#pragma unroll
                for (int w = 0; w < WCOLS; ++w) r += red[w * MB + tid];
                dpart[(size_t)J * Mp + MB * I + tid] = r;
            }
        }
}

// ------------------------------------------------------------------------------------------------ k_post: delta, Abar, vbar, d/d likvar, ELBO
__global__ void __launch_bounds__(SOLVE_THREADS) k_post(DenseParams P) {
    DENSE_SETUP();
    __shared__ double red[32];
    const int warp = tid >> 5, lane = tid & 31;
    const int Mp = dm.Mp, M = dm.M, nT = dm.nT, D = dm.D, N = dm.N;
    double adel = 0.0;
    for (int i = tid; i < Mp; i += SOLVE_THREADS) {
        const int I = i / MB;
        double s = 0.0;
        for (int J = 0; J <= I; ++J) s += vec[VL.dpart + (size_t)J * Mp + i];
        vec[VL.delta + i] = s;
        vec[VL.Abar + i] = s / likvar;
        if (i < M) adel += vec[VL.A + i] * s;
    }
    adel = block_sum(adel, red);
    // vbar = L zbar,  zbar = q / s :  vbar[i] = sum_{j <= i} L[i][j] q[j] / s
    for (int d = 0; d < D; ++d) {
        const double* q = vec + VL.q + (size_t)d * Mp;
        double* vb = vec + VL.vbar + (size_t)d * Mp;
        for (int idx = warp; idx < nT; idx += SOLVE_WARPS) {
            const int it = (idx & 1) ? (nT - 1 - (idx >> 1)) : (idx >> 1);   // interleave long and short tile rows
            double s = 0.0;
            const int sw = (lane & 3) << 2;
            for (int jt = 0; jt <= it; ++jt) {
                const double* row = ((it == jt) ? LCe + (size_t)it * TILE_ELEMS : LC + packed_tile(it, jt) * TILE_ELEMS) + lane * TS;
                const double* qq = q + jt * TS;
#pragma unroll
                for (int k = 0; k < TS; ++k) s += row[k ^ sw] * qq[k];
            }
            vb[it * TS + lane] = s / likvar;
        }
    }
    if (tid == 0) {
        const double s = likvar;
        const double ND = (double)N * (double)D;
        const double logdet = vec[VL.scal + SC_LOGDET], c2 = vec[VL.scal + SC_SUMC2], kl = vec[VL.scal + SC_KL],
                     y2 = vec[VL.scal + SC_SUMY2];
        // assigngp_dense.py:184-199
        const double a1 = -0.5 * ND * log(2.0 * 3.14159265358979323846 * s);
        const double a2 = -0.5 * (double)D * logdet;
        const double a3 = -0.5 * y2 / s;
        const double a4 = 0.5 * c2;
        P.elbo[item] = a1 + a2 + a3 + a4 - P.kl_weight * kl;
        P.grad_hyp[(size_t)item * 4 + 2] = -0.5 * ND / s + 0.5 * y2 / (s * s) - c2 / s - adel / (s * s);
    }
}

// ------------------------------------------------------------------------------------------------ k_adj
__global__ void __launch_bounds__(GEMM_THREADS, 1) k_adj(DenseParams P) {
    DENSE_SETUP();
    Pipe pipe;
    pipe_init(pipe, smem);
    pipe.row_limit = dm.nTv;   // tile rows beyond the last one that holds a row < M are identity padding
    double* red = smem_red(smem);
    const int nM = dm.nM, nT = dm.nTv;   // k-ranges stop at the last tile that holds a row < M
    const double* ka = vec + VL.ka;
    const double* dkal = vec + VL.dkal;
    const double* dkab = vec + VL.dkab;
    const double* KT = P.KT + (size_t)slot * P.slot_stride;
    const double kinv = 1.0 / (kvar + JITTER);
    const bool want_hyp = (P.KConst == nullptr);
    double gl = 0.0, gv = 0.0, gb = 0.0;
    const Opnd opAsym = opnd_packed(LB, ACC_SYM);
    const Opnd opArow = opnd_packed(LB, ACC_ROWK);
    const Opnd opAcol = opnd_packed(LB, ACC_COLK);
    const Opnd opLcol = opnd_packed(LC, ACC_COLK);
    double* S0 = SCR;
    double* S1 = SCR + 16 * TILE_ELEMS;
    Acc acc;
    for (int J = nM - 1; J >= 0; --J) {
        const double* blk = LinvD + (size_t)J * 16 * TILE_ELEMS;
        const Opnd opInvCol = opnd_block(blk, ACC_COLK, TPM * J, TPM * J, 1);
        for (int I = nM - 1; I > J; --I) {
            // E = Lbar[I,J] - sum_{k>J} Atilde_full[I,k] Lc[k,J]
            acc_zero(acc);
            gemm_macro(acc, pipe, opAsym, TPM * I, opLcol, TPM * J, TPM * (J + 1), nT, nullptr);
            acc_foreach(acc, I, J, [&](int it, int jt, int i, int j, double v0, double v1) {
                double* tp = LB + packed_tile(it, jt) * TILE_ELEMS;
                const double2 l = tile_load2(tp, i, j);
                tile_store2(tp, i, j, l.x - v0, l.y - v1);
            });
            fence_proxy_async();
            // Atilde[I,J] = E inv(L_JJ)
            acc_zero(acc);
            gemm_macro(acc, pipe, opArow, TPM * I, opInvCol, TPM * J, TPM * J, TPM * J + TPM, nullptr, 0);   // the skipped operand is B
            acc_foreach(acc, I, J, [&](int it, int jt, int i, int j, double v0, double v1) {
                tile_store2(LB + packed_tile(it, jt) * TILE_ELEMS, i, j, v0, v1);
                if (want_hyp) {
                    const int gi = it * TS + i, gj = jt * TS + j;
                    hyper_accum(P, ka, dkal, dkab, KT, gi, gj, v0, kvar, kinv, gl, gv, gb);
                    hyper_accum(P, ka, dkal, dkab, KT, gi, gj + 1, v1, kvar, kinv, gl, gv, gb);
                }
            }, 0);
            fence_proxy_async();
        }
        // ---- diagonal block:  Leff = tril(Lbar_JJ) - tril(sum_{I>J} Atilde[I,J]^T Lc[I,J])
        acc_zero(acc);
        // only the lower triangle of this block is used (Leff is tril): the ten lower warp tiles, 3/4 of the time of a full macro tile;
        // the upper tiles of S0 are never read (the product below takes S0 as a lower-triangular operand)
        gemm_macro(acc, pipe, opAcol, TPM * J, opLcol, TPM * J, TPM * (J + 1), nT, nullptr, 2);
        acc_foreach(acc, J, J, [&](int it, int jt, int i, int j, double v0, double v1) {
            if (it < jt) return;
            const int a = it - TPM * J, b = jt - TPM * J;
            const double2 l = tile_load2(LB + packed_tile(it, jt) * TILE_ELEMS, i, j);
            const int gi = it * TS + i, gj = jt * TS + j;
            const double o0 = (gi >= gj) ? l.x - v0 : 0.0;
            const double o1 = (gi >= gj + 1) ? l.y - v1 : 0.0;
            tile_store2(S0 + (size_t)(a * TPM + b) * TILE_ELEMS, i, j, o0, o1);
        }, 2);
        fence_proxy_async();
        // U = Phi(L_JJ^T Leff)
        acc_zero(acc);
        gemm_macro(acc, pipe, opLcol, TPM * J, opnd_block(S0, ACC_COLK, TPM * J, TPM * J, 1), TPM * J, TPM * J, TPM * J + TPM, nullptr);
        acc_foreach(acc, J, J, [&](int it, int jt, int i, int j, double v0, double v1) {
            const int a = it - TPM * J, b = jt - TPM * J;
            const int gi = it * TS + i, gj = jt * TS + j;
            const double o0 = (gi > gj) ? v0 : (gi == gj ? 0.5 * v0 : 0.0);
            const double o1 = (gi > gj + 1) ? v1 : (gi == gj + 1 ? 0.5 * v1 : 0.0);
            tile_store2(S1 + (size_t)(a * TPM + b) * TILE_ELEMS, i, j, o0, o1);
        });
        fence_proxy_async();
        // V = inv(L_JJ)^T U
        acc_zero(acc);
        gemm_macro(acc, pipe, opInvCol, TPM * J, opnd_block(S1, ACC_COLK, TPM * J, TPM * J, 1), TPM * J, TPM * J, TPM * J + TPM, nullptr);
        acc_foreach(acc, J, J, [&](int it, int jt, int i, int j, double v0, double v1) {
            const int a = it - TPM * J, b = jt - TPM * J;
            tile_store2(S0 + (size_t)(a * TPM + b) * TILE_ELEMS, i, j, v0, v1);
        });
        fence_proxy_async();
        // S = V inv(L_JJ)
        acc_zero(acc);
        gemm_macro(acc, pipe, opnd_block(S0, ACC_ROWK, TPM * J, TPM * J, 0), TPM * J, opInvCol, TPM * J, TPM * J, TPM * J + TPM, nullptr, 0);
        acc_foreach(acc, J, J, [&](int it, int jt, int i, int j, double v0, double v1) {
            const int a = it - TPM * J, b = jt - TPM * J;
            tile_store2(S1 + (size_t)(a * TPM + b) * TILE_ELEMS, i, j, v0, v1);
        }, 0);
        __syncthreads();
        // Atilde_JJ = S + S^T  (diagonal tiles stored fully symmetric)
        for (int a = 0; a < TPM; ++a)
            for (int b = 0; b <= a; ++b) {
                const double* sab = S1 + (size_t)(a * TPM + b) * TILE_ELEMS;
                const double* sba = S1 + (size_t)(b * TPM + a) * TILE_ELEMS;
                double* tp = LB + packed_tile(TPM * J + a, TPM * J + b) * TILE_ELEMS;
                for (int e = tid; e < TILE_ELEMS; e += GEMM_THREADS) {
                    const int i = e >> 5, j = e & 31;
                    const double w = sab[tswz(i, j)] + sba[tswz(j, i)];
                    tp[tswz(i, j)] = w;
                    if (want_hyp) hyper_accum(P, ka, dkal, dkab, KT, (TPM * J + a) * TS + i, (TPM * J + b) * TS + j, w, kvar, kinv, gl, gv, gb);
                }
            }
        fence_proxy_async();
        __syncthreads();
    }
    gl = block_sum(gl, red);
    gv = block_sum(gv, red);
    gb = block_sum(gb, red);
    if (tid == 0) {
        P.grad_hyp[(size_t)item * 4 + 0] = gl;
        P.grad_hyp[(size_t)item * 4 + 1] = gv;
        P.grad_hyp[(size_t)item * 4 + 3] = gb;
    }
}

// ------------------------------------------------------------------------------------------------ k_predict
// AssignGP.predict_f (assigngp_dense.py:201-240) for up to 128 test inputs per launch, after phases 1-4:
//   tmp1 = (Lc + eps I)^-1 K(X, Xnew),  tmp2 = R^-1 tmp1,  mean = tmp2^T c,  var = Kdiag + colsum(tmp2^2) - colsum(tmp1^2)
// tmp1 / tmp2 are tall tile matrices (nT x 4 tiles) kept in the PB / LB buffers, which the forward pass does not use.
struct PredictArgs {
    const double* Xnew;   // [T][2] (time, function 1..3)
    int T;                // inputs in this launch, <= 128
    int Ttot, s0;         // total number of test inputs of the call and offset of this launch (output row indexing)
    int full_cov;         // only with Ttot <= 128
    int rx_inverse;       // RX already holds R^-1 (state after the gradient phases): tmp2 = R^-1 tmp1 is a product, not a solve
    long long xnew_stride;// doubles between the test inputs of consecutive slots (0: all slots share Xnew)
    int par_chunks;       // grid.y = chunks of 128 test inputs, each with its own scratch inside PB / LB (T, s0 derived per chunk)
    double* mean;         // [slot][T][D]
    double* var;          // [slot][T] or [slot][T][T]
    double* ext;          // optional scratch outside the slot for the tall matrices of every (slot, chunk): full covariance over more than 128 inputs
    long long ext_chunk;  // doubles per (slot, chunk) in ext: two tall matrices + 32 scratch tiles
    const double* neg1;   // Mp times -1.0 (k-scale of the subtracted Gram product in k_predict_cov)
};

__device__ __forceinline__ double kpair(const DenseParams& P, double t1, int f1, double t2, int f2, double ell, double kvar, double kinv,
                                        double b) {
    if (f1 == f2) return kbase(t1 - t2, ell, kvar, P.kind);
    return (kbase(t1 - b, ell, kvar, P.kind) * kinv) * kbase(t2 - b, ell, kvar, P.kind);
}

__global__ void __launch_bounds__(GEMM_THREADS, 1) k_predict(DenseParams P, PredictArgs Q) {
    DENSE_SETUP();
    Pipe pipe;
    pipe_init(pipe, smem);
    double* Dm = pipe.stages;
    const int nM = dm.nM, nT = dm.nT, M = dm.M, D = dm.D, Mp = dm.Mp;
    const int chunk = Q.par_chunks ? (int)blockIdx.y : 0;
    const int qs0 = Q.par_chunks ? chunk * MB : Q.s0;
    const int T = Q.par_chunks ? ((Q.Ttot - qs0 < MB) ? (Q.Ttot - qs0) : MB) : Q.T;
    const double* Xn = Q.Xnew + (size_t)slot * Q.xnew_stride + (Q.par_chunks ? 2 * (size_t)qs0 : 0);
    const size_t tall = (size_t)nT * TPM * TILE_ELEMS;   // one tall nT x 4 tile matrix
    const double b = P.bv[item];
    const double kinv = 1.0 / (kvar + JITTER);
    // scratch blocks: the slot's SCR, or (chunk-parallel) a private pair behind the tall matrices of all chunks in PB
    double* const extc = Q.ext ? Q.ext + ((size_t)slot * gridDim.y + chunk) * Q.ext_chunk : nullptr;
    double* S0 = extc ? extc + 2 * tall : (Q.par_chunks ? PB + (size_t)gridDim.y * tall + (size_t)chunk * 32 * TILE_ELEMS : SCR);
    double* S1 = S0 + 16 * TILE_ELEMS;
    double* T1 = extc ? extc : PB + (size_t)chunk * tall;   // tall tile matrices: tile (r, c) at (r*4 + c)
    double* T2 = extc ? extc + tall : LB + (size_t)chunk * tall;
    const Opnd opL = opnd_packed(LC, ACC_ROWK, LCe);
    const Opnd opR = opnd_packed(RX, ACC_ROWK);
    Acc acc;
    for (int pass = 0; pass < 2; ++pass) {
        double* Tout = pass == 0 ? T1 : T2;
        const Opnd opMat = pass == 0 ? opL : opR;
        const Opnd opPrev = opnd_block(Tout, ACC_COLK, 0, 0, 0);
        for (int I = 0; I < nM; ++I) {
            if (pass == 1 && Q.rx_inverse) {   // tmp2 block row I = sum_{J <= I} R^-1[I][J] tmp1[J]
                acc_zero(acc);
                gemm_macro(acc, pipe, opR, TPM * I, opnd_block(T1, ACC_COLK, 0, 0, 0), 0, 0, TPM * I + TPM, nullptr);
                acc_foreach(acc, I, 0, [&](int it, int jt, int i, int j, double v0, double v1) {
                    tile_store2(Tout + (size_t)(it * TPM + jt) * TILE_ELEMS, i, j, v0, v1);
                });
                __syncthreads();
                continue;
            }
            acc_zero(acc);
            gemm_macro(acc, pipe, opMat, TPM * I, opPrev, 0, 0, TPM * I, nullptr);
            __syncthreads();
            acc_foreach(acc, I, 0, [&](int it, int jt, int i, int j, double v0, double v1) {
                const int a = it - TPM * I;
                double s0, s1;
                if (pass == 0) {
                    const int gi = it * TS + i, s = jt * TS + j;
                    s0 = s1 = 0.0;
                    if (gi < M) {
                        const int ci = gi / 3, fi = gi - 3 * ci + 1;
                        if (s < T) s0 = kpair(P, P.t[ci], fi, Xn[2 * s], (int)Xn[2 * s + 1], ell, kvar, kinv, b);
                        if (s + 1 < T) s1 = kpair(P, P.t[ci], fi, Xn[2 * s + 2], (int)Xn[2 * s + 3], ell, kvar, kinv, b);
                    }
                } else {
                    const double2 p = tile_load2(T1 + (size_t)(it * TPM + jt) * TILE_ELEMS, i, j);
                    s0 = p.x; s1 = p.y;
                }
                tile_store2(S0 + (size_t)(a * TPM + jt) * TILE_ELEMS, i, j, s0 - v0, s1 - v1);
            });
            // inverse of the diagonal macro block: pass 0 needs inv(Lc_II + eps I) (computed here), pass 1 uses the stored inv(R_II)
            const double* invblk;
            if (pass == 0) {
                __syncthreads();
                for (int a = 0; a < TPM; ++a)
                    for (int bb = 0; bb <= a; ++bb) {
                        const double* src = (a == bb) ? LCe + (size_t)(TPM * I + a) * TILE_ELEMS : LC + packed_tile(TPM * I + a, TPM * I + bb) * TILE_ELEMS;
                        for (int e = tid; e < TILE_ELEMS; e += GEMM_THREADS) {
                            const int i = e >> 5, j = e & 31;
                            Dm[(TS * a + i) * DLD + TS * bb + j] = src[tswz(i, j)];
                        }
                    }
                __syncthreads();
                trinv128(Dm);
                store_lower_block(Dm, I, nullptr, S1, nullptr, 0.0);
                invblk = S1;
            } else {
                invblk = RinvD + (size_t)I * 16 * TILE_ELEMS;
            }
            fence_proxy_async();
            __syncthreads();
            acc_zero(acc);
            gemm_macro(acc, pipe, opnd_block(invblk, ACC_ROWK, TPM * I, TPM * I, 1), TPM * I, opnd_block(S0, ACC_COLK, TPM * I, 0, 0), 0,
                       TPM * I, TPM * I + TPM, nullptr);
            acc_foreach(acc, I, 0, [&](int it, int jt, int i, int j, double v0, double v1) {
                tile_store2(Tout + (size_t)(it * TPM + jt) * TILE_ELEMS, i, j, v0, v1);
            });
            fence_proxy_async();
            __syncthreads();
        }
    }
    // mean and marginal variance: column reductions over the Mp rows
    double* red = smem_red(smem);
    {
        const int s = tid & 127, half = tid >> 7;   // `half` = which of the GEMM_THREADS/128 row subsets this thread sums
        const double* cv = vec + VL.c;
        double sq = 0.0;
        for (int d0 = 0; d0 < D; d0 += 1) {
            double m = 0.0;
            for (int i = half; i < Mp; i += GEMM_THREADS / 128) {
                const double x2 = T2[(size_t)((i >> 5) * TPM + (s >> 5)) * TILE_ELEMS + tswz(i & 31, s & 31)];
                m += x2 * cv[(size_t)d0 * Mp + i];
                if (d0 == 0) {
                    const double x1 = T1[(size_t)((i >> 5) * TPM + (s >> 5)) * TILE_ELEMS + tswz(i & 31, s & 31)];
                    sq += x2 * x2 - x1 * x1;
                }
            }
            __syncthreads();
            red[tid] = m;
            __syncthreads();
            if (tid < 128 && s < T) {
                double r = 0.0;
                for (int w = 0; w < GEMM_THREADS / 128; ++w) r += red[tid + 128 * w];
                Q.mean[((size_t)slot * Q.Ttot + qs0 + s) * D + d0] = r;
            }
        }
        __syncthreads();
        red[tid] = sq;
        __syncthreads();
        if (!Q.full_cov && tid < 128 && s < T) {
            double r = kvar + P.square_noise;
            for (int w = 0; w < GEMM_THREADS / 128; ++w) r += red[tid + 128 * w];
            Q.var[(size_t)slot * Q.Ttot + qs0 + s] = r;
        }
    }
    if (Q.full_cov && Q.ext == nullptr) {
        // cov = K(Xnew) + tmp2^T tmp2 - tmp1^T tmp1   (128 x 128 Gram matrices through the tile GEMM)
        const Opnd o2 = opnd_block(T2, ACC_COLK, 0, 0, 0), o1 = opnd_block(T1, ACC_COLK, 0, 0, 0);
        acc_zero(acc);
        gemm_macro(acc, pipe, o2, 0, o2, 0, 0, nT, nullptr);
        acc_foreach(acc, 0, 0, [&](int it, int jt, int i, int j, double v0, double v1) {
            tile_store2(S0 + (size_t)(it * TPM + jt) * TILE_ELEMS, i, j, v0, v1);
        });
        acc_zero(acc);
        gemm_macro(acc, pipe, o1, 0, o1, 0, 0, nT, nullptr);
        acc_foreach(acc, 0, 0, [&](int it, int jt, int i, int j, double v0, double v1) {
            const int s = it * TS + i, r = jt * TS + j;
            const double2 g2 = tile_load2(S0 + (size_t)(it * TPM + jt) * TILE_ELEMS, i, j);
            if (s < T && r < T) {
                double k = kpair(P, Xn[2 * s], (int)Xn[2 * s + 1], Xn[2 * r], (int)Xn[2 * r + 1], ell, kvar, kinv, b);
                if (s == r) k += P.square_noise;
                Q.var[((size_t)slot * T + s) * T + r] = k + g2.x - v0;
            }
            if (s < T && r + 1 < T) {
                double k = kpair(P, Xn[2 * s], (int)Xn[2 * s + 1], Xn[2 * r + 2], (int)Xn[2 * r + 3], ell, kvar, kinv, b);
                if (s == r + 1) k += P.square_noise;
                Q.var[((size_t)slot * T + s) * T + r + 1] = k + g2.y - v1;
            }
        });
    }
}

// Full covariance over more than 128 test inputs: block (c1, c2), c1 >= c2, of  K(Xnew) + tmp2^T tmp2 - tmp1^T tmp1  from the tall
// matrices k_predict left in the external scratch (one CTA per item and block pair; the subtraction rides on the k-scale).
__global__ void __launch_bounds__(GEMM_THREADS, 1) k_predict_cov(DenseParams P, PredictArgs Q, int chunks) {
    DENSE_SETUP();
    Pipe pipe;
    pipe_init(pipe, smem);
    const int nT = dm.nT;
    int c1 = 0;
    while ((c1 + 1) * (c1 + 2) / 2 <= (int)blockIdx.y) ++c1;
    const int c2 = (int)blockIdx.y - c1 * (c1 + 1) / 2;
    const size_t tall = (size_t)nT * TPM * TILE_ELEMS;
    const double* e1 = Q.ext + ((size_t)slot * chunks + c1) * Q.ext_chunk;
    const double* e2 = Q.ext + ((size_t)slot * chunks + c2) * Q.ext_chunk;
    const double* Xn = Q.Xnew + (size_t)slot * Q.xnew_stride;
    const double b = P.bv[item];
    const double kinv = 1.0 / (kvar + JITTER);
    const int Tt = Q.Ttot;
    Acc acc;
    acc_zero(acc);
    gemm_macro(acc, pipe, opnd_block(e1 + tall, ACC_COLK, 0, 0, 0), 0, opnd_block(e2 + tall, ACC_COLK, 0, 0, 0), 0, 0, nT, nullptr);
    gemm_macro(acc, pipe, opnd_block(e1, ACC_COLK, 0, 0, 0), 0, opnd_block(e2, ACC_COLK, 0, 0, 0), 0, 0, nT, Q.neg1);
    double* var = Q.var + (size_t)slot * Tt * Tt;
    acc_foreach(acc, 0, 0, [&](int it, int jt, int i, int j, double v0, double v1) {
        const int s = c1 * MB + it * TS + i, r = c2 * MB + jt * TS + j;
        if (s >= Tt) return;
#pragma unroll
        for (int h2 = 0; h2 < 2; ++h2) {
            const int rr = r + h2;
            if (rr >= Tt) continue;
            double k = kpair(P, Xn[2 * s], (int)Xn[2 * s + 1], Xn[2 * rr], (int)Xn[2 * rr + 1], ell, kvar, kinv, b);
            if (s == rr) k += P.square_noise;
            const double val = k + (h2 ? v1 : v0);
            var[(size_t)s * Tt + rr] = val;
            if (c1 != c2) var[(size_t)rr * Tt + s] = val;
        }
    });
}

__global__ void k_fill(double* p, long long n, double v) {
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) p[e] = v;
}

size_t dense_predict_cov_scratch_doubles(const DenseDims& d, int nitems, int Ttot) {
    const int chunks = (Ttot + MB - 1) / MB;
    const size_t per_chunk = (size_t)(2 * d.nT * TPM + 32) * TILE_ELEMS;
    return (size_t)nitems * chunks * per_chunk + d.Mp;
}

// mean [item][Ttot][D] and the full covariance [item][Ttot][Ttot] for any number of test inputs, after phases 1-4
void dense_fill(double* p, long long n, double v, cudaStream_t st) { k_fill<<<8, 256, 0, st>>>(p, n, v); }

// `ext`: scratch of the first item of this launch (items are dense_predict_cov_scratch_doubles(d, 1, Ttot) - Mp doubles apart);
// `neg1`: Mp times -1.0
void dense_launch_predict_fullcov(const DenseParams& P, int nitems, const double* Xnew, int Ttot, double* mean, double* var, double* ext,
                                  const double* neg1, cudaStream_t st) {
    const int chunks = (Ttot + MB - 1) / MB;
    const long long per_chunk = (long long)(2 * P.d.nT * TPM + 32) * TILE_ELEMS;
    PredictArgs Q{Xnew, 0, Ttot, 0, 1, 0, 0, 1, mean, var, ext, per_chunk, neg1};
    k_predict<<<dim3(nitems, chunks), GEMM_THREADS, SMEM_BYTES, st>>>(P, Q);
    k_predict_cov<<<dim3(nitems, chunks * (chunks + 1) / 2), GEMM_THREADS, SMEM_BYTES, st>>>(P, Q, chunks);
}

void dense_launch_predict(const DenseParams& P, int nitems, const double* Xnew, int T, int Ttot, int s0, int full_cov, double* mean,
                          double* var, cudaStream_t st, int rx_inverse) {
    PredictArgs Q{Xnew, T, Ttot, s0, full_cov, rx_inverse, 0, 0, mean, var, nullptr, 0, nullptr};
    k_predict<<<nitems, GEMM_THREADS, SMEM_BYTES, st>>>(P, Q);
}

// Marginal predictions of `nitems` slots (P.slot_map) at their own Ttot test inputs (Xnew + slot * xnew_stride) from the state
// left by a bound + gradient evaluation.  Chunks of 128 inputs run as separate CTAs when their scratch fits into PB.
void dense_launch_predict_slots(const DenseParams& P, int nitems, const double* Xnew, long long xnew_stride, int Ttot, double* mean,
                                double* var, cudaStream_t st) {
    const int chunks = (Ttot + MB - 1) / MB;
    const long long need = (long long)chunks * ((long long)P.d.nT * TPM + 32) * TILE_ELEMS;
    if (need <= P.d.mat_elems) {
        PredictArgs Q{Xnew, 0, Ttot, 0, 0, 1, xnew_stride, 1, mean, var, nullptr, 0, nullptr};
        k_predict<<<dim3(nitems, chunks), GEMM_THREADS, SMEM_BYTES, st>>>(P, Q);
    } else {
        for (int s0 = 0; s0 < Ttot; s0 += MB) {
            PredictArgs Q{Xnew + 2 * (size_t)s0, (Ttot - s0 < MB) ? (Ttot - s0) : MB, Ttot, s0, 0, 1, xnew_stride, 0, mean, var, nullptr, 0, nullptr};
            k_predict<<<nitems, GEMM_THREADS, SMEM_BYTES, st>>>(P, Q);
        }
    }
}

// ------------------------------------------------------------------------------------------------ debug: unpack a tile matrix to row-major
__global__ void k_unpack(const double* packed, double* out, int Mp, int symmetric) {
    const int nT = Mp / TS;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < (long long)Mp * Mp; e += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(e / Mp), j = (int)(e % Mp);
        const int it = i / TS, jt = j / TS;
        double v = 0.0;
        if (it >= jt) v = packed[packed_tile(it, jt) * TILE_ELEMS + tswz(i % TS, j % TS)];
        else if (symmetric) v = packed[packed_tile(jt, it) * TILE_ELEMS + tswz(j % TS, i % TS)];
        (void)nT;
        out[e] = v;
    }
}

// ------------------------------------------------------------------------------------------------ host launchers
static cudaError_t set_smem(const void* f) { return cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES); }

cudaError_t dense_init_kernels() {
    cudaError_t e;
    if ((e = set_smem((const void*)k_chol<0>)) != cudaSuccess) return e;
    if ((e = set_smem((const void*)k_chol<1>)) != cudaSuccess) return e;
    if ((e = set_smem((const void*)k_syrk)) != cudaSuccess) return e;
    if ((e = set_smem((const void*)k_trtri)) != cudaSuccess) return e;
    if ((e = set_smem((const void*)k_lauum)) != cudaSuccess) return e;
    if ((e = set_smem((const void*)k_trmm)) != cudaSuccess) return e;
    if ((e = set_smem((const void*)k_adj)) != cudaSuccess) return e;
    if ((e = set_smem((const void*)k_predict)) != cudaSuccess) return e;
    if ((e = set_smem((const void*)k_predict_cov)) != cudaSuccess) return e;
    return cudaSuccess;
}

void dense_launch_prep(const DenseParams& P, const double* Apart, const double* vpart, const double* klpart, const double* Y,
                       const int* item_gene, int nitems, cudaStream_t st) {
    PrepArgs Q{Apart, vpart, klpart, Y, item_gene};
    k_prep<<<nitems, NTHREADS, 0, st>>>(P, Q);
    if (P.KConst == nullptr) {
        const int gx = (int)(((long long)P.d.N * P.d.N + NTHREADS * 8 - 1) / (NTHREADS * 8));
        k_ktable<<<dim3(gx < 1 ? 1 : gx, nitems), NTHREADS, 0, st>>>(P);
    }
}

// phase: 1 chol K, 2 P, 3 chol P, 4 solve, 5 trtri, 6 lauum, 7 trmm, 8 post, 9 adjoint
void dense_launch_phase(const DenseParams& P, int nitems, int phase, cudaStream_t st) {
    switch (phase) {
        case 1: k_chol<0><<<nitems, GEMM_THREADS, SMEM_BYTES, st>>>(P); break;
        case 2: k_syrk<<<nitems, GEMM_THREADS, SMEM_BYTES, st>>>(P); break;
        case 3: k_chol<1><<<nitems, GEMM_THREADS, SMEM_BYTES, st>>>(P); break;
        case 4: k_solve<<<nitems, SOLVE_THREADS, 0, st>>>(P); break;
        case 5: k_trtri<<<nitems, GEMM_THREADS, SMEM_BYTES, st>>>(P); break;
        case 6: k_lauum<<<nitems, GEMM_THREADS, SMEM_BYTES, st>>>(P); break;
        case 7: k_trmm<<<nitems, GEMM_THREADS, SMEM_BYTES, st>>>(P); break;
        case 8: k_post<<<nitems, SOLVE_THREADS, 0, st>>>(P); break;
        case 9: k_adj<<<nitems, GEMM_THREADS, SMEM_BYTES, st>>>(P); break;
        default: break;
    }
}

// Cycle counters of the tile GEMM pipeline (only a BGP_TIMING build fills them; tools/gemm_timers.py).
void dense_read_gemm_timers(unsigned long long* out8, int reset) {
#ifdef BGP_TIMING
    cudaMemcpyFromSymbol(out8, g_gemm_timers, 8 * sizeof(unsigned long long));
    if (reset) { unsigned long long z[8] = {0}; cudaMemcpyToSymbol(g_gemm_timers, z, sizeof(z)); }
#else
    for (int i = 0; i < 8; ++i) out8[i] = 0;
    (void)reset;
#endif
}

void dense_unpack(const double* packed, double* out, int Mp, int symmetric, cudaStream_t st) {
    k_unpack<<<296, 256, 0, st>>>(packed, out, Mp, symmetric);
}

}  // namespace bgp
