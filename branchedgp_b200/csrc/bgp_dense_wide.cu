// Dense AssignGP bound + gradient with MANY CTAs per work item ("wide" pipeline), sm_100a.
//
// bgp_dense.cu runs one CTA per (gene, branching point) item: the throughput design for batches of >= one item per SM.  The
// reference's own call pattern (FitBranchingModel.py:124-132) issues ONE evaluation at a time, which there is 1 of 148 SMs.  When
// fewer items than half the SMs are in flight the same phases run here with the 128x128 macro tiles of an item dealt to G CTAs:
//
//   * every phase is one kernel, grid (G, items); a CTA repeatedly takes the next task of its item from an atomic queue head;
//   * the tasks are listed in a topological order and carry their dependencies as per-tile counters in global memory
//     (release / acquire at gpu scope): a task only ever waits for tasks with a smaller queue index, which have been taken by
//     running CTAs, so the scheme cannot deadlock whatever the number of resident CTAs;
//   * the factorisations are restated in scatter (right-looking) form so that a finished 128-wide panel releases O(nM^2)
//     independent tile updates (oracle/blocked_dense_wide.py is the numpy emulation of these task bodies):
//       k_wchol   right-looking tile Cholesky, the last update of a tile fused into its factor / panel task (lookahead);
//       k_wtrtri  block Gauss-Jordan inverse of the row- and column-scaled factor, no serial column sweep;
//       k_wadj    scatter form of the blocked Cholesky adjoint; the long transposed products are split over the panel tasks
//                 (per-row partials summed in a fixed order by the diagonal task);
//       k_wsyrk / k_wlauum / k_wtrmm  independent macro tiles, longest first;
//   * every tile is produced by the same sequence of operations whatever G is and whichever CTA runs the task, so results
//     are bit-identical for any number of items in flight (tests/test_sharded_gpu.py).
// The triangular solves and the vector epilogue (k_wsolve, k_wpost) are dealt out the same way, one 128-row block per task, with
// the block vectors handed from task to task through flags.  k_prep, k_ktable and the assignment passes are shared with the
// one-CTA pipeline.
#include "bgp_dense_dev.cuh"
#include "bgp_launch.h"

namespace bgp {

// ------------------------------------------------------------------------------------------------ control area
// Per slot and per phase: [0] queue head, then cnt[tiles], fin[tiles], col[nM], fail[nM], aux[tiles]   (ints, zeroed before every evaluation)
enum { WPH_CHOLK = 0, WPH_SYRK, WPH_CHOLP, WPH_SOLVE, WPH_TRTRI, WPH_LAUUM, WPH_TRMM, WPH_POST, WPH_ADJ, WPH_COUNT };

__host__ __device__ inline int wide_phase_ints(int nM) { return (8 + 3 * (nM * (nM + 1) / 2) + 2 * nM + 3) & ~3; }
long long dense_wide_ctrl_ints(int nM) { return (long long)WPH_COUNT * wide_phase_ints(nM); }

struct WideCtl {
    int* head;
    int* cnt;
    int* fin;
    int* col;
    int* fail;
    int* aux;
};
__device__ __forceinline__ WideCtl wide_ctl(const DenseParams& P, int slot, int phase) {
    const int nM = P.d.nM, nt = nM * (nM + 1) / 2;
    int* base = P.ctrl + (size_t)slot * P.ctrl_stride + (size_t)phase * wide_phase_ints(nM);
    WideCtl C;
    C.head = base; C.cnt = base + 8; C.fin = C.cnt + nt; C.col = C.fin + nt; C.fail = C.col + nM; C.aux = C.fail + nM;
    return C;
}
__device__ __forceinline__ int tix(int I, int J) { return I * (I + 1) / 2 + J; }

__device__ __forceinline__ int ld_acquire(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(int* p, int v) { asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ void spin_ge(const int* p, int v) {
    while (ld_acquire(p) < v) __nanosleep(32);
}
// after thread 0 has seen its dependencies: make them visible to the whole CTA and to the TMA engine
__device__ __forceinline__ void deps_ready() {
    __syncthreads();
    fence_proxy_async();
}
// all global writes of the task done -> publish
__device__ __forceinline__ void task_signal(int* flag, int v) {
    fence_proxy_async();
    __syncthreads();
    if (threadIdx.x == 0) { __threadfence(); st_release(flag, v); }
}
__device__ __forceinline__ void task_signal_add(int* counter) {
    fence_proxy_async();
    __syncthreads();
    if (threadIdx.x == 0) { __threadfence(); atomicAdd(counter, 1); }
}
__device__ __forceinline__ int wq_next(int* head, int* s_task) {
    __syncthreads();
    if (threadIdx.x == 0) *s_task = atomicAdd(head, 1);
    __syncthreads();
    return *s_task;
}
// tiles written by other CTAs of the same kernel are read past the (incoherent) L1
__device__ __forceinline__ double2 tile_load2_cg(const double* tile, int i, int j) {
    return __ldcg(reinterpret_cast<const double2*>(tile + tswz(i, j)));
}

// Optional cycle counters of the critical-path tasks (BGP_TIMING builds; tools/wide_timers.py): sums over tasks, thread 0's clock
#ifdef BGP_TIMING
__device__ unsigned long long g_wide_timers[48];
#define WT_DECL() long long wt_ = clock64()
#define WT_LAP(slot_) do { if (threadIdx.x == 0) { const long long n_ = clock64(); atomicAdd(&g_wide_timers[slot_], (unsigned long long)(n_ - wt_)); wt_ = n_; } } while (0)
#define WT_COUNT(slot_) do { if (threadIdx.x == 0) atomicAdd(&g_wide_timers[slot_], 1ull); } while (0)
#else
#define WT_DECL()
#define WT_LAP(slot_)
#define WT_COUNT(slot_)
#endif

#define WIDE_SETUP(PHASE)                                   \
    DENSE_SETUP_IDX(blockIdx.y);                            \
    __shared__ int s_task;                                  \
    Pipe pipe;                                              \
    pipe_init(pipe, smem);                                  \
    pipe.row_limit = dm.nTv;                                \
    const int nM = dm.nM;                                   \
    const WideCtl C = wide_ctl(P, slot, PHASE);             \
    (void)nM;

// ------------------------------------------------------------------------------------------------ k_wchol
// Tasks, in queue order:  DIAG(0), PANEL(I,0);  then for J = 0 .. nM-2:  DIAG(J+1), PANEL(I,J+1) (I > J+1)  [they carry update J
// of their own tile],  UPD(I,K;J) for K >= J+2, I >= K.   cnt[tile (I,K)] counts the updates applied; K + 1 means final.
template <int SRC>
__global__ void __launch_bounds__(GEMM_THREADS, 1) k_wchol(DenseParams P) {
    WIDE_SETUP(SRC == 0 ? WPH_CHOLK : WPH_CHOLP);
    double* red = smem_red(smem);
    double* Dm = pipe.stages;
    double* dst = SRC == 0 ? LC : RX;
    double* invD = SRC == 0 ? LinvD : RinvD;
    const double* ka = vec + VL.ka;
    const double* KT = P.KT + (size_t)slot * P.slot_stride;
    const double kinv = 1.0 / (kvar + JITTER);   // inv(k(b,b) + jitter), branch_kernParamGPflow.py:168-189
    const Opnd opL = opnd_packed(dst, ACC_ROWK);
    int ntasks = nM;
    for (int J = 0; J + 1 < nM; ++J) { const int r = nM - 2 - J; ntasks += (nM - 1 - J) + r * (r + 1) / 2; }
    Acc acc;
    for (;;) {
        const int t = wq_next(C.head, &s_task);
        if (t >= ntasks) break;
        int type = 0, I = 0, K = 0, J = 0;   // type 0 DIAG(K), 1 PANEL(I,K), 2 UPD(I,K;J)
        if (t < nM) { type = t == 0 ? 0 : 1; I = t; K = 0; }
        else {
            int u = t - nM;
            for (J = 0;; ++J) {
                const int na = nM - 1 - J, r = nM - 2 - J, nb = r * (r + 1) / 2;
                if (u < na) { type = u == 0 ? 0 : 1; K = J + 1; I = K + u; break; }
                u -= na;
                if (u < nb) { type = 2; K = J + 2; while (u >= nM - K) { u -= nM - K; ++K; } I = K + u; break; }
                u -= nb;
            }
        }
        if (type == 0) {
            // ---- diagonal block K: last update (block column K-1), factor, inverse
            WT_DECL();
            WT_COUNT(0);
            if (tid == 0) {
                if (K >= 2) spin_ge(&C.cnt[tix(K, K)], K - 1);
                if (K >= 1) spin_ge(&C.cnt[tix(K, K - 1)], K);
            }
            deps_ready();
            WT_LAP(1);
            acc_zero(acc);
            if (K >= 1) gemm_macro(acc, pipe, opL, TPM * K, opL, TPM * K, TPM * (K - 1), TPM * K, nullptr, 2);
            __syncthreads();   // the ring doubles as the 128x128 scratch from here on
            WT_LAP(2);
            acc_foreach(acc, K, K, [&](int it, int jt, int i, int j, double v0, double v1) {
                if (it < jt) return;
                const int gi = it * TS + i, gj = jt * TS + j;
                double s0, s1;
                if (SRC == 0 && K <= 1) {
                    s0 = kgen(P, ka, KT, gi, gj, kinv);
                    s1 = kgen(P, ka, KT, gi, gj + 1, kinv);
                } else {
                    const double2 p = tile_load2_cg(dst + packed_tile(it, jt) * TILE_ELEMS, i, j);
                    s0 = p.x; s1 = p.y;
                }
                double* o = Dm + (gi - MB * K) * DLD + (gj - MB * K);
                o[0] = s0 - v0;
                o[1] = s1 - v1;
            }, 2);
            __syncthreads();
            WT_LAP(3);
            int fail = 0;
            chol128(Dm, fail, MB * K);
            WT_LAP(4);
            if (SRC == 1) {
                double l = 0.0;
                if (tid < MB) { const double r = Dm[tid * DLD + tid]; l = log(r * r); }
                l = block_sum(l, red);
                if (tid == 0) vec[VL.dpart + K] = l;   // dpart is free until k_wtrmm
            }
            store_lower_block(Dm, K, dst, nullptr, SRC == 0 ? LCe : nullptr, JITTER);
            __syncthreads();
            WT_LAP(5);
            trinv128(Dm);
            WT_LAP(6);
            store_lower_block(Dm, K, nullptr, invD + (size_t)K * 16 * TILE_ELEMS, nullptr, 0.0);
            if (tid == 0) C.fail[K] = fail;
            if (K == nM - 1) {
                // the last diagonal task follows (transitively) every other one: log-det in block order, first failure
                __syncthreads();
                if (tid == 0) {
                    int f = 0;
                    double ld = 0.0;
                    for (int k = 0; k < nM; ++k) {
                        const int fk = (k == K) ? fail : __ldcg(&C.fail[k]);
                        if (f == 0 && fk != 0) f = fk;
                        if (SRC == 1) ld += (k == K) ? vec[VL.dpart + k] : __ldcg(&vec[VL.dpart + k]);
                    }
                    if (SRC == 1) vec[VL.scal + SC_LOGDET] = ld;
                    if (f != 0 && P.status[item] == 0) P.status[item] = f + (SRC == 1 ? dm.Mp : 0);
                }
            }
            task_signal(&C.cnt[tix(K, K)], K + 1);
            WT_LAP(7);
        } else if (type == 1) {
            // ---- panel tile (I, K): last update, then  X = C * inv(L_KK)^T
            WT_DECL();
            WT_COUNT(8);
            if (K >= 1) {
                if (tid == 0) {
                    if (K >= 2) spin_ge(&C.cnt[tix(I, K)], K - 1);
                    spin_ge(&C.cnt[tix(I, K - 1)], K);
                    spin_ge(&C.cnt[tix(K, K - 1)], K);
                }
                deps_ready();
            }
            acc_zero(acc);
            if (K >= 1) gemm_macro(acc, pipe, opL, TPM * I, opL, TPM * K, TPM * (K - 1), TPM * K, nullptr);
            acc_foreach(acc, I, K, [&](int it, int jt, int i, int j, double v0, double v1) {
                double* tp = dst + packed_tile(it, jt) * TILE_ELEMS;
                double s0, s1;
                if (SRC == 0 && K <= 1) {
                    const int gi = it * TS + i, gj = jt * TS + j;
                    s0 = kgen(P, ka, KT, gi, gj, kinv);
                    s1 = kgen(P, ka, KT, gi, gj + 1, kinv);
                } else {
                    const double2 p = tile_load2_cg(tp, i, j);
                    s0 = p.x; s1 = p.y;
                }
                tile_store2(tp, i, j, s0 - v0, s1 - v1);
            });
            fence_proxy_async();
            WT_LAP(9);
            if (tid == 0) spin_ge(&C.cnt[tix(K, K)], K + 1);
            deps_ready();
            WT_LAP(10);
            const Opnd opInv = opnd_block(invD + (size_t)K * 16 * TILE_ELEMS, ACC_ROWK, TPM * K, TPM * K, 1);
            acc_zero(acc);
            gemm_macro(acc, pipe, opL, TPM * I, opInv, TPM * K, TPM * K, TPM * K + TPM, nullptr, 0);   // the skipped operand is B
            acc_foreach(acc, I, K, [&](int it, int jt, int i, int j, double v0, double v1) {
                tile_store2(dst + packed_tile(it, jt) * TILE_ELEMS, i, j, v0, v1);
            }, 0);
            task_signal(&C.cnt[tix(I, K)], K + 1);
            WT_LAP(11);
        } else {
            // ---- trailing update of tile (I, K) with block column J
            WT_DECL();
            WT_COUNT(12);
            if (tid == 0) {
                spin_ge(&C.cnt[tix(I, J)], J + 1);
                spin_ge(&C.cnt[tix(K, J)], J + 1);
                if (J >= 1) spin_ge(&C.cnt[tix(I, K)], J);
            }
            deps_ready();
            WT_LAP(13);
            const int rowmap = (I == K) ? 2 : 1;
            acc_zero(acc);
            gemm_macro(acc, pipe, opL, TPM * I, opL, TPM * K, TPM * J, TPM * J + TPM, nullptr, rowmap);
            acc_foreach(acc, I, K, [&](int it, int jt, int i, int j, double v0, double v1) {
                if (it < jt) return;
                double* tp = dst + packed_tile(it, jt) * TILE_ELEMS;
                double s0, s1;
                if (SRC == 0 && J == 0) {
                    const int gi = it * TS + i, gj = jt * TS + j;
                    s0 = kgen(P, ka, KT, gi, gj, kinv);
                    s1 = kgen(P, ka, KT, gi, gj + 1, kinv);
                } else {
                    const double2 p = tile_load2_cg(tp, i, j);
                    s0 = p.x; s1 = p.y;
                }
                tile_store2(tp, i, j, s0 - v0, s1 - v1);
            }, rowmap);
            task_signal(&C.cnt[tix(I, K)], J + 1);
            WT_LAP(14);
        }
    }
}

// ------------------------------------------------------------------------------------------------ independent macro tiles
__device__ __forceinline__ void tile_of(int t, int& I, int& J) {
    I = 0;
    while ((I + 1) * (I + 2) / 2 <= t) ++I;
    J = t - I * (I + 1) / 2;
}

// P = L^T Delta L + I  (L = Lc + eps I);  k range [4I, nTv): longest tiles first = ascending I
__global__ void __launch_bounds__(GEMM_THREADS, 1) k_wsyrk(DenseParams P) {
    WIDE_SETUP(WPH_SYRK);
    const Opnd opL = opnd_packed(LC, ACC_COLK, LCe);
    const double* Delta = vec + VL.Delta;
    const int ntasks = nM * (nM + 1) / 2;
    Acc acc;
    for (;;) {
        const int t = wq_next(C.head, &s_task);
        if (t >= ntasks) break;
        int I, J;
        tile_of(t, I, J);
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

// PB = -D/2 X^T X - 1/2 q q^T
__global__ void __launch_bounds__(GEMM_THREADS, 1) k_wlauum(DenseParams P) {
    WIDE_SETUP(WPH_LAUUM);
    const Opnd opX = opnd_packed(RX, ACC_COLK);
    const double* q = vec + VL.q;
    const int Mp = dm.Mp, D = dm.D;
    const double hD = 0.5 * (double)D;
    const int ntasks = nM * (nM + 1) / 2;
    Acc acc;
    for (;;) {
        const int t = wq_next(C.head, &s_task);
        if (t >= ntasks) break;
        int I, J;
        tile_of(t, I, J);
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
}

// G = L Pbar (lower), Lbar = tril(2 Delta G + v zbar^T), delta partials;  k range [0, 4I+4): longest first = descending I
__global__ void __launch_bounds__(GEMM_THREADS, 1) k_wtrmm(DenseParams P) {
    WIDE_SETUP(WPH_TRMM);
    double* red = smem_red(smem);
    const Opnd opL = opnd_packed(LC, ACC_ROWK, LCe);
    const Opnd opP = opnd_packed(PB, ACC_SYM);
    const double* Delta = vec + VL.Delta;
    const double* v = vec + VL.v;
    const double* q = vec + VL.q;
    double* dpart = vec + VL.dpart;
    const int Mp = dm.Mp, D = dm.D;
    const int warp = tid >> 5, lane = tid & 31, g = lane >> 2, c = lane & 3;
    const int ntasks = nM * (nM + 1) / 2;
    Acc acc;
    for (;;) {
        const int t = wq_next(C.head, &s_task);
        if (t >= ntasks) break;
        int I, J;
        tile_of(ntasks - 1 - t, I, J);
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
        if (rowmap == 2)
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
                if (c == 0 && !idle_) red[wb_ * MB + wa_ * TS + 8 * mi + g] = r;
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

// ------------------------------------------------------------------------------------------------ k_wsolve:  z, c, q
// z = (Lc + eps I)^T v,  c = R^-1 z / s,  q = R^-T c  (assigngp_dense.py:174-181) for one item by many CTAs.
// Queue: ZCOL(jt) for the nT tile columns of z (longest first), FWD(J) for J ascending, BWD(J) for J descending.  FWD(J) owns
// block row J of R: it consumes the blocks c_k, k < J, as their flags come up (so the chain per block is one tile product and
// one 128x128 triangular product), then publishes c_J; BWD likewise with the block columns.  Every sum has a fixed order.
// fin[J]: tile columns of z block J done (4);  col[J]: c_J published;  fail[J]: q_J published (arrays reused as plain flags).
constexpr int WS_THREADS = GEMM_THREADS;
constexpr int WS_WARPS = WS_THREADS / 32;

__global__ void __launch_bounds__(WS_THREADS) k_wsolve(DenseParams P) {
    DENSE_SETUP_IDX(blockIdx.y);
    __shared__ int s_task;
    __shared__ double part[WS_WARPS][TS];
    __shared__ double yv[MB];
    __shared__ double red[32];
    const int nM = dm.nM, nT = dm.nT, Mp = dm.Mp, D = dm.D;
    const WideCtl C = wide_ctl(P, slot, WPH_SOLVE);
    int* zcnt = C.fin;
    int* cflag = C.col;
    int* qflag = C.fail;
    const int warp = tid >> 5, lane = tid & 31;
    double* cu = vec + VL.vbar;       // unscaled R^-1 z (the recursion runs on it); vbar is written by k_wpost later
    double* c2part = vec + VL.Abar;   // per-block sums of c^2 (Abar is written by k_wpost later)
    const int ntasks = nT + 2 * nM;
    for (;;) {
        const int t = wq_next(C.head, &s_task);
        if (t >= ntasks) break;
        if (t < nT) {
            // ---- ZCOL(jt): z[jt] = sum_{it >= jt} tile(it,jt)^T v[it]
            const int jt = t;
            for (int d = 0; d < D; ++d) {
                const double* v = vec + VL.v + (size_t)d * Mp;
                double s = 0.0;
                for (int it = jt + warp; it < nT; it += WS_WARPS) {
                    const double* tile = (it == jt) ? LCe + (size_t)it * TILE_ELEMS : LC + packed_tile(it, jt) * TILE_ELEMS;
                    const double* vv = v + it * TS;
#pragma unroll
                    for (int i = 0; i < TS; ++i) s += tile[tswz(i, lane)] * vv[i];
                }
                part[warp][lane] = s;
                __syncthreads();
                if (warp == 0) {
                    double r = 0.0;
#pragma unroll
                    for (int w = 0; w < WS_WARPS; ++w) r += part[w][lane];
                    vec[VL.z + (size_t)d * Mp + jt * TS + lane] = r;
                }
                __syncthreads();
            }
            if (tid == 0) { __threadfence(); atomicAdd(&zcnt[jt / TPM], 1); }
        } else if (t < nT + nM) {
            // ---- FWD(J): cu_J = inv(R_JJ) (z_J - sum_{k<J} R[J,k] cu_k),  c_J = cu_J / s
            const int J = t - nT;
            const int a = warp & 3, bq = warp >> 2;   // tile row inside block J, tile column inside block k
            double c2 = 0.0;
            for (int d = 0; d < D; ++d) {
                double s = 0.0;
                for (int k = 0; k < J; ++k) {
                    if (lane == 0) spin_ge(&cflag[k], 1);
                    __syncwarp();
                    const double* row = RX + packed_tile(TPM * J + a, TPM * k + bq) * TILE_ELEMS + lane * TS;
                    const double* cc = cu + (size_t)d * Mp + (TPM * k + bq) * TS;
                    const int sw = (lane & 3) << 2;
#pragma unroll
                    for (int kk = 0; kk < TS; ++kk) s += row[kk ^ sw] * __ldcg(cc + kk);
                }
                part[warp][lane] = s;
                if (tid == 0) spin_ge(&zcnt[J], TPM);
                __syncthreads();
                if (tid < MB) {
                    double y = __ldcg(vec + VL.z + (size_t)d * Mp + MB * J + tid);
#pragma unroll
                    for (int q4 = 0; q4 < 4; ++q4) y -= part[q4 * 4 + (tid >> 5)][tid & 31];
                    yv[tid] = y;
                }
                __syncthreads();
                double x = 0.0;
                if (tid < MB) {
                    const double* blk = RinvD + (size_t)J * 16 * TILE_ELEMS;
                    const int ar = tid >> 5, i = tid & 31;
                    double sacc = 0.0;
                    for (int b = 0; b <= ar; ++b) {
                        const double* row = blk + (size_t)(ar * TPM + b) * TILE_ELEMS + i * TS;
                        const int sw = (i & 3) << 2;
#pragma unroll
                        for (int kk = 0; kk < TS; ++kk) sacc += row[kk ^ sw] * yv[b * TS + kk];
                    }
                    cu[(size_t)d * Mp + MB * J + tid] = sacc;
                    x = sacc / likvar;
                    vec[VL.c + (size_t)d * Mp + MB * J + tid] = x;
                }
                c2 += block_sum(x * x, red);
                __syncthreads();
            }
            if (tid == 0) c2part[J] = c2;
            task_signal(&cflag[J], 1);
        } else {
            // ---- BWD(J): q_J = inv(R_JJ)^T (c_J - sum_{I>J} R[I,J]^T q_I)
            const int J = nM - 1 - (t - nT - nM);
            const int b = warp & 3, aq = warp >> 2;   // tile column inside block J, tile row inside block I
            if (tid == 0) spin_ge(&cflag[J], 1);
            __syncthreads();
            for (int d = 0; d < D; ++d) {
                double s = 0.0;
                for (int I = nM - 1; I > J; --I) {
                    if (lane == 0) spin_ge(&qflag[I], 1);
                    __syncwarp();
                    const double* tile = RX + packed_tile(TPM * I + aq, TPM * J + b) * TILE_ELEMS;
                    const double* qq = vec + VL.q + (size_t)d * Mp + (TPM * I + aq) * TS;
#pragma unroll
                    for (int i = 0; i < TS; ++i) s += tile[tswz(i, lane)] * __ldcg(qq + i);
                }
                part[warp][lane] = s;
                __syncthreads();
                if (tid < MB) {
                    double y = __ldcg(vec + VL.c + (size_t)d * Mp + MB * J + tid);
#pragma unroll
                    for (int q4 = 0; q4 < 4; ++q4) y -= part[q4 * 4 + (tid >> 5)][tid & 31];
                    yv[tid] = y;
                }
                __syncthreads();
                if (tid < MB) {
                    const double* blk = RinvD + (size_t)J * 16 * TILE_ELEMS;
                    const int bc = tid >> 5, j = tid & 31;
                    double sacc = 0.0;
                    for (int ar = bc; ar < TPM; ++ar) {
                        const double* tile = blk + (size_t)(ar * TPM + bc) * TILE_ELEMS;
#pragma unroll
                        for (int kk = 0; kk < TS; ++kk) sacc += tile[tswz(kk, j)] * yv[ar * TS + kk];
                    }
                    vec[VL.q + (size_t)d * Mp + MB * J + tid] = sacc;
                }
                __syncthreads();
            }
            if (J == 0 && tid == 0) {   // follows every FWD task transitively: sum of c^2 in block order
                double c2 = 0.0;
                for (int k = 0; k < nM; ++k) c2 += __ldcg(&c2part[k]);
                vec[VL.scal + SC_SUMC2] = c2;
            }
            task_signal(&qflag[J], 1);
        }
    }
}

// ------------------------------------------------------------------------------------------------ k_wpost: delta, Abar, vbar, d/d likvar, ELBO
// One task per 32-row tile row (longest first): vbar = L zbar for those rows, delta / Abar from the partials of k_wtrmm and the
// tile row's share of sum_i A_i delta_i.  The task that finishes last adds the shares in row order and writes the scalars.
__global__ void __launch_bounds__(WS_THREADS) k_wpost(DenseParams P) {
    DENSE_SETUP_IDX(blockIdx.y);
    __shared__ int s_task;
    __shared__ int s_last;
    __shared__ double part[WS_WARPS][TS];
    const int nT = dm.nT, Mp = dm.Mp, M = dm.M, D = dm.D, N = dm.N;
    const WideCtl C = wide_ctl(P, slot, WPH_POST);
    const int warp = tid >> 5, lane = tid & 31;
    double* adpart = vec + VL.z;   // per-tile-row shares of sum A delta (z is dead after k_wsolve)
    for (;;) {
        const int t = wq_next(C.head, &s_task);
        if (t >= nT) break;
        const int it = nT - 1 - t;
        // vbar[i] = sum_{j <= i} L[i][j] q[j] / s
        for (int d = 0; d < D; ++d) {
            const double* q = vec + VL.q + (size_t)d * Mp;
            double s = 0.0;
            const int sw = (lane & 3) << 2;
            for (int jt = warp; jt <= it; jt += WS_WARPS) {
                const double* row = ((it == jt) ? LCe + (size_t)it * TILE_ELEMS : LC + packed_tile(it, jt) * TILE_ELEMS) + lane * TS;
                const double* qq = q + jt * TS;
#pragma unroll
                for (int k = 0; k < TS; ++k) s += row[k ^ sw] * qq[k];
            }
            part[warp][lane] = s;
            __syncthreads();
            if (warp == 0) {
                double r = 0.0;
#pragma unroll
                for (int w = 0; w < WS_WARPS; ++w) r += part[w][lane];
                vec[VL.vbar + (size_t)d * Mp + it * TS + lane] = r / likvar;
            }
            __syncthreads();
        }
        if (warp == 0) {
            const int i = it * TS + lane, I = i / MB;
            double s = 0.0;
            for (int J = 0; J <= I; ++J) s += vec[VL.dpart + (size_t)J * Mp + i];
            vec[VL.delta + i] = s;
            vec[VL.Abar + i] = s / likvar;
            double ad = (i < M) ? vec[VL.A + i] * s : 0.0;
            ad = warp_sum(ad);
            if (lane == 0) adpart[it] = ad;
        }
        __syncthreads();
        if (tid == 0) { __threadfence(); s_last = (atomicAdd(&C.col[0], 1) == nT - 1); }
        __syncthreads();
        if (s_last && tid == 0) {
            __threadfence();
            double adel = 0.0;
            for (int k = 0; k < nT; ++k) adel += __ldcg(&adpart[k]);
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
}

// ------------------------------------------------------------------------------------------------ k_wtrtri:  RX <- inv(R)
// With D_J = inv(R_JJ) (left by k_wchol<1>):  PRE(I,J): W0[I,J] = -D_I R[I,J] -> LB,  V[I,J] = W0[I,J] D_J -> RX (in place);
// TR(I,c;J), I > J > c, J ascending:  V[I,c] += W0[I,J] V[J,c]  (row J of V is final by then);  diagonal tiles = D_J.
// cnt[tile]: 1 after PRE, +1 per TR;  tile (J,c) is final at J - c.
__global__ void __launch_bounds__(GEMM_THREADS, 1) k_wtrtri(DenseParams P) {
    WIDE_SETUP(WPH_TRTRI);
    const int nOff = nM * (nM - 1) / 2;
    int ntasks = nOff + nM;
    for (int J = 1; J + 1 < nM; ++J) ntasks += (nM - 1 - J) * J;
    const Opnd opRcol = opnd_packed(RX, ACC_COLK);
    const Opnd opW = opnd_packed(LB, ACC_ROWK);
    Acc acc;
    for (;;) {
        const int t = wq_next(C.head, &s_task);
        if (t >= ntasks) break;
        if (t < nOff) {
            int I = 1;
            while (I * (I + 1) / 2 <= t) ++I;
            const int J = t - I * (I - 1) / 2;
            const Opnd opDI = opnd_block(RinvD + (size_t)I * 16 * TILE_ELEMS, ACC_ROWK, TPM * I, TPM * I, 1);
            acc_zero(acc);
            gemm_macro(acc, pipe, opDI, TPM * I, opRcol, TPM * J, TPM * I, min(TPM * I + TPM, dm.nTv), nullptr);
            acc_foreach(acc, I, J, [&](int it, int jt, int i, int j, double v0, double v1) {
                tile_store2(LB + packed_tile(it, jt) * TILE_ELEMS, i, j, -v0, -v1);
            });
            fence_proxy_async();
            const Opnd opDJ = opnd_block(RinvD + (size_t)J * 16 * TILE_ELEMS, ACC_COLK, TPM * J, TPM * J, 1);
            acc_zero(acc);
            gemm_macro(acc, pipe, opW, TPM * I, opDJ, TPM * J, TPM * J, TPM * J + TPM, nullptr, 0);   // the skipped operand is B
            acc_foreach(acc, I, J, [&](int it, int jt, int i, int j, double v0, double v1) {
                tile_store2(RX + packed_tile(it, jt) * TILE_ELEMS, i, j, v0, v1);
            }, 0);
            task_signal(&C.cnt[tix(I, J)], 1);
        } else if (t < nOff + nM) {
            const int J = t - nOff;
            const double* blk = RinvD + (size_t)J * 16 * TILE_ELEMS;
            for (int a = 0; a < TPM; ++a)
                for (int b = 0; b <= a; ++b) {
                    const double* src = blk + (size_t)(a * TPM + b) * TILE_ELEMS;
                    double* dstp = RX + packed_tile(TPM * J + a, TPM * J + b) * TILE_ELEMS;
                    for (int e = tid; e < TILE_ELEMS; e += GEMM_THREADS) dstp[e] = src[e];
                }
        } else {
            int u = t - nOff - nM, J = 1;
            while (u >= (nM - 1 - J) * J) { u -= (nM - 1 - J) * J; ++J; }
            const int I = J + 1 + u / J, c = u % J;
            if (tid == 0) {
                spin_ge(&C.cnt[tix(I, J)], 1);
                spin_ge(&C.cnt[tix(J, c)], J - c);
                spin_ge(&C.cnt[tix(I, c)], J - c);
            }
            deps_ready();
            acc_zero(acc);
            gemm_macro(acc, pipe, opW, TPM * I, opRcol, TPM * c, TPM * J, TPM * J + TPM, nullptr);
            acc_foreach(acc, I, c, [&](int it, int jt, int i, int j, double v0, double v1) {
                double* tp = RX + packed_tile(it, jt) * TILE_ELEMS;
                const double2 l = tile_load2_cg(tp, i, j);
                tile_store2(tp, i, j, l.x + v0, l.y + v1);
            });
            task_signal(&C.cnt[tix(I, c)], J - c + 1);
        }
    }
}

// ------------------------------------------------------------------------------------------------ k_wadj
// Scatter form of the blocked Cholesky adjoint (derivation and task order: oracle/blocked_dense_wide.py::chol_adjoint_scatter).
// Queue order:  ADIAG(nM-1);  for K = nM-1 .. 1:  APANEL(I, K-1) (I = K first), ADIAG(K-1), BROW(K, J) (J = K-2 .. 0),
// SC(I, c; K) (c = K-2 .. 0, I > K).   fin[tile] = 1: Atilde tile final;  col[J]: panel partials P_I of column J written;
// cnt[tile (I,J)]: 1 after BROW, +1 per SC.   Partial blocks (16 tiles each) live in PB, which is dead after k_wtrmm.
__global__ void __launch_bounds__(GEMM_THREADS, 1) k_wadj(DenseParams P) {
    WIDE_SETUP(WPH_ADJ);
    double* red = smem_red(smem);
    const int nT = dm.nTv;   // k-ranges stop at the last tile that holds a row < M
    const double* ka = vec + VL.ka;
    const double* dkal = vec + VL.dkal;
    const double* dkab = vec + VL.dkab;
    const double* KT = P.KT + (size_t)slot * P.slot_stride;
    const double kinv = 1.0 / (kvar + JITTER);
    const bool want_hyp = (P.KConst == nullptr);
    double* hp = vec + VL.dpart;   // per-tile hyper-gradient partials [tile][3] (dpart is dead after k_post)
    const Opnd opAsym = opnd_packed(LB, ACC_SYM);
    const Opnd opArow = opnd_packed(LB, ACC_ROWK);
    const Opnd opAcol = opnd_packed(LB, ACC_COLK);
    const Opnd opLcol = opnd_packed(LC, ACC_COLK);
    double* S0 = SCR;
    double* S1 = SCR + 16 * TILE_ELEMS;
    auto Pscr = [&](int I) { return PB + (size_t)(I - 1) * 16 * TILE_ELEMS; };
    auto Qscr = [&](int I) { return PB + (size_t)(nM - 1 + I - 1) * 16 * TILE_ELEMS; };
    int ntasks = 1;
    // BROW(K, J) is one long product (up to nT - 4K k-steps on ONE CTA) that the chain needs two columns later; in the last third of
    // the sweep it is longer than two columns take.  Long ones are cut into <= 4 chunks of the k range that run as separate
    // tasks (partial blocks in PB behind the P / Q blocks, two generations) and are added, in chunk order, by the BROW task itself.
    constexpr int BROW_CH = 24, BROW_NC = 4;
    const bool chunk_ok = dm.mat_elems / TILE_ELEMS >= (long long)32 * (nM - 1) + (long long)2 * BROW_NC * 16 * (nM - 2) && nM >= 3;
    auto nchunks = [&](int K_) {
        int nc = (nT - TPM * K_ + BROW_CH - 1) / BROW_CH;
        if (nc > BROW_NC) nc = BROW_NC;
        return (chunk_ok && nc > 1) ? nc : 1;
    };
    auto Bscr = [&](int K_, int J_, int c_) {
        return PB + ((size_t)32 * (nM - 1) + (size_t)((((K_ & 1) * (nM - 2) + J_) * BROW_NC + c_) * 16)) * TILE_ELEMS;
    };
    for (int K = nM - 1; K >= 1; --K) {
        const int ncK = nchunks(K);
        ntasks += (nM - K) + 1 + (K >= 2 ? 1 + (K - 1) * (ncK > 1 ? ncK + 1 : 1) + (K - 1) * (nM - 1 - K) : 0);
    }
    int* hpcnt = C.head + 1;   // hyper-gradient partials written so far
    int* qsflag = C.fail;      // qsflag[J]: the Q partials of column J have been summed into the Q block of row J
    Acc acc;
    for (;;) {
        const int t = wq_next(C.head, &s_task);
        if (t >= ntasks) break;
        int type = 0, I = 0, J = 0, K = 0, ch = 0;   // 0 ADIAG(J), 1 APANEL(I,J), 2 BROW(I=K,J), 3 SC(I,J;K), 4 QSUM(J), 5 BROW chunk (I=K,J,ch)
        if (t == 0) { type = 0; J = nM - 1; }
        else {
            int u = t - 1;
            for (K = nM - 1;; --K) {
                const int na = nM - K;
                if (u < na) { type = 1; I = K + u; J = K - 1; break; }
                u -= na;
                if (u == 0) { type = 0; J = K - 1; break; }
                u -= 1;
                if (K >= 2) {
                    if (u == 0) { type = 4; J = K - 1; break; }
                    u -= 1;
                }
                const int nb = K >= 2 ? K - 1 : 0;
                const int ncK = nchunks(K);
                if (ncK > 1) {
                    if (u < nb * ncK) { type = 5; I = K; J = K - 2 - u / ncK; ch = u % ncK; break; }
                    u -= nb * ncK;
                }
                if (u < nb) { type = 2; I = K; J = K - 2 - u; break; }
                u -= nb;
                const int nc = nb * (nM - 1 - K);
                if (u < nc) { type = 3; J = K - 2 - u / (nM - 1 - K); I = K + 1 + u % (nM - 1 - K); break; }
                u -= nc;
            }
        }
        double gl = 0.0, gv = 0.0, gb = 0.0;
        if (type == 1) {
            // ---- panel tile (I, J) of column J = K - 1
            WT_DECL();
            WT_COUNT(I == K ? 16 : 24);
            const double* blk = LinvD + (size_t)J * 16 * TILE_ELEMS;
            const Opnd opInvCol = opnd_block(blk, ACC_COLK, TPM * J, TPM * J, 1);
            if (I == K) {
                // E = Lbar[K,J] - Atilde_KK Lc[K,J] - sum_{k>K} Q_k   (Q_k = Atilde[k,K]^T Lc[k,J], left by the panel tasks of column K)
                if (tid == 0) {
                    spin_ge(&C.fin[tix(K, K)], 1);
                    if (K + 1 < nM) spin_ge(&qsflag[K], 1);
                }
                deps_ready();
                WT_LAP(17);
                acc_zero(acc);
                gemm_macro(acc, pipe, opAsym, TPM * I, opLcol, TPM * J, TPM * K, min(TPM * K + TPM, nT), nullptr);
                acc_foreach(acc, I, J, [&](int it, int jt, int i, int j, double v0, double v1) {
                    double* tp = LB + packed_tile(it, jt) * TILE_ELEMS;
                    const double2 l = tile_load2_cg(tp, i, j);
                    const size_t qo = (size_t)((it - TPM * I) * TPM + (jt - TPM * J)) * TILE_ELEMS;
                    double2 q = make_double2(0.0, 0.0);
                    if (K + 1 < nM) q = tile_load2_cg(Qscr(K) + qo, i, j);   // sum_{k>K} Q_k, summed in row order by QSUM(K)
                    tile_store2(tp, i, j, l.x - v0 - q.x, l.y - v1 - q.y);
                });
            } else {
                // E = (Lbar[I,J] with the scatters of columns > K already applied) - Atilde[I,K] Lc[K,J]
                if (tid == 0) {
                    spin_ge(&C.fin[tix(I, K)], 1);
                    if (I - J - 1 > 0) spin_ge(&C.cnt[tix(I, J)], I - J - 1);
                }
                deps_ready();
                WT_LAP(25);
                acc_zero(acc);
                gemm_macro(acc, pipe, opAsym, TPM * I, opLcol, TPM * J, TPM * K, min(TPM * K + TPM, nT), nullptr);
                acc_foreach(acc, I, J, [&](int it, int jt, int i, int j, double v0, double v1) {
                    double* tp = LB + packed_tile(it, jt) * TILE_ELEMS;
                    const double2 l = tile_load2_cg(tp, i, j);
                    tile_store2(tp, i, j, l.x - v0, l.y - v1);
                });
            }
            fence_proxy_async();
            WT_LAP(I == K ? 18 : 26);
            // Atilde[I,J] = E inv(L_JJ)
            acc_zero(acc);
            gemm_macro(acc, pipe, opArow, TPM * I, opInvCol, TPM * J, TPM * J, TPM * J + TPM, nullptr, 0);   // the skipped operand is B
            acc_foreach(acc, I, J, [&](int it, int jt, int i, int j, double v0, double v1) {
                tile_store2(LB + packed_tile(it, jt) * TILE_ELEMS, i, j, v0, v1);
            }, 0);
            fence_proxy_async();
            WT_LAP(I == K ? 19 : 27);
            // P_I = Atilde[I,J]^T Lc[I,J]  (lower warp tiles) for the diagonal task of column J
            acc_zero(acc);
            gemm_macro(acc, pipe, opAcol, TPM * J, opLcol, TPM * J, TPM * I, min(TPM * I + TPM, nT), nullptr, 2);
            // the partial block of row I still holds column K's contribution until the diagonal task of column K has summed it
            // (rows run ahead of each other: nothing else ties row I to the diagonal of column K)
            if (I > K) {
                if (tid == 0) spin_ge(&C.fin[tix(K, K)], 1);
                __syncthreads();
            }
            acc_foreach(acc, J, J, [&](int it, int jt, int i, int j, double v0, double v1) {
                if (it < jt) return;
                tile_store2(Pscr(I) + (size_t)((it - TPM * J) * TPM + (jt - TPM * J)) * TILE_ELEMS, i, j, v0, v1);
            }, 2);
            task_signal_add(&C.col[J]);
            WT_LAP(I == K ? 20 : 28);
            // Q_I = Atilde[I,J]^T Lc[I,J-1]  for the panel tile (J, J-1) of the next column
            if (J >= 1) {
                acc_zero(acc);
                gemm_macro(acc, pipe, opAcol, TPM * J, opLcol, TPM * (J - 1), TPM * I, min(TPM * I + TPM, nT), nullptr);
                if (I > K) {   // likewise the Q block of row I is read by the panel task (K, J) of the previous column
                    if (tid == 0) spin_ge(&C.fin[tix(K, J)], 1);
                    __syncthreads();
                }
                acc_foreach(acc, J, J - 1, [&](int it, int jt, int i, int j, double v0, double v1) {
                    tile_store2(Qscr(I) + (size_t)((it - TPM * J) * TPM + (jt - TPM * (J - 1))) * TILE_ELEMS, i, j, v0, v1);
                });
            }
            task_signal(&C.fin[tix(I, J)], 1);
            // off the dependency chain: this tile's share of sum(Atilde o dK/dtheta), read back from the finished tile
            if (want_hyp) {
                acc_foreach(acc, I, J, [&](int it, int jt, int i, int j, double, double) {
                    const double2 w = tile_load2(LB + packed_tile(it, jt) * TILE_ELEMS, i, j);
                    const int gi = it * TS + i, gj = jt * TS + j;
                    hyper_accum(P, ka, dkal, dkab, KT, gi, gj, w.x, kvar, kinv, gl, gv, gb);
                    hyper_accum(P, ka, dkal, dkab, KT, gi, gj + 1, w.y, kvar, kinv, gl, gv, gb);
                });
                gl = block_sum(gl, red); gv = block_sum(gv, red); gb = block_sum(gb, red);
                if (tid == 0) {
                    double* h3 = hp + (size_t)tix(I, J) * 3;
                    h3[0] = gl; h3[1] = gv; h3[2] = gb;
                    __threadfence();
                    atomicAdd(hpcnt, 1);
                }
            }
            WT_LAP(I == K ? 21 : 29);
        } else if (type == 0) {
            // ---- diagonal block J:  Leff = tril(Lbar_JJ) - tril(sum_{I>J} P_I);  Atilde_JJ = sym(inv(L)^T Phi(L^T Leff) inv(L))
            const double* blk = LinvD + (size_t)J * 16 * TILE_ELEMS;
            const Opnd opInvCol = opnd_block(blk, ACC_COLK, TPM * J, TPM * J, 1);
            WT_DECL();
            WT_COUNT(32);
            if (tid == 0) spin_ge(&C.col[J], nM - 1 - J);
            deps_ready();
            WT_LAP(33);
            {
                // 10 lower tiles x 1024 elements = 20 per thread, kept in registers while the partial blocks stream by: the loads of
                // one partial are independent of each other (the element-by-element version spent 24 us here, latency bound)
                constexpr int PER = 10 * TILE_ELEMS / GEMM_THREADS;
                double val[PER];
                int off[PER];
                unsigned live = 0;
#pragma unroll
                for (int r = 0; r < PER; ++r) {
                    const int e = tid + GEMM_THREADS * r, tl = e >> 10, o = e & (TILE_ELEMS - 1);
                    int a = 0;
                    while ((a + 1) * (a + 2) / 2 <= tl) ++a;
                    const int b = tl - a * (a + 1) / 2;
                    const int i = o >> 5, j = (o & 31) ^ ((i & 3) << 2);   // storage position o holds element (i, j)
                    off[r] = (a * TPM + b) * TILE_ELEMS + o;
                    const bool use = b < a || j <= i;
                    live |= (use ? 1u : 0u) << r;
                    val[r] = use ? LB[packed_tile(TPM * J + a, TPM * J + b) * TILE_ELEMS + o] : 0.0;
                }
                for (int I2 = J + 1; I2 < nM; ++I2) {
                    const double* pp = Pscr(I2);
                    double pv[PER];
#pragma unroll
                    for (int r = 0; r < PER; ++r) pv[r] = __ldcg(pp + off[r]);
#pragma unroll
                    for (int r = 0; r < PER; ++r) val[r] -= pv[r];
                }
#pragma unroll
                for (int r = 0; r < PER; ++r) S0[off[r]] = ((live >> r) & 1u) ? val[r] : 0.0;
                // the six tiles above the diagonal are structural zeros of Leff
                for (int e = tid; e < 6 * TILE_ELEMS; e += GEMM_THREADS) {
                    const int t6 = e >> 10, o = e & (TILE_ELEMS - 1);
                    const int a = t6 < 3 ? 0 : (t6 < 5 ? 1 : 2);
                    const int b = t6 < 3 ? t6 + 1 : (t6 < 5 ? t6 - 1 : 3);
                    S0[(size_t)(a * TPM + b) * TILE_ELEMS + o] = 0.0;
                }
            }
            fence_proxy_async();
            WT_LAP(34);
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
            WT_LAP(35);
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
                    }
                }
            task_signal(&C.fin[tix(J, J)], 1);
            WT_LAP(36);
            if (want_hyp) {
                // off the dependency chain: the diagonal block's share of sum(Atilde o dK/dtheta)
                for (int a = 0; a < TPM; ++a)
                    for (int b = 0; b <= a; ++b) {
                        const double* tp = LB + packed_tile(TPM * J + a, TPM * J + b) * TILE_ELEMS;
                        for (int e = tid; e < TILE_ELEMS; e += GEMM_THREADS) {
                            const int i = e >> 5, j = e & 31;
                            hyper_accum(P, ka, dkal, dkab, KT, (TPM * J + a) * TS + i, (TPM * J + b) * TS + j, tp[tswz(i, j)], kvar, kinv, gl, gv, gb);
                        }
                    }
                gl = block_sum(gl, red); gv = block_sum(gv, red); gb = block_sum(gb, red);
                if (tid == 0) { double* h3 = hp + (size_t)tix(J, J) * 3; h3[0] = gl; h3[1] = gv; h3[2] = gb; }
                if (J == 0) {
                    // last task of the phase: wait until every other tile has published its partial, then sum them in a fixed order
                    const int nt = nM * (nM + 1) / 2;
                    if (tid == 0) spin_ge(hpcnt, nt - 1);
                    __syncthreads();
                    double a0 = 0.0, a1 = 0.0, a2 = 0.0;
                    for (int e = tid; e < nt; e += GEMM_THREADS) {
                        a0 += __ldcg(hp + (size_t)e * 3 + 0);
                        a1 += __ldcg(hp + (size_t)e * 3 + 1);
                        a2 += __ldcg(hp + (size_t)e * 3 + 2);
                    }
                    a0 = block_sum(a0, red); a1 = block_sum(a1, red); a2 = block_sum(a2, red);
                    if (tid == 0) {
                        P.grad_hyp[(size_t)item * 4 + 0] = a0;
                        P.grad_hyp[(size_t)item * 4 + 1] = a1;
                        P.grad_hyp[(size_t)item * 4 + 3] = a2;
                    }
                } else if (tid == 0) {
                    __threadfence();
                    atomicAdd(hpcnt, 1);
                }
            } else if (J == 0 && tid == 0) {
                P.grad_hyp[(size_t)item * 4 + 0] = 0.0;
                P.grad_hyp[(size_t)item * 4 + 1] = 0.0;
                P.grad_hyp[(size_t)item * 4 + 3] = 0.0;
            }
        } else if (type == 4) {
            // ---- QSUM(J): Q block of row J <- sum_{k>J} Q_k (row order), beside the diagonal task of column J instead of inside the
            // panel task (J, J-1) that needs it
            if (tid == 0)
                for (int k = J + 1; k < nM; ++k) spin_ge(&C.fin[tix(k, J)], 1);
            __syncthreads();
            constexpr int PERQ = 16 * TILE_ELEMS / (2 * GEMM_THREADS);   // double2 per thread
            double2 q[PERQ];
#pragma unroll
            for (int r = 0; r < PERQ; ++r) q[r] = make_double2(0.0, 0.0);
            for (int k = J + 1; k < nM; ++k) {
                const double2* src = reinterpret_cast<const double2*>(Qscr(k));
                double2 v[PERQ];
#pragma unroll
                for (int r = 0; r < PERQ; ++r) v[r] = __ldcg(src + tid + GEMM_THREADS * r);
#pragma unroll
                for (int r = 0; r < PERQ; ++r) { q[r].x += v[r].x; q[r].y += v[r].y; }
            }
            double2* dstq = reinterpret_cast<double2*>(Qscr(J));
#pragma unroll
            for (int r = 0; r < PERQ; ++r) dstq[tid + GEMM_THREADS * r] = q[r];
            task_signal(&qsflag[J], 1);
        } else if (type == 5) {
            // ---- chunk `ch` of BROW(K, J): partial product over its share of the k range -> scratch block
            const int ncK = nchunks(I);
            const int klen = nT - TPM * I, per = (klen + ncK - 1) / ncK;
            const int k0 = TPM * I + ch * per, k1 = min(k0 + per, nT);
            if (tid == 0) {
                spin_ge(&C.fin[tix(I, I)], 1);
                for (int k = I + 1; k < nM; ++k) spin_ge(&C.fin[tix(k, I)], 1);
                // the scratch block was last used two rows up (same parity): its BROW task must have consumed it
                if (I + 2 <= nM - 1 && nchunks(I + 2) > ch) spin_ge(&C.cnt[tix(I + 2, J)], 1);
            }
            deps_ready();
            acc_zero(acc);
            gemm_macro(acc, pipe, opAsym, TPM * I, opLcol, TPM * J, k0, k1, nullptr);
            double* blkp = Bscr(I, J, ch);
            acc_foreach(acc, I, J, [&](int it, int jt, int i, int j, double v0, double v1) {
                tile_store2(blkp + (size_t)((it - TPM * I) * TPM + (jt - TPM * J)) * TILE_ELEMS, i, j, v0, v1);
            });
            task_signal_add(&C.aux[tix(I, J)]);
        } else if (type == 2) {
            WT_DECL();
            WT_COUNT(38);
            // ---- BROW(K, J): tile (K,J) -= Atilde_sym[K, K:] Lc[K:, J]   (one long product over the rows >= K of column K, or the sum
            // of its chunks when they ran as separate tasks)
            const int ncK = nchunks(I);
            if (ncK > 1) {
                if (tid == 0) spin_ge(&C.aux[tix(I, J)], ncK);
                __syncthreads();
                WT_LAP(39);
                constexpr int PERB = 16 * TILE_ELEMS / (2 * GEMM_THREADS);   // double2 per thread
#pragma unroll 4
                for (int r = 0; r < PERB; ++r) {
                    const int e2 = tid + GEMM_THREADS * r, tl = e2 >> 9, o2 = e2 & 511;
                    double2* tp = reinterpret_cast<double2*>(LB + packed_tile(TPM * I + (tl >> 2), TPM * J + (tl & 3)) * TILE_ELEMS) + o2;
                    double2 v = __ldcg(tp);
                    for (int c2 = 0; c2 < ncK; ++c2) {
                        const double2 pv = __ldcg(reinterpret_cast<const double2*>(Bscr(I, J, c2)) + e2);
                        v.x -= pv.x; v.y -= pv.y;
                    }
                    *tp = v;
                }
            } else {
                if (tid == 0) {
                    spin_ge(&C.fin[tix(I, I)], 1);
                    for (int k = I + 1; k < nM; ++k) spin_ge(&C.fin[tix(k, I)], 1);
                }
                deps_ready();
                WT_LAP(39);
                acc_zero(acc);
                gemm_macro(acc, pipe, opAsym, TPM * I, opLcol, TPM * J, TPM * I, nT, nullptr);
                acc_foreach(acc, I, J, [&](int it, int jt, int i, int j, double v0, double v1) {
                    double* tp = LB + packed_tile(it, jt) * TILE_ELEMS;
                    const double2 l = tile_load2_cg(tp, i, j);
                    tile_store2(tp, i, j, l.x - v0, l.y - v1);
                });
            }
            task_signal(&C.cnt[tix(I, J)], 1);
            WT_LAP(40);
        } else {
            WT_DECL();
            WT_COUNT(42);
            // ---- SC(I, J; K): tile (I,J) -= Atilde[I,K] Lc[K,J],  I > K > J + 1
            if (tid == 0) {
                spin_ge(&C.fin[tix(I, K)], 1);
                spin_ge(&C.cnt[tix(I, J)], I - K);
            }
            deps_ready();
            WT_LAP(43);
            acc_zero(acc);
            gemm_macro(acc, pipe, opAsym, TPM * I, opLcol, TPM * J, TPM * K, min(TPM * K + TPM, nT), nullptr);
            acc_foreach(acc, I, J, [&](int it, int jt, int i, int j, double v0, double v1) {
                double* tp = LB + packed_tile(it, jt) * TILE_ELEMS;
                const double2 l = tile_load2_cg(tp, i, j);
                tile_store2(tp, i, j, l.x - v0, l.y - v1);
            });
            task_signal(&C.cnt[tix(I, J)], I - K + 1);
            WT_LAP(44);
        }
    }
}

void dense_read_wide_timers(unsigned long long* out48, int reset) {
#ifdef BGP_TIMING
    cudaMemcpyFromSymbol(out48, g_wide_timers, 48 * sizeof(unsigned long long));
    if (reset) { unsigned long long z[48] = {0}; cudaMemcpyToSymbol(g_wide_timers, z, sizeof(z)); }
#else
    for (int i = 0; i < 48; ++i) out48[i] = 0;
    (void)reset;
#endif
}

// ------------------------------------------------------------------------------------------------ host launchers
static cudaError_t set_smem_w(const void* f) { return cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES); }

cudaError_t dense_wide_init_kernels() {
    cudaError_t e;
    if ((e = set_smem_w((const void*)k_wchol<0>)) != cudaSuccess) return e;
    if ((e = set_smem_w((const void*)k_wchol<1>)) != cudaSuccess) return e;
    if ((e = set_smem_w((const void*)k_wsyrk)) != cudaSuccess) return e;
    if ((e = set_smem_w((const void*)k_wtrtri)) != cudaSuccess) return e;
    if ((e = set_smem_w((const void*)k_wlauum)) != cudaSuccess) return e;
    if ((e = set_smem_w((const void*)k_wtrmm)) != cudaSuccess) return e;
    if ((e = set_smem_w((const void*)k_wadj)) != cudaSuccess) return e;
    return cudaSuccess;
}

// Same phase numbering as dense_launch_phase.  Returns false for a phase that has no wide kernel.
bool dense_launch_phase_wide(const DenseParams& P, int nitems, int phase, int G, cudaStream_t st) {
    const dim3 grid(G, nitems);
    switch (phase) {
        case 1: k_wchol<0><<<grid, GEMM_THREADS, SMEM_BYTES, st>>>(P); return true;
        case 2: k_wsyrk<<<grid, GEMM_THREADS, SMEM_BYTES, st>>>(P); return true;
        case 3: k_wchol<1><<<grid, GEMM_THREADS, SMEM_BYTES, st>>>(P); return true;
        case 4: k_wsolve<<<grid, WS_THREADS, 0, st>>>(P); return true;
        case 8: k_wpost<<<grid, WS_THREADS, 0, st>>>(P); return true;
        case 5: k_wtrtri<<<grid, GEMM_THREADS, SMEM_BYTES, st>>>(P); return true;
        case 6: k_wlauum<<<grid, GEMM_THREADS, SMEM_BYTES, st>>>(P); return true;
        case 7: k_wtrmm<<<grid, GEMM_THREADS, SMEM_BYTES, st>>>(P); return true;
        case 9: k_wadj<<<grid, GEMM_THREADS, SMEM_BYTES, st>>>(P); return true;
        default: return false;
    }
}

}  // namespace bgp
