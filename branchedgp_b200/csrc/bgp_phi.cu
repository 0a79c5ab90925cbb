// Assignment passes over the N x 3N variational matrix logPhi (HBM traffic + fp64 transcendentals: the per-element exp / log are
// the tailored routines of bgp_math.cuh).
//
// forward  (assigngp_dense.py:160-162, 242-245; MBGP/assigngp.py:217-251, 688-691):
//     S = softmax_rows(logPhi);  Phi = (1-2e-6) S + 1e-6;  A_j = sum_i Phi_ij;  v_jd = sum_i Phi_ij Y_id;
//     KL = sum Phi log Phi - sum Phi log pZ,  with pZ never materialised: it is a function of
//     (i, j, t_i <= b, prior[i]) -- pZ_construction_singleBP.py:6-25.
// backward (SURVEY.md App. B.1 last line):
//     Sbar = a (Abar_j + sum_d Y_id vbar_jd - w (log Phi + 1 - log pZ));  dlogPhi = S (Sbar - sum_k S_ik Sbar_ik)
#include <cooperative_groups.h>
#include <cstdlib>

#include "bgp_dense.cuh"
#include "bgp_math.cuh"
#include "bgp_phi.cuh"
#include "bgp_launch.h"

namespace bgp {

__device__ __forceinline__ double trunk_value(bool inblock, int f) {   // phiTrunk, MBGP/assigngp.py:399-405
    return inblock ? (f == 0 ? (1.0 - 2e-9) : 1e-9) : 0.0;
}

// squashed assignment probability given the softmax value
__device__ __forceinline__ double phi_of(int model, bool act, bool inblock, int f, double S) {
    double w = S;
    if (model == MODEL_MBGP && !act) w = trunk_value(inblock, f);
    return SQ_A * w + SQ_B;
}

__device__ __forceinline__ double log_pz(int model, bool act, bool inblock, int f, const double* __restrict__ prior, int row,
                                         double log_eps) {
    if (model == MODEL_BGP) {
        if (!inblock) return log_eps;
        if (!act) return f == 0 ? 0.0 : log_eps;                 // trunk rows: [1, 1e-6, 1e-6]
        return f == 0 ? log_eps : log(prior[row * 2 + (f - 1)]);  // branch rows: [1e-6, p0, p1]
    }
    double w;
    if (act) w = inblock ? prior[row * 3 + f] : 1e-6;             // eZ0, MBGP/assigngp.py:25-40
    else w = trunk_value(inblock, f);
    return log(SQ_A * w + SQ_B);
}

// PHI_DCH = outputs accumulated per sweep of the column pass: 1 for single-output models (every BGP FitModel call), else 4.
template <int PHI_DCH>
__global__ void __launch_bounds__(NTHREADS) k_phi_forward(PhiParams P) {
    __shared__ double lse_s[PHI_RCH];
    __shared__ double Ys[PHI_RCH][PHI_DCH];
    __shared__ double red[32];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int chunk = blockIdx.x, slot = blockIdx.y, item = P.item0 + slot / P.SD, ycol = (P.SD > 1) ? slot % P.SD : 0;
    const int N = P.N, M = P.M, Mp = P.Mp, D = P.D;
    const int r0 = chunk * PHI_RCH;
    const int nr = min(PHI_RCH, N - r0);
    const double* lp = P.logPhi + ((size_t)item * N + r0) * M;
    const double b = P.bv[(size_t)P.item0 * P.SD + slot];
    // The gradient buffer of the item doubles as a stash between the passes (bound + gradient calls only): pass 1 leaves
    // e = exp(x - max) there, pass 2 turns it into the softmax value S, and the backward kernel starts from S -- one exp per
    // element in the whole evaluation instead of three.  With several sub-problems per item (MBGP multi-output, SD > 1) the
    // CTAs of different sub-problems share the rows, so only S is stashed (by sub-problem 0) and pass 2 recomputes the exp.
    double* stash = P.grad_logPhi ? P.grad_logPhi + ((size_t)item * N + r0) * M : nullptr;
    const bool stash_e = stash != nullptr && P.SD == 1;
    const bool stash_S = stash != nullptr && (P.SD == 1 || slot % P.SD == 0);
    __shared__ double inv_s[PHI_RCH];
    __shared__ int act_s[PHI_RCH];
    if (tid < PHI_RCH) act_s[tid] = (tid < nr && P.t[r0 + tid] > b) ? 1 : 0;
    const int model = P.model;
    const double* prior = P.prior;
    // ---- row log-sum-exp, one warp per row (max pass, then exp-sum pass; a branchy one-pass online variant measured slower)
    for (int rr = warp; rr < nr; rr += NTHREADS / 32) {
        const double* row = lp + (size_t)rr * M;
        double m = -1.0e308;
        for (int j = lane; j < M; j += 32) m = fmax(m, row[j]);
        m = warp_max(m);
        double s = 0.0;
        if (stash_e) {
            double* srow = stash + (size_t)rr * M;
            for (int j = lane; j < M; j += 32) { const double e = fast_exp(row[j] - m); srow[j] = e; s += e; }
        } else {
            for (int j = lane; j < M; j += 32) s += fast_exp(row[j] - m);
        }
        s = warp_sum(s);
        const double l = m + log(s);
        if (lane == 0) {
            lse_s[rr] = l;
            inv_s[rr] = 1.0 / s;
            P.lse[(size_t)slot * P.slot_stride + r0 + rr] = l;
        }
    }
    __syncthreads();
    // ---- column pass: thread per column, rows of the chunk in the inner loop
    const double log_eps = log(1e-6);
    const double* Yg = P.Y + (size_t)P.item_gene[item] * N * P.Dtot + ycol;
    double kl = 0.0;
    for (int d0 = 0; d0 < D; d0 += PHI_DCH) {
        const int nd = min(PHI_DCH, D - d0);
        __syncthreads();
        if (tid < PHI_RCH * PHI_DCH) {
            const int rr = tid / PHI_DCH, dd = tid % PHI_DCH;
            Ys[rr][dd] = (rr < nr && dd < nd) ? Yg[(size_t)(r0 + rr) * P.Dtot + d0 + dd] : 0.0;
        }
        __syncthreads();
        for (int j = tid; j < Mp; j += NTHREADS) {
            double a = 0.0, vv[PHI_DCH];
#pragma unroll
            for (int dd = 0; dd < PHI_DCH; ++dd) vv[dd] = 0.0;
            if (j < M) {
                for (int rr = 0; rr < nr; ++rr) {
                    const int row = r0 + rr;
                    const int f = j - 3 * row;             // column 3*row + f belongs to the row's own cell
                    const bool inblock = (f >= 0 && f < 3);
                    const bool act = act_s[rr] != 0;
                    double S;
                    if (stash_e) {
                        S = stash[(size_t)rr * M + j];
                        if (d0 == 0) { S *= inv_s[rr]; stash[(size_t)rr * M + j] = S; }   // later output chunks find S itself
                    } else {
                        S = fast_exp(lp[(size_t)rr * M + j] - lse_s[rr]);
                        if (stash_S && d0 == 0) stash[(size_t)rr * M + j] = S;
                    }
                    const double phi = phi_of(model, act, inblock, f, S);
                    a += phi;
#pragma unroll
                    for (int dd = 0; dd < PHI_DCH; ++dd) vv[dd] += phi * Ys[rr][dd];
                    if (d0 == 0) kl += phi * (fast_log(phi) - log_pz(model, act, inblock, f, prior, row, log_eps));
                }
            }
            if (d0 == 0) P.Apart[(size_t)slot * P.slot_stride + (size_t)chunk * Mp + j] = a;
            for (int dd = 0; dd < nd; ++dd) P.vpart[(size_t)slot * P.slot_stride + ((size_t)chunk * D + d0 + dd) * Mp + j] = vv[dd];
        }
    }
    kl = block_sum(kl, red);
    if (tid == 0) P.klpart[(size_t)slot * P.slot_stride + chunk] = kl;
}


// ------------------------------------------------------------------------------------------- MBGP multi-output forward pass
// The outputs of the joint model share logPhi; only the trunk mask t_i > b_d differs (MBGP/assigngp.py:217-251).  The reference (and
// k_phi_forward with one CTA column per sub-problem) redoes softmax, exp and log for every output; here a CTA (16 rows, one item)
// does them once, keeps the 16 active-form values of its columns in registers and then only selects and accumulates per output:
//   A_d[j] = sum_i (t_i > b_d ? Phi_ij : PhiTrunk_ij),  v_d[j] = sum_i (...) Y_id,  KL_d = sum_{i: t_i > b_d} KLrow_i
// (a trunk row contributes exactly 0 to KL: its Phi and pZ are the same constant).  grid (row chunks, items), SD <= MO_MAXSD.
constexpr int MO_MAXSD = 256;
__global__ void __launch_bounds__(NTHREADS) k_phi_forward_multi(PhiParams P) {
    __shared__ double lse_s[PHI_RCH], inv_dummy[1], t_s[PHI_RCH], klrow_s[PHI_RCH], red[32];
    __shared__ double bv_s[MO_MAXSD];
    __shared__ double Ys[PHI_RCH][MO_MAXSD];
    (void)inv_dummy;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int chunk = blockIdx.x, litem = blockIdx.y, item = P.item0 + litem;
    const int N = P.N, M = P.M, Mp = P.Mp, SD = P.SD;
    const int r0 = chunk * PHI_RCH, nr = min(PHI_RCH, N - r0);
    const size_t slot0 = (size_t)litem * SD;
    const double* lp = P.logPhi + ((size_t)item * N + r0) * M;
    double* stash = P.grad_logPhi ? P.grad_logPhi + ((size_t)item * N + r0) * M : nullptr;
    const double* Yg = P.Y + (size_t)P.item_gene[item] * N * P.Dtot;
    const double* prior = P.prior;
    for (int e = tid; e < PHI_RCH * SD; e += NTHREADS) {
        const int rr = e / SD, sd = e - rr * SD;
        Ys[rr][sd] = rr < nr ? Yg[(size_t)(r0 + rr) * P.Dtot + sd] : 0.0;
    }
    for (int sd = tid; sd < SD; sd += NTHREADS) bv_s[sd] = P.bv[(size_t)item * SD + sd];
    if (tid < PHI_RCH) t_s[tid] = tid < nr ? P.t[r0 + tid] : -1.0e308;   // padding rows are trunk rows of every output ...
    // ---- row log-sum-exp, one warp per row
    for (int rr = warp; rr < nr; rr += NTHREADS / 32) {
        const double* row = lp + (size_t)rr * M;
        double m = -1.0e308;
        for (int j = lane; j < M; j += 32) m = fmax(m, row[j]);
        m = warp_max(m);
        double sacc = 0.0;
        for (int j = lane; j < M; j += 32) sacc += fast_exp(row[j] - m);
        sacc = warp_sum(sacc);
        const double l = m + log(sacc);
        if (lane == 0) {
            lse_s[rr] = l;
            P.lse[slot0 * P.slot_stride + r0 + rr] = l;
        }
    }
    __syncthreads();
    const double lpz_off = log(SQ_A * 1e-6 + SQ_B);   // log pZ of every column outside the row's own cell (eZ0 = 1e-6 there)
    double klr[PHI_RCH];
#pragma unroll
    for (int rr = 0; rr < PHI_RCH; ++rr) klr[rr] = 0.0;
    for (int j = tid; j < Mp; j += NTHREADS) {
        double pa[PHI_RCH], pt[PHI_RCH];   // active-form / trunk-form squashed probabilities of column j
#pragma unroll
        for (int rr = 0; rr < PHI_RCH; ++rr) {
            pa[rr] = 0.0; pt[rr] = 0.0;    // ... and contribute nothing
            if (rr < nr && j < M) {
                const int row = r0 + rr;
                const int f = j - 3 * row;
                const bool inblock = (f >= 0 && f < 3);
                const double S = fast_exp(lp[(size_t)rr * M + j] - lse_s[rr]);
                if (stash) stash[(size_t)rr * M + j] = S;
                const double phi = SQ_A * S + SQ_B;
                pa[rr] = phi;
                pt[rr] = SQ_A * trunk_value(inblock, f) + SQ_B;
                const double lpz = inblock ? log_pz(MODEL_MBGP, true, true, f, prior, row, 0.0) : lpz_off;
                klr[rr] += phi * (fast_log(phi) - lpz);
            }
        }
        for (int sd = 0; sd < SD; ++sd) {
            const double b = bv_s[sd];
            double a = 0.0, v = 0.0;
#pragma unroll
            for (int rr = 0; rr < PHI_RCH; ++rr) {
                const double p = t_s[rr] > b ? pa[rr] : pt[rr];
                a += p;
                v = fma(p, Ys[rr][sd], v);
            }
            P.Apart[(slot0 + sd) * P.slot_stride + (size_t)chunk * Mp + j] = a;
            P.vpart[(slot0 + sd) * P.slot_stride + (size_t)chunk * Mp + j] = v;
        }
    }
#pragma unroll
    for (int rr = 0; rr < PHI_RCH; ++rr) {
        const double x = block_sum(klr[rr], red);
        if (tid == 0) klrow_s[rr] = x;
    }
    __syncthreads();
    for (int sd = tid; sd < SD; sd += NTHREADS) {
        const double b = bv_s[sd];
        double kl = 0.0;
        for (int rr = 0; rr < nr; ++rr) kl += t_s[rr] > b ? klrow_s[rr] : 0.0;
        P.klpart[(slot0 + sd) * P.slot_stride + chunk] = kl;
    }
}

// ---------------------------------------------------------------------------------------------------- row-resident forward
// Bound + gradient calls with one sub-problem per item: the forward pass is split into a row pass and a column pass.
//   k_phi_rowpass: one CTA (or a cluster of CL CTAs) owns one row of logPhi, which it reads ONCE into registers: max, e = exp(x - max),
//                  sum, S = e / sum (written to the stash in the gradient buffer), Phi, the row's KL terms.  One exp and one log per
//                  element; a cluster combines its CTAs' (max, sum) pairs through distributed shared memory.
//   k_phi_colpass: thread per column over 16-row chunks of the stash: A = sum_i Phi, v = Phi^T Y (no transcendentals).
// DRAM traffic per element: 8 B read + 8 B written + 8 B read, against 40 B for k_phi_forward at rows longer than the L2 can hold.
template <int BT, int EPT, int MINB, int CL>
__global__ void __launch_bounds__(BT, MINB) k_phi_rowpass(PhiParams P) {
    __shared__ double red[32];
    __shared__ double xch[2];
    const int tid = threadIdx.x;
    const int rank = CL > 1 ? (int)(blockIdx.x % CL) : 0;
    const int row = CL > 1 ? (int)(blockIdx.x / CL) : (int)blockIdx.x;
    const int slot = blockIdx.y, item = P.item0 + slot;
    const int N = P.N, M = P.M;
    const int j0 = rank * BT * EPT;
    const int model = P.model;
    const double* prior = P.prior;
    const double* x = P.logPhi + ((size_t)item * N + row) * M;
    double* stash = P.grad_logPhi + ((size_t)item * N + row) * M;
    const bool act = P.t[row] > P.bv[item];
    const double log_eps = log(1e-6);
    double e[EPT];
    double m = -1.0e308;
#pragma unroll
    for (int k = 0; k < EPT; ++k) {
        const int j = j0 + tid + k * BT;
        e[k] = (j < M) ? x[j] : -1.0e308;
        m = fmax(m, e[k]);
    }
    // block maximum
    m = warp_max(m);
    __syncthreads();
    if ((tid & 31) == 0) red[tid >> 5] = m;
    __syncthreads();
    m = (tid & 31) < BT / 32 ? red[tid & 31] : -1.0e308;
    m = warp_max(m);
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < EPT; ++k) {
        const int j = j0 + tid + k * BT;
        e[k] = (j < M) ? fast_exp(e[k] - m) : 0.0;
        s += e[k];
    }
    s = block_sum(s, red);
    double scale = 1.0 / s, lse = m + log(s);
    if (CL > 1) {   // combine the (max, sum) pairs of the row's CTAs in rank order
        namespace cg = cooperative_groups;
        cg::cluster_group cluster = cg::this_cluster();
        if (tid == 0) { xch[0] = m; xch[1] = s; }
        cluster.sync();
        double mg = -1.0e308;
        for (int c = 0; c < CL; ++c) mg = fmax(mg, cluster.map_shared_rank(xch, c)[0]);
        double sg = 0.0;
        for (int c = 0; c < CL; ++c) { const double* o = cluster.map_shared_rank(xch, c); sg += o[1] * exp(o[0] - mg); }
        cluster.sync();   // nobody leaves (or reuses xch) while it may still be read
        scale = exp(m - mg) / sg;
        lse = mg + log(sg);
    }
    if (tid == 0 && rank == 0) P.lse[(size_t)slot * P.slot_stride + row] = lse;
    double kl = 0.0;
#pragma unroll
    for (int k = 0; k < EPT; ++k) {
        const int j = j0 + tid + k * BT;
        if (j < M) {
            const double S = e[k] * scale;
            stash[j] = S;
            const int f = j - 3 * row;
            const bool inblock = (f >= 0 && f < 3);
            const double phi = phi_of(model, act, inblock, f, S);
            kl += phi * (fast_log(phi) - log_pz(model, act, inblock, f, prior, row, log_eps));
        }
    }
    kl = block_sum(kl, red);
    if (tid == 0) P.klrow[(size_t)slot * P.slot_stride + (size_t)row * PHI_KLW + rank] = kl;
}

template <int PHI_DCH>
__global__ void __launch_bounds__(NTHREADS) k_phi_colpass(PhiParams P, int ncl) {
    __shared__ double Ys[PHI_RCH][PHI_DCH];
    __shared__ int act_s[PHI_RCH];
    const int tid = threadIdx.x;
    const int chunk = blockIdx.x, slot = blockIdx.y, item = P.item0 + slot;
    const int N = P.N, M = P.M, Mp = P.Mp, D = P.D;
    const int r0 = chunk * PHI_RCH;
    const int nr = min(PHI_RCH, N - r0);
    const int model = P.model;
    const double b = P.bv[item];
    const double* stash = P.grad_logPhi + ((size_t)item * N + r0) * M;
    const double* Yg = P.Y + (size_t)P.item_gene[item] * N * P.Dtot;
    if (tid < PHI_RCH) act_s[tid] = (tid < nr && P.t[r0 + tid] > b) ? 1 : 0;
    if (tid == 0) {   // KL of the chunk: the rows' partials in a fixed order
        double kl = 0.0;
        for (int rr = 0; rr < nr; ++rr)
            for (int c = 0; c < ncl; ++c) kl += P.klrow[(size_t)slot * P.slot_stride + (size_t)(r0 + rr) * PHI_KLW + c];
        P.klpart[(size_t)slot * P.slot_stride + chunk] = kl;
    }
    for (int d0 = 0; d0 < D; d0 += PHI_DCH) {
        const int nd = min(PHI_DCH, D - d0);
        __syncthreads();
        if (tid < PHI_RCH * PHI_DCH) {
            const int rr = tid / PHI_DCH, dd = tid % PHI_DCH;
            Ys[rr][dd] = (rr < nr && dd < nd) ? Yg[(size_t)(r0 + rr) * P.Dtot + d0 + dd] : 0.0;
        }
        __syncthreads();
        for (int j = tid; j < Mp; j += NTHREADS) {
            double a = 0.0, vv[PHI_DCH];
#pragma unroll
            for (int dd = 0; dd < PHI_DCH; ++dd) vv[dd] = 0.0;
            if (j < M) {
                for (int rr = 0; rr < nr; ++rr) {
                    const int f = j - 3 * (r0 + rr);
                    const double phi = phi_of(model, act_s[rr] != 0, f >= 0 && f < 3, f, stash[(size_t)rr * M + j]);
                    a += phi;
#pragma unroll
                    for (int dd = 0; dd < PHI_DCH; ++dd) vv[dd] += phi * Ys[rr][dd];
                }
            }
            if (d0 == 0) P.Apart[(size_t)slot * P.slot_stride + (size_t)chunk * Mp + j] = a;
            for (int dd = 0; dd < nd; ++dd) P.vpart[(size_t)slot * P.slot_stride + ((size_t)chunk * D + d0 + dd) * Mp + j] = vv[dd];
        }
    }
}

// grid (row chunks, ITEMS of the wave): sums the contributions of the item's SD sub-problems.
__global__ void __launch_bounds__(NTHREADS) k_phi_backward(PhiParams P) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int chunk = blockIdx.x, litem = blockIdx.y, item = P.item0 + litem;
    const int N = P.N, M = P.M, Mp = P.Mp, D = P.D, SD = P.SD;
    const int r0 = chunk * PHI_RCH;
    const int nr = min(PHI_RCH, N - r0);
    const double log_eps = log(1e-6);
    const double* Yg = P.Y + (size_t)P.item_gene[item] * N * P.Dtot;
    const double* bvi = P.bv + (size_t)item * SD;
    const size_t slot0 = (size_t)litem * SD;
    for (int rr = warp; rr < nr; rr += NTHREADS / 32) {
        const int row = r0 + rr;
        const double* x = P.logPhi + ((size_t)item * N + row) * M;
        double* gout = P.grad_logPhi + ((size_t)item * N + row) * M;
        const double ti = P.t[row];
        // sub-problems in which this row is a free (softmax) row; BGP rows are always free, MBGP trunk rows
        // (t_i <= b_d) use the constant phiTrunk and carry no gradient
        int nact = 0;
        for (int sd = 0; sd < SD; ++sd) nact += (P.model == MODEL_BGP || ti > bvi[sd]) ? 1 : 0;
        if (nact == 0) {
            for (int j = lane; j < M; j += 32) gout[j] = 0.0;
            continue;
        }
        const bool act0 = ti > bvi[0];   // BGP: selects the trunk / branch form of pZ;  MBGP active rows: always the eZ0 form
        const double l = P.lse[slot0 * P.slot_stride + row];
        double r = 0.0;
        for (int j = lane; j < M; j += 32) {
            const int f = j - 3 * row;
            const bool inblock = (f >= 0 && f < 3);
            const double S = gout[j];   // softmax value stashed by the forward pass
            const double phi = SQ_A * S + SQ_B;
            double acc = 0.0;
            for (int sd = 0; sd < SD; ++sd) {
                if (!(P.model == MODEL_BGP || ti > bvi[sd])) continue;
                const double* Abar = P.Abar + (slot0 + sd) * P.slot_stride;
                const double* vbar = P.vbar + (slot0 + sd) * P.slot_stride;
                const int yc = (SD > 1) ? sd : 0;
                double yv = 0.0;
                for (int d = 0; d < D; ++d) yv += Yg[(size_t)row * P.Dtot + yc + d] * vbar[(size_t)d * Mp + j];
                acc += Abar[j] + yv;
            }
            const double lpz = log_pz(P.model, P.model == MODEL_BGP ? act0 : true, inblock, f, P.prior, row, log_eps);
            const double sb = SQ_A * (acc - P.kl_weight * (double)nact * (fast_log(phi) + 1.0 - lpz));
            gout[j] = sb;
            r += S * sb;
        }
        r = warp_sum(r);
        for (int j = lane; j < M; j += 32) {
            const double S = fast_exp(x[j] - l);
            gout[j] = S * (gout[j] - r);
        }
    }
}

// Row-resident backward pass: the softmax values S (stashed by the forward pass in the gradient buffer) and Sbar of one row of
// logPhi live in registers (EPT elements per thread), so the stash is read once and the gradient written once over it, with one
// log per element.  A row is owned by one CTA, or -- rows longer than 6144 columns (N > 2048, e.g. the N = 10 000 sparse
// configuration) -- by a thread-block cluster of CL CTAs that exchange their partial sums of S . Sbar through distributed
// shared memory.  grid (N * CL, items).
// SIMPLE: one sub-problem per item and one output (every BGP FitModel call): the loops over sub-problems / outputs and their
// pointer arithmetic disappear from the per-element code (ncu: 248 -> ~120 executed instructions per element).
template <int BT, int EPT, int MINB, int CL, bool SIMPLE>
__global__ void __launch_bounds__(BT, MINB) k_phi_backward_rows(PhiParams P) {
    __shared__ double red[32];
    __shared__ double part;
    const int tid = threadIdx.x;
    const int rank = CL > 1 ? (int)(blockIdx.x % CL) : 0;
    const int row = CL > 1 ? (int)(blockIdx.x / CL) : (int)blockIdx.x;
    const int litem = blockIdx.y, item = P.item0 + litem;
    const int N = P.N, M = P.M, Mp = P.Mp, D = P.D, SD = P.SD;
    const int j0 = rank * BT * EPT;
    const double log_eps = log(1e-6);
    const double* Yg = P.Y + (size_t)P.item_gene[item] * N * P.Dtot;
    const double* bvi = P.bv + (size_t)item * SD;
    const size_t slot0 = (size_t)litem * SD;
    double* gout = P.grad_logPhi + ((size_t)item * N + row) * M;
    const double ti = P.t[row];
    int nact = 0;
    for (int sd = 0; sd < SD; ++sd) nact += (P.model == MODEL_BGP || ti > bvi[sd]) ? 1 : 0;
    if (nact == 0) {   // uniform over the CTAs of the row
        for (int j = j0 + tid; j < M && j < j0 + BT * EPT; j += BT) gout[j] = 0.0;
        return;
    }
    const bool act0 = ti > bvi[0];
    // loop invariants in registers (the parameter block lives in constant memory)
    const int model = P.model;
    const double* prior = P.prior;
    const double klw = P.kl_weight * (double)nact;
    const double* Abar0 = P.Abar + slot0 * P.slot_stride;
    const double* vbar0 = P.vbar + slot0 * P.slot_stride;
    const double y0 = Yg[(size_t)row * P.Dtot];
    double S[EPT], sb[EPT];
    double r = 0.0;
#pragma unroll
    for (int e = 0; e < EPT; ++e) {
        const int j = j0 + tid + e * BT;
        S[e] = 0.0; sb[e] = 0.0;
        if (j < M) {
            const int f = j - 3 * row;
            const bool inblock = (f >= 0 && f < 3);
            const double s_ = gout[j];   // softmax value stashed by the forward pass
            const double phi = SQ_A * s_ + SQ_B;
            double acc = 0.0;
            if (SIMPLE) {
                acc = Abar0[j] + y0 * vbar0[j];
            } else {
                for (int sd = 0; sd < SD; ++sd) {
                    if (!(model == MODEL_BGP || ti > bvi[sd])) continue;
                    const double* Abar = P.Abar + (slot0 + sd) * P.slot_stride;
                    const double* vbar = P.vbar + (slot0 + sd) * P.slot_stride;
                    const int yc = (SD > 1) ? sd : 0;
                    double yv = 0.0;
                    for (int d = 0; d < D; ++d) yv += Yg[(size_t)row * P.Dtot + yc + d] * vbar[(size_t)d * Mp + j];
                    acc += Abar[j] + yv;
                }
            }
            const double lpz = log_pz(model, model == MODEL_BGP ? act0 : true, inblock, f, prior, row, log_eps);
            const double v = SQ_A * (acc - klw * (fast_log(phi) + 1.0 - lpz));
            S[e] = s_; sb[e] = v;
            r += s_ * v;
        }
    }
    r = block_sum(r, red);
    if (CL > 1) {   // sum the partials of the row's CTAs in rank order (deterministic) through distributed shared memory
        namespace cg = cooperative_groups;
        cg::cluster_group cluster = cg::this_cluster();
        if (tid == 0) part = r;
        cluster.sync();
        double rt = 0.0;
        for (int c = 0; c < CL; ++c) rt += *cluster.map_shared_rank(&part, c);
        cluster.sync();   // nobody leaves while its partial may still be read
        r = rt;
    }
#pragma unroll
    for (int e = 0; e < EPT; ++e) {
        const int j = j0 + tid + e * BT;
        if (j < M) gout[j] = S[e] * (sb[e] - r);
    }
}


// ------------------------------------------------------------------------------------------------------ streaming passes
// Rows too long for the registers of a CTA (M > 6144: the N = 10 000 inducing-point configuration has 30 000 columns) in bound +
// gradient calls with one sub-problem and one output per item (every BGP FitModel call).  Nothing is row resident; every pass is
// a plain stream over the N x M matrix with all of the GPU's warps busy, and the row-wide quantities travel through small
// per-(column chunk, row) partial arrays that are summed in a fixed order (deterministic):
//   k_phi_lse_stream  CTA per row, read logPhi once: online (max, sum exp) -> row log-sum-exp                       8 B / element
//   k_phi_fwd_stream  CTA = 1024 columns x rch rows, thread = 4 columns: S = exp(x - lse) -> stash (gradient buffer), Phi, KL,
//                     A = sum_i Phi, v = Phi^T y per column, c_i = sum_j S (log Phi + 1 - log pZ) per row            8 B + 8 B
//   k_phi_kl_stream   KL partials of the column chunks summed per row chunk
//   k_phi_rsum_stream r_i = sum_j S Sbar = a (sum_j S (Abar_j + y_i vbar_j) - w c_i): reads the stash, no transcendental     8 B
//   k_phi_bwd_stream  dlogPhi = S (Sbar - r_i) written over the stash, one log per element                           8 B + 8 B
// 40 B of DRAM traffic, two exps and two logs per element against 2 x 16 B + partial re-reads, one exp and two logs for the
// row-resident kernels -- which at these row lengths run at a quarter of the issue rate (cluster-wide reductions between the
// load, compute and store phases of a CTA, two CTAs per SM).
constexpr int SW = NTHREADS / 32;            // warps per CTA
constexpr int SCPT = PHI_SCOLS / NTHREADS;   // columns per thread

__device__ __forceinline__ void lse_merge(double& m, double& s, double m2, double s2) {
    const double mx = fmax(m, m2);
    s = s * exp(m - mx) + s2 * exp(m2 - mx);
    m = mx;
}

__global__ void __launch_bounds__(NTHREADS) k_phi_lse_stream(PhiParams P) {
    __shared__ double ms[SW], ss[SW];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int row = blockIdx.x, slot = blockIdx.y, item = P.item0 + slot;
    const int M = P.M;
    const double* x = P.logPhi + ((size_t)item * P.N + row) * M;
    constexpr int U = 4;
    double m = -1.0e308, s = 0.0;
    for (int j0 = tid; j0 < M; j0 += NTHREADS * U) {
        double v[U];
        double lm = -1.0e308;
#pragma unroll
        for (int k = 0; k < U; ++k) {
            const int j = j0 + k * NTHREADS;
            v[k] = (j < M) ? x[j] : -1.0e308;
            lm = fmax(lm, v[k]);
        }
        if (lm > m) { s *= fast_exp(m - lm); m = lm; }   // rescale the running sum (rare after the first blocks)
#pragma unroll
        for (int k = 0; k < U; ++k) s += fast_exp(v[k] - m);   // padding (-1e308) contributes exactly 0
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double m2 = __shfl_xor_sync(0xffffffffu, m, o), s2 = __shfl_xor_sync(0xffffffffu, s, o);
        lse_merge(m, s, m2, s2);
    }
    if (lane == 0) { ms[warp] = m; ss[warp] = s; }
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < SW; ++w) lse_merge(m, s, ms[w], ss[w]);
        P.lse[(size_t)slot * P.slot_stride + row] = m + log(s);
    }
}

__global__ void __launch_bounds__(NTHREADS, 3) k_phi_fwd_stream(PhiParams P) {
    __shared__ double lse_s[PHI_SRCH], y_s[PHI_SRCH], cw[SW][PHI_SRCH], red[32];
    __shared__ int act_s[PHI_SRCH];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cc = blockIdx.x, chunk = blockIdx.y, slot = blockIdx.z, item = P.item0 + slot;
    const int N = P.N, M = P.M, Mp = P.Mp, rch = P.rch, ncc = gridDim.x;
    const int r0 = chunk * rch, nr = min(rch, N - r0);
    const int jb = cc * PHI_SCOLS + tid;
    const int model = P.model;
    const double* prior = P.prior;
    const double log_eps = log(1e-6);
    const double b = P.bv[item];
    const double* lp = P.logPhi + ((size_t)item * N + r0) * M;
    double* stash = P.grad_logPhi + ((size_t)item * N + r0) * M;
    const double* Yg = P.Y + (size_t)P.item_gene[item] * N * P.Dtot;
    if (tid < nr) {
        lse_s[tid] = P.lse[(size_t)slot * P.slot_stride + r0 + tid];
        y_s[tid] = Yg[(size_t)(r0 + tid) * P.Dtot];
        act_s[tid] = P.t[r0 + tid] > b ? 1 : 0;
    }
    __syncthreads();
    double a[SCPT], v[SCPT], xn[SCPT];
    double kl = 0.0;
#pragma unroll
    for (int k = 0; k < SCPT; ++k) {
        a[k] = 0.0; v[k] = 0.0;
        const int j = jb + k * NTHREADS;
        xn[k] = (j < M) ? lp[j] : 0.0;
    }
    for (int rr = 0; rr < nr; ++rr) {
        double x[SCPT];
#pragma unroll
        for (int k = 0; k < SCPT; ++k) x[k] = xn[k];
        if (rr + 1 < nr) {   // next row's elements are in flight while this row's are worked on
#pragma unroll
            for (int k = 0; k < SCPT; ++k) {
                const int j = jb + k * NTHREADS;
                if (j < M) xn[k] = lp[(size_t)(rr + 1) * M + j];
            }
        }
        const int row = r0 + rr;
        const double l = lse_s[rr], y = y_s[rr];
        const bool act = act_s[rr] != 0;
        double c = 0.0;
#pragma unroll
        for (int k = 0; k < SCPT; ++k) {
            const int j = jb + k * NTHREADS;
            if (j < M) {
                const double S = fast_exp(x[k] - l);
                stash[(size_t)rr * M + j] = S;
                const int f = j - 3 * row;
                const bool inblock = (f >= 0 && f < 3);
                const double phi = phi_of(model, act, inblock, f, S);
                const double d = fast_log(phi) - log_pz(model, act, inblock, f, prior, row, log_eps);
                a[k] += phi;
                v[k] = fma(phi, y, v[k]);
                kl = fma(phi, d, kl);
                c = fma(S, d + 1.0, c);
            }
        }
        c = warp_sum(c);
        if (lane == 0) cw[warp][rr] = c;
    }
#pragma unroll
    for (int k = 0; k < SCPT; ++k) {
        const int j = jb + k * NTHREADS;
        if (j < Mp) {
            P.Apart[(size_t)slot * P.slot_stride + (size_t)chunk * Mp + j] = a[k];
            P.vpart[(size_t)slot * P.slot_stride + (size_t)chunk * Mp + j] = v[k];
        }
    }
    kl = block_sum(kl, red);   // (contains the __syncthreads that publishes cw)
    if (tid == 0) P.klrow[(size_t)slot * P.slot_stride + (size_t)chunk * ncc + cc] = kl;
    if (tid < nr) {
        double c = 0.0;
        for (int w = 0; w < SW; ++w) c += cw[w][tid];
        P.rowpart[(size_t)slot * P.slot_stride + (size_t)cc * N + r0 + tid] = c;
    }
}

__global__ void __launch_bounds__(NTHREADS) k_phi_kl_stream(PhiParams P, int ncc) {
    const int slot = blockIdx.x;
    for (int ch = threadIdx.x; ch < P.nch; ch += NTHREADS) {
        double kl = 0.0;
        for (int c = 0; c < ncc; ++c) kl += P.klrow[(size_t)slot * P.slot_stride + (size_t)ch * ncc + c];
        P.klpart[(size_t)slot * P.slot_stride + ch] = kl;
    }
}

__global__ void __launch_bounds__(NTHREADS, 3) k_phi_rsum_stream(PhiParams P) {
    __shared__ double y_s[PHI_SRCH], cw[SW][PHI_SRCH];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cc = blockIdx.x, chunk = blockIdx.y, slot = blockIdx.z, item = P.item0 + slot;
    const int N = P.N, M = P.M, rch = P.rch;
    const int r0 = chunk * rch, nr = min(rch, N - r0);
    const int jb = cc * PHI_SCOLS + tid;
    const double* S = P.grad_logPhi + ((size_t)item * N + r0) * M;
    const double* Yg = P.Y + (size_t)P.item_gene[item] * N * P.Dtot;
    const double* Abar = P.Abar + (size_t)slot * P.slot_stride;
    const double* vbar = P.vbar + (size_t)slot * P.slot_stride;
    if (tid < nr) y_s[tid] = Yg[(size_t)(r0 + tid) * P.Dtot];
    double ab[SCPT], vb[SCPT];
#pragma unroll
    for (int k = 0; k < SCPT; ++k) {
        const int j = jb + k * NTHREADS;
        ab[k] = (j < M) ? Abar[j] : 0.0;
        vb[k] = (j < M) ? vbar[j] : 0.0;
    }
    __syncthreads();
    constexpr int RU = 4;   // rows in flight per thread
    for (int rb = 0; rb < nr; rb += RU) {
        double s[RU][SCPT];
#pragma unroll
        for (int u = 0; u < RU; ++u)
#pragma unroll
            for (int k = 0; k < SCPT; ++k) {
                const int j = jb + k * NTHREADS;
                s[u][k] = (rb + u < nr && j < M) ? S[(size_t)(rb + u) * M + j] : 0.0;
            }
#pragma unroll
        for (int u = 0; u < RU; ++u) {
            if (rb + u < nr) {
                const double y = y_s[rb + u];
                double p = 0.0;
#pragma unroll
                for (int k = 0; k < SCPT; ++k) p = fma(s[u][k], fma(y, vb[k], ab[k]), p);
                p = warp_sum(p);
                if (lane == 0) cw[warp][rb + u] = p;
            }
        }
    }
    __syncthreads();
    if (tid < nr) {
        double p = 0.0;
        for (int w = 0; w < SW; ++w) p += cw[w][tid];
        double* rp = P.rowpart + (size_t)slot * P.slot_stride + (size_t)cc * N + r0 + tid;
        *rp = SQ_A * (p - P.kl_weight * *rp);
    }
}

__global__ void __launch_bounds__(NTHREADS, 3) k_phi_bwd_stream(PhiParams P) {
    __shared__ double r_s[PHI_SRCH], y_s[PHI_SRCH];
    __shared__ int act_s[PHI_SRCH];
    const int tid = threadIdx.x;
    const int cc = blockIdx.x, chunk = blockIdx.y, slot = blockIdx.z, item = P.item0 + slot;
    const int N = P.N, M = P.M, rch = P.rch, ncc = gridDim.x;
    const int r0 = chunk * rch, nr = min(rch, N - r0);
    const int jb = cc * PHI_SCOLS + tid;
    const int model = P.model;
    const double* prior = P.prior;
    const double log_eps = log(1e-6);
    const double klw = P.kl_weight;
    const double b = P.bv[item];
    double* g = P.grad_logPhi + ((size_t)item * N + r0) * M;
    const double* Yg = P.Y + (size_t)P.item_gene[item] * N * P.Dtot;
    const double* Abar = P.Abar + (size_t)slot * P.slot_stride;
    const double* vbar = P.vbar + (size_t)slot * P.slot_stride;
    if (tid < nr) {
        double r = 0.0;
        for (int c = 0; c < ncc; ++c) r += P.rowpart[(size_t)slot * P.slot_stride + (size_t)c * N + r0 + tid];
        r_s[tid] = r;
        y_s[tid] = Yg[(size_t)(r0 + tid) * P.Dtot];
        act_s[tid] = P.t[r0 + tid] > b ? 1 : 0;
    }
    double ab[SCPT], vb[SCPT], sn[SCPT];
#pragma unroll
    for (int k = 0; k < SCPT; ++k) {
        const int j = jb + k * NTHREADS;
        ab[k] = (j < M) ? Abar[j] : 0.0;
        vb[k] = (j < M) ? vbar[j] : 0.0;
        sn[k] = (j < M) ? g[j] : 0.0;
    }
    __syncthreads();
    for (int rr = 0; rr < nr; ++rr) {
        double s[SCPT];
#pragma unroll
        for (int k = 0; k < SCPT; ++k) s[k] = sn[k];
        if (rr + 1 < nr) {
#pragma unroll
            for (int k = 0; k < SCPT; ++k) {
                const int j = jb + k * NTHREADS;
                if (j < M) sn[k] = g[(size_t)(rr + 1) * M + j];
            }
        }
        const int row = r0 + rr;
        const double r = r_s[rr], y = y_s[rr];
        const bool act = act_s[rr] != 0;
        if (model != MODEL_BGP && !act) {   // MBGP trunk rows use the constant phiTrunk and carry no gradient
#pragma unroll
            for (int k = 0; k < SCPT; ++k) {
                const int j = jb + k * NTHREADS;
                if (j < M) g[(size_t)rr * M + j] = 0.0;
            }
            continue;
        }
#pragma unroll
        for (int k = 0; k < SCPT; ++k) {
            const int j = jb + k * NTHREADS;
            if (j < M) {
                const int f = j - 3 * row;
                const bool inblock = (f >= 0 && f < 3);
                const double phi = fma(SQ_A, s[k], SQ_B);
                const double lpz = log_pz(model, model == MODEL_BGP ? act : true, inblock, f, prior, row, log_eps);
                const double sb = SQ_A * (fma(y, vb[k], ab[k]) - klw * (fast_log(phi) + 1.0 - lpz));
                g[(size_t)rr * M + j] = s[k] * (sb - r);
            }
        }
    }
}


// ---- BGP specialisations of the two heavy streams: branch-free element code.  Phi = a S + b everywhere, log pZ = log 1e-6 for
// every element outside the 3 own-cell columns of a row; those (at most 3 per row, in one column chunk) are patched by the row's
// thread after the main loop, so the hot loop has no data-dependent branch and the four elements of a thread interleave.
template <bool FULL>
__device__ __forceinline__ void fwd_row_bgp(const double (&x)[SCPT], double* __restrict__ st, int jb, int M, double l, double y,
                                            double log_eps, double (&a)[SCPT], double (&v)[SCPT], double& kl, double& c) {
#pragma unroll
    for (int k = 0; k < SCPT; ++k) {
        const int j = jb + k * NTHREADS;
        if (FULL || j < M) {
            const double S = fast_exp(x[k] - l);
            st[k * NTHREADS] = S;
            const double phi = fma(SQ_A, S, SQ_B);
            const double d = fast_log(phi) - log_eps;
            a[k] += phi;
            v[k] = fma(phi, y, v[k]);
            kl = fma(phi, d, kl);
            c = fma(S, d + 1.0, c);
        }
    }
}

// TMA = true (rows start 16-byte aligned, i.e. M even): the CTA's 8 KB piece of each row arrives through cp.async.bulk in a ring of
// SNST stages filled SNST - 1 rows ahead by thread 0 (full / empty mbarriers, one empty arrival per warp), so 3 CTAs keep
// 72 KB per SM in flight without spending registers on it; TMA = false: register prefetch of the next row.
constexpr int SNST = 4;
struct StreamRing {
    double* ring;      // [SNST][PHI_SCOLS]
    uint64_t* full;    // [SNST]
    uint64_t* empty;   // [SNST]
    const double* src; // element (row 0, column c0) of this CTA's block
    uint32_t bytes;    // per row
    int M, nr;
    __device__ __forceinline__ void init(int tid) {
        if (tid == 0) {
            for (int s = 0; s < SNST; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], SW); }
            fence_mbar_init();
        }
        __syncthreads();
        if (tid == 0)
            for (int r = 0; r < SNST - 1 && r < nr; ++r) issue(r);
    }
    __device__ __forceinline__ void issue(int r) {
        const int s = r % SNST;
        mbar_arrive_expect_tx(&full[s], bytes);
        tma_load_1d(ring + (size_t)s * PHI_SCOLS, src + (size_t)r * M, bytes, &full[s]);
    }
    // top of iteration rr: refill the stage consumed in iteration rr - 1, then wait for row rr
    __device__ __forceinline__ const double* acquire(int rr, int tid) {
        if (tid == 0) {
            const int nxt = rr + SNST - 1;
            if (nxt < nr) {
                if (nxt >= SNST) mbar_wait(&empty[nxt % SNST], (uint32_t)((nxt / SNST - 1) & 1));
                issue(nxt);
            }
        }
        mbar_wait(&full[rr % SNST], (uint32_t)((rr / SNST) & 1));
        return ring + (size_t)(rr % SNST) * PHI_SCOLS;
    }
    __device__ __forceinline__ void release(int rr, int lane) {
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[rr % SNST]);
    }
};

#ifndef BGP_FWD_MINB
#define BGP_FWD_MINB 3
#endif
template <bool TMA>
__global__ void __launch_bounds__(NTHREADS, BGP_FWD_MINB) k_phi_fwd_stream_bgp(PhiParams P) {
    __shared__ double lse_s[PHI_SRCH], y_s[PHI_SRCH], cw[SW][PHI_SRCH], red[32];
    __shared__ int act_s[PHI_SRCH];
    __shared__ __align__(128) double ring_s[TMA ? SNST * PHI_SCOLS : 1];
    __shared__ uint64_t bars[2 * SNST];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cc = blockIdx.x, chunk = blockIdx.y, slot = blockIdx.z, item = P.item0 + slot;
    const int N = P.N, M = P.M, Mp = P.Mp, rch = P.rch, ncc = gridDim.x;
    const int r0 = chunk * rch, nr = min(rch, N - r0);
    const int c0 = cc * PHI_SCOLS, jb = c0 + tid;
    const bool full = c0 + PHI_SCOLS <= M;
    const double log_eps = log(1e-6);
    const double b = P.bv[item];
    const double* lp = P.logPhi + ((size_t)item * N + r0) * M + jb;
    double* st = P.grad_logPhi + ((size_t)item * N + r0) * M + jb;
    const double* Yg = P.Y + (size_t)P.item_gene[item] * N * P.Dtot;
    if (tid < nr) {
        lse_s[tid] = P.lse[(size_t)slot * P.slot_stride + r0 + tid];
        y_s[tid] = Yg[(size_t)(r0 + tid) * P.Dtot];
        act_s[tid] = P.t[r0 + tid] > b ? 1 : 0;
    }
    __syncthreads();
    double a[SCPT], v[SCPT], xn[SCPT];
    double kl = 0.0;
    StreamRing R;
    if (TMA) {
        R.ring = ring_s; R.full = bars; R.empty = bars + SNST; R.src = lp - tid; R.M = M; R.nr = nr;
        R.bytes = (uint32_t)(min(PHI_SCOLS, M - c0) * 8);
        R.init(tid);
    }
#pragma unroll
    for (int k = 0; k < SCPT; ++k) {
        a[k] = 0.0; v[k] = 0.0;
        xn[k] = 0.0;
        if (!TMA && (full || jb + k * NTHREADS < M)) xn[k] = lp[k * NTHREADS];
    }
    for (int rr = 0; rr < nr; ++rr) {
        double x[SCPT];
        if (TMA) {
            const double* rowp = R.acquire(rr, tid);
#pragma unroll
            for (int k = 0; k < SCPT; ++k) x[k] = (full || jb + k * NTHREADS < M) ? rowp[tid + k * NTHREADS] : 0.0;
            R.release(rr, lane);
        } else {
#pragma unroll
            for (int k = 0; k < SCPT; ++k) x[k] = xn[k];
            lp += M;
            if (rr + 1 < nr) {   // next row's elements are in flight while this row's are worked on
#pragma unroll
                for (int k = 0; k < SCPT; ++k)
                    if (full || jb + k * NTHREADS < M) xn[k] = lp[k * NTHREADS];
            }
        }
        double c = 0.0;
        if (full) fwd_row_bgp<true>(x, st, jb, M, lse_s[rr], y_s[rr], log_eps, a, v, kl, c);
        else fwd_row_bgp<false>(x, st, jb, M, lse_s[rr], y_s[rr], log_eps, a, v, kl, c);
        st += M;
        c = warp_sum(c);
        if (lane == 0) cw[warp][rr] = c;
    }
#pragma unroll
    for (int k = 0; k < SCPT; ++k) {
        const int j = jb + k * NTHREADS;
        if (j < Mp) {
            P.Apart[(size_t)slot * P.slot_stride + (size_t)chunk * Mp + j] = a[k];
            P.vpart[(size_t)slot * P.slot_stride + (size_t)chunk * Mp + j] = v[k];
        }
    }
    __syncthreads();   // the stash values written above are visible to the row threads below
    if (tid < nr) {
        const int row = r0 + tid;
        double c = 0.0;
        for (int w = 0; w < SW; ++w) c += cw[w][tid];
        const double* srow = P.grad_logPhi + ((size_t)item * N + row) * M;
        for (int f = 0; f < 3; ++f) {   // own-cell columns: replace log 1e-6 by the true log pZ
            const int j = 3 * row + f;
            if (j >= c0 && j < c0 + PHI_SCOLS) {
                const double S = srow[j];
                const double dl = log_eps - log_pz(MODEL_BGP, act_s[tid] != 0, true, f, P.prior, row, log_eps);
                kl = fma(fma(SQ_A, S, SQ_B), dl, kl);
                c = fma(S, dl, c);
            }
        }
        P.rowpart[(size_t)slot * P.slot_stride + (size_t)cc * N + row] = c;
    }
    kl = block_sum(kl, red);
    if (tid == 0) P.klrow[(size_t)slot * P.slot_stride + (size_t)chunk * ncc + cc] = kl;
}

template <bool FULL>
__device__ __forceinline__ void bwd_row_bgp(const double (&s)[SCPT], double* __restrict__ g, int jb, int M, double r, double y, double klw,
                                            double log_eps, const double (&ab)[SCPT], const double (&vb)[SCPT]) {
#pragma unroll
    for (int k = 0; k < SCPT; ++k) {
        if (FULL || jb + k * NTHREADS < M) {
            const double phi = fma(SQ_A, s[k], SQ_B);
            const double sb = SQ_A * (fma(y, vb[k], ab[k]) - klw * (fast_log(phi) + 1.0 - log_eps));
            g[k * NTHREADS] = s[k] * (sb - r);
        }
    }
}

template <bool TMA>
__global__ void __launch_bounds__(NTHREADS, 3) k_phi_bwd_stream_bgp(PhiParams P) {
    __shared__ double r_s[PHI_SRCH], y_s[PHI_SRCH];
    __shared__ __align__(128) double ring_s[TMA ? SNST * PHI_SCOLS : 1];
    __shared__ uint64_t bars[2 * SNST];
    const int tid = threadIdx.x, lane = tid & 31;
    const int cc = blockIdx.x, chunk = blockIdx.y, slot = blockIdx.z, item = P.item0 + slot;
    const int N = P.N, M = P.M, rch = P.rch, ncc = gridDim.x;
    const int r0 = chunk * rch, nr = min(rch, N - r0);
    const int c0 = cc * PHI_SCOLS, jb = c0 + tid;
    const bool full = c0 + PHI_SCOLS <= M;
    const double log_eps = log(1e-6);
    const double klw = P.kl_weight;
    double* g = P.grad_logPhi + ((size_t)item * N + r0) * M + jb;
    const double* Yg = P.Y + (size_t)P.item_gene[item] * N * P.Dtot;
    const double* Abar = P.Abar + (size_t)slot * P.slot_stride;
    const double* vbar = P.vbar + (size_t)slot * P.slot_stride;
    double sown[3] = {0.0, 0.0, 0.0};   // softmax values of the row's own-cell columns (row thread), saved before they are overwritten
    if (tid < nr) {
        double r = 0.0;
        for (int c = 0; c < ncc; ++c) r += P.rowpart[(size_t)slot * P.slot_stride + (size_t)c * N + r0 + tid];
        r_s[tid] = r;
        y_s[tid] = Yg[(size_t)(r0 + tid) * P.Dtot];
        const int row = r0 + tid;
        const double* srow = P.grad_logPhi + ((size_t)item * N + row) * M;
        for (int f = 0; f < 3; ++f) {
            const int j = 3 * row + f;
            if (j >= c0 && j < c0 + PHI_SCOLS) sown[f] = srow[j];
        }
    }
    double ab[SCPT], vb[SCPT], sn[SCPT];
#pragma unroll
    for (int k = 0; k < SCPT; ++k) {
        const bool ok = full || jb + k * NTHREADS < M;
        ab[k] = ok ? Abar[jb + k * NTHREADS] : 0.0;
        vb[k] = ok ? vbar[jb + k * NTHREADS] : 0.0;
        sn[k] = (!TMA && ok) ? g[k * NTHREADS] : 0.0;
    }
    StreamRing R;
    if (TMA) {   // (init holds the __syncthreads that orders the own-cell reads above before any write below)
        R.ring = ring_s; R.full = bars; R.empty = bars + SNST; R.src = g - tid; R.M = M; R.nr = nr;
        R.bytes = (uint32_t)(min(PHI_SCOLS, M - c0) * 8);
        R.init(tid);
    }
    __syncthreads();
    for (int rr = 0; rr < nr; ++rr) {
        double sv[SCPT];
        if (TMA) {
            const double* rowp = R.acquire(rr, tid);
#pragma unroll
            for (int k = 0; k < SCPT; ++k) sv[k] = (full || jb + k * NTHREADS < M) ? rowp[tid + k * NTHREADS] : 0.0;
            R.release(rr, lane);
        } else {
#pragma unroll
            for (int k = 0; k < SCPT; ++k) sv[k] = sn[k];
            if (rr + 1 < nr) {
#pragma unroll
                for (int k = 0; k < SCPT; ++k)
                    if (full || jb + k * NTHREADS < M) sn[k] = g[(size_t)M + k * NTHREADS];
            }
        }
        if (full) bwd_row_bgp<true>(sv, g, jb, M, r_s[rr], y_s[rr], klw, log_eps, ab, vb);
        else bwd_row_bgp<false>(sv, g, jb, M, r_s[rr], y_s[rr], klw, log_eps, ab, vb);
        g += M;
    }
    __syncthreads();   // the row threads overwrite elements other threads have just written
    if (tid < nr) {
        const int row = r0 + tid;
        const bool act = P.t[row] > P.bv[item];
        double* grow = P.grad_logPhi + ((size_t)item * N + row) * M;
        for (int f = 0; f < 3; ++f) {
            const int j = 3 * row + f;
            if (j >= c0 && j < c0 + PHI_SCOLS) {
                const double S = sown[f];
                const double lpz = log_pz(MODEL_BGP, act, true, f, P.prior, row, log_eps);
                const double sb = SQ_A * (fma(y_s[tid], vbar[j], Abar[j]) - klw * (fast_log(fma(SQ_A, S, SQ_B)) + 1.0 - lpz));
                grow[j] = S * (sb - r_s[tid]);
            }
        }
    }
}

static bool stream_enabled() {   // BGP_PHI_STREAM=0 keeps long rows on the row-resident / two-pass kernels (A/B timing, tests)
    const char* e = getenv("BGP_PHI_STREAM");
    return !(e && e[0] == '0');
}
bool phi_streaming(int N, int M, int SD, int D, bool want_grad, int rch) {
    if (!stream_enabled() || !want_grad || SD != 1 || D != 1 || M <= 512 * 12) return false;
    const long long nch = (N + rch - 1) / rch;
    return nch * phi_col_chunks(M) <= (long long)N * PHI_KLW;   // the KL partials live in klrow
}
// bulk copies need 16-byte aligned row pieces: even M and 16-byte aligned buffers (BGP_PHI_TMA=0: register prefetch, for A/B timing)
static bool stream_tma(const PhiParams& P) {
    const char* e = getenv("BGP_PHI_TMA");
    if (e && e[0] == '0') return false;
    return (P.M % 2 == 0) && ((uintptr_t)P.logPhi % 16 == 0) && ((uintptr_t)P.grad_logPhi % 16 == 0);
}
static bool params_streaming(const PhiParams& P) {
    const int rch = P.rch > 0 ? P.rch : PHI_RCH;
    return rch <= PHI_SRCH && phi_streaming(P.N, P.M, P.SD, P.D, P.grad_logPhi != nullptr, rch) && P.rowpart != nullptr;
}


// MBGP multi-output backward pass, MB_R rows per CTA: the per-element sum over the outputs that treat the row as a branch row,
//     acc_ij = sum_{d: t_i > b_d} (Abar_d[j] + Y_id vbar_d[j]),
// reads two vectors per output and column; with one row per CTA that is 2 SD M doubles from L2 per row (ncu: the pass was L2
// bound).  Here every pair (Abar_d[j], vbar_d[j]) fetched by a thread feeds MB_R rows.  Rows fit the CTA: M <= MB_BT * MB_EPT.
constexpr int MB_R = 4, MB_BT = 256, MB_EPT = 6;
__global__ void __launch_bounds__(MB_BT) k_phi_backward_rows_multi(PhiParams P) {
    __shared__ double y_s[MB_R][MO_MAXSD], bv_s[MO_MAXSD], t_s[MB_R], red[32];
    __shared__ int nact_s[MB_R];
    const int tid = threadIdx.x;
    const int litem = blockIdx.y, item = P.item0 + litem;
    const int N = P.N, M = P.M, SD = P.SD;
    const int r0 = blockIdx.x * MB_R, nr = min(MB_R, N - r0);
    const size_t slot0 = (size_t)litem * SD;
    const double* Yg = P.Y + (size_t)P.item_gene[item] * N * P.Dtot;
    const double* prior = P.prior;
    double* gout = P.grad_logPhi + ((size_t)item * N + r0) * M;
    for (int e = tid; e < MB_R * SD; e += MB_BT) {
        const int r = e / SD, sd = e - r * SD;
        y_s[r][sd] = r < nr ? Yg[(size_t)(r0 + r) * P.Dtot + sd] : 0.0;
    }
    for (int sd = tid; sd < SD; sd += MB_BT) bv_s[sd] = P.bv[(size_t)item * SD + sd];
    if (tid < MB_R) t_s[tid] = tid < nr ? P.t[r0 + tid] : -1.0e308;
    __syncthreads();
    if (tid < MB_R) {
        int na = 0;
        for (int sd = 0; sd < SD; ++sd) na += t_s[tid] > bv_s[sd] ? 1 : 0;
        nact_s[tid] = na;
    }
    double acc[MB_R][MB_EPT], S[MB_R][MB_EPT];
#pragma unroll
    for (int r = 0; r < MB_R; ++r)
#pragma unroll
        for (int e = 0; e < MB_EPT; ++e) acc[r][e] = 0.0;
    for (int sd = 0; sd < SD; ++sd) {
        const double b = bv_s[sd];
        const double* Abar = P.Abar + (slot0 + sd) * P.slot_stride;
        const double* vbar = P.vbar + (slot0 + sd) * P.slot_stride;
        double A[MB_EPT], V[MB_EPT];
#pragma unroll
        for (int e = 0; e < MB_EPT; ++e) {
            const int j = tid + e * MB_BT;
            A[e] = j < M ? Abar[j] : 0.0;
            V[e] = j < M ? vbar[j] : 0.0;
        }
#pragma unroll
        for (int r = 0; r < MB_R; ++r) {
            if (t_s[r] > b) {   // uniform over the CTA
                const double y = y_s[r][sd];
#pragma unroll
                for (int e = 0; e < MB_EPT; ++e) acc[r][e] += fma(y, V[e], A[e]);
            }
        }
    }
    __syncthreads();   // nact_s
    const double lpz_off = log(SQ_A * 1e-6 + SQ_B);
    double rs[MB_R];
#pragma unroll
    for (int r = 0; r < MB_R; ++r) {
        rs[r] = 0.0;
        const int row = r0 + r;
        const double klw = P.kl_weight * (double)nact_s[r];
#pragma unroll
        for (int e = 0; e < MB_EPT; ++e) {
            const int j = tid + e * MB_BT;
            S[r][e] = 0.0;
            if (r < nr && j < M) {
                const int f = j - 3 * row;
                const bool inblock = (f >= 0 && f < 3);
                const double s_ = gout[(size_t)r * M + j];   // softmax value stashed by the forward pass
                const double lpz = inblock ? log_pz(MODEL_MBGP, true, true, f, prior, row, 0.0) : lpz_off;
                const double v = SQ_A * (acc[r][e] - klw * (fast_log(fma(SQ_A, s_, SQ_B)) + 1.0 - lpz));
                S[r][e] = s_;
                acc[r][e] = v;
                rs[r] = fma(s_, v, rs[r]);
            }
        }
    }
#pragma unroll
    for (int r = 0; r < MB_R; ++r) rs[r] = block_sum(rs[r], red);
#pragma unroll
    for (int r = 0; r < MB_R; ++r) {
        if (r < nr) {
            const bool live = nact_s[r] > 0;   // trunk rows of every output carry no gradient
#pragma unroll
            for (int e = 0; e < MB_EPT; ++e) {
                const int j = tid + e * MB_BT;
                if (j < M) gout[(size_t)r * M + j] = live ? S[r][e] * (acc[r][e] - rs[r]) : 0.0;
            }
        }
    }
}

template <int BT, int EPT, int MINB, int CL>
static void launch_rows_cluster(const PhiParams& P, int nitems, cudaStream_t st) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(P.N * CL, nitems);
    cfg.blockDim = dim3(BT);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    if (P.SD == 1 && P.D == 1) cudaLaunchKernelEx(&cfg, k_phi_backward_rows<BT, EPT, MINB, CL, true>, P);
    else cudaLaunchKernelEx(&cfg, k_phi_backward_rows<BT, EPT, MINB, CL, false>, P);
}

void phi_launch_forward(const PhiParams& P, int nsubs, cudaStream_t st) {
    dim3 grid(P.nch, nsubs);
    if (params_streaming(P)) {
        const int ncc = phi_col_chunks(P.M);
        k_phi_lse_stream<<<dim3(P.N, nsubs), NTHREADS, 0, st>>>(P);
        if (P.model == MODEL_BGP && stream_tma(P)) k_phi_fwd_stream_bgp<true><<<dim3(ncc, P.nch, nsubs), NTHREADS, 0, st>>>(P);
        else if (P.model == MODEL_BGP) k_phi_fwd_stream_bgp<false><<<dim3(ncc, P.nch, nsubs), NTHREADS, 0, st>>>(P);
        else k_phi_fwd_stream<<<dim3(ncc, P.nch, nsubs), NTHREADS, 0, st>>>(P);
        k_phi_kl_stream<<<nsubs, NTHREADS, 0, st>>>(P, ncc);
        return;
    }
    // Row pass + column pass through the stash for rows that fit one CTA's registers.  The cluster variants of the row pass
    // (k_phi_rowpass<512, 8, 2, 4 or 8> under a cluster launch) work but were slower than k_phi_forward at N = 10 000 (39 ms against 30 ms for 8 items:
    // two block reductions and a cluster exchange per 4096 elements), so long rows stay with k_phi_forward.
    if (P.grad_logPhi != nullptr && P.SD == 1 && P.M <= 512 * 12) {
        if (P.M <= 256 * 4) k_phi_rowpass<256, 4, 4, 1><<<dim3(P.N, nsubs), 256, 0, st>>>(P);
        else if (P.M <= 512 * 6) k_phi_rowpass<512, 6, 2, 1><<<dim3(P.N, nsubs), 512, 0, st>>>(P);
        else k_phi_rowpass<512, 12, 1, 1><<<dim3(P.N, nsubs), 512, 0, st>>>(P);
        if (P.D == 1) k_phi_colpass<1><<<grid, NTHREADS, 0, st>>>(P, 1);
        else k_phi_colpass<4><<<grid, NTHREADS, 0, st>>>(P, 1);
        return;
    }
    const char* mo = getenv("BGP_PHI_MULTI");
    if (P.SD > 1 && P.SD <= MO_MAXSD && P.D == 1 && P.model == MODEL_MBGP && nsubs % P.SD == 0 && !(mo && mo[0] == '0')) {
        k_phi_forward_multi<<<dim3(P.nch, nsubs / P.SD), NTHREADS, 0, st>>>(P);   // shared softmax for the outputs of a joint model
        return;
    }
    if (P.D == 1) k_phi_forward<1><<<grid, NTHREADS, 0, st>>>(P);
    else k_phi_forward<4><<<grid, NTHREADS, 0, st>>>(P);
}
void phi_launch_backward(const PhiParams& P, int nitems, cudaStream_t st) {
    dim3 rows(P.N, nitems);
    if (params_streaming(P)) {
        const dim3 grid(phi_col_chunks(P.M), P.nch, nitems);
        k_phi_rsum_stream<<<grid, NTHREADS, 0, st>>>(P);
        if (P.model == MODEL_BGP && stream_tma(P)) k_phi_bwd_stream_bgp<true><<<grid, NTHREADS, 0, st>>>(P);
        else if (P.model == MODEL_BGP) k_phi_bwd_stream_bgp<false><<<grid, NTHREADS, 0, st>>>(P);
        else k_phi_bwd_stream<<<grid, NTHREADS, 0, st>>>(P);
        return;
    }
    const bool simple = (P.SD == 1 && P.D == 1);
    {
        const char* mo = getenv("BGP_PHI_MULTI");
        if (P.SD > 1 && P.SD <= MO_MAXSD && P.D == 1 && P.model == MODEL_MBGP && P.M <= MB_BT * MB_EPT && !(mo && mo[0] == '0')) {
            k_phi_backward_rows_multi<<<dim3((P.N + MB_R - 1) / MB_R, nitems), MB_BT, 0, st>>>(P);
            return;
        }
    }
    if (P.M <= 256 * 4) {
        if (simple) k_phi_backward_rows<256, 4, 4, 1, true><<<rows, 256, 0, st>>>(P);
        else k_phi_backward_rows<256, 4, 4, 1, false><<<rows, 256, 0, st>>>(P);
    } else if (P.M <= 512 * 6) {
        if (simple) k_phi_backward_rows<512, 6, 2, 1, true><<<rows, 512, 0, st>>>(P);
        else k_phi_backward_rows<512, 6, 2, 1, false><<<rows, 512, 0, st>>>(P);
    } else if (P.M <= 512 * 12) {
        if (simple) k_phi_backward_rows<512, 12, 1, 1, true><<<rows, 512, 0, st>>>(P);
        else k_phi_backward_rows<512, 12, 1, 1, false><<<rows, 512, 0, st>>>(P);
    }
    // clusters of 4 / 8 CTAs per row; 64 registers so that two CTAs share an SM (measured at N = 10 000, 8 items: 34 ms against
    // 73 ms with one 124-register CTA per SM, 46 ms with 1024 x 4, 50 ms for the two-pass kernel below)
    else if (P.M <= 4 * 512 * 8) launch_rows_cluster<512, 8, 2, 4>(P, nitems, st);
    else if (P.M <= 8 * 512 * 8) launch_rows_cluster<512, 8, 2, 8>(P, nitems, st);
    else {   // rows too long even for a cluster's registers: warp-per-row two-pass kernel
        dim3 grid(P.nch, nitems);
        k_phi_backward<<<grid, NTHREADS, 0, st>>>(P);
    }
}

}  // namespace bgp

namespace bgp {
// AssignGP.GetPhi (assigngp_dense.py:130-140): unsquashed softmax, gathered at columns 3i, 3i+1, 3i+2 of row i.
__global__ void __launch_bounds__(NTHREADS) k_phi_gather(const double* logPhi, double* phi, int N) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int item = blockIdx.y;
    const int M = 3 * N;
    const int row = blockIdx.x * (NTHREADS / 32) + warp;
    if (row >= N) return;
    const double* x = logPhi + ((size_t)item * N + row) * M;
    double m = -1.0e308;
    for (int j = lane; j < M; j += 32) m = fmax(m, x[j]);
    m = warp_max(m);
    double s = 0.0;
    for (int j = lane; j < M; j += 32) s += exp(x[j] - m);
    s = warp_sum(s);
    if (lane < 3) phi[((size_t)item * N + row) * 3 + lane] = exp(x[3 * row + lane] - m) / s;
}
void phi_launch_gather(const double* logPhi, double* phi, int count, int N, cudaStream_t st) {
    dim3 grid((N + NTHREADS / 32 - 1) / (NTHREADS / 32), count);
    k_phi_gather<<<grid, NTHREADS, 0, st>>>(logPhi, phi, N);
}
}  // namespace bgp
