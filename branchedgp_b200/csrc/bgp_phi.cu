// Assignment passes over the N x 3N variational matrix logPhi (HBM-bound, fp64 transcendentals).
//
// forward  (assigngp_dense.py:160-162, 242-245; MBGP/assigngp.py:217-251, 688-691):
//     S = softmax_rows(logPhi);  Phi = (1-2e-6) S + 1e-6;  A_j = sum_i Phi_ij;  v_jd = sum_i Phi_ij Y_id;
//     KL = sum Phi log Phi - sum Phi log pZ,  with pZ never materialised: it is a function of
//     (i, j, t_i <= b, prior[i]) -- pZ_construction_singleBP.py:6-25.
// backward (SURVEY.md App. B.1 last line):
//     Sbar = a (Abar_j + sum_d Y_id vbar_jd - w (log Phi + 1 - log pZ));  dlogPhi = S (Sbar - sum_k S_ik Sbar_ik)
#include "bgp_dense.cuh"
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

constexpr int PHI_DCH = 4;

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
    // ---- row log-sum-exp, one warp per row (max pass, then exp-sum pass; a branchy one-pass online variant measured slower)
    for (int rr = warp; rr < nr; rr += NTHREADS / 32) {
        const double* row = lp + (size_t)rr * M;
        double m = -1.0e308;
        for (int j = lane; j < M; j += 32) m = fmax(m, row[j]);
        m = warp_max(m);
        double s = 0.0;
        for (int j = lane; j < M; j += 32) s += exp(row[j] - m);
        s = warp_sum(s);
        const double l = m + log(s);
        if (lane == 0) {
            lse_s[rr] = l;
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
            double a = 0.0, vv[PHI_DCH] = {0.0, 0.0, 0.0, 0.0};
            if (j < M) {
                const int cell = j / 3, f = j - 3 * cell;
                for (int rr = 0; rr < nr; ++rr) {
                    const int row = r0 + rr;
                    const bool inblock = (cell == row);
                    const bool act = P.t[row] > b;
                    const double S = exp(lp[(size_t)rr * M + j] - lse_s[rr]);
                    const double phi = phi_of(P.model, act, inblock, f, S);
                    a += phi;
#pragma unroll
                    for (int dd = 0; dd < PHI_DCH; ++dd) vv[dd] += phi * Ys[rr][dd];
                    if (d0 == 0) kl += phi * (log(phi) - log_pz(P.model, act, inblock, f, P.prior, row, log_eps));
                }
            }
            if (d0 == 0) P.Apart[(size_t)slot * P.slot_stride + (size_t)chunk * Mp + j] = a;
            for (int dd = 0; dd < nd; ++dd) P.vpart[(size_t)slot * P.slot_stride + ((size_t)chunk * D + d0 + dd) * Mp + j] = vv[dd];
        }
    }
    kl = block_sum(kl, red);
    if (tid == 0) P.klpart[(size_t)slot * P.slot_stride + chunk] = kl;
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
            const int cell = j / 3, f = j - 3 * cell;
            const bool inblock = (cell == row);
            const double S = exp(x[j] - l);
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
            const double sb = SQ_A * (acc - P.kl_weight * (double)nact * (log(phi) + 1.0 - lpz));
            gout[j] = sb;
            r += S * sb;
        }
        r = warp_sum(r);
        for (int j = lane; j < M; j += 32) {
            const double S = exp(x[j] - l);
            gout[j] = S * (gout[j] - r);
        }
    }
}

// nsubs = sub-problems (slots) in the wave;  nitems = nsubs / SD
// Row-resident variant of the backward pass: one CTA per row of logPhi, the row's S and Sbar live in registers
// (EPT elements per thread), so logPhi is read once and the gradient written once, with one exp and one log per element
// (the warp-per-row kernel above re-reads the row and recomputes the exp for its second pass).  grid (N, items).
template <int BT, int EPT>
__global__ void __launch_bounds__(BT) k_phi_backward_rows(PhiParams P) {
    __shared__ double red[32];
    const int tid = threadIdx.x;
    const int row = blockIdx.x, litem = blockIdx.y, item = P.item0 + litem;
    const int N = P.N, M = P.M, Mp = P.Mp, D = P.D, SD = P.SD;
    const double log_eps = log(1e-6);
    const double* Yg = P.Y + (size_t)P.item_gene[item] * N * P.Dtot;
    const double* bvi = P.bv + (size_t)item * SD;
    const size_t slot0 = (size_t)litem * SD;
    const double* x = P.logPhi + ((size_t)item * N + row) * M;
    double* gout = P.grad_logPhi + ((size_t)item * N + row) * M;
    const double ti = P.t[row];
    int nact = 0;
    for (int sd = 0; sd < SD; ++sd) nact += (P.model == MODEL_BGP || ti > bvi[sd]) ? 1 : 0;
    if (nact == 0) {   // uniform over the CTA
        for (int j = tid; j < M; j += BT) gout[j] = 0.0;
        return;
    }
    const bool act0 = ti > bvi[0];
    const double l = P.lse[slot0 * P.slot_stride + row];
    double S[EPT], sb[EPT];
    double r = 0.0;
#pragma unroll
    for (int e = 0; e < EPT; ++e) {
        const int j = tid + e * BT;
        S[e] = 0.0; sb[e] = 0.0;
        if (j < M) {
            const int cell = j / 3, f = j - 3 * cell;
            const bool inblock = (cell == row);
            const double s_ = exp(x[j] - l);
            const double phi = SQ_A * s_ + SQ_B;
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
            const double v = SQ_A * (acc - P.kl_weight * (double)nact * (log(phi) + 1.0 - lpz));
            S[e] = s_; sb[e] = v;
            r += s_ * v;
        }
    }
    r = block_sum(r, red);
#pragma unroll
    for (int e = 0; e < EPT; ++e) {
        const int j = tid + e * BT;
        if (j < M) gout[j] = S[e] * (sb[e] - r);
    }
}

void phi_launch_forward(const PhiParams& P, int nsubs, cudaStream_t st) {
    dim3 grid(P.nch, nsubs);
    k_phi_forward<<<grid, NTHREADS, 0, st>>>(P);
}
void phi_launch_backward(const PhiParams& P, int nitems, cudaStream_t st) {
    dim3 rows(P.N, nitems);
    if (P.M <= 256 * 4) k_phi_backward_rows<256, 4><<<rows, 256, 0, st>>>(P);
    else if (P.M <= 256 * 12) k_phi_backward_rows<256, 12><<<rows, 256, 0, st>>>(P);
    else if (P.M <= 512 * 12) k_phi_backward_rows<512, 12><<<rows, 512, 0, st>>>(P);
    else {   // rows too long for one CTA's registers: warp-per-row two-pass kernel
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
