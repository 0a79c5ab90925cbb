// Inducing-point bounds: AssignGPSparse (assigngp_denseSparse.py:59-107) and MBGP _build_likelihood_sparse
// (MBGP/assigngp.py:492-558).  One "sub-problem" per (item, output group), see bgp_phi.cuh.
#pragma once
#include "bgp_common.cuh"

namespace bgp {

constexpr int SP_SL = 32;          // columns of X handled per slab
constexpr int SP_MAX_MZ = 208;     // packed lower triangle of Mz x Mz must fit in shared memory

struct SparseLayout {   // offsets in doubles inside one sub-problem slot
    long long L, Linv, LinvT, Ppart, R, Rinv, Pbar, Lbpart, T, W, S, V, A, Delta, v, Abar, vbar, zpart, z, c, q, zbar,
        kax, dkalx, dkabx, kaz, dkalz, dkabz, scal, hpart, Apart, vpart, klpart, lse, klrow, rowpart, total;
};

enum { SS_KL = 0, SS_SUMY2, SS_LOGDET, SS_SUMC2, SS_TRACE, SS_ADELTA, SS_SUMA, SS_COUNT = 8 };

__host__ __device__ inline SparseLayout sparse_layout(int N, int Mp, int Mz, int D, int nsplit, int nch) {
    SparseLayout L;
    long long o = 0;
    const long long mm = (long long)Mz * Mz;
    auto take = [&](long long n) { long long r = o; o += (n + 3) & ~3LL; return r; };
    L.L = take(mm); L.Linv = take(mm); L.LinvT = take(mm);
    L.Ppart = take(mm * nsplit);
    L.R = take(mm); L.Rinv = take(mm); L.Pbar = take(mm);
    L.Lbpart = take(mm * nsplit);
    L.T = take(mm); L.W = take(mm); L.S = take(mm);
    L.V = take((long long)Mz * Mp);
    L.A = take(Mp); L.Delta = take(Mp); L.v = take((long long)D * Mp); L.Abar = take(Mp); L.vbar = take((long long)D * Mp);
    L.zpart = take((long long)nsplit * D * Mz); L.z = take((long long)D * Mz); L.c = take((long long)D * Mz);
    L.q = take((long long)D * Mz); L.zbar = take((long long)D * Mz);
    L.kax = take(N); L.dkalx = take(N); L.dkabx = take(N);
    L.kaz = take(Mz); L.dkalz = take(Mz); L.dkabz = take(Mz);
    L.scal = take(SS_COUNT);
    L.hpart = take((long long)nsplit * 8);     // per split: trace, Adelta, g_ell, g_var, g_b
    L.Apart = take((long long)nch * Mp); L.vpart = take((long long)nch * D * Mp); L.klpart = take(nch); L.lse = take(N);
    L.klrow = take((long long)N * 8);   // PHI_KLW per-row KL partials of the row-resident assignment pass
    L.rowpart = take((long long)N * ((3 * N + 1023) / 1024));   // [column chunks][N] of the streaming assignment passes
    L.total = (o + 31) & ~31LL;
    return L;
}

struct SparseParams {
    int N, M, Mp, Mz, D, nch, nsplit;   // D = outputs handled by one sub-problem
    int SD, Dtot;                       // sub-problems per item, outputs per item
    int kind, model, want_grad;
    double square_noise, kl_weight;
    int item0;                          // first ITEM of the wave; slot s <-> sub-problem item0*SD + s
    long long slot_stride;
    SparseLayout lay;                   // offsets inside a slot (sparse_layout(N, Mp, Mz, D, nsplit, nch), computed once on the host)
    const double* t;         // [N]
    const double* Z;         // [Mz][2]  (time, function 1..3)
    const double* hyp;       // [count][3]
    const double* bv;        // [count][SD]
    const double* Y;         // [nY][N][Dtot]
    const int* item_gene;    // [count]
    double* ws;              // workspace base (slot 0)
    int* status;             // [count]
    double* elbo;            // [count]
    double* grad_hyp;        // [count][3]   d/d ell, var, likvar
    double* grad_b;          // [count][SD]
    int dbg_skip;            // timing experiments only (BGP_SP_SKIP bit mask: 1 slab products, 2 outer products, 4 K generation); results are wrong when set
};

}  // namespace bgp
