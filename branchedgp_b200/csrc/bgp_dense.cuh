// Parameter block and workspace layout of the dense AssignGP pipeline (assigngp_dense.py:151-199 + gradient).
#pragma once
#include "bgp_common.cuh"

namespace bgp {

constexpr int MODEL_BGP = 0;    // assigngp_dense.AssignGP: pZ from pZ_construction_singleBP, White(1e-6) summand
constexpr int MODEL_MBGP = 1;   // MBGP.assigngp.AssignGP: per-output masks, phiTrunk, squashed pZ, noise_level 1e-6

// scalar slots (per work item)
enum { SC_LOGDET = 0, SC_SUMC2 = 1, SC_KL = 2, SC_SUMY2 = 3, SC_ADELTA = 4, SC_TRACE = 5, SC_COUNT = 8 };

struct DenseDims {
    int N, M, Mp, nT, nM, D, nch;   // cells, 3N, padded, tiles per side, macro blocks per side, outputs, row chunks of the Phi pass
    int nTv;                        // tiles that contain a row < M: k-ranges stop here (identity padding contributes nothing to valid entries)
    long long mat_elems;            // doubles in one packed block-lower matrix
    long long vec_elems;            // doubles in one per-item vector area
};

// offsets (in doubles) inside the per-item vector area
struct VecLayout {
    long long A, Delta, v, z, c, q, vbar, delta, Abar, ka, dkal, dkab, dpart, scal, total;
};

__host__ __device__ inline VecLayout vec_layout(int Mp, int nM, int D) {
    VecLayout L;
    long long o = 0;
    L.A = o; o += Mp;
    L.Delta = o; o += Mp;
    L.v = o; o += (long long)D * Mp;
    L.z = o; o += (long long)D * Mp;
    L.c = o; o += (long long)D * Mp;
    L.q = o; o += (long long)D * Mp;
    L.vbar = o; o += (long long)D * Mp;
    L.delta = o; o += Mp;
    L.Abar = o; o += Mp;
    L.ka = o; o += Mp;
    L.dkal = o; o += Mp;
    L.dkab = o; o += Mp;
    L.dpart = o; o += (long long)nM * Mp;
    L.scal = o; o += SC_COUNT;
    L.total = (o + 15) & ~15LL;
    return L;
}

struct DenseParams {
    DenseDims d;
    int kind, model;
    double square_noise;      // White(1e-6) summand / MBGP noise_level on the square kernel
    double kl_weight;         // 1 (BGP) or D (MBGP single-b, MBGP/assigngp.py:637)
    int item0;                // slot s of this wave is item item0 + s
    const int* slot_map;      // optional: CTA c works on slot slot_map[c] (subset launches of the batched fit); nullptr = identity
    long long slot_stride;    // doubles between consecutive slots, common to every per-slot array below
    // inputs (device)
    const double* t;          // [N]
    const double* hyp;        // [count][3]  ell, kernel variance, likelihood variance (constrained values)
    const double* bv;         // [count]     branching point
    const double* KConst;     // nullable [M][M] row-major debug covariance (assigngp_dense.py:156-158)
    // per-slot workspace
    double* LC; double* RX; double* PB; double* LB;   // packed block-lower tile matrices
    double* LCe;              // nT diagonal tiles of chol(K) + jitter*I
    double* LinvD; double* RinvD;                     // nM blocks of 16 tiles (inverse diagonal macro blocks)
    double* SCR;              // 2 blocks of 16 tiles
    double* KT;               // [2][N][N] base-kernel table k(t_i - t_j) and its lengthscale derivative (same-function entries of K)
    double* vec;              // vector area, see vec_layout()
    int* status;              // [count]
    int* ctrl;                // task queues / dependency counters of the multi-CTA pipeline (bgp_dense_wide.cu), zeroed per evaluation
    long long ctrl_stride;    // ints between the control areas of consecutive slots
    // outputs (device)
    double* elbo;             // [count]
    double* grad_hyp;         // [count][4]  d/d ell, var, likvar, b
};

}  // namespace bgp
