// Parameters of the assignment (Phi) passes shared by the dense, sparse and MBGP pipelines.
#pragma once
#include "bgp_common.cuh"

namespace bgp {

constexpr int PHI_RCH = 16;   // rows of logPhi handled by one CTA of the forward pass
constexpr int PHI_KLW = 8;    // KL partials kept per row by the row-resident forward pass (one per CTA of the row's cluster)

// A "sub-problem" is one (work item, output group): dense / sparse BGP have one per item carrying all D outputs
// (SD = 1, D = Dtot); the MBGP multi-output bound has one per output (SD = Dtot, D = 1) because the trunk mask
// t_i > b_d differs per output (MBGP/assigngp.py:217-251).  Workspace slot s of a wave is sub-problem s; the slots of
// one item are contiguous.
struct PhiParams {
    int N, M, Mp, D, nch, model;
    int SD, Dtot;              // sub-problems per item; outputs per item (row length of Y)
    int item0;
    long long slot_stride;     // doubles between consecutive slots for lse / Apart / vpart / klpart / Abar / vbar
    double kl_weight;
    const double* logPhi;      // [count][N][M] row-major (M = 3N, cell-major columns 3*cell + f)
    const double* t;           // [N] pseudotime
    const double* prior;       // BGP: [N][2] phiPrior;  MBGP: [N][3]
    const double* bv;          // [count][SD] branching point used for the trunk / branch split (t <= b  => trunk)
    const double* Y;           // [nY][N][Dtot]
    const int* item_gene;      // [count]
    double* lse;               // [slot][N]  row log-sum-exp of logPhi
    // forward outputs (per slot, partial over row chunks; reduced by the consumer in a fixed order => deterministic)
    double* Apart;             // [slot][nch][Mp]
    double* vpart;             // [slot][nch][D][Mp]
    double* klpart;            // [slot][nch]
    double* klrow;             // [slot][N][PHI_KLW]  scratch of the row-resident forward pass
    // backward inputs (per slot) and output
    const double* Abar;        // [Mp]
    const double* vbar;        // [D][Mp]
    double* grad_logPhi;       // [count][N][M]
};

void phi_launch_forward(const PhiParams& P, int nitems, cudaStream_t st);
void phi_launch_backward(const PhiParams& P, int nitems, cudaStream_t st);

}  // namespace bgp
