// Parameters of the assignment (Phi) passes shared by the dense, sparse and MBGP pipelines.
#pragma once
#include "bgp_common.cuh"

namespace bgp {

constexpr int PHI_RCH = 16;   // rows of logPhi handled by one CTA of the forward pass
constexpr int PHI_KLW = 8;    // KL partials kept per row by the row-resident forward pass (one per CTA of the row's cluster)
// Streaming passes (rows too long for a CTA's registers, bound + gradient calls of single-output BGP items): a CTA covers
// PHI_SCOLS columns x rch rows, rch = PHI_SRCH where the caller sizes its partial arrays for it (inducing-point path), else PHI_RCH.
constexpr int PHI_SCOLS = 1024;
constexpr int PHI_SRCH = 64;
inline int phi_col_chunks(int M) { return (M + PHI_SCOLS - 1) / PHI_SCOLS; }
bool phi_streaming(int N, int M, int SD, int D, bool want_grad, int rch);

// A "sub-problem" is one (work item, output group): dense / sparse BGP have one per item carrying all D outputs
// (SD = 1, D = Dtot); the MBGP multi-output bound has one per output (SD = Dtot, D = 1) because the trunk mask
// t_i > b_d differs per output (MBGP/assigngp.py:217-251).  Workspace slot s of a wave is sub-problem s; the slots of
// one item are contiguous.
struct PhiParams {
    int N, M, Mp, D, nch, model;
    int rch;                   // rows per chunk of the partial arrays (nch = ceil(N / rch)); 0 means PHI_RCH
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
    double* klrow;             // [slot][N][PHI_KLW]  scratch of the row-resident forward pass (streaming: [nch][column chunks] KL partials)
    double* rowpart;           // [slot][column chunks][N]  per-row partial sums of the streaming passes
    // backward inputs (per slot) and output
    const double* Abar;        // [Mp]
    const double* vbar;        // [D][Mp]
    double* grad_logPhi;       // [count][N][M]
};

void phi_launch_forward(const PhiParams& P, int nitems, cudaStream_t st);
void phi_launch_backward(const PhiParams& P, int nitems, cudaStream_t st);

}  // namespace bgp
