// Host-side launchers of the kernel files (internal to the shared library).
#pragma once
#include "bgp_dense.cuh"
#include "bgp_phi.cuh"
#include "bgp_sparse.cuh"

namespace bgp {
cudaError_t dense_init_kernels();
void dense_launch_prep(const DenseParams& P, const double* Apart, const double* vpart, const double* klpart, const double* Y,
                       const int* item_gene, int nitems, cudaStream_t st);
void dense_launch_phase(const DenseParams& P, int nitems, int phase, cudaStream_t st);
// multi-CTA-per-item pipeline (bgp_dense_wide.cu): G CTAs per item; false = the phase has no wide kernel (use dense_launch_phase)
cudaError_t dense_wide_init_kernels();
long long dense_wide_ctrl_ints(int nM);
bool dense_launch_phase_wide(const DenseParams& P, int nitems, int phase, int G, cudaStream_t st);
void dense_launch_predict(const DenseParams& P, int nitems, const double* Xnew, int T, int Ttot, int s0, int full_cov, double* mean,
                          double* var, cudaStream_t st, int rx_inverse = 0);
void dense_launch_predict_slots(const DenseParams& P, int nitems, const double* Xnew, long long xnew_stride, int Ttot, double* mean,
                                double* var, cudaStream_t st);
size_t dense_predict_cov_scratch_doubles(const DenseDims& d, int nitems, int Ttot);
void dense_launch_predict_fullcov(const DenseParams& P, int nitems, const double* Xnew, int Ttot, double* mean, double* var, double* ext,
                                  const double* neg1, cudaStream_t st);
void dense_fill(double* p, long long n, double v, cudaStream_t st);
void dense_read_gemm_timers(unsigned long long* out8, int reset);
void dense_read_wide_timers(unsigned long long* out48, int reset);
void dense_unpack(const double* packed, double* out, int Mp, int symmetric, cudaStream_t st);
cudaError_t sparse_init_kernels();
void sparse_launch_phase(const SparseParams& P, int nslots, int phase, double* sub_out, cudaStream_t st);
void sparse_launch_predict(const SparseParams& P, const double* Xnew, int T, int full_cov, double* scratch, double* mean_out,
                           double* var_out, cudaStream_t st);
void phi_launch_gather(const double* logPhi, double* phi, int count, int N, cudaStream_t st);
}  // namespace bgp
