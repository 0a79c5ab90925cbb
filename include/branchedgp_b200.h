/* branchedgp_b200 -- C ABI of the B200-native evaluator for the BranchedGP hot path.
 *
 * The reference (ManchesterBioinference/BranchedGP, pure Python on TensorFlow/GPflow) has no FFI of its own; the
 * boundary this library sits behind is the Python call the L-BFGS driver makes once per function evaluation:
 *
 *   AssignGP.maximum_log_likelihood_objective()            BranchedGP/assigngp_dense.py:151-199   (+ TF autodiff)
 *   AssignGPSparse.maximum_log_likelihood_objective()      BranchedGP/assigngp_denseSparse.py:59-107
 *   MBGP AssignGP._build_likelihood_sparse / _single_b     BranchedGP/MBGP/assigngp.py:492-558, 608-647
 *   AssignGP.GetPhi()                                      BranchedGP/assigngp_dense.py:130-140
 *
 * Each entry point evaluates the bound (and its gradient) for `count` independent work items
 * (gene g, branching point b, hyper-parameters, logPhi) in one call.  Conventions:
 *   - all reals are IEEE float64, row-major; `M = 3*N` columns of logPhi are cell-major (column 3*cell + f), the
 *     order produced by VBHelperFunctions.GetFunctionIndexListGeneral (VBHelperFunctions.py:170-194);
 *   - pointers of the plain entry points are DEVICE pointers owned by the caller; the `_host` variants take HOST
 *     pointers and do the host<->device copies themselves; the library owns only its workspace;
 *   - return value 0 = success, otherwise an error code (bgp_last_error gives the text); never throws;
 *   - status[i] = 0, or k>0 when the Cholesky of item i met a non-positive pivot at row k (<= Mp: K, > Mp: P),
 *     which the Python layer turns into the exception the reference gets from tf.linalg.cholesky;
 *   - one handle per (GPU, host thread); calls on one handle are serialised on the given CUDA stream.
 */
#ifndef BRANCHEDGP_B200_H
#define BRANCHEDGP_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct bgp_handle bgp_handle;

enum { BGP_KERNEL_MATERN32 = 0, BGP_KERNEL_SQEXP = 1 };   /* gpflow.kernels.Matern32 / SquaredExponential */
enum { BGP_MODEL_BGP = 0, BGP_MODEL_MBGP = 1 };           /* assigngp_dense.AssignGP / MBGP.assigngp.AssignGP (multi=False) */

enum {
    BGP_OK = 0,
    BGP_ERR_CUDA = 1,         /* a CUDA runtime call failed */
    BGP_ERR_ARG = 2,          /* invalid argument */
    BGP_ERR_NOMEM = 3         /* workspace does not fit in device memory */
};

int bgp_version(void);
int bgp_create(int device, bgp_handle** out);
int bgp_destroy(bgp_handle* h);
const char* bgp_last_error(const bgp_handle* h);
int bgp_device_sm_count(const bgp_handle* h);

/* Upper bound on work items processed concurrently (one workspace slot each).  0 restores the default (one per SM). */
int bgp_set_max_slots(bgp_handle* h, int max_slots);
/* Bytes of library workspace one dense / sparse work item needs at the given shape. */
size_t bgp_dense_slot_bytes(int N, int D);

/* Pinned host memory for the `_host` entry points (plain malloc'ed buffers work too, just slower). */
int bgp_host_alloc(void** p, size_t bytes);
int bgp_host_free(void* p);

/* Dense AssignGP bound + gradient.  Replaces AssignGP.maximum_log_likelihood_objective (assigngp_dense.py:151-199)
 * and the TF autodiff gradient that gpflow.optimizers.Scipy requests (FitBranchingModel.py:127-132).
 *
 *   t[N]                 pseudotime of each cell
 *   Y[nY][N][D]          expression; item i uses Y[item_gene[i]]
 *   b[count]             candidate branching point of item i (kernel and trunk/branch split t <= b)
 *   prior                BGP: phiPrior[N][2];  MBGP: prior[N][3]
 *   hyp[count][3]        lengthscale, kernel variance, likelihood variance (constrained values)
 *   logPhi[count][N][3N] variational logits
 *   KConst[3N][3N]       optional fixed covariance (AssignGP(KConst=...), assigngp_dense.py:156-158), else NULL
 *   kl_weight            1 for BGP, D for MBGP multi=False (MBGP/assigngp.py:637)
 *   want_grad            0: bound only
 * outputs
 *   elbo[count]          the bound (without hyper-priors)
 *   grad_hyp[count][4]   d elbo / d (lengthscale, kernel variance, likelihood variance, b)   (may be NULL if !want_grad)
 *   grad_logPhi[count][N][3N]                                                                 (may be NULL if !want_grad)
 *                        -- doubles as scratch during the call (the forward assignment pass parks the softmax values there
 *                        for the backward pass), so it must not alias logPhi
 *   status[count]
 */
int bgp_dense_elbo_grad(bgp_handle* h, int count, int N, int D, const double* t, const double* Y, int nY,
                        const int* item_gene, const double* b, const double* prior, const double* hyp, int kernel_kind,
                        int model, double kl_weight, const double* logPhi, const double* KConst, int want_grad,
                        double* elbo, double* grad_hyp, double* grad_logPhi, int* status, void* cuda_stream);

/* Same, HOST pointers; copies in/out inside the call and synchronises before returning. */
int bgp_dense_elbo_grad_host(bgp_handle* h, int count, int N, int D, const double* t, const double* Y, int nY,
                             const int* item_gene, const double* b, const double* prior, const double* hyp,
                             int kernel_kind, int model, double kl_weight, const double* logPhi, const double* KConst,
                             int want_grad, double* elbo, double* grad_hyp, double* grad_logPhi, int* status);

/* Inducing-point bounds + gradients.
 *   multi == 0: AssignGPSparse.maximum_log_likelihood_objective (assigngp_denseSparse.py:59-107): one branching point per
 *               item (b[count]), all Dtot outputs share Phi; model = BGP_MODEL_BGP, prior[N][2].
 *   multi == 1: MBGP AssignGP._build_likelihood_sparse (MBGP/assigngp.py:492-558): one branching point per OUTPUT
 *               (b[count][Dtot]), per-output trunk masks on a shared logPhi; model = BGP_MODEL_MBGP, prior[N][3].
 *   Z[Mz][2]    inducing inputs (time, function in {1,2,3}), Mz <= 208
 * outputs: elbo[count], grad_hyp[count][3] (d/d lengthscale, kernel variance, likelihood variance),
 *          grad_b[count][multi ? Dtot : 1], grad_logPhi[count][N][3N], status[count]. */
int bgp_sparse_elbo_grad(bgp_handle* h, int count, int N, int Dtot, int Mz, const double* t, const double* Z, const double* Y,
                         int nY, const int* item_gene, const double* b, const double* prior, const double* hyp,
                         int kernel_kind, int model, int multi, const double* logPhi, int want_grad, double* elbo,
                         double* grad_hyp, double* grad_b, double* grad_logPhi, int* status, void* cuda_stream);
int bgp_sparse_elbo_grad_host(bgp_handle* h, int count, int N, int Dtot, int Mz, const double* t, const double* Z,
                              const double* Y, int nY, const int* item_gene, const double* b, const double* prior,
                              const double* hyp, int kernel_kind, int model, int multi, const double* logPhi, int want_grad,
                              double* elbo, double* grad_hyp, double* grad_b, double* grad_logPhi, int* status);

/* predict_f (assigngp_dense.py:201-240; MBGP/assigngp.py:560-647 per output): posterior mean[count][T][D] and marginal variance
 * var[count][T] (full_cov: var[count][T][T], T <= 128) of the latent functions at Xnew[T][2] = (time, function in {1,2,3}).
 * HOST pointers; same model arguments as bgp_dense_elbo_grad_host. */
int bgp_dense_predict_host(bgp_handle* h, int count, int N, int D, const double* t, const double* Y, int nY, const int* item_gene,
                           const double* b, const double* prior, const double* hyp, int kernel_kind, int model,
                           const double* logPhi, int T, const double* Xnew, int full_cov, double* mean, double* var);
/* AssignGPSparse.predict_f (assigngp_denseSparse.py:109-155), one model: mean[T][Dtot], var[T] or [T][T]. */
int bgp_sparse_predict_host(bgp_handle* h, int N, int Dtot, int Mz, const double* t, const double* Z, const double* Y, double b,
                            const double* prior, const double* hyp, int kernel_kind, const double* logPhi, int T,
                            const double* Xnew, int full_cov, double* mean, double* var);

/* AssignGP.GetPhi (assigngp_dense.py:130-140): softmax rows of logPhi gathered at the cell's own 3 columns -> Phi[count][N][3].
 * HOST pointers. */
int bgp_get_phi_host(bgp_handle* h, int count, int N, const double* logPhi, double* phi);

/* ---- batched fitting ---------------------------------------------------------------------------------------------
 * Replaces, for `count` work items at once, the optimiser loop of FitBranchingModel.FitModel (FitBranchingModel.py:96-140):
 *   m.UpdateBranchingPoint(b, phiInitial); gpflow.optimizers.Scipy().minimize(m.training_loss, m.trainable_variables,
 *   options=dict(maxiter=maxiter)); ll = m.log_posterior_density(); Phi = m.GetPhi(); predictBranchingModel(m)
 * with logPhi, its gradient and the L-BFGS history resident on the GPU (bgp_fit.cu).  The optimiser restates SciPy's
 * L-BFGS-B (unbounded) and its More'-Thuente line search; defaults below are SciPy's. */
typedef struct bgp_fit_options {
    int maxiter;            /* L-BFGS iterations per item (FitModel's maxiter, default 100) */
    int maxcor;             /* stored correction pairs, <= 16 (SciPy maxcor = 10) */
    int maxls;              /* evaluations per line search (20) */
    int maxfun;             /* evaluations per item (15000) */
    double ftol;            /* stop when (f_k - f_{k+1}) / max(|f_k|, |f_{k+1}|, 1) <= ftol (2.22e-9) */
    double gtol;            /* stop when max |g| <= gtol (1e-5) */
    int train_hyp;          /* 1: lengthscale, kernel variance and likelihood variance are optimised with logPhi (softplus
                               transforms, GPflow positive()); 0: FitModel(fixHyperparameters=True) */
    int has_prior[3];       /* Normal priors on the constrained (lengthscale, kernel variance, likelihood variance) */
    double prior_mean[3];
    double prior_sd[3];
    int max_slots;          /* items fitted concurrently; 0 = one per SM, reduced to what fits in device memory */
} bgp_fit_options;

void bgp_fit_default_options(bgp_fit_options* o);

/* Dense AssignGP (FitModel(M=0)).  HOST pointers.
 *   t[N], Y[nY][N][D], item_gene[count], b[count], prior = phiPrior[N][2], phiInitial[N][2] (InitialiseVariationalPhi,
 *   assigngp_dense.py:92-128, is evaluated on the device per item), hyp0[count][3] initial constrained hyper-parameters;
 *   Xnew[count][T][2] test inputs per item (T = 0: no prediction).
 * outputs per item: objective = log posterior density at the optimum (FitModel's loglik entry), elbo (the bound alone),
 *   hyp[3], phi[N][3] (GetPhi), pred_mean[T][D], pred_var[T] (predict_f), iters, nfev,
 *   status: 0 converged (gradient), 1 converged (function decrease), 2 maxiter, 3 maxfun, 4 abnormal line search,
 *           5 a Cholesky factorisation failed (objective = NaN, like FitModel's exception branch);
 *   logPhi_out[count][N][3N] optional (NULL to skip). */
int bgp_dense_fit_host(bgp_handle* h, int count, int N, int D, const double* t, const double* Y, int nY, const int* item_gene,
                       const double* b, const double* prior, const double* phiInitial, const double* hyp0, int kernel_kind,
                       const bgp_fit_options* opt, int T, const double* Xnew, double* objective, double* elbo, double* hyp,
                       double* phi, double* pred_mean, double* pred_var, int* iters, int* nfev, int* status, double* logPhi_out);
/* Same for AssignGPSparse (FitModel(M>0)) with inducing inputs Z[Mz][2]. */
int bgp_sparse_fit_host(bgp_handle* h, int count, int N, int D, int Mz, const double* t, const double* Z, const double* Y, int nY,
                        const int* item_gene, const double* b, const double* prior, const double* phiInitial, const double* hyp0,
                        int kernel_kind, const bgp_fit_options* opt, int T, const double* Xnew, double* objective, double* elbo,
                        double* hyp, double* phi, double* pred_mean, double* pred_var, int* iters, int* nfev, int* status,
                        double* logPhi_out);

/* Result gather over NCCL (NVLink / NVSwitch) for host programs that only have this C library.  Work items are independent, so the
 * one exchange of a multi-GPU sweep is the collection of the per-item results.  Every rank: bgp_create on its GPU; rank 0:
 * bgp_comm_unique_id (128 bytes, shipped to the other ranks by the caller); every rank: bgp_comm_init; then
 * bgp_gather_results(local rows) -> all[nranks][rows_max][k] on every rank (rows beyond a rank's rows_local are zero).
 * libnccl.so.2 is opened at run time.  HOST pointers.  (branchedgp_b200/sharding.py does the same through torch.distributed.) */
int bgp_comm_unique_id(bgp_handle* h, void* id128);
int bgp_comm_init(bgp_handle* h, int nranks, int rank, const void* id128);
int bgp_comm_destroy(bgp_handle* h);
int bgp_gather_results(bgp_handle* h, const double* local, int rows_local, int k, int rows_max, double* all);

/* Batched Gaussian sampling: MBGP AssignGP.sample_prior (MBGP/assigngp.py:800-835) and the correlated branch of
 * predict_f_samples (MBGP/assigngp.py:741-798).  For every output d = 0 .. batch-1:
 *     C_d = cov_d + jitter I,   L_d = chol(C_d),   samples[s][i][d] = mean[d][i] + sum_j L_d[i][j] z[d][j][s]
 * from_kernel = 1: cov_d = BranchKernelParam.K(X, dim=d) (MBGP/assigngp.py:103-187) built on the device from X[n][2]
 *                  (time, function 1..3), the branching point bv[d], the base kernel (kernel_kind, ell, kvar) and `noise`
 *                  (noise_level) on the diagonal;   from_kernel = 0: cov[batch][n][n] is given (posterior covariance).
 * mean[batch][n] may be NULL (zero mean).  z[batch][n][S] are standard normal draws supplied by the caller (the reference draws
 * tf.random.normal([D, N, S])), samples is [S][n][batch] like the reference's return value.  status[d] != 0: non-positive pivot.
 * All pointers are HOST pointers. */
int bgp_gaussian_samples_host(bgp_handle* h, int batch, int n, int S, int from_kernel, const double* X, const double* bv,
                              double ell, double kvar, int kernel_kind, double noise, const double* cov, const double* mean,
                              double jitter, const double* z, double* samples, int* status);

/* Per-kernel device timing (CUDA events recorded on the caller's stream around every launch).  Phase ids:
 * 0 phi forward, 1 prep, 2 chol K, 3 P = L^T D L + I, 4 chol P, 5 solves, 6 trtri, 7 lauum, 8 trmm, 9 post, 10 adjoint,
 * 11 phi backward.  bgp_profile_read waits for the recorded events, accumulates milliseconds per phase since the last
 * reset into ms_out[BGP_NPHASES] and returns the number of kernel launches made since the last reset. */
#define BGP_NPHASES 12
int bgp_profile_enable(bgp_handle* h, int on);
int bgp_profile_reset(bgp_handle* h);
int bgp_profile_read(bgp_handle* h, double* ms_out, long long* launches);

/* ---- test hooks (used by tests/ only) ------------------------------------------------------------------------
 * Stop the dense pipeline after phase `phase` (1 chol K, 2 P, 3 chol P, 4 solve, 5 trtri, 6 lauum, 7 trmm, 8 post,
 * 9 adjoint; 0 = run everything) and read back workspace matrices / vectors of a slot of the last wave. */
int bgp_debug_set_stop_phase(bgp_handle* h, int phase);
int bgp_debug_padded_size(int N);
/* which: 0 LC (chol K), 1 RX (P / R / R^-1), 2 PB (Pbar), 3 LB (Lbar / Ktilde); out_host[Mp*Mp] row-major */
int bgp_debug_dense_matrix(bgp_handle* h, int slot, int which, int symmetric, double* out_host);
/* which: 0 A, 1 Delta, 2 v, 3 z, 4 c, 5 q, 6 vbar, 7 delta, 8 Abar; out_host[len] (len = Mp or D*Mp) */
int bgp_debug_dense_vector(bgp_handle* h, int slot, int which, double* out_host);

/* Cycle counters of the tile GEMM pipeline in k_lauum, filled only by a -DBGP_TIMING build (make timing; tools/gemm_timers.py):
 * [0] cycles a consumer warp waited for the first stages of a GEMM call, [1] for later stages, [2] producer waits for a free
 * stage, [3] cycles inside the GEMM calls, [4] calls, [5] k-steps; summed over CTAs. */
int bgp_debug_gemm_timers(bgp_handle* h, unsigned long long* out8, int reset);
/* Same for the task sections of the multi-CTA dense pipeline (bgp_dense_wide.cu): 48 counters, zeros unless BGP_TIMING. */
int bgp_debug_wide_timers(bgp_handle* h, unsigned long long* out48, int reset);

/* Host-only hooks for the optimiser pieces of bgp_fit.cu (no GPU needed): one More'-Thuente line-search call on caller-owned
 * state[32] (returns 0 evaluate at *stp, 1 convergence, 2 warning, 3 error), and the limited-memory direction d = -H g
 * computed from inner products of S[col][n], Y[col][n] (oldest pair first) and g[n]. */
int bgp_debug_linesearch(double* state, int start, double* stp, double f, double g);
int bgp_debug_lbfgs_direction(int col, int n, const double* S, const double* Y, const double* g, double* d);

#ifdef __cplusplus
}
#endif
#endif /* BRANCHEDGP_B200_H */
