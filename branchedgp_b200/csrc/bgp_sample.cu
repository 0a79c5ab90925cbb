// Batched Gaussian sampling: MBGP AssignGP.sample_prior and predict_f_samples (MBGP/assigngp.py:741-835).
//
// Both reference functions do the same three steps for every output d of a model:  C_d = cov_d + jitter I,  L_d = chol(C_d),
// samples[s, :, d] = mean_d + L_d z[d, :, s].  Here one CTA owns one output: it builds C_d in a private row-major workspace --
// from the branching kernel with that output's branching point (prior samples: BranchKernelParam.K(x, dim=d),
// MBGP/assigngp.py:103-187, noise_level on the square) or from a given covariance (posterior samples) --, factors it with a
// blocked right-looking Cholesky (32x32 diagonal block in shared memory by one warp, panel and rank-32 trailing update as
// mma.sync.m8n8k4.f64 tiles read straight from L2: the matrices are a few hundred rows) and applies the factor to the normal
// draws the caller passes in (so the host decides the random stream and the tests can compare with numpy draw for draw).
#include "bgp_dense_dev.cuh"
#include "bgp_internal.h"

namespace bgp {

constexpr int SMP_THREADS = 512;

struct SampleArgs {
    int n, np, S, batch;       // rows, rows padded to a multiple of 32, samples per output, outputs
    int from_kernel;           // 1: covariance generated from X / bv / hyp;  0: read from cov
    int kind;
    double ell, kvar, noise, jitter;
    const double* X;           // [n][2] (time, function 1..3)         (from_kernel)
    const double* bv;          // [batch] branching point per output    (from_kernel)
    const double* cov;         // [batch][n][n]                         (!from_kernel)
    const double* mean;        // [batch][n] or nullptr
    const double* z;           // [batch][n][S] standard normal draws
    double* W;                 // [batch][np][np] workspace
    double* out;               // [S][n][batch]
    int* status;               // [batch]
};

__global__ void __launch_bounds__(SMP_THREADS) k_sample(SampleArgs Q) {
    __shared__ double T[TS * (TS + 1)];
    __shared__ double Winv[TS * (TS + 1)];
    __shared__ int fail_s;
    const int d = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nwarps = SMP_THREADS / 32;
    const int n = Q.n, np = Q.np;
    double* A = Q.W + (size_t)d * np * np;
    // ---- covariance (lower triangle incl. diagonal; identity padding)
    const double b = Q.from_kernel ? Q.bv[d] : 0.0;
    const double kinv = 1.0 / (Q.kvar + JITTER);   // cross-covariance denominator k(b,b) + 1e-6 (MBGP/assigngp.py:150-176)
    for (long long e = tid; e < (long long)np * np; e += SMP_THREADS) {
        const int i = (int)(e / np), j = (int)(e % np);
        double v = 0.0;
        if (j <= i) {
            if (i >= n) v = (i == j) ? 1.0 : 0.0;
            else if (Q.from_kernel) {
                const double ti = Q.X[2 * i], tj = Q.X[2 * j];
                const int fi = (int)Q.X[2 * i + 1], fj = (int)Q.X[2 * j + 1];
                if (fi == fj) v = kbase(ti - tj, Q.ell, Q.kvar, Q.kind);
                else v = (kbase(ti - b, Q.ell, Q.kvar, Q.kind) * kinv) * kbase(tj - b, Q.ell, Q.kvar, Q.kind);
                if (i == j) v += Q.noise + Q.jitter;
            } else {
                v = Q.cov[((size_t)d * n + i) * n + j];
                if (i == j) v += Q.jitter;
            }
        }
        A[e] = v;
    }
    if (tid == 0) fail_s = 0;
    __syncthreads();
    // ---- blocked right-looking Cholesky, 32-wide panels
    const int nb = np / TS;
    const int g = lane >> 2, c = lane & 3;
    for (int k = 0; k < nb; ++k) {
        const int k0 = k * TS;
        for (int e = tid; e < TS * TS; e += SMP_THREADS) T[(e >> 5) * (TS + 1) + (e & 31)] = A[(size_t)(k0 + (e >> 5)) * np + k0 + (e & 31)];
        __syncthreads();
        if (warp == 0) {
            int fail = 0;
            chol32_warp(T, TS + 1, fail, k0, lane);
            trinv32_warp(T, TS + 1, Winv, TS + 1, lane);
            if (lane == 0 && fail != 0 && fail_s == 0) fail_s = fail;
        }
        __syncthreads();
        for (int e = tid; e < TS * TS; e += SMP_THREADS) {
            const int i = e >> 5, j = e & 31;
            A[(size_t)(k0 + i) * np + k0 + j] = (j <= i) ? T[i * (TS + 1) + j] : 0.0;
        }
        // panel: L[i, k0:k0+32] = A[i, k0:k0+32] inv(L_kk)^T, one row per thread
        for (int i = k0 + TS + tid; i < np; i += SMP_THREADS) {
            double* row = A + (size_t)i * np + k0;
            for (int j = TS - 1; j >= 0; --j) {   // descending: entry j only needs the still original entries p <= j
                double s0 = 0.0, s1 = 0.0;
                int p = 0;
                for (; p + 1 <= j; p += 2) {
                    s0 = fma(row[p], Winv[j * (TS + 1) + p], s0);       // inv(L)^T[p][j] = inv(L)[j][p], zero for p > j
                    s1 = fma(row[p + 1], Winv[j * (TS + 1) + p + 1], s1);
                }
                if (p <= j) s0 = fma(row[p], Winv[j * (TS + 1) + p], s0);
                row[j] = s0 + s1;
            }
        }
        __syncthreads();
        // trailing update: 32x32 tiles (ti >= tj > k), one warp each, C -= Li Lj^T as 16 DMMA 8x8 tiles over 8 k-steps
        const int r = nb - 1 - k;
        const int ntile = r * (r + 1) / 2;
        for (int tt = warp; tt < ntile; tt += nwarps) {
            int ti = 0;
            while ((ti + 1) * (ti + 2) / 2 <= tt) ++ti;
            const int tj = tt - ti * (ti + 1) / 2;
            const int i0 = (k + 1 + ti) * TS, j0 = (k + 1 + tj) * TS;
            double acc[4][4][2];
#pragma unroll
            for (int mi = 0; mi < 4; ++mi)
#pragma unroll
                for (int ni = 0; ni < 4; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
#pragma unroll 2
            for (int k4 = 0; k4 < 8; ++k4) {
                double af[4], bf[4];
#pragma unroll
                for (int mi = 0; mi < 4; ++mi) af[mi] = A[(size_t)(i0 + 8 * mi + g) * np + k0 + 4 * k4 + c];
#pragma unroll
                for (int ni = 0; ni < 4; ++ni) bf[ni] = A[(size_t)(j0 + 8 * ni + g) * np + k0 + 4 * k4 + c];
#pragma unroll
                for (int mi = 0; mi < 4; ++mi)
#pragma unroll
                    for (int ni = 0; ni < 4; ++ni) dmma(acc[mi][ni][0], acc[mi][ni][1], af[mi], bf[ni]);
            }
#pragma unroll
            for (int mi = 0; mi < 4; ++mi)
#pragma unroll
                for (int ni = 0; ni < 4; ++ni) {
                    double2* p = reinterpret_cast<double2*>(A + (size_t)(i0 + 8 * mi + g) * np + j0 + 8 * ni + 2 * c);
                    double2 v = *p;
                    v.x -= acc[mi][ni][0];
                    v.y -= acc[mi][ni][1];
                    *p = v;
                }
        }
        __syncthreads();
    }
    if (tid == 0) Q.status[d] = fail_s;
    // ---- samples[s][i][d] = mean[d][i] + sum_{j <= i} L[i][j] z[d][j][s]
    const double* zd = Q.z + (size_t)d * n * Q.S;
    for (long long e = tid; e < (long long)n * Q.S; e += SMP_THREADS) {
        const int i = (int)(e / Q.S), s = (int)(e % Q.S);
        const double* row = A + (size_t)i * np;
        double a0 = 0.0, a1 = 0.0;
        int j = 0;
        for (; j + 1 <= i; j += 2) {
            a0 = fma(row[j], zd[(size_t)j * Q.S + s], a0);
            a1 = fma(row[j + 1], zd[(size_t)(j + 1) * Q.S + s], a1);
        }
        if (j <= i) a0 = fma(row[j], zd[(size_t)j * Q.S + s], a0);
        const double m = Q.mean ? Q.mean[(size_t)d * n + i] : 0.0;
        Q.out[((size_t)s * n + i) * Q.batch + d] = m + (a0 + a1);
    }
}

}  // namespace bgp

using namespace bgp;

extern "C" int bgp_gaussian_samples_host(bgp_handle* h, int batch, int n, int S, int from_kernel, const double* X, const double* bv,
                                         double ell, double kvar, int kernel_kind, double noise, const double* cov,
                                         const double* mean, double jitter, const double* z, double* samples, int* status) {
    if (!h) return BGP_ERR_ARG;
    if (batch <= 0 || n <= 0 || S <= 0 || !z || !samples || !status) return fail(h, BGP_ERR_ARG, "batch, n, S must be positive; z, samples, status must be given");
    if (from_kernel && (!X || !bv)) return fail(h, BGP_ERR_ARG, "from_kernel needs X and bv");
    if (!from_kernel && !cov) return fail(h, BGP_ERR_ARG, "a covariance [batch][n][n] is needed when from_kernel = 0");
    if (kernel_kind != BGP_KERNEL_MATERN32 && kernel_kind != BGP_KERNEL_SQEXP) return fail(h, BGP_ERR_ARG, "unknown kernel kind %d", kernel_kind);
    CU(cudaSetDevice(h->device));
    const int np = (n + TS - 1) / TS * TS;
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t r = o; o = align_up(o + bytes, 256); return r; };
    const size_t o_W = take((size_t)batch * np * np * 8), o_z = take((size_t)batch * n * S * 8), o_out = take((size_t)S * n * batch * 8);
    const size_t o_X = take(from_kernel ? (size_t)n * 16 : 0), o_bv = take(from_kernel ? (size_t)batch * 8 : 0);
    const size_t o_cov = take(from_kernel ? 0 : (size_t)batch * n * n * 8), o_mean = take(mean ? (size_t)batch * n * 8 : 0);
    const size_t o_st = take((size_t)batch * 4);
    int rc = ensure_stage(h, o);
    if (rc != BGP_OK) return rc;
    char* s = (char*)h->stage;
    cudaStream_t st = 0;
    CU(cudaMemcpyAsync(s + o_z, z, (size_t)batch * n * S * 8, cudaMemcpyHostToDevice, st));
    if (from_kernel) {
        CU(cudaMemcpyAsync(s + o_X, X, (size_t)n * 16, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(s + o_bv, bv, (size_t)batch * 8, cudaMemcpyHostToDevice, st));
    } else {
        CU(cudaMemcpyAsync(s + o_cov, cov, (size_t)batch * n * n * 8, cudaMemcpyHostToDevice, st));
    }
    if (mean) CU(cudaMemcpyAsync(s + o_mean, mean, (size_t)batch * n * 8, cudaMemcpyHostToDevice, st));
    SampleArgs Q;
    Q.n = n; Q.np = np; Q.S = S; Q.batch = batch; Q.from_kernel = from_kernel; Q.kind = kernel_kind;
    Q.ell = ell; Q.kvar = kvar; Q.noise = noise; Q.jitter = jitter;
    Q.X = (const double*)(s + o_X); Q.bv = (const double*)(s + o_bv); Q.cov = (const double*)(s + o_cov);
    Q.mean = mean ? (const double*)(s + o_mean) : nullptr;
    Q.z = (const double*)(s + o_z); Q.W = (double*)(s + o_W); Q.out = (double*)(s + o_out); Q.status = (int*)(s + o_st);
    k_sample<<<batch, SMP_THREADS, 0, st>>>(Q);
    h->launches++;
    cudaError_t e0 = cudaGetLastError();
    cudaError_t e1 = cudaMemcpyAsync(samples, s + o_out, (size_t)S * n * batch * 8, cudaMemcpyDeviceToHost, st);
    cudaError_t e2 = cudaMemcpyAsync(status, s + o_st, (size_t)batch * 4, cudaMemcpyDeviceToHost, st);
    cudaError_t e3 = cudaStreamSynchronize(st);
    for (cudaError_t e : {e0, e1, e2, e3})
        if (e != cudaSuccess) return fail(h, BGP_ERR_CUDA, "bgp_gaussian_samples_host: %s", cudaGetErrorString(e));
    return BGP_OK;
}
