// Stand-alone check of the shared-memory 128x128 block routines (chol128, trinv128 of csrc/bgp_dense_dev.cuh): one CTA factors and
// inverts a random SPD block; the host compares with a plain double-precision Cholesky / inverse and prints timings.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I branchedgp_b200/csrc -o tools/microbench/block_kernels_test tools/microbench/block_kernels_test.cu
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "bgp_dense_dev.cuh"

using namespace bgp;

__global__ void __launch_bounds__(GEMM_THREADS, 1) k_block(const double* A, double* Lout, double* Xout, int* failout, long long* cyc, int reps, int mode) {
    extern __shared__ __align__(128) unsigned char smem[];
    double* Dm = reinterpret_cast<double*>(smem);
    const int tid = threadIdx.x;
    long long c_chol = 0, c_inv = 0;
    int fail = 0;
    for (int r = 0; r < reps; ++r) {
        for (int e = tid; e < MB * MB; e += GEMM_THREADS) Dm[(e / MB) * DLD + (e % MB)] = A[e];
        __syncthreads();
        long long t0 = clock64();
        fail = 0;
        if (mode != 2) chol128(Dm, fail, 0);
        long long t1 = clock64();
        if (r == reps - 1)
            for (int e = tid; e < MB * MB; e += GEMM_THREADS) Lout[e] = ((e % MB) <= (e / MB)) ? Dm[(e / MB) * DLD + (e % MB)] : 0.0;
        __syncthreads();
        long long t2 = clock64();
        if (mode != 1) trinv128(Dm);
        long long t3 = clock64();
        c_chol += t1 - t0; c_inv += t3 - t2;
        __syncthreads();
    }
    for (int e = tid; e < MB * MB; e += GEMM_THREADS) Xout[e] = ((e % MB) <= (e / MB)) ? Dm[(e / MB) * DLD + (e % MB)] : 0.0;
    if (tid == 0) { *failout = fail; cyc[0] = c_chol / reps; cyc[1] = c_inv / reps; }
}

int main(int argc, char** argv) {
    const int mode = argc > 1 ? atoi(argv[1]) : 0;   // 0 both, 1 chol128 only, 2 trinv128 only (on the input matrix's lower triangle)
    const int n = MB;
    std::vector<double> A(n * n), L(n * n), X(n * n);
    srand(1);
    std::vector<double> B(n * n);
    for (auto& v : B) v = rand() / (double)RAND_MAX - 0.5;
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) {
            double s = (i == j) ? 0.5 : 0.0;
            for (int k = 0; k < n; ++k) s += B[i * n + k] * B[j * n + k] / n;
            A[i * n + j] = s;
        }
    double *dA, *dL, *dX; int* dF; long long* dC;
    cudaMalloc(&dA, n * n * 8); cudaMalloc(&dL, n * n * 8); cudaMalloc(&dX, n * n * 8); cudaMalloc(&dF, 4); cudaMalloc(&dC, 16);
    cudaMemcpy(dA, A.data(), n * n * 8, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(k_block, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    k_block<<<1, GEMM_THREADS, SMEM_BYTES>>>(dA, dL, dX, dF, dC, 20, mode);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("{\"error\": \"%s\"}\n", cudaGetErrorString(e)); return 1; }
    int fail; long long cyc[2];
    cudaMemcpy(L.data(), dL, n * n * 8, cudaMemcpyDeviceToHost); cudaMemcpy(X.data(), dX, n * n * 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(&fail, dF, 4, cudaMemcpyDeviceToHost); cudaMemcpy(cyc, dC, 16, cudaMemcpyDeviceToHost);
    // host reference
    std::vector<double> R(n * n, 0.0);
    for (int j = 0; j < n; ++j) {
        double d = A[j * n + j];
        for (int k = 0; k < j; ++k) d -= R[j * n + k] * R[j * n + k];
        R[j * n + j] = sqrt(d);
        for (int i = j + 1; i < n; ++i) {
            double s = A[i * n + j];
            for (int k = 0; k < j; ++k) s -= R[i * n + k] * R[j * n + k];
            R[i * n + j] = s / R[j * n + j];
        }
    }
    double eL = 0.0, eX = 0.0;
    for (int i = 0; i < n * n; ++i) eL = fmax(eL, fabs(L[i] - R[i]));
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) {
            double s = 0.0;
            for (int k = 0; k < n; ++k) s += X[i * n + k] * R[k * n + j];
            eX = fmax(eX, fabs(s - (i == j ? 1.0 : 0.0)));
        }
    printf("{\"fail\": %d, \"max_abs_err_chol\": %.3e, \"max_abs_err_inv_times_L_minus_I\": %.3e, \"chol128_cycles\": %lld, \"trinv128_cycles\": %lld}\n",
           fail, eL, eX, cyc[0], cyc[1]);
    return 0;
}
