// FP64 peak micro-benchmarks for B200 (sm_100a): DFMA issue rate, DMMA.8x8x4 issue rate,
// DMMA fed from shared memory. Prints one JSON line. Used to fill the FP64 roofline denominator
// that MEASURED_PEAKS.json does not carry.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

template <int ILP>
__global__ void __launch_bounds__(256) dfma_kernel(double* out, int iters, double x, double y) {
    double acc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = threadIdx.x + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) acc[i] = fma(acc[i], x, y);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void __launch_bounds__(256) dmma_kernel(double* out, int iters, double x, double y) {
    double c[NACC][2];
#pragma unroll
    for (int i = 0; i < NACC; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; }
    double a = x + threadIdx.x, b = y;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// mixed: DMMA + DFMA interleaved, to see whether they share a pipe
template <int NACC>
__global__ void __launch_bounds__(256) mixed_kernel(double* out, int iters, double x, double y) {
    double c[NACC][2], f[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; f[i] = i; }
    double a = x + threadIdx.x, b = y;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) {
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
            f[i] = fma(f[i], x, y);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1] + f[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
float time_ms(F launch, int reps) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    launch(); launch(); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int sms = p.multiProcessorCount;
    double* out; CK(cudaMalloc(&out, sizeof(double) * sms * 8 * 256));
    const int iters = 20000;
    printf("{\"gpu\": \"%s\", \"sms\": %d", p.name, sms);
    for (int bps = 1; bps <= 8; bps *= 2) {
        int grid = sms * bps;
        float ms = time_ms([&] { dfma_kernel<8><<<grid, 256>>>(out, iters, 1.0000001, 1e-9); }, 5);
        double tf = 2.0 * 8 * iters * 256.0 * grid / (ms * 1e-3) / 1e12;
        printf(", \"dfma_tflops_bps%d\": %.2f", bps, tf);
    }
    for (int bps = 1; bps <= 8; bps *= 2) {
        int grid = sms * bps;
        float ms = time_ms([&] { dmma_kernel<8><<<grid, 256>>>(out, iters, 1.0000001, 1e-9); }, 5);
        double tf = 2.0 * 256 * 8 * iters * 8.0 * grid / (ms * 1e-3) / 1e12;   // 256 FMA per warp-MMA, 8 warps
        printf(", \"dmma_tflops_bps%d\": %.2f", bps, tf);
    }
    {
        int grid = sms * 4;
        float ms = time_ms([&] { dmma_kernel<2><<<grid, 256>>>(out, iters, 1.0000001, 1e-9); }, 5);
        double tf = 2.0 * 256 * 2 * iters * 8.0 * grid / (ms * 1e-3) / 1e12;
        printf(", \"dmma_tflops_acc2_bps4\": %.2f", tf);
        ms = time_ms([&] { dmma_kernel<1><<<sms, 32>>>(out, iters, 1.0000001, 1e-9); }, 5);
        printf(", \"dmma_dep_latency_ns\": %.2f", ms * 1e6 / iters);
        ms = time_ms([&] { dfma_kernel<1><<<sms, 32>>>(out, iters, 1.0000001, 1e-9); }, 5);
        printf(", \"dfma_dep_latency_ns\": %.2f", ms * 1e6 / iters);
    }
    {
        int grid = sms * 4;
        float ms = time_ms([&] { mixed_kernel<8><<<grid, 256>>>(out, iters, 1.0000001, 1e-9); }, 5);
        double tf_mma = 2.0 * 256 * 8 * iters * 8.0 * grid / (ms * 1e-3) / 1e12;
        double tf_fma = 2.0 * 8 * iters * 256.0 * grid / (ms * 1e-3) / 1e12;
        printf(", \"mixed_dmma_tflops\": %.2f, \"mixed_dfma_tflops\": %.2f", tf_mma, tf_fma);
    }
    printf("}\n");
    return 0;
}
