// Shared device helpers for the BranchedGP B200 evaluator (sm_100a only).
//
// Tile storage.  Every M x M matrix of the dense pipeline (Lc, R / R^-1, Pbar, Lbar / Ktilde) lives in HBM as
// 32x32 fp64 tiles (8 KB, contiguous) in packed block-lower order: tile (r,c), r>=c, at index r(r+1)/2+c.
// Inside a tile element (i,j) sits at  i*32 + (j ^ ((i&3)<<2)).  That XOR swizzle makes both fragment access
// patterns of mma.sync.m8n8k4.f64 (tile[o][k] and tile[k][o]) shared-memory bank-conflict free, and one tile is
// one cp.async.bulk (TMA, UBLKCP) transaction.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace bgp {

constexpr int TS = 32;                 // tile side
constexpr int TILE_ELEMS = TS * TS;    // 1024 doubles
constexpr int TILE_BYTES = TILE_ELEMS * 8;
constexpr int TPM = 4;                 // tiles per macro-block side
constexpr int MB = TS * TPM;           // macro block side (128)
constexpr int NTHREADS = 256;          // assignment passes, vector kernels, sparse path
#ifndef BGP_GEMM_WARPS
#define BGP_GEMM_WARPS 16
#endif
// Tile GEMM kernels: 4 warps along the 128 rows x (GEMM_WARPS/4) warps along the 128 columns.  16 warps (32x32 warp
// tiles, 4 warps per SM sub-partition) keep the DMMA pipe fed while one warp waits on LDS / mbarrier / instruction fetch.
constexpr int GEMM_WARPS = BGP_GEMM_WARPS;
constexpr int GEMM_THREADS = 32 * GEMM_WARPS;
constexpr int WCOLS = GEMM_WARPS / 4;          // warps along the columns (2 or 4)
constexpr int WN = MB / WCOLS;                 // columns per warp tile (64 or 32)
constexpr int NI = WN / 8;                     // 8-column MMA tiles per warp (8 or 4)
constexpr int NBT = WN / TS;                   // B tiles per warp (2 or 1)
constexpr int NSTAGES = 3;
constexpr int STAGE_ELEMS = 8 * TILE_ELEMS;          // 4 A tiles + 4 B tiles
constexpr int STAGE_BYTES = STAGE_ELEMS * 8;          // 64 KB
constexpr int PIPE_BYTES = NSTAGES * STAGE_BYTES;     // 192 KB
#ifndef BGP_DLD
#define BGP_DLD (MB + 4)
#endif
// leading dimension of the 128x128 shared-memory scratch of the diagonal-block routines: 132 = 4 mod 16 eight-byte banks makes both
// DMMA fragment patterns ([g][c] and [c][g]) conflict free (tools/microbench/block_kernels_test: chol128 + trinv128 99k cycles at
// 129, 72k at 132)
constexpr int DLD = BGP_DLD;
constexpr int SMEM_RED_DOUBLES = 3 * 512;
constexpr int SMEM_BYTES = PIPE_BYTES + SMEM_RED_DOUBLES * 8 + 64;   // + reduction scratch + 2*NSTAGES mbarriers

constexpr double JITTER = 1e-6;        // gpflow.default_jitter()
constexpr double SQ_A = 1.0 - 2e-6;    // assigngp_dense.py:162
constexpr double SQ_B = 1e-6;

__host__ __device__ __forceinline__ int tswz(int i, int j) { return i * TS + (j ^ ((i & 3) << 2)); }
__host__ __device__ __forceinline__ long long packed_tile(int r, int c) { return (long long)r * (r + 1) / 2 + c; }
__host__ __device__ __forceinline__ long long packed_tiles(int nT) { return (long long)nT * (nT + 1) / 2; }

// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred P;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, P;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
// One tile (or any 16B-multiple run) global -> shared through the TMA engine, completing on an mbarrier.
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// D(8x8) += A(8x4, row) * B(4x8, col).  lane = 4*g + c:  a = A[g][c],  b = B[c][g],  d = {D[g][2c], D[g][2c+1]}.
__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d"(d0), "+d"(d1)
        : "d"(a), "d"(b));
}

// ---------------------------------------------------------------- reductions
__device__ __forceinline__ double warp_sum(double x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    return x;
}
__device__ __forceinline__ double warp_max(double x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x = fmax(x, __shfl_xor_sync(0xffffffffu, x, o));
    return x;
}
// Sum over the whole CTA (blockDim.x multiple of 32, <= 1024).  `red` holds >= 32 doubles.  Result valid in all threads.
__device__ __forceinline__ double block_sum(double x, double* red) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    x = warp_sum(x);
    __syncthreads();
    if (lane == 0) red[w] = x;
    __syncthreads();
    double r = (lane < nw) ? red[lane] : 0.0;
    r = warp_sum(r);
    return r;
}

// ---------------------------------------------------------------- base kernels (branch_kernParamGPflow.py:117; gpflow Matern32 / SquaredExponential)
constexpr int KIND_MATERN32 = 0;
constexpr int KIND_SQEXP = 1;
#define BGP_SQRT3 1.7320508075688772

// k(d) for d = t - t'
__device__ __forceinline__ double kbase(double d, double ell, double var, int kind) {
    const double r = fabs(d) / ell;
    if (kind == KIND_SQEXP) return var * exp(-0.5 * r * r);
    const double s = BGP_SQRT3 * r;
    return var * (1.0 + s) * exp(-s);
}
// k, dk/d ell, dk/dd in one go
__device__ __forceinline__ void kbase_grads(double d, double ell, double var, int kind, double& k, double& dl, double& dd) {
    const double r = fabs(d) / ell;
    if (kind == KIND_SQEXP) {
        const double e = var * exp(-0.5 * r * r);
        k = e;
        dl = e * r * r / ell;
        dd = -e * d / (ell * ell);
    } else {
        const double s = BGP_SQRT3 * r;
        const double e = var * exp(-s);
        k = (1.0 + s) * e;
        dl = 3.0 * r * r * e / ell;
        dd = -3.0 * r * e * (d > 0.0 ? 1.0 : (d < 0.0 ? -1.0 : 0.0)) / ell;
    }
}

}  // namespace bgp
