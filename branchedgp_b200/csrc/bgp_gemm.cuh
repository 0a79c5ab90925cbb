// Per-CTA fp64 tile GEMM core:  C(128x128, registers) += Aop(128 x k) * Bop(128 x k)^T
//
// * operands are streamed from HBM/L2 as 32x32 swizzled tiles by cp.async.bulk (TMA) into a 3-stage shared-memory
//   ring, completion signalled on mbarriers (expect_tx);
// * 8 warps, warp tile 32 x 64, mma.sync.aligned.m8n8k4.f64 (DMMA) -- tcgen05 has no fp64 kind;
// * each logical operand is "outer x k" and is resolved tile by tile from a block-lower matrix in one of three
//   ways (ROWK: op[o][k] = Mat[o][k];  COLK: op[o][k] = Mat[k][o];  SYM: symmetric matrix stored lower), so every
//   heavy phase of the dense pipeline (potrf, L^T D L, trtri, lauum, trmm, Cholesky adjoint) is this one routine
//   with a different operand description, k-range and epilogue.
#pragma once
#include "bgp_common.cuh"

#ifndef BGP_KUNROLL
#define BGP_KUNROLL 2
#endif
#ifndef BGP_SWPIPE
#define BGP_SWPIPE 0
#endif

namespace bgp {

constexpr int KUNROLL = BGP_KUNROLL;
#ifndef BGP_PRODUCER_LAST
#define BGP_PRODUCER_LAST 1
#endif
constexpr int PRODUCER_TID = BGP_PRODUCER_LAST ? (GEMM_THREADS - 32) : 0;   // unroll factor of the k4 loop of warp_kstep

// Warp -> 32x32 warp tile of the 128x128 macro tile.  Warp w runs on SM sub-partition w & 3.  Every structural-zero skip
// of the dense pipeline is on the A (row) operand (triangular factors in syrk / trtri / lauum / trmm), so the four warps
// that share an A tile row are spread over the four sub-partitions: a skipped row tile then frees one warp's share of the
// DMMA pipe on each of them instead of idling one sub-partition while the others still run four warps.
#ifndef BGP_ROWS_ACROSS_SMSP
#define BGP_ROWS_ACROSS_SMSP 1
#endif
// rowmap = 1 (default): A tile rows across sub-partitions; rowmap = 0: B tile columns across sub-partitions, for the calls whose
// skipped operand is B (multiplications by an inverse diagonal block from the right).  gemm_macro and the acc_foreach that
// reads its accumulators must be given the same value.
__device__ __forceinline__ int warp_row_tile(int warp, int rowmap = 1) {
    return (BGP_ROWS_ACROSS_SMSP && WCOLS == 4 && rowmap) ? (warp >> 2) : (warp & 3);
}
__device__ __forceinline__ int warp_col_group(int warp, int rowmap = 1) {
    return (BGP_ROWS_ACROSS_SMSP && WCOLS == 4 && rowmap) ? (warp & 3) : (warp >> 2);
}

// rowmap = 2: diagonal macro tiles whose upper triangle is never used.  Only the 10 lower warp tiles are computed, dealt
// 3/3/2/2 over the sub-partitions (warp w sits on sub-partition w & 3), so such a tile costs 3/4 of a full one.
//   warp:      0      1      2      3      4      5      6      7      8      9     10..15
//   (wa, wb): (0,0)  (1,1)  (2,2)  (3,3)  (1,0)  (2,1)  (3,1)  (3,2)  (2,0)  (3,0)   idle
__device__ __forceinline__ void warp_tile(int warp, int rowmap, int& wa, int& wb, bool& idle) {
    idle = false;
    if (rowmap == 2 && WCOLS == 4) {
        const int a = (int)((0xFFFFFF3233213210ull >> (4 * warp)) & 15), b = (int)((0xFFFFFF0021103210ull >> (4 * warp)) & 15);
        idle = (a == 15);
        wa = idle ? 0 : a;
        wb = idle ? 0 : b;
    } else {
        wa = warp_row_tile(warp, rowmap);
        wb = warp_col_group(warp, rowmap);
    }
}

enum { ACC_ROWK = 0, ACC_COLK = 1, ACC_SYM = 2 };
enum { LAY_PACKED = 0, LAY_BLOCK = 1 };

struct Opnd {
    const double* base;   // tile storage
    const double* dalt;   // optional replacement for the diagonal tiles (packed layout): tile r at dalt + r*1024
    int layout;           // LAY_PACKED: global block-lower matrix;  LAY_BLOCK: 4x4 tile block, tile (r,c) at (r-r0)*4+(c-c0)
    int access;           // ACC_*
    int r0, c0;           // origin of a LAY_BLOCK operand, in tiles
    int lower;            // tiles with r < c are structural zeros (never loaded, MMAs skipped)
};

__device__ __forceinline__ Opnd opnd_packed(const double* base, int access, const double* dalt = nullptr) {
    Opnd o;
    o.base = base; o.dalt = dalt; o.layout = LAY_PACKED; o.access = access; o.r0 = 0; o.c0 = 0; o.lower = 1;
    return o;
}
__device__ __forceinline__ Opnd opnd_block(const double* base, int access, int r0, int c0, int lower) {
    Opnd o;
    o.base = base; o.dalt = nullptr; o.layout = LAY_BLOCK; o.access = access; o.r0 = r0; o.c0 = c0; o.lower = lower;
    return o;
}

// Tile that holds op[o-tile][k-tile]; `trans` tells whether the tile is stored [k][o] (true) or [o][k] (false).
__device__ __forceinline__ const double* resolve(const Opnd& op, int o, int kt, bool& trans) {
    int r, c;
    if (op.access == ACC_ROWK) { r = o; c = kt; trans = false; }
    else if (op.access == ACC_COLK) { r = kt; c = o; trans = true; }
    else if (o >= kt) { r = o; c = kt; trans = false; }
    else { r = kt; c = o; trans = true; }
    if (op.lower && r < c) return nullptr;
    if (op.layout == LAY_PACKED) {
        if (r == c && op.dalt != nullptr) return op.dalt + (size_t)r * TILE_ELEMS;
        return op.base + packed_tile(r, c) * TILE_ELEMS;
    }
    return op.base + (size_t)((r - op.r0) * TPM + (c - op.c0)) * TILE_ELEMS;
}

struct Pipe {
    double* stages;    // NSTAGES * STAGE_ELEMS doubles of shared memory
    uint64_t* full;    // NSTAGES mbarriers: TMA bytes of the stage have landed
    uint64_t* empty;   // NSTAGES mbarriers: all 8 warps are done reading the stage
    uint32_t it;       // running k-step counter; identical in every thread of the CTA
    int row_limit;     // A (= output) tile rows >= row_limit hold only identity padding: their warps skip the MMAs (accumulators stay 0)
#ifdef BGP_TIMING
    long long t_first, t_steady, t_empty, t_macro, n_calls, n_ksteps;   // cycle counters of the calling thread (tools/gemm_timers.py)
#endif
};

#ifdef BGP_TIMING
__device__ unsigned long long g_gemm_timers[8];
#define BGP_TIC() const long long tic_ = clock64()
#define BGP_TOC(field) pipe.field += clock64() - tic_
// warp 0 lane 0 reports the consumer-side waits, the producer thread the empty-barrier waits
__device__ __forceinline__ void pipe_report(const Pipe& p) {
    if (threadIdx.x == 0) {
        atomicAdd(&g_gemm_timers[0], (unsigned long long)p.t_first);
        atomicAdd(&g_gemm_timers[1], (unsigned long long)p.t_steady);
        atomicAdd(&g_gemm_timers[3], (unsigned long long)p.t_macro);
        atomicAdd(&g_gemm_timers[4], (unsigned long long)p.n_calls);
        atomicAdd(&g_gemm_timers[5], (unsigned long long)p.n_ksteps);
    }
    if (threadIdx.x == GEMM_THREADS - 32) atomicAdd(&g_gemm_timers[2], (unsigned long long)p.t_empty);
}
#else
#define BGP_TIC()
#define BGP_TOC(field)
__device__ __forceinline__ void pipe_report(const Pipe&) {}
#endif

__device__ __forceinline__ void pipe_init(Pipe& p, unsigned char* smem) {
    p.stages = reinterpret_cast<double*>(smem);
    p.full = reinterpret_cast<uint64_t*>(smem + PIPE_BYTES + SMEM_RED_DOUBLES * 8);
    p.empty = p.full + NSTAGES;
    p.it = 0;
    p.row_limit = 1 << 30;
#ifdef BGP_TIMING
    p.t_first = p.t_steady = p.t_empty = p.t_macro = p.n_calls = p.n_ksteps = 0;
#endif
    if (threadIdx.x == 0) {
        for (int s = 0; s < NSTAGES; ++s) { mbar_init(p.full + s, 1); mbar_init(p.empty + s, GEMM_WARPS); }
        fence_mbar_init();
    }
    __syncthreads();
}
__device__ __forceinline__ double* smem_red(unsigned char* smem) { return reinterpret_cast<double*>(smem + PIPE_BYTES); }

typedef double Acc[4][NI][2];

__device__ __forceinline__ void acc_zero(Acc& acc) {
#pragma unroll
    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
        for (int ni = 0; ni < NI; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
}

// One 32-wide k-step of the warp tile.  FULL: both 32-column halves of the warp's 64 columns are live (no predication);
// otherwise nmask bit h enables half h.  The k4 loop is only partially unrolled so that the body stays inside the L0
// instruction cache (ncu: a fully unrolled body cost ~10 % of issue slots in no_instruction stalls).
template <bool TA, bool TB, bool FULL>
__device__ __forceinline__ void warp_kstep(Acc& acc, const double* __restrict__ As, const double* __restrict__ Bs0,
                                           const double* __restrict__ Bs1, int nmask, const double* __restrict__ ksc,
                                           int g, int c) {
    const int g3 = (g & 3) << 2;
    const int c2 = c << 2;
    // lane-dependent parts of the swizzled addresses
    //   non-transposed  tile[8*o8+g][4*k4+c]  ->  (8*o8+g)*32 + (((4*k4) ^ g3) + c)
    //   transposed      tile[4*k4+c][8*o8+g]  ->  (4*k4+c)*32 + ((8*o8+g) ^ c2)
    const int nbase = g * TS + c;
    const int tbase = c * TS;
    int toff[4];
#pragma unroll
    for (int o8 = 0; o8 < 4; ++o8) toff[o8] = tbase + ((8 * o8 + g) ^ c2);
    const bool h0 = FULL || (nmask & 1), h1 = (NBT == 2) && (FULL || (nmask & 2));
    auto load = [&](int k4, double (&af)[4], double (&bf)[NI]) {
        const int kx = ((4 * k4) ^ g3);
#pragma unroll
        for (int mi = 0; mi < 4; ++mi) af[mi] = As[TA ? (toff[mi] + 4 * k4 * TS) : (nbase + 8 * mi * TS + kx)];
        if (ksc != nullptr) {
            const double s = ksc[4 * k4 + c];
#pragma unroll
            for (int mi = 0; mi < 4; ++mi) af[mi] *= s;
        }
#pragma unroll
        for (int ni = 0; ni < NI; ++ni) {
            const int o8 = ni & 3;
            const double* Bt = (ni >> 2) ? Bs1 : Bs0;
            const int off = TB ? (toff[o8] + 4 * k4 * TS) : (nbase + 8 * o8 * TS + kx);
            if (FULL) bf[ni] = Bt[off];
            else bf[ni] = ((ni >> 2) ? h1 : h0) ? Bt[off] : 0.0;
        }
    };
    auto mma = [&](const double (&af)[4], const double (&bf)[NI]) {
        if (h0) {
#pragma unroll
            for (int mi = 0; mi < 4; ++mi)
#pragma unroll
                for (int ni = 0; ni < 4; ++ni) dmma(acc[mi][ni][0], acc[mi][ni][1], af[mi], bf[ni]);
        }
        if (NBT == 2 && h1) {
#pragma unroll
            for (int mi = 0; mi < 4; ++mi)
#pragma unroll
                for (int ni = 4; ni < NI; ++ni) dmma(acc[mi][ni][0], acc[mi][ni][1], af[mi], bf[ni]);
        }
    };
#if BGP_SWPIPE
    // Experiment kept for reference (off): software pipelining over the back edge -- fragments of k4+1 / k4+2 requested before
    // the MMAs of k4 / k4+1.  Measured on B200: 8911 cycles per k-step against 8623 for the plain loop (tools/gemm_timers.py;
    // small spills, and the LDS wait was never the limiter), so the plain loop below is the shipped one.
    double a0[4], b0[NI], a1[4], b1[NI];
    load(0, a0, b0);
#pragma unroll 1
    for (int k4 = 0; k4 < 8; k4 += 2) {
        load(k4 + 1, a1, b1);
        mma(a0, b0);
        load((k4 + 2) & 7, a0, b0);   // the wrap on the last trip re-reads k4 = 0 (unused) instead of branching
        mma(a1, b1);
    }
#else
#pragma unroll KUNROLL
    for (int k4 = 0; k4 < 8; ++k4) {
        double af[4], bf[NI];
        load(k4, af, bf);
        mma(af, bf);
    }
#endif
}

template <bool FULL>
__device__ __forceinline__ void warp_kstep_dispatch(Acc& acc, bool ta, bool tb, const double* As, const double* Bs0, const double* Bs1,
                                                    int nmask, const double* ksc, int g, int c) {
    if (ta) { if (tb) warp_kstep<true, true, FULL>(acc, As, Bs0, Bs1, nmask, ksc, g, c); else warp_kstep<true, false, FULL>(acc, As, Bs0, Bs1, nmask, ksc, g, c); }
    else    { if (tb) warp_kstep<false, true, FULL>(acc, As, Bs0, Bs1, nmask, ksc, g, c); else warp_kstep<false, false, FULL>(acc, As, Bs0, Bs1, nmask, ksc, g, c); }
}

// acc += Aop[ao0*32 .. +128][kt0*32 .. kt1*32) * Bop[bo0*32 .. +128][same k]^T, optionally with the k index scaled
// by kscale[k].  Must be called by all GEMM_THREADS threads with identical arguments.
__device__ __forceinline__ void gemm_macro(Acc& acc, Pipe& pipe, const Opnd& A, int ao0, const Opnd& B, int bo0, int kt0,
                                           int kt1, const double* __restrict__ kscale, int rowmap = 1) {
    const int nk = kt1 - kt0;
    if (nk <= 0) return;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, c = lane & 3;
    int wa, wb;
    bool idle;
    warp_tile(warp, rowmap, wa, wb, idle);
    wb *= NBT;
    const bool row_pad = idle || (ao0 + wa >= pipe.row_limit);

    auto issue = [&](int ks) {
        const uint32_t it = pipe.it + (uint32_t)ks;
        const int st = (int)(it % NSTAGES);
        double* sb = pipe.stages + (size_t)st * STAGE_ELEMS;
        uint64_t* bar = pipe.full + st;
        const int kt = kt0 + ks;
        const double* src[8];
        uint32_t n = 0;
        bool tr;
#pragma unroll
        for (int a = 0; a < 4; ++a) { src[a] = resolve(A, ao0 + a, kt, tr); n += (src[a] != nullptr); }
#pragma unroll
        for (int b = 0; b < 4; ++b) { src[4 + b] = resolve(B, bo0 + b, kt, tr); n += (src[4 + b] != nullptr); }
        mbar_arrive_expect_tx(bar, n * (uint32_t)TILE_BYTES);
#pragma unroll
        for (int i = 0; i < 8; ++i)
            if (src[i] != nullptr) tma_load_1d(sb + (size_t)i * TILE_ELEMS, src[i], TILE_BYTES, bar);
    };

    // The producer is lane 0 of the LAST warp: the warp arbiter favours high warp ids, so that warp reaches each k-step
    // first and the refill is requested as early as the ring allows (with warp 0 as producer the loads were issued by the
    // slowest warp and every other warp then waited for the data: ncu long_scoreboard 17 % on the full-barrier spin).
    __syncthreads();   // every warp is done with the ring (previous GEMM or scratch use)
#ifdef BGP_TIMING
    const long long t_macro0 = clock64();
    pipe.n_calls += 1; pipe.n_ksteps += nk;
#endif
    if (tid == PRODUCER_TID) {
        const int pre = nk < (NSTAGES - 1) ? nk : (NSTAGES - 1);
        for (int p = 0; p < pre; ++p) {
            const uint32_t itn = pipe.it + (uint32_t)p;
            if (itn >= NSTAGES) mbar_wait(pipe.empty + itn % NSTAGES, ((itn / NSTAGES) - 1u) & 1u);
            issue(p);
        }
    }
    for (int ks = 0; ks < nk; ++ks) {
        if (ks + NSTAGES - 1 < nk) {
            if (tid == PRODUCER_TID) {   // refill the slot released NSTAGES-1 k-steps ago once all 8 warps have arrived on its `empty` barrier
                const uint32_t itn = pipe.it + (uint32_t)(ks + NSTAGES - 1);
                { BGP_TIC(); if (itn >= NSTAGES) mbar_wait(pipe.empty + itn % NSTAGES, ((itn / NSTAGES) - 1u) & 1u); BGP_TOC(t_empty); }
                issue(ks + NSTAGES - 1);
            }
        }
        const uint32_t it = pipe.it + (uint32_t)ks;
        const int st = (int)(it % NSTAGES);
#ifdef BGP_TIMING
        { BGP_TIC(); mbar_wait(pipe.full + st, (it / NSTAGES) & 1u); if (ks < NSTAGES - 1) { BGP_TOC(t_first); } else { BGP_TOC(t_steady); } }
#else
        mbar_wait(pipe.full + st, (it / NSTAGES) & 1u);
#endif
        const int kt = kt0 + ks;
        bool ta, tb0, tb1;
        const double* pa = resolve(A, ao0 + wa, kt, ta);
        const double* pb0 = resolve(B, bo0 + wb, kt, tb0);
        const double* pb1 = nullptr;
        tb1 = tb0;
        if (NBT == 2) pb1 = resolve(B, bo0 + wb + 1, kt, tb1);
        if (pa == nullptr || row_pad) { __syncwarp(); if (lane == 0) mbar_arrive(pipe.empty + st); continue; }
        int nmask = (pb0 != nullptr ? 1 : 0) | (pb1 != nullptr ? 2 : 0);
        if (nmask == 0) { __syncwarp(); if (lane == 0) mbar_arrive(pipe.empty + st); continue; }
        const double* sb = pipe.stages + (size_t)st * STAGE_ELEMS;
        const double* As = sb + (size_t)wa * TILE_ELEMS;
        const double* Bs0 = sb + (size_t)(4 + wb) * TILE_ELEMS;
        const double* Bs1 = Bs0 + TILE_ELEMS;
        const double* ksc = kscale ? kscale + (size_t)kt * TS : nullptr;
        if (nmask == (NBT == 2 ? 3 : 1) && tb0 == tb1) {
            warp_kstep_dispatch<true>(acc, ta, tb0, As, Bs0, Bs1, 3, ksc, g, c);
        } else {
            if (nmask & 1) warp_kstep_dispatch<false>(acc, ta, tb0, As, Bs0, Bs1, 1, ksc, g, c);
            if (nmask & 2) warp_kstep_dispatch<false>(acc, ta, tb1, As, Bs0, Bs1, 2, ksc, g, c);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(pipe.empty + st);
    }
    pipe.it += (uint32_t)nk;
#ifdef BGP_TIMING
    pipe.t_macro += clock64() - t_macro0;
#endif
}

// Visit every accumulator pair of the calling thread: f(tile_row, tile_col, i_in_tile, j_in_tile (even), v0, v1)
// for elements (i, j) and (i, j+1) of tile (4*I + ..., 4*J + ...).
template <class F>
__device__ __forceinline__ void acc_foreach(Acc& acc, int I, int J, F f, int rowmap = 1) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, c = lane & 3;
    int wa, wb;
    bool idle;
    warp_tile(warp, rowmap, wa, wb, idle);
    wb *= NBT;
    if (idle) return;
#pragma unroll
    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
        for (int ni = 0; ni < NI; ++ni)
            f(TPM * I + wa, TPM * J + wb + (ni >> 2), 8 * mi + g, 8 * (ni & 3) + 2 * c, acc[mi][ni][0], acc[mi][ni][1]);
}

__device__ __forceinline__ void tile_store2(double* tile, int i, int j, double v0, double v1) {
    *reinterpret_cast<double2*>(tile + tswz(i, j)) = make_double2(v0, v1);
}
__device__ __forceinline__ double2 tile_load2(const double* tile, int i, int j) {
    return *reinterpret_cast<const double2*>(tile + tswz(i, j));
}

}  // namespace bgp
