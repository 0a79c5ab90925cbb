// Device-resident batched L-BFGS for the AssignGP models: many (gene, branching point) work items are fitted at once, one
// workspace slot each, with logPhi, its gradient and the quasi-Newton history living in HBM for the whole fit.
//
// Replaces, for a batch of items, the per-item loop  gpflow.optimizers.Scipy().minimize(m.training_loss, m.trainable_variables,
// options=dict(maxiter=...))  of FitBranchingModel.py:120-132 (SciPy L-BFGS-B, unbounded, on the packed unconstrained
// vector [lengthscale, kernel variance, likelihood variance, logPhi]).  The optimiser restates the published algorithms
// SciPy wraps -- Byrd/Lu/Nocedal/Zhu L-BFGS-B 3.0 in its unconstrained branch (steepest descent first step of length
// 1/|g|, unit steps afterwards, curvature-skip rule, memory reset after a failed line search, PGTOL and FACTR*EPSMCH
// tests) and the More'-Thuente line search of MINPACK-2 (dcsrch / dcstep, ftol 1e-3, gtol 0.9, xtol 0.1, at most 20
// evaluations) -- with the quasi-Newton direction computed from inner products only:
//
//   * the 3N^2-long vectors never leave the GPU: one fused pass forms the inner products of the newest (s, y) pair and of
//     the gradient with the whole history (k_fit_gram), the host runs the two-loop recursion in that (2m+1)-dimensional
//     coefficient space per slot, and one fused pass forms the direction and the trial point (k_fit_dir);
//   * every round evaluates the bound + gradient of all slots with the same kernels as bgp_dense_elbo_grad /
//     bgp_sparse_elbo_grad; a slot whose item has converged is refilled with the next pending item, so the one-CTA-per-item
//     kernels stay full until the tail of the batch;
//   * the three hyper-parameters of a slot (softplus-transformed, with their Normal priors) are carried on the
//     host and join every inner product.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <limits>

#include "bgp_internal.h"

using namespace bgp;

namespace {

constexpr int FIT_PARTS = 16;      // CTAs per slot in the vector kernels
constexpr int FIT_CHUNK = 2048;    // elements staged per step of k_fit_gram
constexpr int FIT_MAXCOR = 16;
constexpr double EPSMCH = 2.220446049250313e-16;

// =============================================================================================== More'-Thuente line search
struct LineSearch {
    enum Task { FG = 0, CONVERGENCE = 1, WARNING = 2, ERROR = 3 };
    double ftol = 1e-3, gtol = 0.9, xtol = 0.1, stpmin = 0.0, stpmax = 1e10;
    bool brackt = false;
    int stage = 1;
    double ginit = 0, gtest = 0, gx = 0, gy = 0, finit = 0, fx = 0, fy = 0, stx = 0, sty = 0, stmin = 0, stmax = 0, width = 0, width1 = 0;

    static void cstep(double& stx, double& fx, double& dx, double& sty, double& fy, double& dy, double& stp, double fp, double dp,
                      bool& brackt, double stpmin, double stpmax) {
        const double sgnd = dp * (dx / std::fabs(dx));
        double stpf;
        if (fp > fx) {   // higher function value: the minimum is bracketed
            const double theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp;
            const double s = std::max(std::fabs(theta), std::max(std::fabs(dx), std::fabs(dp)));
            double gamma = s * std::sqrt((theta / s) * (theta / s) - (dx / s) * (dp / s));
            if (stp < stx) gamma = -gamma;
            const double p = (gamma - dx) + theta, q = ((gamma - dx) + gamma) + dp, r = p / q;
            const double stpc = stx + r * (stp - stx);
            const double stpq = stx + ((dx / ((fx - fp) / (stp - stx) + dx)) / 2.0) * (stp - stx);
            stpf = (std::fabs(stpc - stx) < std::fabs(stpq - stx)) ? stpc : stpc + (stpq - stpc) / 2.0;
            brackt = true;
        } else if (sgnd < 0.0) {   // lower value, derivatives of opposite sign: bracketed
            const double theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp;
            const double s = std::max(std::fabs(theta), std::max(std::fabs(dx), std::fabs(dp)));
            double gamma = s * std::sqrt((theta / s) * (theta / s) - (dx / s) * (dp / s));
            if (stp > stx) gamma = -gamma;
            const double p = (gamma - dp) + theta, q = ((gamma - dp) + gamma) + dx, r = p / q;
            const double stpc = stp + r * (stx - stp);
            const double stpq = stp + (dp / (dp - dx)) * (stx - stp);
            stpf = (std::fabs(stpc - stp) > std::fabs(stpq - stp)) ? stpc : stpq;
            brackt = true;
        } else if (std::fabs(dp) < std::fabs(dx)) {   // lower value, same sign, derivative magnitude decreases
            const double theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp;
            const double s = std::max(std::fabs(theta), std::max(std::fabs(dx), std::fabs(dp)));
            double gamma = s * std::sqrt(std::max(0.0, (theta / s) * (theta / s) - (dx / s) * (dp / s)));
            if (stp > stx) gamma = -gamma;
            const double p = (gamma - dp) + theta, q = (gamma + (dx - dp)) + gamma, r = p / q;
            double stpc;
            if (r < 0.0 && gamma != 0.0) stpc = stp + r * (stx - stp);
            else if (stp > stx) stpc = stpmax;
            else stpc = stpmin;
            const double stpq = stp + (dp / (dp - dx)) * (stx - stp);
            if (brackt) {
                stpf = (std::fabs(stpc - stp) < std::fabs(stpq - stp)) ? stpc : stpq;
                if (stp > stx) stpf = std::min(stp + 0.66 * (sty - stp), stpf);
                else stpf = std::max(stp + 0.66 * (sty - stp), stpf);
            } else {
                stpf = (std::fabs(stpc - stp) > std::fabs(stpq - stp)) ? stpc : stpq;
                stpf = std::min(stpmax, stpf);
                stpf = std::max(stpmin, stpf);
            }
        } else {   // lower value, same sign, derivative magnitude does not decrease
            if (brackt) {
                const double theta = 3.0 * (fp - fy) / (sty - stp) + dy + dp;
                const double s = std::max(std::fabs(theta), std::max(std::fabs(dy), std::fabs(dp)));
                double gamma = s * std::sqrt((theta / s) * (theta / s) - (dy / s) * (dp / s));
                if (stp > sty) gamma = -gamma;
                const double p = (gamma - dp) + theta, q = ((gamma - dp) + gamma) + dy, r = p / q;
                stpf = stp + r * (sty - stp);
            } else if (stp > stx) stpf = stpmax;
            else stpf = stpmin;
        }
        if (fp > fx) { sty = stp; fy = fp; dy = dp; }
        else {
            if (sgnd < 0.0) { sty = stx; fy = fx; dy = dx; }
            stx = stp; fx = fp; dx = dp;
        }
        stp = stpf;
    }

    // First call of a search: f, g at step 0 and the initial trial step.
    Task start(double& stp, double f, double g) {
        if (stp < stpmin || stp > stpmax || g >= 0.0) return ERROR;
        brackt = false; stage = 1; finit = f; ginit = g; gtest = ftol * ginit;
        width = stpmax - stpmin; width1 = width / 0.5;
        stx = 0.0; fx = finit; gx = ginit; sty = 0.0; fy = finit; gy = ginit;
        stmin = 0.0; stmax = stp + 4.0 * stp;
        return FG;
    }
    // f, g at the trial step `stp`; returns FG with the next trial step in `stp`, or CONVERGENCE / WARNING (stp stays).
    Task next(double& stp, double f, double g) {
        const double ftest = finit + stp * gtest;
        if (stage == 1 && f <= ftest && g >= 0.0) stage = 2;
        Task task = FG;
        if (brackt && (stp <= stmin || stp >= stmax)) task = WARNING;            // rounding errors prevent progress
        if (brackt && stmax - stmin <= xtol * stmax) task = WARNING;             // xtol test satisfied
        if (stp == stpmax && f <= ftest && g <= gtest) task = WARNING;           // stp = stpmax
        if (stp == stpmin && (f > ftest || g >= gtest)) task = WARNING;          // stp = stpmin
        if (f <= ftest && std::fabs(g) <= gtol * (-ginit)) task = CONVERGENCE;
        if (task != FG) return task;
        if (stage == 1 && f <= fx && f > ftest) {
            const double fm = f - stp * gtest, gm = g - gtest;
            double fxm = fx - stx * gtest, fym = fy - sty * gtest, gxm = gx - gtest, gym = gy - gtest;
            cstep(stx, fxm, gxm, sty, fym, gym, stp, fm, gm, brackt, stmin, stmax);
            fx = fxm + stx * gtest; fy = fym + sty * gtest; gx = gxm + gtest; gy = gym + gtest;
        } else {
            cstep(stx, fx, gx, sty, fy, gy, stp, f, g, brackt, stmin, stmax);
        }
        if (brackt) {
            if (std::fabs(sty - stx) >= 0.66 * width1) stp = stx + 0.5 * (sty - stx);
            width1 = width;
            width = std::fabs(sty - stx);
        }
        if (brackt) { stmin = std::min(stx, sty); stmax = std::max(stx, sty); }
        else { stmin = stp + 1.1 * (stp - stx); stmax = stp + 4.0 * (stp - stx); }
        stp = std::max(stp, stpmin);
        stp = std::min(stp, stpmax);
        if ((brackt && (stp <= stmin || stp >= stmax)) || (brackt && stmax - stmin <= xtol * stmax)) stp = stx;
        return FG;
    }
};

// =============================================================================================== limited-memory direction
// History of one slot in coefficient space.  Ring position p holds pair p; `order` lists the stored positions oldest first.
// Inner products are over the FULL variable vector (logPhi part from the device + hyper part from the host).
struct History {
    int m = 10, col = 0;
    int order[FIT_MAXCOR];
    double SS[FIT_MAXCOR][FIT_MAXCOR], SY[FIT_MAXCOR][FIT_MAXCOR], YY[FIT_MAXCOR][FIT_MAXCOR];   // SY[i][j] = s_i . y_j
    double su[FIT_MAXCOR][3], yu[FIT_MAXCOR][3];   // hyper-parameter components of the pairs
    void reset() { col = 0; }
    int next_pos() const {   // where the next pair goes (overwrites the oldest when full)
        if (col < m) {
            bool used[FIT_MAXCOR] = {false};
            for (int i = 0; i < col; ++i) used[order[i]] = true;
            for (int p = 0; p < m; ++p) if (!used[p]) return p;
        }
        return order[0];
    }
    void push(int p) {
        if (col == m) { for (int i = 1; i < col; ++i) order[i - 1] = order[i]; order[col - 1] = p; }
        else order[col++] = p;
    }
};

// d = -H g by the two-loop recursion carried out on coefficients over the basis {g, s_p, y_p}.
//   bS[p] = s_p . g, bY[p] = y_p . g, gg = g . g.  Output: d = cg * g + sum_p cs[p] s_p + cy[p] y_p, and gd = g . d.
void lbfgs_direction(const History& H, const double* bS, const double* bY, double gg, double& cg, double* cs, double* cy, double& gd) {
    const int m = H.m;
    for (int p = 0; p < m; ++p) cs[p] = cy[p] = 0.0;
    cg = 1.0;   // q = g
    double alpha[FIT_MAXCOR];
    auto dot_s = [&](int i) {   // s_i . q
        double v = cg * bS[i];
        for (int k = 0; k < H.col; ++k) { const int p = H.order[k]; v += cs[p] * H.SS[i][p] + cy[p] * H.SY[i][p]; }
        return v;
    };
    auto dot_y = [&](int i) {   // y_i . q
        double v = cg * bY[i];
        for (int k = 0; k < H.col; ++k) { const int p = H.order[k]; v += cs[p] * H.SY[p][i] + cy[p] * H.YY[i][p]; }
        return v;
    };
    for (int k = H.col - 1; k >= 0; --k) {
        const int i = H.order[k];
        alpha[i] = dot_s(i) / H.SY[i][i];
        cy[i] -= alpha[i];
    }
    if (H.col > 0) {
        const int nw = H.order[H.col - 1];
        const double gamma = H.SY[nw][nw] / H.YY[nw][nw];
        cg *= gamma;
        for (int p = 0; p < m; ++p) { cs[p] *= gamma; cy[p] *= gamma; }
    }
    for (int k = 0; k < H.col; ++k) {
        const int i = H.order[k];
        const double beta = dot_y(i) / H.SY[i][i];
        cs[i] += alpha[i] - beta;
    }
    cg = -cg;
    for (int p = 0; p < m; ++p) { cs[p] = -cs[p]; cy[p] = -cy[p]; }
    gd = cg * gg;
    for (int k = 0; k < H.col; ++k) { const int p = H.order[k]; gd += cs[p] * bS[p] + cy[p] * bY[p]; }
}

// =============================================================================================== device kernels
// All vector kernels run grid (FIT_PARTS, slots) x 256 threads over per-slot vectors of n = N*3N doubles stored as
// [slot][n] arrays (history: [slot][m][n]).  G holds the gradient of the BOUND (the objective is its negative): signs are
// applied on the host except where a difference of gradients is stored (y = g_new - g_old = G0 - G).
struct FitVecs {
    long long n;
    int m;
    double *X, *G, *X0, *G0, *Dv, *Sh, *Yh;
};

// InitialiseVariationalPhi (assigngp_dense.py:92-128): off-block -9, own block log([1e-9, phi0 - 1e-9, phi1 - 1e-9]) after the
// branching point, log([1 - 2e-9, 1e-9, 1e-9]) on the trunk (t <= b).
__global__ void __launch_bounds__(256) k_fit_init_logphi(double* X, long long n, int slot, int N, const double* __restrict__ t,
                                                         const double* __restrict__ phiInitial, double b) {
    double* x = X + (size_t)slot * n;
    const int M = 3 * N;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(e / M), j = (int)(e - (long long)i * M);
        double v = -9.0;
        const int f = j - 3 * i;
        if (f >= 0 && f < 3) {
            const double eps = 1e-9;
            if (t[i] > b) v = log(f == 0 ? eps : phiInitial[2 * i + f - 1] - eps);
            else v = log(f == 0 ? 1.0 - 2.0 * eps : eps);
        }
        x[e] = v;
    }
}

// After an evaluation: gd = G . D, gg = G . G, gmax = max |G| per slot -> part[slot][FIT_PARTS][3]
__global__ void __launch_bounds__(256) k_fit_gd(FitVecs V, double* __restrict__ part) {
    __shared__ double red[32];
    const int slot = blockIdx.y;
    const double* g = V.G + (size_t)slot * V.n;
    const double* d = V.Dv + (size_t)slot * V.n;
    double s0 = 0.0, s1 = 0.0, mx = 0.0;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < V.n; e += (long long)gridDim.x * blockDim.x) {
        const double gv = g[e];
        s0 += gv * d[e];
        s1 += gv * gv;
        mx = fmax(mx, fabs(gv));
    }
    s0 = block_sum(s0, red);
    s1 = block_sum(s1, red);
    mx = warp_max(mx);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
        double r = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) r = fmax(r, red[w]);
        double* o = part + ((size_t)slot * FIT_PARTS + blockIdx.x) * 3;
        o[0] = s0; o[1] = s1; o[2] = r;
    }
}

// Accepted step of a slot: act 0 nothing; 1: X0 <- X, G0 <- G; 2: additionally s = X - X0 and y = G0 - G into ring position pos;
// 3: restart from the stored X0 / G0 after a failed line search (nothing to copy, only a new direction).
__global__ void __launch_bounds__(256) k_fit_accept(FitVecs V, const int* __restrict__ act, const int* __restrict__ pos) {
    const int slot = blockIdx.y, a = act[slot];
    if (a == 0 || a == 3) return;
    const size_t o = (size_t)slot * V.n;
    const double* x = V.X + o;
    const double* g = V.G + o;
    double* x0 = V.X0 + o;
    double* g0 = V.G0 + o;
    double* sp = nullptr;
    double* yp = nullptr;
    if (a == 2) {
        sp = V.Sh + ((size_t)slot * V.m + pos[slot]) * V.n;
        yp = V.Yh + ((size_t)slot * V.m + pos[slot]) * V.n;
    }
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < V.n; e += (long long)gridDim.x * blockDim.x) {
        const double xv = x[e], gv = g[e];
        if (a == 2) { sp[e] = xv - x0[e]; yp[e] = g0[e] - gv; }
        x0[e] = xv;
        g0[e] = gv;
    }
}

// Inner products of {s_new, y_new, G0} with the whole history: part[slot][FIT_PARTS][3][2m+1]
//   column j < m: S_j;  m <= j < 2m: Y_{j-m};  2m: G0.   Row 0: with s_new, 1: with y_new, 2: with G0.
// valid[slot] = bit mask of stored ring positions, newpos[slot] = position of the newest pair or -1 (rows 0/1 skipped).
__global__ void __launch_bounds__(256) k_fit_gram(FitVecs V, const int* __restrict__ act, const int* __restrict__ valid,
                                                  const int* __restrict__ newpos, double* __restrict__ part) {
    __shared__ double sh[3][FIT_CHUNK];
    const int slot = blockIdx.y;
    if (act[slot] == 0) return;
    const int m = V.m, W = 2 * m + 1;
    const int mask = valid[slot], np = newpos[slot];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t o = (size_t)slot * V.n;
    const double* g0 = V.G0 + o;
    const double* sn = np >= 0 ? V.Sh + ((size_t)slot * m + np) * V.n : nullptr;
    const double* yn = np >= 0 ? V.Yh + ((size_t)slot * m + np) * V.n : nullptr;
    // the columns this warp owns: j = warp, warp + 8, ... over the 2m history vectors; the last warp also owns column 2m
    double acc[3][(2 * FIT_MAXCOR + 8) / 8 + 1];
    constexpr int NJ = (2 * FIT_MAXCOR + 8) / 8 + 1;
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int q = 0; q < NJ; ++q) acc[r][q] = 0.0;
    const long long nchunks = (V.n + FIT_CHUNK - 1) / FIT_CHUNK;
    for (long long c = blockIdx.x; c < nchunks; c += gridDim.x) {
        const long long e0 = c * FIT_CHUNK;
        const int len = (int)((V.n - e0 < FIT_CHUNK) ? (V.n - e0) : FIT_CHUNK);
        __syncthreads();
        for (int k = threadIdx.x; k < FIT_CHUNK; k += blockDim.x) {
            const bool in = k < len;
            sh[0][k] = (in && sn) ? sn[e0 + k] : 0.0;
            sh[1][k] = (in && yn) ? yn[e0 + k] : 0.0;
            sh[2][k] = in ? g0[e0 + k] : 0.0;
        }
        __syncthreads();
#pragma unroll
        for (int q = 0; q < NJ; ++q) {
            const int j = warp + 8 * q;
            if (j > 2 * m) break;
            const double* bj;
            if (j == 2 * m) bj = g0;
            else {
                const int p = j < m ? j : j - m;
                if (!((mask >> p) & 1)) continue;
                bj = (j < m ? V.Sh : V.Yh) + ((size_t)slot * m + p) * V.n;
            }
            double a0 = 0.0, a1 = 0.0, a2 = 0.0;
            for (int k = lane; k < len; k += 32) {
                const double bv = bj[e0 + k];
                a0 += bv * sh[0][k];
                a1 += bv * sh[1][k];
                a2 += bv * sh[2][k];
            }
            acc[0][q] += a0; acc[1][q] += a1; acc[2][q] += a2;
        }
    }
#pragma unroll
    for (int q = 0; q < NJ; ++q) {
        const int j = warp + 8 * q;
        if (j > 2 * m) break;
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const double v = warp_sum(acc[r][q]);
            if (lane == 0) part[(((size_t)slot * FIT_PARTS + blockIdx.x) * 3 + r) * W + j] = v;
        }
    }
}

// New direction and first trial point of the slots with act != 0:
//   D = cg * G0 + sum_p cs[p] S_p + cy[p] Y_p  (coef[slot][2m+1] = cs, cy, cg);   X = X0 + stp * D.
__global__ void __launch_bounds__(256) k_fit_dir(FitVecs V, const int* __restrict__ act, const int* __restrict__ valid,
                                                 const double* __restrict__ coef, const double* __restrict__ stp) {
    const int slot = blockIdx.y;
    if (act[slot] == 0) return;
    const int m = V.m, W = 2 * m + 1;
    const int mask = valid[slot];
    const double* cf = coef + (size_t)slot * W;
    const double cg = cf[2 * m], st = stp[slot];
    const size_t o = (size_t)slot * V.n;
    const double* g0 = V.G0 + o;
    const double* x0 = V.X0 + o;
    double* d = V.Dv + o;
    double* x = V.X + o;
    const double* Sb = V.Sh + (size_t)slot * m * V.n;
    const double* Yb = V.Yh + (size_t)slot * m * V.n;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < V.n; e += (long long)gridDim.x * blockDim.x) {
        double v = cg * g0[e];
        for (int p = 0; p < m; ++p)
            if ((mask >> p) & 1) v += cf[p] * Sb[(size_t)p * V.n + e] + cf[m + p] * Yb[(size_t)p * V.n + e];
        d[e] = v;
        x[e] = x0[e] + st * v;
    }
}

// Next trial point of a running line search (act != 0): X = X0 + stp * D
__global__ void __launch_bounds__(256) k_fit_step(FitVecs V, const int* __restrict__ act, const double* __restrict__ stp) {
    const int slot = blockIdx.y;
    if (act[slot] == 0) return;
    const double st = stp[slot];
    const size_t o = (size_t)slot * V.n;
    const double* x0 = V.X0 + o;
    const double* d = V.Dv + o;
    double* x = V.X + o;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < V.n; e += (long long)gridDim.x * blockDim.x)
        x[e] = (st == 0.0) ? x0[e] : x0[e] + st * d[e];
}

// =============================================================================================== hyper-parameter transforms
// GPflow positive(): softplus, with lower bound 1e-6 for the Gaussian likelihood variance.
inline double softplus(double u) { return u > 0 ? u + std::log1p(std::exp(-u)) : std::log1p(std::exp(u)); }
inline double softplus_inv(double z) { return z + std::log(-std::expm1(-z)); }
inline double sigmoid(double u) { return 1.0 / (1.0 + std::exp(-u)); }
const double HYP_LOWER[3] = {0.0, 0.0, 1e-6};

enum Phase { PH_IDLE = 0, PH_INITIAL, PH_SEARCH, PH_FINAL, };

struct Slot {
    int item = -1;
    Phase phase = PH_IDLE;
    int iter = 0, nfev = 0, ifun = 0;
    double u[3], u0[3], du[3], gu[3], gu0[3];   // unconstrained hypers: trial, start of the search, direction, gradients
    double f = 0, fold = 0, gd = 0, gdold = 0, stp = 0, sbgnrm = 0;
    int result = 0;
    LineSearch ls;
    History H;
};

struct Evaluator {
    virtual int eval(int nslots, const double* X, double* G, cudaStream_t st) = 0;   // fills elbo / grad_hyp / status device arrays
    virtual ~Evaluator() {}
};

}  // namespace

// result codes written to status[] by the fit entry points
enum {
    FIT_CONV_PGTOL = 0,        // max |g| <= gtol
    FIT_CONV_FTOL = 1,         // relative reduction of f <= ftol
    FIT_STOP_MAXITER = 2,
    FIT_STOP_MAXFUN = 3,
    FIT_ABNORMAL_LNSRCH = 4,   // line search failed with an empty memory
    FIT_CHOLESKY_FAILED = 5    // a factorisation met a non-positive pivot (the reference raises; FitModel returns NaNs)
};

struct FitShared {
    bgp_handle* h;
    int count, N, D;
    long long n;
    const bgp_fit_options* opt;
    // host inputs / outputs
    const int* item_gene; const double* b; const double* hyp0;
    double *objective, *elbo_out, *hyp_out, *phi_out, *pred_mean, *pred_var, *logPhi_out;
    int *iters, *nfev, *status;
    int T; int gh_stride;
    int want_predict;
};

// Generic driver.  `ev` evaluates all slots; `finish_slots(slots, items)` is called once per round with the slots whose last
// evaluation was at their final point (gather Phi, predict from the slots' workspaces, copy the results out).
template <class InitX, class FinishSlot>
static int fit_core(FitShared& F, Evaluator& ev, int S, FitVecs V, cudaStream_t st, double* d_hyp, double* d_b, int* d_gene,
                    double* d_elbo, double* d_gh, int* d_status, int b_per_item, InitX init_x, FinishSlot finish_slots) {
    bgp_handle* h = F.h;
    const bgp_fit_options& O = *F.opt;
    const int m = V.m, W = 2 * m + 1;
    const double factr_eps = O.ftol;   // SciPy: factr = ftol / eps, test against factr * epsmch
    // small device / pinned buffers
    double *d_part1 = nullptr, *d_partG = nullptr, *d_coef = nullptr, *d_stp = nullptr;
    int *d_act = nullptr, *d_pos = nullptr, *d_valid = nullptr, *d_newpos = nullptr;
    // freed on every way out of this function, including the early returns of CU() inside the round loop
    struct DevGuard { std::vector<void*> p; ~DevGuard() { for (void* q : p) cudaFree(q); } } guard;
    auto dmalloc = [&](void** q, size_t bytes) -> cudaError_t { cudaError_t e = cudaMalloc(q, bytes); if (e == cudaSuccess) guard.p.push_back(*q); return e; };
    CU(dmalloc((void**)&d_part1, (size_t)S * FIT_PARTS * 3 * 8));
    CU(dmalloc((void**)&d_partG, (size_t)S * FIT_PARTS * 3 * W * 8));
    CU(dmalloc((void**)&d_coef, (size_t)S * W * 8));
    CU(dmalloc((void**)&d_stp, (size_t)S * 8));
    CU(dmalloc((void**)&d_act, (size_t)S * 4 * 4));
    d_pos = d_act + S; d_valid = d_pos + S; d_newpos = d_valid + S;
    std::vector<double> part1((size_t)S * FIT_PARTS * 3), partG((size_t)S * FIT_PARTS * 3 * W), coef((size_t)S * W), stpv(S);
    std::vector<double> hyp_c((size_t)S * 3), bv((size_t)S * b_per_item), elbo(S), gh((size_t)S * F.gh_stride);
    std::vector<int> gene(S), status(S), ints((size_t)S * 4);
    int* act = ints.data(); int* pos = act + S; int* valid = pos + S; int* newpos = valid + S;
    std::vector<Slot> slots(S);
    for (auto& s : slots) { s.H.m = m; s.ls = LineSearch(); }
    int next_item = 0, done = 0;
    const dim3 grid(FIT_PARTS, S);
    auto cleanup = [&]() {};   // the guard above owns the buffers

    auto constrained = [&](const Slot& s, double* out) { for (int k = 0; k < 3; ++k) out[k] = softplus(s.u[k]) + HYP_LOWER[k]; };
    auto log_prior = [&](const double* hc) {
        double lp = 0.0;
        if (O.train_hyp)
            for (int k = 0; k < 3; ++k)
                if (O.has_prior[k]) {
                    const double z = (hc[k] - O.prior_mean[k]) / O.prior_sd[k];
                    lp += -0.5 * z * z - std::log(O.prior_sd[k]) - 0.5 * std::log(2.0 * M_PI);
                }
        return lp;
    };
    auto load_item = [&](int sidx) -> int {
        Slot& s = slots[sidx];
        s = Slot();
        s.H.m = m;
        if (next_item >= F.count) { s.phase = PH_IDLE; return BGP_OK; }
        s.item = next_item++;
        s.phase = PH_INITIAL;
        for (int k = 0; k < 3; ++k) s.u[k] = softplus_inv(F.hyp0[(size_t)s.item * 3 + k] - HYP_LOWER[k]);
        gene[sidx] = F.item_gene[s.item];
        for (int k = 0; k < b_per_item; ++k) bv[(size_t)sidx * b_per_item + k] = F.b[(size_t)s.item * b_per_item + k];
        return init_x(sidx, s.item);
    };
    for (int sidx = 0; sidx < S; ++sidx) {
        gene[sidx] = 0;
        for (int k = 0; k < b_per_item; ++k) bv[(size_t)sidx * b_per_item + k] = 0.5;
        int rc = load_item(sidx);
        if (rc != BGP_OK) { cleanup(); return rc; }
    }

    std::vector<int> fin_slots, fin_items, refill;   // slots that finished in this round (outputs pending / to be refilled)
    auto finish = [&](int sidx, int result) -> int {
        Slot& s = slots[sidx];
        const int it = s.item;
        double hc[3];
        constrained(s, hc);
        F.status[it] = result;
        F.iters[it] = s.iter;
        F.nfev[it] = s.nfev;
        if (result == FIT_CHOLESKY_FAILED) {
            const double nan = std::numeric_limits<double>::quiet_NaN();
            F.objective[it] = nan; F.elbo_out[it] = nan;
            for (int k = 0; k < 3; ++k) F.hyp_out[(size_t)it * 3 + k] = hc[k];
        } else {
            const double lp = log_prior(hc);
            F.objective[it] = -s.f;          // log posterior density = bound + log prior
            F.elbo_out[it] = -s.f - lp;
            for (int k = 0; k < 3; ++k) F.hyp_out[(size_t)it * 3 + k] = hc[k];
            fin_slots.push_back(sidx);
            fin_items.push_back(it);
        }
        ++done;
        refill.push_back(sidx);
        s.phase = PH_IDLE;
        return BGP_OK;
    };

    long long rounds = 0;
    double s_eval = 0.0, s_opt = 0.0;   // wall seconds in evaluation rounds / optimiser passes (BGP_FIT_STATS=1 prints them)
    auto now = []() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    while (done < F.count) {
        ++rounds;
        const double t_round0 = now();
        // ---- evaluate every slot at its trial point
        for (int sidx = 0; sidx < S; ++sidx) {
            const Slot& s = slots[sidx];
            if (s.phase == PH_IDLE) { for (int k = 0; k < 3; ++k) hyp_c[(size_t)sidx * 3 + k] = F.hyp0[k]; }
            else constrained(s, &hyp_c[(size_t)sidx * 3]);
        }
        CU(cudaMemcpyAsync(d_hyp, hyp_c.data(), (size_t)S * 24, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(d_b, bv.data(), (size_t)S * b_per_item * 8, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(d_gene, gene.data(), (size_t)S * 4, cudaMemcpyHostToDevice, st));
        int rc = ev.eval(S, V.X, V.G, st);
        if (rc != BGP_OK) { cleanup(); return rc; }
        k_fit_gd<<<grid, 256, 0, st>>>(V, d_part1);
        CU(cudaMemcpyAsync(part1.data(), d_part1, part1.size() * 8, cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(elbo.data(), d_elbo, (size_t)S * 8, cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(gh.data(), d_gh, gh.size() * 8, cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(status.data(), d_status, (size_t)S * 4, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        const double t_round1 = now();
        s_eval += t_round1 - t_round0;

        // ---- per-slot decisions
        bool any_accept = false;
        for (int sidx = 0; sidx < S; ++sidx) { act[sidx] = 0; pos[sidx] = 0; valid[sidx] = 0; newpos[sidx] = -1; }
        std::vector<int> step_act(S, 0);
        for (int sidx = 0; sidx < S; ++sidx) {
            Slot& s = slots[sidx];
            if (s.phase == PH_IDLE) continue;
            s.nfev++;
            if (status[sidx] != 0) {
                rc = finish(sidx, FIT_CHOLESKY_FAILED);
                if (rc != BGP_OK) { cleanup(); return rc; }
                continue;
            }
            double hc[3];
            constrained(s, hc);
            double gdp = 0.0, ggp = 0.0, gmx = 0.0;
            for (int p = 0; p < FIT_PARTS; ++p) {
                const double* o = &part1[((size_t)sidx * FIT_PARTS + p) * 3];
                gdp += o[0]; ggp += o[1]; gmx = std::max(gmx, o[2]);
            }
            s.f = -(elbo[sidx] + log_prior(hc));
            for (int k = 0; k < 3; ++k) {
                if (!O.train_hyp) { s.gu[k] = 0.0; continue; }
                double g = gh[(size_t)sidx * F.gh_stride + k];
                if (O.has_prior[k]) g += -(hc[k] - O.prior_mean[k]) / (O.prior_sd[k] * O.prior_sd[k]);
                s.gu[k] = -g * sigmoid(s.u[k]);
            }
            double gnorm_inf = gmx, gg = ggp;
            for (int k = 0; k < 3; ++k) { gnorm_inf = std::max(gnorm_inf, std::fabs(s.gu[k])); gg += s.gu[k] * s.gu[k]; }

            if (s.phase == PH_FINAL) {   // re-evaluation at the restored point after an abnormal line search
                rc = finish(sidx, s.result);
                if (rc != BGP_OK) { cleanup(); return rc; }
                continue;
            }
            bool new_iteration = false;   // start a search from the point just evaluated
            if (s.phase == PH_INITIAL) {
                s.sbgnrm = gnorm_inf;
                if (s.sbgnrm <= O.gtol) { rc = finish(sidx, FIT_CONV_PGTOL); if (rc != BGP_OK) { cleanup(); return rc; } continue; }
                new_iteration = true;
            } else {   // PH_SEARCH
                s.gd = -gdp;
                for (int k = 0; k < 3; ++k) s.gd += s.gu[k] * s.du[k];
                double stp = s.stp;
                const LineSearch::Task task = s.ls.next(stp, s.f, s.gd);
                if (task == LineSearch::FG) {
                    s.ifun++;
                    if (s.ifun - 1 >= O.maxls) {
                        // restore the previous iterate; with an empty memory this is an abnormal termination, otherwise the
                        // memory is dropped and the iteration restarts along steepest descent
                        for (int k = 0; k < 3; ++k) { s.u[k] = s.u0[k]; s.gu[k] = s.gu0[k]; }
                        s.f = s.fold;
                        if (s.H.col == 0) {   // re-evaluate at the restored point so that the slot's workspace matches it, then finish
                            s.result = FIT_ABNORMAL_LNSRCH; s.phase = PH_FINAL;
                            stpv[sidx] = 0.0; step_act[sidx] = 1;
                        } else {              // X0 / G0 still hold the previous iterate: new direction from them, no copy
                            s.H.reset();
                            act[sidx] = 3; valid[sidx] = 0; any_accept = true;
                        }
                        continue;
                    }
                    s.stp = stp;
                    for (int k = 0; k < 3; ++k) s.u[k] = s.u0[k] + stp * s.du[k];
                    stpv[sidx] = stp; step_act[sidx] = 1;
                    continue;
                }
                // line search done: new iterate
                s.iter++;
                s.sbgnrm = gnorm_inf;
                if (s.iter >= O.maxiter) { rc = finish(sidx, FIT_STOP_MAXITER); if (rc != BGP_OK) { cleanup(); return rc; } continue; }
                if (s.nfev > O.maxfun) { rc = finish(sidx, FIT_STOP_MAXFUN); if (rc != BGP_OK) { cleanup(); return rc; } continue; }
                if (s.sbgnrm <= O.gtol) { rc = finish(sidx, FIT_CONV_PGTOL); if (rc != BGP_OK) { cleanup(); return rc; } continue; }
                const double ddum = std::max(std::fabs(s.fold), std::max(std::fabs(s.f), 1.0));
                if (s.fold - s.f <= factr_eps * ddum) { rc = finish(sidx, FIT_CONV_FTOL); if (rc != BGP_OK) { cleanup(); return rc; } continue; }
                // curvature test for the update (dr = y.s through the directional derivatives)
                const double dr = (s.gd - s.gdold) * s.stp, dd = -s.gdold * s.stp;
                if (dr > EPSMCH * dd) {
                    const int p = s.H.next_pos();
                    for (int k = 0; k < 3; ++k) { s.H.su[p][k] = s.u[k] - s.u0[k]; s.H.yu[p][k] = s.gu[k] - s.gu0[k]; }
                    s.H.push(p);
                    act[sidx] = 2; pos[sidx] = p; newpos[sidx] = p;
                }
                new_iteration = true;
            }
            if (new_iteration) {
                if (act[sidx] == 0) act[sidx] = 1;
                int mask = 0;
                for (int k = 0; k < s.H.col; ++k) mask |= 1 << s.H.order[k];
                valid[sidx] = mask;
                any_accept = true;
            }
        }
        // ---- outputs of the slots that finished (their last evaluation was at the final point), then refill them
        if (!fin_slots.empty()) {
            rc = finish_slots(fin_slots, fin_items);
            if (rc != BGP_OK) { cleanup(); return rc; }
            fin_slots.clear(); fin_items.clear();
        }
        for (int sidx : refill) {
            rc = load_item(sidx);
            if (rc != BGP_OK) { cleanup(); return rc; }
        }
        refill.clear();
        if (done >= F.count) break;

        // ---- history update + inner products for the slots that start a new iteration
        if (any_accept) {
            CU(cudaMemcpyAsync(d_act, act, (size_t)S * 16, cudaMemcpyHostToDevice, st));
            k_fit_accept<<<grid, 256, 0, st>>>(V, d_act, d_pos);
            k_fit_gram<<<grid, 256, 0, st>>>(V, d_act, d_valid, d_newpos, d_partG);
            CU(cudaMemcpyAsync(partG.data(), d_partG, partG.size() * 8, cudaMemcpyDeviceToHost, st));
            CU(cudaStreamSynchronize(st));
            for (int sidx = 0; sidx < S; ++sidx) {
                if (act[sidx] == 0) continue;
                Slot& s = slots[sidx];
                History& H = s.H;
                double row[3][2 * FIT_MAXCOR + 1];
                for (int r = 0; r < 3; ++r)
                    for (int j = 0; j < W; ++j) {
                        double v = 0.0;
                        for (int p = 0; p < FIT_PARTS; ++p) v += partG[(((size_t)sidx * FIT_PARTS + p) * 3 + r) * W + j];
                        row[r][j] = v;
                    }
                if (act[sidx] == 2) {
                    const int p = pos[sidx];
                    for (int k = 0; k < H.col; ++k) {
                        const int j = H.order[k];
                        double ss = row[0][j], sy_pj = row[0][m + j], sy_jp = row[1][j], yy = row[1][m + j];
                        for (int c = 0; c < 3; ++c) {
                            ss += H.su[p][c] * H.su[j][c];
                            sy_pj += H.su[p][c] * H.yu[j][c];
                            sy_jp += H.su[j][c] * H.yu[p][c];
                            yy += H.yu[p][c] * H.yu[j][c];
                        }
                        H.SS[p][j] = H.SS[j][p] = ss;
                        H.SY[p][j] = sy_pj;
                        H.SY[j][p] = sy_jp;
                        H.YY[p][j] = H.YY[j][p] = yy;
                    }
                }
                // products with the gradient g = -G0 (logPhi part) and gu (hyper part)
                double bS[FIT_MAXCOR], bY[FIT_MAXCOR];
                for (int p = 0; p < m; ++p) { bS[p] = 0.0; bY[p] = 0.0; }
                for (int k = 0; k < H.col; ++k) {
                    const int p = H.order[k];
                    bS[p] = -row[2][p]; bY[p] = -row[2][m + p];
                    for (int c = 0; c < 3; ++c) { bS[p] += H.su[p][c] * s.gu[c]; bY[p] += H.yu[p][c] * s.gu[c]; }
                }
                double gg = row[2][2 * m];
                for (int c = 0; c < 3; ++c) gg += s.gu[c] * s.gu[c];
                double cg, cs[FIT_MAXCOR], cy[FIT_MAXCOR], gd;
                lbfgs_direction(H, bS, bY, gg, cg, cs, cy, gd);
                if (!(gd < 0.0) && H.col > 0) {   // not a descent direction: drop the memory (L-BFGS-B info = -4)
                    H.reset();
                    valid[sidx] = 0;
                    lbfgs_direction(H, bS, bY, gg, cg, cs, cy, gd);
                }
                // direction in both parts; on the device d = -cg' ... with g = -G0:  cg * g = (-cg) * G0
                double* cf = &coef[(size_t)sidx * W];
                for (int p = 0; p < m; ++p) { cf[p] = cs[p]; cf[m + p] = cy[p]; }
                cf[2 * m] = -cg;
                double dtd = 0.0;
                for (int c = 0; c < 3; ++c) {
                    double v = cg * s.gu[c];
                    for (int k = 0; k < H.col; ++k) { const int p = H.order[k]; v += cs[p] * H.su[p][c] + cy[p] * H.yu[p][c]; }
                    s.du[c] = O.train_hyp ? v : 0.0;
                }
                // |d|^2 from the coefficients (only needed for the very first step)
                if (H.col == 0) dtd = gg;
                for (int c = 0; c < 3; ++c) { s.u0[c] = s.u[c]; s.gu0[c] = s.gu[c]; }
                s.fold = s.f;
                s.gdold = gd;
                s.gd = gd;
                s.ifun = 0;
                double stp = (s.iter == 0 && H.col == 0) ? std::min(1.0 / std::sqrt(dtd), 1e10) : 1.0;
                s.ls = LineSearch();
                if (s.ls.start(stp, s.f, gd) != LineSearch::FG) {   // cannot happen for gd < 0; treat like a failed search
                    s.result = FIT_ABNORMAL_LNSRCH; s.phase = PH_FINAL; act[sidx] = 0; stpv[sidx] = 0.0; step_act[sidx] = 1;
                    continue;
                }
                s.ifun = 1;
                s.stp = stp;
                stpv[sidx] = stp;
                for (int c = 0; c < 3; ++c) s.u[c] = s.u0[c] + stp * s.du[c];
                s.phase = PH_SEARCH;
            }
            CU(cudaMemcpyAsync(d_act, act, (size_t)S * 16, cudaMemcpyHostToDevice, st));
            CU(cudaMemcpyAsync(d_coef, coef.data(), coef.size() * 8, cudaMemcpyHostToDevice, st));
            CU(cudaMemcpyAsync(d_stp, stpv.data(), (size_t)S * 8, cudaMemcpyHostToDevice, st));
            k_fit_dir<<<grid, 256, 0, st>>>(V, d_act, d_valid, d_coef, d_stp);
            CU(cudaStreamSynchronize(st));   // act / coef / stpv are reused below
        }
        bool any_step = false;
        for (int sidx = 0; sidx < S; ++sidx) { act[sidx] = step_act[sidx]; any_step = any_step || step_act[sidx]; }
        if (any_step) {
            CU(cudaMemcpyAsync(d_act, act, (size_t)S * 4, cudaMemcpyHostToDevice, st));
            CU(cudaMemcpyAsync(d_stp, stpv.data(), (size_t)S * 8, cudaMemcpyHostToDevice, st));
            k_fit_step<<<grid, 256, 0, st>>>(V, d_act, d_stp);
            CU(cudaStreamSynchronize(st));
        }
        CU(cudaGetLastError());
        s_opt += now() - t_round1;
    }
    cleanup();
    if (const char* e = getenv("BGP_FIT_STATS"))
        if (e[0] == '1')
            fprintf(stderr, "bgp_fit: items %d slots %d rounds %lld  evaluation %.3f s  optimiser passes %.3f s\n", F.count, S, rounds,
                    s_eval, s_opt);
    return BGP_OK;
}

// =============================================================================================== entry points
namespace {

struct DenseEval : Evaluator {
    bgp_handle* h; int N, D, kind, model; double klw;
    const double *t, *Y, *prior, *hyp, *b; const int* gene;
    double *elbo, *gh; int* status; int slots;
    int eval(int nslots, const double* X, double* G, cudaStream_t st) override {
        return dense_run(h, nslots, N, D, t, Y, gene, b, prior, hyp, kind, model, klw, X, nullptr, 1, elbo, gh, G, status, (char*)h->ws,
                         slots, st, false, 0);
    }
};

struct SparseEval : Evaluator {
    bgp_handle* h; int N, D, Mz, nY, kind;
    const double *t, *Z, *Y, *prior, *hyp, *b; const int* gene;
    double *elbo, *gh, *gb; int* status;
    int eval(int nslots, const double* X, double* G, cudaStream_t st) override {
        return bgp_sparse_elbo_grad(h, nslots, N, D, Mz, t, Z, Y, nY, gene, b, prior, hyp, kind, BGP_MODEL_BGP, 0, X, 1, elbo, gh, gb, G,
                                    status, (void*)st);
    }
};

struct DevBuf {   // one allocation carved into aligned pieces
    char* base = nullptr; size_t off = 0, cap = 0;
    template <class T> T* take(size_t count) { T* p = (T*)(base + off); off = align_up(off + count * sizeof(T), 256); return p; }
};

int check_fit_options(bgp_handle* h, const bgp_fit_options* o) {
    if (!o) return fail(h, BGP_ERR_ARG, "null options");
    if (o->maxcor < 1 || o->maxcor > FIT_MAXCOR) return fail(h, BGP_ERR_ARG, "maxcor must be in 1..%d", FIT_MAXCOR);
    if (o->maxiter < 0 || o->maxls < 1 || o->maxfun < 1) return fail(h, BGP_ERR_ARG, "maxiter, maxls, maxfun must be positive");
    for (int k = 0; k < 3; ++k)
        if (o->train_hyp && o->has_prior[k] && !(o->prior_sd[k] > 0.0)) return fail(h, BGP_ERR_ARG, "prior_sd must be positive");
    return BGP_OK;
}

}  // namespace

extern "C" {

void bgp_fit_default_options(bgp_fit_options* o) {
    if (!o) return;
    memset(o, 0, sizeof(*o));
    o->maxiter = 100; o->maxcor = 10; o->maxls = 20; o->maxfun = 15000;
    o->ftol = 2.220446049250313e-09; o->gtol = 1e-5;
    o->train_hyp = 1;
    // FitBranchingModel.py:109-117: Normal(2, 1) on the lengthscale, Normal(3, 1) on the kernel variance, Normal(0.1, 0.1) on the noise
    const double mu[3] = {2.0, 3.0, 0.1}, sd[3] = {1.0, 1.0, 0.1};
    for (int k = 0; k < 3; ++k) { o->has_prior[k] = 1; o->prior_mean[k] = mu[k]; o->prior_sd[k] = sd[k]; }
    o->max_slots = 0;
}

static int fit_common(bgp_handle* h, bool sparse, int count, int N, int D, int Mz, const double* t, const double* Z, const double* Y, int nY,
                      const int* item_gene, const double* b, const double* prior, const double* phiInitial, const double* hyp0,
                      int kernel_kind, const bgp_fit_options* opt, int T, const double* Xnew, double* objective, double* elbo_out,
                      double* hyp_out, double* phi_out, double* pred_mean, double* pred_var, int* iters, int* nfev, int* status,
                      double* logPhi_out) {
    if (!h) return BGP_ERR_ARG;
    int rc = check_fit_options(h, opt);
    if (rc != BGP_OK) return rc;
    if (count <= 0 || N <= 0 || D <= 0 || nY <= 0) return fail(h, BGP_ERR_ARG, "count, N, D, nY must be positive");
    if (!t || !Y || !item_gene || !b || !prior || !phiInitial || !hyp0 || !objective || !elbo_out || !hyp_out || !phi_out || !iters ||
        !nfev || !status)
        return fail(h, BGP_ERR_ARG, "null pointer");
    if (T < 0 || (T > 0 && (!Xnew || !pred_mean || !pred_var))) return fail(h, BGP_ERR_ARG, "prediction needs Xnew, pred_mean and pred_var");
    if (sparse && (!Z || Mz <= 0 || Mz > SP_MAX_MZ)) return fail(h, BGP_ERR_ARG, "sparse fit needs 0 < Mz <= %d inducing rows", SP_MAX_MZ);
    if (kernel_kind != BGP_KERNEL_MATERN32 && kernel_kind != BGP_KERNEL_SQEXP) return fail(h, BGP_ERR_ARG, "unknown kernel kind %d", kernel_kind);
    for (int i = 0; i < count; ++i)
        if (item_gene[i] < 0 || item_gene[i] >= nY) return fail(h, BGP_ERR_ARG, "item_gene[%d] out of range", i);
    CU(cudaSetDevice(h->device));
    cudaStream_t st = 0;
    const long long n = (long long)N * 3 * N;
    const int m = opt->maxcor;
    const DenseDims dd = make_dims(N, D);
    const SlotLayout SL = slot_layout(dd);
    // slots: one per SM at most; bounded by memory (optimiser state + evaluator workspace)
    int S = opt->max_slots > 0 ? opt->max_slots : (h->max_slots > 0 ? h->max_slots : h->sm_count);
    if (S > count) S = count;
    const size_t state_per_slot = (size_t)(5 + 2 * m) * n * 8;
    size_t eval_per_slot;
    if (sparse) {
        const int Mp = (3 * N + SP_SL - 1) / SP_SL * SP_SL;
        eval_per_slot = (size_t)sparse_layout(N, Mp, Mz, D, 16, dd.nch).total * 8 + 4096;
    } else {
        eval_per_slot = SL.total;
    }
    if (h->stage) { cudaFree(h->stage); h->stage = nullptr; h->stage_bytes = 0; }
    size_t free_b = 0, total_b = 0;
    CU(cudaMemGetInfo(&free_b, &total_b));
    const size_t avail = (size_t)(0.92 * (double)(free_b + h->ws_bytes + h->fit_state_bytes));
    const size_t fixed = (size_t)nY * N * D * 8 + ((size_t)64 << 20);
    if (avail < fixed + state_per_slot + eval_per_slot) return fail(h, BGP_ERR_NOMEM, "one work item at N=%d does not fit in device memory", N);
    const size_t fit = (avail - fixed) / (state_per_slot + eval_per_slot);
    if ((size_t)S > fit) S = (int)fit;
    if (!sparse) {
        // the workspace is grow-only; a larger one left by a host-pointer call would not be covered by the budget above
        if (h->ws && h->ws_bytes > (size_t)S * SL.total) { cudaFree(h->ws); h->ws = nullptr; h->ws_bytes = 0; }
        rc = ensure_ws(h, (size_t)S * SL.total);
        if (rc != BGP_OK) return rc;
    }
    // ---- device state
    DevBuf B;
    const int Tq = T > 0 ? T : 1;
    size_t need = 0;
    {
        DevBuf probe;   // measure
        probe.take<double>((size_t)S * n * (5 + 2 * m) + 64);
        probe.take<double>(N); probe.take<double>((size_t)nY * N * D); probe.take<double>((size_t)N * 2); probe.take<double>((size_t)N * 2);
        probe.take<double>((size_t)Mz * 2 + 2);
        probe.take<double>((size_t)S * 3); probe.take<double>(S); probe.take<int>(S); probe.take<double>(S); probe.take<double>((size_t)S * 4);
        probe.take<double>(S); probe.take<int>(S);
        probe.take<double>((size_t)S * N * 3); probe.take<double>((size_t)S * Tq * 2); probe.take<double>((size_t)S * Tq * D);
        probe.take<double>((size_t)S * Tq); probe.take<double>((size_t)3 * (Mz + 1) * Tq);
        need = probe.off + 4096;
    }
    // the optimiser state is kept in the handle between calls (FitModelBatch calls once per grid point); grow-only
    if (h->fit_state_bytes < need) {
        if (h->fit_state) { cudaFree(h->fit_state); h->fit_state = nullptr; h->fit_state_bytes = 0; }
        cudaError_t ce = cudaMalloc(&h->fit_state, need);
        if (ce != cudaSuccess) { cudaGetLastError(); return fail(h, BGP_ERR_NOMEM, "optimiser state of %zu bytes: %s", need, cudaGetErrorString(ce)); }
        h->fit_state_bytes = need;
    }
    B.base = (char*)h->fit_state;
    B.cap = need;
    FitVecs V;
    V.n = n; V.m = m;
    double* vecs = B.take<double>((size_t)S * n * (5 + 2 * m) + 64);
    V.X = vecs; V.G = V.X + (size_t)S * n; V.X0 = V.G + (size_t)S * n; V.G0 = V.X0 + (size_t)S * n; V.Dv = V.G0 + (size_t)S * n;
    V.Sh = V.Dv + (size_t)S * n; V.Yh = V.Sh + (size_t)S * m * n;
    double* d_t = B.take<double>(N);
    double* d_Y = B.take<double>((size_t)nY * N * D);
    double* d_prior = B.take<double>((size_t)N * 2);
    double* d_phi0 = B.take<double>((size_t)N * 2);
    double* d_Z = B.take<double>((size_t)Mz * 2 + 2);
    double* d_hyp = B.take<double>((size_t)S * 3);
    double* d_b = B.take<double>(S);
    int* d_gene = B.take<int>(S);
    double* d_elbo = B.take<double>(S);
    double* d_gh = B.take<double>((size_t)S * 4);
    double* d_gb = B.take<double>(S);
    int* d_status = B.take<int>(S);
    double* d_phi = B.take<double>((size_t)S * N * 3);
    double* d_X = B.take<double>((size_t)S * Tq * 2);
    double* d_mean = B.take<double>((size_t)S * Tq * D);
    double* d_var = B.take<double>((size_t)S * Tq);
    double* d_scr = B.take<double>((size_t)3 * (Mz + 1) * Tq);
    auto bail = [&](int code) { return code; };
#define CUF(call)                                                                                                   \
    do {                                                                                                            \
        cudaError_t e_ = (call);                                                                                    \
        if (e_ != cudaSuccess) return bail(fail(h, BGP_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e_)));         \
    } while (0)
    CUF(cudaMemcpyAsync(d_t, t, (size_t)N * 8, cudaMemcpyHostToDevice, st));
    CUF(cudaMemcpyAsync(d_Y, Y, (size_t)nY * N * D * 8, cudaMemcpyHostToDevice, st));
    CUF(cudaMemcpyAsync(d_prior, prior, (size_t)N * 16, cudaMemcpyHostToDevice, st));
    CUF(cudaMemcpyAsync(d_phi0, phiInitial, (size_t)N * 16, cudaMemcpyHostToDevice, st));
    if (sparse) CUF(cudaMemcpyAsync(d_Z, Z, (size_t)Mz * 16, cudaMemcpyHostToDevice, st));
    CUF(cudaMemsetAsync(d_gh, 0, (size_t)S * 32, st));

    FitShared F;
    F.h = h; F.count = count; F.N = N; F.D = D; F.n = n; F.opt = opt;
    F.item_gene = item_gene; F.b = b; F.hyp0 = hyp0;
    F.objective = objective; F.elbo_out = elbo_out; F.hyp_out = hyp_out; F.phi_out = phi_out; F.pred_mean = pred_mean; F.pred_var = pred_var;
    F.logPhi_out = logPhi_out; F.iters = iters; F.nfev = nfev; F.status = status; F.T = T; F.want_predict = T > 0;
    F.gh_stride = sparse ? 3 : 4;

    DenseEval de;
    SparseEval se;
    Evaluator* ev;
    if (sparse) {
        se.h = h; se.N = N; se.D = D; se.Mz = Mz; se.nY = nY; se.kind = kernel_kind; se.t = d_t; se.Z = d_Z; se.Y = d_Y; se.prior = d_prior;
        se.hyp = d_hyp; se.b = d_b; se.gene = d_gene; se.elbo = d_elbo; se.gh = d_gh; se.gb = d_gb; se.status = d_status;
        ev = &se;
    } else {
        de.h = h; de.N = N; de.D = D; de.kind = kernel_kind; de.model = BGP_MODEL_BGP; de.klw = 1.0; de.t = d_t; de.Y = d_Y; de.prior = d_prior;
        de.hyp = d_hyp; de.b = d_b; de.gene = d_gene; de.elbo = d_elbo; de.gh = d_gh; de.status = d_status; de.slots = S;
        ev = &de;
    }
    const int saved_slots = h->max_slots;
    if (sparse) h->max_slots = S;   // bgp_sparse_elbo_grad sizes its waves from the handle

    auto init_x = [&](int slot, int item) -> int {
        k_fit_init_logphi<<<FIT_PARTS * 4, 256, 0, st>>>(V.X, n, slot, N, d_t, d_phi0, b[item]);
        CU(cudaGetLastError());
        return BGP_OK;
    };
    int* d_map = nullptr;
    CUF(cudaMalloc(&d_map, (size_t)S * 4));
    auto finish_slots = [&](const std::vector<int>& fslots, const std::vector<int>& fitems) -> int {
        // the slots' workspaces hold the factorisations at their final points (state after a bound + gradient evaluation)
        const int nf = (int)fslots.size();
        for (int k = 0; k < nf; ++k) {
            const int slot = fslots[k], item = fitems[k];
            phi_launch_gather(V.X + (size_t)slot * n, d_phi + (size_t)slot * N * 3, 1, N, st);
            CU(cudaMemcpyAsync(phi_out + (size_t)item * N * 3, d_phi + (size_t)slot * N * 3, (size_t)N * 24, cudaMemcpyDeviceToHost, st));
            if (T > 0) CU(cudaMemcpyAsync(d_X + (size_t)slot * T * 2, Xnew + (size_t)item * T * 2, (size_t)T * 16, cudaMemcpyHostToDevice, st));
            if (logPhi_out) CU(cudaMemcpyAsync(logPhi_out + (size_t)item * n, V.X + (size_t)slot * n, (size_t)n * 8, cudaMemcpyDeviceToHost, st));
        }
        if (T > 0) {
            if (sparse) {
                for (int k = 0; k < nf; ++k) {
                    const int slot = fslots[k];
                    SparseParams P = h->lastS;
                    P.ws = P.ws + (size_t)slot * P.slot_stride;
                    P.item0 = slot;
                    sparse_launch_predict(P, d_X + (size_t)slot * T * 2, T, 0, d_scr, d_mean + (size_t)slot * T * D, d_var + (size_t)slot * T, st);
                }
            } else {
                CU(cudaMemcpyAsync(d_map, fslots.data(), (size_t)nf * 4, cudaMemcpyHostToDevice, st));
                DenseParams P = h->lastP;
                P.item0 = 0;
                P.slot_map = d_map;
                dense_launch_predict_slots(P, nf, d_X, (long long)T * 2, T, d_mean, d_var, st);
            }
            for (int k = 0; k < nf; ++k) {
                const int slot = fslots[k], item = fitems[k];
                CU(cudaMemcpyAsync(pred_mean + (size_t)item * T * D, d_mean + (size_t)slot * T * D, (size_t)T * D * 8, cudaMemcpyDeviceToHost, st));
                CU(cudaMemcpyAsync(pred_var + (size_t)item * T, d_var + (size_t)slot * T, (size_t)T * 8, cudaMemcpyDeviceToHost, st));
            }
        }
        CU(cudaGetLastError());
        CU(cudaStreamSynchronize(st));   // host vectors above go out of scope; the slots are refilled next
        return BGP_OK;
    };
    h->in_fit = true;
    rc = fit_core(F, *ev, S, V, st, d_hyp, d_b, d_gene, d_elbo, d_gh, d_status, 1, init_x, finish_slots);
    h->in_fit = false;
    cudaFree(d_map);
    h->max_slots = saved_slots;
    cudaError_t e2 = cudaStreamSynchronize(st);
    if (rc == BGP_OK && e2 != cudaSuccess) rc = fail(h, BGP_ERR_CUDA, "fit: %s", cudaGetErrorString(e2));
    return rc;
#undef CUF
}

int bgp_dense_fit_host(bgp_handle* h, int count, int N, int D, const double* t, const double* Y, int nY, const int* item_gene,
                       const double* b, const double* prior, const double* phiInitial, const double* hyp0, int kernel_kind,
                       const bgp_fit_options* opt, int T, const double* Xnew, double* objective, double* elbo, double* hyp,
                       double* phi, double* pred_mean, double* pred_var, int* iters, int* nfev, int* status, double* logPhi_out) {
    return fit_common(h, false, count, N, D, 0, t, nullptr, Y, nY, item_gene, b, prior, phiInitial, hyp0, kernel_kind, opt, T, Xnew,
                      objective, elbo, hyp, phi, pred_mean, pred_var, iters, nfev, status, logPhi_out);
}

int bgp_sparse_fit_host(bgp_handle* h, int count, int N, int D, int Mz, const double* t, const double* Z, const double* Y, int nY,
                        const int* item_gene, const double* b, const double* prior, const double* phiInitial, const double* hyp0,
                        int kernel_kind, const bgp_fit_options* opt, int T, const double* Xnew, double* objective, double* elbo,
                        double* hyp, double* phi, double* pred_mean, double* pred_var, int* iters, int* nfev, int* status,
                        double* logPhi_out) {
    return fit_common(h, true, count, N, D, Mz, t, Z, Y, nY, item_gene, b, prior, phiInitial, hyp0, kernel_kind, opt, T, Xnew, objective,
                      elbo, hyp, phi, pred_mean, pred_var, iters, nfev, status, logPhi_out);
}

// ---- host-only test hooks (no GPU needed) -------------------------------------------------------------------------
// One More'-Thuente call.  state[16] is opaque storage owned by the caller; start != 0 begins a search at step 0 with
// (f, g) and the initial trial *stp.  Returns 0 FG (evaluate at *stp), 1 convergence, 2 warning, 3 error.
int bgp_debug_linesearch(double* state, int start, double* stp, double f, double g) {
    static_assert(sizeof(LineSearch) <= 32 * sizeof(double), "state too small");
    LineSearch ls;
    if (!start) memcpy(&ls, state, sizeof(ls));
    const int task = start ? (int)ls.start(*stp, f, g) : (int)ls.next(*stp, f, g);
    memcpy(state, &ls, sizeof(ls));
    return task;
}

// Two-loop direction from inner products: S, Y are [col][n] (oldest first), g [n]; out d[n] = -H g.
int bgp_debug_lbfgs_direction(int col, int n, const double* S, const double* Y, const double* g, double* d) {
    if (col < 0 || col > FIT_MAXCOR || n <= 0) return BGP_ERR_ARG;
    History H;
    H.m = FIT_MAXCOR;
    H.col = col;
    auto dot = [&](const double* a, const double* b_) { double v = 0; for (int k = 0; k < n; ++k) v += a[k] * b_[k]; return v; };
    double bS[FIT_MAXCOR] = {0}, bY[FIT_MAXCOR] = {0};
    for (int i = 0; i < col; ++i) {
        H.order[i] = i;
        for (int j = 0; j < col; ++j) {
            H.SS[i][j] = dot(S + (size_t)i * n, S + (size_t)j * n);
            H.SY[i][j] = dot(S + (size_t)i * n, Y + (size_t)j * n);
            H.YY[i][j] = dot(Y + (size_t)i * n, Y + (size_t)j * n);
        }
        bS[i] = dot(S + (size_t)i * n, g);
        bY[i] = dot(Y + (size_t)i * n, g);
    }
    double cg, cs[FIT_MAXCOR], cy[FIT_MAXCOR], gd;
    lbfgs_direction(H, bS, bY, dot(g, g), cg, cs, cy, gd);
    for (int k = 0; k < n; ++k) {
        double v = cg * g[k];
        for (int i = 0; i < col; ++i) v += cs[i] * S[(size_t)i * n + k] + cy[i] * Y[(size_t)i * n + k];
        d[k] = v;
    }
    return BGP_OK;
}

}  // extern "C"
