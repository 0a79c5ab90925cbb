// Internals shared by the translation units behind the C ABI (bgp_api.cu, bgp_fit.cu): handle, workspace layout, wave loop.
#pragma once
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../include/branchedgp_b200.h"
#include "bgp_launch.h"

struct bgp_handle {
    int device = 0;
    int sm_count = 0;
    int max_slots = 0;         // 0 = default
    int stop_phase = 0;
    int wide_mode = -1;        // multi-CTA-per-item dense pipeline: -1 automatic (few items in flight), 0 never, 1 always
    char err[512] = {0};
    // workspace (grow-only)
    void* ws = nullptr;
    size_t ws_bytes = 0;
    // staging for the _host entry points (grow-only)
    void* stage = nullptr;
    size_t stage_bytes = 0;
    // optimiser state of the batched fit (bgp_fit.cu), kept between calls (grow-only)
    void* fit_state = nullptr;
    size_t fit_state_bytes = 0;
    bool in_fit = false;       // set while a fit call runs: its evaluator calls must not drop the state to make room
    // description of the last dense wave, for the debug readers
    bgp::DenseParams lastP;
    bgp::SparseParams lastS;        // parameters of the last sparse wave (device pointers into the staging area stay valid
    bool keep_sparse = false;  //  until the next call), used by bgp_sparse_predict_host
    int last_slots = 0;
    bool kernels_ready = false;
    std::vector<cudaStream_t> lane_streams;   // streams of the host-pointer entry points (one per in-flight item group)
    cudaEvent_t shared_ready = nullptr;
    std::vector<cudaEvent_t> sp_events;       // stage hand-off events of the pipelined sparse host entry
    void* pin = nullptr;       // pinned host scratch for the small per-item results of the host-pointer entry points
    size_t pin_bytes = 0;
    // optional per-kernel timing (bench.py): events recorded around every launch of a call
    bool profiling = false;
    std::vector<cudaEvent_t> ev_pool;
    std::vector<std::pair<int, int>> ev_marks;   // (phase id, index into ev_pool of the event recorded AFTER the launch)
    int ev_used = 0;
    // set by bgp_dense_predict_host around a forward-only dense call
    const double* pred_X = nullptr; int pred_T = 0; int pred_full = 0; double* pred_mean = nullptr; double* pred_var = nullptr;
    void* pred_buf = nullptr; size_t pred_buf_bytes = 0;   // predict_f: test inputs, outputs and the full-covariance scratch (grow-only)
    double* pred_ext = nullptr; double* pred_neg1 = nullptr;   // parts of pred_buf: tall matrices of the tiled full covariance, Mp x -1.0
    // optional NCCL communicator of bgp_comm_init / bgp_gather_results (bgp_comm.cu)
    void* comm = nullptr; int comm_ranks = 0, comm_rank = 0; void* comm_buf = nullptr; size_t comm_buf_bytes = 0;
    double prof_ms[BGP_NPHASES] = {0};
    long long launches = 0;
};

void prof_mark(bgp_handle* h, int phase, cudaStream_t st);
int fail(bgp_handle* h, int code, const char* fmt, ...);
#define CU(call)                                                                                       \
    do {                                                                                               \
        cudaError_t e_ = (call);                                                                       \
        if (e_ != cudaSuccess) return fail(h, BGP_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e_)); \
    } while (0)

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

bgp::DenseDims make_dims(int N, int D);

struct SlotLayout {   // byte offsets inside one slot
    size_t LC, RX, PB, LB, LCe, LinvD, RinvD, SCR, KT, vec, Apart, vpart, klpart, lse, klrow, rowpart, ctrl, ctrl_bytes, total;
};
SlotLayout slot_layout(const bgp::DenseDims& d);
int ensure_ws(bgp_handle* h, size_t bytes);
int ensure_stage(bgp_handle* h, size_t bytes);

// One call's worth of waves of the dense pipeline on `st` (see bgp_api.cu).
int dense_run(bgp_handle* h, int count, int N, int D, const double* t, const double* Y, const int* item_gene, const double* b,
              const double* prior, const double* hyp, int kernel_kind, int model, double kl_weight, const double* logPhi,
              const double* KConst, int want_grad, double* elbo, double* grad_hyp, double* grad_logPhi, int* status,
              char* base, int slots, cudaStream_t st, bool mark, int global_item0, bool allow_wide = true);
