// extern "C" boundary of libbgp_b200.so (declared in include/branchedgp_b200.h): handle, workspace, wave loop.
#include "bgp_internal.h"

using namespace bgp;

extern "C" {
static int ensure_lane_streams(bgp_handle* h, int lanes);
static int ensure_pin(bgp_handle* h, size_t bytes);
}

void prof_mark(bgp_handle* h, int phase, cudaStream_t st) {
    if (phase >= 0) h->launches++;
    if (!h->profiling) return;
    if (h->ev_used == (int)h->ev_pool.size()) {
        cudaEvent_t e;
        cudaEventCreate(&e);
        h->ev_pool.push_back(e);
    }
    cudaEventRecord(h->ev_pool[h->ev_used], st);
    h->ev_marks.push_back({phase, h->ev_used});
    h->ev_used++;
}

int fail(bgp_handle* h, int code, const char* fmt, ...) {
    if (h) {
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(h->err, sizeof(h->err), fmt, ap);
        va_end(ap);
    }
    return code;
}
DenseDims make_dims(int N, int D) {
    DenseDims d;
    d.N = N;
    d.M = 3 * N;
    d.Mp = (d.M + MB - 1) / MB * MB;
    d.nT = d.Mp / TS;
    d.nM = d.Mp / MB;
    d.nTv = (d.M + TS - 1) / TS;
    d.D = D;
    d.nch = (N + PHI_RCH - 1) / PHI_RCH;
    // PB / LB double as tall nT x 4 tile matrices in k_predict, which is larger than the packed triangle when nT < 7
    const long long tiles = packed_tiles(d.nT) > 4LL * d.nT ? packed_tiles(d.nT) : 4LL * d.nT;
    d.mat_elems = tiles * TILE_ELEMS;
    d.vec_elems = vec_layout(d.Mp, d.nM, D).total;
    return d;
}

SlotLayout slot_layout(const DenseDims& d) {
    SlotLayout L;
    size_t o = 0;
    auto take = [&](size_t doubles) { size_t r = o; o = align_up(o + doubles * 8, 256); return r; };
    L.LC = take(d.mat_elems); L.RX = take(d.mat_elems); L.PB = take(d.mat_elems); L.LB = take(d.mat_elems);
    L.LCe = take((size_t)d.nT * TILE_ELEMS);
    L.LinvD = take((size_t)d.nM * 16 * TILE_ELEMS);
    L.RinvD = take((size_t)d.nM * 16 * TILE_ELEMS);
    L.SCR = take((size_t)2 * 16 * TILE_ELEMS);
    L.KT = take((size_t)2 * d.N * d.N);
    L.vec = take(d.vec_elems);
    L.Apart = take((size_t)d.nch * d.Mp);
    L.vpart = take((size_t)d.nch * d.D * d.Mp);
    L.klpart = take(d.nch);
    L.lse = take(d.N);
    L.klrow = take((size_t)d.N * PHI_KLW);
    L.rowpart = take((size_t)d.N * phi_col_chunks(d.M));
    L.ctrl_bytes = (size_t)dense_wide_ctrl_ints(d.nM) * 4;
    L.ctrl = take((L.ctrl_bytes + 7) / 8);
    L.total = o;
    return L;
}

extern "C" {

int bgp_version(void) { return 100; }

int bgp_create(int device, bgp_handle** out) {
    if (!out) return BGP_ERR_ARG;
    bgp_handle* h = new bgp_handle();
    *out = h;
    h->device = device;
    CU(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    h->sm_count = prop.multiProcessorCount;
    if (prop.major < 10) return fail(h, BGP_ERR_CUDA, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
    CU(dense_init_kernels());
    CU(dense_wide_init_kernels());
    CU(sparse_init_kernels());
    h->kernels_ready = true;
    return BGP_OK;
}

int bgp_destroy(bgp_handle* h) {
    if (!h) return BGP_OK;
    cudaSetDevice(h->device);
    for (cudaEvent_t e : h->ev_pool) cudaEventDestroy(e);
    for (cudaStream_t s_ : h->lane_streams) cudaStreamDestroy(s_);
    if (h->shared_ready) cudaEventDestroy(h->shared_ready);
    bgp_comm_destroy(h);
    if (h->comm_buf) cudaFree(h->comm_buf);
    if (h->pin) cudaFreeHost(h->pin);
    if (h->pred_buf) cudaFree(h->pred_buf);
    if (h->ws) cudaFree(h->ws);
    if (h->stage) cudaFree(h->stage);
    if (h->fit_state) cudaFree(h->fit_state);
    delete h;
    return BGP_OK;
}

const char* bgp_last_error(const bgp_handle* h) { return h ? h->err : "null handle"; }
int bgp_device_sm_count(const bgp_handle* h) { return h ? h->sm_count : 0; }
int bgp_set_max_slots(bgp_handle* h, int max_slots) {
    if (!h || max_slots < 0) return BGP_ERR_ARG;
    h->max_slots = max_slots;
    return BGP_OK;
}
size_t bgp_dense_slot_bytes(int N, int D) { return slot_layout(make_dims(N, D)).total; }

int bgp_host_alloc(void** p, size_t bytes) { return cudaHostAlloc(p, bytes, cudaHostAllocDefault) == cudaSuccess ? BGP_OK : BGP_ERR_NOMEM; }
int bgp_host_free(void* p) { return cudaFreeHost(p) == cudaSuccess ? BGP_OK : BGP_ERR_CUDA; }

}  // extern "C"
// The optimiser state cached by the last fit is given back before an evaluation entry point budgets device memory.
static void drop_fit_state(bgp_handle* h) {
    if (h->fit_state && !h->in_fit) { cudaFree(h->fit_state); h->fit_state = nullptr; h->fit_state_bytes = 0; }
}

int ensure_ws(bgp_handle* h, size_t bytes) {
    if (h->ws_bytes >= bytes) return BGP_OK;
    if (h->ws) { cudaFree(h->ws); h->ws = nullptr; h->ws_bytes = 0; }
    cudaError_t e = cudaMalloc(&h->ws, bytes);
    if (e != cudaSuccess) { cudaGetLastError(); return fail(h, BGP_ERR_NOMEM, "workspace of %zu bytes: %s", bytes, cudaGetErrorString(e)); }
    h->ws_bytes = bytes;
    return BGP_OK;
}
extern "C" {

static int pick_slots(bgp_handle* h, int count, size_t slot_bytes) {
    int cap = h->max_slots > 0 ? h->max_slots : h->sm_count;
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) {
        const size_t avail = free_b + h->ws_bytes;
        const size_t fit = (size_t)(0.9 * (double)avail) / slot_bytes;
        if ((size_t)cap > fit) cap = (int)fit;
    }
    if (cap > count) cap = count;
    return cap;
}

}  // extern "C"
// One call's worth of waves on `st`, using `slots` workspace slots starting at `base`.  `global_item0` is the index of the
// call's first item in caller-global arrays that are not staged per call (the predict outputs).
int dense_run(bgp_handle* h, int count, int N, int D, const double* t, const double* Y, const int* item_gene, const double* b,
                     const double* prior, const double* hyp, int kernel_kind, int model, double kl_weight, const double* logPhi,
                     const double* KConst, int want_grad, double* elbo, double* grad_hyp, double* grad_logPhi, int* status,
                     char* base, int slots, cudaStream_t st, bool mark, int global_item0, bool allow_wide) {
    const DenseDims d = make_dims(N, D);
    const SlotLayout SL = slot_layout(d);
    const size_t stride_d = SL.total / 8;   // slot stride in doubles (SL.total is a multiple of 256)

    DenseParams P;
    memset(&P, 0, sizeof(P));
    P.d = d;
    P.kind = kernel_kind;
    P.model = model;
    P.square_noise = 1e-6;
    P.kl_weight = kl_weight;
    P.t = t; P.hyp = hyp; P.bv = b; P.KConst = KConst;
    P.LC = (double*)(base + SL.LC); P.RX = (double*)(base + SL.RX); P.PB = (double*)(base + SL.PB); P.LB = (double*)(base + SL.LB);
    P.LCe = (double*)(base + SL.LCe); P.LinvD = (double*)(base + SL.LinvD); P.RinvD = (double*)(base + SL.RinvD);
    P.SCR = (double*)(base + SL.SCR); P.KT = (double*)(base + SL.KT); P.vec = (double*)(base + SL.vec);
    P.status = status; P.elbo = elbo; P.grad_hyp = grad_hyp;
    P.slot_stride = (long long)stride_d;
    P.ctrl = (int*)(base + SL.ctrl);
    P.ctrl_stride = (long long)(SL.total / 4);
    // Few items in flight: deal the macro tiles of every item to G CTAs (bgp_dense_wide.cu).  BGP_DENSE_WIDE=0 / 1 forces a path.
    int wide_mode = h->wide_mode;
    if (const char* e = getenv("BGP_DENSE_WIDE")) { if (e[0] == '0') wide_mode = 0; else if (e[0] == '1') wide_mode = 1; }

    PhiParams F;
    memset(&F, 0, sizeof(F));
    F.N = N; F.M = d.M; F.Mp = d.Mp; F.D = D; F.nch = d.nch; F.model = model; F.kl_weight = kl_weight;
    F.SD = 1; F.Dtot = D;
    F.logPhi = logPhi; F.t = t; F.prior = prior; F.bv = b; F.Y = Y; F.item_gene = item_gene;
    F.lse = (double*)(base + SL.lse); F.Apart = (double*)(base + SL.Apart); F.vpart = (double*)(base + SL.vpart);
    F.klpart = (double*)(base + SL.klpart);
    F.klrow = (double*)(base + SL.klrow);
    F.rowpart = (double*)(base + SL.rowpart);
    F.rch = PHI_RCH;
    F.slot_stride = (long long)stride_d;
    const VecLayout VL = vec_layout(d.Mp, d.nM, D);
    F.Abar = P.vec + VL.Abar; F.vbar = P.vec + VL.vbar;
    F.grad_logPhi = grad_logPhi;

    auto pm = [&](int phase) { if (mark) prof_mark(h, phase, st); else if (phase >= 0) h->launches++; };
    for (int item0 = 0, n = 0; item0 < count; item0 += n) {
        n = (count - item0 < slots) ? (count - item0) : slots;
        // between half and three quarters of an SM count of items: two many-CTA waves of half the items beat one one-CTA wave
        // (N = 1000: 74 items take 262 ms with two CTAs each, 402 ms with one)
        // (not inside a fit: there slot == item for the whole call, the slots keep the factorisations of their items)
        if (wide_mode < 0 && allow_wide && !h->in_fit && 2 * n > h->sm_count && 4 * n <= 3 * h->sm_count) n = (n + 1) / 2;
        P.item0 = item0;
        F.item0 = item0;
        const int last = h->stop_phase > 0 ? h->stop_phase : 99;
        pm(-1);
        phi_launch_forward(F, n, st);
        pm(0);
        dense_launch_prep(P, F.Apart, F.vpart, F.klpart, Y, item_gene, n, st);   // k_prep (+ k_ktable unless KConst)
        if (KConst == nullptr) h->launches++;
        pm(1);
        const bool wide = wide_mode == 1 || (wide_mode < 0 && allow_wide && 2 * n <= h->sm_count);
        int G = h->sm_count / n;
        if (G < 1) G = 1;
        if (wide) CU(cudaMemset2DAsync(P.ctrl, SL.total, 0, SL.ctrl_bytes, (size_t)n, st));
        for (int phase = 1; phase <= 9 && phase <= last; ++phase) {
            const bool grad_phase = (phase == 5 || phase == 6 || phase == 7 || phase == 9);
            if (grad_phase && !want_grad) continue;
            if (!wide || !dense_launch_phase_wide(P, n, phase, G, st)) dense_launch_phase(P, n, phase, st);
            pm(phase + 1);
        }
        if (want_grad && last >= 9) {
            phi_launch_backward(F, n, st);
            pm(11);
        }
        if (h->pred_T > 0) {   // predict_f on this wave's forward state
            const size_t vstride = h->pred_full ? (size_t)h->pred_T * h->pred_T : (size_t)h->pred_T;
            const size_t g0 = (size_t)global_item0 + item0;
            if (h->pred_full && h->pred_T > MB) {   // full covariance over more than one 128-input chunk: tall matrices in pred_ext
                const size_t per_item = dense_predict_cov_scratch_doubles(d, 1, h->pred_T) - d.Mp;
                dense_launch_predict_fullcov(P, n, h->pred_X, h->pred_T, h->pred_mean + g0 * h->pred_T * D, h->pred_var + g0 * vstride,
                                             h->pred_ext + g0 * per_item, h->pred_neg1, st);
                h->launches += 2;
            } else {
                for (int s0 = 0; s0 < h->pred_T; s0 += MB) {   // 128 test inputs per launch
                    const int tc = (h->pred_T - s0 < MB) ? (h->pred_T - s0) : MB;
                    dense_launch_predict(P, n, h->pred_X + 2 * (size_t)s0, tc, h->pred_T, s0, h->pred_full,
                                         h->pred_mean + g0 * h->pred_T * D, h->pred_var + g0 * vstride, st);
                    h->launches++;
                }
            }
        }
        CU(cudaGetLastError());
        h->last_slots = n;
    }
    h->lastP = P;
    return BGP_OK;
}
extern "C" {

static int dense_check_args(bgp_handle* h, int count, int N, int D, int nY, const void* t, const void* Y, const void* item_gene,
                            const void* b, const void* prior, const void* hyp, int kernel_kind, int model, const void* logPhi,
                            int want_grad, const void* elbo, const void* grad_hyp, const void* grad_logPhi, const void* status) {
    if (count <= 0 || N <= 0 || D <= 0 || nY <= 0) return fail(h, BGP_ERR_ARG, "count, N, D, nY must be positive");
    if (!t || !Y || !item_gene || !b || !prior || !hyp || !logPhi || !elbo || !status) return fail(h, BGP_ERR_ARG, "null input pointer");
    if (want_grad && (!grad_hyp || !grad_logPhi)) return fail(h, BGP_ERR_ARG, "want_grad needs grad_hyp and grad_logPhi");
    if (kernel_kind != BGP_KERNEL_MATERN32 && kernel_kind != BGP_KERNEL_SQEXP) return fail(h, BGP_ERR_ARG, "unknown kernel kind %d", kernel_kind);
    if (model != BGP_MODEL_BGP && model != BGP_MODEL_MBGP) return fail(h, BGP_ERR_ARG, "unknown model %d", model);
    return BGP_OK;
}

int bgp_dense_elbo_grad(bgp_handle* h, int count, int N, int D, const double* t, const double* Y, int nY,
                        const int* item_gene, const double* b, const double* prior, const double* hyp, int kernel_kind,
                        int model, double kl_weight, const double* logPhi, const double* KConst, int want_grad,
                        double* elbo, double* grad_hyp, double* grad_logPhi, int* status, void* cuda_stream) {
    if (!h) return BGP_ERR_ARG;
    int rc = dense_check_args(h, count, N, D, nY, t, Y, item_gene, b, prior, hyp, kernel_kind, model, logPhi, want_grad, elbo, grad_hyp,
                              grad_logPhi, status);
    if (rc != BGP_OK) return rc;
    CU(cudaSetDevice(h->device));
    drop_fit_state(h);
    const SlotLayout SL = slot_layout(make_dims(N, D));
    const int slots = pick_slots(h, count, SL.total);
    if (slots < 1) return fail(h, BGP_ERR_NOMEM, "one dense work item at N=%d needs %zu bytes of workspace", N, SL.total);
    rc = ensure_ws(h, (size_t)slots * SL.total);
    if (rc != BGP_OK) return rc;
    return dense_run(h, count, N, D, t, Y, item_gene, b, prior, hyp, kernel_kind, model, kl_weight, logPhi, KConst, want_grad, elbo,
                     grad_hyp, grad_logPhi, status, (char*)h->ws, slots, (cudaStream_t)cuda_stream, true, 0);
}

// ---- host-pointer variant ---------------------------------------------------------------------------------------
}  // extern "C"
int ensure_stage(bgp_handle* h, size_t bytes) {
    if (h->stage_bytes >= bytes) return BGP_OK;
    if (h->stage) { cudaFree(h->stage); h->stage = nullptr; h->stage_bytes = 0; }
    cudaError_t e = cudaMalloc(&h->stage, bytes);
    if (e != cudaSuccess) { cudaGetLastError(); return fail(h, BGP_ERR_NOMEM, "staging of %zu bytes: %s", bytes, cudaGetErrorString(e)); }
    h->stage_bytes = bytes;
    return BGP_OK;
}
extern "C" {

// Host-pointer entry.  The items are cut into groups of a quarter of the SMs; each group owns a lane (CUDA stream +
// workspace slots + staging buffers) on which its H2D copy, kernels and D2H copy are queued in order.  Up to four groups
// compute concurrently (their one-CTA-per-item kernels share the SMs) while the copy engines move the next / previous
// groups, so for calls with more than one SM-count of items the PCIe time hides behind the arithmetic.
int bgp_dense_elbo_grad_host(bgp_handle* h, int count, int N, int D, const double* t, const double* Y, int nY,
                             const int* item_gene, const double* b, const double* prior, const double* hyp,
                             int kernel_kind, int model, double kl_weight, const double* logPhi, const double* KConst,
                             int want_grad, double* elbo, double* grad_hyp, double* grad_logPhi, int* status) {
    if (!h) return BGP_ERR_ARG;
    int rc = dense_check_args(h, count, N, D, nY, t, Y, item_gene, b, prior, hyp, kernel_kind, model, logPhi, want_grad, elbo, grad_hyp,
                              grad_logPhi, status);
    if (rc != BGP_OK) return rc;
    CU(cudaSetDevice(h->device));
    drop_fit_state(h);
    const size_t M = 3 * (size_t)N;
    const int pcols = model == BGP_MODEL_MBGP ? 3 : 2;
    const SlotLayout SL = slot_layout(make_dims(N, D));
    const size_t per_item = (size_t)N * M * 8;
    int conc = 4;   // measured on B200 with 32 hardware work queues (tools/e2e_sweep.sh, 296 items at N = 1000): 2-6 groups 311-315 items/s, 8 groups 293-306, 10 groups 295
    int sm_cap = h->max_slots > 0 ? h->max_slots : h->sm_count;
    if (sm_cap > count) sm_cap = count;
    if (const char* e = getenv("BGP_HOST_CONC")) conc = atoi(e) > 0 ? atoi(e) : conc;
    int gs = sm_cap / conc;                              // items per group: conc groups never ask for more CTAs than there are SMs
    if (gs < 1) gs = 1;
    const bool few = 2 * count <= h->sm_count;           // few items: one group, its macro tiles dealt to all SMs (bgp_dense_wide.cu)
    if (few) gs = count;
    int ngroups = (count + gs - 1) / gs;
    int lanes = ngroups < 2 * conc ? ngroups : 2 * conc;
    if (const char* e = getenv("BGP_HOST_LANES")) { const int l_ = atoi(e); if (l_ > 0) lanes = l_ < ngroups ? l_ : ngroups; }
    // per-lane staging: per-item inputs, outputs, logPhi and gradient
    size_t lo = 0;
    auto ltake = [&](size_t bytes) { size_t r = lo; lo = align_up(lo + bytes, 256); return r; };
    size_t l_gene, l_b, l_hyp, l_elbo, l_gh, l_st, l_lp, l_g;
    auto lane_layout = [&]() {
        lo = 0;
        l_gene = ltake((size_t)gs * 4); l_b = ltake((size_t)gs * 8); l_hyp = ltake((size_t)gs * 24);
        l_elbo = ltake((size_t)gs * 8); l_gh = ltake((size_t)gs * 32); l_st = ltake((size_t)gs * 4);
        l_lp = ltake((size_t)gs * per_item); l_g = ltake(want_grad ? (size_t)gs * per_item : 0);
    };
    lane_layout();
    size_t free_b = 0, total_b = 0;
    CU(cudaMemGetInfo(&free_b, &total_b));
    const size_t avail = (size_t)(0.85 * (double)(free_b + h->ws_bytes + h->stage_bytes));
    while (lanes > 1 && (size_t)lanes * ((size_t)gs * SL.total + lo) > avail) --lanes;
    while (gs > 1 && (size_t)lanes * ((size_t)gs * SL.total + lo) > avail) { gs = (gs + 1) / 2; lane_layout(); }
    if ((size_t)lanes * ((size_t)gs * SL.total + lo) > avail) return fail(h, BGP_ERR_NOMEM, "one dense work item at N=%d does not fit in device memory", N);
    ngroups = (count + gs - 1) / gs;
    rc = ensure_ws(h, (size_t)lanes * gs * SL.total);
    if (rc != BGP_OK) return rc;
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t r = o; o = align_up(o + bytes, 256); return r; };
    const size_t o_t = take((size_t)N * 8), o_Y = take((size_t)nY * N * D * 8), o_prior = take((size_t)N * pcols * 8);
    const size_t o_K = take(KConst ? M * M * 8 : 0);
    const size_t o_lanes = take((size_t)lanes * lo);
    rc = ensure_stage(h, o);
    if (rc != BGP_OK) return rc;
    rc = ensure_lane_streams(h, lanes);
    if (rc != BGP_OK) return rc;
    // small results land in pinned scratch (an async copy into pageable memory would block the host until the group is done)
    rc = ensure_pin(h, (size_t)count * 48 + (size_t)count * 40);
    if (rc != BGP_OK) return rc;
    double* pin_elbo = (double*)h->pin;
    double* pin_gh = pin_elbo + count;
    int* pin_st = (int*)(pin_gh + (size_t)4 * count);
    // ... and so do the small per-item inputs (an async copy from pageable memory synchronises with the stream first)
    double* pin_b = (double*)((char*)h->pin + (size_t)count * 48);
    double* pin_hyp = pin_b + count;
    int* pin_gene = (int*)(pin_hyp + (size_t)3 * count);
    memcpy(pin_b, b, (size_t)count * 8);
    memcpy(pin_hyp, hyp, (size_t)count * 24);
    memcpy(pin_gene, item_gene, (size_t)count * 4);
    char* s = (char*)h->stage;
    cudaStream_t s0 = h->lane_streams[0];
    CU(cudaMemcpyAsync(s + o_t, t, (size_t)N * 8, cudaMemcpyHostToDevice, s0));
    CU(cudaMemcpyAsync(s + o_Y, Y, (size_t)nY * N * D * 8, cudaMemcpyHostToDevice, s0));
    CU(cudaMemcpyAsync(s + o_prior, prior, (size_t)N * pcols * 8, cudaMemcpyHostToDevice, s0));
    if (KConst) CU(cudaMemcpyAsync(s + o_K, KConst, M * M * 8, cudaMemcpyHostToDevice, s0));
    CU(cudaEventRecord(h->shared_ready, s0));
    for (int l = 1; l < lanes; ++l) CU(cudaStreamWaitEvent(h->lane_streams[l], h->shared_ready, 0));
    // a failing call inside the loop must not return before the lanes have drained (copies into the caller's arrays may be in flight)
    auto run = [&]() -> int {
    for (int g = 0; g < ngroups; ++g) {
        const int l = g % lanes;
        cudaStream_t st = h->lane_streams[l];
        const int i0 = g * gs;
        const int n = (count - i0 < gs) ? (count - i0) : gs;
        char* ls = s + o_lanes + (size_t)l * lo;
        CU(cudaMemcpyAsync(ls + l_gene, pin_gene + i0, (size_t)n * 4, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(ls + l_b, pin_b + i0, (size_t)n * 8, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(ls + l_hyp, pin_hyp + (size_t)i0 * 3, (size_t)n * 24, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(ls + l_lp, logPhi + (size_t)i0 * N * M, (size_t)n * per_item, cudaMemcpyHostToDevice, st));
        const int r2 = dense_run(h, n, N, D, (double*)(s + o_t), (double*)(s + o_Y), (int*)(ls + l_gene), (double*)(ls + l_b), (double*)(s + o_prior),
                       (double*)(ls + l_hyp), kernel_kind, model, kl_weight, (double*)(ls + l_lp), KConst ? (double*)(s + o_K) : nullptr,
                       want_grad, (double*)(ls + l_elbo), (double*)(ls + l_gh), want_grad ? (double*)(ls + l_g) : nullptr,
                       (int*)(ls + l_st), (char*)h->ws + (size_t)l * gs * SL.total, gs, st, false, i0, few);
        if (r2 != BGP_OK) return r2;
        CU(cudaMemcpyAsync(pin_elbo + i0, ls + l_elbo, (size_t)n * 8, cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(pin_st + i0, ls + l_st, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
        if (want_grad) {
            CU(cudaMemcpyAsync(pin_gh + (size_t)i0 * 4, ls + l_gh, (size_t)n * 32, cudaMemcpyDeviceToHost, st));
            CU(cudaMemcpyAsync(grad_logPhi + (size_t)i0 * N * M, ls + l_g, (size_t)n * per_item, cudaMemcpyDeviceToHost, st));
        }
    }
    return BGP_OK;
    };
    rc = run();
    for (int l = 0; l < lanes; ++l) {
        cudaError_t e_ = cudaStreamSynchronize(h->lane_streams[l]);
        if (e_ != cudaSuccess && rc == BGP_OK) rc = fail(h, BGP_ERR_CUDA, "cudaStreamSynchronize: %s", cudaGetErrorString(e_));
    }
    if (rc == BGP_OK) {
        memcpy(elbo, pin_elbo, (size_t)count * 8);
        memcpy(status, pin_st, (size_t)count * 4);
        if (want_grad) memcpy(grad_hyp, pin_gh, (size_t)count * 32);
    }
    return rc;
}

// ---- inducing-point bounds --------------------------------------------------------------------------------------
int bgp_sparse_elbo_grad(bgp_handle* h, int count, int N, int Dtot, int Mz, const double* t, const double* Z, const double* Y,
                         int nY, const int* item_gene, const double* b, const double* prior, const double* hyp,
                         int kernel_kind, int model, int multi, const double* logPhi, int want_grad, double* elbo,
                         double* grad_hyp, double* grad_b, double* grad_logPhi, int* status, void* cuda_stream) {
    if (!h) return BGP_ERR_ARG;
    if (count <= 0 || N <= 0 || Dtot <= 0 || nY <= 0 || Mz <= 0) return fail(h, BGP_ERR_ARG, "count, N, Dtot, nY, Mz must be positive");
    if (Mz > SP_MAX_MZ) return fail(h, BGP_ERR_ARG, "Mz = %d exceeds the supported maximum of %d inducing rows", Mz, SP_MAX_MZ);
    if (!t || !Z || !Y || !item_gene || !b || !prior || !hyp || !logPhi || !elbo || !status) return fail(h, BGP_ERR_ARG, "null input pointer");
    if (want_grad && (!grad_hyp || !grad_logPhi)) return fail(h, BGP_ERR_ARG, "want_grad needs grad_hyp and grad_logPhi");
    if (kernel_kind != BGP_KERNEL_MATERN32 && kernel_kind != BGP_KERNEL_SQEXP) return fail(h, BGP_ERR_ARG, "unknown kernel kind %d", kernel_kind);
    CU(cudaSetDevice(h->device));
    drop_fit_state(h);
    cudaStream_t st = (cudaStream_t)cuda_stream;
    const int SD = multi ? Dtot : 1, D = multi ? 1 : Dtot;
    const int M = 3 * N, Mp = (M + SP_SL - 1) / SP_SL * SP_SL;
    // rows per chunk of the assignment-pass partials: the streaming passes (long rows, bound + gradient) take larger chunks
    const int rch = phi_streaming(N, M, SD, D, want_grad != 0, PHI_SRCH) ? PHI_SRCH : PHI_RCH;
    const int nch = (N + rch - 1) / rch;
    const int nslab = Mp / SP_SL;
    // items per wave: bounded by memory; column splits so that a wave fills the GPU about twice
    auto slot_bytes_for = [&](int nsplit) { return (size_t)sparse_layout(N, Mp, Mz, D, nsplit, nch).total * 8; };
    int items_cap = h->max_slots > 0 ? h->max_slots : h->sm_count;
    if (items_cap > count) items_cap = count;
    // column splits per sub-problem: the slab kernels launch items x SD x nsplit CTAs of equal length, so what matters is how
    // close that count comes to a whole number of rounds over the SMs (400 sub-problems unsplit = 2.7 rounds, i.e. three
    // rounds with the last one two thirds empty; split in four = 10.8 rounds).  Rounds per split, plus a small charge per
    // split for its fixed work (constants, zero-initialised partials, the reduction over splits), is minimised.
    int nsplit = 1;
    {
        const long long subs = (long long)items_cap * SD;
        double best = 1e300;
        for (int n = 1; n <= 16 && n <= nslab; ++n) {
            const long long rounds = (subs * n + h->sm_count - 1) / h->sm_count;
            const double cost = (double)rounds / n + 0.015 * n * (double)subs / h->sm_count;
            if (cost < best - 1e-12) { best = cost; nsplit = n; }
        }
        if (const char* e = getenv("BGP_SP_NSPLIT")) { const int n = atoi(e); if (n >= 1 && n <= 16 && n <= nslab) nsplit = n; }
    }
    size_t free_b = 0, total_b = 0;
    CU(cudaMemGetInfo(&free_b, &total_b));
    const size_t avail = (size_t)(0.9 * (double)(free_b + h->ws_bytes));
    const size_t per_item = slot_bytes_for(nsplit) * SD + (size_t)SD * 64;
    if ((size_t)items_cap * per_item > avail) items_cap = (int)(avail / per_item);
    if (items_cap < 1) return fail(h, BGP_ERR_NOMEM, "one sparse work item at N=%d, Mz=%d needs %zu bytes of workspace", N, Mz, per_item);
    const SparseLayout SL = sparse_layout(N, Mp, Mz, D, nsplit, nch);
    const size_t sub_out_off = align_up((size_t)items_cap * SD * SL.total * 8, 256);
    int rc = ensure_ws(h, sub_out_off + (size_t)items_cap * SD * 64);
    if (rc != BGP_OK) return rc;
    double* wsd = (double*)h->ws;
    double* sub_out = (double*)((char*)h->ws + sub_out_off);

    SparseParams P;
    memset(&P, 0, sizeof(P));
    P.N = N; P.M = M; P.Mp = Mp; P.Mz = Mz; P.D = D; P.nch = nch; P.nsplit = nsplit; P.SD = SD; P.Dtot = Dtot;
    P.kind = kernel_kind; P.model = model; P.want_grad = want_grad; P.square_noise = 1e-6; P.kl_weight = 1.0;
    P.slot_stride = SL.total;
    P.lay = SL;
    P.t = t; P.Z = Z; P.hyp = hyp; P.bv = b; P.Y = Y; P.item_gene = item_gene;
    P.dbg_skip = 0;
    if (const char* e = getenv("BGP_SP_SKIP")) P.dbg_skip = atoi(e);
    P.ws = wsd; P.status = status; P.elbo = elbo; P.grad_hyp = want_grad ? grad_hyp : nullptr; P.grad_b = want_grad ? grad_b : nullptr;

    PhiParams F;
    memset(&F, 0, sizeof(F));
    F.N = N; F.M = M; F.Mp = Mp; F.D = D; F.nch = nch; F.rch = rch; F.model = model; F.kl_weight = 1.0; F.SD = SD; F.Dtot = Dtot;
    F.slot_stride = SL.total;
    F.logPhi = logPhi; F.t = t; F.prior = prior; F.bv = b; F.Y = Y; F.item_gene = item_gene;
    F.lse = wsd + SL.lse; F.Apart = wsd + SL.Apart; F.vpart = wsd + SL.vpart; F.klpart = wsd + SL.klpart; F.klrow = wsd + SL.klrow;
    F.rowpart = wsd + SL.rowpart;
    F.Abar = wsd + SL.Abar; F.vbar = wsd + SL.vbar; F.grad_logPhi = grad_logPhi;

    for (int item0 = 0; item0 < count; item0 += items_cap) {
        const int n = (count - item0 < items_cap) ? (count - item0) : items_cap;
        const int nslots = n * SD;
        P.item0 = item0; F.item0 = item0;
        // profiling slots reuse the dense phase ids: 0 phi forward, 1 prep, 2 Kuu, 3 forward slabs, 4 mid, 5 backward slabs,
        // 6 end + reduce, 11 phi backward
        prof_mark(h, -1, st);
        phi_launch_forward(F, nslots, st);
        prof_mark(h, 0, st);
        for (int phase = 0; phase <= 5; ++phase) {
            if (phase == 4 && !want_grad) continue;
            sparse_launch_phase(P, nslots, phase, sub_out, st);
            if (phase == 5 || phase == 0) h->launches++;
            prof_mark(h, phase + 1, st);
        }
        if (want_grad) { phi_launch_backward(F, n, st); prof_mark(h, 11, st); }
        CU(cudaGetLastError());
    }
    h->lastS = P;
    return BGP_OK;
}

// Lane streams of the host-pointer entry points; earlier lanes get higher priority so that groups complete roughly in order
// (their D2H then overlaps later groups).
static int ensure_lane_streams(bgp_handle* h, int lanes) {
    while ((int)h->lane_streams.size() < lanes) {
        int lo_p = 0, hi_p = 0;
        CU(cudaDeviceGetStreamPriorityRange(&lo_p, &hi_p));   // hi_p is numerically the smallest
        int prio = hi_p + (int)h->lane_streams.size();
        if (prio > lo_p) prio = lo_p;
        cudaStream_t s_;
        CU(cudaStreamCreateWithPriority(&s_, cudaStreamNonBlocking, prio));
        h->lane_streams.push_back(s_);
    }
    if (!h->shared_ready) CU(cudaEventCreateWithFlags(&h->shared_ready, cudaEventDisableTiming));
    return BGP_OK;
}
static int ensure_pin(bgp_handle* h, size_t bytes) {
    if (h->pin_bytes >= bytes) return BGP_OK;
    if (h->pin) cudaFreeHost(h->pin);
    h->pin = nullptr; h->pin_bytes = 0;
    CU(cudaHostAlloc(&h->pin, bytes, cudaHostAllocDefault));
    h->pin_bytes = bytes;
    return BGP_OK;
}

// Host-pointer entry of the inducing-point bounds.  The items are cut into groups that pass through three stages on three
// streams -- H2D of the group's logPhi, the kernels, D2H of its gradient -- with two staging buffers, so the copies of
// neighbouring groups run in both PCIe directions at once and hide the (much shorter) arithmetic: at C3 (N = 10 000) an item
// moves 2.4 GB each way for 3 ms of kernels, i.e. the call is bound by one PCIe direction instead of the sum of both.
int bgp_sparse_elbo_grad_host(bgp_handle* h, int count, int N, int Dtot, int Mz, const double* t, const double* Z,
                              const double* Y, int nY, const int* item_gene, const double* b, const double* prior,
                              const double* hyp, int kernel_kind, int model, int multi, const double* logPhi, int want_grad,
                              double* elbo, double* grad_hyp, double* grad_b, double* grad_logPhi, int* status) {
    if (!h) return BGP_ERR_ARG;
    if (count <= 0 || N <= 0 || Dtot <= 0 || nY <= 0 || Mz <= 0) return fail(h, BGP_ERR_ARG, "count, N, Dtot, nY, Mz must be positive");
    if (!t || !Z || !Y || !item_gene || !b || !prior || !hyp || !logPhi || !elbo || !status) return fail(h, BGP_ERR_ARG, "null input pointer");
    if (want_grad && (!grad_hyp || !grad_logPhi)) return fail(h, BGP_ERR_ARG, "want_grad needs grad_hyp and grad_logPhi");
    CU(cudaSetDevice(h->device));
    const size_t M = 3 * (size_t)N;
    const int pcols = model == BGP_MODEL_MBGP ? 3 : 2;
    const int SD = multi ? Dtot : 1;
    const size_t per_item = (size_t)N * M * 8;
    size_t free_b = 0, total_b = 0;
    CU(cudaMemGetInfo(&free_b, &total_b));
    const size_t avail = free_b + h->ws_bytes + h->stage_bytes;
    int cap = h->max_slots > 0 ? h->max_slots : h->sm_count;
    if (cap > count) cap = count;
    // half of the memory for staging (two buffers of logPhi + gradient), the rest for the workspace
    const size_t stage_budget = (size_t)(0.45 * (double)avail);
    if ((size_t)cap * 2 * per_item > stage_budget) cap = (int)(stage_budget / (2 * per_item));
    if (cap < 1) return fail(h, BGP_ERR_NOMEM, "one sparse work item at N=%d does not fit in device memory", N);
    // large transfers: at least four groups so that H2D, kernels and D2H of neighbouring groups overlap
    if (count > 1 && (size_t)count * per_item > ((size_t)256 << 20)) {
        const int want = count < 4 ? count : 4;
        const int c2 = (count + want - 1) / want;
        if (c2 < cap) cap = c2;
    }
    int ngroups = (count + cap - 1) / cap;
    if (ngroups > 1) {   // several groups: two staging buffers, groups of equal size
        if ((size_t)cap * 4 * per_item > stage_budget) cap = (int)(stage_budget / (4 * per_item));
        if (cap < 1) cap = 1;
        ngroups = (count + cap - 1) / cap;
        cap = (count + ngroups - 1) / ngroups;
    }
    const int nbuf = ngroups > 1 ? 2 : 1;
    const int saved_slots = h->max_slots;
    size_t lo = 0;
    auto ltake = [&](size_t bytes) { size_t r = lo; lo = align_up(lo + bytes, 256); return r; };
    const size_t l_gene = ltake((size_t)cap * 4), l_b = ltake((size_t)cap * SD * 8), l_hyp = ltake((size_t)cap * 24);
    const size_t l_elbo = ltake((size_t)cap * 8), l_gh = ltake((size_t)cap * 24), l_gb = ltake((size_t)cap * SD * 8), l_st = ltake((size_t)cap * 4);
    const size_t l_lp = ltake((size_t)cap * per_item), l_g = ltake(want_grad ? (size_t)cap * per_item : 0);
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t r = o; o = align_up(o + bytes, 256); return r; };
    const size_t o_t = take((size_t)N * 8), o_Z = take((size_t)Mz * 16), o_Y = take((size_t)nY * N * Dtot * 8);
    const size_t o_prior = take((size_t)N * pcols * 8);
    const size_t o_lanes = take((size_t)nbuf * lo);
    int rc = ensure_stage(h, o);
    if (rc != BGP_OK) return rc;
    rc = ensure_lane_streams(h, 3);
    if (rc != BGP_OK) return rc;
    while (h->sp_events.size() < 6) {
        cudaEvent_t e;
        CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        h->sp_events.push_back(e);
    }
    // small per-item inputs / results travel through pinned scratch (async copies from / to pageable memory would serialise the lanes)
    const size_t n_out = (size_t)count * (8 + 24 + (size_t)SD * 8 + 4), n_in = (size_t)count * ((size_t)SD * 8 + 24 + 4);
    rc = ensure_pin(h, align_up(n_out, 64) + n_in + 64);
    if (rc != BGP_OK) return rc;
    double* pin_elbo = (double*)h->pin;
    double* pin_gh = pin_elbo + count;
    double* pin_gb = pin_gh + (size_t)3 * count;
    int* pin_st = (int*)(pin_gb + (size_t)SD * count);
    double* pin_b = (double*)((char*)h->pin + align_up(n_out, 64));
    double* pin_hyp = pin_b + (size_t)SD * count;
    int* pin_gene = (int*)(pin_hyp + (size_t)3 * count);
    memcpy(pin_b, b, (size_t)count * SD * 8);
    memcpy(pin_hyp, hyp, (size_t)count * 24);
    memcpy(pin_gene, item_gene, (size_t)count * 4);
    char* s = (char*)h->stage;
    cudaStream_t s_in = h->lane_streams[0], s_cmp = h->lane_streams[1], s_out = h->lane_streams[2];
    CU(cudaMemcpyAsync(s + o_t, t, (size_t)N * 8, cudaMemcpyHostToDevice, s_in));
    CU(cudaMemcpyAsync(s + o_Z, Z, (size_t)Mz * 16, cudaMemcpyHostToDevice, s_in));
    CU(cudaMemcpyAsync(s + o_Y, Y, (size_t)nY * N * Dtot * 8, cudaMemcpyHostToDevice, s_in));
    CU(cudaMemcpyAsync(s + o_prior, prior, (size_t)N * pcols * 8, cudaMemcpyHostToDevice, s_in));
    h->max_slots = cap;
    auto run = [&]() -> int {
        for (int g = 0; g < ngroups; ++g) {
            const int buf = g % nbuf;
            cudaEvent_t in_ev = h->sp_events[buf], done_ev = h->sp_events[2 + buf], free_ev = h->sp_events[4 + buf];
            const int i0 = g * cap;
            const int n = (count - i0 < cap) ? (count - i0) : cap;
            char* ls = s + o_lanes + (size_t)buf * lo;
            if (g >= nbuf) CU(cudaStreamWaitEvent(s_in, free_ev, 0));   // the buffer's previous gradient is home
            CU(cudaMemcpyAsync(ls + l_gene, pin_gene + i0, (size_t)n * 4, cudaMemcpyHostToDevice, s_in));
            CU(cudaMemcpyAsync(ls + l_b, pin_b + (size_t)i0 * SD, (size_t)n * SD * 8, cudaMemcpyHostToDevice, s_in));
            CU(cudaMemcpyAsync(ls + l_hyp, pin_hyp + (size_t)i0 * 3, (size_t)n * 24, cudaMemcpyHostToDevice, s_in));
            CU(cudaMemcpyAsync(ls + l_lp, logPhi + (size_t)i0 * N * M, (size_t)n * per_item, cudaMemcpyHostToDevice, s_in));
            CU(cudaEventRecord(in_ev, s_in));
            CU(cudaStreamWaitEvent(s_cmp, in_ev, 0));
            int r2 = bgp_sparse_elbo_grad(h, n, N, Dtot, Mz, (double*)(s + o_t), (double*)(s + o_Z), (double*)(s + o_Y), nY, (int*)(ls + l_gene),
                                          (double*)(ls + l_b), (double*)(s + o_prior), (double*)(ls + l_hyp), kernel_kind, model, multi,
                                          (double*)(ls + l_lp), want_grad, (double*)(ls + l_elbo), (double*)(ls + l_gh), (double*)(ls + l_gb),
                                          want_grad ? (double*)(ls + l_g) : nullptr, (int*)(ls + l_st), s_cmp);
            if (r2 != BGP_OK) return r2;
            CU(cudaEventRecord(done_ev, s_cmp));
            CU(cudaStreamWaitEvent(s_out, done_ev, 0));
            CU(cudaMemcpyAsync(pin_elbo + i0, ls + l_elbo, (size_t)n * 8, cudaMemcpyDeviceToHost, s_out));
            CU(cudaMemcpyAsync(pin_st + i0, ls + l_st, (size_t)n * 4, cudaMemcpyDeviceToHost, s_out));
            if (want_grad) {
                CU(cudaMemcpyAsync(pin_gh + (size_t)i0 * 3, ls + l_gh, (size_t)n * 24, cudaMemcpyDeviceToHost, s_out));
                CU(cudaMemcpyAsync(pin_gb + (size_t)i0 * SD, ls + l_gb, (size_t)n * SD * 8, cudaMemcpyDeviceToHost, s_out));
                CU(cudaMemcpyAsync(grad_logPhi + (size_t)i0 * N * M, ls + l_g, (size_t)n * per_item, cudaMemcpyDeviceToHost, s_out));
            }
            CU(cudaEventRecord(free_ev, s_out));
        }
        return BGP_OK;
    };
    rc = run();
    // common epilogue, also after an error: nothing may still be writing into the caller's arrays when this function returns
    for (cudaStream_t st : {s_in, s_cmp, s_out}) {
        cudaError_t e_ = cudaStreamSynchronize(st);
        if (e_ != cudaSuccess && rc == BGP_OK) rc = fail(h, BGP_ERR_CUDA, "cudaStreamSynchronize: %s", cudaGetErrorString(e_));
    }
    h->max_slots = saved_slots;
    if (rc == BGP_OK) {
        memcpy(elbo, pin_elbo, (size_t)count * 8);
        memcpy(status, pin_st, (size_t)count * 4);
        if (want_grad) {
            memcpy(grad_hyp, pin_gh, (size_t)count * 24);
            if (grad_b) memcpy(grad_b, pin_gb, (size_t)count * SD * 8);
        }
    }
    return rc;
}

// ---- predict_f ----------------------------------------------------------------------------------------------------
int bgp_dense_predict_host(bgp_handle* h, int count, int N, int D, const double* t, const double* Y, int nY, const int* item_gene,
                           const double* b, const double* prior, const double* hyp, int kernel_kind, int model,
                           const double* logPhi, int T, const double* Xnew, int full_cov, double* mean, double* var) {
    if (!h) return BGP_ERR_ARG;
    if (T <= 0 || !Xnew || !mean || !var) return fail(h, BGP_ERR_ARG, "predict needs T > 0, Xnew, mean and var");
    CU(cudaSetDevice(h->device));
    const size_t vlen = full_cov ? (size_t)T * T : (size_t)T;
    // one grow-only device buffer: test inputs, mean, variance and (full covariance over more than 128 inputs) the tall matrices
    // of every item (the host entry below runs several item groups concurrently, so every item has its own)
    const DenseDims dd = make_dims(N, D);
    const bool tiled = full_cov && T > MB;
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t r = o; o = align_up(o + bytes, 256); return r; };
    const size_t o_X = take((size_t)T * 16), o_mean = take((size_t)count * T * D * 8), o_var = take((size_t)count * vlen * 8);
    const size_t o_ext = take(tiled ? dense_predict_cov_scratch_doubles(dd, count, T) * 8 : 0);
    if (h->pred_buf_bytes < o) {
        if (h->pred_buf) cudaFree(h->pred_buf);
        h->pred_buf = nullptr; h->pred_buf_bytes = 0;
        cudaError_t e = cudaMalloc(&h->pred_buf, o);
        if (e != cudaSuccess) { cudaGetLastError(); return fail(h, BGP_ERR_NOMEM, "predict buffers of %zu bytes: %s", o, cudaGetErrorString(e)); }
        h->pred_buf_bytes = o;
    }
    char* pb = (char*)h->pred_buf;
    double* dX = (double*)(pb + o_X); double* dmean = (double*)(pb + o_mean); double* dvar = (double*)(pb + o_var);
    CU(cudaMemcpy(dX, Xnew, (size_t)T * 16, cudaMemcpyHostToDevice));
    h->pred_X = dX; h->pred_T = T; h->pred_full = full_cov; h->pred_mean = dmean; h->pred_var = dvar;
    h->pred_ext = tiled ? (double*)(pb + o_ext) : nullptr;
    h->pred_neg1 = nullptr;
    if (tiled) {
        h->pred_neg1 = h->pred_ext + dense_predict_cov_scratch_doubles(dd, count, T) - dd.Mp;
        dense_fill(h->pred_neg1, dd.Mp, -1.0, 0);
        CU(cudaStreamSynchronize(0));
    }
    const int saved = h->stop_phase;
    h->stop_phase = 4;
    std::vector<double> elbo(count);
    std::vector<int> status(count);
    int rc = bgp_dense_elbo_grad_host(h, count, N, D, t, Y, nY, item_gene, b, prior, hyp, kernel_kind, model, 1.0, logPhi, nullptr, 0,
                                      elbo.data(), nullptr, nullptr, status.data());
    h->stop_phase = saved;
    h->pred_T = 0; h->pred_X = nullptr; h->pred_mean = nullptr; h->pred_var = nullptr; h->pred_ext = nullptr;
    if (rc == BGP_OK) {
        cudaError_t e1 = cudaMemcpy(mean, dmean, (size_t)count * T * D * 8, cudaMemcpyDeviceToHost);
        cudaError_t e2 = cudaMemcpy(var, dvar, (size_t)count * vlen * 8, cudaMemcpyDeviceToHost);
        if (e1 != cudaSuccess || e2 != cudaSuccess) rc = fail(h, BGP_ERR_CUDA, "predict copy back failed");
        for (int i = 0; i < count && rc == BGP_OK; ++i)
            if (status[i] != 0) rc = fail(h, BGP_ERR_CUDA, "Cholesky decomposition was not successful for item %d (status %d)", i, status[i]);
    }
    return rc;
}

int bgp_sparse_predict_host(bgp_handle* h, int N, int Dtot, int Mz, const double* t, const double* Z, const double* Y, double b,
                            const double* prior, const double* hyp, int kernel_kind, const double* logPhi, int T,
                            const double* Xnew, int full_cov, double* mean, double* var) {
    if (!h) return BGP_ERR_ARG;
    if (T <= 0 || !Xnew || !mean || !var) return fail(h, BGP_ERR_ARG, "predict needs T > 0, Xnew, mean and var");
    CU(cudaSetDevice(h->device));
    // forward state of the single item: run the bound without gradient, then predict from the workspace of slot 0
    double elbo = 0.0; int status = 0, gene = 0;
    h->keep_sparse = true;
    int rc = bgp_sparse_elbo_grad_host(h, 1, N, Dtot, Mz, t, Z, Y, 1, &gene, &b, prior, hyp, kernel_kind, BGP_MODEL_BGP, 0, logPhi, 0,
                                       &elbo, nullptr, nullptr, nullptr, &status);
    h->keep_sparse = false;
    if (rc != BGP_OK) return rc;
    if (status != 0) return fail(h, BGP_ERR_CUDA, "Cholesky decomposition was not successful (status %d)", status);
    const size_t vlen = full_cov ? (size_t)T * T : (size_t)T;
    double *dX = nullptr, *dscr = nullptr, *dmean = nullptr, *dvar = nullptr;
    CU(cudaMalloc(&dX, (size_t)T * 16));
    CU(cudaMalloc(&dscr, (size_t)3 * Mz * T * 8));
    CU(cudaMalloc(&dmean, (size_t)T * Dtot * 8));
    CU(cudaMalloc(&dvar, vlen * 8));
    CU(cudaMemcpy(dX, Xnew, (size_t)T * 16, cudaMemcpyHostToDevice));
    sparse_launch_predict(h->lastS, dX, T, full_cov, dscr, dmean, dvar, 0);
    cudaError_t e0 = cudaGetLastError();
    cudaError_t e1 = cudaMemcpy(mean, dmean, (size_t)T * Dtot * 8, cudaMemcpyDeviceToHost);
    cudaError_t e2 = cudaMemcpy(var, dvar, vlen * 8, cudaMemcpyDeviceToHost);
    cudaFree(dX); cudaFree(dscr); cudaFree(dmean); cudaFree(dvar);
    if (e0 != cudaSuccess || e1 != cudaSuccess || e2 != cudaSuccess) return fail(h, BGP_ERR_CUDA, "sparse predict failed: %s", cudaGetErrorString(e0 != cudaSuccess ? e0 : (e1 != cudaSuccess ? e1 : e2)));
    return BGP_OK;
}

int bgp_get_phi_host(bgp_handle* h, int count, int N, const double* logPhi, double* phi) {
    if (!h || count <= 0 || N <= 0 || !logPhi || !phi) return BGP_ERR_ARG;
    CU(cudaSetDevice(h->device));
    const size_t per_item = (size_t)N * 3 * N * 8, per_out = (size_t)N * 3 * 8;
    // as many items per trip as a 2 GB staging budget holds (one H2D, one gather launch over all of them, one D2H)
    size_t chunk = ((size_t)2 << 30) / per_item;
    if (chunk < 1) chunk = 1;
    if (chunk > (size_t)count) chunk = (size_t)count;
    const size_t o_out = align_up(chunk * per_item, 256);
    int rc = ensure_stage(h, o_out + chunk * per_out);
    if (rc != BGP_OK) return rc;
    char* s = (char*)h->stage;
    for (size_t i0 = 0; i0 < (size_t)count; i0 += chunk) {
        const size_t n = ((size_t)count - i0 < chunk) ? ((size_t)count - i0) : chunk;
        CU(cudaMemcpyAsync(s, logPhi + i0 * (size_t)N * 3 * N, n * per_item, cudaMemcpyHostToDevice, 0));
        phi_launch_gather((double*)s, (double*)(s + o_out), (int)n, N, 0);
        CU(cudaGetLastError());
        CU(cudaMemcpyAsync(phi + i0 * (size_t)N * 3, s + o_out, n * per_out, cudaMemcpyDeviceToHost, 0));
        CU(cudaStreamSynchronize(0));
    }
    return BGP_OK;
}

// ---- per-kernel timing ------------------------------------------------------------------------------------------
int bgp_profile_enable(bgp_handle* h, int on) {
    if (!h) return BGP_ERR_ARG;
    h->profiling = on != 0;
    return BGP_OK;
}
int bgp_profile_reset(bgp_handle* h) {
    if (!h) return BGP_ERR_ARG;
    for (int i = 0; i < BGP_NPHASES; ++i) h->prof_ms[i] = 0.0;
    h->launches = 0;
    h->ev_marks.clear();
    h->ev_used = 0;
    return BGP_OK;
}
int bgp_profile_read(bgp_handle* h, double* ms_out, long long* launches) {
    if (!h) return BGP_ERR_ARG;
    CU(cudaSetDevice(h->device));
    if (!h->ev_marks.empty()) {
        CU(cudaEventSynchronize(h->ev_pool[h->ev_marks.back().second]));
        for (size_t i = 1; i < h->ev_marks.size(); ++i) {
            const int phase = h->ev_marks[i].first;
            if (phase < 0) continue;   // start-of-wave marker
            float ms = 0.f;
            CU(cudaEventElapsedTime(&ms, h->ev_pool[h->ev_marks[i - 1].second], h->ev_pool[h->ev_marks[i].second]));
            h->prof_ms[phase] += ms;
        }
        h->ev_marks.clear();
        h->ev_used = 0;
    }
    if (ms_out) for (int i = 0; i < BGP_NPHASES; ++i) ms_out[i] = h->prof_ms[i];
    if (launches) *launches = h->launches;
    return BGP_OK;
}

// ---- test hooks -------------------------------------------------------------------------------------------------
int bgp_debug_set_stop_phase(bgp_handle* h, int phase) {
    if (!h) return BGP_ERR_ARG;
    h->stop_phase = phase;
    return BGP_OK;
}
int bgp_debug_padded_size(int N) { return make_dims(N, 1).Mp; }

int bgp_debug_dense_matrix(bgp_handle* h, int slot, int which, int symmetric, double* out_host) {
    if (!h || !out_host || slot < 0 || slot >= h->last_slots) return BGP_ERR_ARG;
    CU(cudaSetDevice(h->device));
    const DenseParams& P = h->lastP;
    const double* basep = which == 0 ? P.LC : which == 1 ? P.RX : which == 2 ? P.PB : P.LB;
    const double* src = basep + (size_t)slot * P.slot_stride;
    const size_t n = (size_t)P.d.Mp * P.d.Mp;
    double* tmp = nullptr;
    CU(cudaMalloc(&tmp, n * 8));
    dense_unpack(src, tmp, P.d.Mp, symmetric, 0);
    cudaError_t e = cudaMemcpy(out_host, tmp, n * 8, cudaMemcpyDeviceToHost);
    cudaFree(tmp);
    if (e != cudaSuccess) return fail(h, BGP_ERR_CUDA, "debug copy: %s", cudaGetErrorString(e));
    return BGP_OK;
}

int bgp_debug_gemm_timers(bgp_handle* h, unsigned long long* out8, int reset) {
    if (!h || !out8) return BGP_ERR_ARG;
    CU(cudaSetDevice(h->device));
    CU(cudaDeviceSynchronize());
    dense_read_gemm_timers(out8, reset);
    return BGP_OK;
}

int bgp_debug_wide_timers(bgp_handle* h, unsigned long long* out48, int reset) {
    if (!h || !out48) return BGP_ERR_ARG;
    CU(cudaSetDevice(h->device));
    CU(cudaDeviceSynchronize());
    dense_read_wide_timers(out48, reset);
    return BGP_OK;
}

int bgp_debug_dense_vector(bgp_handle* h, int slot, int which, double* out_host) {
    if (!h || !out_host || slot < 0 || slot >= h->last_slots) return BGP_ERR_ARG;
    CU(cudaSetDevice(h->device));
    const DenseParams& P = h->lastP;
    const VecLayout VL = vec_layout(P.d.Mp, P.d.nM, P.d.D);
    long long off, len = P.d.Mp;
    switch (which) {
        case 0: off = VL.A; break;
        case 1: off = VL.Delta; break;
        case 2: off = VL.v; len *= P.d.D; break;
        case 3: off = VL.z; len *= P.d.D; break;
        case 4: off = VL.c; len *= P.d.D; break;
        case 5: off = VL.q; len *= P.d.D; break;
        case 6: off = VL.vbar; len *= P.d.D; break;
        case 7: off = VL.delta; break;
        case 8: off = VL.Abar; break;
        default: return BGP_ERR_ARG;
    }
    CU(cudaMemcpy(out_host, P.vec + (size_t)slot * P.slot_stride + off, (size_t)len * 8, cudaMemcpyDeviceToHost));
    return BGP_OK;
}

}  // extern "C"
