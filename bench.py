#!/usr/bin/env python
"""Benchmark of the BranchedGP hot path on B200: ELBO + gradient evaluations per second (gene x branch-point work
items) for the dense AssignGP bound at N = 1000 cells (BASELINE.json configs[1], "C2").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One JSON line on stdout (rank 0).  A step is one pass of the hot path (bound + full gradient) over one batch of
`--items` distinct work items of the C2 grid (512 genes x 20 branching points); every rank owns its own batch (weak
scaling, no data-path collective) and the per-item ELBOs / hyper-gradients are gathered over NCCL at the end of a
step.  See DESIGN.md section "Measurement" for how each field is produced.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

# bgp_dense_elbo_grad_host keeps up to 16 item groups in flight on their own streams (must be set before CUDA initialises)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
# stdout carries exactly one JSON line: NCCL's version / debug banner goes to stderr
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")

import numpy as np  # noqa: E402

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_GENES, N_B = 512, 20
METRIC = "elbo_grad_evals_per_s"
UNIT = "items/s"


def workload_name(N):
    return (f"C2: BGP dense AssignGP bound+grad, N={N} cells (M=3N={3 * N}), grid of {N_GENES} genes x {N_B} branching points; "
            "a step = one batch of distinct (gene, branching point) work items")


def c2_grid():
    return np.concatenate([np.linspace(0.05, 0.95, N_B - 1), [1.1]])   # last entry = "no branching" null (VBHelperFunctions.py:13)


def phase_flops(N):
    """Algorithmic fp64 flops per work item of each GEMM-shaped kernel, M = 3N."""
    M3 = float(3 * N) ** 3
    return {"chol_K": M3 / 3, "syrk_P": M3 / 3, "chol_P": M3 / 3, "trtri": M3 / 3, "lauum": M3 / 3, "trmm": 2 * M3 / 3,
            "adjoint": 2 * M3 / 3}


class ClockSampler(threading.Thread):
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self._stop_evt = index, [], threading.Event()

    def run(self):
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self._stop_evt.wait(0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=6)
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows)
        reasons = []
        for k, name in enumerate(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]):
            if any(r[3 + k].lower().startswith("active") for r in self.rows):
                reasons.append(name)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.rows[0][1]), "power_w_max": max(float(r[2]) for r in self.rows),
                "samples": len(self.rows), "reasons": reasons}


def fp64_peak(device_index=0, torch=None, skip=False):
    """FP64 roofline denominator.  MEASURED_PEAKS.json (driver-written) carries only HBM and bf16, so the fp64 figures are
    measured HERE, in this run, before the timed region: tools/microbench/fp64_peak (compiled by __graft_entry__.build():
    DFMA and DMMA m8n8k4 issue rates over all SMs) and a cuBLAS DGEMM 4096^3 through torch.matmul, with the SM clock sampled
    while they run.  The committed round-1 figures are the fallback when the binary is missing."""
    out = {"dmma_tflops": 37.14, "cublas_dgemm_tflops": 35.44, "measured_this_run": None,
           "source": "fallback: profiles/r01_fp64_peak.json (DMMA m8n8k4 issue rate), profiles/r01_lib_peaks.json (cuBLAS DGEMM 8192^3); "
                     "MEASURED_PEAKS.json has no fp64 entry"}
    try:
        with open(os.path.join(ROOT, "profiles", "r01_fp64_peak.json")) as f:
            out["dmma_tflops"] = float(json.load(f)["dmma_tflops_bps8"])
        with open(os.path.join(ROOT, "profiles", "r01_lib_peaks.json")) as f:
            out["cublas_dgemm_tflops"] = float(json.load(f)["dgemm_8192_tflops"])
    except Exception:
        pass
    exe = os.path.join(ROOT, "tools", "microbench", "fp64_peak")
    if skip or not os.path.exists(exe):
        return out
    try:
        sampler = ClockSampler(device_index)
        sampler.start()
        vis = [v for v in os.environ.get("CUDA_VISIBLE_DEVICES", "").split(",") if v]
        env = dict(os.environ, CUDA_VISIBLE_DEVICES=vis[device_index] if device_index < len(vis) else str(device_index))
        res = subprocess.run([exe], capture_output=True, text=True, timeout=120, env=env)
        mb = json.loads(res.stdout.strip().splitlines()[-1])
        dgemm = None
        if torch is not None:
            a = torch.randn(4096, 4096, dtype=torch.float64, device="cuda")
            b = torch.randn(4096, 4096, dtype=torch.float64, device="cuda")
            torch.matmul(a, b)
            best = 1e30
            for _ in range(5):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                torch.matmul(a, b)
                e1.record()
                torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
            dgemm = 2.0 * 4096 ** 3 / (best * 1e-3) / 1e12
            del a, b
        clk = sampler.stop()
        out["measured_this_run"] = {"dmma_tflops": mb["dmma_tflops_bps8"], "dfma_tflops": mb["dfma_tflops_bps8"],
                                    "cublas_dgemm_4096_tflops": None if dgemm is None else round(dgemm, 2),
                                    "sm_mhz": clk.get("sm_mhz"), "reasons": clk.get("reasons"),
                                    "how": "tools/microbench/fp64_peak (mma.sync.m8n8k4.f64 / fma.rn.f64 issue rate, 8 CTAs of 256 threads per SM, "
                                           "best of 5) and torch.matmul fp64 4096^3 (best of 5), run inside bench.py before the timed region"}
        out["dmma_tflops"] = float(mb["dmma_tflops_bps8"])
        if dgemm is not None:
            out["cublas_dgemm_tflops"] = round(dgemm, 2)
        out["source"] = "measured in this run (roofline.peak_measured_this_run)"
    except Exception as ex:
        out["measured_this_run"] = {"error": repr(ex)}
    return out


# ----------------------------------------------------------------------------------------------------------------------
def synth_host(N, seed=0):
    """Item-independent inputs of SURVEY.md 8d: pseudotime, labels, FitModel's prior / initial phi, Y for all genes."""
    from branchedgp_b200.FitBranchingModel import GetInitialConditionsAndPrior
    rng = np.random.default_rng(seed)
    t = np.sort(rng.uniform(0.0, 1.0, N))
    labels = np.where(t < 0.3, 1, rng.integers(2, 4, N))
    phi0, prior = GetInitialConditionsAndPrior(labels, 0.8, True)
    bg = rng.uniform(0.1, 0.9, N_GENES)
    ag = rng.uniform(0.5, 3.0, N_GENES)
    sign = np.where(labels == 3, -1.0, 1.0)
    Y = sign[None, :] * ag[:, None] * np.maximum(t[None, :] - bg[:, None], 0.0) + 0.3 * rng.standard_normal((N_GENES, N))
    Y -= Y.mean(axis=1, keepdims=True)
    hyp = np.stack([rng.uniform(0.5, 3.0, N_GENES * N_B), rng.uniform(1.0, 6.0, N_GENES * N_B),
                    rng.uniform(0.05, 1.0, N_GENES * N_B)], axis=1)
    return dict(t=t, labels=labels, phi0=phi0, prior=prior, Y=Y[:, :, None].copy(), hyp=hyp)


def item_arrays(first, n):
    """(gene index, branching point) of n consecutive items of the C2 grid starting at `first` (wrapping)."""
    idx = (first + np.arange(n)) % (N_GENES * N_B)
    grid = c2_grid()
    return idx, (idx // N_B).astype(np.int32), grid[idx % N_B]


def init_logphi_device(torch, dev, N, phi0, t, b_items, seed):
    """InitialiseVariationalPhi (assigngp_dense.py:92-128) for every item, plus 0.1 N(0,1) so items are distinct."""
    n = len(b_items)
    g = torch.Generator(device=dev).manual_seed(seed)
    lp = torch.full((n, N, 3 * N), -9.0, dtype=torch.float64, device=dev)
    eps = 1e-9
    tt = torch.as_tensor(t, device=dev)
    bb = torch.as_tensor(b_items, device=dev)
    act = tt[None, :] > bb[:, None]                                  # [n, N]
    p0 = torch.as_tensor(phi0, device=dev)
    branch = torch.log(torch.stack([torch.full((N,), eps, dtype=torch.float64, device=dev), p0[:, 0] - eps, p0[:, 1] - eps], 1))
    trunk = torch.log(torch.tensor([1 - 2 * eps, eps, eps], dtype=torch.float64, device=dev))
    blk = torch.where(act[:, :, None], branch[None], trunk[None, None, :])   # [n, N, 3]
    ar = torch.arange(N, device=dev)
    for f in range(3):
        lp[:, ar, 3 * ar + f] = blk[:, :, f]
    lp += 0.1 * torch.randn(lp.shape, generator=g, dtype=torch.float64, device=dev)
    return lp


def bind_to_gpu_numa_node(gpu_index):
    """Multi-rank runs: keep this rank's host threads (and with first-touch its pinned buffers) on the CPUs next to its GPU, so
    that the host<->device copies of the e2e leg do not cross the socket interconnect."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * w + b for w, m in enumerate(words) for b in range(64) if (m >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
    except Exception as ex:   # affinity is an optimisation only
        print(f"[bench] no NUMA binding for GPU {gpu_index}: {ex}", file=sys.stderr)



def c5_side(eng, torch, dist, world, rank, dev, stream, peak_info, N=2000, items=148):
    """One step of 148 items per GPU at the C5 shape, device-resident inputs; returns the side line."""
    H5 = synth_host(N, seed=5)
    d5 = {k: torch.as_tensor(H5[k], device=dev).contiguous() for k in ("t", "prior", "Y")}
    first = rank * items
    idx, gene, bb = item_arrays(first, items)
    logPhi = init_logphi_device(torch, dev, N, H5["phi0"], H5["t"], bb, seed=77 + rank)
    grad = torch.empty_like(logPhi)
    elbo = torch.empty(items, dtype=torch.float64, device=dev)
    gh = torch.zeros(items, 4, dtype=torch.float64, device=dev)
    status = torch.zeros(items, dtype=torch.int32, device=dev)
    gene_d, b_d = torch.as_tensor(gene, device=dev), torch.as_tensor(bb, device=dev)
    hyp_d = torch.as_tensor(H5["hyp"][idx], device=dev).contiguous()
    res_local = torch.empty(items, 5, dtype=torch.float64, device=dev)
    res_all = torch.empty(world * items, 5, dtype=torch.float64, device=dev) if world > 1 else None

    def step():
        eng.dense_elbo_grad_device(items, N, 1, d5["t"].data_ptr(), d5["Y"].data_ptr(), N_GENES, gene_d.data_ptr(), b_d.data_ptr(),
                                   d5["prior"].data_ptr(), hyp_d.data_ptr(), logPhi.data_ptr(), elbo.data_ptr(), gh.data_ptr(),
                                   grad.data_ptr(), status.data_ptr(), stream=stream)
        if world > 1:
            res_local[:, 0] = elbo
            res_local[:, 1:] = gh
            dist.all_gather_into_tensor(res_all, res_local)

    step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    step()
    e1.record()
    torch.cuda.synchronize()
    tms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms = float(tms.item())
    flops = 3.0 * float(3 * N) ** 3 * items * world
    tf = flops / (ms * 1e-3) / 1e12
    return {"config": f"C5 shape: N={N} cells (M={3 * N}), {items} (gene, branching point) items per GPU, 1 warm-up + 1 timed step",
            "value": round(world * items / (ms * 1e-3), 3), "unit": UNIT, "n_gpus": world, "ms_per_step": round(ms, 2),
            "tflops_total": round(tf, 2), "frac_dmma_peak_per_gpu": round(tf / world / peak_info["dmma_tflops"], 4),
            "status_nonzero": int((status != 0).sum().item()),
            "genome_wide_sweep_minutes": round(20000 * 40 / (world * items / (ms * 1e-3)) / 60.0, 1)}


def single_gene_fitmodel(eng, H, N=1000, grid=(0.3, 0.6, 1.1), maxiter=5):
    """The reference's own call pattern at the headline size: FitModel on ONE gene (dense model, M = 0), one bound + gradient
    evaluation at a time -- every evaluation has the whole GPU (multi-CTA-per-item pipeline, bgp_dense_wide.cu)."""
    from branchedgp_b200 import FitBranchingModel
    t, labels = H["t"], H["labels"].astype(float)
    Y = H["Y"][3]                                  # [N, 1]
    out = {"config": f"FitModel(M=0), 1 gene, N={N}, {len(grid)} grid points, maxiter={maxiter}, trainable hyper-parameters, fPredict=True"}
    for name, kw in (("FitModel_scipy_driver", {}), ("FitModel_device_optimiser", {"optimizer": "device"})):
        t0 = time.perf_counter()
        r = FitBranchingModel.FitModel(list(grid), t, Y, labels, M=0, maxiter=maxiter, **kw)
        out[name] = {"seconds": round(time.perf_counter() - t0, 3), "argmax_b": float(grid[int(np.argmax(r["loglik"]))])}
    # latency of one evaluation through the host-array entry (what the SciPy driver pays per function evaluation)
    from branchedgp_b200 import _lib
    lp = _lib.pinned_empty((1, N, 3 * N))
    lp[:] = -9.0
    gp = _lib.pinned_empty((1, N, 3 * N))
    args = (t, Y[None], np.zeros(1, dtype=np.int32), np.array([0.5]), H["prior"], np.array([[2.0, 3.0, 0.1]]), lp)
    eng.dense_elbo_grad(*args, grad_out=gp)
    t0 = time.perf_counter()
    for _ in range(5):
        eng.dense_elbo_grad(*args, grad_out=gp)
    out["one_evaluation_ms_host_arrays_pinned"] = round((time.perf_counter() - t0) / 5 * 1e3, 2)
    return out


def hematopoiesis_recipe():
    """The reference's own timed recipe (notebooks/Hematopoiesis.py:107-118: genes MPO and CTSG, every 20th of 3854 cells = 193
    cells, 6 candidate branching points, FitModel defaults M = 10 inducing points, maxiter = 100, fPredict) on the committed
    sub-sample (tests/golden/hematopoiesis_subsample.npz), wall seconds beside the notebook's stored 58.0 s / 64.9 s."""
    from branchedgp_b200 import FitBranchingModel
    z = np.load(os.path.join(ROOT, "tests", "golden", "hematopoiesis_subsample.npz"))
    grid = [float(x) for x in z["Bsearch"]]
    GPt, labels = z["GPt"], z["globalBranching"]
    out = {"config": "notebooks/Hematopoiesis.py FitGene(g, ns=20): N=193 cells, 6 grid points, M=10, maxiter=100",
           "reference_notebook_seconds": {"MPO": 58.0, "CTSG": 64.9, "hardware": "unnamed CPU (notebooks/Hematopoiesis.ipynb:435,498)"}}
    FitBranchingModel.FitModel(grid[:2], GPt, z["GPy_MPO"], labels, maxiter=2)   # warm-up (library workspace)
    for g in ("MPO", "CTSG"):
        t0 = time.perf_counter()
        r = FitBranchingModel.FitModel(grid, GPt, z["GPy_" + g], labels)
        out[g] = {"FitModel_seconds": round(time.perf_counter() - t0, 3), "argmax_b": grid[int(np.argmax(r["loglik"]))]}
    GPY = np.hstack([z["GPy_MPO"], z["GPy_CTSG"]])
    t0 = time.perf_counter()
    rb = FitBranchingModel.FitModelBatch(grid, GPt, GPY, labels)
    out["FitModelBatch_both_genes_seconds"] = round(time.perf_counter() - t0, 3)
    out["FitModelBatch_argmax_b"] = [grid[int(np.argmax(r["loglik"]))] for r in rb]
    return out


def pcie_probe(torch, dist, world, dev, mb=1024):
    """H2D and D2H bandwidth of this rank's GPU from pinned memory while every rank copies at the same time (GB/s; min / mean
    over ranks).  The e2e leg moves 24 MB per item each way, so on a multi-GPU box it is bounded by what the host's memory system and
    PCIe root complexes deliver to all GPUs at once, not by one link."""
    n = mb * (1 << 20) // 8
    hbuf = torch.empty(n, dtype=torch.float64).pin_memory()
    dbuf = torch.empty(n, dtype=torch.float64, device=dev)
    res = []
    for src, dst in ((hbuf, dbuf), (dbuf, hbuf)):
        dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            dst.copy_(src, non_blocking=True)
        e1.record()
        torch.cuda.synchronize()
        res.append(3 * n * 8 / (e0.elapsed_time(e1) * 1e-3) / 1e9)
    tt = torch.tensor(res, dtype=torch.float64, device=dev)
    lo, mean = tt.clone(), tt.clone()
    if world > 1:
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(mean, op=dist.ReduceOp.SUM)
        mean /= world
    return {"h2d_GBps_min": round(float(lo[0]), 1), "h2d_GBps_mean": round(float(mean[0]), 1), "d2h_GBps_min": round(float(lo[1]), 1),
            "d2h_GBps_mean": round(float(mean[1]), 1), "how": f"{mb} MiB pinned copies, all {world} rank(s) at the same time, one direction at a time"}


def run_ours(args):
    import torch
    import torch.distributed as dist
    from branchedgp_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        bind_to_gpu_numa_node(local)
        dist.init_process_group("nccl", device_id=dev)
    N, ips = args.cells, args.items
    # fp64 peak of this GPU, measured before anything else is resident (rank-local; the line reports rank 0's)
    peak_info = fp64_peak(local, None if args.no_peak else torch, skip=args.no_peak)
    if world > 1:
        dist.barrier()
    eng = _lib.Engine(local)
    H = synth_host(N)
    d = {k: torch.as_tensor(H[k], device=dev).contiguous() for k in ("t", "prior", "Y")}
    hyp_all = torch.as_tensor(H["hyp"], device=dev)
    stream = torch.cuda.current_stream().cuda_stream

    # this rank's batch of items; logPhi (ips x 24 MB) is far larger than the 126 MB L2, so no flush is needed
    first = rank * ips
    _, gene0, b0 = item_arrays(first, ips)
    logPhi = init_logphi_device(torch, dev, N, H["phi0"], H["t"], b0, seed=1234 + rank)
    grad = torch.empty_like(logPhi)
    elbo = torch.empty(ips, dtype=torch.float64, device=dev)
    gh = torch.zeros(ips, 4, dtype=torch.float64, device=dev)
    status = torch.zeros(ips, dtype=torch.int32, device=dev)
    res_local = torch.empty(ips, 5, dtype=torch.float64, device=dev)
    res_all = torch.empty(world * ips, 5, dtype=torch.float64, device=dev) if world > 1 else None

    def step(k):
        idx, gene, bb = item_arrays(first + k * world * ips, ips)
        gene_d = torch.as_tensor(gene, device=dev)
        b_d = torch.as_tensor(bb, device=dev)
        hyp_d = hyp_all[torch.as_tensor(idx, device=dev)].contiguous()
        eng.dense_elbo_grad_device(ips, N, 1, d["t"].data_ptr(), d["Y"].data_ptr(), N_GENES, gene_d.data_ptr(), b_d.data_ptr(),
                                   d["prior"].data_ptr(), hyp_d.data_ptr(), logPhi.data_ptr(), elbo.data_ptr(), gh.data_ptr(),
                                   grad.data_ptr(), status.data_ptr(), stream=stream)
        if world > 1:   # the only collective of the path: gather per-item ELBO + hyper-gradient
            res_local[:, 0] = elbo
            res_local[:, 1:] = gh
            dist.all_gather_into_tensor(res_all, res_local)
        return gene_d, b_d, hyp_d   # keep alive until the stream has consumed them

    keep = []
    for k in range(args.warmup):
        keep.append(step(k))
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    eng.profile_enable(True)
    eng.profile_reset()
    sampler = ClockSampler(local)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for k in range(args.steps):
        keep.append(step(args.warmup + k))
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    clocks = sampler.stop()
    ms = e0.elapsed_time(e1)
    phase_ms, launches = eng.profile_read()
    eng.profile_enable(False)
    bad = int((status != 0).sum().item())
    tms = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms = float(tms.item())
    value = world * ips * args.steps / (ms * 1e-3)

    # ---- roofline of the dominant kernel (rank 0's own launches)
    fl = phase_flops(N)
    dom = max(fl, key=lambda k: phase_ms[k])
    slots = min(ips, eng.sm_count)                       # one workspace slot per SM: a launch covers `slots` items
    waves_per_step = math.ceil(ips / slots)
    launch_ms = phase_ms[dom] / (args.steps * waves_per_step)
    items_per_launch = ips / waves_per_step
    achieved = fl[dom] * items_per_launch / (launch_ms * 1e-3) / 1e12
    pk = peak_info
    gemm_ms = sum(phase_ms[k] for k in fl)
    total_flops = sum(fl.values()) * ips * args.steps
    traffic = None   # dram bytes of the dominant kernel per launch, from the committed ncu --set full capture (148-item launch)
    try:
        summ = os.path.join(ROOT, "profiles", "r02_ncu_summary.json")
        with open(summ if os.path.exists(summ) else os.path.join(ROOT, "profiles", "r01_ncu_summary.json")) as f:
            cap = json.load(f).get({"adjoint": "k_adj", "lauum": "k_lauum"}.get(dom, dom), {})
        if "dram_bytes_read" in cap and abs(items_per_launch - 148.0) < 1e-9 and N == 1000:
            traffic = cap["dram_bytes_read"] + cap["dram_bytes_write"]
    except Exception:
        pass
    roofline = {"bound": "tensor", "kernel": dom, "achieved": round(achieved, 3), "peak": pk["dmma_tflops"], "unit": "TFLOP/s",
                "frac": round(achieved / pk["dmma_tflops"], 4), "traffic": traffic,
                "launch_ms": round(launch_ms, 3), "items_per_launch": items_per_launch,
                "flops_per_item": fl[dom], "peak_source": pk["source"], "peak_measured_this_run": pk["measured_this_run"],
                "cublas_dgemm_tflops": pk["cublas_dgemm_tflops"],
                "whole_step_tflops": round(total_flops / (ms * 1e-3) / 1e12, 3),
                "whole_step_frac": round(total_flops / (ms * 1e-3) / 1e12 / pk["dmma_tflops"], 4),
                "phase_ms_per_step": {k: round(v / args.steps, 3) for k, v in phase_ms.items()},
                "gemm_kernels_share_of_step": round(gemm_ms / ms, 4)}

    # ---- end to end through the host-pointer C ABI (pinned host buffers, copies inside the timed region)
    e2e = None
    if not args.no_e2e:
        ne = min(args.e2e_items, ips)
        h_lp = torch.empty((ne, N, 3 * N), dtype=torch.float64).pin_memory()
        h_lp.copy_(logPhi[:ne])
        idx, gene, bb = item_arrays(first, ne)
        hyp_h = H["hyp"][idx]
        del grad, logPhi
        torch.cuda.empty_cache()
        import ctypes
        gptr = ctypes.c_void_p()
        nbytes = ne * N * 3 * N * 8
        assert eng._lib.bgp_host_alloc(ctypes.byref(gptr), nbytes) == 0
        g_host = np.ctypeslib.as_array(ctypes.cast(gptr, ctypes.POINTER(ctypes.c_double)), shape=(ne, N, 3 * N))
        out = None
        reps = max(1, min(args.steps, 3))
        tbest = []
        for r in range(1 + reps):   # first repetition is a warm-up
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            out = eng.dense_elbo_grad(H["t"], H["Y"], gene, bb, H["prior"], hyp_h, h_lp.numpy(), grad_out=g_host)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            if r > 0:
                tbest.append(dt)
        tt = torch.tensor([sum(tbest) / len(tbest)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e = {"value": round(world * ne / float(tt.item()), 3), "unit": UNIT, "items_per_step": ne,
               "h2d_bytes_per_step": int(nbytes + ne * 40), "d2h_bytes_per_step": int(nbytes + ne * 44),
               "api": "bgp_dense_elbo_grad_host (ctypes, pinned host buffers)", "timer": "host wall clock around the call, max over ranks",
               "elbo_checksum": float(out["elbo"].sum())}
        eng._lib.bgp_host_free(gptr)
        del h_lp
        try:
            e2e["pcie_under_load"] = pcie_probe(torch, dist, world, dev)
        except Exception as exc:
            e2e["pcie_under_load"] = {"error": repr(exc)}

    # ---- the same evaluations inside the device-resident batched L-BFGS (bgp_dense_fit_host): a few iterations of one wave
    fit = None
    if not args.no_fit:
        nf = min(args.fit_items, ips)
        idx, gene, bb = item_arrays(first, nf)
        opts = eng.fit_options(maxiter=args.fit_maxiter, train_hyp=0)
        torch.cuda.empty_cache()   # the optimiser state (0.6 GB per item at N = 1000) is sized from the free device memory
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fo = eng.fit(H["t"], H["Y"], gene, bb, H["prior"], H["phi0"], H["hyp"][idx], options=opts)
        tf = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        nev = torch.tensor([float(fo["nfev"].sum())], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tf, op=dist.ReduceOp.MAX)
            dist.all_reduce(nev, op=dist.ReduceOp.SUM)
        fit = {"api": "bgp_dense_fit_host (device-resident batched L-BFGS, logPhi / gradient / history never leave HBM)",
               "items": int(world * nf), "maxiter": args.fit_maxiter, "evaluations": int(nev.item()), "seconds": round(float(tf.item()), 3),
               "evals_per_s": round(float(nev.item()) / float(tf.item()), 2), "status_counts": np.bincount(fo["status"], minlength=6).tolist(),
               "note": "wall clock of the whole call incl. allocation of the optimiser state and the final Phi / prediction outputs"}

    # ---- side measurements.  The engine of the headline measurement is closed first (it gives back the dense workspace and the
    # optimiser state of the fit, 116 GB).
    del keep
    eng.close()
    torch.cuda.empty_cache()
    eng = _lib.Engine(local)

    # C5 shape (BASELINE.json configs[4]: N = 2000 cells, genome-wide sweep sharded over the GPUs): one wave of 148 items per GPU,
    # timed like the headline (device-resident inputs, CUDA events, max over ranks, NCCL gather of the per-item results)
    c5 = None
    if not args.no_c5:
        try:
            c5 = c5_side(eng, torch, dist, world, rank, dev, stream, peak_info)
        except Exception as exc:
            c5 = {"error": repr(exc)}
        eng.close()
        torch.cuda.empty_cache()
        eng = _lib.Engine(local)

    # the inducing-point configurations of BASELINE.json (C3 sparse BGP, C4 MBGP) and the drop-in FitModel calls, N = 1 only:
    # parity-test cases, reported beside the headline so that their figures come from the same run
    other = None
    if world == 1 and not args.no_sparse:
        try:
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            import time_sparse
            other = {"note": "bound + gradient evaluations/s with inputs resident in HBM (tools/time_sparse.py), `e2e` = the same through the "
                             "host-pointer entry with pinned host buffers (copies inside the timed region); C3: HBM bound, fraction = "
                             "algorithmic logPhi + gradient bytes (16 B per element) over the measured copy bandwidth, its e2e is bound by "
                             "PCIe (2.4 GB per item each way); C4: algebra flops D (4 Mz^2 3N + Mz^3) over the measured DMMA peak",
                     "C3": time_sparse.run("C3 sparse BGP: N=10000, Mz=30", 10000, 1, 30, 8, False, eng=eng, verbose=False, e2e_items=4),
                     "C4": time_sparse.run("C4 MBGP: N=500, D=100, Mz=180", 500, 100, 180, 4, True, eng=eng, verbose=False, e2e_items=4),
                     "C4_single_model": time_sparse.run("C4 MBGP, ONE joint model in flight (the MBGP optimiser's own call pattern): "
                                                        "N=500, D=100, Mz=180", 500, 100, 180, 1, True, eng=eng, verbose=False)}
            # C1 (BASELINE.json configs[0]): the drop-in FitModel call on one synthetic gene, N = 100, 10-point grid -- wall
            # clock of the whole call through the public API, SciPy driver and device-resident batched optimiser
            sys.path.insert(0, os.path.join(ROOT, "tests"))
            from golden.make_golden_c1 import MAXITER, c1_problem
            from branchedgp_b200 import FitBranchingModel
            t1, Y1, lab1, grid1 = c1_problem()
            c1 = {"config": "C1: FitModel, 1 gene, N=100, 10 grid points, maxiter=%d, trainable hyper-parameters" % MAXITER}
            for name, fn in (("FitModel_scipy_driver", lambda: FitBranchingModel.FitModel(grid1, t1, Y1, lab1, M=0, maxiter=MAXITER)),
                             ("FitModelBatch_device_optimiser", lambda: FitBranchingModel.FitModelBatch(grid1, t1, Y1, lab1, M=0, maxiter=MAXITER)[0])):
                fn()   # warm-up (workspace allocation)
                t0 = time.perf_counter()
                r1 = fn()
                c1[name] = {"seconds": round(time.perf_counter() - t0, 3), "argmax_b": round(float(grid1[int(np.argmax(r1["loglik"]))]), 6)}
            other["C1"] = c1
            other["single_gene_N1000"] = single_gene_fitmodel(eng, H)
            other["hematopoiesis"] = hematopoiesis_recipe()
        except Exception as exc:   # never lose the headline line to a side measurement
            other = {"error": repr(exc)} if other is None else dict(other, error=repr(exc))

    line = {"metric": METRIC, "value": round(value, 3), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(ms / args.steps, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(N),
                       "items_per_step_per_gpu": ips, "cells": N, "D": 1, "kernel": "Matern32",
                       "l2": "inputs larger than L2 (logPhi+grad = %.1f GB per step), no flush" % (2 * ips * N * 3 * N * 8 / 1e9),
                       "parallelism": f"items sharded over {world} GPU(s), NCCL all_gather of [items,5] results per step"},
            "e2e": e2e, "fit": fit, "c5_side": c5, "other_configs": other, "gpu_launches": launches, "roofline": roofline, "clocks": clocks, "status_nonzero": bad}
    if rank == 0 and world == 1 and not args.no_cpu:   # reported at N = 1 only
        line["cpu_baseline"] = cpu_baseline(N, H, kind="port", budget_s=args.cpu_seconds)
    if rank == 0:
        emit(line)
    if world > 1:
        dist.destroy_process_group()


# ----------------------------------------------------------------------------------------------------------------------
def _host_problem(N, H, k):
    """Dense numpy inputs of item k for the CPU oracle."""
    from oracle import host_ref
    idx, gene, bb = item_arrays(k, 1)
    b = float(bb[0])
    X, _ = host_ref.function_index_list(H["t"])
    pZ = host_ref.prior_with_trunk_bgp(host_ref.expand_prior_bgp(H["prior"]), b, H["t"])
    rng = np.random.default_rng(1000 + k)
    lp = host_ref.init_logphi_bgp(H["phi0"], H["t"], b) + 0.1 * rng.standard_normal((N, 3 * N))
    return lp, H["Y"][gene[0]], X, pZ, b, H["hyp"][idx[0]]


def _use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1; the CPU arms are meant to use every host core."""
    import torch
    n = os.cpu_count() or 1
    try:
        torch.set_num_threads(n)
    except Exception:
        pass
    try:
        import threadpoolctl
        threadpoolctl.threadpool_limits(limits=n)
    except Exception:
        pass
    return n


def cpu_baseline(N, H, kind, budget_s=20.0, impl="numpy"):
    """Time the CPU oracle (test infrastructure; the one sanctioned use outside tests) on a bounded sample of the workload."""
    import torch
    from oracle import bgp_numpy, bgp_torch
    _use_all_host_threads()
    fn = bgp_numpy.dense_value_and_grad if impl == "numpy" else bgp_torch.dense_value_and_grad
    times, k = [], 0
    t_start = time.perf_counter()
    while True:
        lp, Y, X, pZ, b, hyp = _host_problem(N, H, k)
        t0 = time.perf_counter()
        fn(lp, Y, X, pZ, b, hyp[0], hyp[1], hyp[2])
        times.append(time.perf_counter() - t0)
        k += 1
        if k >= 2 and time.perf_counter() - t_start > budget_s:
            break
    per = float(np.mean(times[1:])) if len(times) > 1 else times[0]
    return {"value": round(1.0 / per, 4), "unit": UNIT, "cores": torch.get_num_threads(), "host_cpus": os.cpu_count(), "kind": kind,
            "sample": f"{len(times) - 1} timed + 1 warm-up work items of the same workload (N={N}), "
                      f"oracle.{'bgp_numpy (LAPACK + hand adjoint)' if impl == 'numpy' else 'bgp_torch (op-for-op restatement + autograd)'}",
            "s_per_item": round(per, 4)}


def run_reference(args):
    """Reference arm: the reference's own formulation of the path (op-for-op float64 restatement of the TF/GPflow graph with
    reverse-mode autodiff, oracle/bgp_torch.py) on the host cores.  TensorFlow/GPflow cannot be installed here (no wheels in
    /opt/wheelhouse, no network), so oracle/_ref does not exist and kind = "port"."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    N = args.cells
    H = synth_host(N)
    from oracle import bgp_torch
    _use_all_host_threads()
    k = 0
    times = []
    per_step = max(1, args.ref_items)
    for s in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        for _ in range(per_step):
            lp, Y, X, pZ, b, hyp = _host_problem(N, H, k)
            bgp_torch.dense_value_and_grad(lp, Y, X, pZ, b, hyp[0], hyp[1], hyp[2])
            k += 1
        if s >= args.warmup:
            times.append(time.perf_counter() - t0)
    tot = sum(times)
    value = per_step * args.steps / tot
    cb = {"value": round(value, 4), "unit": UNIT, "cores": torch.get_num_threads(), "host_cpus": os.cpu_count(), "kind": "port",
          "sample": f"{per_step} work item(s) per step of the same workload (N={N}); oracle.bgp_torch = restatement of the TF 2.4 / "
                    "GPflow 2.1.2 graph + autodiff (reference not installable here)"}
    line = {"impl": "reference", "metric": METRIC, "value": round(value, 4), "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(tot / args.steps * 1e3, 3), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(N), "items_per_step": per_step},
            "cpu_baseline": cb, "e2e": {"value": round(value, 4), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


_REAL_STDOUT = None


def emit(line):
    """The one JSON line of the contract, on the process's original stdout."""
    out = _REAL_STDOUT if _REAL_STDOUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    global _REAL_STDOUT
    # NCCL prints its version banner on fd 1 when NCCL_DEBUG is set in the environment: send fd 1 to stderr for the whole run
    # and keep a private handle on the original stdout for the result line
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cells", type=int, default=1000, help="N (C2 = 1000)")
    ap.add_argument("--items", type=int, default=296, help="work items per step per GPU")
    ap.add_argument("--e2e-items", type=int, default=296)
    ap.add_argument("--ref-items", type=int, default=1)
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-fit", action="store_true")
    ap.add_argument("--no-sparse", action="store_true", help="skip the C3 / C4 / FitModel side measurements")
    ap.add_argument("--no-c5", action="store_true", help="skip the C5-shaped (N = 2000) side line")
    ap.add_argument("--no-peak", action="store_true", help="skip the fp64 peak micro-benchmark (profiler runs)")
    ap.add_argument("--fit-items", type=int, default=148)
    ap.add_argument("--fit-maxiter", type=int, default=3)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
