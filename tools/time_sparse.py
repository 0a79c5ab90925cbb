"""Throughput of the inducing-point paths at the BASELINE shapes C3 (sparse BGP) and C4 (MBGP) with device-resident inputs."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from branchedgp_b200 import _lib  # noqa: E402


def run(name, N, Dtot, Mz, count, multi, reps=3, eng=None, verbose=True, e2e_items=0):
    """Times `count` work items of one inducing-point configuration with device-resident inputs; returns the result line as a dict
    (phases_ms = CUDA-event time per pipeline phase of one call)."""
    dev = torch.device("cuda", torch.cuda.current_device())
    eng = eng if eng is not None else _lib.Engine(torch.cuda.current_device())
    lib, h = eng._lib, eng._h
    g = torch.Generator().manual_seed(0)
    M = 3 * N
    t = torch.sort(torch.rand(N, generator=g, dtype=torch.float64)).values.to(dev)
    nz = Mz // 3 if multi else Mz
    if multi:
        zt = torch.linspace(0, 1, nz + 1, dtype=torch.float64)[:-1]
        Z = torch.stack([zt.repeat_interleave(3), torch.tensor([1.0, 2.0, 3.0], dtype=torch.float64).repeat(nz)], 1).to(dev)
    else:
        zt = torch.linspace(0, 1, Mz + 1, dtype=torch.float64)[:-1]
        Z = torch.stack([zt, torch.tensor([1.0, 2.0, 3.0], dtype=torch.float64).repeat(Mz)[:Mz]], 1).to(dev)
    Y = torch.randn(1, N, Dtot, generator=g, dtype=torch.float64).to(dev)
    SD = Dtot if multi else 1
    b = (0.1 + 0.8 * torch.rand(count, SD, generator=g, dtype=torch.float64)).to(dev)
    prior = torch.full((N, 3 if multi else 2), 0.5, dtype=torch.float64, device=dev)
    if multi:
        prior[:, 0] = 0.0
    hyp = torch.tensor([[0.5, 1.0, 0.05]], dtype=torch.float64).repeat(count, 1).to(dev)
    gene = torch.zeros(count, dtype=torch.int32, device=dev)
    logPhi = torch.full((count, N, M), -9.0, dtype=torch.float64, device=dev)
    logPhi += 0.1 * torch.randn(count, N, M, dtype=torch.float64, device=dev)
    ar = torch.arange(N, device=dev)
    for f in range(3):
        logPhi[:, ar, 3 * ar + f] = float(np.log(1 / 3))
    elbo = torch.empty(count, dtype=torch.float64, device=dev)
    gh = torch.zeros(count, 3, dtype=torch.float64, device=dev)
    gb = torch.zeros(count, SD, dtype=torch.float64, device=dev)
    gl = torch.empty_like(logPhi)
    st = torch.zeros(count, dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream().cuda_stream

    def call():
        rc = lib.bgp_sparse_elbo_grad(h, count, N, Dtot, Mz, t.data_ptr(), Z.data_ptr(), Y.data_ptr(), 1, gene.data_ptr(), b.data_ptr(),
                                      prior.data_ptr(), hyp.data_ptr(), 1 if multi else 0, 1 if multi else 0, 1 if multi else 0,
                                      logPhi.data_ptr(), 1, elbo.data_ptr(), gh.data_ptr(), gb.data_ptr(), gl.data_ptr(), st.data_ptr(),
                                      stream)
        assert rc == 0, lib.bgp_last_error(h)

    call()
    torch.cuda.synchronize()
    eng.profile_enable(True)
    eng.profile_reset()
    call()
    ph, _ = eng.profile_read()
    eng.profile_enable(False)
    names = dict(phi_forward="phi_fwd", prep="prep", chol_K="kuu", syrk_P="fwd_slabs", chol_P="mid", solve="bwd_slabs", trtri="end",
                 phi_backward="phi_bwd")
    phases = {names[k]: round(v, 3) for k, v in ph.items() if k in names}
    if verbose:
        print(json.dumps(phases))
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        call()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    # roofline figures of SURVEY.md 8(d): algorithmic bytes = logPhi read + gradient write (2 x 8 x 3N^2 per item); algorithmic flops
    # of the inducing-point algebra = D (4 Mz^2 3N + Mz^3) per item.  Peaks: MEASURED_PEAKS.json (HBM copy), profiles/r01_fp64_peak.json.
    bytes_alg = 2.0 * 8 * N * M * count
    flops_alg = float(Dtot) * (4.0 * Mz * Mz * M + float(Mz) ** 3) * count
    hbm_peak, dmma_peak = 6545.6, 37.14
    try:
        hbm_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    gbps = bytes_alg / best / 1e6
    tflops = flops_alg / best / 1e9
    out = {"config": name, "N": N, "D": Dtot, "Mz": Mz, "items": count, "ms": round(best, 3),
           "items_per_s": round(count / best * 1e3, 2), "logPhi_GBps": round(gbps, 1),
           "frac_hbm_peak_algorithmic": round(gbps / hbm_peak, 4), "algebra_TFLOPs": round(tflops, 3),
           "frac_dmma_peak": round(tflops / dmma_peak, 4),
           "status_nonzero": int((st != 0).sum().item()), "elbo0": float(elbo[0].item())}
    if e2e_items > 0:
        # end to end: the same evaluations through the host-pointer entry (bgp_sparse_elbo_grad_host), logits and gradient in
        # pinned host memory, host<->device copies inside the timed region
        import time
        ne = min(e2e_items, count)
        h_lp = _lib.pinned_empty((ne, N, M))
        h_lp[:] = logPhi[:ne].cpu().numpy()
        h_g = _lib.pinned_empty((ne, N, M))
        hb = b[:ne].cpu().numpy()
        hargs = (t.cpu().numpy(), Z.cpu().numpy(), Y.cpu().numpy(), np.zeros(ne, dtype=np.int32), hb if multi else hb[:, 0],
                 prior.cpu().numpy(), hyp[:ne].cpu().numpy(), h_lp)
        kw = dict(kernel=1 if multi else 0, model=1 if multi else 0, multi=multi, grad_out=h_g)
        del gl
        torch.cuda.empty_cache()
        eng.sparse_elbo_grad(*hargs, **kw)
        t0 = time.perf_counter()
        r = eng.sparse_elbo_grad(*hargs, **kw)
        dt = time.perf_counter() - t0
        out["e2e"] = {"items_per_s": round(ne / dt, 2), "items": ne, "h2d_bytes": int(ne * N * M * 8), "d2h_bytes": int(ne * N * M * 8),
                      "GBps_each_way": round(ne * N * M * 8 / dt / 1e9, 1), "elbo0": float(r["elbo"][0]),
                      "api": "bgp_sparse_elbo_grad_host (ctypes, pinned host buffers)"}
    if verbose:
        print(json.dumps(out))
    out["phases_ms"] = phases
    return out


if __name__ == "__main__":
    # usage: time_sparse.py [C3 items in flight = 8] [which = c3,c4]
    which = sys.argv[2] if len(sys.argv) > 2 else "c3,c4"
    if "c3" in which:
        run("C3 sparse BGP", 10000, 1, 30, int(sys.argv[1]) if len(sys.argv) > 1 else 8, False)
    if "c4" in which:
        run("C4 MBGP", 500, 100, 180, 4, True)
