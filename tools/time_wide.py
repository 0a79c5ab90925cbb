"""Latency of bound + gradient evaluations with FEW items in flight: the multi-CTA-per-item pipeline (bgp_dense_wide.cu) against
the one-CTA-per-item pipeline, device-resident inputs, CUDA events, per-phase times from the library's profiler.
usage: python tools/time_wide.py [N=1000] [count=1] [reps=3] [modes=1,0]   (GPU box only)"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from branchedgp_b200 import _lib  # noqa: E402


def problem(N, count, dev):
    g = torch.Generator(device="cpu").manual_seed(0)
    M = 3 * N
    t = torch.sort(torch.rand(N, generator=g, dtype=torch.float64)).values
    Y = torch.randn(8, N, 1, generator=g, dtype=torch.float64)
    prior = torch.full((N, 2), 0.5, dtype=torch.float64)
    item_gene = (torch.arange(count) % 8).to(torch.int32)
    b = torch.linspace(0.05, 0.95, count, dtype=torch.float64) if count > 1 else torch.tensor([0.5], dtype=torch.float64)
    hyp = torch.stack([0.5 + 2.5 * torch.rand(count, generator=g, dtype=torch.float64),
                       1 + 5 * torch.rand(count, generator=g, dtype=torch.float64),
                       0.05 + 0.95 * torch.rand(count, generator=g, dtype=torch.float64)], 1)
    t, Y, prior, item_gene, b, hyp = [x.to(dev).contiguous() for x in (t, Y, prior, item_gene, b, hyp)]
    logPhi = torch.full((count, N, M), -9.0, dtype=torch.float64, device=dev)
    logPhi += 0.1 * torch.randn(count, N, M, dtype=torch.float64, device=dev)
    ar = torch.arange(N, device=dev)
    for f in range(3):
        logPhi[:, ar, 3 * ar + f] = float(np.log(1.0 / 3)) + 0.1 * torch.randn(count, N, dtype=torch.float64, device=dev)
    return t, Y, prior, item_gene, b, hyp, logPhi


def measure(eng, N, count, reps, mode):
    dev = torch.device("cuda:0")
    t, Y, prior, item_gene, b, hyp, logPhi = problem(N, count, dev)
    elbo = torch.empty(count, dtype=torch.float64, device=dev)
    gh = torch.zeros(count, 4, dtype=torch.float64, device=dev)
    gl = torch.empty_like(logPhi)
    status = torch.zeros(count, dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    os.environ["BGP_DENSE_WIDE"] = mode

    def run():
        eng.dense_elbo_grad_device(count, N, 1, t.data_ptr(), Y.data_ptr(), 8, item_gene.data_ptr(), b.data_ptr(), prior.data_ptr(),
                                   hyp.data_ptr(), logPhi.data_ptr(), elbo.data_ptr(), gh.data_ptr(), gl.data_ptr(), status.data_ptr(),
                                   stream=st)
    run()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        run()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    eng.profile_enable(True)
    eng.profile_reset()
    run()
    phases, launches = eng.profile_read()
    eng.profile_enable(False)
    os.environ.pop("BGP_DENSE_WIDE")
    flops = 3.0 * float(3 * N) ** 3 * count
    return dict(mode="wide" if mode == "1" else "narrow", N=N, count=count, ms=round(best, 3), evals_per_s=round(count / best * 1e3, 2),
                tflops=round(flops / best / 1e9, 2), status_nonzero=int((status != 0).sum().item()), elbo0=float(elbo[0].item()),
                grad_hyp0=[float(x) for x in gh[0].tolist()], phases_ms={k: round(v, 3) for k, v in phases.items() if v > 0}, launches=launches)


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
    count = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    modes = sys.argv[4].split(",") if len(sys.argv) > 4 else ["1", "0"]
    eng = _lib.Engine(0)
    for mode in modes:
        print(json.dumps(measure(eng, N, count, reps, mode)), flush=True)


if __name__ == "__main__":
    main()
