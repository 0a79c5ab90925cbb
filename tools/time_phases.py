"""Time the dense pipeline phase by phase at a given N (device-resident inputs, CUDA events). GPU box only."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from branchedgp_b200 import _lib  # noqa: E402


def main(N=1000, count=148, reps=2):
    dev = torch.device("cuda:0")
    eng = _lib.Engine(0)
    g = torch.Generator(device="cpu").manual_seed(0)
    M = 3 * N
    t = torch.sort(torch.rand(N, generator=g, dtype=torch.float64)).values
    Y = torch.randn(8, N, 1, generator=g, dtype=torch.float64)
    prior = torch.full((N, 2), 0.5, dtype=torch.float64)
    item_gene = (torch.arange(count) % 8).to(torch.int32)
    b = torch.linspace(0.05, 0.95, count, dtype=torch.float64)
    hyp = torch.stack([0.5 + 2.5 * torch.rand(count, generator=g, dtype=torch.float64),
                       1 + 5 * torch.rand(count, generator=g, dtype=torch.float64),
                       0.05 + 0.95 * torch.rand(count, generator=g, dtype=torch.float64)], 1)
    t, Y, prior, item_gene, b, hyp = [x.to(dev).contiguous() for x in (t, Y, prior, item_gene, b, hyp)]
    logPhi = torch.full((count, N, M), -9.0, dtype=torch.float64, device=dev)
    logPhi += 0.1 * torch.randn(count, N, M, dtype=torch.float64, device=dev)
    ar = torch.arange(N, device=dev)
    for f in range(3):
        logPhi[:, ar, 3 * ar + f] = float(np.log(1.0 / 3)) + 0.1 * torch.randn(count, N, dtype=torch.float64, device=dev)
    elbo = torch.empty(count, dtype=torch.float64, device=dev)
    gh = torch.zeros(count, 4, dtype=torch.float64, device=dev)
    gl = torch.empty_like(logPhi)
    status = torch.zeros(count, dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream().cuda_stream

    def run():
        eng.dense_elbo_grad_device(count, N, 1, t.data_ptr(), Y.data_ptr(), 8, item_gene.data_ptr(), b.data_ptr(), prior.data_ptr(),
                                   hyp.data_ptr(), logPhi.data_ptr(), elbo.data_ptr(), gh.data_ptr(), gl.data_ptr(), status.data_ptr(),
                                   stream=st)

    names = ["phi_fwd+prep+chol_K", "syrk_P", "chol_P", "solve", "trtri", "lauum", "trmm", "post", "adj+phi_bwd"]
    cum = []
    for phase in list(range(1, 10)):
        eng.debug_set_stop_phase(phase if phase < 9 else 0)
        best = 1e30
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            run()
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        cum.append(best)
    eng.debug_set_stop_phase(0)
    out = {"N": N, "count": count, "status_nonzero": int((status != 0).sum().item()), "total_ms": cum[-1],
           "items_per_s": count / cum[-1] * 1e3}
    prev = 0.0
    M3 = float(3 * N) ** 3
    for n, c in zip(names, cum):
        out[n + "_ms"] = round(c - prev, 3)
        prev = c
    out["tflops_3M3"] = 3 * M3 * count / cum[-1] / 1e9
    print(json.dumps(out))


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 1000, int(sys.argv[2]) if len(sys.argv) > 2 else 148)
