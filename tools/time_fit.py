"""Throughput of the device-resident batched fit (bgp_dense_fit_host / bgp_sparse_fit_host) on synthetic genes.

usage: python tools/time_fit.py N genes grid_points M maxiter [fix]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from branchedgp_b200 import FitBranchingModel, _lib  # noqa: E402
from common import make_problem  # noqa: E402


def main():
    N, G, B, M, maxiter = (int(x) for x in sys.argv[1:6])
    fix = len(sys.argv) > 6 and sys.argv[6] == "fix"
    os.environ["BGP_FIT_STATS"] = "1"
    pr = make_problem(N, D=1, n_genes=G, bs=(0.4,), seed=3)
    grid = list(np.linspace(0.05, 0.95, B - 1)) + [1.1]
    GPY = pr["Y"][:, :, 0].T.copy()
    eng = _lib.default_engine()
    t0 = time.time()
    res = FitBranchingModel.FitModelBatch(grid, pr["t"], GPY, pr["labels"].astype(float), M=M, maxiter=maxiter, fixHyperparameters=fix,
                                          fPredict=True, engine=eng)
    dt = time.time() - t0
    nfev = int(sum(r["fit"]["nfev"].sum() for r in res))
    iters = int(sum(r["fit"]["iters"].sum() for r in res))
    status = np.concatenate([r["fit"]["status"] for r in res])
    print(json.dumps(dict(N=N, genes=G, grid=B, M=M, maxiter=maxiter, fixHyperparameters=fix, seconds=round(dt, 3), items=G * B,
                          fits_per_s=round(G * B / dt, 3), evaluations=nfev, evals_per_s=round(nfev / dt, 2), iterations=iters,
                          status_counts=np.bincount(status, minlength=6).tolist(),
                          argmax_b=[float(grid[int(np.argmax(r["loglik"]))]) for r in res[:8]])))


if __name__ == "__main__":
    main()
