"""Latency of ONE bound + gradient evaluation through the host-array entry points (what FitModel's SciPy driver pays per function
evaluation) and of the Python layers above it.  usage: python tools/time_latency.py [N=100] [reps=200]"""
import cProfile
import io
import os
import pstats
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from branchedgp_b200 import _lib  # noqa: E402
from common import make_problem  # noqa: E402


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 200
    eng = _lib.default_engine()
    pr = make_problem(N, D=1, n_genes=1, bs=(0.5,), seed=1)
    args = (pr["t"], pr["Y"], pr["item_gene"], pr["b"], pr["prior"], pr["hyp"], pr["logPhi"])
    for _ in range(5):
        eng.dense_elbo_grad(*args)
    t0 = time.perf_counter()
    for _ in range(reps):
        eng.dense_elbo_grad(*args)
    dense_ms = (time.perf_counter() - t0) / reps * 1e3
    Z = np.ones((10, 2))
    Z[:, 0] = np.linspace(0, 1, 10, endpoint=False)
    Z[:, 1] = np.array([i for j in range(10) for i in range(1, 4)])[:10]
    sargs = (pr["t"], Z, pr["Y"], pr["item_gene"], pr["b"], pr["prior"], pr["hyp"], pr["logPhi"])
    for _ in range(5):
        eng.sparse_elbo_grad(*sargs)
    t0 = time.perf_counter()
    for _ in range(reps):
        eng.sparse_elbo_grad(*sargs)
    sparse_ms = (time.perf_counter() - t0) / reps * 1e3
    print({"N": N, "dense_eval_ms": round(dense_ms, 3), "sparse_eval_ms": round(sparse_ms, 3)})
    # the whole FitModel call under cProfile: where the host time goes
    from golden.make_golden_c1 import MAXITER, c1_problem
    from branchedgp_b200 import FitBranchingModel
    t, Y, lab, grid = c1_problem()
    FitBranchingModel.FitModel(grid[:2], t, Y, lab, M=0, maxiter=5)
    prof = cProfile.Profile()
    prof.enable()
    t0 = time.perf_counter()
    FitBranchingModel.FitModel(grid, t, Y, lab, M=0, maxiter=MAXITER)
    wall = time.perf_counter() - t0
    prof.disable()
    out = io.StringIO()
    pstats.Stats(prof, stream=out).sort_stats("cumulative").print_stats(28)
    print("FitModel C1 wall %.3f s" % wall)
    print(out.getvalue()[:6000])


if __name__ == "__main__":
    main()
