"""Device-resident batched fit (FitModelBatch) against the SciPy-driven FitModel on the same GPU objective.

Prints, per configuration, the largest deviations of the log-likelihood table, hyper-parameters, Phi and predictions, the
iteration / evaluation counts of both optimisers and the wall time of each.  Usage: python tools/compare_fit.py [N] [genes]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from branchedgp_b200 import FitBranchingModel  # noqa: E402
from common import make_problem  # noqa: E402


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    G = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    pr = make_problem(N, D=1, n_genes=G, bs=(0.4,), seed=3)
    t, labels = pr["t"], pr["labels"].astype(float)
    GPY = pr["Y"][:, :, 0].T.copy()
    grid = [0.2, 0.5, 0.8, 1.1]
    for M in (0, 10):
        for fix in (True, False):
            t0 = time.time()
            batch = FitBranchingModel.FitModelBatch(grid, t, GPY, labels, M=M, maxiter=30, fixHyperparameters=fix, fPredict=True)
            t_batch = time.time() - t0
            t0 = time.time()
            single = [FitBranchingModel.FitModel(grid, t, GPY[:, g:g + 1], labels, M=M, maxiter=30, fixHyperparameters=fix, fPredict=True)
                      for g in range(G)]
            t_single = time.time() - t0
            dev = dict(loglik=0.0, loglik_rel=0.0, phi=0.0, hyp=0.0, mu=0.0, var=0.0)
            same_argmax = True
            for g in range(G):
                a, s = batch[g], single[g]
                dev["loglik"] = max(dev["loglik"], float(np.abs(a["loglik"] - s["loglik"]).max()))
                dev["loglik_rel"] = max(dev["loglik_rel"], float(np.abs((a["loglik"] - s["loglik"]) / s["loglik"]).max()))
                same_argmax &= int(np.argmax(a["loglik"])) == int(np.argmax(s["loglik"]))
                dev["phi"] = max(dev["phi"], float(np.abs(a["Phi"] - s["Phi"]).max()))
                for k in ("likvar", "kerlen", "kervar"):
                    dev["hyp"] = max(dev["hyp"], float(abs(a["hyperparameters"][k] - s["hyperparameters"][k])))
                for f in range(3):
                    dev["mu"] = max(dev["mu"], float(np.abs(a["prediction"]["mu"][f] - s["prediction"]["mu"][f]).max()))
                    dev["var"] = max(dev["var"], float(np.abs(a["prediction"]["var"][f] - s["prediction"]["var"][f]).max()))
            print(json.dumps(dict(N=N, genes=G, M=M, fixHyperparameters=fix, same_argmax=bool(same_argmax), dev=dev,
                                  batch_iters=[batch[g]["fit"]["iters"].tolist() for g in range(G)],
                                  batch_nfev=[batch[g]["fit"]["nfev"].tolist() for g in range(G)],
                                  batch_status=[batch[g]["fit"]["status"].tolist() for g in range(G)],
                                  loglik_batch=batch[0]["loglik"].tolist(), loglik_single=single[0]["loglik"].tolist(),
                                  s_batch=round(t_batch, 2), s_single=round(t_single, 2))), flush=True)


if __name__ == "__main__":
    main()
