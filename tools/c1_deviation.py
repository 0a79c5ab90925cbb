import os, sys, numpy as np
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
from golden.make_golden_c1 import MAXITER, c1_problem
from branchedgp_b200 import FitBranchingModel
GOLD = np.load("/root/repo/tests/golden/c1_fitmodel.npz")
t, Y, labels, grid = c1_problem()
for which, fix in (("train", False), ("fixed", True)):
    for driver in ("FitModel", "FitModelBatch"):
        if driver == "FitModel":
            d = FitBranchingModel.FitModel(grid, t, Y, labels, M=0, maxiter=MAXITER, fixHyperparameters=fix, fPredict=False)
        else:
            d = FitBranchingModel.FitModelBatch(grid, t, Y, labels, M=0, maxiter=MAXITER, fixHyperparameters=fix, fPredict=False)[0]
        ll, want = d["loglik"], GOLD[which + "/loglik"]
        h = d["hyperparameters"]
        print(which, driver, "rel dev loglik", np.abs((ll - want) / want).max(), "argmax", np.argmax(ll), np.argmax(want),
              "Phi dev", np.abs(d["Phi"] - GOLD[which + "/Phi"]).max(),
              "hyp dev", np.abs(np.array([float(h["kerlen"]), float(h["kervar"]), float(h["likvar"])]) / GOLD[which + "/hyp"] - 1).max())
