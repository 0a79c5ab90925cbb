"""Generates tests/golden/synthetic20_reference_run.npz from the only outputs of the REAL reference that ship with it:
``/root/reference/notebooks/syntheticdata/syntheticDataRun.p``, written by ``notebooks/runsynteticData.py:27-52`` (for each of
the 40 genes of ``notebooks/syntheticdata/synthetic20.csv``: ``FitModel(Bsearch, GPt, GPy, globalBranching, maxiter=20, M=10)``
with ``Bsearch = [0.1, 0.2, 0.3, 0.5, 0.8, 1.1]``, i.e. the inducing-point model of assigngp_denseSparse.py) and the inputs of
that run (150 cells: ``Time``, ``MonocleState``, 40 expression columns whose names carry the true branching time).

The fixture holds inputs + the reference's ``loglik[40, 6]``, ``Phi[40, 150, 3]``, hyper-parameters and ``Bmode``.  The pickle
was produced by an unknown (older) TF / GPflow build after 20 unconverged L-BFGS iterations per grid point, so it pins the
drop-in contract of BASELINE.json (argmax branching point, sign and size of the log Bayes factor) loosely, not 1e-8 parity.
Run from the repo root in the build container (needs /root/reference and pandas): python tests/golden/make_golden_synthetic20.py
"""
import os
import pickle

import numpy as np

REF = "/root/reference/notebooks/syntheticdata"


def main():
    import pandas as pd
    run = pickle.load(open(os.path.join(REF, "syntheticDataRun.p"), "rb"))
    data = pd.read_csv(os.path.join(REF, "synthetic20.csv"), index_col=[0])
    Y = data.iloc[:, 2:]
    models = run["gpmodels"]
    assert len(models) == Y.shape[1] == 40 and run["M"] == 10 and run["maxiter"] == 20
    out = dict(
        t=data["Time"].values.astype(np.float64),
        labels=data["MonocleState"].values.astype(np.int64),
        Y=Y.values.astype(np.float64),
        true_b=np.array([float(c[-3:]) for c in Y.columns]),
        Bsearch=np.array(run["Bsearch"], dtype=np.float64),
        M=np.int64(run["M"]), maxiter=np.int64(run["maxiter"]),
        ref_loglik=np.stack([np.asarray(m["loglik"], dtype=np.float64) for m in models]),
        ref_Phi=np.stack([np.asarray(m["Phi"], dtype=np.float64) for m in models]),
        ref_hyp=np.array([[float(m["hyperparameters"][k][0]) for k in ("kerlen", "kervar", "likvar")] for m in models]),
        ref_Bmode=np.array([float(m["posteriorB"]["Bmode"]) for m in models]),
    )
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "synthetic20_reference_run.npz"), **out)
    print({k: getattr(v, "shape", v) for k, v in out.items()})


if __name__ == "__main__":
    main()
