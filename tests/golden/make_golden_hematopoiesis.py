"""Generates tests/golden/hematopoiesis_subsample.npz: the inputs of the reference's own timed recipe
(/root/reference/notebooks/Hematopoiesis.py:107-118, `FitGene(g, ns=20)`): every 20th of the 3854 cells of
notebooks/singlecelldata/hematoData.csv / hematoMonocle.csv (193 cells), genes MPO and CTSG, mean-removed expression,
Monocle's stretched pseudotime and state labels.  The notebook's stored output reports 58.0 s (MPO) and 64.9 s (CTSG) for
`FitModel(Bsearch, GPt, GPy, globalBranching)` with its defaults (M = 10 inducing points, maxiter = 100) on an unnamed CPU
(notebooks/Hematopoiesis.ipynb:435,498).  bench.py times the same call on these inputs.  Needs pandas; run from the repo root."""
import os

import numpy as np
import pandas as pd

REF = "/root/reference/notebooks/singlecelldata"


def main():
    Y = pd.read_csv(os.path.join(REF, "hematoData.csv"), index_col=[0])
    monocle = pd.read_csv(os.path.join(REF, "hematoMonocle.csv"), index_col=[0])
    ns = 20
    out = {"genes": np.array(["MPO", "CTSG"]), "Bsearch": np.array(list(np.linspace(0.05, 0.95, 5)) + [1.1]),
           "GPt": monocle["StretchedPseudotime"].values[::ns], "globalBranching": monocle["State"].values[::ns].astype(int)}
    for g in out["genes"]:
        v = Y[g].iloc[::ns].values
        out["GPy_" + g] = (v - v.mean())[:, None]
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "hematopoiesis_subsample.npz"), **out)
    print({k: getattr(v, "shape", None) for k, v in out.items()})


if __name__ == "__main__":
    main()
