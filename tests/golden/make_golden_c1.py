"""Generates tests/golden/c1_fitmodel.npz: the oracle's FitModel results on BASELINE.json configs[0] ("C1": one synthetic gene,
N = 100 cells, 10-point branching grid, float64, CPU) for the dense model (M = 0) with trainable and with fixed
hyper-parameters.  Run from the repo root: python tests/golden/make_golden_c1.py (about a minute on 8 cores)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

MAXITER = 30


def c1_problem():
    """Seeded C1 inputs (shapes and generator of SURVEY.md 8d): sorted uniform pseudotime, trunk before 0.3, one gene that
    branches at 0.5, grid = linspace(0.05, 0.95, 9) + [1.1]."""
    rng = np.random.default_rng(100)
    N = 100
    t = np.sort(rng.uniform(0.0, 1.0, N))
    labels = np.where(t < 0.3, 1, rng.integers(2, 4, N)).astype(float)
    sign = np.where(labels == 3, -1.0, 1.0)
    y = sign * 4.0 * np.maximum(t - 0.5, 0.0) + 0.25 * rng.standard_normal(N)
    y -= y.mean()
    grid = list(np.linspace(0.05, 0.95, 9)) + [1.1]
    return t, y[:, None], labels, grid


def main():
    from oracle import fit_ref
    t, Y, labels, grid = c1_problem()
    out = {}
    for name, fix in (("train", False), ("fixed", True)):
        r = fit_ref.fit_model(grid, t, Y, labels, M=0, maxiter=MAXITER, fixHyperparameters=fix)
        out[name + "/loglik"] = r["loglik"]
        out[name + "/Phi"] = r["Phi"]
        out[name + "/hyp"] = r["hyperparameters"]
        out[name + "/hyp_all"] = r["hyp_all"]
        print(name, r["loglik"], r["hyperparameters"])
    # the first grid point on its own (trainable hyper-parameters): no warm-start chain has amplified anything yet, so the
    # assignment probabilities there are held to north_star's 1e-6
    r = fit_ref.fit_model(grid[:1], t, Y, labels, M=0, maxiter=MAXITER, fixHyperparameters=False)
    out["train_first/loglik"] = r["loglik"]
    out["train_first/Phi"] = r["Phi"]
    out["train_first/hyp"] = r["hyperparameters"]
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "c1_fitmodel.npz"), **out)


if __name__ == "__main__":
    main()
