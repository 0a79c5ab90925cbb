"""GPU: the reference's own timed recipe (notebooks/Hematopoiesis.py:107-118 -- genes MPO and CTSG, every 20th of the 3854 cells, six
candidate branching points, FitModel defaults M = 10, maxiter = 100) on the committed sub-sample of the reference's data
(tests/golden/hematopoiesis_subsample.npz, made by tests/golden/make_golden_hematopoiesis.py).  The notebook stores no numbers besides
the wall time, so what is pinned is internal consistency of the two drivers: the drop-in FitModel (SciPy L-BFGS-B calling the CUDA bound)
and FitModelBatch (device-resident optimiser, both genes at once) must agree on the most likely branching point and on the sign and
rough size of the log Bayes factor."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
Z = np.load(os.path.join(os.path.dirname(__file__), "golden", "hematopoiesis_subsample.npz"))


def test_hematopoiesis_recipe_scipy_and_device_drivers_agree():
    from branchedgp_b200 import FitBranchingModel, VBHelperFunctions
    grid = [float(x) for x in Z["Bsearch"]]
    GPt, labels = Z["GPt"], Z["globalBranching"]
    assert GPt.shape == (193,) and set(np.unique(labels)) <= {1, 2, 3}
    GPY = np.hstack([Z["GPy_MPO"], Z["GPy_CTSG"]])
    batch = FitBranchingModel.FitModelBatch(grid, GPt, GPY, labels)
    for k, g in enumerate(("MPO", "CTSG")):
        d = FitBranchingModel.FitModel(grid, GPt, Z["GPy_" + g], labels)
        assert set(d) == {"loglik", "Phi", "prediction", "hyperparameters", "posteriorB"}
        assert np.all(np.isfinite(d["loglik"])) and np.all(np.isfinite(batch[k]["loglik"]))
        assert d["Phi"].shape == (193, 3) and len(d["prediction"]["mu"]) == 3
        assert int(np.argmax(d["loglik"])) == int(np.argmax(batch[k]["loglik"]))
        bf_s = VBHelperFunctions.CalculateBranchingEvidence(d, grid)["logBayesFactor"]
        bf_d = VBHelperFunctions.CalculateBranchingEvidence(batch[k], grid)["logBayesFactor"]
        assert np.sign(bf_s) == np.sign(bf_d)
        assert abs(bf_s - bf_d) <= 0.1 * max(abs(bf_s), abs(bf_d), 10.0)
