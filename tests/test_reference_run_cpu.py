"""CPU: the oracle's FitModel restatement against outputs of the REAL reference.

tests/golden/synthetic20_reference_run.npz (made by tests/golden/make_golden_synthetic20.py) holds the inputs and results of the
one run of the reference that ships with it: notebooks/runsynteticData.py:27-52 -> notebooks/syntheticdata/syntheticDataRun.p,
``FitModel(Bsearch, GPt, GPy, globalBranching, maxiter=20, M=10)`` for 40 genes x 6 candidate branching points (N = 150 cells,
inducing-point bound of assigngp_denseSparse.py).  The pickle comes from an older TF / GPflow build and every fit stops
unconverged after 20 L-BFGS iterations, so log-likelihoods differ by O(10) while the quantities the drop-in contract names
(BASELINE.json: argmax branching point, sign of the log Bayes factor) and the size of the Bayes factor are reproduced.
This is the only pin the reference offers; 1e-8 parity of the bound itself stays "oracle = restatement" (DESIGN.md)."""
import os
import warnings

import numpy as np
import pytest

from oracle import fit_ref

RUN = np.load(os.path.join(os.path.dirname(__file__), "golden", "synthetic20_reference_run.npz"))
GRID = [float(b) for b in RUN["Bsearch"]]
# genes of the three true-branching-time groups (0.1 / 0.8 / 1.1 = no branching) whose reference fits are well behaved
CPU_GENES = [0, 6, 8, 11, 20, 24, 31, 33, 37]


def log_bayes_factor(loglik):
    """VBHelperFunctions.py:18-34 (CalculateBranchingEvidence) as one logsumexp."""
    o = np.asarray(loglik, dtype=float)
    m = o[:-1].max()
    return m + np.log(np.exp(o[:-1] - m).sum()) - np.log(o.size - 1) - o[-1]


def well_behaved(ref_ll):
    """The reference run itself diverged for three genes (|loglik| of 4e4 ... 2e5 at one grid point)."""
    return bool(np.all(np.abs(ref_ll) < 1e3))


def check_against_reference(ll, ref_ll, strict):
    bf, bf_ref = log_bayes_factor(ll), log_bayes_factor(ref_ll)
    same_b = int(np.argmax(ll)) == int(np.argmax(ref_ll))
    if strict:      # clear-cut genes: |log Bayes factor| of the reference above 20
        assert same_b
        assert np.sign(bf) == np.sign(bf_ref)
        assert abs(bf - bf_ref) <= 0.10 * abs(bf_ref) + 2.0
    return same_b, np.sign(bf) == np.sign(bf_ref)


def test_fixture_is_the_reference_run():
    assert RUN["ref_loglik"].shape == (40, 6) and RUN["Y"].shape == (150, 40) and int(RUN["M"]) == 10 and int(RUN["maxiter"]) == 20
    assert sum(well_behaved(r) for r in RUN["ref_loglik"]) == 37
    # the reference finds the true branching time (column name) for the early / late branching genes
    ok = [GRID[int(np.argmax(r))] == b for r, b in zip(RUN["ref_loglik"], RUN["true_b"]) if well_behaved(r) and b < 1.0]
    assert np.mean(ok) == 1.0


@pytest.fixture(scope="module")
def oracle_fits():
    """The oracle's FitModel on every gene whose reference fit is well behaved (37 fits of 6 grid points, ~40 s on 8 cores)."""
    fits = {}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for g in range(40):
            if well_behaved(RUN["ref_loglik"][g]):
                fits[g] = fit_ref.fit_model(GRID, RUN["t"], RUN["Y"][:, g:g + 1], RUN["labels"], M=10, maxiter=20)
    return fits


def test_oracle_fit_reproduces_reference_run(oracle_fits):
    same_b = same_sign = 0
    for g, out in oracle_fits.items():
        ref_ll, ll = RUN["ref_loglik"][g], out["loglik"]
        assert np.all(np.isfinite(ll)), g
        sb, ss = check_against_reference(ll, ref_ll, strict=abs(log_bayes_factor(ref_ll)) > 20.0)
        same_b, same_sign = same_b + sb, same_sign + ss
    # every clear-cut gene is reproduced (asserted above); of the borderline ones (no true branching, |log Bayes factor| < 20)
    # one flips its argmax between the last two grid points
    assert len(oracle_fits) == 37 and same_b >= 36 and same_sign >= 36, (same_b, same_sign)


@pytest.mark.parametrize("g", CPU_GENES)
def test_oracle_assignment_agrees_with_reference_posterior(oracle_fits, g):
    """Phi of the best model: the branch assignment of the cells beyond the branching point matches the reference's."""
    ref_ll, out = RUN["ref_loglik"][g], oracle_fits[g]
    ib = int(np.argmax(ref_ll))
    assert int(np.argmax(out["loglik"])) == ib
    if GRID[ib] < 1.0:
        late = RUN["t"] > GRID[ib] + 0.05
        agree = np.argmax(out["Phi"][late], axis=1) == np.argmax(RUN["ref_Phi"][g][late], axis=1)
        assert agree.mean() >= 0.9
