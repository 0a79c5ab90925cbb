"""GPU: predict_f (dense, sparse, MBGP) against the oracle, and the drop-in drivers end to end with their defaults."""
import numpy as np
import pytest

from branchedgp_b200 import _lib
from common import make_problem, relerr
from oracle import bgp_numpy, host_ref

pytestmark = pytest.mark.gpu


def _xnew(T, seed=0):
    rng = np.random.default_rng(seed)
    return np.stack([np.sort(rng.uniform(0, 1, T)), rng.integers(1, 4, T).astype(float)], 1)


@pytest.mark.parametrize("N,D,T,full", [(20, 1, 100, False), (50, 2, 150, False), (90, 1, 300, False), (45, 1, 60, True),
                                        (45, 1, 150, True), (30, 2, 300, True), (100, 1, 257, True)])
def test_dense_predict_matches_oracle(engine, N, D, T, full):
    pr = make_problem(N, D=D, seed=300 + N)
    Xn = _xnew(T)
    b, h = pr["b"][0], pr["hyp"][0]
    mean, var = engine.dense_predict(pr["t"], pr["Y"], pr["item_gene"], pr["b"], pr["prior"], pr["hyp"], pr["logPhi"], Xn, full_cov=full)
    m0, v0 = bgp_numpy.dense_predict(pr["logPhi"][0], pr["Y"][0], pr["X"], Xn, b, h[0], h[1], h[2], full_cov=full)
    assert relerr(mean[0], m0) <= 1e-6
    assert relerr(var[0], v0) <= 1e-6


def test_sparse_predict_matches_oracle(engine):
    pr = make_problem(60, D=2, seed=7)
    Mz = 21
    Z = np.ones((Mz, 2))
    Z[:, 0] = np.linspace(0, 1, Mz, endpoint=False)
    Z[:, 1] = np.array([i for j in range(Mz) for i in range(1, 4)])[:Mz]
    b, h = pr["b"][0], pr["hyp"][0]
    for full, T in ((False, 200), (True, 50)):
        Xn = _xnew(T, 3)
        mean, var = engine.sparse_predict(pr["t"], Z, pr["Y"][0], b, pr["prior"], h, pr["logPhi"][0], Xn, full_cov=full)
        m0, v0 = bgp_numpy.sparse_predict(pr["logPhi"][0], pr["Y"][0], pr["X"], Z, Xn, b, h[0], h[1], h[2], full_cov=full)
        assert relerr(mean, m0) <= 1e-6
        assert relerr(var, v0) <= 1e-6


def test_fitmodel_with_reference_defaults_sparse_and_prediction():
    """testing/test_objective.py verbatim (FitModel defaults: M = 10 inducing points, fPredict = True)."""
    from branchedgp_b200 import FitBranchingModel
    np.random.seed(43)
    N = 20
    t = np.linspace(0, 1, N)
    trueB = np.ones((1, 1)) * 0.4
    Y = np.zeros((N, 1))
    idx = np.nonzero(t > 0.5)[0]
    idxA, idxB = idx[::2], idx[1::2]
    Y[idxA, 0] = 2 * t[idxA]
    Y[idxB, 0] = -2 * t[idxB]
    labels = np.ones(N)
    istart = np.min(np.flatnonzero(t > 0.5))
    labels[istart::2] = 2
    labels[(istart + 1)::2] = 3
    ptb = np.min([np.min(t[labels == 2]), np.min(t[labels == 3])])
    grid = [ptb, 0.1, 0.4, np.round(ptb, 1), np.round(ptb + 0.01, 2), 0.7, 1.1]
    d = FitBranchingModel.FitModel(grid, t, Y, labels, maxiter=50, priorConfidence=0.65, kervar=1.0, kerlen=1.0, likvar=0.01)
    iw = np.argmax(d["loglik"])
    Bmode = grid[iw]
    assignment = np.ones(N)
    assignment[d["Phi"][:, 1] > 0.5] = 2
    assignment[d["Phi"][:, 1] < 0.5] = 3
    assignment[d["Phi"][:, 0] > 0.99] = 1
    alt = np.ones(N)
    alt[d["Phi"][:, 1] > 0.5] = 3
    alt[d["Phi"][:, 1] < 0.5] = 2
    alt[d["Phi"][:, 0] > 0.99] = 1
    truth = np.ones(N)
    truth[idxA] = 2
    truth[idxB] = 3
    acc = np.max([(truth == assignment).sum(), (truth == alt).sum()])
    assert Bmode == trueB, "Picked B=%f" % Bmode
    assert acc >= N - 2, "Got accuracy of %g" % acc
    iobjSort = np.argsort(d["loglik"])
    assert iobjSort.size == 7
    assert np.all(iobjSort == np.array([6, 5, 4, 0, 3, 1, 2])), (iobjSort, d["loglik"])
    assert len(d["prediction"]["mu"]) == 3 and d["prediction"]["mu"][0].shape == (100, 1)
    assert np.all(np.asarray(d["prediction"]["var"][0]) > 0)


@pytest.mark.parametrize("D", [1, 2])
def test_mbgp_equivalence_multi_to_single(D):
    """testing/MBGP/test_equivalence_multi_to_single.py: dense single-b and inducing-point multi-b paths agree."""
    from branchedgp_b200 import gpflow_shim as gpflow
    from branchedgp_b200.MBGP import VBHelperFunctions
    from branchedgp_b200.MBGP.assigngp import AssignGP
    t = np.array([0.0, 0.25, 0.5, 0.5, 0.9, 0.9])
    Y = np.array([[0.0, 0, 1, -1, 2.0, -3], [0, 0, 0, 0, 1, -1]]).T
    XExpanded, indices, _ = VBHelperFunctions.GetFunctionIndexListGeneral(t)
    phiInitial = np.array([[0.5, 0.5], [0.5, 0.5], [0.9, 0.1], [0.1, 0.9], [0.9, 0.1], [0.1, 0.9]])
    phiPrior = np.repeat(np.array([0.2, 0.4, 0.4])[None, :], t.size, axis=0)
    Yd = Y[:, :D]
    common = dict(phiPrior=phiPrior, phiInitial=phiInitial, fDebug=True)
    m_single = AssignGP(t, XExpanded, Yd, indices, **common, multi=False)
    gpflow.set_trainable(m_single.likelihood.variance, False)
    ll_s = m_single.maximum_log_likelihood_objective()
    mu_s, var_s = m_single.predict_f(XExpanded)
    m_multi = AssignGP(t, XExpanded, Yd, indices, **common, multi=True)
    ll_m = m_multi.maximum_log_likelihood_objective()
    mu_m, var_m = m_multi.predict_f(XExpanded)
    assert np.allclose(ll_s, ll_m)
    assert np.allclose(mu_s, mu_m)
    assert np.allclose(var_s, var_m)
    # and against the oracle
    _, trunk = host_ref.init_logphi_mbgp(phiInitial, t)
    eZ0 = host_ref.expand_prior_mbgp(phiPrior)
    Bv = np.full(D, 0.0001)
    e, _ = bgp_numpy.mbgp_dense_value_and_grad(m_single.logPhi.numpy(), Yd, XExpanded, t, eZ0, trunk, Bv, 1.0, 1.0, 1.0)
    assert abs(ll_s - e) <= 1e-8 * abs(e)


@pytest.mark.parametrize("num_inputs", [1, 2, 3])
def test_mbgp_fitmodel_learns_branching_points(num_inputs):
    """testing/MBGP/test_synthetic_data.py::test_demo with ToyBranchedData (MBGP/data_generation.py:289-308) rebuilt inline."""
    from branchedgp_b200.MBGP.assigngp import get_branching_point_kernel
    from branchedgp_b200.MBGP.FitBranchingModel import FitModel
    true_bps = [0.2, 0.4, 0.5, 0.6][:num_inputs]
    starts = [0.01, 0.15, 0.4, 0.55]
    N = 30
    t = np.linspace(0, 1, N)
    Y = np.zeros((N, num_inputs))
    state = np.ones(N, dtype=int)
    state[::2] = 2
    state[1::2] = 3
    for ib, b in enumerate(true_bps):
        idx2 = np.logical_and(t > b, state == 2)
        idx3 = np.logical_and(t > b, state == 3)
        Y[idx2, ib] = t[idx2] ** 2 - b ** 2
        Y[idx3, ib] = -t[idx3] ** 2 + b ** 2
    kern = get_branching_point_kernel(starts[:num_inputs])
    fit = FitModel(starts, t, Y, state, M=0, maxiter=100, priorConfidence=0.65, kervar=1.0, kerlen=1.0, likvar=0.01,
                   optimisation_method="L-BFGS-B", kern=kern)
    assert np.all(np.isfinite(fit["loglik"]))
    assert fit["prediction"]["mu"][0].shape == (100, num_inputs)
    for true_bp, inferred_bp in zip(true_bps, fit["Bmode"][:num_inputs]):
        assert abs(true_bp - inferred_bp) < 0.01, f"True branching points: {true_bps}. Inferred: {fit['Bmode']}"


@pytest.mark.parametrize("which", ["scipy", "amazing", "randomising"])
def test_mbgp_training_helpers_optimisers(which):
    """testing/MBGP/test_synthetic_data.py::test_demo__using_other_training_methods (two genes)."""
    from branchedgp_b200.MBGP.data_generation import ToyBranchedData
    from branchedgp_b200.MBGP.training_helpers import (ElvijsAmazingOptimiser, ElvijsRandomisingOptimiser, ScipyOptimiser,
                                                       SimplePhiConstructor, construct_and_fit_assigngp_model, get_training_outcome)
    np.random.seed(0)
    true_bps = [0.2, 0.4]
    data = ToyBranchedData(B=true_bps, N=30)
    opt = {"scipy": ScipyOptimiser(), "amazing": ElvijsAmazingOptimiser(),
           "randomising": ElvijsRandomisingOptimiser(base_optimiser=ScipyOptimiser(), num_samples=300)}[which]
    model = construct_and_fit_assigngp_model(data, phi_constructor=SimplePhiConstructor(data, prior_confidence=0.65),
                                             initial_branching_points=[0.01, 0.15], optimiser=opt)
    outcome = get_training_outcome(model)
    assert outcome.predictions.y_mean[0].shape == (100, 2)
    assert np.all(np.isfinite(outcome.predictions.y_var[1]))
    if which != "scipy":   # plain L-BFGS from early branching points is not expected to find them (the reference only tests the others)
        assert np.all(np.abs(np.array(true_bps) - outcome.learned_branching_points) < 0.01), outcome.learned_branching_points


def test_batched_branching_point_scan_matches_sequential():
    from branchedgp_b200.MBGP.data_generation import ToyBranchedData
    from branchedgp_b200.MBGP.training_helpers import SimplePhiConstructor, construct_assigngp_model
    np.random.seed(1)
    data = ToyBranchedData(B=[0.3, 0.6], N=20)
    m = construct_assigngp_model(data, SimplePhiConstructor(data, prior_confidence=0.65), [0.2, 0.2])
    cand = np.random.uniform(size=(5, 2))
    batched = m.training_losses_for_branching_points(cand)
    seq = []
    for c in cand:
        m.kernel.Bv.assign(c)
        seq.append(m.training_loss())
    assert np.allclose(batched, seq, rtol=1e-6, atol=0)   # Bv round-trips through the sigmoid transform in the sequential path
