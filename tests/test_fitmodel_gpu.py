"""GPU: the drop-in FitModel driver (dense model, M = 0) on the reference's own synthetic test problem."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _toy(N=20):
    t = np.linspace(0, 1, N)
    Y = np.zeros((N, 1))
    idx = np.nonzero(t > 0.5)[0]
    idxA, idxB = idx[::2], idx[1::2]
    Y[idxA, 0] = 2 * t[idxA]
    Y[idxB, 0] = -2 * t[idxB]
    labels = np.ones(N)
    istart = np.min(np.flatnonzero(t > 0.5))
    labels[istart::2] = 2
    labels[(istart + 1)::2] = 3
    return t, Y, labels, idxA, idxB


def test_fitmodel_dense_recovers_branching_point_and_assignment():
    """Mirror of testing/test_objective.py (dense variant, M = 0, no prediction)."""
    from branchedgp_b200 import FitBranchingModel, VBHelperFunctions
    np.random.seed(43)
    N = 20
    t, Y, labels, idxA, idxB = _toy(N)
    ptb = np.min([np.min(t[labels == 2]), np.min(t[labels == 3])])
    grid = [ptb, 0.1, 0.4, np.round(ptb, 1), np.round(ptb + 0.01, 2), 0.7, 1.1]
    d = FitBranchingModel.FitModel(grid, t, Y, labels, M=0, maxiter=50, priorConfidence=0.65, kervar=1.0, kerlen=1.0, likvar=0.01,
                                   fPredict=False)
    assert np.all(np.isfinite(d["loglik"]))
    Bmode = grid[int(np.argmax(d["loglik"]))]
    assign = np.ones(N)
    assign[d["Phi"][:, 1] > 0.5] = 2
    assign[d["Phi"][:, 1] < 0.5] = 3
    assign[d["Phi"][:, 0] > 0.99] = 1
    alt = np.ones(N)
    alt[d["Phi"][:, 1] > 0.5] = 3
    alt[d["Phi"][:, 1] < 0.5] = 2
    alt[d["Phi"][:, 0] > 0.99] = 1
    truth = np.ones(N)
    truth[idxA] = 2
    truth[idxB] = 3
    acc = max((truth == assign).sum(), (truth == alt).sum())
    assert Bmode == 0.4, "Picked B=%f" % Bmode
    assert acc >= N - 2, "Got accuracy of %g" % acc
    # the null (b = 1.1, last) must lose against branching: positive log Bayes factor
    ev = VBHelperFunctions.CalculateBranchingEvidence(d, grid)
    assert ev["logBayesFactor"] > 0
    assert set(d) == {"loglik", "Phi", "prediction", "hyperparameters", "posteriorB"}
    assert set(d["hyperparameters"]) == {"likvar", "kerlen", "kervar"}


def test_training_loss_gradient_matches_finite_differences():
    from branchedgp_b200 import BranchingTree as bt
    from branchedgp_b200 import FitBranchingModel, VBHelperFunctions, assigngp_dense
    from branchedgp_b200 import branch_kernParamGPflow as bk
    from branchedgp_b200 import gpflow_shim as gpflow
    np.random.seed(1)
    t, Y, labels, _, _ = _toy(16)
    phi0, prior = FitBranchingModel.GetInitialConditionsAndPrior(labels, 0.8, True)
    X, indices, _ = VBHelperFunctions.GetFunctionIndexListGeneral(t)
    tree = bt.BinaryBranchingTree(0, 1)
    tree.add(None, 1, np.ones((1, 1)) * 0.5)
    fm, _ = tree.GetFunctionBranchTensor()
    kb = bk.BranchKernelParam(gpflow.kernels.Matern32(1), fm, b=np.zeros((1, 1))) + gpflow.kernels.White(1)
    kb.kernels[1].variance.assign(1e-6)
    gpflow.set_trainable(kb.kernels[1].variance, False)
    m = assigngp_dense.AssignGP(t, X, Y, kb, indices, np.ones((1, 1)) * 0.4, phiInitial=phi0, phiPrior=prior)
    m.likelihood.variance.assign(0.3)
    m.kernel.kernels[0].kern.lengthscales.assign(1.5)
    m.kernel.kernels[0].kern.variance.assign(2.0)
    m.kernel.kernels[0].kern.lengthscales.prior = gpflow.Normal(2.0, 1.0)
    m.kernel.kernels[0].kern.variance.prior = gpflow.Normal(3.0, 1.0)
    m.likelihood.variance.prior = gpflow.Normal(0.1, 0.1)
    vs = m.trainable_variables
    assert len(vs) == 4
    loss, grads = m._loss_and_grad(vs)
    assert np.isclose(loss, m.training_loss(), rtol=1e-12)
    h = 1e-6
    for p, g in zip(vs[:3], grads[:3]):
        u0 = p.unconstrained_variable.copy()
        p.set_unconstrained(u0 + h)
        lp = m.training_loss()
        p.set_unconstrained(u0 - h)
        lm = m.training_loss()
        p.set_unconstrained(u0)
        assert np.isclose((lp - lm) / (2 * h), float(g), rtol=2e-5, atol=1e-6), (p.name, (lp - lm) / (2 * h), g)
