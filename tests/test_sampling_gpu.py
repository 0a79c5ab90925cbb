"""GPU: SURVEY row f4 -- MBGP AssignGP.sample_prior / predict_f_samples (MBGP/assigngp.py:741-835) on the batched CUDA sampler
(bgp_gaussian_samples_host: covariance, Cholesky and L z per output in one CTA) against the numpy restatement of the same three
steps fed with the SAME normal draws, plus the reference's own shape tests (testing/MBGP/test_assigngp.py:35: 300 rows;
testing/MBGP/test_helpers.py:23: 150 rows)."""
import numpy as np
import pytest

from branchedgp_b200 import VBHelperFunctions as vbh
from branchedgp_b200 import _lib
from branchedgp_b200.MBGP import assigngp as mbgp

pytestmark = pytest.mark.gpu
JITTER = 1e-6


def _model(N=40, D=3, seed=0, multi=True):
    rng = np.random.default_rng(seed)
    t = np.sort(rng.uniform(0.01, 0.99, N))
    XExpanded, indices, _ = vbh.GetFunctionIndexListGeneral(t)
    Y = rng.standard_normal((N, D))
    prior = np.tile([0.0, 0.5, 0.5], (N, 1))
    m = mbgp.AssignGP(t, XExpanded, Y, indices, phiPrior=prior, phiInitial=np.tile([0.5, 0.5], (N, 1)), multi=multi,
                      kern=mbgp.get_branching_point_kernel(list(rng.uniform(0.2, 0.8, D))))
    m.kernel.kern.lengthscales.assign(0.6)
    m.kernel.kern.variance.assign(1.7)
    m.likelihood.variance.assign(0.1)
    return m, rng


def _xexp(rows, seed=1):
    rng = np.random.default_rng(seed)
    return np.stack([np.sort(rng.uniform(0.0, 1.0, rows)), rng.integers(1, 4, rows).astype(float)], 1)


@pytest.mark.parametrize("rows,S", [(31, 2), (150, 3), (300, 4)])
def test_sample_prior_matches_numpy_draw_for_draw(rows, S):
    m, _ = _model()
    x = _xexp(rows)
    np.random.seed(11)
    got = m.sample_prior(x, num_samples=S)
    np.random.seed(11)
    z = np.random.standard_normal((m.D, rows, S))
    cov = np.stack([m.kernel.K(x, dim=d) for d in range(m.D)])          # BranchKernelParam.K incl. noise_level on the square
    L = np.linalg.cholesky(cov + JITTER * np.eye(rows)[None])
    want = np.transpose(L @ z, (2, 1, 0))
    assert got.shape == (S, rows, m.D) and np.all(np.isfinite(got))
    # a smooth kernel with a 2e-6 nugget has condition number ~1e8, so two correct Cholesky factors differ by ~1e8 x round-off
    # (measured 2.6e-9 at 300 rows); the factorisation itself is held to 1e-12 in the reconstruction test below
    assert np.abs(got - want).max() <= 1e-6 * np.abs(want).max()


@pytest.mark.parametrize("full_cov", [True, False])
@pytest.mark.parametrize("multi", [True, False])
def test_predict_f_samples_shapes_and_values(full_cov, multi):
    """testing/MBGP/test_assigngp.py::test_f_samples shape contract ([S, 300, D], finite) plus values against numpy."""
    m, _ = _model(N=35, D=2, seed=3, multi=multi)
    x = _xexp(300, seed=5)
    np.random.seed(5)
    got = m.predict_f_samples(x, 3, full_cov=full_cov)
    assert got.shape == (3, 300, 2) and np.all(np.isfinite(got))
    mean, cov = m.predict_f(x, full_cov=full_cov)
    np.random.seed(5)
    if full_cov:
        assert cov.shape == (300, 300, 2)
        z = np.random.standard_normal((2, 300, 3))
        L = np.linalg.cholesky(np.transpose(cov, (2, 0, 1)) + JITTER * np.eye(300)[None])
        want = np.transpose(mean.T[..., None] + L @ z, (2, 1, 0))
        # the posterior covariance is nearly singular away from the data (only the 1e-6 jitter keeps it positive definite), so the
        # two Cholesky factors differ by round-off amplified by its condition number; compare at the scale of the samples
        assert np.abs(got - want).max() <= 1e-6 * max(1.0, np.abs(want).max())
    else:
        want = mean + np.sqrt(cov) * np.random.standard_normal((3, 300, 2))
        assert np.allclose(got, want, rtol=0, atol=1e-12)
    single = m.predict_f_samples(x, None, full_cov=full_cov)
    assert single.shape == (300, 2)


def test_gaussian_samples_reports_non_positive_definite_covariance(engine):
    cov = -np.eye(40)[None]
    with pytest.raises(_lib.CholeskyError):
        engine.gaussian_samples(np.zeros((1, 40, 2)), cov=cov)


def test_gaussian_samples_given_covariance_reconstructs_it(engine):
    """z = identity turns the samples into the Cholesky factor itself: L L^T must give back cov + jitter I (n = 200, 5 outputs)."""
    rng = np.random.default_rng(2)
    n, B = 200, 5
    A = rng.standard_normal((B, n, n))
    cov = A @ np.transpose(A, (0, 2, 1)) / n + 0.1 * np.eye(n)[None]
    z = np.broadcast_to(np.eye(n)[None], (B, n, n)).copy()
    mean = rng.standard_normal((B, n))
    out = engine.gaussian_samples(z, mean=mean, cov=cov, jitter=JITTER)          # [S = n, n, B]
    L = np.transpose(out - mean.T[None], (2, 1, 0))                               # [B, n, S]: column s = L e_s
    assert np.abs(np.triu(L, 1)).max() == 0.0
    rec = L @ np.transpose(L, (0, 2, 1))
    assert np.abs(rec - cov - JITTER * np.eye(n)[None]).max() <= 1e-12 * np.abs(cov).max()


def test_predict_branching_model_single_posterior_evaluation():
    """VBHelperFunctions.predictBranchingModel (one stacked predict_f call here) equals three separate predict_f calls, for
    marginal variances and for full covariances (300 inputs: three 128-input chunks of the tiled covariance)."""
    from branchedgp_b200 import FitBranchingModel
    rng = np.random.default_rng(4)
    N = 30
    t = np.sort(rng.uniform(0.02, 0.98, N))
    labels = np.where(t < 0.4, 1, rng.integers(2, 4, N)).astype(float)
    y = np.where(labels == 3, -1.0, 1.0) * 3 * np.maximum(t - 0.4, 0) + 0.2 * rng.standard_normal(N)
    d = FitBranchingModel.FitModel([0.4, 1.1], t, (y - y.mean())[:, None], labels, M=0, maxiter=5, fPredict=True)
    assert len(d["prediction"]["mu"]) == 3 and d["prediction"]["mu"][0].shape == (100, 1)
    # rebuild the model at the best point and compare stacked vs separate calls
    from branchedgp_b200 import assigngp_dense, branch_kernParamGPflow as bk, gpflow_shim as gpflow, BranchingTree as bt
    XExpanded, indices, _ = vbh.GetFunctionIndexListGeneral(t)
    tree = bt.BinaryBranchingTree(0, 1)
    tree.add(None, 1, np.ones((1, 1)) * 0.4)
    fm, _ = tree.GetFunctionBranchTensor()
    kb = bk.BranchKernelParam(gpflow.kernels.Matern32(1), fm, b=np.zeros((1, 1))) + gpflow.kernels.White(1)
    kb.kernels[1].variance.assign(1e-6)
    phi0, prior = FitBranchingModel.GetInitialConditionsAndPrior(labels, 0.8, True)
    m = assigngp_dense.AssignGP(t, XExpanded, (y - y.mean())[:, None], kb, indices, np.ones((1, 1)) * 0.4, phiInitial=phi0, phiPrior=prior)
    for full in (False, True):
        tt, mu, var = vbh.predictBranchingModel(m, full_cov=full)
        for f in range(3):
            Xt = np.hstack((tt[f], tt[f] * 0 + f + 1))
            m1, v1 = m.predict_f(Xt, full_cov=full)
            assert np.allclose(mu[f], m1, rtol=1e-9, atol=1e-10)
            assert np.allclose(var[f], v1, rtol=1e-8, atol=1e-9)
