"""CPU restatement of the FitModel driver (FitBranchingModel.py:11-178) on top of the numpy oracle.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  PARITY UNPINNED: oracle = restatement of TF 2.4.0 / GPflow 2.1.2
semantics; the optimiser is the real third-party one the reference calls (SciPy L-BFGS-B through
``gpflow.optimizers.Scipy``, default options plus ``maxiter``).

Per grid point b (FitBranchingModel.py:96-140): ``UpdateBranchingPoint`` re-initialises logPhi
(assigngp_dense.py:92-128), the trainable variables [lengthscale, kernel variance, likelihood variance, logPhi] are
minimised in GPflow's unconstrained space (softplus; likelihood variance with lower bound 1e-6) on
``training_loss = -(bound + log prior)`` with the Normal priors of FitBranchingModel.py:109-117, hyper-parameters
carry over to the next grid point, and ``loglik[b] = log_posterior_density`` at the optimum.
"""
import math

import numpy as np
import scipy.optimize

from . import bgp_numpy, host_ref

PRIORS = ((2.0, 1.0), (3.0, 1.0), (0.1, 0.1))   # Normal(mean, sd) on lengthscale, kernel variance, likelihood variance
LOWER = (0.0, 0.0, 1e-6)


def _softplus(u):
    return np.logaddexp(0.0, u)


def _softplus_inv(z):
    return z + np.log(-np.expm1(-z))


def _log_prior(h):
    return sum(-0.5 * ((h[k] - m) / s) ** 2 - math.log(s) - 0.5 * math.log(2 * math.pi) for k, (m, s) in enumerate(PRIORS))


def fit_model(bConsider, t, Y, labels, priorConfidence=0.8, M=0, likvar=1.0, kerlen=2.0, kervar=5.0, maxiter=100,
              fixHyperparameters=False):
    """Returns dict(loglik[B], Phi[N,3] at the best grid point, hyperparameters (kerlen, kervar, likvar) there, hyp_all[B,3])."""
    t = np.asarray(t, dtype=np.float64)
    Y = np.asarray(Y, dtype=np.float64)
    N = t.size
    phi0, prior = host_ref.initial_conditions_and_prior(np.asarray(labels), priorConfidence)
    X, _ = host_ref.function_index_list(t)
    eZ0 = host_ref.expand_prior_bgp(prior)
    Z = None
    if M > 0:   # FitBranchingModel.py:83-85
        Z = np.ones((M, 2))
        Z[:, 0] = np.linspace(0, 1, M, endpoint=False)
        Z[:, 1] = np.array([i for j in range(M) for i in range(1, 4)])[:M]
    hyp = np.array([kerlen, kervar, likvar], dtype=np.float64)
    ll = np.zeros(len(bConsider))
    phis, hyps = [], []
    for ib, b in enumerate(bConsider):
        pZ = host_ref.prior_with_trunk_bgp(eZ0, b, t)
        lp0 = host_ref.init_logphi_bgp(phi0, t, b)

        def value_and_grad(logPhi, h):
            if Z is None:
                return bgp_numpy.dense_value_and_grad(logPhi, Y, X, pZ, b, h[0], h[1], h[2])
            return bgp_numpy.sparse_value_and_grad(logPhi, Y, X, Z, pZ, b, h[0], h[1], h[2])

        if fixHyperparameters:
            def fun(x):
                e, g = value_and_grad(x.reshape(N, 3 * N), hyp)
                return -e, -g["logPhi"].ravel()
            x0 = lp0.ravel()
        else:
            def fun(x):
                u = x[:3]
                h = _softplus(u) + np.array(LOWER)
                e, g = value_and_grad(x[3:].reshape(N, 3 * N), h)
                gh = np.array([g["ell"], g["var"], g["likvar"]])
                gh = gh + np.array([-(h[k] - m) / (s * s) for k, (m, s) in enumerate(PRIORS)])
                gu = gh / (1.0 + np.exp(-u))
                return -(e + _log_prior(h)), np.concatenate([-gu, -g["logPhi"].ravel()])
            x0 = np.concatenate([_softplus_inv(hyp - np.array(LOWER)), lp0.ravel()])
        res = scipy.optimize.minimize(fun, x0, jac=True, method="L-BFGS-B", options=dict(maxiter=maxiter))
        if fixHyperparameters:
            logPhi = res.x.reshape(N, 3 * N)
            ll[ib] = -res.fun
        else:
            hyp = _softplus(res.x[:3]) + np.array(LOWER)
            logPhi = res.x[3:].reshape(N, 3 * N)
            ll[ib] = -res.fun   # log posterior density = bound + log prior
        S = bgp_numpy.softmax_rows(logPhi)
        phis.append(np.stack([S[np.arange(N), 3 * np.arange(N) + f] for f in range(3)], axis=1))
        hyps.append(hyp.copy())
    iw = int(np.argmax(ll))
    return dict(loglik=ll, Phi=phis[iw], hyperparameters=hyps[iw], hyp_all=np.array(hyps), Phi_all=np.array(phis))
