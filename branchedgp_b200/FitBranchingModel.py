"""Grid-search driver for the BGP model (FitBranchingModel.py:11-257): same signature and return dictionary."""
import numpy as np

from . import BranchingTree as bt
from . import VBHelperFunctions, assigngp_dense
from . import branch_kernParamGPflow as bk
from . import gpflow_shim as gpflow
from .gpflow_shim import set_trainable, to_default_float
from .gpflow_shim import distributions as _tfd


def FitModel(bConsider, GPt, GPy, globalBranching, priorConfidence=0.80, M=10, likvar=1.0, kerlen=2.0, kervar=5.0,
             fDebug=False, maxiter=100, fPredict=True, fixHyperparameters=False):
    """Fit the BGP model for every candidate branching point in ``bConsider`` (warm-starting the hyper-parameters from
    one grid point to the next, as the reference does) and return
    ``{loglik, Phi, prediction{xtest, mu, var}, hyperparameters, posteriorB}``."""
    assert isinstance(bConsider, list), "Candidate B must be list"
    assert GPt.ndim == 1
    assert GPy.ndim == 2
    assert GPt.size == GPy.size, "pseudotime and gene expression data must be the same size"
    assert globalBranching.size == GPy.size, "state space must be same size as number of cells"
    assert M >= 0, "at least 0 or more inducing points should be given"
    phiInitial, phiPrior = GetInitialConditionsAndPrior(globalBranching, priorConfidence, infPriorPhi=True)
    XExpanded, indices, _ = VBHelperFunctions.GetFunctionIndexListGeneral(GPt)
    ptb = np.min([np.min(GPt[globalBranching == 2]), np.min(GPt[globalBranching == 3])])
    tree = bt.BinaryBranchingTree(0, 1, fDebug=False)
    tree.add(None, 1, np.ones((1, 1)) * ptb)
    (fm, _) = tree.GetFunctionBranchTensor()
    kb = bk.BranchKernelParam(gpflow.kernels.Matern32(1), fm, b=np.zeros((1, 1))) + gpflow.kernels.White(1)
    kb.kernels[1].variance.assign(1e-6)          # discontinuity magnitude at the branching point / jitter
    set_trainable(kb.kernels[1].variance, False)
    if M == 0:
        m = assigngp_dense.AssignGP(GPt, XExpanded, GPy, kb, indices, np.ones((1, 1)) * ptb, phiInitial=phiInitial,
                                    phiPrior=phiPrior)
    else:
        from . import assigngp_denseSparse
        ZExpanded = np.ones((M, 2))
        ZExpanded[:, 0] = np.linspace(0, 1, M, endpoint=False)
        ZExpanded[:, 1] = np.array([i for j in range(M) for i in range(1, 4)])[:M]
        m = assigngp_denseSparse.AssignGPSparse(GPt, XExpanded, GPy, kb, indices, np.ones((1, 1)) * ptb, ZExpanded,
                                                phiInitial=phiInitial, phiPrior=phiPrior)
    m.likelihood.variance.assign(likvar)
    m.kernel.kernels[0].kern.lengthscales.assign(kerlen)
    m.kernel.kernels[0].kern.variance.assign(kervar)
    if fixHyperparameters:
        print("Fixing hyperparameters")
        set_trainable(m.kernel.kernels[0].kern.lengthscales, False)
        set_trainable(m.likelihood.variance, False)
        set_trainable(m.kernel.kernels[0].kern.variance, False)
    else:
        if fDebug:
            print("Adding prior logistic on length scale to avoid numerical problems")
        m.kernel.kernels[0].kern.lengthscales.prior = _tfd.Normal(to_default_float(2.0), to_default_float(1.0))
        m.kernel.kernels[0].kern.variance.prior = _tfd.Normal(to_default_float(3.0), to_default_float(1.0))
        m.likelihood.variance.prior = _tfd.Normal(to_default_float(0.1), to_default_float(0.1))

    ll = np.zeros(len(bConsider))
    Phi_l = list()
    ttestl_l, mul_l, varl_l = list(), list(), list()
    hyps = list()
    for ib, b in enumerate(bConsider):
        m.UpdateBranchingPoint(np.ones((1, 1)) * b, phiInitial)
        try:
            opt = gpflow.optimizers.Scipy()
            opt.minimize(m.training_loss, variables=m.trainable_variables, options=dict(disp=True, maxiter=maxiter))
            hyps.append({"likvar": m.likelihood.variance.numpy(),
                         "kerlen": m.kernel.kernels[0].kern.lengthscales.numpy(),
                         "kervar": m.kernel.kernels[0].kern.variance.numpy()})
            ll[ib] = m.log_posterior_density()
        except Exception as ex:
            print(f"Unexpected error: {ex} {'-' * 60}\nCaused by model: {m} {'-' * 60}")
            ll[0] = np.nan
            return {"loglik": ll, "model": m, "Phi": np.nan, "prediction": {"xtest": np.nan, "mu": np.nan, "var": np.nan},
                    "hyperparameters": np.nan, "posteriorB": np.nan}
        Phi_l.append(m.GetPhi())
        if fPredict:
            ttestl, mul, varl = VBHelperFunctions.predictBranchingModel(m)
            ttestl_l.append(ttestl), mul_l.append(mul), varl_l.append(varl)
        else:
            ttestl_l.append([]), mul_l.append([]), varl_l.append([])
    iw = np.argmax(ll)
    postB = GetPosteriorB(ll, bConsider)
    if fDebug:
        print("BGP Maximum at b=%.2f" % bConsider[iw], "CI= [%.2f, %.2f]" % (postB["B_CI"][0], postB["B_CI"][1]))
    assert np.allclose(bConsider[iw], postB["Bmode"]), "%s-%s" % (str(postB["B_CI"]), bConsider[iw])
    return {"loglik": ll, "Phi": Phi_l[iw], "prediction": {"xtest": ttestl_l[iw], "mu": mul_l[iw], "var": varl_l[iw]},
            "hyperparameters": hyps[iw], "posteriorB": postB}


def GetPosteriorB(objUnsorted, BgridSearch, ciLimits=[0.01, 0.99]):
    """Posterior over the grid of branching points, its mode and the indices of the confidence limits
    (FitBranchingModel.py:181-222)."""
    assert objUnsorted.size == len(BgridSearch), "size do not match %g-%g" % (objUnsorted.size, len(BgridSearch))
    grid = np.array(BgridSearch)
    isort = np.argsort(grid)
    grid = grid[isort]
    obj = objUnsorted[isort].copy()
    imode = np.argmax(obj)
    w = np.exp(obj - np.max(obj))
    p = w / w.sum()
    assert np.any(~np.isnan(p)), "Nans in p! %s" % str(p)
    assert np.any(~np.isinf(p)), "Infinities in p! %s" % str(p)
    cdf = np.cumsum(p)
    confInt = np.zeros(len(ciLimits), dtype=int)
    for k, lim in enumerate(ciLimits):
        below = np.flatnonzero(cdf <= lim)
        confInt[k] = 0 if below.size == 0 else np.max(below)
    assert np.all(confInt < obj.size), confInt
    return {"B_CI": grid[confInt], "Bmode": grid[imode], "idx_confInt": confInt, "idx_mode": imode,
            "BgridSearch_sort": grid, "isort": isort}


def GetInitialConditionsAndPrior(globalBranching, v, infPriorPhi):
    """Initial assignment probabilities and label prior from the global cell labels {1 trunk, 2, 3}
    (FitBranchingModel.py:225-257).  Seeds numpy's global RNG with 42 and draws in the reference's order:
    N uniforms for the first column, then one uniform per branch-labelled cell in cell order."""
    np.random.seed(42)
    assert isinstance(v, float), "v should be scalar is %s" % str(type(v))
    labels = np.asarray(globalBranching).reshape(-1)
    N = labels.size
    phiInitial = np.empty((N, 2))
    phiInitial[:, 0] = np.random.rand(N)
    phiInitial[:, 1] = 1 - phiInitial[:, 0]
    phiPrior = np.full((N, 2), 0.5)
    branch_cells = np.flatnonzero(labels != 1)
    col = (labels[branch_cells] - 2).astype(int)
    if infPriorPhi:
        phiPrior[branch_cells, :] = 1 - v
        phiPrior[branch_cells, col] = v
    draws = np.random.random(branch_cells.size)
    own = 0.5 + draws / 2.0                     # in [0.5, 1]
    phiInitial[branch_cells, col] = own
    phiInitial[branch_cells, 1 - col] = 1 - own
    assert np.allclose(phiPrior.sum(1), 1), "Phi Prior should be close to 1 but got %s" % str(phiPrior)
    assert np.allclose(phiInitial.sum(1), 1), "Phi Initial should be close to 1 but got %s" % str(phiInitial)
    assert np.all(~np.isnan(phiInitial)), "No nans please!"
    assert np.all(~np.isnan(phiPrior)), "No nans please!"
    return phiInitial, phiPrior
