"""Grid-search driver for the BGP model (FitBranchingModel.py:11-257): same signature and return dictionary."""
import numpy as np

from . import BranchingTree as bt
from . import VBHelperFunctions, assigngp_dense
from . import branch_kernParamGPflow as bk
from . import gpflow_shim as gpflow
from .gpflow_shim import set_trainable, to_default_float
from .gpflow_shim import distributions as _tfd


def FitModel(bConsider, GPt, GPy, globalBranching, priorConfidence=0.80, M=10, likvar=1.0, kerlen=2.0, kervar=5.0,
             fDebug=False, maxiter=100, fPredict=True, fixHyperparameters=False, *, optimizer="scipy"):
    """Fit the BGP model for every candidate branching point in ``bConsider`` (warm-starting the hyper-parameters from
    one grid point to the next, as the reference does) and return
    ``{loglik, Phi, prediction{xtest, mu, var}, hyperparameters, posteriorB}``.

    ``optimizer="scipy"`` (default) is the reference's own driver: SciPy L-BFGS-B calling the CUDA bound once per function
    evaluation.  ``optimizer="device"`` (keyword only, not in the reference) hands the same problem to the device-resident
    batched L-BFGS of ``FitModelBatch`` -- with ``fixHyperparameters`` all grid points are then fitted concurrently."""
    if optimizer == "device":
        assert GPy.ndim == 2 and GPy.shape[1] == 1, "the BGP model fits one gene (FitBranchingModel.py:47-49)"
        res = FitModelBatch(bConsider, GPt, GPy, globalBranching, priorConfidence=priorConfidence, M=M, likvar=likvar, kerlen=kerlen,
                            kervar=kervar, fDebug=fDebug, maxiter=maxiter, fPredict=fPredict, fixHyperparameters=fixHyperparameters)[0]
        res.pop("fit", None)
        return res
    if optimizer != "scipy":
        raise ValueError("optimizer must be 'scipy' or 'device'")
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


def FitModelBatch(bConsider, GPt, GPY, globalBranching, priorConfidence=0.80, M=10, likvar=1.0, kerlen=2.0, kervar=5.0,
                  fDebug=False, maxiter=100, fPredict=True, fixHyperparameters=False, engine=None, max_slots=0, maxcor=10):
    """``FitModel`` for every gene (column of ``GPY`` [N, G]) at once: returns a list of G dictionaries with FitModel's keys.

    Per gene the semantics are FitModel's (FitBranchingModel.py:11-178): the same initial conditions and label prior for
    every gene, ``logPhi`` re-initialised at every candidate branching point, trainable hyper-parameters warm-started from
    one grid point to the next, log posterior density as ``loglik``, and the error branch (NaN in ``loglik[0]``) when a
    factorisation fails.  What changes is where the optimiser runs: all genes are fitted together by the device-resident
    batched L-BFGS of ``libbgp_b200`` (bgp_dense_fit_host for M = 0, bgp_sparse_fit_host otherwise) -- one call per grid
    point with G work items when the hyper-parameters are trainable (the warm start makes the grid a serial chain per
    gene), one call with G x len(bConsider) items when ``fixHyperparameters``.

    ``maxcor`` is SciPy's number of stored L-BFGS pairs (10).  The optimiser state of an item is (5 + 2 maxcor) vectors of
    3 N^2 doubles (0.6 GB at N = 1000, 2.4 GB at N = 2000), which bounds the number of items fitted concurrently (148 at
    N = 1000, 52 at N = 2000 on a 180 GB B200); a smaller ``maxcor`` trades optimiser memory for more items in flight."""
    from . import _lib
    assert isinstance(bConsider, list), "Candidate B must be list"
    GPt = np.asarray(GPt, dtype=np.float64)
    GPY = np.asarray(GPY, dtype=np.float64)
    assert GPt.ndim == 1
    assert GPY.ndim == 2 and GPY.shape[0] == GPt.size, "expression must be [cells, genes]"
    assert globalBranching.size == GPt.size, "state space must be same size as number of cells"
    assert M >= 0, "at least 0 or more inducing points should be given"
    N, G = GPY.shape
    B = len(bConsider)
    eng = engine if engine is not None else _lib.default_engine()
    phiInitial, phiPrior = GetInitialConditionsAndPrior(globalBranching, priorConfidence, infPriorPhi=True)
    VBHelperFunctions.GetFunctionIndexListGeneral(GPt)   # consumes numpy's global RNG exactly like FitModel
    Z = None
    if M > 0:
        Z = np.ones((M, 2))
        Z[:, 0] = np.linspace(0, 1, M, endpoint=False)
        Z[:, 1] = np.array([i for j in range(M) for i in range(1, 4)])[:M]
    opts = eng.fit_options(maxiter=int(maxiter), train_hyp=0 if fixHyperparameters else 1, max_slots=int(max_slots), maxcor=int(maxcor))
    Y = np.ascontiguousarray(GPY.T[:, :, None])          # [G, N, 1]
    lo, hi = float(np.min(GPt)), float(np.max(GPt))

    def test_inputs(b):
        cols = [np.linspace(lo, b, 100), np.linspace(b, hi, 100), np.linspace(b, hi, 100)]
        return np.vstack([np.stack([c, np.full(100, float(f))], axis=1) for f, c in zip((1, 2, 3), cols)])

    ll = np.zeros((G, B))
    hyp_at = np.empty((G, B, 3))
    phi_at = np.empty((G, B, N, 3))
    pred_mu = np.empty((G, B, 300, 1)) if fPredict else None
    pred_var = np.empty((G, B, 300)) if fPredict else None
    failed = np.zeros(G, dtype=bool)
    failed_at = np.full(G, B, dtype=int)      # first grid index whose factorisation failed
    fit_info = [dict(iters=np.zeros(B, dtype=int), nfev=np.zeros(B, dtype=int), status=np.zeros(B, dtype=int)) for _ in range(G)]

    def run(genes, ibs, hyp0):
        bs = np.array([bConsider[ib] for ib in ibs], dtype=np.float64)
        Xnew = np.stack([test_inputs(b) for b in bs]) if fPredict else None
        out = eng.fit(GPt, Y, genes, bs, phiPrior, phiInitial, hyp0, Z=Z, options=opts, Xnew=Xnew)
        for k, (g, ib) in enumerate(zip(genes, ibs)):
            ll[g, ib] = out["objective"][k]
            hyp_at[g, ib] = out["hyp"][k]
            phi_at[g, ib] = out["phi"][k]
            if fPredict:
                pred_mu[g, ib] = out["pred_mean"][k]
                pred_var[g, ib] = out["pred_var"][k]
            for key in ("iters", "nfev", "status"):
                fit_info[g][key][ib] = out[key][k]
            if out["status"][k] == 5:
                failed_at[g] = min(failed_at[g], int(ib))
            if out["status"][k] == 5 and not failed[g]:
                failed[g] = True
                print(f"Unexpected error: Cholesky decomposition was not successful (gene {g}, b={bConsider[ib]})")
        return out

    if fixHyperparameters:
        genes = np.repeat(np.arange(G), B)
        ibs = np.tile(np.arange(B), G)
        run(genes, ibs, np.tile([kerlen, kervar, likvar], (G * B, 1)))
    else:
        cur = np.tile([kerlen, kervar, likvar], (G, 1)).astype(np.float64)
        for ib in range(B):
            genes = np.flatnonzero(~failed)
            if genes.size == 0:
                break
            out = run(genes, [ib] * genes.size, cur[genes])
            cur[genes] = out["hyp"]

    results = []
    for g in range(G):
        if failed[g]:
            # FitModel's error branch (FitBranchingModel.py:142-153): the grid entries computed before the failing point are kept,
            # the failing one and everything after it stay zero, and only ll[0] is overwritten with NaN
            bad = ll[g].copy()
            bad[failed_at[g]:] = 0.0
            bad[0] = np.nan
            results.append({"loglik": bad, "model": None, "Phi": np.nan, "prediction": {"xtest": np.nan, "mu": np.nan, "var": np.nan},
                            "hyperparameters": np.nan, "posteriorB": np.nan, "fit": fit_info[g]})
            continue
        iw = int(np.argmax(ll[g]))
        postB = GetPosteriorB(ll[g], bConsider)
        assert np.allclose(bConsider[iw], postB["Bmode"]), "%s-%s" % (str(postB["B_CI"]), bConsider[iw])
        if fPredict:
            X = test_inputs(bConsider[iw])
            pred = {"xtest": [X[100 * f:100 * (f + 1), :1] for f in range(3)],
                    "mu": [pred_mu[g, iw, 100 * f:100 * (f + 1)] for f in range(3)],
                    "var": [pred_var[g, iw, 100 * f:100 * (f + 1), None] for f in range(3)]}
        else:
            pred = {"xtest": [], "mu": [], "var": []}
        results.append({"loglik": ll[g].copy(), "Phi": phi_at[g, iw],
                        "prediction": pred,
                        "hyperparameters": {"likvar": np.array(hyp_at[g, iw, 2]), "kerlen": np.array(hyp_at[g, iw, 0]),
                                            "kervar": np.array(hyp_at[g, iw, 1])},
                        "posteriorB": postB, "fit": fit_info[g]})
    return results


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
