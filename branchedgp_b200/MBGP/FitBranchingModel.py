"""Driver for the multi-output model (MBGP/FitBranchingModel.py:31-226): same signature and return dictionary."""
import logging
from typing import Optional

import numpy as np

from .. import gpflow_shim as gpflow
from ..FitBranchingModel import GetInitialConditionsAndPrior, GetPosteriorB  # noqa: F401  (verbatim copies in the reference)
from . import VBHelperFunctions
from .assigngp import AssignGP, BranchKernelParam

LOG = logging.getLogger("mBGP")


def FitModel(bConsider, GPt, GPy, globalBranching, priorConfidence=0.80, M=10, likvar=1.0, kerlen=2.0, kervar=5.0, fDebug=False,
             maxiter=1000, fPredict=True, fixHyperparameters=False, optimisation_method="L-BFGS-B",
             kern: Optional[BranchKernelParam] = None):
    """Train the multi-output model from each starting branching point in ``bConsider`` (only M == 0 is supported, as in
    the reference) and return the best restart."""
    assert isinstance(bConsider, list), "Candidate B must be list"
    assert GPt.ndim == 1
    assert GPy.ndim == 2
    assert GPt.size == GPy.shape[0], "pseudotime and gene expression data must be the same size"
    assert globalBranching.size == GPy.shape[0], "state space must be same size as number of cells"
    assert M >= 0, "at least 0 or more inducing points should be given"
    phiInitial, phiPrior = GetInitialConditionsAndPrior(globalBranching, priorConfidence, infPriorPhi=True)
    phiPrior = np.c_[np.zeros(phiPrior.shape[0])[:, None], phiPrior]   # prepend 0 for trunk
    XExpanded, indices, _ = VBHelperFunctions.GetFunctionIndexListGeneral(GPt)
    if M == 0:
        LOG.info("Experimental assign")
        m = AssignGP(GPt, XExpanded, GPy, indices, kern=kern, phiInitial=phiInitial, phiPrior=phiPrior, multi=True)
    else:
        raise NotImplementedError
    ll = np.zeros(len(bConsider))
    Phi_l, branching_points = list(), list()
    ttestl_l, mul_l, varl_l = list(), list(), list()
    hyps, models = list(), list()
    for ib, b in enumerate(bConsider):
        m.UpdateBranchingPoint(np.ones((GPy.shape[1], 1)) * b)
        trained_model = trainModel(m, maxiter, method=optimisation_method)
        hyps.append({"likvar": trained_model.likelihood.variance.numpy()})
        ll[ib] = trained_model.log_posterior_density()
        models.append(trained_model)
        branching_points.append(trained_model.kernel.Bv.numpy().flatten())
        Phi_l.append(trained_model.GetPhi())
        if fPredict:
            ttestl, mul, varl = VBHelperFunctions.predictBranchingModel(trained_model)
            ttestl_l.append(ttestl), mul_l.append(mul), varl_l.append(varl)
        else:
            ttestl_l.append([]), mul_l.append([]), varl_l.append([])
    iw = np.argmax(ll)
    postB = GetPosteriorB(ll, bConsider)
    return {"loglik": ll, "Bmode": branching_points[iw], "Phi": Phi_l[iw], "model": models[iw],
            "prediction": {"xtest": ttestl_l[iw], "mu": mul_l[iw], "var": varl_l[iw]}, "hyperparameters": hyps[iw],
            "posteriorB": postB, "globalB": branching_points[0]}


def trainModel(gpflow_model, maxiter=100, method="L-BFGS-B"):
    LOG.info(f"Starting training. Initial loss: {gpflow_model.training_loss()}.")
    if method == "Adam":
        raise NotImplementedError("only the SciPy optimisers are supported")
    opt = gpflow.optimizers.Scipy()
    result = opt.minimize(gpflow_model.training_loss, variables=gpflow_model.trainable_variables, options=dict(maxiter=maxiter),
                          method=method)
    LOG.info(result)
    LOG.info(f"Training complete. Final loss: {gpflow_model.training_loss()}.")
    return gpflow_model
