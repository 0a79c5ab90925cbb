"""MBGP copies of the host helpers (MBGP/VBHelperFunctions.py); only predictBranchingModel differs from the BGP module."""
import numpy as np

from ..VBHelperFunctions import (CalculateBranchingEvidence, GetFunctionIndexListGeneral,  # noqa: F401
                                 SetXExpandedBranchingPoint)


def predictBranchingModel(m, Bi=None, full_cov=False):
    """Predictions of the three functions on 100 points spanning the whole pseudotime range (MBGP/VBHelperFunctions.py:214-242)."""
    pt = m.t
    lo, hi = np.min(pt), np.max(pt)
    mul, varl, ttestl = [], [], []
    for f in range(1, 4):
        ttest = np.linspace(lo, hi, 100)[:, None]
        Xtest = np.hstack((ttest, ttest * 0 + f))
        mu, var = m.predict_f(Xtest, full_cov=full_cov)
        assert np.all(np.isfinite(mu)), "All elements should be finite but are " + str(mu)
        assert np.all(np.isfinite(var)), "All elements should be finite but are " + str(var)
        mul.append(mu)
        varl.append(var)
        ttestl.append(ttest)
    return ttestl, mul, varl
