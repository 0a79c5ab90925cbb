"""Host-side helpers around the fitted models (VBHelperFunctions.py:5-36, 141-207); plotting is out of scope."""
import numpy as np


def CalculateBranchingEvidence(d, Bsearch=None):
    """Posterior over the candidate branching points and log Bayes factor of branching vs. not branching.

    ``d['loglik']`` holds one log posterior density per grid point; by convention the LAST grid point is the
    "no branching" null (VBHelperFunctions.py:5-36)."""
    if Bsearch is None:
        Bsearch = list(np.linspace(0.05, 0.95, 5)) + [1.1]
    ll = np.asarray(d["loglik"], dtype=np.float64)
    cand = ll[:-1]
    w = np.exp(cand - np.max(cand))
    posterior = w / w.sum()
    Nb = ll.size - 1
    if Nb != len(Bsearch) - 1:
        raise NameError("Passed in wrong length of Bsearch is %g- should be %g" % (len(Bsearch), Nb))
    imax = int(np.argmax(cand))
    top = cand[imax]
    rest = np.delete(cand, imax)
    log_bf = top + np.log(1.0 + np.exp(rest - top).sum()) - ll[-1] - np.log(Nb)
    return {"posteriorBranching": posterior, "logBayesFactor": log_bf}


def GetFunctionIndexListGeneral(Xin):
    """Cell-major expansion of pseudotime over the three latent functions (VBHelperFunctions.py:170-194).

    Returns (XExpanded [3N, 2] with rows (t_i, f), f = 1, 2, 3;  indices[i] = [3i, 3i+1, 3i+2];  XSample [N, 2] with a
    random function per cell).  Like the reference it draws N values from numpy's global RNG for XSample."""
    Xin = np.asarray(Xin)
    assert Xin.shape[0] == np.size(Xin)
    t = Xin.reshape(-1).astype(np.float64)
    n = t.size
    XSample = np.zeros((n, 2), dtype=float)
    XSample[:, 0] = t
    for i in range(n):
        XSample[i, 1] = np.random.choice([1, 2, 3])
    X = np.empty((3 * n, 2), dtype=float)
    X[:, 0] = np.repeat(t, 3)
    X[:, 1] = np.tile(np.array([1.0, 2.0, 3.0]), n)
    indices = [[3 * i, 3 * i + 1, 3 * i + 2] for i in range(n)]
    return X, indices, XSample


def SetXExpandedBranchingPoint(XExpanded, B):
    """Rows available under branching point B: function 1 up to B, functions 2 and 3 after it."""
    XExpanded = np.asarray(XExpanded)
    B = float(np.asarray(B).ravel()[0])
    keep_trunk = (XExpanded[:, 0] <= B) & (XExpanded[:, 1] == 1)
    keep_branch = (XExpanded[:, 0] > B) & (XExpanded[:, 1] != 1)
    return np.vstack([XExpanded[keep_trunk], XExpanded[keep_branch]])


def predictBranchingModel(m, full_cov=False):
    """Predictive mean / variance of the three functions on 100 test points each (VBHelperFunctions.py:141-167).  The reference
    calls ``predict_f`` once per function, each call refactorising the model; here the 300 test inputs go through ONE posterior
    evaluation (one pair of Cholesky factorisations on the GPU) and the result is cut into the reference's three lists."""
    pt = m.t
    B = np.asarray(m.kernel.kernels[0].Bv).flatten()
    lo, hi = np.min(pt), np.max(pt)
    ttestl = [np.linspace(lo, B, 100) if f == 1 else np.linspace(B, hi, 100) for f in range(1, 4)]
    Xtest = np.vstack([np.hstack((tt, tt * 0 + f)) for f, tt in zip(range(1, 4), ttestl)])
    mu, var = m.predict_f(Xtest, full_cov=full_cov)
    assert np.all(np.isfinite(mu)), "All elements should be finite but are " + str(mu)
    assert np.all(np.isfinite(var)), "All elements should be finite but are " + str(var)
    mul = [mu[100 * f:100 * (f + 1)] for f in range(3)]
    if full_cov:
        varl = [var[100 * f:100 * (f + 1), 100 * f:100 * (f + 1)] for f in range(3)]
    else:
        varl = [var[100 * f:100 * (f + 1)] for f in range(3)]
    return ttestl, mul, varl
