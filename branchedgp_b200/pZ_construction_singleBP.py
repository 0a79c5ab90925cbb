"""Dense views of the assignment prior p(Z) (pZ_construction_singleBP.py:6-25).

The CUDA kernels never read these matrices -- they evaluate the same rule per element from (prior[i], t_i <= b)
(csrc/bgp_phi.cu: log_pz).  The dense N x 3N versions exist for callers that inspect ``model.pZ`` / ``model.eZ0``.
"""
import numpy as np


def expand_pZ0Zeros(pZ0, epsilon=1e-6):
    pZ0 = np.asarray(pZ0, dtype=np.float64)
    assert pZ0.shape[1] == 2, "Should have exactly two cols got %g " % pZ0.shape[1]
    n = pZ0.shape[0]
    for z0 in pZ0:
        assert z0.sum() == 1, "should sum to 1 is %s=%.3f" % (str(z0), z0.sum())
    r = np.full((n, 3 * n), float(epsilon))
    rows = np.arange(n)
    r[rows, 3 * rows + 1] = pZ0[:, 0]
    r[rows, 3 * rows + 2] = pZ0[:, 1]
    return r


def expand_pZ0PureNumpyZeros(eZ0, BP, X, epsilon=1e-6):
    r = np.array(eZ0, dtype=np.float64, copy=True)
    trunk = np.flatnonzero(np.asarray(X).ravel() <= np.asarray(BP).ravel()[0])   # convention: t <= b is trunk
    r[trunk, 3 * trunk] = 1.0
    r[trunk, 3 * trunk + 1] = epsilon
    r[trunk, 3 * trunk + 2] = epsilon
    return r
