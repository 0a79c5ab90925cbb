"""Loop-for-loop restatement of the reference's host-side constructors (dense N x 3N matrices).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  The product never materialises pZ / phiTrunk; these
dense versions exist so the tests can check the in-kernel rules against what the reference builds.
"""
import numpy as np


def function_index_list(t):
    """VBHelperFunctions.GetFunctionIndexListGeneral (VBHelperFunctions.py:170-194): cell-major expansion."""
    rows, idx, k = [], [], 0
    for x in t:
        cur = []
        for f in (1, 2, 3):
            rows.append([x, f])
            cur.append(k)
            k += 1
        idx.append(cur)
    return np.array(rows, dtype=float), idx


def initial_conditions_and_prior(labels, v, inf_prior=True):
    """FitBranchingModel.GetInitialConditionsAndPrior (FitBranchingModel.py:225-257). Seeds numpy with 42."""
    np.random.seed(42)
    N = labels.size
    phi0 = np.ones((N, 2)) * 0.5
    phi0[:, 0] = np.random.rand(N)
    phi0[:, 1] = 1 - phi0[:, 0]
    prior = np.ones((N, 2)) * 0.5
    for i in range(N):
        br = int(labels[i]) - 2
        if br == -1:
            prior[i, :] = 0.5
        else:
            if inf_prior:
                prior[i, :] = 1 - v
                prior[i, br] = v
            phi0[i, br] = 0.5 + np.random.random() / 2.0
            phi0[i, 1 - br] = 1 - phi0[i, br]
    return phi0, prior


def expand_prior_bgp(prior2, eps=1e-6):
    """pZ_construction_singleBP.expand_pZ0Zeros (pZ_construction_singleBP.py:6-15)."""
    N = prior2.shape[0]
    r = np.zeros((N, 3 * N)) + eps
    for i in range(N):
        r[i, 3 * i + 1:3 * i + 3] = prior2[i]
    return r


def prior_with_trunk_bgp(eZ0, b, t, eps=1e-6):
    """pZ_construction_singleBP.expand_pZ0PureNumpyZeros (pZ_construction_singleBP.py:18-25)."""
    r = eZ0.copy()
    for i in np.flatnonzero(t <= b):
        r[i, 3 * i] = 1
        r[i, 3 * i + 1] = eps
        r[i, 3 * i + 2] = eps
    return r


def init_logphi_bgp(phi0, t, b):
    """AssignGP.InitialiseVariationalPhi (assigngp_dense.py:92-128)."""
    N = t.size
    out = -9.0 * np.ones((N, 3 * N))
    eps = 1e-9
    for i in range(N):
        if t[i] > b:
            blk = np.hstack([eps, phi0[i, :] - eps])
        else:
            blk = np.array([1 - 2 * eps, eps, eps])
        out[i, 3 * i:3 * i + 3] = np.log(blk)
    return out


def expand_prior_mbgp(prior3, eps=1e-6):
    """MBGP expand_pZ0Zeros (MBGP/assigngp.py:25-40): three-column prior, trunk column explicit."""
    N = prior3.shape[0]
    r = np.zeros((N, 3 * N)) + eps
    for i in range(N):
        r[i, 3 * i:3 * i + 3] = prior3[i]
    return r


def init_logphi_mbgp(phi0, t):
    """MBGP AssignGP.InitialiseVariationalPhi (MBGP/assigngp.py:369-405): always the branch form, + phiTrunk."""
    N = t.size
    eps = 1e-9
    out = -9.0 * np.ones((N, 3 * N))
    trunk = np.zeros((N, 3 * N))
    for i in range(N):
        out[i, 3 * i:3 * i + 3] = np.log(np.hstack([eps, phi0[i, :] - eps]))
        trunk[i, 3 * i:3 * i + 3] = np.array([1 - 2 * eps, eps, eps])
    return out, trunk
