"""Op-for-op float64 restatement of the reference's TF/GPflow hot path on torch-CPU (autograd gradients).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  PARITY UNPINNED: oracle = restatement of
TF 2.4.0 / GPflow 2.1.2 semantics, not the real thing.

Each function cites the reference lines it restates (paths relative to /root/reference).
Third-party semantics restated from the published GPflow 2.1.2 / TFP 0.11.1 behaviour:

* ``gpflow.kernels.Matern32``:  v (1 + sqrt3 r) exp(-sqrt3 r),  r = sqrt(max(r2, 1e-36))
* ``gpflow.kernels.SquaredExponential``:  v exp(-r2 / 2)
* ``r2`` = |x/l|^2 + |x'/l|^2 - 2 (x/l).(x'/l)   (expanded form used by ``square_distance``)
* ``gpflow.kernels.White``:  K(X) = variance * I,  K(X, X2) = 0,  K_diag = variance
* ``gpflow.default_jitter()`` = 1e-6
"""
import math

import numpy as np
import torch

JITTER = 1e-6          # gpflow.default_jitter()
SQUASH_A = 1.0 - 2e-6  # assigngp_dense.py:162
SQUASH_B = 1e-6

MATERN32 = 0
SQEXP = 1


def _t(x):
    return torch.as_tensor(x, dtype=torch.float64)


def base_kernel(x1, x2, ell, var, kind):
    """gpflow stationary kernel K(x1, x2) on column vectors [n,1], [m,1] (GPflow 2.1.2 semantics)."""
    a = x1 / ell
    b = x2 / ell
    r2 = -2.0 * (a @ b.T) + ((a * a).sum(-1)[:, None] + (b * b).sum(-1)[None, :])
    if kind == SQEXP:
        return var * torch.exp(-0.5 * r2)
    r = torch.sqrt(torch.clamp(r2, min=1e-36))
    s3 = math.sqrt(3.0)
    return var * (1.0 + s3 * r) * torch.exp(-s3 * r)


def branch_K(X, X2, b, ell, var, kind):
    """BranchKernelParam.K for one branch point (branch_kernParamGPflow.py:88-199; MBGP/assigngp.py:103-187).

    ``fm[fi,fj,0] = 1`` for every ordered pair fi != fj (BranchingTree.py:277-371 with one branch point),
    so each off-function block is k(t,b) inv(k(b,b)+jitter) k(b,t')  (:168-196).
    """
    t1, f1 = X[:, 0:1], X[:, 1]
    t2, f2 = X2[:, 0:1], X2[:, 1]
    Ktt = base_kernel(t1, t2, ell, var, kind)
    Bs = b.reshape(1, 1)
    kbb = base_kernel(Bs, Bs, ell, var, kind) + JITTER
    kbb_inv = torch.linalg.inv(kbb)
    kb1 = base_kernel(t1, Bs, ell, var, kind)
    kb2 = base_kernel(t2, Bs, ell, var, kind)
    Kcross = (kb1 @ kbb_inv) @ kb2.T
    same = f1[:, None] == f2[None, :]
    return torch.where(same, Ktt, Kcross)


def squash(Phi):
    return SQUASH_A * Phi + SQUASH_B


def dense_bound(logPhi, Y, X, pZ, b, ell, var, likvar, kind=MATERN32, square_noise=1e-6, KConst=None,
                kl_weight=1.0, Phi_override=None):
    """AssignGP.maximum_log_likelihood_objective, dense (assigngp_dense.py:151-199, build_KL :242-245).

    ``square_noise`` is the White(1e-6) summand of the Sum kernel (FitBranchingModel.py:64-70); MBGP's
    dense variant uses ``noise_level`` 1e-6 on the square kernel instead (MBGP/assigngp.py:180-187) and
    ``kl_weight = D`` (:637).
    """
    N, D = Y.shape
    M = X.shape[0]
    if KConst is not None:
        K = _t(KConst)
    else:
        K = branch_K(X, X, b, ell, var, kind) + square_noise * torch.eye(M, dtype=torch.float64)
    Phi = squash(torch.softmax(logPhi, dim=1)) if Phi_override is None else Phi_override
    L = torch.linalg.cholesky(K) + JITTER * torch.eye(M, dtype=torch.float64)
    A = Phi.sum(0)
    W = L.T * torch.sqrt(A) / torch.sqrt(likvar)
    P = W @ W.T + torch.eye(M, dtype=torch.float64)
    R = torch.linalg.cholesky(P)
    LPhiY = L.T @ (Phi.T @ Y)
    c = torch.linalg.solve_triangular(R, LPhiY, upper=False) / likvar
    KL = (Phi * torch.log(Phi)).sum() - (Phi * torch.log(pZ)).sum()
    tau = 1.0 / likvar
    a1 = -0.5 * N * D * torch.log(2.0 * math.pi / tau)
    a2 = -0.5 * D * torch.log(torch.diagonal(R) ** 2).sum()
    a3 = -0.5 * (Y * Y).sum() / likvar
    a4 = 0.5 * (c * c).sum()
    return a1 + a2 + a3 + a4 - kl_weight * KL


def sparse_bound_core(Phi, pZ, Y, X, Z, b, ell, var, likvar, kind, square_noise):
    """Body shared by AssignGPSparse (assigngp_denseSparse.py:59-107) and one output of
    MBGP._build_likelihood_sparse (MBGP/assigngp.py:512-543).  Returns (a2 + a4 - KL + trace) pieces."""
    Mz = Z.shape[0]
    eye = torch.eye(Mz, dtype=torch.float64)
    Kuu = branch_K(Z, Z, b, ell, var, kind) + square_noise * eye + JITTER * eye
    Kuf = branch_K(Z, X, b, ell, var, kind)
    Kdiag = var + square_noise                       # K_diag: branch_kernParamGPflow.py:201-204 (+White / noise_level)
    L = torch.linalg.cholesky(Kuu)
    A = Phi.sum(0)
    V = torch.linalg.solve_triangular(L, Kuf, upper=False)
    W = V * torch.sqrt(A) / torch.sqrt(likvar)
    P = W @ W.T + eye
    trace = -0.5 * (Kdiag * A).sum() / likvar + 0.5 * (W * W).sum()
    R = torch.linalg.cholesky(P)
    tmp = V @ (Phi.T @ Y)
    c = torch.linalg.solve_triangular(R, tmp, upper=False) / likvar
    KL = (Phi * torch.log(Phi)).sum() - (Phi * torch.log(pZ)).sum()
    a2 = -0.5 * torch.log(torch.diagonal(R) ** 2).sum()
    a4 = 0.5 * (c * c).sum()
    return trace, a2, a4, KL


def sparse_bound(logPhi, Y, X, Z, pZ, b, ell, var, likvar, kind=MATERN32, square_noise=1e-6):
    """AssignGPSparse.maximum_log_likelihood_objective (assigngp_denseSparse.py:59-107)."""
    N, D = Y.shape
    Phi = squash(torch.softmax(logPhi, dim=1))
    trace, a2, a4, KL = sparse_bound_core(Phi, pZ, Y, X, Z, b, ell, var, likvar, kind, square_noise)
    return (trace - 0.5 * N * D * torch.log(2 * math.pi * likvar) + D * a2
            - 0.5 * (Y * Y).sum() / likvar + a4 - KL)


def mbgp_phi_beta(logPhi, t, b_d, phiTrunk):
    """MBGP AssignGP._GetPhiBeta (MBGP/assigngp.py:217-233)."""
    S = torch.softmax(logPhi, dim=1)
    return squash(torch.where((t > b_d)[:, None], S, phiTrunk))


def mbgp_pz(eZ0, t, b_d, phiTrunk):
    """MBGP AssignGP._GetPZ (MBGP/assigngp.py:235-251)."""
    return squash(torch.where((t > b_d)[:, None], eZ0, phiTrunk))


def mbgp_sparse_bound(logPhi, Y, X, Z, t, eZ0, phiTrunk, Bv, ell, var, likvar, kind=SQEXP, noise_level=1e-6):
    """MBGP AssignGP._build_likelihood_sparse (MBGP/assigngp.py:492-558)."""
    N, D = Y.shape
    acc = torch.zeros((), dtype=torch.float64)
    for d in range(D):
        b_d = Bv[d]
        Phi_d = mbgp_phi_beta(logPhi, t, b_d, phiTrunk)
        pZ_d = mbgp_pz(eZ0, t, b_d, phiTrunk)
        trace, a2, a4, KL = sparse_bound_core(Phi_d, pZ_d, Y[:, d:d + 1], X, Z, b_d, ell, var, likvar, kind,
                                              noise_level)
        acc = acc + a2 + a4 - KL + trace
    a1 = -0.5 * N * D * torch.log(2.0 * math.pi * likvar)
    a3 = -0.5 * (Y * Y).sum() / likvar
    return a1 + a3 + acc


def mbgp_dense_bound(logPhi, Y, X, t, eZ0, phiTrunk, Bv, ell, var, likvar, kind=SQEXP, noise_level=1e-6):
    """MBGP AssignGP._build_likelihood_single_b (MBGP/assigngp.py:608-647): mask and kernel from Bv[0], -D*KL."""
    D = Y.shape[1]
    b0 = Bv[0]
    Phi = mbgp_phi_beta(logPhi, t, b0, phiTrunk)
    pZ = mbgp_pz(eZ0, t, b0, phiTrunk)
    return dense_bound(logPhi, Y, X, pZ, b0, ell, var, likvar, kind=kind, square_noise=noise_level,
                       kl_weight=float(D), Phi_override=Phi)


# ----------------------------------------------------------------------------------------------------
# value + gradient helpers (gradients with respect to the CONSTRAINED hyper-parameters and logPhi)
# ----------------------------------------------------------------------------------------------------
def _leaf(x):
    return _t(np.asarray(x, dtype=np.float64)).clone().requires_grad_(True)


def dense_value_and_grad(logPhi, Y, X, pZ, b, ell, var, likvar, kind=MATERN32, square_noise=1e-6):
    lp, l_, v_, s_ = _leaf(logPhi), _leaf(ell), _leaf(var), _leaf(likvar)
    f = dense_bound(lp, _t(Y), _t(X), _t(pZ), _t(b), l_, v_, s_, kind, square_noise)
    f.backward()
    return f.item(), dict(logPhi=lp.grad.numpy(), ell=l_.grad.item(), var=v_.grad.item(), likvar=s_.grad.item())


def sparse_value_and_grad(logPhi, Y, X, Z, pZ, b, ell, var, likvar, kind=MATERN32, square_noise=1e-6):
    lp, l_, v_, s_ = _leaf(logPhi), _leaf(ell), _leaf(var), _leaf(likvar)
    f = sparse_bound(lp, _t(Y), _t(X), _t(Z), _t(pZ), _t(b), l_, v_, s_, kind, square_noise)
    f.backward()
    return f.item(), dict(logPhi=lp.grad.numpy(), ell=l_.grad.item(), var=v_.grad.item(), likvar=s_.grad.item())


def mbgp_sparse_value_and_grad(logPhi, Y, X, Z, t, eZ0, phiTrunk, Bv, ell, var, likvar, kind=SQEXP,
                               noise_level=1e-6):
    lp, l_, v_, s_, b_ = _leaf(logPhi), _leaf(ell), _leaf(var), _leaf(likvar), _leaf(Bv)
    f = mbgp_sparse_bound(lp, _t(Y), _t(X), _t(Z), _t(t), _t(eZ0), _t(phiTrunk), b_, l_, v_, s_, kind, noise_level)
    f.backward()
    return f.item(), dict(logPhi=lp.grad.numpy(), ell=l_.grad.item(), var=v_.grad.item(), likvar=s_.grad.item(),
                          Bv=b_.grad.numpy())


def mbgp_dense_value_and_grad(logPhi, Y, X, t, eZ0, phiTrunk, Bv, ell, var, likvar, kind=SQEXP, noise_level=1e-6):
    lp, l_, v_, s_, b_ = _leaf(logPhi), _leaf(ell), _leaf(var), _leaf(likvar), _leaf(Bv)
    f = mbgp_dense_bound(lp, _t(Y), _t(X), _t(t), _t(eZ0), _t(phiTrunk), b_, l_, v_, s_, kind, noise_level)
    f.backward()
    return f.item(), dict(logPhi=lp.grad.numpy(), ell=l_.grad.item(), var=v_.grad.item(), likvar=s_.grad.item(),
                          Bv=b_.grad.numpy())
