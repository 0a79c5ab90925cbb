"""numpy / SciPy-LAPACK float64 forward + hand-derived adjoint of the AssignGP bounds.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  PARITY UNPINNED: oracle = restatement of
TF 2.4.0 / GPflow 2.1.2 semantics.  Independent of ``oracle.bgp_torch`` (different kernel formula,
different linear-algebra backend, analytic instead of automatic differentiation); the two must agree.

Reference lines restated: assigngp_dense.py:151-199,242-245; assigngp_denseSparse.py:59-107;
branch_kernParamGPflow.py:88-204; MBGP/assigngp.py:103-191,217-251,492-558,608-647,688-691.
"""
import math

import numpy as np
from scipy.linalg import cholesky, solve_triangular

JITTER = 1e-6
SQ_A = 1.0 - 2e-6
SQ_B = 1e-6
MATERN32 = 0
SQEXP = 1
_S3 = math.sqrt(3.0)


# ------------------------------------------------------------------ base kernel and derivatives
def k_base(d, ell, var, kind):
    """k(t,t') as a function of d = t - t' (direct-difference form)."""
    r = np.abs(d) / ell
    if kind == SQEXP:
        return var * np.exp(-0.5 * r * r)
    return var * (1.0 + _S3 * r) * np.exp(-_S3 * r)


def dk_dell(d, ell, var, kind):
    r = np.abs(d) / ell
    if kind == SQEXP:
        return var * np.exp(-0.5 * r * r) * r * r / ell
    return var * 3.0 * r * r * np.exp(-_S3 * r) / ell


def dk_dd(d, ell, var, kind):
    """d k / d d  (d = t - t')."""
    r = np.abs(d) / ell
    if kind == SQEXP:
        return -var * np.exp(-0.5 * r * r) * d / (ell * ell)
    return -3.0 * var * r * np.exp(-_S3 * r) * np.sign(d) / ell


def branch_K(X, X2, b, ell, var, kind):
    """BranchKernelParam.K, one branch point (branch_kernParamGPflow.py:117-198)."""
    t1, f1 = X[:, 0], X[:, 1]
    t2, f2 = X2[:, 0], X2[:, 1]
    Ktt = k_base(t1[:, None] - t2[None, :], ell, var, kind)
    ka = k_base(t1 - b, ell, var, kind)
    kb = k_base(t2 - b, ell, var, kind)
    kbb = var + JITTER
    cross = np.outer(ka, kb) / kbb
    return np.where(f1[:, None] == f2[None, :], Ktt, cross)


def branch_K_grads(Kbar, X, X2, b, ell, var, kind):
    """(d/d ell, d/d var, d/d b) of sum(Kbar * branch_K(X, X2))."""
    t1, f1 = X[:, 0], X[:, 1]
    t2, f2 = X2[:, 0], X2[:, 1]
    same = f1[:, None] == f2[None, :]
    d12 = t1[:, None] - t2[None, :]
    Ktt = k_base(d12, ell, var, kind)
    ka, kb = k_base(t1 - b, ell, var, kind), k_base(t2 - b, ell, var, kind)
    dka, dkb = dk_dell(t1 - b, ell, var, kind), dk_dell(t2 - b, ell, var, kind)
    kbb = var + JITTER
    cross = np.outer(ka, kb) / kbb
    dK_l = np.where(same, dk_dell(d12, ell, var, kind), (np.outer(dka, kb) + np.outer(ka, dkb)) / kbb)
    dK_v = np.where(same, Ktt / var, cross * (2.0 / var - 1.0 / kbb))
    # d/db: k(t - b) -> -dk/dd
    dka_b, dkb_b = -dk_dd(t1 - b, ell, var, kind), -dk_dd(t2 - b, ell, var, kind)
    dK_b = np.where(same, 0.0, (np.outer(dka_b, kb) + np.outer(ka, dkb_b)) / kbb)
    return float((Kbar * dK_l).sum()), float((Kbar * dK_v).sum()), float((Kbar * dK_b).sum())


# ------------------------------------------------------------------ small helpers
def softmax_rows(x):
    m = x.max(axis=1, keepdims=True)
    e = np.exp(x - m)
    return e / e.sum(axis=1, keepdims=True)


def chol_adjoint(Lc, Lbar):
    """Reverse-mode Cholesky: Kbar = 1/2 (S + S^T), S = Lc^-T Phi(Lc^T tril(Lbar)) Lc^-1."""
    T = np.tril(Lc.T @ np.tril(Lbar))
    T[np.diag_indices_from(T)] *= 0.5
    S = solve_triangular(Lc, T, lower=True, trans="T")          # Lc^-T T
    S = solve_triangular(Lc, S.T, lower=True, trans="T").T      # (.) Lc^-1
    return 0.5 * (S + S.T)


def _phi_backward(S, Phi, pZ, Abar, Y, vbar, kl_weight, active=None):
    """dELBO/dlogPhi from dELBO/dA (Abar[M]), dELBO/dv (vbar[M,D]) and the KL term."""
    Phibar = Abar[None, :] + Y @ vbar.T - kl_weight * (np.log(Phi) + 1.0 - np.log(pZ))
    Sbar = SQ_A * Phibar
    if active is not None:
        Sbar = Sbar * active[:, None]
    return S * (Sbar - (S * Sbar).sum(axis=1, keepdims=True))


# ------------------------------------------------------------------ dense bound (A.1, B.1)
def dense_forward(logPhi, Y, X, pZ, b, ell, var, likvar, kind=MATERN32, square_noise=1e-6, KConst=None,
                  kl_weight=1.0, Phi_override=None, want_state=False):
    N, D = Y.shape
    M = X.shape[0]
    s = likvar
    S = softmax_rows(logPhi)
    Phi = SQ_A * S + SQ_B if Phi_override is None else Phi_override
    A = Phi.sum(axis=0)
    Delta = A / s
    if KConst is not None:
        K = np.asarray(KConst, dtype=np.float64)
    else:
        K = branch_K(X, X, b, ell, var, kind) + square_noise * np.eye(M)
    Lc = cholesky(K, lower=True)
    L = Lc + JITTER * np.eye(M)
    P = L.T @ (Delta[:, None] * L) + np.eye(M)
    R = cholesky(P, lower=True)
    v = Phi.T @ Y
    z = L.T @ v
    c = solve_triangular(R, z, lower=True) / s
    KL = (Phi * np.log(Phi)).sum() - (Phi * np.log(pZ)).sum()
    terms = dict(a1=-0.5 * N * D * math.log(2.0 * math.pi * s),
                 a2=-0.5 * D * np.log(np.diag(R) ** 2).sum(),
                 a3=-0.5 * (Y * Y).sum() / s,
                 a4=0.5 * (c * c).sum(),
                 a5=-kl_weight * KL)
    elbo = terms["a1"] + terms["a2"] + terms["a3"] + terms["a4"] + terms["a5"]
    if not want_state:
        return elbo
    return elbo, dict(S=S, Phi=Phi, A=A, Delta=Delta, K=K, Lc=Lc, L=L, P=P, R=R, v=v, z=z, c=c, KL=KL, terms=terms)


def dense_value_and_grad(logPhi, Y, X, pZ, b, ell, var, likvar, kind=MATERN32, square_noise=1e-6, kl_weight=1.0,
                         Phi_override=None, active=None, want_state=False):
    """ELBO and gradient wrt (logPhi, ell, var, likvar, b) -- all constrained values."""
    N, D = Y.shape
    s = likvar
    elbo, st = dense_forward(logPhi, Y, X, pZ, b, ell, var, likvar, kind, square_noise, None, kl_weight,
                             Phi_override, want_state=True)
    L, Lc, R, c, v, A, Delta = st["L"], st["Lc"], st["R"], st["c"], st["v"], st["A"], st["Delta"]
    M = L.shape[0]
    q = solve_triangular(R, c, lower=True, trans="T")
    Rinv = solve_triangular(R, np.eye(M), lower=True)
    Pinv = Rinv.T @ Rinv
    Pbar = -0.5 * D * Pinv - 0.5 * (q @ q.T)
    zbar = q / s
    G = L @ Pbar
    Lbar = 2.0 * Delta[:, None] * G + v @ zbar.T
    delta = (G * L).sum(axis=1)
    Abar = delta / s
    sbar = -0.5 * N * D / s + 0.5 * (Y * Y).sum() / s ** 2 - (c * c).sum() / s - (A * delta).sum() / s ** 2
    vbar = L @ zbar
    Kbar = chol_adjoint(Lc, Lbar)
    g_ell, g_var, g_b = branch_K_grads(Kbar, X, X, b, ell, var, kind)
    g_logPhi = _phi_backward(st["S"], st["Phi"], pZ, Abar, Y, vbar, kl_weight, active)
    grads = dict(logPhi=g_logPhi, ell=g_ell, var=g_var, likvar=float(sbar), b=g_b)
    if want_state:
        st.update(q=q, Pinv=Pinv, Pbar=Pbar, G=G, Lbar=Lbar, delta=delta, Kbar=Kbar, vbar=vbar)
        return elbo, grads, st
    return elbo, grads


# ------------------------------------------------------------------ sparse bound (A.2, B.2)
def _sparse_core(Phi, pZ, Y, X, Z, b, ell, var, s, kind, square_noise, kl_weight, D_logdet, want_grad):
    """Returns (trace + D_logdet*a2 + a4 - kl_weight*KL, partial grads)."""
    N, D = Y.shape
    Mz = Z.shape[0]
    A = Phi.sum(axis=0)
    Delta = A / s
    Kuu = branch_K(Z, Z, b, ell, var, kind) + (square_noise + JITTER) * np.eye(Mz)
    Kuf = branch_K(Z, X, b, ell, var, kind)
    Kdiag = var + square_noise
    L = cholesky(Kuu, lower=True)
    V = solve_triangular(L, Kuf, lower=True)
    P = (V * Delta[None, :]) @ V.T + np.eye(Mz)
    R = cholesky(P, lower=True)
    v = Phi.T @ Y
    z = V @ v
    c = solve_triangular(R, z, lower=True) / s
    trace = -0.5 * Kdiag * A.sum() / s + 0.5 * ((V * V) * Delta[None, :]).sum()
    KL = (Phi * np.log(Phi)).sum() - (Phi * np.log(pZ)).sum()
    a2 = -0.5 * np.log(np.diag(R) ** 2).sum()
    a4 = 0.5 * (c * c).sum()
    val = trace + D_logdet * a2 + a4 - kl_weight * KL
    if not want_grad:
        return val, None
    q = solve_triangular(R, c, lower=True, trans="T")
    Rinv = solve_triangular(R, np.eye(Mz), lower=True)
    Pinv = Rinv.T @ Rinv
    Pbar = -0.5 * D_logdet * Pinv - 0.5 * (q @ q.T)
    zbar = q / s
    Vbar = 2.0 * (Pbar @ V) * Delta[None, :] + zbar @ v.T + V * Delta[None, :]
    delta = ((Pbar @ V) * V).sum(axis=0) + 0.5 * (V * V).sum(axis=0) - 0.5 * Kdiag
    Abar = delta / s
    # explicit likvar terms of this block (the -ND/2 log and Y^2 terms are added by the caller)
    sbar = -(c * c).sum() / s - (A * delta).sum() / s ** 2
    vbar = V.T @ zbar
    Kuf_bar = solve_triangular(L, Vbar, lower=True, trans="T")
    Lbar = -np.tril(Kuf_bar @ V.T)
    Kuu_bar = chol_adjoint(L, Lbar)
    gl1, gv1, gb1 = branch_K_grads(Kuu_bar, Z, Z, b, ell, var, kind)
    gl2, gv2, gb2 = branch_K_grads(Kuf_bar, Z, X, b, ell, var, kind)
    g_var = gv1 + gv2 - 0.5 * Delta.sum()
    return val, dict(Abar=Abar, vbar=vbar, sbar=float(sbar), ell=gl1 + gl2, var=g_var, b=gb1 + gb2)


def sparse_value_and_grad(logPhi, Y, X, Z, pZ, b, ell, var, likvar, kind=MATERN32, square_noise=1e-6,
                          want_grad=True):
    """AssignGPSparse bound (assigngp_denseSparse.py:59-107) and gradient."""
    N, D = Y.shape
    s = likvar
    S = softmax_rows(logPhi)
    Phi = SQ_A * S + SQ_B
    val, g = _sparse_core(Phi, pZ, Y, X, Z, b, ell, var, s, kind, square_noise, 1.0, float(D), want_grad)
    elbo = val - 0.5 * N * D * math.log(2.0 * math.pi * s) - 0.5 * (Y * Y).sum() / s
    if not want_grad:
        return elbo
    sbar = g["sbar"] - 0.5 * N * D / s + 0.5 * (Y * Y).sum() / s ** 2
    g_logPhi = _phi_backward(S, Phi, pZ, g["Abar"], Y, g["vbar"], 1.0)
    return elbo, dict(logPhi=g_logPhi, ell=g["ell"], var=g["var"], likvar=float(sbar), b=g["b"])


# ------------------------------------------------------------------ MBGP (A.2b)
def mbgp_masks(logPhi, t, b_d, eZ0, phiTrunk):
    """_GetPhiBeta / _GetPZ (MBGP/assigngp.py:217-251)."""
    S = softmax_rows(logPhi)
    act = t > b_d
    Phi = SQ_A * np.where(act[:, None], S, phiTrunk) + SQ_B
    pZ = SQ_A * np.where(act[:, None], eZ0, phiTrunk) + SQ_B
    return S, act, Phi, pZ


def mbgp_sparse_value_and_grad(logPhi, Y, X, Z, t, eZ0, phiTrunk, Bv, ell, var, likvar, kind=SQEXP,
                               noise_level=1e-6, want_grad=True):
    """MBGP _build_likelihood_sparse (MBGP/assigngp.py:492-558) and gradient wrt logPhi, ell, var, likvar, Bv."""
    N, D = Y.shape
    s = likvar
    acc = 0.0
    g_logPhi = np.zeros_like(logPhi)
    g_ell = g_var = 0.0
    sbar = -0.5 * N * D / s + 0.5 * (Y * Y).sum() / s ** 2
    g_b = np.zeros(D)
    for d in range(D):
        S, act, Phi_d, pZ_d = mbgp_masks(logPhi, t, Bv[d], eZ0, phiTrunk)
        Yd = Y[:, d:d + 1]
        val, g = _sparse_core(Phi_d, pZ_d, Yd, X, Z, Bv[d], ell, var, s, kind, noise_level, 1.0, 1.0, want_grad)
        acc += val
        if want_grad:
            g_logPhi += _phi_backward(S, Phi_d, pZ_d, g["Abar"], Yd, g["vbar"], 1.0, active=act.astype(float))
            g_ell += g["ell"]
            g_var += g["var"]
            sbar += g["sbar"]
            g_b[d] = g["b"]
    elbo = -0.5 * N * D * math.log(2.0 * math.pi * s) - 0.5 * (Y * Y).sum() / s + acc
    if not want_grad:
        return elbo
    return elbo, dict(logPhi=g_logPhi, ell=g_ell, var=g_var, likvar=float(sbar), Bv=g_b)


def mbgp_dense_value_and_grad(logPhi, Y, X, t, eZ0, phiTrunk, Bv, ell, var, likvar, kind=SQEXP, noise_level=1e-6):
    """MBGP _build_likelihood_single_b (MBGP/assigngp.py:608-647): mask / kernel from Bv[0], -D*KL."""
    D = Y.shape[1]
    S, act, Phi, pZ = mbgp_masks(logPhi, t, Bv[0], eZ0, phiTrunk)
    elbo, g = dense_value_and_grad(logPhi, Y, X, pZ, Bv[0], ell, var, likvar, kind, noise_level, kl_weight=float(D),
                                   Phi_override=Phi, active=act.astype(float))
    gb = np.zeros(D)
    gb[0] = g.pop("b")
    g["Bv"] = gb
    return elbo, g


# ------------------------------------------------------------------ predict_f (assigngp_dense.py:201-240, assigngp_denseSparse.py:109-155)
def dense_predict(logPhi, Y, X, Xnew, b, ell, var, likvar, kind=MATERN32, square_noise=1e-6, full_cov=False, Phi_override=None):
    M = X.shape[0]
    s = likvar
    Phi = SQ_A * softmax_rows(logPhi) + SQ_B if Phi_override is None else Phi_override
    K = branch_K(X, X, b, ell, var, kind) + square_noise * np.eye(M)
    L = cholesky(K, lower=True) + JITTER * np.eye(M)
    A = Phi.sum(axis=0)
    P = L.T @ ((A / s)[:, None] * L) + np.eye(M)
    R = cholesky(P, lower=True)
    c = solve_triangular(R, L.T @ (Phi.T @ Y), lower=True) / s
    Kus = branch_K(X, Xnew, b, ell, var, kind)
    tmp1 = solve_triangular(L, Kus, lower=True)
    tmp2 = solve_triangular(R, tmp1, lower=True)
    mean = tmp2.T @ c
    if full_cov:
        v = branch_K(Xnew, Xnew, b, ell, var, kind) + square_noise * np.eye(Xnew.shape[0]) + tmp2.T @ tmp2 - tmp1.T @ tmp1
    else:
        v = (var + square_noise) + (tmp2 ** 2).sum(0) - (tmp1 ** 2).sum(0)
    return mean, v


def sparse_predict(logPhi, Y, X, Z, Xnew, b, ell, var, likvar, kind=MATERN32, square_noise=1e-6, full_cov=False):
    Mz = Z.shape[0]
    s = likvar
    Phi = SQ_A * softmax_rows(logPhi) + SQ_B
    Kuu = branch_K(Z, Z, b, ell, var, kind) + (square_noise + JITTER) * np.eye(Mz)
    Kuf = branch_K(Z, X, b, ell, var, kind)
    L = cholesky(Kuu, lower=True)
    A = Phi.sum(axis=0)
    V = solve_triangular(L, Kuf, lower=True)
    P = (V * (A / s)[None, :]) @ V.T + np.eye(Mz)
    R = cholesky(P, lower=True)
    c = solve_triangular(R, V @ (Phi.T @ Y), lower=True) / s
    Kus = branch_K(Z, Xnew, b, ell, var, kind)
    tmp1 = solve_triangular(L, Kus, lower=True)
    tmp2 = solve_triangular(R, tmp1, lower=True)
    mean = tmp2.T @ c
    if full_cov:
        v = branch_K(Xnew, Xnew, b, ell, var, kind) + square_noise * np.eye(Xnew.shape[0]) + tmp2.T @ tmp2 - tmp1.T @ tmp1
    else:
        v = (var + square_noise) + (tmp2 ** 2).sum(0) - (tmp1 ** 2).sum(0)
    return mean, v
