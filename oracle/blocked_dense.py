"""Block-level numpy emulator of the CUDA dense pipeline (branchedgp_b200/csrc/bgp_dense.cu).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  It restates, with plain numpy block products, the exact
sequence of *gather-form* block operations the per-CTA CUDA pipeline performs (left-looking Cholesky with
explicit diagonal-block inverses, in-place right-to-left trtri, lauum, lower-only trmm, and the gather form of
Murray's blocked Cholesky adjoint) on identity-padded matrices.  Its only purpose is to validate those
derivations against ``oracle.bgp_numpy`` on the CPU, where there is no GPU to run the kernels.
"""
import math

import numpy as np

from . import bgp_numpy as ref


def _blk(A, I, J, nb):
    return A[I * nb:(I + 1) * nb, J * nb:(J + 1) * nb]


def chol_left_looking(A, nb):
    """Lower Cholesky of SPD A (Mp x Mp, Mp % nb == 0).  Returns (L, [inv(L_JJ)])."""
    Mp = A.shape[0]
    n = Mp // nb
    L = np.zeros_like(A)
    invd = []
    for J in range(n):
        acc = _blk(L, J, 0, nb)[:, :0]
        C = _blk(A, J, J, nb) - L[J * nb:(J + 1) * nb, :J * nb] @ L[J * nb:(J + 1) * nb, :J * nb].T
        LJJ = np.linalg.cholesky(C)
        Linv = np.linalg.inv(LJJ)
        invd.append(Linv)
        _blk(L, J, J, nb)[:] = LJJ
        for I in range(J + 1, n):
            C = _blk(A, I, J, nb) - L[I * nb:(I + 1) * nb, :J * nb] @ L[J * nb:(J + 1) * nb, :J * nb].T
            _blk(L, I, J, nb)[:] = C @ Linv.T
    return L, invd


def trtri_inplace(R, invd, nb):
    """X = inv(R) (lower), right-to-left gather form, in place over a copy of R."""
    X = R.copy()
    n = R.shape[0] // nb
    for J in range(n - 1, -1, -1):
        for I in range(n - 1, J, -1):
            # acc = sum_{J<k<=I} X[I,k] R[k,J]  (X of later columns is final; R[k,J] of this column still original)
            acc = X[I * nb:(I + 1) * nb, (J + 1) * nb:(I + 1) * nb] @ X[(J + 1) * nb:(I + 1) * nb, J * nb:(J + 1) * nb]
            _blk(X, I, J, nb)[:] = -acc @ invd[J]
        _blk(X, J, J, nb)[:] = invd[J]
    return X


def lauum(X, nb):
    """lower(X^T X) for lower-triangular X; diagonal blocks returned fully symmetric."""
    n = X.shape[0] // nb
    out = np.zeros_like(X)
    for I in range(n):
        for J in range(I + 1):
            _blk(out, I, J, nb)[:] = X[I * nb:, I * nb:(I + 1) * nb].T @ X[I * nb:, J * nb:(J + 1) * nb]
    return out


def sym_from_lower(AL, nb):
    """Full symmetric matrix from block-lower storage whose diagonal blocks are already symmetric."""
    n = AL.shape[0] // nb
    F = AL.copy()
    for I in range(n):
        for J in range(I + 1, n):
            _blk(F, I, J, nb)[:] = _blk(AL, J, I, nb).T
    return F


def chol_adjoint_gather(Lc, invd, Lbar, nb):
    """Gather form of the blocked reverse-mode Cholesky.  Returns At (block lower storage) with
    off-diagonal entries (i>j) = 2*Kbar_ij and diagonal blocks = S+S^T (so diagonal entries = 2*Kbar_ii)."""
    n = Lc.shape[0] // nb
    At = np.tril(Lbar).copy()
    for J in range(n - 1, -1, -1):
        lo, hi = J * nb, (J + 1) * nb
        Afull = sym_from_lower(At, nb)          # only columns > J of the symmetric matrix are read below
        for I in range(n - 1, J, -1):
            E = _blk(At, I, J, nb) - Afull[I * nb:(I + 1) * nb, hi:] @ Lc[hi:, lo:hi]
            _blk(At, I, J, nb)[:] = E @ invd[J]
        T = At[hi:, lo:hi].T @ Lc[hi:, lo:hi]
        LJJ = _blk(Lc, J, J, nb)
        Leff = np.tril(_blk(At, J, J, nb)) - np.tril(T)
        U = np.tril(LJJ.T @ Leff)
        U[np.diag_indices(nb)] *= 0.5
        S = invd[J].T @ U @ invd[J]
        _blk(At, J, J, nb)[:] = S + S.T
    return At


def dense_value_and_grad_blocked(logPhi, Y, X, pZ, b, ell, var, likvar, kind=ref.MATERN32, nb=8,
                                 square_noise=1e-6, kl_weight=1.0):
    """Same contract as oracle.bgp_numpy.dense_value_and_grad, computed with the CUDA pipeline's block algorithm."""
    N, D = Y.shape
    M = X.shape[0]
    Mp = ((M + nb - 1) // nb) * nb
    s = likvar
    S = ref.softmax_rows(logPhi)
    Phi = ref.SQ_A * S + ref.SQ_B
    A = np.zeros(Mp)
    A[:M] = Phi.sum(axis=0)
    v = np.zeros((Mp, D))
    v[:M] = Phi.T @ Y
    KL = (Phi * np.log(Phi)).sum() - (Phi * np.log(pZ)).sum()
    Delta = A / s
    K = np.eye(Mp)
    K[:M, :M] = ref.branch_K(X, X, b, ell, var, kind) + square_noise * np.eye(M)
    Lc, linvd = chol_left_looking(K, nb)
    L = Lc + ref.JITTER * np.eye(Mp)
    P = np.tril(L.T @ (Delta[:, None] * L)) + np.eye(Mp)
    P = P + np.tril(P, -1).T
    R, rinvd = chol_left_looking(P, nb)
    z = L.T @ v
    # forward solve by block rows with the diagonal-block inverses
    n = Mp // nb
    c = np.zeros_like(z)
    for J in range(n):
        c[J * nb:(J + 1) * nb] = rinvd[J] @ (z[J * nb:(J + 1) * nb] - R[J * nb:(J + 1) * nb, :J * nb] @ c[:J * nb])
    c /= s
    q = np.zeros_like(c)
    for J in range(n - 1, -1, -1):
        q[J * nb:(J + 1) * nb] = rinvd[J].T @ (c[J * nb:(J + 1) * nb] - R[(J + 1) * nb:, J * nb:(J + 1) * nb].T @ q[(J + 1) * nb:])
    elbo = (-0.5 * N * D * math.log(2 * math.pi * s) - 0.5 * D * np.log(np.diag(R) ** 2).sum()
            - 0.5 * (Y * Y).sum() / s + 0.5 * (c * c).sum() - kl_weight * KL)
    Xinv = trtri_inplace(R, rinvd, nb)
    Pinv = sym_from_lower(lauum(Xinv, nb), nb)
    Pbar = -0.5 * D * Pinv - 0.5 * (q @ q.T)
    zbar = q / s
    G = np.tril(L @ Pbar)
    Lbar = np.tril(2.0 * Delta[:, None] * G + v @ zbar.T)
    delta = (G * L).sum(axis=1)
    Abar = delta / s
    sbar = -0.5 * N * D / s + 0.5 * (Y * Y).sum() / s ** 2 - (c * c).sum() / s - (A * delta).sum() / s ** 2
    vbar = L @ zbar
    At = chol_adjoint_gather(Lc, linvd, Lbar, nb)
    # hyper-gradient: sum over the stored lower triangle; off-diagonal entries carry 2*Kbar, diagonal entries 2*Kbar too
    W = np.tril(At[:M, :M], -1) + 0.5 * np.diag(np.diag(At[:M, :M]))
    g_ell, g_var, g_b = ref.branch_K_grads(W, X, X, b, ell, var, kind)
    g_logPhi = ref._phi_backward(S, Phi, pZ, Abar[:M], Y, vbar[:M], kl_weight)
    return elbo, dict(logPhi=g_logPhi, ell=g_ell, var=g_var, likvar=float(sbar), b=g_b), dict(At=At, Lc=Lc, R=R)
