"""Block-level numpy emulator of the multi-CTA ("wide") dense pipeline (branchedgp_b200/csrc/bgp_dense_wide.cu).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  When few work items are in flight the CUDA library spreads the macro tiles of
ONE item over many CTAs; the factorisations then run as *scatter-form* (right-looking) tile tasks ordered by dependency counters
instead of the gather-form loops of one CTA.  This file restates those task bodies with numpy block products, in the task order
the kernels use, so that the derivations can be validated on the CPU against the gather forms of ``oracle.blocked_dense``
(and through them against ``oracle.bgp_numpy``)."""
import numpy as np

from .blocked_dense import _blk, sym_from_lower


def chol_right_looking(A, nb):
    """Right-looking tile Cholesky with the last update of a tile fused into its factor / panel task.
    Returns (L, [inv(L_JJ)])."""
    n = A.shape[0] // nb
    W = np.tril(A).copy()          # tiles are updated in place; only the block-lower part exists
    invd = [None] * n

    def diag(J):
        C = _blk(W, J, J, nb).copy()
        if J >= 1:
            C -= _blk(W, J, J - 1, nb) @ _blk(W, J, J - 1, nb).T
        LJJ = np.linalg.cholesky(C)
        _blk(W, J, J, nb)[:] = LJJ
        invd[J] = np.linalg.inv(LJJ)

    def panel(I, J):
        C = _blk(W, I, J, nb).copy()
        if J >= 1:
            C -= _blk(W, I, J - 1, nb) @ _blk(W, J, J - 1, nb).T
        _blk(W, I, J, nb)[:] = C @ invd[J].T

    def upd(I, K, J):
        _blk(W, I, K, nb)[:] -= _blk(W, I, J, nb) @ _blk(W, K, J, nb).T

    diag(0)
    for I in range(1, n):
        panel(I, 0)
    for J in range(n - 1):
        diag(J + 1)
        for I in range(J + 2, n):
            panel(I, J + 1)
        for K in range(J + 2, n):
            for I in range(K, n):
                upd(I, K, J)
    return np.tril(W), invd


def trtri_scatter(R, invd, nb):
    """X = inv(R) by block Gauss-Jordan elimination on the row- and column-scaled factor.
    PRE(I,J):  W0[I,J] = -D_I R[I,J] (kept aside),  V[I,J] = W0[I,J] D_J (in place of R);
    TR(I,c;J): V[I,c] += W0[I,J] V[J,c]  for I > J > c, J ascending;   diagonal: D_J."""
    n = R.shape[0] // nb
    V = np.zeros_like(R)
    W0 = np.zeros_like(R)
    for J in range(n):
        _blk(V, J, J, nb)[:] = invd[J]
        for I in range(J + 1, n):
            _blk(W0, I, J, nb)[:] = -invd[I] @ _blk(R, I, J, nb)
            _blk(V, I, J, nb)[:] = _blk(W0, I, J, nb) @ invd[J]
    for J in range(1, n - 1):
        for I in range(J + 1, n):
            for c in range(J):
                _blk(V, I, c, nb)[:] += _blk(W0, I, J, nb) @ _blk(V, J, c, nb)
    return V


def chol_adjoint_scatter(Lc, invd, Lbar, nb):
    """Scatter form of the blocked reverse-mode Cholesky, same output convention as blocked_dense.chol_adjoint_gather.

    Tile (I, J), I > J, of  E = Lbar - Atilde_sym[:, J+1:] Lc[J+1:, J]  collects three kinds of terms:
      k >= I : Atilde_sym[I, k] Lc[k, J] uses the tiles of block column I (rows k >= I), final once column I is done
               -> BROW(I, J), one long product, for J <= I - 2; for J = I - 1 (needed by the very next column) the panel tasks
               of column I leave per-row partials Q_k = Atilde[k, I]^T Lc[k, I-1] that the panel task (I, I-1) sums in order;
      J+1 < k < I : Atilde[I, k] Lc[k, J] -> SC(I, J; k), scattered when column k is final;
      k = J + 1 : fused into the panel task APANEL(I, J) itself.
    The panel tasks of column J also leave P_I = Atilde[I, J]^T Lc[I, J] for the diagonal task to sum in order."""
    n = Lc.shape[0] // nb
    At = np.tril(Lbar).copy()
    P = [None] * n
    Q = [None] * n

    def adiag(J):
        T = np.zeros((nb, nb))
        for I in range(J + 1, n):
            T += P[I]
        LJJ = _blk(Lc, J, J, nb)
        Leff = np.tril(_blk(At, J, J, nb)) - np.tril(T)
        U = np.tril(LJJ.T @ Leff)
        U[np.diag_indices(nb)] *= 0.5
        S = invd[J].T @ U @ invd[J]
        _blk(At, J, J, nb)[:] = S + S.T

    def apanel(I, J):
        if I == J + 1:
            acc = _blk(At, I, I, nb) @ _blk(Lc, I, J, nb)      # diagonal block, stored fully symmetric
            for k in range(I + 1, n):
                acc = acc + Q[k]
        else:
            acc = _blk(At, I, J + 1, nb) @ _blk(Lc, J + 1, J, nb)
        E = _blk(At, I, J, nb) - acc
        _blk(At, I, J, nb)[:] = E @ invd[J]
        P[I] = _blk(At, I, J, nb).T @ _blk(Lc, I, J, nb)
        if J >= 1:
            Q[I] = _blk(At, I, J, nb).T @ _blk(Lc, I, J - 1, nb)

    def brow(I, J):
        acc = _blk(At, I, I, nb) @ _blk(Lc, I, J, nb)
        for k in range(I + 1, n):
            acc = acc + _blk(At, k, I, nb).T @ _blk(Lc, k, J, nb)
        _blk(At, I, J, nb)[:] -= acc

    def scatter(I, c, K):
        _blk(At, I, c, nb)[:] -= _blk(At, I, K, nb) @ _blk(Lc, K, c, nb)

    adiag(n - 1)
    for K in range(n - 1, 0, -1):
        for I in range(K, n):
            apanel(I, K - 1)
        adiag(K - 1)
        for J in range(K - 2, -1, -1):
            brow(K, J)
        for c in range(K - 2, -1, -1):
            for I in range(K + 1, n):
                scatter(I, c, K)
    return At
