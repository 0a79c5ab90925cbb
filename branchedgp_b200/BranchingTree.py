"""Branching-tree bookkeeping reduced to what the single-branch-point models need.

The reference's BranchingTree.py (465 lines) builds, for an arbitrary binary tree of branch points, the F x F x B tensor
``fm`` of branch-point ids shared by each pair of functions (BranchingTree.py:277-371).  Every model in the reference
uses exactly one branch point (FitBranchingModel.py:60-62; MBGP/assigngp.py:846-877), for which the tensor is the constant
``fm[i, j, 0] = 1`` for i != j and NaN on the diagonal, F = 3.  That case is what is provided here, together with the
function domains and the domain-restricted input expansion of that one-branch-point tree.
"""
import numpy as np


class BinaryBranchingTree:
    def __init__(self, lb, ub, fDebug=False):
        self.lb, self.ub, self.fDebug = lb, ub, fDebug
        self.branch_points = []

    def add(self, parent, idB, val):
        if self.branch_points:
            raise NotImplementedError("only a single branching point is supported (all reference models use one)")
        self.branch_points.append((idB, np.asarray(val, dtype=np.float64)))

    def GetFunctionBranchTensor(self):
        """(fm, fmb): fm[i, j, 0] = id of the branch point functions i and j split at (NaN for i == j)."""
        assert len(self.branch_points) == 1, "add the branching point first"
        idB = self.branch_points[0][0]
        fm = np.full((3, 3, 1), np.nan)
        for i in range(3):
            for j in range(3):
                if i != j:
                    fm[i, j, 0] = idB
        fmb = fm.copy()
        return fm, fmb

    def GetFunctionDomains(self):
        """[3, 2] array of (lower, upper] domains: trunk (lb, b], both branches (b, ub]  (BranchingTree.py:427-465)."""
        assert len(self.branch_points) == 1, "add the branching point first"
        b = float(np.ravel(self.branch_points[0][1])[0])
        return np.array([[self.lb, b], [b, self.ub], [b, self.ub]], dtype=np.float64)

    def GetFunctionIndexList(self, Xin, fReturnXtrue=False):
        """Each x is expanded only over the functions whose domain (lower, upper] contains it (BranchingTree.py:384-425):
        the trunk for x <= b, the two branches for x > b.  One ``np.random.choice`` per input, as in the reference; values at
        or below the lower bound or above the upper bound raise NameError."""
        Xin = np.asarray(Xin, dtype=np.float64)
        assert Xin.shape[0] == np.size(Xin)
        x = Xin.reshape(-1)
        dom = self.GetFunctionDomains()
        if np.any(x <= dom[0, 0]):
            raise NameError("Value passed in at or less than lower bound " + str(dom[0, 0]))
        if np.any(x > dom[-1, 1]):
            raise NameError("Value passed in greater than upper bound " + str(dom[-1, 1]))
        member = (x[:, None] > dom[None, :, 0]) & (x[:, None] <= dom[None, :, 1])   # [n, 3]
        counts = member.sum(axis=1)
        starts = np.concatenate(([0], np.cumsum(counts)[:-1]))
        rows, funcs = np.nonzero(member)                      # row-major: inputs in order, functions ascending
        Xnew = np.stack([x[rows], funcs + 1.0], axis=1)
        indices = [list(range(int(s0), int(s0 + c))) for s0, c in zip(starts, counts)]
        Xtrue = np.zeros((x.size, 2))
        Xtrue[:, 0] = x
        for i in range(x.size):                               # same draws from the global RNG as the reference
            Xtrue[i, 1] = np.random.choice(list(np.flatnonzero(member[i]))) + 1
        assert np.allclose(Xnew[[ix[0] for ix in indices], 0], x)
        return (Xnew, indices, Xtrue) if fReturnXtrue else (Xnew, indices)
