"""Branching-tree bookkeeping reduced to what the single-branch-point models need.

The reference's BranchingTree.py (465 lines) builds, for an arbitrary binary tree of branch points, the F x F x B tensor
``fm`` of branch-point ids shared by each pair of functions (BranchingTree.py:277-371).  Every model in the reference
uses exactly one branch point (FitBranchingModel.py:60-62; MBGP/assigngp.py:846-877), for which the tensor is the constant
``fm[i, j, 0] = 1`` for i != j and NaN on the diagonal, F = 3.  That case is what is provided here.
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

    def GetFunctionIndexList(self, Xin, fReturnXtrue=False):
        from .VBHelperFunctions import GetFunctionIndexListGeneral
        X, idx, Xs = GetFunctionIndexListGeneral(Xin)
        return (X, idx, Xs) if fReturnXtrue else (X, idx)
