"""BranchKernelParam: the branching covariance over (pseudotime, function) pairs (branch_kernParamGPflow.py:60-204).

On the bound's path the covariance is generated inside the CUDA kernels (csrc/bgp_dense.cu: kgen) from the parameters
this object carries (base kernel kind, lengthscale, variance, branching point ``Bv``).  ``K`` / ``K_diag`` below are
host-side numpy conveniences for inspecting the kernel; the models never call them."""
import numpy as np

from . import gpflow_shim as gpflow
from .gpflow_shim import Kernel


class BranchKernelParam(Kernel):
    name = "branch_kernel_param"

    def __init__(self, base_kern, branchPtTensor, b, fDebug=False):
        assert isinstance(b, np.ndarray)
        fm = np.asarray(branchPtTensor)
        assert fm.shape[0] == fm.shape[1]
        assert fm.shape[2] > 0
        if fm.shape[0] != 3 or fm.shape[2] != 1:
            raise NotImplementedError("only one branching point (3 functions) is supported, as in every reference model")
        self.kern = base_kern
        self.fm = fm
        self.fDebug = fDebug
        self.Bv = b

    @property
    def kind(self):
        return self.kern.KIND

    def K(self, X, Y=None):
        X = np.asarray(X, dtype=np.float64)
        Y = X if Y is None else np.asarray(Y, dtype=np.float64)
        b = float(np.asarray(self.Bv).ravel()[0])
        same = X[:, 1][:, None] == Y[:, 1][None, :]
        Ktt = self.kern.K_r(X[:, 0][:, None] - Y[:, 0][None, :])
        kbb = float(self.kern.variance.numpy()) + gpflow.default_jitter()
        cross = np.outer(self.kern.K_r(X[:, 0] - b) / kbb, self.kern.K_r(Y[:, 0] - b))
        return np.where(same, Ktt, cross)

    def K_diag(self, X):
        return self.kern.K_diag(X)
