"""AssignGP: GP regression with unknown assignment of cells to latent functions, dense bound (assigngp_dense.py:9-245).

Same constructor, attributes and methods as the reference class; the bound (``maximum_log_likelihood_objective``,
assigngp_dense.py:151-199) and its gradient are evaluated by the CUDA library through ``_lib.Engine``.
"""
import numpy as np

from . import _lib, pZ_construction_singleBP
from . import gpflow_shim as gpflow
from .gpflow_shim import Parameter


def _check_cell_major(t, XExpanded):
    X = np.asarray(XExpanded, dtype=np.float64)
    N = t.shape[0]
    ok = X.shape == (3 * N, 2) and np.array_equal(X[:, 0], np.repeat(t, 3)) and np.array_equal(X[:, 1], np.tile([1.0, 2.0, 3.0], N))
    if not ok:
        raise NotImplementedError("XExpanded must be the cell-major expansion of t produced by "
                                  "VBHelperFunctions.GetFunctionIndexListGeneral (rows (t_i, 1), (t_i, 2), (t_i, 3))")
    return X


class AssignGP:
    MODEL_KIND = _lib.MODEL_BGP

    def __init__(self, t, XExpanded, Y, kern, indices, b, phiPrior=None, phiInitial=None, fDebug=False, KConst=None):
        assert len(indices) == t.size, "indices must be size N"
        assert len(t.shape) == 1, "pseudotime should be 1D"
        self.kernel = kern
        self.likelihood = gpflow.Gaussian()
        self.num_latent_gps = Y.shape[-1]
        self.Y = np.ascontiguousarray(Y, dtype=np.float64)
        self.t = t.astype(gpflow.default_float())
        self.X = _check_cell_major(self.t, XExpanded)
        self.N = t.shape[0]
        self.indices = indices
        self.logPhi = Parameter(np.random.randn(t.shape[0], t.shape[0] * 3), name="logPhi")
        if phiInitial is None:
            phiInitial = np.ones((self.N, 2)) * 0.5
            phiInitial[:, 0] = np.random.rand(self.N)
            phiInitial[:, 1] = 1 - phiInitial[:, 0]
        self.fDebug = fDebug
        if phiPrior is None:
            phiPrior = np.ones((self.N, 2)) * 0.5
        self._engine = None
        self.UpdateBranchingPoint(b, phiInitial, prior=phiPrior)
        self.KConst = KConst
        if not fDebug:
            assert KConst is None, "KConst only for debugging"

    # ------------------------------------------------------------------ plumbing
    @property
    def engine(self):
        if self._engine is None:
            self._engine = _lib.default_engine()
        return self._engine

    @property
    def pZ(self):
        """Dense N x 3N prior (built on demand; the kernels use the rule, not this matrix)."""
        return pZ_construction_singleBP.expand_pZ0PureNumpyZeros(self.eZ0, self.b, self.t)

    @property
    def eZ0(self):
        return pZ_construction_singleBP.expand_pZ0Zeros(self._prior)

    def _branch_kernel(self):
        k = self.kernel.kernels[0]
        assert k.name == "branch_kernel_param"
        return k

    def _square_noise_ok(self):
        if self.KConst is not None:
            return
        ks = self.kernel.kernels
        if len(ks) != 2 or ks[1].name != "white" or not np.isclose(float(ks[1].variance.numpy()), 1e-6, rtol=1e-6, atol=0):
            raise NotImplementedError("kernel must be BranchKernelParam + White(variance=1e-6), as FitModel builds it")

    # ------------------------------------------------------------------ reference API
    def UpdateBranchingPoint(self, b, phiInitial, prior=None):
        """Update the branching point and reset the variational parameters (assigngp_dense.py:77-90)."""
        assert isinstance(b, np.ndarray)
        assert b.size == 1, "Must have scalar branching point"
        self.b = b.astype(gpflow.default_float())
        k = self._branch_kernel()
        k.Bv = b
        assert self.logPhi.trainable is True, "Phi should not be constant when changing branching location"
        if prior is not None:
            prior = np.asarray(prior, dtype=np.float64)
            pZ_construction_singleBP.expand_pZ0Zeros(prior[:1])   # same shape / sum-to-one assertions as the reference
            for z0 in prior:
                assert z0.sum() == 1, "should sum to 1 is %s=%.3f" % (str(z0), z0.sum())
            self._prior = np.ascontiguousarray(prior)
        self.InitialiseVariationalPhi(phiInitial)

    def InitialiseVariationalPhi(self, phiInitialIn):
        """logPhi <- -9 off the cell's own block; block = log([1e-9, phi - 1e-9]) after the branching point,
        log([1 - 2e-9, 1e-9, 1e-9]) on the trunk (t <= b)   (assigngp_dense.py:92-128)."""
        phiInitialIn = np.asarray(phiInitialIn, dtype=np.float64)
        assert np.allclose(phiInitialIn.sum(1), 1), "probs must sum to 1 %s" % str(phiInitialIn)
        assert self.b == self._branch_kernel().Bv, "Need to call UpdateBranchingPoint"
        N = self.Y.shape[0]
        assert phiInitialIn.shape[0] == N
        assert phiInitialIn.shape[1] == 2
        eps = 1e-9
        block = np.empty((N, 3))
        after = self.t > float(self.b.ravel()[0])
        block[after, 0] = eps
        block[after, 1:] = phiInitialIn[after] - eps
        block[~after] = np.array([1 - 2 * eps, eps, eps])
        assert not np.any(np.isnan(block)), "no nans please"
        assert not np.any(block < -eps), "no negatives please"
        lp = np.full((N, 3 * N), -9.0)
        rows = np.arange(N)
        with np.errstate(divide="ignore", invalid="ignore"):
            lb = np.log(block)
        for f in range(3):
            lp[rows, 3 * rows + f] = lb[:, f]
        self.logPhi.assign(lp)

    def GetPhi(self):
        """Phi collapsed to the cell's own three columns, [N, 3] (assigngp_dense.py:130-140)."""
        assert self.b == self._branch_kernel().Bv, "Need to call UpdateBranchingPoint"
        phi = self.engine.get_phi(self.logPhi.numpy())[0]
        tol = 1e-6
        assert np.all(phi.sum(1) <= 1 + tol)
        assert np.all(phi >= 0 - tol)
        assert np.all(phi <= 1 + tol)
        return phi

    def GetPhiExpanded(self):
        x = self.logPhi.numpy()
        e = np.exp(x - x.max(axis=1, keepdims=True))
        return e / e.sum(axis=1, keepdims=True)

    # ------------------------------------------------------------------ parameters
    def _hyper_parameters(self):
        k = self._branch_kernel().kern
        return [k.lengthscales, k.variance, self.likelihood.variance]

    @property
    def parameters(self):
        return self._hyper_parameters() + [self.kernel.kernels[1].variance, self.logPhi]

    @property
    def trainable_parameters(self):
        return [p for p in self.parameters if p.trainable]

    @property
    def trainable_variables(self):
        return self.trainable_parameters

    def log_prior_density(self):
        s = 0.0
        for p in self.trainable_parameters:
            if p.prior is not None:
                s += float(np.sum(p.prior.log_prob(p.numpy())))
        return s

    # ------------------------------------------------------------------ bound
    def _evaluate(self, want_grad):
        self._square_noise_ok()
        k = self._branch_kernel()
        hyp = np.array([[float(k.kern.lengthscales.numpy()), float(k.kern.variance.numpy()),
                         float(self.likelihood.variance.numpy())]])
        # logits and gradient live in page-locked host memory: the host-pointer entry then moves them by DMA (24 MB each way at
        # N = 1000) instead of through pageable staging copies, and nothing is copied on the Python side
        grad_out = None
        try:
            self.logPhi.pin(_lib.pinned_empty)
            if want_grad:
                if getattr(self, "_grad_pin", None) is None or self._grad_pin.shape != (1,) + self.logPhi.shape:
                    self._grad_pin = _lib.pinned_empty((1,) + tuple(self.logPhi.shape))
                grad_out = self._grad_pin
        except _lib.BgpError:
            grad_out = None
        return self.engine.dense_elbo_grad(self.t, self.Y[None], np.zeros(1, dtype=np.int32), np.array([float(self.b.ravel()[0])]),
                                           self._prior, hyp, self.logPhi.view()[None], kernel=k.kind, model=self.MODEL_KIND,
                                           kl_weight=1.0, KConst=self.KConst, want_grad=want_grad, grad_out=grad_out)

    def maximum_log_likelihood_objective(self):
        return float(self._evaluate(False)["elbo"][0])

    def log_posterior_density(self):
        return self.maximum_log_likelihood_objective() + self.log_prior_density()

    def training_loss(self):
        return -self.log_posterior_density()

    def objectiveFun(self):
        """Objective to minimise, -log likelihood - log prior, as written in the reference (assigngp_dense.py:146-149)."""
        return -self.log_posterior_density() - self.log_prior_density()

    def _loss_and_grad(self, variables, out=None):
        """training_loss and its gradient with respect to the unconstrained value of each parameter in `variables`.
        ``out``: optional list of arrays (one per variable) that receive the gradients in place."""
        out_ev = self._evaluate(True)
        k = self._branch_kernel().kern
        g_constrained = {id(k.lengthscales): out_ev["grad_hyp"][0, 0], id(k.variance): out_ev["grad_hyp"][0, 1],
                         id(self.likelihood.variance): out_ev["grad_hyp"][0, 2], id(self.logPhi): out_ev["grad_logPhi"][0]}
        loss = -(float(out_ev["elbo"][0]) + self.log_prior_density())
        grads = []
        for i, p in enumerate(variables):
            if id(p) not in g_constrained:
                raise ValueError(f"no gradient available for {p!r}")
            g = np.asarray(g_constrained[id(p)], dtype=np.float64).reshape(p.shape)
            plain = p.prior is None and isinstance(p.transform, gpflow.Identity)
            if plain and out is not None:
                np.negative(g, out=out[i])          # one pass, no temporaries
                grads.append(out[i])
                continue
            if p.prior is not None:
                g = g + p.prior.dlog_prob(p.numpy())
            gu = -(g * p.transform.dforward(p.unconstrained_variable))
            if out is not None:
                np.copyto(out[i], gu)
                gu = out[i]
            grads.append(gu)
        return loss, grads

    def _hyp_row(self):
        k = self._branch_kernel()
        return np.array([[float(k.kern.lengthscales.numpy()), float(k.kern.variance.numpy()), float(self.likelihood.variance.numpy())]])

    def predict_f(self, Xnew, full_cov=False):
        """Posterior mean [T, D] and variance [T, D] (full_cov: [T, T, D]) at Xnew [T, 2] (assigngp_dense.py:201-240)."""
        self._square_noise_ok()
        if self.KConst is not None:
            raise NotImplementedError("predict_f with KConst")
        Xnew = np.asarray(Xnew, dtype=np.float64)
        mean, var = self.engine.dense_predict(self.t, self.Y[None], np.zeros(1, dtype=np.int32), np.array([float(self.b.ravel()[0])]),
                                              self._prior, self._hyp_row(), self.logPhi.numpy()[None], Xnew,
                                              kernel=self._branch_kernel().kind, model=self.MODEL_KIND, full_cov=full_cov)
        D = self.Y.shape[1]
        if full_cov:
            return mean[0], np.repeat(var[0][:, :, None], D, axis=2)
        return mean[0], np.repeat(var[0][:, None], D, axis=1)
