"""MBGP AssignGP (MBGP/assigngp.py:67-877): D outputs, one branching point per output (``kernel.Bv``, sigmoid
transform), one shared variational assignment ``logPhi``.

``multi=True``  -> inducing-point bound summed over outputs (_build_likelihood_sparse, :492-558) on the CUDA sparse path;
``multi=False`` -> dense bound with the mask / kernel of ``Bv[0]`` and ``-D * KL`` (_build_likelihood_single_b, :608-647)."""
from enum import IntEnum
from typing import Optional, Sequence

import numpy as np

from .. import _lib
from .. import gpflow_shim as gpflow
from ..gpflow_shim import Kernel, Parameter, set_trainable
from . import BranchingTree as bt
from . import VBHelperFunctions

_DEFAULT_BRANCH_POINT = 0.0001   # earlier than anything in the data


class Branches(IntEnum):
    TRUNK = 0
    BRANCH1 = 1
    BRANCH2 = 2


def expand_pZ0Zeros(pZ0, epsilon=1e-6):
    """Dense N x 3N view of the three-column prior (trunk column explicit), MBGP/assigngp.py:25-40."""
    pZ0 = np.asarray(pZ0, dtype=np.float64)
    assert pZ0.shape[1] == 3, "Should have exactly three cols got %g " % pZ0.shape[1]
    n = pZ0.shape[0]
    for z0 in pZ0:
        assert z0.sum() == 1, "should sum to 1 is %s=%.3f" % (str(z0), z0.sum())
    r = np.full((n, 3 * n), float(epsilon))
    rows = np.arange(n)
    for f in range(3):
        r[rows, 3 * rows + f] = pZ0[:, f]
    return r


class BranchKernelParam(Kernel):
    """Branching-point kernel of mBGP: all functions share ``base_kern``; ``Bv`` holds one branching point per output."""
    name = "branch_kernel_param"

    def __init__(self, base_kern, branchPtTensor, branchParam, fDebug=False, noise_level=1e-6):
        fm = np.asarray(branchPtTensor)
        assert fm.shape[0] == fm.shape[1]
        assert fm.shape[2] > 0
        self.kern = base_kern
        self.fm = fm
        self.fDebug = fDebug
        self.Bv = branchParam
        self.noise_level = noise_level
        set_trainable(self.Bv, True)

    @property
    def kind(self):
        return self.kern.KIND

    def K(self, X, Y=None, dim=0):
        """Host-side numpy view of the covariance for output ``dim`` (inspection only; the bound never calls it)."""
        X = np.asarray(X, dtype=np.float64)
        square = Y is None
        Y = X if Y is None else np.asarray(Y, dtype=np.float64)
        b = float(np.asarray(self.Bv.numpy()).ravel()[dim])
        same = X[:, 1][:, None] == Y[:, 1][None, :]
        Ktt = self.kern.K_r(X[:, 0][:, None] - Y[:, 0][None, :])
        kbb = float(self.kern.variance.numpy()) + gpflow.default_jitter()
        cross = np.outer(self.kern.K_r(X[:, 0] - b) / kbb, self.kern.K_r(Y[:, 0] - b))
        K = np.where(same, Ktt, cross)
        return K + self.noise_level * np.eye(len(X)) if square else K

    def K_diag(self, X, dim=0):
        return self.kern.K_diag(X) + self.noise_level


def get_branch_point_tensor(initial_branching_point: float):
    tree = bt.BinaryBranchingTree(0.0, 1.0, fDebug=False)
    tree.add(None, 1, np.ones((1, 1)) * initial_branching_point)
    (fm, _) = tree.GetFunctionBranchTensor()
    return fm


def get_branching_point_kernel(branching_points: Sequence[float], base_kernel=None, transform=None, prior=None) -> BranchKernelParam:
    """Branching kernel around the given initial branching points; SquaredExponential base kernel and Sigmoid transform
    (pseudotime normalised to [0, 1]) by default (MBGP/assigngp.py:846-877)."""
    base_kernel = base_kernel or gpflow.kernels.SquaredExponential()
    transform = transform or gpflow.Sigmoid()
    locs = np.array(branching_points).flatten().astype(gpflow.default_float())
    bp_param = Parameter(locs, transform=transform, prior=prior, name="Bv")
    return BranchKernelParam(base_kern=base_kernel, branchPtTensor=get_branch_point_tensor(locs.mean()), branchParam=bp_param)


class AssignGP:
    def __init__(self, t, XExpanded, Y, indices, phiPrior=None, phiInitial=None, logPhi=None, fDebug=False, multi=False,
                 sparse=True, kern: Optional[BranchKernelParam] = None):
        assert len(t.shape) == 1, "pseudotime should be 1D"
        from ..assigngp_dense import _check_cell_major
        if sparse:
            M = 60
            Z = np.linspace(0.0, 1.0, M, endpoint=False)
            ZExpanded, _, _ = VBHelperFunctions.GetFunctionIndexListGeneral(Z)
            self.ZExpanded = Parameter(ZExpanded, trainable=False, name="ZExpanded")
            assert ZExpanded.shape[1] == XExpanded.shape[1]
        num_outputs = Y.shape[1]
        default_bps = _DEFAULT_BRANCH_POINT * np.ones(shape=(num_outputs,))
        kern = kern or get_branching_point_kernel(base_kernel=gpflow.kernels.SquaredExponential(), branching_points=default_bps)
        self.kernel = kern
        self.likelihood = gpflow.Gaussian()
        self.num_latent_gps = Y.shape[-1]
        self.Y = np.ascontiguousarray(Y, dtype=np.float64)
        self.t = t.astype(gpflow.default_float())
        self.X = _check_cell_major(self.t, XExpanded)
        self.multi = multi
        self.N = t.shape[0]
        self.D = Y.shape[1]
        self.indices = indices
        self.logPhi = Parameter(np.random.randn(t.shape[0], t.shape[0] * 3).astype(gpflow.default_float()), name="logPhi")
        if phiInitial is None:
            phiInitial = np.ones((self.N, 2)) * 0.5
            phiInitial[:, 0] = np.random.rand(self.N)
            phiInitial[:, 1] = 1 - phiInitial[:, 0]
        self.fDebug = fDebug
        if phiPrior is None:
            phiPrior = np.ones((self.N, 2)) * 0.5
            phiPrior = np.c_[np.zeros(phiPrior.shape[0])[:, None], phiPrior]
        self._engine = None
        if logPhi is not None:
            self.update_phi(prior=phiPrior)
            self.logPhi = logPhi if isinstance(logPhi, Parameter) else Parameter(logPhi, name="logPhi")
        else:
            self.update_phi(phiInitial, prior=phiPrior)

    # ------------------------------------------------------------------ plumbing
    @property
    def engine(self):
        if self._engine is None:
            self._engine = _lib.default_engine()
        return self._engine

    @property
    def phiTrunk(self):
        """Dense N x 3N constant used for cells before the branching point (built on demand)."""
        N = self.N
        out = np.zeros((N, 3 * N))
        rows = np.arange(N)
        eps = 1e-9
        out[rows, 3 * rows] = 1 - 2 * eps
        out[rows, 3 * rows + 1] = eps
        out[rows, 3 * rows + 2] = eps
        return out

    @property
    def eZ0(self):
        return expand_pZ0Zeros(self._prior)

    @property
    def BranchingPoints(self) -> np.ndarray:
        return self.kernel.Bv.numpy()

    # ------------------------------------------------------------------ reference API
    def UpdateBranchingPoint(self, b, phiInitial=None, prior=None):
        assert isinstance(b, np.ndarray)
        self.kernel.Bv.assign(b.flatten())
        self.update_phi(phi_initial=phiInitial, prior=prior)

    def update_phi(self, phi_initial=None, prior=None) -> None:
        if prior is not None:
            prior = np.asarray(prior, dtype=np.float64)
            assert prior.shape[1] == 3, "Should have exactly three cols got %g " % prior.shape[1]
            for z0 in prior:
                assert z0.sum() == 1, "should sum to 1 is %s=%.3f" % (str(z0), z0.sum())
            self._prior = np.ascontiguousarray(prior)
        if phi_initial is not None:
            self.InitialiseVariationalPhi(phi_initial)

    def InitialiseVariationalPhi(self, phiInitialIn):
        """Always the branch form log([1e-9, phi - 1e-9]) in the cell's block, -9 elsewhere (MBGP/assigngp.py:369-405)."""
        phiInitialIn = np.asarray(phiInitialIn, dtype=np.float64)
        assert np.allclose(phiInitialIn.sum(1), 1), "probs must sum to 1 %s" % str(phiInitialIn)
        N = self.Y.shape[0]
        assert phiInitialIn.shape[0] == N
        assert phiInitialIn.shape[1] == 2
        eps = 1e-9
        lp = np.full((N, 3 * N), -9.0)
        rows = np.arange(N)
        with np.errstate(divide="ignore", invalid="ignore"):
            block = np.log(np.hstack([np.full((N, 1), eps), phiInitialIn - eps]))
        assert not np.any(np.isnan(block)), "no nans please"
        for f in range(3):
            lp[rows, 3 * rows + f] = block[:, f]
        self.logPhi = Parameter(lp, name="logPhi")

    def GetPhi(self):
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
    @property
    def parameters(self):
        k = self.kernel
        return [k.Bv, k.kern.lengthscales, k.kern.variance, self.likelihood.variance, self.logPhi]

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
    def _hyp_row(self):
        k = self.kernel.kern
        return np.array([[float(k.lengthscales.numpy()), float(k.variance.numpy()), float(self.likelihood.variance.numpy())]])

    def _evaluate(self, want_grad):
        if not np.isclose(self.kernel.noise_level, 1e-6, rtol=1e-9, atol=0):
            raise NotImplementedError("noise_level other than 1e-6")
        Bv = np.asarray(self.kernel.Bv.numpy(), dtype=np.float64).ravel()
        gene = np.zeros(1, dtype=np.int32)
        if self.multi:
            Z = self.ZExpanded.numpy()
            out = self.engine.sparse_elbo_grad(self.t, Z, self.Y[None], gene, Bv[None], self._prior, self._hyp_row(),
                                               self.logPhi.numpy()[None], kernel=self.kernel.kind, model=_lib.MODEL_MBGP, multi=True,
                                               want_grad=want_grad)
            gb = out["grad_b"][0] if want_grad else None
        else:
            out = self.engine.dense_elbo_grad(self.t, self.Y[None], gene, Bv[:1], self._prior, self._hyp_row(), self.logPhi.numpy()[None],
                                              kernel=self.kernel.kind, model=_lib.MODEL_MBGP, kl_weight=float(self.D), want_grad=want_grad)
            gb = None
            if want_grad:
                gb = np.zeros(self.D)
                gb[0] = out["grad_hyp"][0, 3]
        out["grad_Bv"] = gb
        return out

    def maximum_log_likelihood_objective(self):
        return float(self._evaluate(False)["elbo"][0])

    def log_posterior_density(self):
        return self.maximum_log_likelihood_objective() + self.log_prior_density()

    def training_loss(self):
        return -self.log_posterior_density()

    def training_losses_for_branching_points(self, candidates):
        """training_loss for every row of `candidates` [n, D] used as the branching points, everything else unchanged:
        one batched CUDA call (n independent work items) instead of n sequential evaluations."""
        cand = np.ascontiguousarray(candidates, dtype=np.float64).reshape(-1, self.D)
        n = cand.shape[0]
        gene = np.zeros(n, dtype=np.int32)
        hyp = np.repeat(self._hyp_row(), n, axis=0)
        lp = np.broadcast_to(self.logPhi.numpy(), (n, self.N, 3 * self.N))
        if self.multi:
            out = self.engine.sparse_elbo_grad(self.t, self.ZExpanded.numpy(), self.Y[None], gene, cand, self._prior, hyp, lp,
                                               kernel=self.kernel.kind, model=_lib.MODEL_MBGP, multi=True, want_grad=False)
        else:
            out = self.engine.dense_elbo_grad(self.t, self.Y[None], gene, cand[:, 0], self._prior, hyp, lp, kernel=self.kernel.kind,
                                              model=_lib.MODEL_MBGP, kl_weight=float(self.D), want_grad=False)
        prior = 0.0
        for p in self.trainable_parameters:
            if p.prior is not None and p is not self.kernel.Bv:
                prior += float(np.sum(p.prior.log_prob(p.numpy())))
        bv = self.kernel.Bv
        bprior = np.zeros(n)
        if bv.trainable and bv.prior is not None:
            bprior = np.sum(bv.prior.log_prob(cand), axis=1)
        return -(out["elbo"] + prior + bprior)

    def _loss_and_grad(self, variables):
        out = self._evaluate(True)
        k = self.kernel
        gc = {id(k.kern.lengthscales): out["grad_hyp"][0, 0], id(k.kern.variance): out["grad_hyp"][0, 1],
              id(self.likelihood.variance): out["grad_hyp"][0, 2], id(self.logPhi): out["grad_logPhi"][0], id(k.Bv): out["grad_Bv"]}
        loss = -(float(out["elbo"][0]) + self.log_prior_density())
        grads = []
        for p in variables:
            if id(p) not in gc:
                raise ValueError(f"no gradient available for {p!r}")
            g = np.asarray(gc[id(p)], dtype=np.float64).reshape(p.shape)
            if p.prior is not None:
                g = g + p.prior.dlog_prob(p.numpy())
            grads.append(-(g * p.transform.dforward(p.unconstrained_variable)))
        return loss, grads

    # ------------------------------------------------------------------ prediction
    def predict_f(self, Xnew, full_cov=False, full_output_cov: bool = False):
        """multi=True: per-output dense posterior with that output's branching point (_build_predict_multi_b, :560-606);
        multi=False: dense posterior with Bv[0] (_build_predict_single_b, :608-647)."""
        assert not full_output_cov, "Not supported"
        Xnew = np.asarray(Xnew, dtype=np.float64)
        Bv = np.asarray(self.kernel.Bv.numpy(), dtype=np.float64).ravel()
        if self.multi:
            D = self.D
            Yt = np.ascontiguousarray(self.Y.T[:, :, None])           # [D, N, 1]: every output is its own single-output model
            lp = np.broadcast_to(self.logPhi.numpy(), (D, self.N, 3 * self.N))
            mean, var = self.engine.dense_predict(self.t, Yt, np.arange(D, dtype=np.int32), Bv, self._prior, np.repeat(self._hyp_row(), D, 0),
                                                  lp, Xnew, kernel=self.kernel.kind, model=_lib.MODEL_MBGP, full_cov=full_cov)
            mean = mean[:, :, 0].T                                      # [T, D]
            var = np.moveaxis(var, 0, -1)                               # [T, D] or [T, T, D]
            return mean, var
        mean, var = self.engine.dense_predict(self.t, self.Y[None], np.zeros(1, dtype=np.int32), Bv[:1], self._prior, self._hyp_row(),
                                              self.logPhi.numpy()[None], Xnew, kernel=self.kernel.kind, model=_lib.MODEL_MBGP,
                                              full_cov=full_cov)
        if full_cov:
            return mean[0], np.repeat(var[0][:, :, None], self.D, axis=2)
        return mean[0], np.repeat(var[0][:, None], self.D, axis=1)

    # ------------------------------------------------------------------ sampling (MBGP/assigngp.py:741-835)
    def predict_f_samples(self, Xnew, num_samples: Optional[int] = 1, full_cov: bool = True, full_output_cov: bool = False):
        """Samples of the posterior latent functions at Xnew, [S, N, D] (MBGP/assigngp.py:741-798).  Posterior moments from the
        CUDA predict kernels; with ``full_cov`` the D Cholesky factorisations of cov_d + jitter I and the products with the
        normal draws run batched on the GPU (bgp_gaussian_samples_host, one CTA per output).  The draws themselves come from
        numpy's global generator in the reference's shapes ([S, N, D] / [D, N, S])."""
        if full_cov and full_output_cov:
            raise NotImplementedError("The combination of both `full_cov` and `full_output_cov` is not supported.")
        S = 1 if num_samples is None else num_samples
        Xnew = np.asarray(Xnew, dtype=np.float64)
        N, D = Xnew.shape[0], self.Y.shape[-1]
        mean, cov = self.predict_f(Xnew, full_cov=full_cov, full_output_cov=full_output_cov)
        if not full_cov:
            samples = mean + np.sqrt(cov) * np.random.standard_normal((S, N, D))
        else:
            z = np.random.standard_normal((D, N, S))
            samples = self.engine.gaussian_samples(z, mean=np.ascontiguousarray(mean.T), cov=np.ascontiguousarray(np.transpose(cov, (2, 0, 1))),
                                                   jitter=gpflow.default_jitter())
        return samples[0] if num_samples is None else samples

    def sample_prior(self, x_expanded: np.ndarray, num_samples: int = 1) -> np.ndarray:
        """Samples of (f, g, h) from the prior at x_expanded, [S, N, D], each output with its own branching point
        (MBGP/assigngp.py:800-835): covariance generation, Cholesky and the product with the draws run on the GPU, one CTA per
        output (bgp_gaussian_samples_host with from_kernel = 1)."""
        x_expanded = np.asarray(x_expanded, dtype=np.float64)
        N, D, S = x_expanded.shape[0], self.Y.shape[1], num_samples
        z = np.random.standard_normal((D, N, S))
        k = self.kernel
        return self.engine.gaussian_samples(z, X=x_expanded, bv=np.asarray(k.Bv.numpy(), dtype=np.float64).ravel(),
                                            ell=float(k.kern.lengthscales.numpy()), kvar=float(k.kern.variance.numpy()), kernel=k.kind,
                                            noise=float(k.noise_level), jitter=gpflow.default_jitter())

    def predict_y(self, Xnew, full_cov=False):
        mean, var = self.predict_f(Xnew, full_cov=full_cov)
        s = float(self.likelihood.variance.numpy())
        if full_cov:
            T = var.shape[0]
            return mean, var + s * np.eye(T)[:, :, None]
        return mean, var + s
