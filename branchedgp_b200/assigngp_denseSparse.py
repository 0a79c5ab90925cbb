"""AssignGPSparse: AssignGP with fixed inducing points ZExpanded (assigngp_denseSparse.py:9-155)."""
import numpy as np

from . import assigngp_dense


class AssignGPSparse(assigngp_dense.AssignGP):
    def __init__(self, t, XExpanded, Y, kern, indices, b, ZExpanded, fDebug=False, phiInitial=None, phiPrior=None):
        assigngp_dense.AssignGP.__init__(self, t, XExpanded, Y, kern, indices, b, fDebug=fDebug, phiInitial=phiInitial,
                                         phiPrior=phiPrior)
        # inducing points are data, never parameters
        self.ZExpanded = np.ascontiguousarray(ZExpanded, dtype=np.float64)
        assert self.ZExpanded.shape[1] == self.X.shape[1]

    def _evaluate(self, want_grad):
        self._square_noise_ok()
        k = self._branch_kernel()
        out = self.engine.sparse_elbo_grad(self.t, self.ZExpanded, self.Y[None], np.zeros(1, dtype=np.int32),
                                           np.array([float(self.b.ravel()[0])]), self._prior, self._hyp_row(), self.logPhi.numpy()[None],
                                           kernel=k.kind, model=self.MODEL_KIND, multi=False, want_grad=want_grad)
        if want_grad:   # same layout as the dense engine: [count, 4] = ell, var, likvar, b
            out["grad_hyp"] = np.hstack([out["grad_hyp"], out["grad_b"][:, :1]])
        self.bound = float(out["elbo"][0])
        return out

    def predict_f(self, Xnew, full_cov=False):
        """assigngp_denseSparse.py:109-155."""
        self._square_noise_ok()
        Xnew = np.asarray(Xnew, dtype=np.float64)
        mean, var = self.engine.sparse_predict(self.t, self.ZExpanded, self.Y, float(self.b.ravel()[0]), self._prior, self._hyp_row()[0],
                                               self.logPhi.numpy(), Xnew, kernel=self._branch_kernel().kind, full_cov=full_cov)
        D = self.Y.shape[1]
        if full_cov:
            return mean, np.repeat(var[:, :, None], D, axis=2)
        return mean, np.repeat(var[:, None], D, axis=1)
