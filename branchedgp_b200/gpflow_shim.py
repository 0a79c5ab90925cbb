"""The handful of GPflow / TFP names the reference's callers touch, without TensorFlow.

The reference builds its models from ``gpflow.Parameter``, ``gpflow.kernels.*``, ``gpflow.utilities.set_trainable``,
``tfp.distributions.Normal`` priors and ``gpflow.optimizers.Scipy`` (FitBranchingModel.py:64-132).  These light
containers keep that vocabulary so reference call sites and tests read unchanged; they hold numbers only -- all bound /
gradient arithmetic happens in libbgp_b200.so.  Semantics restated from GPflow 2.1.2:

* positive parameters: softplus transform, ``y = log(1 + exp(u)) + lower`` (Gaussian likelihood variance: lower = 1e-6);
* ``Sigmoid`` transform for MBGP's branching points (MBGP/assigngp.py:860-866);
* ``log_prior_density`` = sum of ``prior.log_prob(constrained value)`` over trainable parameters that carry a prior;
* ``optimizers.Scipy().minimize(closure, variables, options=...)`` runs SciPy L-BFGS-B on the packed unconstrained vector.
"""
import math

import numpy as np
import scipy.optimize


def default_float():
    return np.float64


def default_jitter():
    return 1e-6


def to_default_float(x):
    return np.asarray(x, dtype=np.float64)


# ---------------------------------------------------------------------------------------------------- transforms
class Identity:
    def forward(self, u):
        return u

    def inverse(self, y):
        return y

    def dforward(self, u):
        return np.ones_like(u)


class Softplus:
    def __init__(self, lower=0.0):
        self.lower = float(lower)

    def forward(self, u):
        return np.logaddexp(0.0, u) + self.lower

    def inverse(self, y):
        z = np.asarray(y, dtype=np.float64) - self.lower
        return z + np.log(-np.expm1(-z))

    def dforward(self, u):
        return 1.0 / (1.0 + np.exp(-u))


class Sigmoid:
    """tfp.bijectors.Sigmoid(low, high)."""

    def __init__(self, low=0.0, high=1.0):
        self.low, self.high = float(low), float(high)

    def forward(self, u):
        return self.low + (self.high - self.low) / (1.0 + np.exp(-u))

    def inverse(self, y):
        z = (np.asarray(y, dtype=np.float64) - self.low) / (self.high - self.low)
        return np.log(z) - np.log1p(-z)

    def dforward(self, u):
        s = 1.0 / (1.0 + np.exp(-u))
        return (self.high - self.low) * s * (1.0 - s)


def positive(lower=0.0):
    return Softplus(lower)


# ---------------------------------------------------------------------------------------------------- priors
class Normal:
    """tfp.distributions.Normal(loc, scale): only log_prob and its derivative are needed."""

    def __init__(self, loc, scale):
        self.loc, self.scale = float(loc), float(scale)

    def log_prob(self, y):
        z = (np.asarray(y, dtype=np.float64) - self.loc) / self.scale
        return -0.5 * z * z - math.log(self.scale) - 0.5 * math.log(2.0 * math.pi)

    def dlog_prob(self, y):
        return -(np.asarray(y, dtype=np.float64) - self.loc) / (self.scale * self.scale)


class distributions:   # so that ``tfp.distributions.Normal`` reads the same
    Normal = Normal


# ---------------------------------------------------------------------------------------------------- Parameter
class Parameter:
    def __init__(self, value, transform=None, prior=None, trainable=True, name=None):
        self.transform = transform if transform is not None else Identity()
        self.prior = prior
        self.trainable = bool(trainable)
        self.name = name
        self._u = None
        self._pinned = False
        self.assign(value)

    # constrained view
    def numpy(self):
        return np.array(self.transform.forward(self._u), dtype=np.float64)

    def view(self):
        """Constrained value WITHOUT a copy when there is no transform (the N x 3N assignment logits: 24 MB at N = 1000 that
        would otherwise be copied on every function evaluation); read-only use."""
        if isinstance(self.transform, Identity):
            return self._u
        return self.numpy()

    def assign(self, value):
        v = np.array(value, dtype=np.float64)
        u = np.asarray(self.transform.inverse(v), dtype=np.float64)
        if self._pinned and self._u is not None and u.shape == self._u.shape:
            np.copyto(self._u, u)          # keep the page-locked storage
        else:
            self._u = np.array(u, dtype=np.float64)
            self._pinned = False
        return self

    def pin(self, alloc):
        """Move the storage into page-locked host memory (``alloc(shape)`` -> float64 array, e.g. ``_lib.pinned_empty``) so
        that the host-pointer CUDA entry points read it by DMA.  No-op when already pinned."""
        if not self._pinned:
            buf = alloc(self._u.shape)
            np.copyto(buf, self._u)
            self._u = buf
            self._pinned = True
        return self

    @property
    def unconstrained_variable(self):
        return self._u

    def set_unconstrained(self, u):
        u = np.asarray(u, dtype=np.float64).reshape(self._u.shape)
        if self._pinned:
            np.copyto(self._u, u)
        else:
            self._u = np.array(u, dtype=np.float64)

    @property
    def shape(self):
        return self._u.shape

    @property
    def size(self):
        return int(self._u.size)

    def __array__(self, dtype=None, copy=None):
        a = self.numpy()
        return a.astype(dtype) if dtype is not None else a

    def __float__(self):
        return float(self.numpy())

    def __repr__(self):
        return f"Parameter({self.name or ''} value={self.numpy()!r}, trainable={self.trainable})"


def set_trainable(obj, flag):
    if isinstance(obj, Parameter):
        obj.trainable = bool(flag)
        return
    for p in getattr(obj, "parameters", lambda: [])():
        p.trainable = bool(flag)


class utilities:
    set_trainable = staticmethod(set_trainable)
    to_default_float = staticmethod(to_default_float)


# ---------------------------------------------------------------------------------------------------- kernels (containers)
class Kernel:
    name = "kernel"

    def parameters(self):
        out = []
        for v in self.__dict__.values():
            if isinstance(v, Parameter):
                out.append(v)
            elif isinstance(v, Kernel):
                out.extend(v.parameters())
            elif isinstance(v, (list, tuple)):
                for k in v:
                    if isinstance(k, Kernel):
                        out.extend(k.parameters())
        return out

    def __add__(self, other):
        return Sum([self, other])


class Sum(Kernel):
    name = "sum"

    def __init__(self, kernels):
        self.kernels = list(kernels)


class _Stationary(Kernel):
    KIND = None

    def __init__(self, variance=1.0, lengthscales=1.0, **_ignored):
        # note: the reference writes gpflow.kernels.Matern32(1): the positional argument is the variance
        self.variance = Parameter(variance, transform=positive(), name="variance")
        self.lengthscales = Parameter(lengthscales, transform=positive(), name="lengthscales")

    def K_r(self, d):
        """k as a function of the signed distance d = t - t' (host helper; never on the bound's path)."""
        ell, var = float(self.lengthscales.numpy()), float(self.variance.numpy())
        r = np.abs(d) / ell
        if self.KIND == 1:
            return var * np.exp(-0.5 * r * r)
        s = math.sqrt(3.0) * r
        return var * (1.0 + s) * np.exp(-s)

    def K(self, X, X2=None):
        X = np.asarray(X, dtype=np.float64).reshape(len(X), -1)
        X2 = X if X2 is None else np.asarray(X2, dtype=np.float64).reshape(len(X2), -1)
        return self.K_r(X[:, 0][:, None] - X2[:, 0][None, :])

    def K_diag(self, X):
        return np.full(len(X), float(self.variance.numpy()))


class Matern32(_Stationary):
    name = "matern32"
    KIND = 0


class SquaredExponential(_Stationary):
    name = "squared_exponential"
    KIND = 1


RBF = SquaredExponential


class White(Kernel):
    name = "white"

    def __init__(self, variance=1.0, **_ignored):
        self.variance = Parameter(variance, transform=positive(), name="variance")

    def K(self, X, X2=None):
        n = len(X)
        if X2 is None:
            return float(self.variance.numpy()) * np.eye(n)
        return np.zeros((n, len(X2)))

    def K_diag(self, X):
        return np.full(len(X), float(self.variance.numpy()))


class kernels:
    Kernel = Kernel
    Sum = Sum
    Matern32 = Matern32
    SquaredExponential = SquaredExponential
    RBF = RBF
    White = White


# ---------------------------------------------------------------------------------------------------- likelihood
class Gaussian:
    DEFAULT_VARIANCE_LOWER_BOUND = 1e-6

    def __init__(self, variance=1.0):
        self.variance = Parameter(variance, transform=positive(self.DEFAULT_VARIANCE_LOWER_BOUND), name="likelihood_variance")

    def parameters(self):
        return [self.variance]


class likelihoods:
    Gaussian = Gaussian


# ---------------------------------------------------------------------------------------------------- optimiser
class Scipy:
    """gpflow.optimizers.Scipy: L-BFGS-B on the packed unconstrained variables.

    ``closure`` must be the bound method ``model.training_loss`` of one of this package's models (the model supplies
    value and gradient together from one CUDA evaluation); ``variables`` is ``model.trainable_variables``."""

    def minimize(self, closure, variables, method="L-BFGS-B", step_callback=None, compile=True, **scipy_kwargs):
        model = getattr(closure, "__self__", None)
        if model is None or not hasattr(model, "_loss_and_grad"):
            raise TypeError("Scipy.minimize needs closure = <model>.training_loss of a branchedgp_b200 model")
        variables = list(variables)
        if not variables:
            raise ValueError("no trainable variables")
        sizes = [p.size for p in variables]
        x0 = np.concatenate([p.unconstrained_variable.ravel() for p in variables])

        def unpack(x):
            o = 0
            for p, n in zip(variables, sizes):
                p.set_unconstrained(x[o:o + n])
                o += n

        # the packed gradient is written in place, variable by variable (for the N x 3N logits that is one pass over 24 MB at
        # N = 1000 instead of a negation, a product with the transform's derivative and a concatenation); two buffers are used
        # in turn because SciPy may still hold the array returned by the previous evaluation
        gpack = [np.empty(sum(sizes)), np.empty(sum(sizes))]
        turn = [0]
        import inspect
        in_place = "out" in inspect.signature(model._loss_and_grad).parameters

        def fun(x):
            unpack(x)
            buf = gpack[turn[0]]
            turn[0] ^= 1
            outs, o = [], 0
            for p, n in zip(variables, sizes):
                outs.append(buf[o:o + n].reshape(p.shape))
                o += n
            if in_place:
                loss, _ = model._loss_and_grad(variables, out=outs)
            else:                      # a model without the in-place variant
                loss, grads = model._loss_and_grad(variables)
                for dst, g in zip(outs, grads):
                    np.copyto(dst, np.asarray(g).reshape(dst.shape))
            return loss, buf

        cb = None
        if step_callback is not None:
            it = [0]

            def cb(x):
                unpack(x)
                step_callback(it[0], variables, [p.unconstrained_variable for p in variables])
                it[0] += 1

        options = dict(scipy_kwargs.pop("options", {}) or {})
        options.pop("disp", None)   # the Fortran L-BFGS-B progress print is noise here
        res = scipy.optimize.minimize(fun, x0, jac=True, method=method, callback=cb, options=options, **scipy_kwargs)
        unpack(res.x)
        return res


class optimizers:
    Scipy = Scipy
