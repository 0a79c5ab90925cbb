"""ctypes binding of libbgp_b200.so (C ABI declared in include/branchedgp_b200.h).

There is deliberately NO fallback: if the CUDA library is missing or no sm_100 GPU is visible, creating an
``Engine`` raises.  numpy is the only Python dependency of this module; torch is never imported here.
"""
import ctypes
import os
import threading

import numpy as np

KERNEL_MATERN32 = 0
KERNEL_SQEXP = 1
MODEL_BGP = 0
MODEL_MBGP = 1

_LIB_NAME = os.environ.get("BGP_LIB_NAME", "libbgp_b200.so")   # test hook: alternative builds of the same library
_lib = None
_lib_lock = threading.Lock()

c_double_p = ctypes.POINTER(ctypes.c_double)
c_int_p = ctypes.POINTER(ctypes.c_int)

# name -> (restype, argtypes); must list every symbol include/branchedgp_b200.h declares (tests check this).
SIGNATURES = {
    "bgp_version": (ctypes.c_int, []),
    "bgp_create": (ctypes.c_int, [ctypes.c_int, ctypes.POINTER(ctypes.c_void_p)]),
    "bgp_destroy": (ctypes.c_int, [ctypes.c_void_p]),
    "bgp_last_error": (ctypes.c_char_p, [ctypes.c_void_p]),
    "bgp_device_sm_count": (ctypes.c_int, [ctypes.c_void_p]),
    "bgp_set_max_slots": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int]),
    "bgp_dense_slot_bytes": (ctypes.c_size_t, [ctypes.c_int, ctypes.c_int]),
    "bgp_host_alloc": (ctypes.c_int, [ctypes.POINTER(ctypes.c_void_p), ctypes.c_size_t]),
    "bgp_host_free": (ctypes.c_int, [ctypes.c_void_p]),
    "bgp_dense_elbo_grad": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                           ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                           ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_void_p,
                                           ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                           ctypes.c_void_p, ctypes.c_void_p]),
    "bgp_dense_elbo_grad_host": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                                ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                                ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_double,
                                                ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
                                                ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "bgp_sparse_elbo_grad": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                            ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "bgp_sparse_elbo_grad_host": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                                 ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
                                                 ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                                                 ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                                 ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "bgp_dense_predict_host": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                              ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                              ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int,
                                              ctypes.c_void_p, ctypes.c_void_p]),
    "bgp_sparse_predict_host": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                               ctypes.c_void_p, ctypes.c_double, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
                                               ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
                                               ctypes.c_void_p]),
    "bgp_comm_unique_id": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    "bgp_comm_init": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]),
    "bgp_comm_destroy": (ctypes.c_int, [ctypes.c_void_p]),
    "bgp_gather_results": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]),
    "bgp_gaussian_samples_host": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                                 ctypes.c_void_p, ctypes.c_double, ctypes.c_double, ctypes.c_int, ctypes.c_double,
                                                 ctypes.c_void_p, ctypes.c_void_p, ctypes.c_double, ctypes.c_void_p, ctypes.c_void_p,
                                                 ctypes.c_void_p]),
    "bgp_get_phi_host": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]),
    "bgp_fit_default_options": (None, [ctypes.c_void_p]),
    "bgp_dense_fit_host": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int] + [ctypes.c_void_p] * 2 + [ctypes.c_int]
                           + [ctypes.c_void_p] * 5 + [ctypes.c_int, ctypes.c_void_p, ctypes.c_int] + [ctypes.c_void_p] * 11),
    "bgp_sparse_fit_host": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int] + [ctypes.c_void_p] * 3
                            + [ctypes.c_int] + [ctypes.c_void_p] * 5 + [ctypes.c_int, ctypes.c_void_p, ctypes.c_int] + [ctypes.c_void_p] * 11),
    "bgp_debug_gemm_timers": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]),
    "bgp_debug_wide_timers": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]),
    "bgp_debug_linesearch": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(ctypes.c_double), ctypes.c_double, ctypes.c_double]),
    "bgp_debug_lbfgs_direction": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "bgp_profile_enable": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int]),
    "bgp_profile_reset": (ctypes.c_int, [ctypes.c_void_p]),
    "bgp_profile_read": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.POINTER(ctypes.c_longlong)]),
    "bgp_debug_set_stop_phase": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int]),
    "bgp_debug_padded_size": (ctypes.c_int, [ctypes.c_int]),
    "bgp_debug_dense_matrix": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]),
    "bgp_debug_dense_vector": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]),
}


class FitOptions(ctypes.Structure):
    """bgp_fit_options of include/branchedgp_b200.h (SciPy L-BFGS-B defaults, FitModel's hyper-priors)."""
    _fields_ = [("maxiter", ctypes.c_int), ("maxcor", ctypes.c_int), ("maxls", ctypes.c_int), ("maxfun", ctypes.c_int),
                ("ftol", ctypes.c_double), ("gtol", ctypes.c_double), ("train_hyp", ctypes.c_int), ("has_prior", ctypes.c_int * 3),
                ("prior_mean", ctypes.c_double * 3), ("prior_sd", ctypes.c_double * 3), ("max_slots", ctypes.c_int)]


FIT_STATUS = ("converged: max |g| <= gtol", "converged: relative reduction of f <= ftol", "stopped: maxiter", "stopped: maxfun",
              "abnormal termination in line search", "Cholesky decomposition was not successful")


class BgpError(RuntimeError):
    """The CUDA library reported an error (or is not available)."""


class CholeskyError(BgpError):
    """Non-positive pivot in a Cholesky factorisation (the reference's tf.linalg.cholesky InvalidArgumentError)."""


def library_path():
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), _LIB_NAME)


def load_library():
    """dlopen the in-tree shared library and declare the signatures.  Raises BgpError when it is missing."""
    global _lib
    with _lib_lock:
        if _lib is not None:
            return _lib
        # the host-pointer entry points keep up to 16 item groups in flight on their own streams; the default of 8 hardware
        # work queues would serialise some of them (only effective if the CUDA context does not exist yet)
        os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
        path = library_path()
        if not os.path.exists(path):
            raise BgpError(f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(make -C branchedgp_b200/csrc); there is no CPU fallback")
        lib = ctypes.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
        return lib


def pinned_empty(shape):
    """float64 numpy array in page-locked host memory (bgp_host_alloc): the `_host` entry points then move it at full PCIe speed
    instead of through the driver's staging copies.  Freed when the last view is garbage collected."""
    import weakref
    lib = load_library()
    shape = tuple(int(x) for x in np.atleast_1d(shape))
    n = int(np.prod(shape))
    p = ctypes.c_void_p()
    if lib.bgp_host_alloc(ctypes.byref(p), max(n, 1) * 8) != 0 or not p.value:
        raise BgpError(f"bgp_host_alloc of {n * 8} bytes failed")
    flat = np.ctypeslib.as_array(ctypes.cast(p, ctypes.POINTER(ctypes.c_double)), shape=(max(n, 1),))
    weakref.finalize(flat, lib.bgp_host_free, p)
    return flat[:n].reshape(shape)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _ptr(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


class Engine:
    """One handle on one GPU.  Not thread-safe; create one per host thread."""

    def __init__(self, device=0):
        self._lib = load_library()
        h = ctypes.c_void_p()
        rc = self._lib.bgp_create(int(device), ctypes.byref(h))
        self._h = h
        if rc != 0:
            msg = self._lib.bgp_last_error(h).decode() if h else "bgp_create failed"
            raise BgpError(f"bgp_create(device={device}) failed: {msg}")
        self.device = int(device)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.bgp_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ helpers
    def _check(self, rc, what):
        if rc != 0:
            raise BgpError(f"{what} failed (code {rc}): {self._lib.bgp_last_error(self._h).decode()}")

    @property
    def sm_count(self):
        return self._lib.bgp_device_sm_count(self._h)

    def set_max_slots(self, n):
        self._check(self._lib.bgp_set_max_slots(self._h, int(n)), "bgp_set_max_slots")

    # ------------------------------------------------------------------ dense bound
    def dense_elbo_grad(self, t, Y, item_gene, b, prior, hyp, logPhi, kernel=KERNEL_MATERN32, model=MODEL_BGP,
                        kl_weight=1.0, KConst=None, want_grad=True, check=True, grad_out=None):
        """Host-array entry: Y [nY,N,D], item_gene [count], b [count], hyp [count,3], logPhi [count,N,3N]."""
        t = _f64(t)
        Y = _f64(Y)
        if Y.ndim == 2:
            Y = Y[None]
        nY, N, D = Y.shape
        item_gene = np.ascontiguousarray(item_gene, dtype=np.int32)
        count = item_gene.size
        b = _f64(b).reshape(count)
        hyp = _f64(hyp).reshape(count, 3)
        prior = _f64(prior)
        logPhi = _f64(logPhi).reshape(count, N, 3 * N)
        assert t.shape == (N,)
        assert prior.shape == (N, 3 if model == MODEL_MBGP else 2)
        assert item_gene.min() >= 0 and item_gene.max() < nY
        if KConst is not None:
            KConst = _f64(KConst)
            assert KConst.shape == (3 * N, 3 * N)
        elbo = np.empty(count)
        status = np.empty(count, dtype=np.int32)
        grad_hyp = np.zeros((count, 4)) if want_grad else None
        grad_logPhi = None
        if want_grad:
            grad_logPhi = grad_out if grad_out is not None else np.empty((count, N, 3 * N))
            assert grad_logPhi.shape == (count, N, 3 * N) and grad_logPhi.dtype == np.float64 and grad_logPhi.flags.c_contiguous
        rc = self._lib.bgp_dense_elbo_grad_host(self._h, count, N, D, _ptr(t), _ptr(Y), nY, _ptr(item_gene), _ptr(b),
                                                _ptr(prior), _ptr(hyp), int(kernel), int(model), float(kl_weight),
                                                _ptr(logPhi), _ptr(KConst), int(bool(want_grad)), _ptr(elbo),
                                                _ptr(grad_hyp), _ptr(grad_logPhi), _ptr(status))
        self._check(rc, "bgp_dense_elbo_grad_host")
        if check and (status != 0).any():
            i = int(np.flatnonzero(status)[0])
            raise CholeskyError(f"Cholesky decomposition was not successful for work item {i} (status {int(status[i])}): "
                                "the input might not be valid")
        return dict(elbo=elbo, grad_hyp=grad_hyp, grad_logPhi=grad_logPhi, status=status)

    def dense_elbo_grad_device(self, count, N, D, t, Y, nY, item_gene, b, prior, hyp, logPhi, elbo, grad_hyp, grad_logPhi,
                               status, kernel=KERNEL_MATERN32, model=MODEL_BGP, kl_weight=1.0, KConst=0, want_grad=True,
                               stream=0):
        """Device-pointer entry (integers from e.g. torch.Tensor.data_ptr()); asynchronous on `stream`."""
        rc = self._lib.bgp_dense_elbo_grad(self._h, count, N, D, t, Y, nY, item_gene, b, prior, hyp, int(kernel), int(model),
                                           float(kl_weight), logPhi, KConst or None, int(bool(want_grad)), elbo,
                                           grad_hyp or None, grad_logPhi or None, status, stream or None)
        self._check(rc, "bgp_dense_elbo_grad")

    # ------------------------------------------------------------------ inducing-point bounds
    def sparse_elbo_grad(self, t, Z, Y, item_gene, b, prior, hyp, logPhi, kernel=KERNEL_MATERN32, model=MODEL_BGP, multi=False,
                         want_grad=True, check=True, grad_out=None):
        """Host-array entry.  Y [nY,N,Dtot]; b [count] (multi=False) or [count,Dtot] (multi=True); Z [Mz,2]."""
        t = _f64(t)
        Z = _f64(Z)
        Y = _f64(Y)
        if Y.ndim == 2:
            Y = Y[None]
        nY, N, Dtot = Y.shape
        Mz = Z.shape[0]
        item_gene = np.ascontiguousarray(item_gene, dtype=np.int32)
        count = item_gene.size
        SD = Dtot if multi else 1
        b = _f64(b).reshape(count, SD)
        hyp = _f64(hyp).reshape(count, 3)
        prior = _f64(prior)
        logPhi = _f64(logPhi).reshape(count, N, 3 * N)
        assert t.shape == (N,) and Z.shape == (Mz, 2)
        assert prior.shape == (N, 3 if model == MODEL_MBGP else 2)
        elbo = np.empty(count)
        status = np.empty(count, dtype=np.int32)
        grad_hyp = np.zeros((count, 3)) if want_grad else None
        grad_b = np.zeros((count, SD)) if want_grad else None
        grad_logPhi = None
        if want_grad:
            grad_logPhi = grad_out if grad_out is not None else np.empty((count, N, 3 * N))
        rc = self._lib.bgp_sparse_elbo_grad_host(self._h, count, N, Dtot, Mz, _ptr(t), _ptr(Z), _ptr(Y), nY, _ptr(item_gene), _ptr(b),
                                                 _ptr(prior), _ptr(hyp), int(kernel), int(model), int(bool(multi)), _ptr(logPhi),
                                                 int(bool(want_grad)), _ptr(elbo), _ptr(grad_hyp), _ptr(grad_b), _ptr(grad_logPhi),
                                                 _ptr(status))
        self._check(rc, "bgp_sparse_elbo_grad_host")
        if check and (status != 0).any():
            i = int(np.flatnonzero(status)[0])
            raise CholeskyError(f"Cholesky decomposition was not successful for work item {i} (status {int(status[i])}): "
                                "the input might not be valid")
        return dict(elbo=elbo, grad_hyp=grad_hyp, grad_b=grad_b, grad_logPhi=grad_logPhi, status=status)

    # ------------------------------------------------------------------ predict_f
    def dense_predict(self, t, Y, item_gene, b, prior, hyp, logPhi, Xnew, kernel=KERNEL_MATERN32, model=MODEL_BGP, full_cov=False):
        """Posterior mean [count,T,D] and variance [count,T] (or [count,T,T]) of the latent functions at Xnew [T,2]."""
        t = _f64(t)
        Y = _f64(Y)
        if Y.ndim == 2:
            Y = Y[None]
        nY, N, D = Y.shape
        item_gene = np.ascontiguousarray(item_gene, dtype=np.int32)
        count = item_gene.size
        b = _f64(b).reshape(count)
        hyp = _f64(hyp).reshape(count, 3)
        prior = _f64(prior)
        logPhi = _f64(logPhi).reshape(count, N, 3 * N)
        Xnew = _f64(Xnew)
        T = Xnew.shape[0]
        assert Xnew.shape == (T, 2)
        mean = np.empty((count, T, D))
        var = np.empty((count, T, T) if full_cov else (count, T))
        rc = self._lib.bgp_dense_predict_host(self._h, count, N, D, _ptr(t), _ptr(Y), nY, _ptr(item_gene), _ptr(b), _ptr(prior),
                                              _ptr(hyp), int(kernel), int(model), _ptr(logPhi), T, _ptr(Xnew), int(bool(full_cov)),
                                              _ptr(mean), _ptr(var))
        self._check(rc, "bgp_dense_predict_host")
        return mean, var

    def sparse_predict(self, t, Z, Y, b, prior, hyp, logPhi, Xnew, kernel=KERNEL_MATERN32, full_cov=False):
        t, Z, Y, prior, Xnew = _f64(t), _f64(Z), _f64(Y), _f64(prior), _f64(Xnew)
        N, D = Y.shape
        hyp = _f64(hyp).reshape(3)
        logPhi = _f64(logPhi).reshape(N, 3 * N)
        T = Xnew.shape[0]
        mean = np.empty((T, D))
        var = np.empty((T, T) if full_cov else (T,))
        rc = self._lib.bgp_sparse_predict_host(self._h, N, D, Z.shape[0], _ptr(t), _ptr(Z), _ptr(Y), float(b), _ptr(prior), _ptr(hyp),
                                               int(kernel), _ptr(logPhi), T, _ptr(Xnew), int(bool(full_cov)), _ptr(mean), _ptr(var))
        self._check(rc, "bgp_sparse_predict_host")
        return mean, var

    # ------------------------------------------------------------------ batched fitting (device-resident L-BFGS)
    # ------------------------------------------------------------------ NCCL gather inside the C ABI
    def comm_unique_id(self):
        buf = ctypes.create_string_buffer(128)
        self._check(self._lib.bgp_comm_unique_id(self._h, buf), "bgp_comm_unique_id")
        return buf.raw

    def comm_init(self, nranks, rank, unique_id):
        buf = ctypes.create_string_buffer(bytes(unique_id), 128)
        self._check(self._lib.bgp_comm_init(self._h, int(nranks), int(rank), buf), "bgp_comm_init")
        self._comm_ranks = int(nranks)

    def gather_results(self, local, rows_max):
        """all_gather of this rank's [rows, k] float64 block -> [nranks, rows_max, k] (zero padded), over the library's own NCCL
        communicator (bgp_gather_results)."""
        local = _f64(local)
        rows, k = local.shape
        out = np.empty((self._comm_ranks, int(rows_max), k))
        self._check(self._lib.bgp_gather_results(self._h, _ptr(local) if rows else None, rows, k, int(rows_max), _ptr(out)),
                    "bgp_gather_results")
        return out

    def gaussian_samples(self, z, mean=None, cov=None, X=None, bv=None, ell=1.0, kvar=1.0, kernel=KERNEL_SQEXP, noise=1e-6,
                         jitter=1e-6):
        """samples[S, n, batch] = mean + chol(cov_d + jitter I) z[d]  for every output d (bgp_gaussian_samples_host).
        z [batch, n, S] standard normal draws; either cov [batch, n, n] or (X [n, 2], bv [batch]) for the prior covariance of
        the branching kernel with one branching point per output."""
        z = _f64(z)
        batch, n, S = z.shape
        from_kernel = cov is None
        if from_kernel:
            X = _f64(X)
            bv = _f64(bv).reshape(batch)
            assert X.shape == (n, 2)
        else:
            cov = _f64(cov)
            assert cov.shape == (batch, n, n)
        if mean is not None:
            mean = _f64(mean)
            assert mean.shape == (batch, n)
        out = np.empty((S, n, batch))
        status = np.empty(batch, dtype=np.int32)
        rc = self._lib.bgp_gaussian_samples_host(self._h, batch, n, S, int(from_kernel), _ptr(X) if from_kernel else None,
                                                 _ptr(bv) if from_kernel else None, float(ell), float(kvar), int(kernel), float(noise),
                                                 None if from_kernel else _ptr(cov), _ptr(mean), float(jitter), _ptr(z), _ptr(out),
                                                 _ptr(status))
        self._check(rc, "bgp_gaussian_samples_host")
        if (status != 0).any():
            d = int(np.flatnonzero(status)[0])
            raise CholeskyError(f"Cholesky decomposition was not successful for output {d} (status {int(status[d])})")
        return out

    def fit_options(self, **kw):
        """bgp_fit_default_options with keyword overrides (maxiter, maxcor, maxls, maxfun, ftol, gtol, train_hyp, has_prior,
        prior_mean, prior_sd, max_slots)."""
        o = FitOptions()
        self._lib.bgp_fit_default_options(ctypes.byref(o))
        for k, v in kw.items():
            if k in ("has_prior", "prior_mean", "prior_sd"):
                arr = getattr(o, k)
                for i in range(3):
                    arr[i] = v[i]
            else:
                if not hasattr(o, k):
                    raise TypeError(f"unknown fit option {k!r}")
                setattr(o, k, v)
        return o

    def fit(self, t, Y, item_gene, b, prior, phiInitial, hyp0, Z=None, kernel=KERNEL_MATERN32, options=None, Xnew=None,
            return_logPhi=False):
        """Fit `count` work items (gene item_gene[i], branching point b[i], initial hyper-parameters hyp0[i]) at once.

        Dense AssignGP when Z is None, AssignGPSparse with inducing inputs Z [Mz,2] otherwise.  Xnew [count,T,2]: test inputs
        of predict_f per item.  Returns dict(objective, elbo, hyp, phi, pred_mean, pred_var, iters, nfev, status[, logPhi])."""
        t = _f64(t)
        Y = _f64(Y)
        if Y.ndim == 2:
            Y = Y[None]
        nY, N, D = Y.shape
        item_gene = np.ascontiguousarray(item_gene, dtype=np.int32)
        count = item_gene.size
        b = _f64(b).reshape(count)
        hyp0 = _f64(hyp0).reshape(count, 3)
        prior = _f64(prior)
        phiInitial = _f64(phiInitial)
        assert t.shape == (N,) and prior.shape == (N, 2) and phiInitial.shape == (N, 2)
        assert count > 0 and item_gene.min() >= 0 and item_gene.max() < nY
        if options is None:
            options = self.fit_options()
        T = 0
        if Xnew is not None:
            Xnew = _f64(Xnew)
            assert Xnew.ndim == 3 and Xnew.shape[0] == count and Xnew.shape[2] == 2
            T = Xnew.shape[1]
        out = dict(objective=np.empty(count), elbo=np.empty(count), hyp=np.empty((count, 3)), phi=np.empty((count, N, 3)),
                   pred_mean=np.empty((count, T, D)) if T else None, pred_var=np.empty((count, T)) if T else None,
                   iters=np.empty(count, dtype=np.int32), nfev=np.empty(count, dtype=np.int32), status=np.empty(count, dtype=np.int32))
        lp = np.empty((count, N, 3 * N)) if return_logPhi else None
        tail = [_ptr(b), _ptr(prior), _ptr(phiInitial), _ptr(hyp0), int(kernel), ctypes.addressof(options), T, _ptr(Xnew),
                _ptr(out["objective"]), _ptr(out["elbo"]), _ptr(out["hyp"]), _ptr(out["phi"]), _ptr(out["pred_mean"]),
                _ptr(out["pred_var"]), _ptr(out["iters"]), _ptr(out["nfev"]), _ptr(out["status"]), _ptr(lp)]
        if Z is None:
            rc = self._lib.bgp_dense_fit_host(self._h, count, N, D, _ptr(t), _ptr(Y), nY, _ptr(item_gene), *tail)
            self._check(rc, "bgp_dense_fit_host")
        else:
            Z = _f64(Z)
            assert Z.ndim == 2 and Z.shape[1] == 2
            rc = self._lib.bgp_sparse_fit_host(self._h, count, N, D, Z.shape[0], _ptr(t), _ptr(Z), _ptr(Y), nY, _ptr(item_gene), *tail)
            self._check(rc, "bgp_sparse_fit_host")
        if return_logPhi:
            out["logPhi"] = lp
        return out

    def get_phi(self, logPhi):
        logPhi = _f64(logPhi)
        if logPhi.ndim == 2:
            logPhi = logPhi[None]
        count, N, M = logPhi.shape
        assert M == 3 * N
        out = np.empty((count, N, 3))
        self._check(self._lib.bgp_get_phi_host(self._h, count, N, _ptr(logPhi), _ptr(out)), "bgp_get_phi_host")
        return out

    # ------------------------------------------------------------------ per-kernel timing
    PHASES = ("phi_forward", "prep", "chol_K", "syrk_P", "chol_P", "solve", "trtri", "lauum", "trmm", "post", "adjoint",
              "phi_backward")

    def profile_enable(self, on=True):
        self._check(self._lib.bgp_profile_enable(self._h, int(bool(on))), "bgp_profile_enable")

    def profile_reset(self):
        self._check(self._lib.bgp_profile_reset(self._h), "bgp_profile_reset")

    def profile_read(self):
        """(dict phase -> accumulated ms since the last reset, kernel launches since the last reset); waits for the GPU."""
        ms = np.zeros(len(self.PHASES))
        n = ctypes.c_longlong(0)
        self._check(self._lib.bgp_profile_read(self._h, _ptr(ms), ctypes.byref(n)), "bgp_profile_read")
        return dict(zip(self.PHASES, ms.tolist())), int(n.value)

    # ------------------------------------------------------------------ test hooks
    def debug_set_stop_phase(self, phase):
        self._check(self._lib.bgp_debug_set_stop_phase(self._h, int(phase)), "bgp_debug_set_stop_phase")

    def debug_matrix(self, N, slot, which, symmetric=False):
        Mp = self._lib.bgp_debug_padded_size(int(N))
        out = np.empty((Mp, Mp))
        self._check(self._lib.bgp_debug_dense_matrix(self._h, int(slot), int(which), int(symmetric), _ptr(out)),
                    "bgp_debug_dense_matrix")
        return out

    def debug_vector(self, N, D, slot, which):
        Mp = self._lib.bgp_debug_padded_size(int(N))
        n = Mp * (D if which in (2, 3, 4, 5, 6) else 1)
        out = np.empty(n)
        self._check(self._lib.bgp_debug_dense_vector(self._h, int(slot), int(which), _ptr(out)), "bgp_debug_dense_vector")
        return out.reshape(-1, Mp) if which in (2, 3, 4, 5, 6) else out


_default_engine = None


def default_engine():
    """Process-wide engine on the GPU named by BGP_DEVICE / LOCAL_RANK (default 0)."""
    global _default_engine
    if _default_engine is None:
        dev = int(os.environ.get("BGP_DEVICE", os.environ.get("LOCAL_RANK", "0")))
        _default_engine = Engine(dev)
    return _default_engine
