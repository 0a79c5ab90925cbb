"""CPU suite: host-side mirror of the reference API (no GPU calls)."""
import numpy as np
import pytest

from branchedgp_b200 import BranchingTree as bt
from branchedgp_b200 import FitBranchingModel, VBHelperFunctions, pZ_construction_singleBP
from branchedgp_b200 import branch_kernParamGPflow as bk
from branchedgp_b200 import gpflow_shim as gpflow
from oracle import bgp_numpy, host_ref


def test_initial_conditions_and_prior_match_loop_restatement():
    rng = np.random.default_rng(1)
    labels = rng.integers(1, 4, 57).astype(float)
    for inf in (True, False):
        a, b = FitBranchingModel.GetInitialConditionsAndPrior(labels, 0.8, inf)
        c, d = host_ref.initial_conditions_and_prior(labels, 0.8, inf)
        assert np.array_equal(a, c) and np.array_equal(b, d)


def test_function_index_list_is_cell_major():
    t = np.linspace(0, 1, 7)
    X, idx, XS = VBHelperFunctions.GetFunctionIndexListGeneral(t)
    X2, idx2 = host_ref.function_index_list(t)
    assert np.array_equal(X, X2) and idx == idx2
    assert XS.shape == (7, 2) and set(XS[:, 1]) <= {1.0, 2.0, 3.0}


def test_pZ_construction_matches_reference_rules():
    """testing/test_pZ_construct.py: trunk rows are [1, eps, eps]; branch rows keep the prior in columns 1, 2."""
    t = np.linspace(0, 1, 9)
    prior = np.tile([0.3, 0.7], (9, 1))
    e = pZ_construction_singleBP.expand_pZ0Zeros(prior)
    assert np.array_equal(e, host_ref.expand_prior_bgp(prior))
    p = pZ_construction_singleBP.expand_pZ0PureNumpyZeros(e, np.ones((1, 1)) * 0.5, t)
    assert np.array_equal(p, host_ref.prior_with_trunk_bgp(e, 0.5, t))
    for i in range(9):
        blk = p[i, 3 * i:3 * i + 3]
        if t[i] <= 0.5:
            assert np.allclose(blk, [1, 1e-6, 1e-6], atol=1e-12)
        else:
            assert np.allclose(blk, [1e-6, 0.3, 0.7], atol=1e-12)
    with pytest.raises(AssertionError):
        pZ_construction_singleBP.expand_pZ0Zeros(np.ones((3, 3)) / 3)


def test_branching_evidence_and_posterior():
    ll = np.array([-10.0, -3.0, -7.0, -20.0])
    B = [0.1, 0.4, 0.7, 1.1]
    ev = VBHelperFunctions.CalculateBranchingEvidence({"loglik": ll}, B)
    w = np.exp(ll[:-1] - ll[:-1].max())
    assert np.allclose(ev["posteriorBranching"], w / w.sum())
    assert np.isclose(ev["logBayesFactor"], np.log(np.exp(ll[:-1]).sum()) - ll[-1] - np.log(3))
    with pytest.raises(NameError):
        VBHelperFunctions.CalculateBranchingEvidence({"loglik": ll}, B[:-1])
    post = FitBranchingModel.GetPosteriorB(ll, B)
    assert post["Bmode"] == 0.4 and post["idx_mode"] == 1
    assert list(post["BgridSearch_sort"]) == B


def test_tree_tensor_for_one_branch_point():
    tree = bt.BinaryBranchingTree(0, 1)
    tree.add(None, 1, np.ones((1, 1)) * 0.3)
    fm, _ = tree.GetFunctionBranchTensor()
    assert fm.shape == (3, 3, 1)
    assert np.all(fm[~np.eye(3, dtype=bool), 0] == 1) and np.all(np.isnan(np.diag(fm[:, :, 0])))


def test_parameter_transforms_and_kernel_container():
    p = gpflow.Parameter(2.5, transform=gpflow.positive(1e-6))
    assert np.isclose(float(p.numpy()), 2.5)
    u = p.unconstrained_variable
    h = 1e-6
    num = (gpflow.positive(1e-6).forward(u + h) - gpflow.positive(1e-6).forward(u - h)) / (2 * h)
    assert np.isclose(num, p.transform.dforward(u), rtol=1e-6)
    s = gpflow.Sigmoid()
    assert np.isclose(s.forward(s.inverse(0.3)), 0.3)
    tree = bt.BinaryBranchingTree(0, 1)
    tree.add(None, 1, np.ones((1, 1)) * 0.3)
    fm, _ = tree.GetFunctionBranchTensor()
    kb = bk.BranchKernelParam(gpflow.kernels.Matern32(1), fm, b=np.ones((1, 1)) * 0.4) + gpflow.kernels.White(1)
    kb.kernels[1].variance.assign(1e-6)
    gpflow.set_trainable(kb.kernels[1].variance, False)
    assert kb.kernels[0].name == "branch_kernel_param" and not kb.kernels[1].variance.trainable
    kb.kernels[0].kern.lengthscales.assign(1.7)
    X, _, _ = VBHelperFunctions.GetFunctionIndexListGeneral(np.linspace(0, 1, 5))
    K = kb.kernels[0].K(X)
    assert np.allclose(K, bgp_numpy.branch_K(X, X, 0.4, 1.7, 1.0, 0), rtol=1e-13)
    n = gpflow.Normal(2.0, 1.0)
    assert np.isclose(n.log_prob(2.0), -0.5 * np.log(2 * np.pi))


def test_linesearch_hook_matches_scipy_dcsrch():
    """The More'-Thuente line search of the batched fit (bgp_fit.cu) visits the same trial steps as SciPy's DCSRCH."""
    import ctypes
    from scipy.optimize._dcsrch import DCSRCH
    from branchedgp_b200 import _lib
    lib = _lib.load_library()
    rng = np.random.default_rng(1)
    checked = 0
    for _ in range(60):
        a, b, c, e = rng.standard_normal(4)

        def phi(x):
            return -abs(a) * x + (abs(b) + 0.1) * x * x + 0.3 * c * np.sin(3 * x) * x * x + 0.05 * e * x ** 3 * np.cos(x)

        def dphi(x):
            return (-abs(a) + 2 * (abs(b) + 0.1) * x + 0.3 * c * (3 * np.cos(3 * x) * x * x + 2 * x * np.sin(3 * x))
                    + 0.05 * e * (3 * x * x * np.cos(x) - x ** 3 * np.sin(x)))
        if dphi(0.0) >= 0:
            continue
        stp0 = float(rng.choice([1.0, 0.1, 5.0, 1e-3]))
        ref_trace = []

        def phi_logged(x):
            ref_trace.append(x)
            return phi(x)
        DCSRCH(phi_logged, dphi, 1e-3, 0.9, 0.1, 0.0, 1e10)(stp0, phi0=phi(0.0), derphi0=dphi(0.0), maxiter=30)
        state = np.zeros(32)
        stp = ctypes.c_double(stp0)
        task = lib.bgp_debug_linesearch(state.ctypes.data, 1, ctypes.byref(stp), phi(0.0), dphi(0.0))
        trace = []
        while task == 0 and len(trace) < 30:
            trace.append(stp.value)
            task = lib.bgp_debug_linesearch(state.ctypes.data, 0, ctypes.byref(stp), phi(stp.value), dphi(stp.value))
        assert len(trace) == len(ref_trace)
        np.testing.assert_allclose(trace, ref_trace, rtol=1e-13)
        checked += 1
    assert checked > 20


def test_direction_hook_is_the_two_loop_recursion():
    from branchedgp_b200 import _lib
    lib = _lib.load_library()
    rng = np.random.default_rng(0)
    n = 60
    A = rng.standard_normal((n, n))
    H = A @ A.T + n * np.eye(n)
    for col in (0, 1, 4, 10):
        S = rng.standard_normal((max(col, 1), n))
        Y = S @ H
        g = rng.standard_normal(n)
        d = np.empty(n)
        assert lib.bgp_debug_lbfgs_direction(col, n, S.ctypes.data, Y.ctypes.data, g.ctypes.data, d.ctypes.data) == 0
        q = g.copy()
        alphas = []
        for s, y in zip(S[:col][::-1], Y[:col][::-1]):
            al = (s @ q) / (s @ y)
            q -= al * y
            alphas.append(al)
        if col:
            q *= (S[col - 1] @ Y[col - 1]) / (Y[col - 1] @ Y[col - 1])
        for (s, y), al in zip(zip(S[:col], Y[:col]), alphas[::-1]):
            q += (al - (y @ q) / (s @ y)) * s
        np.testing.assert_allclose(d, -q, rtol=1e-10, atol=1e-14)


def test_branching_tree_domain_rule_matches_reference_golden():
    """BinaryBranchingTree.GetFunctionIndexList / GetFunctionDomains of the one-branch-point tree against values produced by the
    reference's own class (tests/golden/make_golden_tree.py): trunk for x <= b, both branches after it, the same draws from
    numpy's global generator, NameError outside (lb, ub]."""
    import os
    from branchedgp_b200 import BranchingTree as bt
    G = np.load(os.path.join(os.path.dirname(__file__), "golden", "tree_index_list.npz"))
    tree = bt.BinaryBranchingTree(0, 1)
    tree.add(None, 1, 0.4)
    np.random.seed(3)
    Xnew, idx, Xtrue = tree.GetFunctionIndexList(G["x"], fReturnXtrue=True)
    assert np.array_equal(Xnew, G["Xnew"]) and np.array_equal(Xtrue, G["Xtrue"])
    assert np.array_equal(np.array([i for ix in idx for i in ix]), G["idx_flat"])
    assert np.array_equal(np.array([len(ix) for ix in idx]), G["idx_len"])
    assert np.array_equal(tree.GetFunctionDomains(), G["domains"])
    fm, _ = tree.GetFunctionBranchTensor()
    assert np.array_equal(np.isnan(fm), np.isnan(G["fm"])) and np.array_equal(fm[~np.isnan(fm)], G["fm"][~np.isnan(G["fm"])])
    for bad in ([0.0, 0.5], [0.5, 1.2]):
        with pytest.raises(NameError):
            tree.GetFunctionIndexList(np.array(bad))


def test_scipy_shim_in_place_gradients_and_pinned_parameter_storage():
    """gpflow_shim.optimizers.Scipy with a model that writes its gradients in place (the path the CUDA models take) reaches the
    same minimiser as plain SciPy; a Parameter whose storage was moved into caller-provided ("pinned") memory keeps that storage
    across assign / set_unconstrained."""
    import scipy.optimize
    from branchedgp_b200 import gpflow_shim as gpflow
    rng = np.random.default_rng(0)
    A = rng.standard_normal((30, 30))
    A = A @ A.T + 30 * np.eye(30)
    bvec = rng.standard_normal(30)

    class Quad:
        def __init__(self):
            self.x = gpflow.Parameter(np.zeros((5, 6)), name="x")
            self.s = gpflow.Parameter(1.5, transform=gpflow.positive(), name="s")
            self.calls = 0

        def training_loss(self):
            raise AssertionError("the shim must call _loss_and_grad")

        def _loss_and_grad(self, variables, out=None):
            self.calls += 1
            x = self.x.view().ravel()
            s = float(self.s.numpy())
            f = 0.5 * x @ A @ x - bvec @ x + (s - 2.0) ** 2
            gx = (A @ x - bvec).reshape(5, 6)
            gs = np.array(2.0 * (s - 2.0)) * self.s.transform.dforward(self.s.unconstrained_variable)
            res = []
            for i, p in enumerate(variables):
                g = gx if p is self.x else gs
                np.copyto(out[i], g)
                res.append(out[i])
            return f, res

    m = Quad()
    store = np.empty((5, 6))
    m.x.pin(lambda shape: store)
    assert m.x.view() is store
    gpflow.optimizers.Scipy().minimize(m.training_loss, [m.x, m.s], options=dict(maxiter=200))
    want = np.linalg.solve(A, bvec)
    assert np.allclose(m.x.numpy().ravel(), want, atol=1e-5) and abs(float(m.s.numpy()) - 2.0) < 1e-5
    assert m.x.view() is store and m.calls > 3                     # still the caller's storage after the optimiser's updates
    m.x.assign(np.ones((5, 6)))
    assert m.x.view() is store and np.all(store == 1.0)
