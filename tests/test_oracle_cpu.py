"""CPU suite: the oracle against itself, against the reference's relational tests and against the golden fixtures."""
import os

import numpy as np
import pytest

from common import make_problem, relerr
from oracle import bgp_numpy, bgp_torch, blocked_dense, host_ref

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "dense_golden.npz"))


def _pz(pr, b):
    return host_ref.prior_with_trunk_bgp(host_ref.expand_prior_bgp(pr["prior"]), b, pr["t"])


@pytest.mark.parametrize("kind", [0, 1])
def test_dense_numpy_adjoint_matches_torch_autograd(kind):
    pr = make_problem(35, D=2, seed=2)
    b, h = pr["b"][0], pr["hyp"][0]
    e_t, g_t = bgp_torch.dense_value_and_grad(pr["logPhi"][0], pr["Y"][0], pr["X"], _pz(pr, b), b, h[0], h[1], h[2], kind)
    e_n, g_n = bgp_numpy.dense_value_and_grad(pr["logPhi"][0], pr["Y"][0], pr["X"], _pz(pr, b), b, h[0], h[1], h[2], kind)
    assert abs(e_t - e_n) <= 1e-12 * abs(e_t)
    assert relerr(g_n["logPhi"], g_t["logPhi"]) <= 1e-9
    for nm in ("ell", "var", "likvar"):
        assert abs(g_t[nm] - g_n[nm]) <= 1e-8 * max(1.0, abs(g_t[nm]))


def test_sparse_numpy_adjoint_matches_torch_autograd():
    pr = make_problem(40, D=1, seed=4)
    b, h = pr["b"][0], pr["hyp"][0]
    Mz = 12
    Z = np.ones((Mz, 2))
    Z[:, 0] = np.linspace(0, 1, Mz, endpoint=False)
    Z[:, 1] = np.array([i for j in range(Mz) for i in range(1, 4)])[:Mz]
    e_t, g_t = bgp_torch.sparse_value_and_grad(pr["logPhi"][0], pr["Y"][0], pr["X"], Z, _pz(pr, b), b, h[0], h[1], h[2])
    e_n, g_n = bgp_numpy.sparse_value_and_grad(pr["logPhi"][0], pr["Y"][0], pr["X"], Z, _pz(pr, b), b, h[0], h[1], h[2])
    assert abs(e_t - e_n) <= 1e-12 * abs(e_t)
    assert relerr(g_n["logPhi"], g_t["logPhi"]) <= 1e-9
    for nm in ("ell", "var", "likvar"):
        assert abs(g_t[nm] - g_n[nm]) <= 1e-8 * max(1.0, abs(g_t[nm]))


def test_mbgp_sparse_numpy_adjoint_matches_torch_autograd():
    pr = make_problem(24, D=3, seed=6, mbgp=True)
    t = pr["t"]
    _, trunk = host_ref.init_logphi_mbgp(pr["phi0"], t)
    eZ0 = host_ref.expand_prior_mbgp(pr["prior3"])
    Zt = np.linspace(0, 1, 8, endpoint=False)
    Z = np.array([[z, f] for z in Zt for f in (1, 2, 3)], dtype=float)
    Bv = np.array([0.3, 0.55, 0.7])
    e_t, g_t = bgp_torch.mbgp_sparse_value_and_grad(pr["logPhi"][0], pr["Y"][0], pr["X"], Z, t, eZ0, trunk, Bv, 0.7, 1.3, 0.2)
    e_n, g_n = bgp_numpy.mbgp_sparse_value_and_grad(pr["logPhi"][0], pr["Y"][0], pr["X"], Z, t, eZ0, trunk, Bv, 0.7, 1.3, 0.2)
    assert abs(e_t - e_n) <= 1e-12 * abs(e_t)
    assert relerr(g_n["logPhi"], g_t["logPhi"]) <= 1e-9
    assert relerr(g_n["Bv"], g_t["Bv"]) <= 1e-8
    for nm in ("ell", "var", "likvar"):
        assert abs(g_t[nm] - g_n[nm]) <= 1e-8 * max(1.0, abs(g_t[nm]))


def test_mbgp_single_dense_equals_multi_sparse_with_Z_equal_X():
    """testing/MBGP/test_equivalence_multi_to_single.py:14-56 (allclose at rtol 1e-5), N = 6, D in {1, 2}."""
    for D in (1, 2):
        pr = make_problem(6, D=D, seed=9, mbgp=True)
        t = pr["t"]
        _, trunk = host_ref.init_logphi_mbgp(pr["phi0"], t)
        eZ0 = host_ref.expand_prior_mbgp(pr["prior3"])
        Bv = np.full(D, 0.5)
        e_s = bgp_numpy.mbgp_sparse_value_and_grad(pr["logPhi"][0], pr["Y"][0], pr["X"], pr["X"], t, eZ0, trunk, Bv, 1.0, 1.0, 1.0,
                                                   want_grad=False)
        e_d, _ = bgp_numpy.mbgp_dense_value_and_grad(pr["logPhi"][0], pr["Y"][0], pr["X"], t, eZ0, trunk, Bv, 1.0, 1.0, 1.0)
        assert np.allclose(e_s, e_d, rtol=1e-5)


def test_relational_KL_convention():
    """testing/test_KL.py:15-103: with a fixed covariance, b = 0.20 scores below b = ptb, and b = ptb equals b = 0.22
    (no cell lies between them), i.e. the trunk convention is t <= b."""
    N = 20
    t = np.linspace(0, 1, N)
    Y = np.zeros((N, 1))
    idx = np.nonzero(t > 0.5)[0]
    Y[idx[::2], 0] = 2 * t[idx[::2]]
    Y[idx[1::2], 0] = -2 * t[idx[1::2]]
    labels = np.ones(N)
    labels[4::2] = 2
    labels[5::2] = 3
    phi0, prior = host_ref.initial_conditions_and_prior(labels, 0.51, False)
    X, _ = host_ref.function_index_list(t)
    ptb = min(t[labels == 2].min(), t[labels == 3].min())
    K1 = bgp_numpy.branch_K(X, X, ptb, 1.0, 1.0, bgp_numpy.MATERN32)
    eZ0 = host_ref.expand_prior_bgp(prior)

    def ll(b):
        pZ = host_ref.prior_with_trunk_bgp(eZ0, b, t)
        lp = host_ref.init_logphi_bgp(phi0, t, b)
        return bgp_numpy.dense_forward(lp, Y, X, pZ, b, 1.0, 1.0, 1.0, KConst=K1)

    assert ll(0.20) < ll(ptb)
    assert np.allclose(ll(ptb), ll(0.22))


def test_block_algorithm_of_the_cuda_pipeline_matches_the_oracle():
    pr = make_problem(23, D=2, seed=0)
    b, h = pr["b"][0], pr["hyp"][0]
    for kind in (0, 1):
        e0, g0 = bgp_numpy.dense_value_and_grad(pr["logPhi"][0], pr["Y"][0], pr["X"], _pz(pr, b), b, h[0], h[1], h[2], kind)
        e1, g1, _ = blocked_dense.dense_value_and_grad_blocked(pr["logPhi"][0], pr["Y"][0], pr["X"], _pz(pr, b), b, h[0], h[1], h[2],
                                                               kind, nb=8)
        assert abs(e0 - e1) <= 1e-12 * abs(e0)
        assert relerr(g1["logPhi"], g0["logPhi"]) <= 1e-9
        for nm in ("ell", "var", "likvar", "b"):
            assert abs(g0[nm] - g1[nm]) <= 1e-8 * max(1.0, abs(g0[nm]))


def test_golden_fixture_is_reproduced_by_the_numpy_oracle():
    from golden.make_golden import CASES
    name = "bgp_n30_matern"
    kw = dict(CASES[name])
    kind = kw.pop("kind")
    pr = make_problem(**kw)
    for k in range(pr["count"]):
        b, h = pr["b"][k], pr["hyp"][k]
        e, g = bgp_numpy.dense_value_and_grad(pr["logPhi"][k], pr["Y"][pr["item_gene"][k]], pr["X"], _pz(pr, b), b, h[0], h[1], h[2], kind)
        assert abs(e - GOLD[name + "/elbo"][k]) <= 1e-11 * abs(e)
        assert np.allclose([g["ell"], g["var"], g["likvar"]], GOLD[name + "/grad_hyp"][k], rtol=1e-7, atol=1e-9)
        N = pr["N"]
        assert relerr(g["logPhi"][np.arange(N), 3 * np.arange(N) + 1], GOLD[name + "/grad_logphi_probe"][k]) <= 1e-8


def test_oracle_fitmodel_reproduces_the_c1_golden_fixture():
    """tests/golden/c1_fitmodel.npz is what oracle.fit_ref produces (two of the ten independent fixed-hyper-parameter fits are
    re-run here; the full fixture takes minutes on CPU)."""
    import os
    sys_path_golden = os.path.join(os.path.dirname(__file__), "golden")
    gold = np.load(os.path.join(sys_path_golden, "c1_fitmodel.npz"))
    from golden.make_golden_c1 import MAXITER, c1_problem
    from oracle import fit_ref
    t, Y, labels, grid = c1_problem()
    sub = [grid[4], grid[9]]
    r = fit_ref.fit_model(sub, t, Y, labels, M=0, maxiter=MAXITER, fixHyperparameters=True)
    np.testing.assert_allclose(r["loglik"], gold["fixed/loglik"][[4, 9]], rtol=1e-9)


@pytest.mark.parametrize("n,nb", [(1, 4), (2, 4), (3, 4), (4, 3), (6, 3), (9, 2)])
def test_scatter_form_task_bodies_equal_the_gather_forms(n, nb):
    """oracle/blocked_dense_wide.py restates, task by task and in queue order, what the multi-CTA pipeline (csrc/bgp_dense_wide.cu)
    does: right-looking tile Cholesky with the fused last update, Gauss-Jordan inverse of the scaled factor, scatter form of the
    blocked Cholesky adjoint with its P / Q partials and BROW rows.  They must give what the gather forms of the one-CTA pipeline
    give (which tests above tie to oracle.bgp_numpy)."""
    from oracle import blocked_dense as bd, blocked_dense_wide as bw
    rng = np.random.default_rng(n * 10 + nb)
    Mp = n * nb
    A = rng.standard_normal((Mp, Mp))
    A = A @ A.T + Mp * np.eye(Mp)
    L1, i1 = bd.chol_left_looking(A, nb)
    L2, i2 = bw.chol_right_looking(A, nb)
    assert np.abs(L1 - L2).max() <= 1e-13 * np.abs(L1).max()
    X1 = np.tril(bd.trtri_inplace(L1, i1, nb))
    X2 = np.tril(bw.trtri_scatter(L1, i1, nb))
    assert np.abs(X1 - X2).max() <= 1e-13 * max(1.0, np.abs(X1).max())
    Lbar = np.tril(rng.standard_normal((Mp, Mp)))
    A1 = bd.chol_adjoint_gather(L1, i1, Lbar, nb)
    A2 = bw.chol_adjoint_scatter(L1, i1, Lbar, nb)
    assert np.abs(A1 - A2).max() <= 1e-12 * max(1.0, np.abs(A1).max())
