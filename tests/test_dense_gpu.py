"""GPU parity tests of the dense AssignGP bound + gradient (through the C ABI) against the CPU oracle.

Tolerances are BASELINE.json's: ELBO <= 1e-8 relative, gradient <= 1e-6 relative, Phi <= 1e-6.
oracle = restatement of TF 2.4.0 / GPflow 2.1.2 semantics (parity unpinned by the reference itself)."""
import os

import numpy as np
import pytest

from common import make_problem, relerr
from oracle import bgp_numpy, bgp_torch, host_ref

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "dense_golden.npz"))
ELBO_RTOL, GRAD_RTOL = 1e-8, 1e-6


@pytest.fixture(autouse=True, params=["one CTA per item", "many CTAs per item"])
def dense_path(request, monkeypatch):
    """Every test of this file runs on both dense pipelines: bgp_dense.cu (one CTA per work item, what batches of >= 74 items
    take) and bgp_dense_wide.cu (the macro tiles of an item dealt to many CTAs, what few items in flight take)."""
    monkeypatch.setenv("BGP_DENSE_WIDE", "0" if request.param == "one CTA per item" else "1")
    return request.param


def _pz(pr, b):
    return host_ref.prior_with_trunk_bgp(host_ref.expand_prior_bgp(pr["prior"]), b, pr["t"])


def _check_item(out, k, e, g, gb=False):
    assert out["status"][k] == 0
    assert abs(out["elbo"][k] - e) <= ELBO_RTOL * abs(e), (out["elbo"][k], e)
    assert relerr(out["grad_logPhi"][k], g["logPhi"]) <= GRAD_RTOL
    names = ("ell", "var", "likvar", "b") if gb else ("ell", "var", "likvar")
    for j, nm in enumerate(names):
        assert abs(out["grad_hyp"][k, j] - g[nm]) <= GRAD_RTOL * max(abs(g[nm]), 1e-3 * np.abs(out["grad_hyp"][k, :3]).max()), \
            (nm, out["grad_hyp"][k, j], g[nm])


@pytest.mark.parametrize("N,D,kind", [(7, 1, 0), (20, 1, 0), (43, 2, 1), (50, 1, 0), (86, 3, 0), (128, 1, 1), (343, 1, 0), (555, 2, 1)])
def test_dense_matches_oracle(engine, N, D, kind):
    """Sizes straddle the 128-row macro block (M = 3N = 21 .. 1665): padding, one block, several blocks, ragged last blocks (M = 1029 in
    nine blocks, 1665 in fourteen) where the task graphs of the many-CTA pipeline have every kind of task."""
    pr = make_problem(N, D=D, n_genes=2, bs=(0.25, 0.6), seed=N)
    out = engine.dense_elbo_grad(pr["t"], pr["Y"], pr["item_gene"], pr["b"], pr["prior"], pr["hyp"], pr["logPhi"], kernel=kind)
    for k in range(pr["count"]):
        b, h = pr["b"][k], pr["hyp"][k]
        e, g = bgp_numpy.dense_value_and_grad(pr["logPhi"][k], pr["Y"][pr["item_gene"][k]], pr["X"], _pz(pr, b), b, h[0], h[1], h[2], kind)
        _check_item(out, k, e, g, gb=True)


def test_dense_matches_golden_fixtures(engine):
    from golden.make_golden import CASES
    for name, kw in CASES.items():
        kw = dict(kw)
        kind = kw.pop("kind")
        pr = make_problem(**kw)
        out = engine.dense_elbo_grad(pr["t"], pr["Y"], pr["item_gene"], pr["b"], pr["prior"], pr["hyp"], pr["logPhi"], kernel=kind)
        N = pr["N"]
        assert np.all(out["status"] == 0)
        assert np.allclose(out["elbo"], GOLD[name + "/elbo"], rtol=ELBO_RTOL, atol=0)
        assert np.allclose(out["grad_hyp"][:, :3], GOLD[name + "/grad_hyp"], rtol=GRAD_RTOL, atol=1e-7)
        probe = out["grad_logPhi"][:, np.arange(N), 3 * np.arange(N) + 1]
        assert relerr(probe, GOLD[name + "/grad_logphi_probe"]) <= GRAD_RTOL
        assert np.allclose(np.abs(out["grad_logPhi"]).sum(axis=(1, 2)), GOLD[name + "/grad_logphi_abs"], rtol=GRAD_RTOL)


def test_bound_only_equals_bound_with_gradient(engine):
    pr = make_problem(40, D=1, n_genes=1, bs=(0.3, 0.5, 1.1), seed=3)
    a = engine.dense_elbo_grad(pr["t"], pr["Y"], pr["item_gene"], pr["b"], pr["prior"], pr["hyp"], pr["logPhi"], want_grad=False)
    b = engine.dense_elbo_grad(pr["t"], pr["Y"], pr["item_gene"], pr["b"], pr["prior"], pr["hyp"], pr["logPhi"], want_grad=True)
    # the bound-only call runs the one-kernel assignment pass, the gradient call the row-resident one (different summation
    # order of the same terms): equal to rounding, not bit for bit
    np.testing.assert_allclose(a["elbo"], b["elbo"], rtol=1e-13)
    assert a["grad_logPhi"] is None


def test_many_items_in_waves_are_independent_and_reproducible(engine):
    """More items than workspace slots: results must not depend on the slot / wave an item lands in."""
    pr = make_problem(16, D=1, n_genes=3, bs=tuple(np.linspace(0.1, 0.9, 7)), seed=8)
    engine.set_max_slots(4)
    try:
        a = engine.dense_elbo_grad(pr["t"], pr["Y"], pr["item_gene"], pr["b"], pr["prior"], pr["hyp"], pr["logPhi"])
    finally:
        engine.set_max_slots(0)
    b = engine.dense_elbo_grad(pr["t"], pr["Y"], pr["item_gene"], pr["b"], pr["prior"], pr["hyp"], pr["logPhi"])
    assert np.array_equal(a["elbo"], b["elbo"])
    assert np.array_equal(a["grad_logPhi"], b["grad_logPhi"])
    assert np.array_equal(a["grad_hyp"], b["grad_hyp"])
    k = 13
    bb, h = pr["b"][k], pr["hyp"][k]
    e, g = bgp_numpy.dense_value_and_grad(pr["logPhi"][k], pr["Y"][pr["item_gene"][k]], pr["X"], _pz(pr, bb), bb, h[0], h[1], h[2])
    _check_item(b, k, e, g)


def test_mbgp_dense_single_b(engine):
    """MBGP _build_likelihood_single_b (MBGP/assigngp.py:608-647): masks from b, squashed pZ, -D*KL, SE kernel."""
    from branchedgp_b200 import _lib
    pr = make_problem(30, D=2, seed=21, mbgp=True)
    t = pr["t"]
    _, trunk = host_ref.init_logphi_mbgp(pr["phi0"], t)
    eZ0 = host_ref.expand_prior_mbgp(pr["prior3"])
    Bv = np.array([0.45, 0.45])
    h = np.array([[0.8, 1.4, 0.3]])
    e, g = bgp_numpy.mbgp_dense_value_and_grad(pr["logPhi"][0], pr["Y"][0], pr["X"], t, eZ0, trunk, Bv, h[0, 0], h[0, 1], h[0, 2])
    out = engine.dense_elbo_grad(t, pr["Y"], pr["item_gene"][:1], Bv[:1], pr["prior3"], h, pr["logPhi"][:1], kernel=_lib.KERNEL_SQEXP,
                                 model=_lib.MODEL_MBGP, kl_weight=2.0)
    assert abs(out["elbo"][0] - e) <= ELBO_RTOL * abs(e)
    assert relerr(out["grad_logPhi"][0], g["logPhi"]) <= GRAD_RTOL
    assert np.allclose(out["grad_hyp"][0], [g["ell"], g["var"], g["likvar"], g["Bv"][0]], rtol=GRAD_RTOL, atol=1e-8)


def test_kconst_relational_test_of_the_reference(engine):
    """testing/test_KL.py:15-103 through the drop-in classes (KConst + fDebug)."""
    from branchedgp_b200 import BranchingTree as bt
    from branchedgp_b200 import FitBranchingModel, VBHelperFunctions, assigngp_dense
    from branchedgp_b200 import branch_kernParamGPflow as bk
    from branchedgp_b200 import gpflow_shim as gpflow
    np.random.seed(43)
    N = 20
    t = np.linspace(0, 1, N)
    Y = np.zeros((N, 1))
    idx = np.nonzero(t > 0.5)[0]
    Y[idx[::2], 0] = 2 * t[idx[::2]]
    Y[idx[1::2], 0] = -2 * t[idx[1::2]]
    labels = np.ones(N)
    labels[4::2] = 2
    labels[5::2] = 3
    XExpanded, indices, _ = VBHelperFunctions.GetFunctionIndexListGeneral(t)
    phiInitial, phiPrior = FitBranchingModel.GetInitialConditionsAndPrior(labels, 0.51, False)
    ptb = np.min([np.min(t[labels == 2]), np.min(t[labels == 3])])
    tree = bt.BinaryBranchingTree(0, 1, fDebug=False)
    tree.add(None, 1, np.ones((1, 1)) * ptb)
    (fm1, _) = tree.GetFunctionBranchTensor()
    K1 = bk.BranchKernelParam(gpflow.kernels.Matern32(), fm1, b=np.ones((1, 1)) * ptb).K(XExpanded, XExpanded)
    kb = bk.BranchKernelParam(gpflow.kernels.Matern32(), fm1, b=np.zeros((1, 1))) + gpflow.kernels.White()
    kb.kernels[1].variance.assign(1e-6)
    gpflow.set_trainable(kb.kernels[1].variance, False)
    m = assigngp_dense.AssignGP(t, XExpanded, Y, kb, indices, np.ones((1, 1)), phiInitial=phiInitial, phiPrior=phiPrior,
                                KConst=K1, fDebug=True)
    m.UpdateBranchingPoint(np.ones((1, 1)) * ptb, phiInitial.copy())
    ptbLL = m.log_posterior_density()
    m.UpdateBranchingPoint(np.ones((1, 1)) * 0.20, phiInitial.copy())
    eLL = m.log_posterior_density()
    m.UpdateBranchingPoint(np.ones((1, 1)) * 0.22, phiInitial.copy())
    lll = m.log_posterior_density()
    assert eLL < ptbLL
    assert np.allclose(ptbLL, lll)
    # and the absolute value against the oracle on the same KConst
    pZ = host_ref.prior_with_trunk_bgp(host_ref.expand_prior_bgp(phiPrior), 0.22, t)
    want = bgp_numpy.dense_forward(m.logPhi.numpy(), Y, XExpanded, pZ, 0.22, 1.0, 1.0, 1.0, KConst=K1)
    assert abs(lll - want) <= ELBO_RTOL * abs(want)


def test_get_phi_matches_softmax_gather(engine):
    pr = make_problem(33, seed=2)
    phi = engine.get_phi(pr["logPhi"])
    S = bgp_numpy.softmax_rows(pr["logPhi"][0])
    want = np.stack([S[np.arange(33), 3 * np.arange(33) + f] for f in range(3)], 1)
    assert np.abs(phi[0] - want).max() <= 1e-12


def test_non_positive_definite_input_is_reported(engine):
    from branchedgp_b200 import _lib
    pr = make_problem(12, seed=2)
    K = -np.eye(36)
    out = engine.dense_elbo_grad(pr["t"], pr["Y"], pr["item_gene"], pr["b"], pr["prior"], pr["hyp"], pr["logPhi"], KConst=K, check=False)
    assert out["status"][0] != 0
    with pytest.raises(_lib.CholeskyError):
        engine.dense_elbo_grad(pr["t"], pr["Y"], pr["item_gene"], pr["b"], pr["prior"], pr["hyp"], pr["logPhi"], KConst=K)


def test_full_size_n1000_item_matches_both_oracles(engine):
    """BASELINE configs[1] shape: N = 1000 cells (M = 3000), one work item against numpy hand-adjoint AND torch autograd."""
    pr = make_problem(1000, D=1, seed=3)
    b, h = pr["b"][0], pr["hyp"][0]
    out = engine.dense_elbo_grad(pr["t"], pr["Y"], pr["item_gene"], pr["b"], pr["prior"], pr["hyp"], pr["logPhi"])
    e, g = bgp_numpy.dense_value_and_grad(pr["logPhi"][0], pr["Y"][0], pr["X"], _pz(pr, b), b, h[0], h[1], h[2])
    _check_item(out, 0, e, g)
    e2, g2 = bgp_torch.dense_value_and_grad(pr["logPhi"][0], pr["Y"][0], pr["X"], _pz(pr, b), b, h[0], h[1], h[2])
    _check_item(out, 0, e2, g2)


def test_full_size_properties_without_oracle(engine):
    """Size-independent properties at N = 1000: (i) the softmax gradient of every row sums to ~0 (softmax invariance to
    a per-row shift); (ii) shifting every logit of a row by a constant leaves the bound unchanged; (iii) doubling Y and
    quadrupling nothing else changes only the data terms in the known way is NOT used -- instead a directional
    finite difference along the hyper-parameters matches grad_hyp."""
    pr = make_problem(1000, D=1, n_genes=1, bs=(0.4,), seed=5)
    args = (pr["t"], pr["Y"], pr["item_gene"], pr["b"], pr["prior"])
    out = engine.dense_elbo_grad(*args, pr["hyp"], pr["logPhi"])
    g = out["grad_logPhi"][0]
    assert np.abs(g.sum(axis=1)).max() <= 1e-9 * np.abs(g).sum(axis=1).max()
    lp2 = pr["logPhi"] + np.random.default_rng(0).standard_normal((1, 1000, 1))
    out2 = engine.dense_elbo_grad(*args, pr["hyp"], lp2, want_grad=False)
    assert abs(out2["elbo"][0] - out["elbo"][0]) <= 1e-9 * abs(out["elbo"][0])
    d = np.array([[0.3, -0.2, 0.1]])
    eps = 1e-5
    ep = engine.dense_elbo_grad(*args, pr["hyp"] + eps * d, pr["logPhi"], want_grad=False)["elbo"][0]
    em = engine.dense_elbo_grad(*args, pr["hyp"] - eps * d, pr["logPhi"], want_grad=False)["elbo"][0]
    fd = (ep - em) / (2 * eps)
    an = float((out["grad_hyp"][0, :3] * d[0]).sum())
    assert abs(fd - an) <= 1e-5 * max(1.0, abs(an)), (fd, an)


def test_c5_shape_n2000_item_matches_oracle_and_properties(engine):
    """BASELINE configs[4] shape: N = 2000 cells (M = 6000, 47 macro blocks, 0.7 GB workspace per item).  Two items: one against
    the numpy oracle, and the size-independent properties (row-shift invariance, zero row sums of the softmax gradient,
    directional finite difference along the hyper-parameters) on both."""
    pr = make_problem(2000, D=1, n_genes=1, bs=(0.35, 0.7), seed=9)
    args = (pr["t"], pr["Y"], pr["item_gene"], pr["b"], pr["prior"])
    out = engine.dense_elbo_grad(*args, pr["hyp"], pr["logPhi"])
    assert np.all(out["status"] == 0)
    b, h = pr["b"][0], pr["hyp"][0]
    e, g = bgp_numpy.dense_value_and_grad(pr["logPhi"][0], pr["Y"][0], pr["X"], _pz(pr, b), b, h[0], h[1], h[2])
    _check_item(out, 0, e, g)
    for k in range(2):
        gk = out["grad_logPhi"][k]
        assert np.abs(gk.sum(axis=1)).max() <= 1e-9 * np.abs(gk).sum(axis=1).max()
    lp2 = pr["logPhi"] + np.random.default_rng(0).standard_normal((2, 2000, 1))
    out2 = engine.dense_elbo_grad(*args, pr["hyp"], lp2, want_grad=False)
    assert np.all(np.abs(out2["elbo"] - out["elbo"]) <= 1e-9 * np.abs(out["elbo"]))
    d = np.array([[0.3, -0.2, 0.1]])
    eps = 1e-5
    ep = engine.dense_elbo_grad(*args, pr["hyp"] + eps * d, pr["logPhi"], want_grad=False)["elbo"]
    em = engine.dense_elbo_grad(*args, pr["hyp"] - eps * d, pr["logPhi"], want_grad=False)["elbo"]
    fd = (ep - em) / (2 * eps)
    an = (out["grad_hyp"][:, :3] * d).sum(axis=1)
    assert np.all(np.abs(fd - an) <= 1e-5 * np.maximum(1.0, np.abs(an))), (fd, an)


def test_dense_long_rows_streaming_passes(engine, monkeypatch):
    """N = 2100 (rows of 6300 columns, beyond one CTA's registers): the streaming assignment passes (16-row chunks in the dense
    workspace) AND the two-pass / cluster kernels they replace (BGP_PHI_STREAM=0), both against the numpy oracle, plus the
    zero row sums of a softmax gradient."""
    pr = make_problem(2100, D=1, n_genes=1, bs=(0.55,), seed=19)
    args = (pr["t"], pr["Y"], pr["item_gene"], pr["b"], pr["prior"])
    outs = []
    for flag in ("1", "0"):
        monkeypatch.setenv("BGP_PHI_STREAM", flag)
        outs.append(engine.dense_elbo_grad(*args, pr["hyp"], pr["logPhi"]))
    monkeypatch.delenv("BGP_PHI_STREAM")
    a, b = outs
    bb, h = pr["b"][0], pr["hyp"][0]
    e, g = bgp_numpy.dense_value_and_grad(pr["logPhi"][0], pr["Y"][0], pr["X"], _pz(pr, bb), bb, h[0], h[1], h[2])
    for out in outs:
        assert out["status"][0] == 0
        _check_item(out, 0, e, g)
    assert abs(a["elbo"][0] - b["elbo"][0]) <= 1e-12 * abs(b["elbo"][0])
    assert relerr(a["grad_logPhi"][0], b["grad_logPhi"][0]) <= 1e-10
    assert np.allclose(a["grad_hyp"], b["grad_hyp"], rtol=1e-9, atol=0)
    gl = a["grad_logPhi"][0]
    assert np.abs(gl.sum(axis=1)).max() <= 1e-9 * np.abs(gl).sum(axis=1).max()


def test_automatic_path_selection_and_wave_split(engine, monkeypatch):
    """Without BGP_DENSE_WIDE the library picks the pipeline from the number of items in flight: many CTAs per item up to half an
    SM count of items, two such waves between half and three quarters, one CTA per item above.  100 items (between the two
    thresholds on a 148-SM part) and 5 items must agree with the forced one-CTA pipeline to round-off, and 3 of them with the oracle."""
    pr = make_problem(14, D=1, n_genes=10, bs=tuple(np.linspace(0.1, 0.9, 10)), seed=21)
    args = (pr["t"], pr["Y"], pr["item_gene"], pr["b"], pr["prior"], pr["hyp"], pr["logPhi"])
    assert pr["count"] == 100
    monkeypatch.setenv("BGP_DENSE_WIDE", "0")
    ref = engine.dense_elbo_grad(*args)
    monkeypatch.delenv("BGP_DENSE_WIDE")
    auto = engine.dense_elbo_grad(*args)
    sub = engine.dense_elbo_grad(pr["t"], pr["Y"], pr["item_gene"][:5], pr["b"][:5], pr["prior"], pr["hyp"][:5], pr["logPhi"][:5])
    assert np.all(ref["status"] == 0) and np.all(auto["status"] == 0) and np.all(sub["status"] == 0)
    assert np.allclose(auto["elbo"], ref["elbo"], rtol=1e-12, atol=0)
    assert np.allclose(sub["elbo"], ref["elbo"][:5], rtol=1e-12, atol=0)
    assert relerr(auto["grad_logPhi"], ref["grad_logPhi"]) <= 1e-10
    assert np.allclose(auto["grad_hyp"], ref["grad_hyp"], rtol=1e-8, atol=1e-9 * np.abs(ref["grad_hyp"]).max())
    for k in (0, 57, 99):
        b, h = pr["b"][k], pr["hyp"][k]
        e, g = bgp_numpy.dense_value_and_grad(pr["logPhi"][k], pr["Y"][pr["item_gene"][k]], pr["X"], _pz(pr, b), b, h[0], h[1], h[2])
        _check_item(auto, k, e, g, gb=True)


@pytest.mark.parametrize("N,reps", [(300, 30), (1000, 8)])
def test_repeated_evaluations_are_bit_identical(engine, N, reps):
    """Determinism / race stress: the same evaluation repeated must return the same bits every time.  On the many-CTA pipeline the
    tile tasks are taken by whichever CTA is free and hand their results on through release / acquire counters, TMA reads and
    L2-only loads; a missing dependency or a stale read would show up here as a run-to-run difference."""
    pr = make_problem(N, D=1, n_genes=1, bs=(0.35, 0.6) if N <= 300 else (0.5,), seed=123)
    args = (pr["t"], pr["Y"], pr["item_gene"], pr["b"], pr["prior"], pr["hyp"], pr["logPhi"])
    first = engine.dense_elbo_grad(*args)
    assert np.all(first["status"] == 0)
    for _ in range(reps):
        again = engine.dense_elbo_grad(*args)
        assert np.array_equal(again["elbo"], first["elbo"])
        assert np.array_equal(again["grad_hyp"], first["grad_hyp"])
        assert np.array_equal(again["grad_logPhi"], first["grad_logPhi"])
