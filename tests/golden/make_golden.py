"""Generates tests/golden/dense_golden.npz from the CPU oracle (run in the build container, CPU only).

The reference (TensorFlow 2.4 / GPflow 2.1.2) is not installable here, so the "golden" values are produced by the
op-for-op torch restatement (oracle/bgp_torch.py, gradients by autograd) and cross-checked against the independent numpy
hand-adjoint (oracle/bgp_numpy.py) before being written: PARITY UNPINNED by the reference itself (oracle/__init__.py).
Inputs are regenerated from seeds by tests/common.make_problem, so only outputs are stored.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from common import make_problem, relerr  # noqa: E402
from oracle import bgp_numpy, bgp_torch, host_ref  # noqa: E402

CASES = {   # name -> make_problem kwargs + kernel kind
    "bgp_n30_matern": dict(N=30, D=1, n_genes=2, bs=(0.2, 0.5, 1.1), seed=11, kind=0),
    "bgp_n45_sqexp_d2": dict(N=45, D=2, n_genes=1, bs=(0.35, 0.8), seed=12, kind=1),
    "bgp_n100_matern": dict(N=100, D=1, n_genes=1, bs=(0.45,), seed=13, kind=0),
}


def main():
    out = {}
    for name, kw in CASES.items():
        kw = dict(kw)
        kind = kw.pop("kind")
        pr = make_problem(**kw)
        elbo, gh, gsum, gabs, gprobe = [], [], [], [], []
        for k in range(pr["count"]):
            b, h = pr["b"][k], pr["hyp"][k]
            pZ = host_ref.prior_with_trunk_bgp(host_ref.expand_prior_bgp(pr["prior"]), b, pr["t"])
            Y = pr["Y"][pr["item_gene"][k]]
            e_t, g_t = bgp_torch.dense_value_and_grad(pr["logPhi"][k], Y, pr["X"], pZ, b, h[0], h[1], h[2], kind)
            e_n, g_n = bgp_numpy.dense_value_and_grad(pr["logPhi"][k], Y, pr["X"], pZ, b, h[0], h[1], h[2], kind)
            assert abs(e_t - e_n) <= 1e-11 * abs(e_t), (name, k, e_t, e_n)
            assert relerr(g_n["logPhi"], g_t["logPhi"]) < 1e-8
            for nm in ("ell", "var", "likvar"):
                assert abs(g_t[nm] - g_n[nm]) <= 1e-7 * max(1.0, abs(g_t[nm])), (name, k, nm, g_t[nm], g_n[nm])
            elbo.append(e_t)
            gh.append([g_t["ell"], g_t["var"], g_t["likvar"]])
            gsum.append(g_t["logPhi"].sum())
            gabs.append(np.abs(g_t["logPhi"]).sum())
            N = pr["N"]
            gprobe.append(g_t["logPhi"][np.arange(N), 3 * np.arange(N) + 1].copy())   # gradient at each cell's own branch-1 logit
        out[name + "/elbo"] = np.array(elbo)
        out[name + "/grad_hyp"] = np.array(gh)
        out[name + "/grad_logphi_sum"] = np.array(gsum)
        out[name + "/grad_logphi_abs"] = np.array(gabs)
        out[name + "/grad_logphi_probe"] = np.array(gprobe)
        print(name, elbo)
    np.savez(os.path.join(HERE, "dense_golden.npz"), **out)


if __name__ == "__main__":
    main()
