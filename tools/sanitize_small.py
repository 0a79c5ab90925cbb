"""Small end-to-end calls for compute-sanitizer (memcheck): both dense pipelines, predict with a tiled full covariance, the batched
sampler, GetPhi and the inducing-point bound at sizes that finish in seconds under the tool.
usage (GPU box): compute-sanitizer --tool memcheck python tools/sanitize_small.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from branchedgp_b200 import _lib  # noqa: E402
from common import make_problem  # noqa: E402


def main():
    eng = _lib.Engine(0)
    for N, wide in ((20, "1"), (20, "0"), (90, "1"), (90, "0")):
        os.environ["BGP_DENSE_WIDE"] = wide
        pr = make_problem(N, D=1, n_genes=2, bs=(0.3, 0.7), seed=N)
        out = eng.dense_elbo_grad(pr["t"], pr["Y"], pr["item_gene"], pr["b"], pr["prior"], pr["hyp"], pr["logPhi"])
        assert np.all(out["status"] == 0) and np.all(np.isfinite(out["grad_logPhi"]))
        print("dense", N, "wide" if wide == "1" else "narrow", out["elbo"])
    os.environ.pop("BGP_DENSE_WIDE")
    rng = np.random.default_rng(0)
    Xn = np.stack([np.sort(rng.uniform(0, 1, 150)), rng.integers(1, 4, 150).astype(float)], 1)
    mean, var = eng.dense_predict(pr["t"], pr["Y"], pr["item_gene"][:1], pr["b"][:1], pr["prior"], pr["hyp"][:1], pr["logPhi"][:1], Xn, full_cov=True)
    print("predict full covariance", var.shape, float(np.abs(var - np.transpose(var, (0, 2, 1))).max()))
    z = rng.standard_normal((3, 70, 2))
    s = eng.gaussian_samples(z, X=Xn[:70], bv=np.array([0.3, 0.5, 0.7]), ell=0.7, kvar=1.3)
    print("samples", s.shape, bool(np.all(np.isfinite(s))))
    print("phi", eng.get_phi(pr["logPhi"]).shape)
    Z = np.ones((12, 2))
    Z[:, 0] = np.linspace(0, 1, 12, endpoint=False)
    Z[:, 1] = np.array([i for j in range(12) for i in range(1, 4)])[:12]
    so = eng.sparse_elbo_grad(pr["t"], Z, pr["Y"], pr["item_gene"], pr["b"], pr["prior"], pr["hyp"], pr["logPhi"])
    print("sparse", so["elbo"])


if __name__ == "__main__":
    main()
