"""Run the dense CUDA pipeline phase by phase on a small problem and print the error of every intermediate
against the numpy oracle's state (GPU box only; used while bringing kernels up)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from branchedgp_b200 import _lib  # noqa: E402
from common import make_problem, relerr  # noqa: E402
from oracle import bgp_numpy as ref, host_ref  # noqa: E402


def main(N=50, D=1, kind=0):
    pr = make_problem(N, D=D, seed=1)
    eng = _lib.Engine(0)
    t, b, hyp = pr["t"], pr["b"][0], pr["hyp"][0]
    M = 3 * N
    pZ = host_ref.prior_with_trunk_bgp(host_ref.expand_prior_bgp(pr["prior"]), b, t)
    e0, g0, st = ref.dense_value_and_grad(pr["logPhi"][0], pr["Y"][0], pr["X"], pZ, b, hyp[0], hyp[1], hyp[2], kind, want_state=True)
    Mp = eng._lib.bgp_debug_padded_size(N)

    def pad(A, ident=False):
        out = np.eye(Mp) if ident else np.zeros((Mp, Mp))
        out[:M, :M] = A
        return out

    def run(phase):
        eng.debug_set_stop_phase(phase)
        return eng.dense_elbo_grad(t, pr["Y"], pr["item_gene"][:1], pr["b"][:1], pr["prior"], pr["hyp"][:1], pr["logPhi"][:1],
                                   kernel=kind, check=False)

    out = run(1)
    print("status", out["status"])
    print("A     ", relerr(eng.debug_vector(N, D, 0, 0)[:M], st["A"]))
    print("v     ", relerr(eng.debug_vector(N, D, 0, 2)[:, :M], st["v"].T))
    LC = eng.debug_matrix(N, 0, 0)
    print("Lc    ", relerr(LC, pad(st["Lc"], True)))
    run(2)
    P = eng.debug_matrix(N, 0, 1)
    print("P     ", relerr(np.tril(P), np.tril(pad(st["P"], True))))
    run(3)
    R = eng.debug_matrix(N, 0, 1)
    print("R     ", relerr(R, pad(st["R"], True)))
    run(4)
    print("z     ", relerr(eng.debug_vector(N, D, 0, 3)[:, :M], st["z"].T))
    print("c     ", relerr(eng.debug_vector(N, D, 0, 4)[:, :M], st["c"].T))
    print("q     ", relerr(eng.debug_vector(N, D, 0, 5)[:, :M], st["q"].T))
    run(5)
    Xi = eng.debug_matrix(N, 0, 1)
    print("Rinv  ", relerr(Xi[:M, :M], np.linalg.inv(st["R"])))
    run(6)
    PB = eng.debug_matrix(N, 0, 2, True)
    print("Pbar  ", relerr(PB[:M, :M], st["Pbar"]))
    run(7)
    LB = eng.debug_matrix(N, 0, 3)
    print("Lbar  ", relerr(LB[:M, :M], np.tril(st["Lbar"])))
    run(8)
    print("delta ", relerr(eng.debug_vector(N, D, 0, 7)[:M], st["delta"]))
    print("vbar  ", relerr(eng.debug_vector(N, D, 0, 6)[:, :M], st["vbar"].T))
    out = run(0)
    KB = eng.debug_matrix(N, 0, 3, True)
    print("2Kbar ", relerr(KB[:M, :M], 2 * st["Kbar"]))
    print("elbo  ", out["elbo"][0], e0, abs(out["elbo"][0] - e0) / abs(e0))
    gh = out["grad_hyp"][0]
    print("g_ell ", gh[0], g0["ell"], "g_var", gh[1], g0["var"], "g_lik", gh[2], g0["likvar"], "g_b", gh[3], g0["b"])
    print("g_phi ", relerr(out["grad_logPhi"][0], g0["logPhi"]))
    eng.debug_set_stop_phase(0)


if __name__ == "__main__":
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 50
    D = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    kind = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    main(N, D, kind)
