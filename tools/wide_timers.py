"""Cycle breakdown of the critical-path tasks of the multi-CTA Cholesky (needs the instrumented library: make -C branchedgp_b200/csrc timing;
BGP_LIB_NAME=libbgp_b200_timing.so python tools/wide_timers.py [N=1000])."""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from branchedgp_b200 import _lib  # noqa: E402
from common import make_problem  # noqa: E402

NAMES = {0: "DIAG tasks", 1: "DIAG wait deps", 2: "DIAG update gemm", 3: "DIAG epilogue -> smem", 4: "DIAG chol128", 5: "DIAG logdet + store L",
         6: "DIAG trinv128", 7: "DIAG store inverse + signal", 8: "PANEL tasks", 9: "PANEL wait + update gemm + store", 10: "PANEL wait diag",
         11: "PANEL x inv gemm + store + signal", 12: "UPD tasks", 13: "UPD wait deps", 14: "UPD gemm + RMW + signal",
         16: "adj APANEL(K,K-1) tasks (critical)", 17: "wait deps", 18: "diag-term gemm + Q sum + store E", 19: "x inv gemm + store", 20: "P gemm + store + signal",
         21: "Q gemm + hyper partial + signal",
         24: "adj APANEL(I>K) tasks", 25: "wait deps", 26: "scatter gemm + store E", 27: "x inv gemm + store", 28: "P gemm + wait + store + signal",
         29: "Q gemm + wait + hyper partial + signal",
         32: "adj ADIAG tasks", 33: "wait deps", 34: "sum of P partials -> Leff", 35: "three 128^3 products", 36: "symmetrise + hyper partial + signal",
         38: "adj BROW tasks", 39: "wait deps", 40: "long gemm + RMW + signal", 42: "adj SC tasks", 43: "wait deps", 44: "gemm + RMW + signal"}


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
    os.environ["BGP_DENSE_WIDE"] = "1"
    eng = _lib.default_engine()
    pr = make_problem(N, D=1, n_genes=1, bs=(0.5,), seed=1)
    args = (pr["t"], pr["Y"], pr["item_gene"], pr["b"], pr["prior"], pr["hyp"], pr["logPhi"])
    eng.dense_elbo_grad(*args)
    out = np.zeros(48, dtype=np.uint64)
    eng._lib.bgp_debug_wide_timers(eng._h, out.ctypes.data_as(ctypes.c_void_p), 1)
    eng.dense_elbo_grad(*args)
    eng._lib.bgp_debug_wide_timers(eng._h, out.ctypes.data_as(ctypes.c_void_p), 1)
    mhz = 1965.0
    for grp, cnt in ((range(1, 8), 0), (range(9, 12), 8), (range(13, 15), 12), (range(17, 22), 16), (range(25, 30), 24), (range(33, 37), 32),
                     (range(39, 41), 38), (range(43, 45), 42)):
        n = max(int(out[cnt]), 1)
        print("%s: %d" % (NAMES[cnt], int(out[cnt])))
        for k in grp:
            print("   %-40s %9.2f us per task" % (NAMES[k], float(out[k]) / n / mhz))


if __name__ == "__main__":
    main()
