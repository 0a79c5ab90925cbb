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
         11: "PANEL x inv gemm + store + signal", 12: "UPD tasks", 13: "UPD wait deps", 14: "UPD gemm + RMW + signal"}


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
    for grp, cnt in ((range(1, 8), 0), (range(9, 12), 8), (range(13, 15), 12)):
        n = max(int(out[cnt]), 1)
        print("%s: %d (two Cholesky kernels)" % (NAMES[cnt], int(out[cnt])))
        for k in grp:
            print("   %-40s %9.2f us per task" % (NAMES[k], float(out[k]) / n / mhz))


if __name__ == "__main__":
    main()
