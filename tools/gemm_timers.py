"""Where the tile GEMM pipeline of k_lauum waits: cycle counters of a -DBGP_TIMING build (make -C branchedgp_b200/csrc timing).

usage: BGP_LIB_NAME=libbgp_b200_timing.so python tools/gemm_timers.py [N] [items]"""
import ctypes
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from branchedgp_b200 import _lib  # noqa: E402
from common import make_problem  # noqa: E402


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
    items = int(sys.argv[2]) if len(sys.argv) > 2 else 148
    pr = make_problem(N, D=1, n_genes=1, bs=(0.4,), seed=3)
    eng = _lib.Engine(0)
    rep = lambda a: np.repeat(a, items, axis=0)
    args = (pr["t"], pr["Y"], rep(pr["item_gene"]), rep(pr["b"]), pr["prior"], rep(pr["hyp"]), rep(pr["logPhi"]))
    eng.dense_elbo_grad(*args)
    out = (ctypes.c_ulonglong * 8)()
    eng._lib.bgp_debug_gemm_timers(eng._h, out, 1)
    eng.dense_elbo_grad(*args)
    eng._lib.bgp_debug_gemm_timers(eng._h, out, 1)
    v = [int(x) for x in out]
    ctas = items
    print(json.dumps(dict(lib=os.environ.get("BGP_LIB_NAME", "libbgp_b200.so"), N=N, items=items,
                          calls_per_cta=v[4] / ctas, ksteps_per_cta=v[5] / ctas,
                          cycles_in_gemm_calls_per_cta=v[3] / ctas,
                          cycles_per_kstep=round(v[3] / max(v[5], 1), 1), ideal_cycles_per_kstep=8192,
                          frac_wait_first_stages=round(v[0] / max(v[3], 1), 4),
                          frac_wait_later_stages=round(v[1] / max(v[3], 1), 4),
                          frac_producer_wait_empty=round(v[2] / max(v[3], 1), 4),
                          chol128_us_per_block=round(v[6] / 1.965e3 / (2 * ctas * ((3 * N + 127) // 128)), 2),
                          trinv128_us_per_block=round(v[7] / 1.965e3 / (2 * ctas * ((3 * N + 127) // 128)), 2))))


if __name__ == "__main__":
    main()
