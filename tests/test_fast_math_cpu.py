"""CPU: the constants of csrc/bgp_math.cuh (tailored fp64 exp / log of the assignment passes) are parsed from the shipped header
and checked against numpy / math references; the kernels' arithmetic is emulated in float64 (fma steps via longdouble where the
exactness of the range reduction matters).  Guards the tables against typos and the documented 2.3e-16 accuracy."""
import math
import os
import re

import numpy as np

HDR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "branchedgp_b200", "csrc", "bgp_math.cuh")
SRC = open(HDR).read()


def _block(name):
    m = re.search(re.escape(name) + r"\s*\[\d+\]\s*=\s*\{(.*?)\};", SRC, re.S)
    assert m, name
    body = re.sub(r"\s*/\s*", "/", re.sub(r"//.*", "", m.group(1)))
    vals = []
    for tok in re.split(r"[,{}\s]+", body):
        if not tok:
            continue
        if "/" in tok:
            a, b = tok.split("/")
            vals.append(float(a) / float(b))
        else:
            vals.append(float(tok))
    return np.array(vals)


def test_exp_table_and_reduction_constants():
    tab = _block("FAST_EXP_TAB")
    assert tab.size == 32
    want = np.array([2.0 ** (j / 32.0) for j in range(32)])
    assert np.max(np.abs(tab - want) / want) <= 2.3e-16
    c = _block("FAST_EXPT_C")
    assert abs(c[0] - 32.0 / math.log(2.0)) <= 1e-14
    hi, lo = -c[1], -c[2]
    assert abs((hi + lo) - math.log(2.0) / 32.0) <= 1e-18
    # the high part must carry few enough bits that n * hi is exact for every n the reduction can produce (|n| < 2^16)
    mant = int(np.frexp(hi)[0] * 2 ** 53)
    assert mant % (1 << 16) == 0
    assert c[7] == 1.5 * 2.0 ** 52
    np.testing.assert_allclose(c[3:7], [1 / 720, 1 / 120, 1 / 24, 1 / 6], rtol=1e-15)


def _fast_exp(x):
    c = _block("FAST_EXPT_C")
    tab = _block("FAST_EXP_TAB")
    x = np.asarray(x, dtype=np.float64)
    n = np.rint(x * c[0])
    r = np.longdouble(x) + np.longdouble(n) * np.longdouble(c[1])
    r = (r + np.longdouble(n) * np.longdouble(c[2])).astype(np.float64)
    h = np.full_like(r, c[3])
    for k in (c[4], c[5], c[6], 0.5, 1.0):
        h = h * r + k
    ni = n.astype(np.int64)
    T = tab[ni & 31]
    y = (np.longdouble(T) * np.longdouble(r * h) + np.longdouble(T)).astype(np.float64)
    return np.where(x < -708.0, 0.0, np.ldexp(y, (ni >> 5).astype(int)))


def test_fast_exp_accuracy_on_softmax_arguments():
    rng = np.random.default_rng(0)
    x = np.concatenate([-np.abs(rng.standard_normal(200000)) * 30.0, -rng.uniform(0.0, 708.0, 200000), [0.0, -1e-300, -707.9, -709.0, -1e308]])
    with np.errstate(all="ignore"):      # the -1e308 probe overflows the emulated range reduction; the kernel discards it
        got, ref = _fast_exp(x), np.exp(x)
    ok = x >= -708.0
    assert np.max(np.abs(got[ok] - ref[ok]) / ref[ok]) <= 4.5e-16      # numpy's exp itself is good to ~1 ulp
    assert np.all(got[~ok] == 0.0)


def test_log_table_is_consistent():
    tab = _block("FAST_LOG_TAB").reshape(128, 2)
    ci = 1.0 + (np.arange(128) + 0.5) / 128.0
    assert np.max(np.abs(tab[:, 0] - 1.0 / ci) * ci) <= 1.2e-16          # r_i = fl(1 / c_i)
    assert np.max(np.abs(tab[:, 1] + np.log(tab[:, 0]))) <= 1.2e-16      # l_i = -log(r_i) of the ROUNDED reciprocal
    c = _block("FAST_LOG_C")
    np.testing.assert_allclose(c[:6], [-0.125, 1 / 7, -1 / 6, 0.2, -0.25, 1 / 3], rtol=1e-15)
    assert abs((c[6] + c[7]) - math.log(2.0)) <= 1e-17


def test_fast_log_accuracy_on_squashed_probabilities():
    tab = _block("FAST_LOG_TAB").reshape(128, 2)
    c = _block("FAST_LOG_C")
    rng = np.random.default_rng(1)
    y = np.concatenate([np.exp(-rng.uniform(0.0, 14.0, 300000)), [1e-6, 1.0 - 2e-6 + 1e-6, 1.0, 0.5]])
    m, e = np.frexp(y)
    m, e = m * 2.0, e - 1                    # m in [1, 2)
    idx = np.minimum(((m - 1.0) * 128.0).astype(int), 127)
    r, l = tab[idx, 0], tab[idx, 1]
    u = (np.longdouble(m) * np.longdouble(r) - 1).astype(np.float64)      # one exact fma in the kernel
    p = np.full_like(u, c[0])
    for k in (c[1], c[2], c[3], c[4], c[5], -0.5, 1.0):
        p = p * u + k
    got = e * c[6] + (l + (u * p + e * c[7]))
    assert np.max(np.abs(got - np.log(y))) <= 4.5e-16                      # absolute accuracy is what the KL sums need
