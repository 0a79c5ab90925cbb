"""Generates and checks the constants of branchedgp_b200/csrc/bgp_math.cuh (fast_exp polynomial, fast_log table) against
60-digit mpmath references; prints the coefficients, the measured errors of a float64 emulation and writes the table rows
to /tmp/tab.txt.  usage: python tools/gen_fast_math.py"""
import mpmath as mp, numpy as np
mp.mp.dps = 60
# ---- exp polynomial on [-ln2/2, ln2/2]: Chebyshev interpolation degree 12
a = mp.log(2) / 2
deg = 12
nodes = [a * mp.cos(mp.pi * (2 * k + 1) / (2 * (deg + 1))) for k in range(deg + 1)]
A = mp.matrix(deg + 1, deg + 1)
bvec = mp.matrix(deg + 1, 1)
for i, x in enumerate(nodes):
    for j in range(deg + 1):
        A[i, j] = x ** j
    bvec[i] = mp.e ** x
coef = mp.lu_solve(A, bvec)
c = [float(coef[j]) for j in range(deg + 1)]
print("exp coeffs:", ["%.17g" % v for v in c])
# emulate in float64
def exp_emul(x):
    x = np.asarray(x, dtype=np.float64)
    L2E = 1.4426950408889634
    LN2_HI = 6.93147180369123816490e-01
    LN2_LO = 1.90821492927058770002e-10
    n = np.rint(x * L2E)
    r = (x - n * LN2_HI) - n * LN2_LO      # fma in the kernel; plain here (slightly worse)
    p = np.full_like(r, c[deg])
    for j in range(deg - 1, -1, -1):
        p = p * r + c[j]
    return np.ldexp(p, n.astype(int))
xs = -np.abs(np.random.default_rng(0).standard_normal(2000000)) * 30
xs = np.concatenate([xs, -np.linspace(0, 700, 100001)])
ref = np.array([float(mp.e ** mp.mpf(float(v))) for v in xs[:20000]])
got = exp_emul(xs[:20000])
print("exp max rel err (vs mpmath, 20000 pts):", np.max(np.abs(got - ref) / ref))
print("exp vs numpy max rel:", np.max(np.abs(exp_emul(xs) - np.exp(xs)) / np.exp(xs)))
# ---- log table: 128 entries, c_i = 1 + (i + 0.5)/128 over m in [1,2)
T = 128
rtab, ltab = [], []
for i in range(T):
    ci = mp.mpf(1) + (mp.mpf(i) + mp.mpf("0.5")) / T
    r = float(1 / ci)                 # rounded reciprocal
    rtab.append(r)
    ltab.append(float(-mp.log(mp.mpf(r))))   # -log of the ROUNDED reciprocal: log(m) = ltab + log1p(m*r - 1) exactly
lc = [1.0, -0.5, 1/3, -0.25, 0.2, -1/6, 1/7, -0.125]   # log1p(u)/u Taylor
def log_emul(y):
    y = np.asarray(y, dtype=np.float64)
    m, e = np.frexp(y)          # m in [0.5,1)
    m = m * 2; e = e - 1        # m in [1,2)
    idx = ((m - 1) * T).astype(int)
    r = np.array(rtab)[idx]; l = np.array(ltab)[idx]
    u = m * r - 1.0             # fma in the kernel (exact); here emulate with extended: use np.longdouble
    u = (np.longdouble(m) * np.longdouble(r) - 1).astype(np.float64)
    p = np.full_like(u, lc[7])
    for j in range(6, -1, -1):
        p = p * u + lc[j]
    LN2_HI = 6.93147180369123816490e-01
    LN2_LO = 1.90821492927058770002e-10
    return e * LN2_HI + (l + (u * p + e * LN2_LO))
ys = np.exp(-np.random.default_rng(1).uniform(0, 14, 200000))
ys = np.concatenate([ys, 1 - np.logspace(-12, -1, 2000), [1.0, 1e-6, 0.5, 0.999999]])
refl = np.array([float(mp.log(mp.mpf(float(v)))) for v in ys[:20000]])
gotl = log_emul(ys[:20000])
print("log max abs err:", np.max(np.abs(gotl - refl)), "max rel err where |log|>0.01:", np.max(np.abs(gotl - refl)[np.abs(refl) > 0.01] / np.abs(refl)[np.abs(refl) > 0.01]))
tail = log_emul(ys[-2004:]); reft = np.array([float(mp.log(mp.mpf(float(v)))) for v in ys[-2004:]])
print("log near 1: max abs err", np.max(np.abs(tail - reft)))
np.save("/tmp/rtab.npy", np.array(rtab)); np.save("/tmp/ltab.npy", np.array(ltab))
open("/tmp/tab.txt", "w").write(",\n".join("    {%.17g, %.17g}" % (r, l) for r, l in zip(rtab, ltab)))

# ---- table exp (the shipped fast_exp): x = (32 n + j) ln2/32 + r, exp x = 2^n T[j] (1 + r h(r)), T[j] = 2^(j/32), degree-5 h
import struct
Kt = 32
ln2_K = mp.log(2) / Kt
hi = float(ln2_K)
bits = struct.unpack('<Q', struct.pack('<d', hi))[0] & ~((1 << 17) - 1)      # 36 significant bits: n * hi is exact for |n| < 2^16
hi = struct.unpack('<d', struct.pack('<Q', bits))[0]
lo = float(ln2_K - mp.mpf(hi))
inv = float(Kt / mp.log(2))
ttab = [float(mp.mpf(2) ** (mp.mpf(j) / Kt)) for j in range(Kt)]
print("table exp constants: 32/ln2 = %.17g, ln2/32 hi = %.17g, lo = %.17g" % (inv, hi, lo))
print("T[j]:", ", ".join("%.17g" % v for v in ttab))
hc = [1.0, 0.5, 1 / 6, 1 / 24, 1 / 120, 1 / 720]
def exp_table_emul(x):
    x = np.asarray(x, dtype=np.float64)
    n = np.rint(x * inv)
    r = np.longdouble(x) - np.longdouble(n) * np.longdouble(hi)
    r = (r - np.longdouble(n) * np.longdouble(lo)).astype(np.float64)
    h = np.full_like(r, hc[5])
    for k in range(4, -1, -1):
        h = h * r + hc[k]
    ni = n.astype(np.int64)
    T = np.array(ttab)[ni & 31]
    y = (np.longdouble(T) * np.longdouble(r * h) + np.longdouble(T)).astype(np.float64)
    return np.ldexp(y, (ni >> 5).astype(int))
xt = np.concatenate([-np.abs(np.random.default_rng(0).standard_normal(20000)) * 30, -np.random.default_rng(1).uniform(0, 700, 20000)])
reft = np.array([float(mp.e ** mp.mpf(float(v))) for v in xt])
print("table exp max rel err (vs mpmath, 40000 pts):", np.max(np.abs(exp_table_emul(xt) - reft) / reft))
