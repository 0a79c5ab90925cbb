"""Seeded synthetic problems shared by the tests (shapes follow SURVEY.md section 8d)."""
import numpy as np

from oracle import host_ref


def make_problem(N, D=1, n_genes=1, bs=(0.4,), seed=0, jitter_logphi=0.1, mbgp=False):
    """Returns a dict with t, labels, prior, phi0, X, Y[n_genes,N,D] and per-item arrays for genes x bs."""
    rng = np.random.default_rng(seed)
    t = np.sort(rng.uniform(0.0, 1.0, N))
    labels = np.where(t < 0.3, 1, rng.integers(2, 4, N))
    phi0, prior = host_ref.initial_conditions_and_prior(labels, 0.8)
    X, idx = host_ref.function_index_list(t)
    Y = np.empty((n_genes, N, D))
    for g in range(n_genes):
        for d in range(D):
            bg, ag = rng.uniform(0.1, 0.9), rng.uniform(0.5, 3.0)
            sign = np.where(labels == 3, -1.0, 1.0)
            y = sign * ag * np.maximum(t - bg, 0.0) + 0.3 * rng.standard_normal(N)
            Y[g, :, d] = y - y.mean()
    items = [(g, b) for g in range(n_genes) for b in bs]
    count = len(items)
    item_gene = np.array([g for g, _ in items], dtype=np.int32)
    b = np.array([bb for _, bb in items])
    hyp = np.stack([rng.uniform(0.5, 3.0, count), rng.uniform(1.0, 6.0, count), rng.uniform(0.05, 1.0, count)], axis=1)
    logPhi = np.empty((count, N, 3 * N))
    for k, (g, bb) in enumerate(items):
        if mbgp:
            base, _ = host_ref.init_logphi_mbgp(phi0, t)
        else:
            base = host_ref.init_logphi_bgp(phi0, t, bb)
        logPhi[k] = base + jitter_logphi * rng.standard_normal((N, 3 * N))
    prior3 = np.hstack([np.zeros((N, 1)), prior])
    return dict(N=N, D=D, t=t, labels=labels, prior=prior, prior3=prior3, phi0=phi0, X=X, indices=idx, Y=Y,
                item_gene=item_gene, b=b, hyp=hyp, logPhi=logPhi, count=count)


def relerr(a, b):
    a = np.asarray(a, dtype=float)
    b = np.asarray(b, dtype=float)
    den = max(np.abs(b).max(), 1e-300)
    return float(np.abs(a - b).max() / den)
