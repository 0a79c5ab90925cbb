"""Sharding of (gene, branching point) work items over the GPUs of one box.

Work items are independent, so every rank evaluates a contiguous block with no data-path collective; the only exchange is
one all_gather of the per-item results (ELBO, hyper-gradients, status) -- NCCL over NVLink on the GPU box, gloo in the
CPU tests.  One process per GPU (torchrun); `torch.distributed` is used for the plumbing only.
"""
import numpy as np


def shard_range(count, world, rank):
    """Contiguous block [lo, hi) of `count` items owned by `rank`; block sizes differ by at most one."""
    base, extra = divmod(int(count), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_rows(local, count, group=None, device=None):
    """all_gather of row blocks: `local` is this rank's [hi-lo, k] float64 block of a [count, k] array -> the full array on
    every rank.  Blocks are padded to the largest block so that one fixed-size collective suffices."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    local = np.ascontiguousarray(local, dtype=np.float64)
    k = local.shape[1]
    biggest = -(-count // world)
    buf = torch.zeros((biggest, k), dtype=torch.float64, device=device)
    if local.shape[0]:
        buf[:local.shape[0]] = torch.as_tensor(local, device=device)
    out = torch.empty((world * biggest, k), dtype=torch.float64, device=device)
    dist.all_gather_into_tensor(out, buf, group=group)
    out = out.cpu().numpy().reshape(world, biggest, k)
    full = np.empty((count, k))
    for r in range(world):
        lo, hi = shard_range(count, world, r)
        full[lo:hi] = out[r, :hi - lo]
    return full


def evaluate_items_sharded(evaluate_local, count, group=None, device=None):
    """Run `evaluate_local(lo, hi) -> [hi-lo, k] array` on this rank's block and gather the [count, k] result everywhere."""
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    lo, hi = shard_range(count, world, rank)
    local = evaluate_local(lo, hi) if hi > lo else np.zeros((0, 1))
    k = local.shape[1] if hi > lo else None
    if k is None:   # a rank without work still has to know the row width
        import torch
        kk = torch.zeros(1, dtype=torch.int64, device=device)
        dist.all_reduce(kk, op=dist.ReduceOp.MAX, group=group)
        local = np.zeros((0, int(kk.item())))
    else:
        import torch
        kk = torch.tensor([k], dtype=torch.int64, device=device)
        dist.all_reduce(kk, op=dist.ReduceOp.MAX, group=group)
    return gather_rows(local, count, group=group, device=device)


def dense_bound_sharded(engine, t, Y, item_gene, b, prior, hyp, logPhi_of, group=None, device=None, want_grad_logphi=False, **kw):
    """Dense bound (+ hyper-gradient) of `count` items sharded over the ranks of `group`.

    `logPhi_of(lo, hi)` returns the [hi-lo, N, 3N] logits of this rank's items (they are never exchanged: 24 MB per item at
    N = 1000).  Returns dict(elbo[count], grad_hyp[count,4], status[count]) on every rank, plus this rank's grad_logPhi block
    when want_grad_logphi."""
    item_gene = np.asarray(item_gene)
    count = item_gene.size
    keep = {}

    def local(lo, hi):
        out = engine.dense_elbo_grad(t, Y, item_gene[lo:hi], np.asarray(b)[lo:hi], prior, np.asarray(hyp)[lo:hi], logPhi_of(lo, hi),
                                     check=False, **kw)
        if want_grad_logphi:
            keep["grad_logPhi"] = out["grad_logPhi"]
        return np.hstack([out["elbo"][:, None], out["grad_hyp"], out["status"][:, None].astype(np.float64)])

    full = evaluate_items_sharded(local, count, group=group, device=device)
    res = dict(elbo=full[:, 0], grad_hyp=full[:, 1:5], status=full[:, 5].astype(np.int32))
    res.update(keep)
    return res


def fit_genes_sharded(fit_one, n_genes, n_grid, group=None, device=None):
    """Genome-wide sweep: `fit_one(g) -> loglik[n_grid]` (e.g. FitModel(...)['loglik'] for gene g) is run for this rank's
    genes; returns the gathered [n_genes, n_grid] log-likelihood table on every rank."""
    def local(lo, hi):
        return np.stack([np.asarray(fit_one(g), dtype=np.float64).reshape(n_grid) for g in range(lo, hi)])
    return evaluate_items_sharded(local, n_genes, group=group, device=device)


def fit_model_batch_sharded(bConsider, GPt, GPY, globalBranching, group=None, device=None, **fit_kwargs):
    """Genome-wide sweep with the device-resident batched fit: every rank runs ``FitBranchingModel.FitModelBatch`` on its
    contiguous block of genes (columns of ``GPY`` [N, G]) and the per-gene results that FitModel reports as scalars / small
    vectors are gathered on every rank: ``loglik`` [G, B], ``hyperparameters`` [G, 3] (kerlen, kervar, likvar at the best
    branching point; NaN for a failed gene) and ``Bmode`` [G].  The rank-local list of full result dictionaries (Phi,
    prediction, posteriorB) is returned under ``local`` with its gene range ``local_range``."""
    import torch.distributed as dist

    from . import FitBranchingModel
    G = GPY.shape[1]
    B = len(bConsider)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    lo, hi = shard_range(G, world, rank)
    keep = {}

    def local(lo_, hi_):
        res = FitBranchingModel.FitModelBatch(bConsider, GPt, GPY[:, lo_:hi_], globalBranching, **fit_kwargs)
        keep["local"] = res
        rows = np.full((hi_ - lo_, B + 4), np.nan)
        for k, r in enumerate(res):
            rows[k, :B] = r["loglik"]
            if isinstance(r["hyperparameters"], dict):
                h = r["hyperparameters"]
                rows[k, B:B + 3] = [float(h["kerlen"]), float(h["kervar"]), float(h["likvar"])]
                rows[k, B + 3] = r["posteriorB"]["Bmode"]
        return rows

    full = evaluate_items_sharded(local, G, group=group, device=device)
    return dict(loglik=full[:, :B], hyperparameters=full[:, B:B + 3], Bmode=full[:, B + 3], local=keep.get("local", []),
                local_range=(lo, hi))
