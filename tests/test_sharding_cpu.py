"""CPU suite: the multi-rank path (block partition + result gather) with world_size = 2 on gloo."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

from branchedgp_b200 import sharding


def test_shard_range_partitions_exactly():
    for count in (0, 1, 7, 148, 10240):
        for world in (1, 2, 3, 8):
            blocks = [sharding.shard_range(count, world, r) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == count
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in blocks]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, count, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    calls = []

    def fake_eval(lo, hi):   # stands in for the CUDA evaluation of items lo..hi: row i = f(i)
        calls.append((lo, hi))
        i = np.arange(lo, hi, dtype=np.float64)
        return np.stack([-100.0 - i, np.sin(i), np.cos(i), i * i, 0.5 * i, np.zeros_like(i)], axis=1)

    full = sharding.evaluate_items_sharded(fake_eval, count)
    table = sharding.fit_genes_sharded(lambda g: np.array([g, g + 0.5, -g]), count, 3)

    # the batched fit, sharded by gene: a stand-in for FitModelBatch (which needs the GPU) that encodes the gene's data
    from branchedgp_b200 import FitBranchingModel
    grid = [0.2, 0.6, 1.1]

    def fake_batch(bConsider, GPt, GPY, labels, **kw):
        assert kw == {"maxiter": 7}
        res = []
        for g in range(GPY.shape[1]):
            key = float(GPY[0, g])
            if key == 0.0:    # gene 0 "fails" like FitModel's exception branch
                res.append({"loglik": np.array([np.nan, 0.0, 0.0]), "hyperparameters": np.nan, "posteriorB": np.nan})
            else:
                res.append({"loglik": np.array([key, 2 * key, -key]), "posteriorB": {"Bmode": bConsider[1]},
                            "hyperparameters": {"kerlen": np.array(key + 1), "kervar": np.array(key + 2), "likvar": np.array(key + 3)}})
        return res
    FitBranchingModel.FitModelBatch = fake_batch
    GPY = np.tile(np.arange(count, dtype=np.float64), (4, 1))
    sh = sharding.fit_model_batch_sharded(grid, np.linspace(0, 1, 4), GPY, np.ones(4), maxiter=7)
    q.put((rank, calls, full, table, sh))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("count", [5, 1])
def test_gather_over_two_ranks_gloo(count):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, count, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    i = np.arange(count, dtype=np.float64)
    want = np.stack([-100.0 - i, np.sin(i), np.cos(i), i * i, 0.5 * i, np.zeros_like(i)], axis=1)
    for rank, calls, full, table, sh in got:
        lo, hi = sharding.shard_range(count, world, rank)
        assert calls == ([(lo, hi)] if hi > lo else [])          # every rank evaluated only its own block
        assert np.array_equal(full, want)                          # and every rank ends up with all results
        assert np.array_equal(table, np.stack([i, i + 0.5, -i], axis=1))
        assert sh["local_range"] == (lo, hi) and len(sh["local"]) == hi - lo
        assert np.isnan(sh["loglik"][0, 0]) and np.all(np.isnan(sh["hyperparameters"][0])) and np.isnan(sh["Bmode"][0])
        assert np.array_equal(sh["loglik"][1:], np.stack([i, 2 * i, -i], axis=1)[1:])
        assert np.array_equal(sh["hyperparameters"][1:], np.stack([i + 1, i + 2, i + 3], axis=1)[1:])
        assert np.all(sh["Bmode"][1:] == 0.6)
