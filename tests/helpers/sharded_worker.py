"""Worker of tests/test_sharded_gpu.py, launched as `python -m torch.distributed.run --nproc-per-node 2 ... sharded_worker.py OUT`.

Every rank evaluates its contiguous block of (gene, branching point) items through sharding.dense_bound_sharded and the per-item
results are gathered (NCCL when every rank has its own GPU, gloo when the ranks share the box's single GPU); rank 0 then repeats
the whole batch in ONE call on its own GPU and writes both tables for the test to compare bit for bit."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    out_path = sys.argv[1]
    N = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    import torch
    import torch.distributed as dist
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    ngpu = torch.cuda.device_count()
    own_gpu = ngpu >= world
    dev = rank if own_gpu else 0
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl" if own_gpu else "gloo", rank=rank, world_size=world)
    from branchedgp_b200 import _lib, sharding
    from common import make_problem
    eng = _lib.Engine(dev)
    pr = make_problem(N, D=1, n_genes=3, bs=(0.3, 0.5, 0.8), seed=7)   # 9 items: blocks of 5 and 4
    res = sharding.dense_bound_sharded(eng, pr["t"], pr["Y"], pr["item_gene"], pr["b"], pr["prior"], pr["hyp"],
                                       lambda lo, hi: pr["logPhi"][lo:hi], device=torch.device("cuda", dev) if own_gpu else None,
                                       want_grad_logphi=True)
    lo, hi = sharding.shard_range(pr["count"], world, rank)
    # the same gather through the C ABI's own NCCL communicator (bgp_comm_init / bgp_gather_results); needs one GPU per rank
    capi = np.zeros((0, 6))
    if own_gpu:
        ids = [eng.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        eng.comm_init(world, rank, ids[0])
        mine = np.hstack([res["elbo"][lo:hi, None], res["grad_hyp"][lo:hi], res["status"][lo:hi, None].astype(np.float64)])
        blocks = eng.gather_results(mine, -(-pr["count"] // world))
        capi = np.vstack([blocks[r, :sharding.shard_range(pr["count"], world, r)[1] - sharding.shard_range(pr["count"], world, r)[0]]
                          for r in range(world)])
    if rank == 0:
        one = eng.dense_elbo_grad(pr["t"], pr["Y"], pr["item_gene"], pr["b"], pr["prior"], pr["hyp"], pr["logPhi"], check=False)
        np.savez(out_path, elbo=res["elbo"], grad_hyp=res["grad_hyp"], status=res["status"], one_elbo=one["elbo"],
                 one_grad_hyp=one["grad_hyp"], one_status=one["status"], local_grad=res["grad_logPhi"],
                 one_grad_block=one["grad_logPhi"][lo:hi], backend=np.array(dist.get_backend()), world=world, capi=capi)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
