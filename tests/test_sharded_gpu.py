"""GPU: the multi-rank path end to end -- two ranks under torchrun, each evaluating its block of work items on the GPU, results
gathered over the process group (NCCL with one GPU per rank, gloo when both ranks share the box's only GPU) -- must return
bit-identical ELBOs / hyper-gradients to the single-GPU call on the same items (work items are independent and every item's
arithmetic is independent of how many items share a launch)."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("N,wide", [(40, "1"), (150, "0"), (150, "1")])
def test_two_rank_sharded_bound_equals_single_gpu_call(tmp_path, N, wide):
    out = str(tmp_path / "sharded.npz")
    env = dict(os.environ, BGP_DENSE_WIDE=wide)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(HERE, "helpers", "sharded_worker.py"), out, str(N)]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    z = np.load(out)
    assert int(z["world"]) == 2
    assert np.all(z["status"] == 0) and np.all(z["one_status"] == 0)
    assert np.array_equal(z["elbo"], z["one_elbo"])            # bit for bit
    assert np.array_equal(z["grad_hyp"], z["one_grad_hyp"])
    assert np.array_equal(z["local_grad"], z["one_grad_block"])
    if str(z["backend"]) == "nccl":   # one GPU per rank: the C ABI's own NCCL gather returned the same table
        want = np.hstack([z["one_elbo"][:, None], z["one_grad_hyp"], z["one_status"][:, None].astype(np.float64)])
        assert np.array_equal(z["capi"], want)
