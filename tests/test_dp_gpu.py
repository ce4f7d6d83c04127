"""Two-GPU data-parallel equivalence: 2 ranks x N/2 environments == 1 GPU x N environments, for both
exchange paths (in-kernel NVLink peer-memory all-reduce, and NCCL).  Needs >= 2 GPUs (gpurun --gpus 2)."""
import os
import socket
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu
import torch.multiprocessing as mp            # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
IN_NODES = [[0, 1, 2, 3], [1, 2, 3, 4], [2, 3, 4, 0], [3, 4, 0, 1], [4, 0, 1, 2]]


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, peer_comm, out):
    for p in (ROOT, os.path.join(ROOT, "resilient-consensus-based-marl_b200"), os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port), RCMARL_PEER_COMM="1" if peer_comm else "0")
    import torch.distributed as dist
    from golden_util import pretrained
    from rcmarl import dist_util
    from rcmarl.trainer import Trainer
    torch.cuda.set_device(rank)
    dist_util.init_from_env("nccl")
    w, desired, labels = pretrained()
    N = 64
    kw = dict(labels=labels, in_nodes=IN_NODES, weights=w, desired=desired, gamma=0.9, H=1, fast_lr=0.01, slow_lr=0.002,
              max_ep_len=8, n_ep_fixed=9, n_epochs=2, buffer_size=100, seed=5)
    single = None
    if rank == 0:                                    # the single-GPU reference run, before any exchange context exists
        tr1 = Trainer(n_envs=N, **kw)
        for _ in range(2):
            tr1.rollout_block(); tr1.update_round()
        single = [tr1.get_weights(i) for i in range(5)]
    dist.barrier()
    tr = Trainer(n_envs=N // world, rank=rank, world=world, **kw)
    assert (tr.comm is not None) == bool(peer_comm)
    for _ in range(2):
        tr.rollout_block(); tr.update_round()
    if tr.comm is not None:
        tr.comm.check()
    mine = torch.cat([tr.actor.flatten(), tr.critic.flatten(), tr.tr.flatten(), tr.critic_local.flatten()])
    both = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(both, mine)
    if rank == 0:
        assert torch.equal(both[0], both[1]), "replicated parameters diverged between ranks"
        got = [tr.get_weights(i) for i in range(5)]
        for i in range(5):
            for n in range(len(single[i])):
                for k in range(6):
                    np.testing.assert_allclose(got[i][n][k], single[i][n][k], rtol=2e-3, atol=2e-5)
        open(out, "w").write("ok")
    if tr.comm is not None:
        tr.comm.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("peer_comm", [1, 0])
def test_two_ranks_equal_one_gpu(peer_comm, tmp_path):
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    out = str(tmp_path / "ok.txt")
    mp.spawn(_worker, args=(2, _free_port(), peer_comm, out), nprocs=2, join=True)
    assert open(out).read() == "ok"
