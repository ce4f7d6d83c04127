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


def _workload(shape):
    """(labels, in_nodes, weights, desired, grid) of the two tested shapes: the C2 team (4 cooperative + 1 malicious) and
    a C3-shaped team (16 cooperative agents, 10x10 grid, in-neighbourhoods of 6, H = 2: 32 fit jobs x 45 blocks = 1 440
    exchange items per step, more than one resident wave -- the case the round-1 grid-wide rendezvous could not run)."""
    from golden_util import pretrained
    from rcmarl import nets
    if shape == "c2":
        w, desired, labels = pretrained()
        return labels, IN_NODES, w, desired, 5, 1
    rs = np.random.RandomState(2)
    NA = 16
    w = [[nets.glorot_uniform(32, 5, rs), nets.glorot_uniform(32, 1, rs), nets.glorot_uniform(48, 1, rs)] for _ in range(NA)]
    return (["Cooperative"] * NA, [[(i + k) % NA for k in range(6)] for i in range(NA)], w, rs.randint(0, 10, size=(NA, 2)),
            10, 2)


def _worker(rank, world, port, peer_comm, out, shape="c2"):
    for p in (ROOT, os.path.join(ROOT, "resilient-consensus-based-marl_b200"), os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port), RCMARL_PEER_COMM="1" if peer_comm else "0")
    import torch.distributed as dist
    from rcmarl import dist_util
    from rcmarl.trainer import Trainer
    torch.cuda.set_device(rank)
    dist_util.init_from_env("nccl")
    labels, in_nodes, w, desired, grid, H = _workload(shape)
    NA = len(labels)
    N = 64
    kw = dict(labels=labels, in_nodes=in_nodes, weights=w, desired=desired, nrow=grid, ncol=grid, gamma=0.9, H=H,
              fast_lr=0.01, slow_lr=0.002, max_ep_len=8, n_ep_fixed=9, n_epochs=2, buffer_size=100, seed=5)
    single = None
    if rank == 0:                                    # the single-GPU reference run, before any exchange context exists
        tr1 = Trainer(n_envs=N, **kw)
        for _ in range(2):
            tr1.rollout_block(); tr1.update_round()
        single = [tr1.get_weights(i) for i in range(NA)]
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
        got = [tr.get_weights(i) for i in range(NA)]
        for i in range(NA):
            for n in range(len(single[i])):
                for k in range(6):
                    np.testing.assert_allclose(got[i][n][k], single[i][n][k], rtol=2e-3, atol=2e-5)
        open(out, "w").write("ok")
    if tr.comm is not None:
        tr.comm.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("peer_comm,shape", [(1, "c2"), (0, "c2"), (1, "c3")])
def test_two_ranks_equal_one_gpu(peer_comm, shape, tmp_path):
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    out = str(tmp_path / "ok.txt")
    mp.spawn(_worker, args=(2, _free_port(), peer_comm, out, shape), nprocs=2, join=True)
    assert open(out).read() == "ok"
