"""World-size-2 (gloo, CPU) coverage of the data-parallel host logic: the env shards tile the batch,
and `all-reduce of UNSCALED per-shard sums, then divide by the GLOBAL batch` reproduces the single-process full-batch
SGD step.  Per-shard sums come from the oracle here (no GPU in this test); on the GPU the same helper
(rcmarl.dist_util.allreduce_sums) reduces the sums written by rcmarl_grad."""
import os
import socket
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.multiprocessing as mp            # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    for p in (ROOT, os.path.join(ROOT, "resilient-consensus-based-marl_b200"), os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    from rcmarl import dist_util
    from oracle import rpbcac_oracle as O
    from rcmarl import nets
    r, w, _ = dist_util.init_from_env("gloo")
    assert (r, w) == (rank, world)
    rs = np.random.RandomState(0)                       # identical data on both ranks; each takes its env shard
    N, T = 6, 10
    x = rs.randn(T, N, 15).astype(np.float64)
    y = rs.randn(T, N, 1).astype(np.float64)
    wts = O.cast_weights(nets.glorot_uniform(15, 1, rs), np.float64)
    first, cnt = dist_util.shard_envs(N, rank, world)
    xs, ys = x[:, first:first + cnt].reshape(-1, 15), y[:, first:first + cnt].reshape(-1, 1)
    outp, cch = O.mlp_forward(wts, xs, cache=True)
    e = outp - ys
    g = O.mlp_backward(wts, cch, e)
    sums = torch.tensor(np.concatenate([a.reshape(-1) for a in g] + [[(e * e).sum()]]))
    dist_util.allreduce_sums(sums, world)
    B = T * N
    lr = 0.01
    new = nets.pack(wts).astype(np.float64) - lr * 2.0 / B * sums[:-1].numpy()
    want, loss = O.mse_step(wts, x.reshape(-1, 15), y.reshape(-1, 1), lr)
    np.testing.assert_allclose(new, np.concatenate([a.reshape(-1) for a in want]), rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(float(sums[-1]) / B, loss, rtol=1e-12)
    if rank == 0:
        open(out, "w").write("ok")
    import torch.distributed as dist
    dist.destroy_process_group()


def test_sharded_sums_allreduce_equals_full_batch(tmp_path):
    out = str(tmp_path / "ok.txt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    assert open(out).read() == "ok"


def test_env_shards_tile_the_batch():
    sys.path.insert(0, os.path.join(ROOT, "resilient-consensus-based-marl_b200"))
    from rcmarl.dist_util import shard_envs
    for n, world in ((4096, 8), (10, 4), (7, 2), (1, 1), (3, 8)):
        got = [shard_envs(n, r, world) for r in range(world)]
        assert got[0][0] == 0 and sum(c for _, c in got) == n
        for (f0, c0), (f1, _c1) in zip(got, got[1:]):
            assert f1 == f0 + c0
