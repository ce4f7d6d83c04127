"""Exchange-latency probe (run under torchrun with N ranks): 2000 mini-batch SGD steps on a tiny row set, so that the
step time is launch + reduce (+ NVLink exchange) overhead.  Compare N = 1 and N = 2..8; RCMARL_PEER_COMM=0 -> NCCL."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "resilient-consensus-based-marl_b200"))
sys.path.insert(0, ROOT)
import bench                                    # noqa: E402
from rcmarl import dist_util                    # noqa: E402
from rcmarl.trainer import Trainer              # noqa: E402

rank, world, local = dist_util.init_from_env()
torch.cuda.set_device(local)
w, desired, labels = bench.load_pretrained()
n_envs = int(sys.argv[1]) if len(sys.argv) > 1 else 64
tr = Trainer(labels=labels, in_nodes=bench.IN_NODES5, weights=w, desired=desired, n_envs=n_envs, rank=rank, world=world, H=1,
             gamma=0.9, fast_lr=0.01, slow_lr=0.002, max_ep_len=20, n_ep_fixed=50, n_epochs=1, buffer_size=2000, seed=1)
tr.rollout_block()
for _ in range(2):
    tr.profile = {}
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    tr.update_round()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    mb = sum(a.elapsed_time(b) for a, b in tr.profile.get("minibatch_sgd", []))
    steps = 10 * ((tr.t_filled + 31) // 32) if tr.t_filled else 0
if rank == 0:
    T = 1000
    steps = 10 * ((T + 31) // 32)
    print(f"world={world} peer_comm={tr.comm is not None} n_envs/rank={n_envs}: mini-batch chain {mb:.2f} ms for {steps} steps "
          f"= {1e3 * mb / steps:.1f} us/step; whole round {dt * 1e3:.1f} ms", flush=True)
if tr.comm is not None:
    tr.comm.close()
if world > 1:
    torch.distributed.destroy_process_group()
