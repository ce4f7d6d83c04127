"""One C2-shaped update round (n_epochs = 1) after three rollout blocks, for `ncu -k regex:<kernel> -c 1` captures of every
kernel of the hot path at its real shapes (4096 environments, 12.29 M buffer rows):
    ncu --set full --clock-control none --import-source on -k regex:team_kernel -c 1 -o gpurun_out/prof_team python tools/prof_round.py
Without a profiler it prints the CUDA-event time of the round."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "resilient-consensus-based-marl_b200"))
import bench                                           # noqa: E402
from rcmarl.trainer import Trainer                     # noqa: E402

cfg = bench.workload(sys.argv[1] if len(sys.argv) > 1 else "C2", int(sys.argv[2]) if len(sys.argv) > 2 else 0)
cfg["n_epochs"] = 1
tr = Trainer(seed=1, **cfg)
for _ in range(3):
    tr.rollout_block()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
tr.update_round()
b.record()
torch.cuda.synchronize()
print(f"one update round with n_epochs=1: {a.elapsed_time(b):.2f} ms, {tr.launches} launches")
