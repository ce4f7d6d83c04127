"""C5 microbench driver: 64 x 1M clip-mean, H in {0,1,2,4}, L2 flushed between launches.
   python tools/prof_clip_mean.py            (CUDA-event timings)"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "resilient-consensus-based-marl_b200"))
from rcmarl import ops   # noqa: E402

g = torch.Generator(device="cuda"); g.manual_seed(0)
X = torch.randn(64, 1 << 20, device="cuda", generator=g)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
out = torch.empty(1 << 20, device="cuda")
for H in [int(a) for a in sys.argv[1:]] or [0, 1, 2, 4, 7]:
    for _ in range(3):
        ops.clip_mean(X, H, out)
    ts = []
    for _ in range(10):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); ops.clip_mean(X, H, out); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    t = float(np.median(ts))
    print(f"H={H}: {t * 1e3:.1f} us  {4.0 * (1 << 20) * 65 / (t * 1e-3) / 1e9:.0f} GB/s algorithmic")
