"""Stage timeline of one producer thread of grad_kernel_ws (debug build: make -C .../csrc variant_tl):
    RCMARL_LIB=.../librcmarl_tl.so python tools/ws_timeline.py
Prints the mean cycles between the stage boundaries of tiles 8..63 of CTA 0 / thread 0."""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "resilient-consensus-based-marl_b200"))
from rcmarl import ops, nets, _lib as L   # noqa: E402

NA, rows_n, n_jobs = 5, 4096 * 1000, 8
dev = torch.device("cuda:0")
g = torch.Generator(device="cuda"); g.manual_seed(0)
sa = torch.randn(rows_n, 3 * NA, device=dev, generator=g)
ns = torch.randn(rows_n, 2 * NA, device=dev, generator=g)
r = torch.randn(rows_n, NA, device=dev, generator=g)
tgt = torch.randn(rows_n, device=dev, generator=g)
rs = np.random.RandomState(0)
sums = torch.zeros(n_jobs, 762, device=dev)
jobs = []
keep = []
for j in range(n_jobs):
    kind, din = (L.IN_SA, 15) if j % 2 == 0 else (L.IN_S, 10)
    w = torch.as_tensor(nets.pack(nets.glorot_uniform(din, 1, rs))).to(dev)
    keep.append(w)
    jobs.append(ops.grad_job(w, tgt, sums[j][:L.param_count(din, 1) + 1], kind))
rows = ops.make_rows(sa, ns, r, NA)
for _ in range(3):
    ops.grad(rows, jobs, L.LOSS_MSE)
torch.cuda.synchronize()
lib = L.lib()
buf = (C.c_longlong * (64 * 16))()
lib.rcmarl_debug_timeline.restype = C.c_int
st = lib.rcmarl_debug_timeline(buf, 64 * 16)
t = np.array(buf[:], np.int64).reshape(64, 16)
names = ["h1 split, STTM", "group barrier B", "L2 MMAs (in their shadow: claim buffer, x / h1 -> buffer, next loads)", "LDTM",
         "lrelu, head, delta2, split (+ next x split), STTM", "group barrier C", "L3 + next L1 MMAs (delta2 -> buffer)",
         "LDTM, delta1 -> buffer, arrive", "LDTM next h1"]
t = t[:, :10]
d = np.diff(t[8:60], axis=1)
per_tile = np.diff(t[8:60, 0])
print("status", st, "tile period of this group (cycles): mean", per_tile.mean(), "min", per_tile.min(), "max", per_tile.max())
for k, n in enumerate(names):
    print(f"{n:42s} mean {d[:, k].mean():8.0f}  min {d[:, k].min():6d}  max {d[:, k].max():6d}")
print("sum of stages", d.mean(0).sum())
