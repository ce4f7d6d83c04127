"""Stand-alone driver for profiling rcmarl_grad / rcmarl_team under ncu (one GPU):
   ncu --set full --clock-control none --import-source on -k regex:grad_kernel -s 2 -c 1 -o gpurun_out/prof python tools/prof_grad.py
Prints CUDA-event timings when run without a profiler."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "resilient-consensus-based-marl_b200"))
from rcmarl import ops, nets, _lib as L   # noqa: E402

NA = 5
rows_n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096 * 1000
n_jobs = int(sys.argv[2]) if len(sys.argv) > 2 else 8
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
dev = torch.device("cuda:0")
g = torch.Generator(device="cuda"); g.manual_seed(0)
sa = torch.randn(rows_n, 3 * NA, device=dev, generator=g)
ns = torch.randn(rows_n, 2 * NA, device=dev, generator=g)
r = torch.randn(rows_n, NA, device=dev, generator=g)
tgt = torch.randn(rows_n, device=dev, generator=g)
rs = np.random.RandomState(0)
PC, PT = L.param_count(10, 1), L.param_count(15, 1)
wc = [torch.as_tensor(nets.pack(nets.glorot_uniform(10, 1, rs))).to(dev) for _ in range(n_jobs)]
wt = [torch.as_tensor(nets.pack(nets.glorot_uniform(15, 1, rs))).to(dev) for _ in range(n_jobs)]
sums = torch.zeros(n_jobs, PT + 1, device=dev)
rows = ops.make_rows(sa, ns, r, NA)
jobs = []
for j in range(n_jobs):
    if j % 2 == 0:
        jobs.append(ops.grad_job(wt[j], tgt, sums[j], L.IN_SA))
    else:
        jobs.append(ops.grad_job(wc[j], tgt, sums[j][:PC + 1], L.IN_S))
jobs = (L.GradJob * n_jobs)(*jobs)
ts = []
for i in range(reps):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); ops.grad(rows, jobs, L.LOSS_MSE); b.record()
    torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
mac = (n_jobs // 2) * (1860 + 1660) + (n_jobs % 2) * 1860
print(f"grad: rows={rows_n} jobs={n_jobs} ms={np.median(ts):.3f} (incl. reduce)  "
      f"{2 * mac * rows_n / (np.median(ts) * 1e-3) / 1e12:.2f} TFLOP/s fp32")
